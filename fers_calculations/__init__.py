"""Drop-in replacement for the reference's native engine module ``fers_calculations``.

The reference front end imports this name and calls exactly two functions
(fers_core/fers/fers.py:5, :260, :300; tests/nafems/nafems_harness.py:26, :38):

    calculate_from_json(input_json: str) -> str
    calculate_from_file(file_path: str)  -> str

Both take / return the wire document of ``fers_core/types/pydantic_models.py`` (``FERS``
2241-2258 in, same document with ``results`` = ``ResultsBundle`` 2171-2185 out).  Put this repo
on ``PYTHONPATH`` ahead of the PyPI wheel and ``FERS.run_analysis()`` runs on the B200.
Errors surface as exceptions, which the caller wraps in ``RuntimeError("Failed to run
calculation: ...")`` (fers.py:301-302).
"""
from fers_core_b200.engine import calculate_from_file, calculate_from_json, solve_document, solve_fers, solve_lazy  # noqa: F401

# beyond the reference's two functions: dict in / dict out without the JSON round trip (solve_document) and a lazy
# ResultsBundle view over the result arrays for large models (solve_lazy); solve_fers(calc) takes the reference's FERS object
# itself (input fast path: no to_dict / pydantic / JSON round trip)
__all__ = ["calculate_from_json", "calculate_from_file", "solve_document", "solve_lazy", "solve_fers"]
__version__ = "0.2.43+b200"
