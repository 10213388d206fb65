"""fers_core_b200 -- B200-native (sm_100a) implementation of the FERS analysis hot path.

Drop-in boundary: the ``fers_calculations`` module at the repo root
(``calculate_from_json`` / ``calculate_from_file``; reference call sites
fers_core/fers/fers.py:260,300).  See DESIGN.md.
"""
__version__ = "0.1.0"
