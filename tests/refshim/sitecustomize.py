"""Loaded automatically when tests/refshim is on PYTHONPATH: stub out absent optional deps of the
reference front end so ``import fers_core`` works in this container (test shim only)."""
import importlib.util
import sys
import os

sys.path.insert(0, os.path.dirname(__file__))
from _permissive import install  # noqa: E402

_missing = [n for n in ("matplotlib", "pyvista", "sectionproperties", "ezdxf")
            if importlib.util.find_spec(n) is None]
_sub = {
    "matplotlib": ["matplotlib.pyplot", "matplotlib.patches", "matplotlib.path", "matplotlib.collections",
                   "matplotlib.colors", "matplotlib.cm", "matplotlib.lines", "matplotlib.figure",
                   "matplotlib.axes", "matplotlib.transforms", "matplotlib.tri",
                   "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"],
    "pyvista": [], "ezdxf": [],
    "sectionproperties": [],
}
for n in _missing:
    install([n] + _sub[n])
