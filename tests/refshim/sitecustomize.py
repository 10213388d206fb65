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


# Section.from_name builds named sections with `sectionproperties` (absent here).  The reference ships the properties of the
# same catalogue as data (fers_core/sections/steel_sections.generated.json, written by its own scripts/export_sections.py):
# when the package is missing, named sections are read from that file instead (test shim only).
if "sectionproperties" in _missing:
    import importlib.abc
    import importlib.machinery

    class _PatchSections(importlib.abc.MetaPathFinder):
        target = "fers_core.sections.steel_sections_en"

        def find_spec(self, name, path, target=None):
            if name != self.target:
                return None
            sys.meta_path.remove(self)
            try:
                spec = importlib.util.find_spec(name)
            finally:
                sys.meta_path.insert(0, self)
            if spec is None or spec.loader is None:
                return None
            inner = spec.loader

            class Loader(importlib.abc.Loader):
                def create_module(self, s):
                    return inner.create_module(s)

                def exec_module(self, module):
                    inner.exec_module(module)
                    import json
                    here = os.path.dirname(module.__file__)
                    cat = json.load(open(os.path.join(here, "steel_sections.generated.json")))["sections"]
                    keys = {k.replace(" ", "").upper(): k for k in cat}

                    def resolve_section(name, material):
                        from fers_core.members.section import Section
                        k = keys.get(str(name).replace(" ", "").upper())
                        if k is None:
                            raise ValueError(f"Section '{name}' not found in library.")
                        p = {f: v for f, v in cat[k].items() if f not in ("centroid_y", "centroid_z")}
                        return Section(name=str(name), material=material, **p)
                    module.resolve_section = resolve_section
            spec.loader = Loader()
            return spec
    sys.meta_path.insert(0, _PatchSections())
