"""Permissive stub module factory for absent plotting/geometry dependencies (test shim only)."""
import sys
import types


class Permissive(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = Permissive(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return self


def install(names):
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = Permissive(name)
