"""json-backed stand-in for the absent ``ujson`` (test shim only)."""
from json import dumps, loads, dump, load  # noqa: F401
