"""Input fast path: wire.from_fers(calc) == wire.parse_document(calc.to_dict()) on every golden model,
checked in a subprocess that has the reference front end importable (build container only; tests/from_fers_check.py)."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import HAVE_REFERENCE, REFERENCE, ROOT

pytestmark = pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present")


def test_object_graph_flattens_to_the_same_arrays_as_the_document():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "refshim"), REFERENCE, ROOT])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "from_fers_check.py")], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    rep = json.loads(p.stdout.strip().splitlines()[-1])
    assert rep["checked"] >= 20
    # the fast path must at least beat the document route it replaces (to_dict + JSON round trip + parse)
    assert rep["from_fers_s"] < rep["to_dict_s"] + rep["json_round_trip_s"] + rep["parse_document_s"]
