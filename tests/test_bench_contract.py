"""bench.py contract (no GPU): the reference arm prints exactly one JSON line on stdout with the keys the driver
reads; the B200 arm fails loudly without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


SMALL = ["--nx", "4", "--nz", "4", "--ny", "3", "--combos", "8"]


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", *SMALL])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "solved_dof_x_load_combinations_per_second"
    assert d["unit"] == "DOF*combos/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # same workload as the GPU arm would report for these arguments (the driver compares the two config.workload strings)
    assert d["config"]["workload"].startswith("configs[3] synthetic 3D space frame 4x4 bays x 3 storeys")
    assert d["config"]["dof"] == 6 * 5 * 5 * 4 and d["config"]["max_rel_residual"] < 1e-10


def test_reference_arm_honours_steps_and_the_wall_budget():
    p = _run(["--impl", "reference", "--steps", "3", "--warmup", "2", *SMALL])
    d = json.loads(p.stdout.strip())
    assert (d["steps"], d["warmup"]) == (3, 2)
    p = _run(["--impl", "reference", "--steps", "50", "--warmup", "5", *SMALL], {"FERS_REF_BUDGET_S": "0.0001"})
    d = json.loads(p.stdout.strip())
    assert d["steps"] == 1 and d["warmup"] == 1 and d["steps_requested"] == 50 and d["warmup_requested"] == 5


def test_other_ranks_of_the_reference_arm_stay_silent():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", *SMALL], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = _run(["--steps", "1", "--warmup", "1", "--nx", "2", "--nz", "2", "--ny", "2", "--no-cpu-baseline", "--no-e2e"])
    assert p.returncode != 0 and p.stdout.strip() == ""
