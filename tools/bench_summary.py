"""Print a compact summary of bench.py JSON lines read from stdin (developer convenience)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        if line:
            print(line)
        continue
    d = json.loads(line)
    cfg = d.get("config", {})
    print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} value={d.get('value'):.4g} {d.get('unit')} ms/step={d.get('ms_per_step'):.2f} "
          f"iters={cfg.get('pcg_iterations')} resid={cfg.get('max_rel_residual')} launches={d.get('gpu_launches')}")
    roof = d.get("roofline") or {}
    for k in roof.get("kernels", []):
        print(f"    {k['kernel']:24s} {k['ms']:.4f} ms  {k['GB/s']:.0f} GB/s  frac={k['frac']:.3f}  share={k['share_of_step']:.3f}")
    if d.get("e2e"):
        print("    e2e", d["e2e"])
    if d.get("cpu_baseline"):
        print("    cpu", d["cpu_baseline"])
    if d.get("clocks"):
        print("    clocks", d["clocks"])
