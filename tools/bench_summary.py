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
        rate = (f"{k['GB/s']:.0f} GB/s" if "GB/s" in k else f"{k['TFLOP/s']:.2f} TFLOP/s" if "TFLOP/s" in k
                else f"{k.get('int8_TOP/s', 0.0):.0f} TOP/s")
        print(f"    {k['kernel'][:60]:60s} {k['ms']:.4f} ms  {rate}  frac={k['frac']:.3f}  share={k['share_of_step']:.4f}")
    for key in ("factor_ms", "substitution_ms", "factor_tflops_all_kernels", "class_ms_profiled_step"):
        if key in roof:
            print(f"    {key}: {roof[key]}")
    if d.get("e2e"):
        print("    e2e", d["e2e"])
    if d.get("cpu_baseline"):
        print("    cpu", d["cpu_baseline"])
    if d.get("clocks"):
        print("    clocks", d["clocks"])
