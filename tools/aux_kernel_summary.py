"""Summarise an ncu metric list (gpu__time_duration, dram bytes, fp64 pipe instructions) of the kernels around the factorisation
kernels -- extend-add (k_front_init), row panels, substitution (k_fwd_* / k_bwd_* / k_wave2) -- per kernel over ONE pass
(the launches between the first and the second k_front_init of the smallest grid), against the HBM and fp64 peaks.
Usage: tools/aux_kernel_summary.py gpurun_out/<tag>_aux_kernels.csv.gz [hbm_gbs] [fp64_tflops] > profiles/<tag>_aux_kernels.json"""
import collections
import csv
import gzip
import io
import json
import sys

path = sys.argv[1]
hbm = float(sys.argv[2]) if len(sys.argv) > 2 else 6556.0
fp64 = float(sys.argv[3]) if len(sys.argv) > 3 else 37.1
rows = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
hdr, data = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        key = (int(d["ID"]), d["Kernel Name"].split("(")[0].replace("fers::", ""), d["Grid Size"])
        data.setdefault(key, {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
keys = list(data)
# one pass = from the first launch to the launch before the first grid repeats for k_front_init
first = next(k for k in keys if k[1] == "k_front_init")
end = next((i for i, k in enumerate(keys) if i > 0 and k[1] == "k_front_init" and k[2] == first[2]), len(keys))
per = collections.OrderedDict()
for k in keys[:end]:
    m = data[k]
    p = per.setdefault(k[1], dict(launches=0, ms=0.0, dram_gb=0.0, fp64_warp_inst=0.0, rows=[]))
    p["launches"] += 1
    p["ms"] += m["gpu__time_duration.sum"] * 1e-6
    p["dram_gb"] += (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) * 1e-9
    p["fp64_warp_inst"] += m["sm__inst_executed_pipe_fp64.sum"]
    p["rows"].append(dict(grid=k[2], us=round(m["gpu__time_duration.sum"] * 1e-3, 1),
                          dram_gb=round((m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) * 1e-9, 3),
                          gbs=round((m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / m["gpu__time_duration.sum"], 0)))
out = dict(source=path, note="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum "
           "--clock-control none on bench.py --steps 1 --warmup 1; one pass; serialised, cold-cache launch times.  The fp64 pipe "
           "counter does not see DMMA (tensor pipe) instructions: it is ~0 for the substitution and row-panel kernels, whose "
           "arithmetic runs there; their flops are stated from the factor size instead (substitution_*)", hbm_peak_gbs=hbm, fp64_peak_tflops=fp64, kernels=[])
for name, p in per.items():
    gbs = p["dram_gb"] / (p["ms"] * 1e-3) if p["ms"] else 0.0
    out["kernels"].append(dict(kernel=name, launches=p["launches"], ms=round(p["ms"], 3), dram_gb=round(p["dram_gb"], 3),
                               dram_gbs=round(gbs, 0), frac_hbm=round(gbs / hbm, 3),
                               per_launch=p["rows"] if name in ("k_front_init", "k_fwd_bnd", "k_bwd_bnd") else None))
# substitution: 2 flops per factor entry and right-hand side in each direction (L21 dominates)
factor_entries = float(sys.argv[4]) if len(sys.argv) > 4 else 16.145431648e9 / 8.0
nrhs = 64
for direction, names in (("forward", ("k_fwd_gather", "void k_wave2<0>", "k_fwd_bnd")), ("backward", ("k_bwd_gather", "void k_wave2<1>", "k_bwd_bnd"))):
    ms = sum(per[n]["ms"] for n in names if n in per)
    gf = 2.0 * factor_entries * nrhs * 1e-9
    out["substitution_" + direction] = dict(kernels=list(names), ms=round(ms, 3), gflop=round(gf, 1), tflops=round(gf / ms, 2) if ms else 0.0,
                                            frac_fp64_peak=round(gf / ms / fp64, 3) if ms else 0.0,
                                            upper_levels_only=dict(kernel=names[2], ms=round(per[names[2]]["ms"], 3),
                                                                   tflops_if_all_flops_were_here=round(gf / per[names[2]]["ms"], 2)))
json.dump(out, sys.stdout, indent=1)
