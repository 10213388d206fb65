"""Per-level, per-kind summary of an ncu launch list of one factorisation (tools/factor_once.py):
python tools/launch_levels.py gpurun_out/r1q_factor_launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, ig, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
start = [i for i, r in enumerate(rows) if "k_chol_scale" in r[ik]][0]
level = -1
prev = ""
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])        # (level, kind) -> launches, ms, tile-flops
for r in rows[start + 1:]:
    name = re.sub(r"^void ", "", r[ik]).split("(")[0].split("<")[0].replace("fers::", "")
    g = [int(x) for x in re.findall(r"\d+", r[ig])]
    ns = float(r[iv].replace(",", ""))
    if name == "k_front_init":
        level += 1
    kind = name
    flops = 0.0
    if name == "k_update":
        kind = "upd_inner" if prev == "k_trsm" else ("upd_next" if prev == "k_row_panel" else "upd_rest")
        kw = 64 if kind == "upd_inner" else 256
        flops = 2.0 * g[0] * 128 * g[1] * 64 * g[2] * kw
    elif name == "k_syrk":
        flops = 2.0 * g[0] * 128 * g[1] * 64 * g[2]          # x K (unknown here)
    a = agg[(level, kind)]
    a[0] += 1; a[1] += ns / 1e6; a[2] += flops
    prev = name
kinds = ["k_front_init", "k_potrf_inv", "k_trsm", "upd_inner", "k_row_panel", "upd_next", "upd_rest", "k_syrk"]
print("level " + " ".join(f"{k:>22s}" for k in kinds))
tot = collections.Counter()
for l in range(level + 1):
    cells = []
    for k in kinds:
        n, ms, fl = agg.get((l, k), (0, 0.0, 0.0))
        tot[k] += ms
        tf = f"{fl / ms / 1e9:5.1f}" if (ms > 0 and fl > 0 and k != "k_syrk") else "     "
        cells.append(f"{n:4d} {ms:8.2f}ms {tf}")
    print(f"{l:5d} " + " ".join(f"{c:>22s}" for c in cells))
print("total " + " ".join(f"{tot[k]:20.2f}ms" for k in kinds))
