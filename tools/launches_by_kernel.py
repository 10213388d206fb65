"""By-kernel shares of ONE pass of the hot path out of an ncu launch list of bench.py:
python tools/launches_by_kernel.py launches.csv[.gz] "<the ncu command>" > profiles/rNN_launches_by_kernel.json
A pass = the launches from the last k_chol_scale but one to the last one (factorisation, substitution, residual check,
recovery, then load vectors / assembly / BC of the next pass)."""
import collections
import csv
import gzip
import json
import re
import sys

path = sys.argv[1]
fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
marks = [i for i, r in enumerate(rows) if "k_chol_scale" in r[ik]]
if len(marks) < 2:
    sys.exit("need two k_chol_scale launches (two passes) in the list")
seg = [r for r in rows[marks[-2]:marks[-1]] if "k_peak_" not in r[ik]]      # the live fp64-peak measurement is not part of a pass
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in seg:
    name = re.sub(r"^void ", "", r[ik]).replace("fers::", "")
    name = re.sub(r"\(.*$", "", name) if "<" not in name else name[:name.rfind(">") + 1] if ">" in name else name
    a = agg[name]
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) * scale.get(r[iu].strip(), 1e-6)
total = sum(a[1] for a in agg.values())
print(json.dumps({
    "command": sys.argv[2] if len(sys.argv) > 2 else "",
    "note": "launches from one k_chol_scale to the next = one full pass of the hot path; per-launch times are cold-cache and "
            "serialised (no look-ahead overlap under ncu), so shares, not absolutes, compare with bench.py's live CUDA-event numbers",
    "launches": len(seg), "total_ms": total,
    "kernels": [{"kernel": k, "launches": n, "ms": round(ms, 4), "share": round(ms / total, 5)}
                for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}, indent=1))
