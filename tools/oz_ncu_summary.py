"""Summarise an ncu --csv metrics pass over the k_oz_syrk launches of ONE factorisation into the capture summary bench.py
reads for roofline.traffic:
python tools/oz_ncu_summary.py metrics.csv "<ncu command>" <commit> > profiles/rNN_oz_syrk_ncu_summary.json"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
iid, ik, im, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "%": 1.0}
per = collections.OrderedDict()
for r in rows:
    if "k_oz_syrk" not in r[ik]:
        continue
    d = per.setdefault(r[iid], {"kernel": r[ik].split("(")[0]})
    d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu].strip(), 1.0)
launches = []
for d in per.values():
    launches.append({"dram_read_GB": d.get("dram__bytes_read.sum", 0) / 1e9, "dram_write_GB": d.get("dram__bytes_write.sum", 0) / 1e9,
                     "ms": d.get("gpu__time_duration.sum", 0),
                     "tensor_pipe_active_pct": d.get("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                     "l2_to_sm_GB": d.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0) / 1e9})
tot = sum(x["dram_read_GB"] + x["dram_write_GB"] for x in launches) * 1e9
ms = sum(x["ms"] for x in launches)
print(json.dumps({
    "kernel": "k_oz_syrk<7> (tcgen05.mma kind::i8 Schur updates of the upper tree levels of one factorisation of configs[3])",
    "commands": [sys.argv[2]],
    "note": "one launch per front of the levels whose Schur update takes the int8 path; DRAM bytes against the algorithmic minimum "
            "(digit planes once + read / write of the Schur complement) show the L2 reuse of the band tile order; the kernel is "
            "bound by the L2 -> shared memory operand stream (l2_to_sm_GB / ms), not by DRAM and not by the tensor pipe",
    "launches_per_factorisation": len(launches), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / max(len(launches), 1),
    "duration_ms_total": ms, "dram_GBps_average": tot / 1e9 / (ms * 1e-3) if ms else None,
    "captured_config": {"dof": 692106, "nrhs": 64, "frame": [60, 60, 30], "kernel_id": "k_oz_syrk<S=7>", "commit": sys.argv[3] if len(sys.argv) > 3 else ""},
    "launches": launches}, indent=1))
