"""Summarise the raw page of an `ncu --set full` capture of the k_oz_syrk launches of one factorisation
(ncu -i X.ncu-rep --page raw --csv > raw.csv) into the capture summary bench.py reads for roofline.traffic:
python tools/oz_ncu_raw_summary.py raw.csv "<ncu command>" <commit> "<note>" > profiles/rNN_oz_syrk_ncu_summary.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name):
            return i
    raise KeyError(name)


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0, "": 1.0}


def val(r, name):
    i = col(name)
    try:
        return float(r[i].replace(",", "")) * SCALE.get(units[i].strip(), 1.0) if r[i] else None
    except ValueError:              # "no data" for a counter ncu could not collect on that launch
        return None


launches = []
kernels = set()
for r in data:
    if "k_oz_syrk" not in r[col("Kernel Name")]:
        continue
    kernels.add(r[col("Kernel Name")].split("(")[0])
    ms = val(r, "gpu__time_duration.sum")
    l2sm = val(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
    launches.append({
        "ms": ms, "dram_read_GB": val(r, "dram__bytes_read.sum") / 1e9, "dram_write_GB": val(r, "dram__bytes_write.sum") / 1e9,
        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        "l2_to_sm_GB": l2sm / 1e9, "l2_to_sm_TBps": l2sm / 1e12 / (ms * 1e-3),
        "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
        "smem_data_bank_reads_pct": val(r, "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed"),
        "smem_data_bank_writes_pct": val(r, "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed"),
        "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")})
tot = sum(x["dram_read_GB"] + x["dram_write_GB"] for x in launches) * 1e9
ms = sum(x["ms"] for x in launches)
note = sys.argv[4] if len(sys.argv) > 4 else ""
name = launches_kernel = "k_oz_syrk"
print(json.dumps({
    "kernel": "int8 Schur / panel-group updates (tcgen05.mma kind::i8) of one factorisation of configs[3]: " + ", ".join(sorted(kernels)),
    "commands": [sys.argv[2]],
    "note": note,
    "launches_per_factorisation": len(launches), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / max(len(launches), 1),
    "duration_ms_total": ms, "dram_GBps_average": tot / 1e9 / (ms * 1e-3) if ms else None,
    "tensor_pipe_active_pct_time_weighted": (sum(x["tensor_pipe_active_pct"] * x["ms"] for x in launches if x["tensor_pipe_active_pct"] is not None)
                                             / max(sum(x["ms"] for x in launches if x["tensor_pipe_active_pct"] is not None), 1e-30)),
    "l2_to_sm_TBps_time_weighted": sum(x["l2_to_sm_TBps"] * x["ms"] for x in launches) / ms if ms else None,
    "captured_config": {"dof": 692106, "nrhs": 64, "frame": [60, 60, 30], "kernel_id": "k_oz_syrk<S=7>", "commit": sys.argv[3] if len(sys.argv) > 3 else ""},
    "launches": launches}, indent=1))
