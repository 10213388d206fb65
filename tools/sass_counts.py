#!/usr/bin/env python
"""Instruction counts per kernel of libfers_b200.so from `cuobjdump -sass` (no GPU needed): the mnemonics that show which
hardware path a kernel uses -- DMMA (fp64 tensor pipe, mma.sync.m8n8k4.f64), UTCIMMA (tcgen05.mma kind::i8), LDTM (tcgen05.ld),
UBLKCP (cp.async.bulk: TMA bulk copies), LDGSTS (cp.async), SYNCS (mbarrier), DFMA.
    python tools/sass_counts.py [--out profiles/r2_sass_counts.json]"""
import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fers_core_b200", "libfers_b200.so")
WATCH = ("DMMA", "UTCIMMA", "UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "DFMA", "HMMA", "UTCBAR", "LDG", "STG", "LDS", "STS")

p = argparse.ArgumentParser()
p.add_argument("--out", default="")
a = p.parse_args()
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
cur = None
counts = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in WATCH:
            counts[cur][op] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
out = {}
for mangled, nice in zip(counts, demangled):
    nice = re.sub(r"\(.*", "", nice).replace("fers::", "").replace("void ", "")
    c = counts[mangled]
    if any(c[k] for k in ("DMMA", "UTCIMMA", "UBLKCP", "LDGSTS", "LDTM")) or c["_total"] > 800:
        out[nice] = {k: v for k, v in sorted(c.items())}
txt = json.dumps({"library": "fers_core_b200/libfers_b200.so (sm_100a)", "command": "cuobjdump -sass, mnemonic counts per kernel",
                  "kernels": out}, indent=1)
if a.out:
    open(os.path.join(ROOT, a.out), "w").write(txt + "\n")
for k, v in out.items():
    print(k, {x: y for x, y in v.items() if x != "_total" and x in ("DMMA", "UTCIMMA", "LDTM", "UBLKCP", "LDGSTS", "DFMA")}, v["_total"])
