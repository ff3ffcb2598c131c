#!/usr/bin/env python3
"""Dynamic opcode mix and hottest basic blocks of a profiled kernel (ncu --import-source on):
    python tools/ncu_opmix.py report.ncu-rep [n_blocks]"""
import collections
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launch, ie = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        launch.append([r[1], []])
        continue
    if r and r[0] == "Address":
        ie = [i for i, h in enumerate(r) if h.startswith("Instructions Executed")][0]
        continue
    if launch and ie is not None and len(r) > ie:
        try:
            launch[-1][1].append((r[0], r[1].strip(), int(r[ie])))
        except ValueError:
            pass
def opcode(sass):
    m = re.match(r"(@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)", sass)
    return m.group(2) if m else "?"


name, L = launch[0]
tot = sum(x[2] for x in L)
print("kernel:", name[:120])
print("warp instructions executed:", tot)
op = collections.Counter()
for _, s, n in L:
    op[opcode(s)] += n
print("opcode mix:", ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in op.most_common(14)))
segs, cur = [], [L[0]]
for x in L[1:]:
    if abs(x[2] - cur[-1][2]) <= 0.02 * max(cur[-1][2], 1):
        cur.append(x)
    else:
        segs.append(cur)
        cur = [x]
segs.append(cur)
for s in sorted(segs, key=lambda s: -sum(x[2] for x in s))[: int(sys.argv[2]) if len(sys.argv) > 2 else 6]:
    n = sum(x[2] for x in s)
    ops = collections.Counter(opcode(x[1]) for x in s)
    print(f"  block @{s[0][0][-5:]}: {n / tot * 100:5.1f}% of executed, {len(s)} instrs x {s[0][2]:.3e}: "
          + ", ".join(f"{k} {v}" for k, v in ops.most_common(8)))
