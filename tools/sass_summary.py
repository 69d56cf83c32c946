#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libgiga_b200.so (cuobjdump -sass): the evidence that the hot kernels are Blackwell-native
(tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, cp.async.bulk.tensor -> UTMALDG, f32x2 FMAs -> FFMA2).
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "giga_b200", "libgiga_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "HMMA", "REDG", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
fn, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r"\(.*", "", fn)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        for o in OPS:
            if op == o or (o in ("LDG", "STG", "LDS", "STS", "BAR", "SYNCS", "SHFL", "REDG") and op.startswith(o)):
                counts[fn][o] += 1
                total[o] += 1
                break
print("SASS opcode counts per kernel, libgiga_b200.so (sm_100a); cuobjdump -sass | tools/sass_summary.py")
print("%-64s " % "kernel" + " ".join("%7s" % o for o in OPS))
for fn, c in counts.items():
    print("%-64s " % fn[:64] + " ".join("%7d" % c[o] for o in OPS))
print("%-64s " % "TOTAL" + " ".join("%7d" % total[o] for o in OPS))
