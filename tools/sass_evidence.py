#!/usr/bin/env python
"""Counts the Blackwell-native SASS mnemonics per kernel of libsetok_b200.so (B200_PROFILING.md, 'What proves a Blackwell-native
kernel'): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA loads, HMMA = legacy mma.sync.  CPU only (cuobjdump)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "setok_b200", "libsetok_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("setok::(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\((CUtensorMap|setok|__nv|float|int|unsigned|long|void|const).*", "", cur)
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    for key, pat in (("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UTCBAR", r"\bUTCBAR\b"),
                     ("SYNCS", r"\bSYNCS\b"), ("HMMA", r"\bHMMA\b"), ("MUFU", r"\bMUFU\b")):
        if re.search(pat, line):
            counts[cur][key] += 1
keys = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "HMMA"]
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---:|" * len(keys))
for k, c in counts.items():
    if c["UTC*MMA"] or c["UTMALDG"] or c["HMMA"] or "--all" in sys.argv:
        print(f"| `{k[:70]}` | " + " | ".join(str(c[x]) for x in keys) + " |")
