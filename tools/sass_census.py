#!/usr/bin/env python
"""Per-kernel census of the SASS mnemonics that identify the execution path (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA loads, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync.  usage: sass_census.py <lib.so> <out.txt>"""
import collections
import re
import subprocess
import sys

lib, out = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|HMMA|LDGSTS|MUFU)\b")
counts, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        m = pat.search(line)
        if m:
            counts[fn][m.group(1)] += 1
with open(out, "w") as f:
    f.write(f"# SASS mnemonic census of {lib} (cuobjdump -sass; static instruction counts per kernel)\n")
    f.write("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, HMMA = mma.sync\n")
    for fn, c in counts.items():
        if not c:
            continue
        name = re.sub(r"\(anonymous namespace\)::", "", demangle(fn))
        name = re.sub(r"\(.*$", "", name)
        f.write(f"{name:70s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())) + "\n")
print("wrote", out)
