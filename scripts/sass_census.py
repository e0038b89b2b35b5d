#!/usr/bin/env python
"""SASS opcode census of libcir_b200.so: per kernel, how many tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM), TMA tensor / bulk
copies (UTMALDG / UBLKCP), legacy tensor instructions (HMMA) etc. the shipped binary contains.
    python scripts/sass_census.py > profiles/r2_sass_census.txt        (no GPU needed: cuobjdump -sass)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200", "cirtorch_b200", "libcir_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "MUFU", "MATCH", "REDUX",
         "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "LDGSTS", "SHFL", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
counts, total, cur = collections.OrderedDict(), collections.Counter(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
                total[w] += 1
print("SASS opcode census of %s" % os.path.relpath(LIB, ROOT))
print("architectures in the binary: %s" % ", ".join(arch))
print("tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, mma.sync -> HMMA (none expected)\n")
cols = [w for w in WATCH if total[w]]
print("%-58s %7s " % ("kernel", "instrs") + " ".join("%7s" % c for c in cols))
for k, c in counts.items():
    print("%-58s %7d " % (k[:58], c["_all"]) + " ".join("%7s" % (c[w] or ".") for w in cols))
print("%-58s %7s " % ("total", "") + " ".join("%7d" % total[w] for w in cols))
