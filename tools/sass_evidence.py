#!/usr/bin/env python
"""tools/sass_evidence.py — per-kernel SASS mnemonic counts of libtsg.so (cuobjdump -sass), the
proof that the tensor-core path is tcgen05/TMEM/TMA and not a recompiled mma.sync kernel:
tcgen05.mma -> UTCHMMA, tcgen05.st/ld -> STTM/LDTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG (tensor) / UBLKCP (bulk),
mbarrier -> SYNCS, cluster barrier -> UCGABAR_*, fma.rn.f32x2 -> FFMA2."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ternary-spgemm_b200/libtsg.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = ["UTCHMMA", "STTM", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "ELECT",
        "FFMA2", "FFMA", "FADD", "LDS", "STS", "LDG", "STG", "ST", "SHFL", "POPC", "VOTE", "HMMA", "LOP3", "IMAD", "SHF"]
cnt, total, name = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        total[name] += 1
        if m.group(1) in KEEP:
            cnt[name][m.group(1)] += 1
print(f"# {lib}: instructions per kernel and selected mnemonics")
for k in sorted(total):
    print(f"{k}: {total[k]} instr; " + ", ".join(f"{op} {n}" for op, n in sorted(cnt[k].items())))
