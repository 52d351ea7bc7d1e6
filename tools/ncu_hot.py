#!/usr/bin/env python
"""tools/ncu_hot.py — hottest SASS instructions of an `ncu --set full --import-source on` capture:
   python tools/ncu_hot.py x.ncu-rep [top_n]   (stall samples per instruction with the stall reasons)"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:100], " total samples:", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]; n = int(r[idx["# Samples"]] or 0)
    why = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {n:7d} {100.0*n/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} " + " ".join(f"{s}={c}" for c, s in why if c))
