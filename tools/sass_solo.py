#!/usr/bin/env python
"""Print the SASS of the chain loop of pl_k2_solo<FPW> (the innermost loop that holds the VOTE of the channel fix-up)."""
import re, subprocess, sys
fpw = sys.argv[1] if len(sys.argv) > 1 else "5"
lib = "pngloss_b200/libpngloss_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
body = [f for f in out.split("Function : ") if f.startswith(f"_Z10pl_k2_soloILi{fpw}E")][0]
ins = []
for ln in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
vi = [k for k, (a, t) in enumerate(ins) if t.startswith("VOTE") or " VOTE" in t][0]
# smallest loop (back-edge after vi to a head before vi)
best = None
for k in range(vi, len(ins)):
    m = re.search(r"BRA(\.U)?\s+.*0x([0-9a-f]+)", ins[k][1])
    if m and int(m.group(2), 16) <= ins[vi][0]:
        best = (int(m.group(2), 16), k)
        break
head, be = best
hi = [k for k, (a, t) in enumerate(ins) if a == head][0]
print(f"loop {head:#x}..{ins[be][0]:#x}: {be - hi + 1} static instructions")
if "-v" in sys.argv:
    for a, t in ins[hi:be + 1]:
        print(f"{a:05x} {t}")
