#!/usr/bin/env python
"""Print the SASS of the K2 pixel loop (from the loop head to the back-edge) for one LPC."""
import re, subprocess, sys
lpc = sys.argv[1] if len(sys.argv) > 1 else "8"   # lanes per channel; append "b" for the bucket-maxima variant (e.g. 1b)
bm = "1" if lpc.endswith("b") else "0"
lpc = lpc.rstrip("b")
lib = [a for a in sys.argv[2:] if a != "-v"][0] if [a for a in sys.argv[2:] if a != "-v"] else "pngloss_b200/libpngloss_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
body = [f for f in funcs if f.startswith(f"_Z14pl_k2_quantizeILi{lpc}ELb{bm}E")][0]
ins = []
for ln in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# pixel loop = the innermost loop containing ATOMS.POPC.INC; find back-edge after it
ai = [k for k, (a, t) in enumerate(ins) if "ATOMS" in t][0]
be = None
for k in range(ai, len(ins)):
    m = re.search(r"BRA(\.U)?\s+.*0x([0-9a-f]+)", ins[k][1])
    if m and int(m.group(2), 16) < ins[ai][0]:
        be = k; head = int(m.group(2), 16); break
hi = [k for k, (a, t) in enumerate(ins) if a == head][0]
print(f"loop {head:#x}..{ins[be][0]:#x}: {be - hi + 1} static instructions")
if "-v" in sys.argv:
    for a, t in ins[hi:be + 1]:
        print(f"{a:05x} {t}")
