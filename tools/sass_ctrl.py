"""Decode the scheduling control bits (stall, write/read scoreboard, wait mask) of one kernel's SASS.
usage: python tools/sass_ctrl.py lib.so <mangled-name-substring> [grep-regex]
Layout (Volta..Blackwell 128-bit encoding, high word): stall = bits 41-44, yield 45, write barrier 46-48,
read barrier 49-51, wait mask 52-57, reuse 58-61."""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
rx = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on, cur = False, None
for l in txt:
    if "Function :" in l:
        on = pat in l
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if m:
        cur = (m.group(1), re.sub(r"\s+", " ", m.group(2)).strip())
        continue
    m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if m and cur:
        hi = int(m.group(1), 16)
        stall, wr, rd, wait = (hi >> 41) & 0xF, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3F
        s = f"{cur[0]} st{stall:2d} wr{'-' if wr == 7 else wr} rd{'-' if rd == 7 else rd} wait{wait:06b}  {cur[1]}"
        if rx is None or rx.search(cur[1]):
            print(s)
        cur = None
