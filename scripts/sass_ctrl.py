"""Decode the scheduling control fields (stall count, yield, barriers) of a kernel's SASS:
    cuobjdump -sass -fun <mangled> lib.so | python scripts/sass_ctrl.py [pattern]"""
import re, sys
lines = sys.stdin.read().split("\n")
out = []
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            ctrl = int(m2.group(1), 16) >> 41
            out.append((m.group(1), m.group(2).strip(), ctrl & 0xF, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3F))
            i += 2
            continue
    i += 1
pat = sys.argv[1] if len(sys.argv) > 1 else None
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 6
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 40
start = 0
if pat:
    start = [k for k, o in enumerate(out) if pat in o[1]][0]
for o in out[max(0, start - lo):start + hi]:
    print(o[0], o[1][:56].ljust(56), "stall", o[2], "y", o[3], "wb", o[4], "rb", o[5], "wait", format(o[6], "06b"))
