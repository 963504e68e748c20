"""Opcode histogram of an address range of one kernel's SASS.  usage: python tools/sass_loop.py <object> <mangled kernel> <lo hex> <hi hex>"""
import collections, re, subprocess, sys
obj, fun, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ops = collections.Counter(); n = 0
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if not m: continue
    a = int(m.group(1), 16)
    if lo <= a <= hi:
        t = re.sub(r"^@!?U?P\w+\s+", "", m.group(2).strip())
        ops[t.split()[0].split(".")[0]] += 1; n += 1
print(n, "instructions:", ", ".join(f"{k} {v}" for k, v in ops.most_common()))
