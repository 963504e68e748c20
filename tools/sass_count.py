#!/usr/bin/env python
"""Instruction mix of the kernels in libfemb200.so whose mangled name contains every given substring:
   python tools/sass_count.py rowsILi4ELi4ELb0ELb1ELb1ELb0ELb0ELb1   (static counts; loops are counted once)"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", "fempy_b200/csrc/libfemb200.so"], capture_output=True, text=True).stdout
name, mix = None, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name and all(k in name for k in sys.argv[1:]):
            tot = sum(mix.values())
            print(name, tot)
            print("  " + "  ".join(f"{k}:{v}" for k, v in mix.most_common(24)))
        name, mix = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and mix is not None:
        mix[m.group(1)] += 1
if name and all(k in name for k in sys.argv[1:]):
    print(name, sum(mix.values()))
    print("  " + "  ".join(f"{k}:{v}" for k, v in mix.most_common(24)))
