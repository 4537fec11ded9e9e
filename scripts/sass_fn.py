#!/usr/bin/env python3
"""Extract one kernel's SASS from `cuobjdump -sass` output: python scripts/sass_fn.py lib.sass <mangled prefix> > out"""
import re, sys
fn = None
for line in open(sys.argv[1]):
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    if fn and fn.startswith(sys.argv[2]) and re.match(r'\s+/\*[0-9a-f]{4,6}\*/', line):
        print(line.rstrip()[:110])
