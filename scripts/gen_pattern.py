#!/usr/bin/env python3
"""Extract the 256x4 rBRIEF sampling table (data only) from the reference.

Reads the integer literals of `bit_pattern_31_` (reference
src/ORBextractor.cc:149-407, originally OpenCV's learned ORB pattern) and
emits them as a bare comma-separated list, 16 ints (= 4 test pairs) per line,
usable as the initialiser of an `int[1024]`.  Only run in the build container
(the reference tree does not exist on the GPU box); the output files are
committed.
"""
import re, sys, hashlib
src = open('/root/reference/src/ORBextractor.cc').read().split('\n')[148:407]
txt = '\n'.join(src)
txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
body = txt[txt.index('{') + 1: txt.rindex('}')]
vals = [int(v) for v in re.findall(r'-?\d+', body)]
assert len(vals) == 1024, len(vals)
lines = [', '.join(str(v) for v in vals[i:i + 16]) + ',' for i in range(0, 1024, 16)]
out = '\n'.join(lines) + '\n'
for path in sys.argv[1:]:
    open(path, 'w').write(out)
print('sha256', hashlib.sha256(out.encode()).hexdigest(), 'max|v|', max(abs(v) for v in vals))
