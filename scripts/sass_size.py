#!/usr/bin/env python
"""SASS footprint per kernel of an object file / .so (instructions are 16 B on sm_100a)."""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cnt = collections.Counter(); name = None
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m: name = m.group(1); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): cnt[name] += 1
for n, c in sorted(cnt.items(), key=lambda x: x[1]):
    print(f'{c*16/1024:8.1f} KB  {c:7d} instr  {n}')
