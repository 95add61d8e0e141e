#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo): where the code size goes."""
import collections, os, re, subprocess, sys, tempfile
binary, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
cnt = collections.Counter(); first = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    insec, cur = False, None
    for l in dis.splitlines():
        if l.startswith("\t.section"):
            insec = l.split(",")[0].endswith(".text." + kern); continue
        if not insec: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m:
            cnt[cur] += 1; first.setdefault(cur, int(m.group(1), 16))
tot = sum(cnt.values())
print(kern, tot, "instructions", tot * 16 / 1024, "KB")
for k, v in cnt.most_common(top):
    src = ""
    if k:
        p = os.path.join("alphazero-connect6_b200/csrc", k[0])
        if os.path.exists(p):
            L = open(p).read().splitlines(); src = L[k[1]-1].strip()[:90] if k[1]-1 < len(L) else ""
    print(f"{v:5d} {100*v/tot:5.1f}%  first@{first[k]:#06x} {k}  {src}")
