#!/usr/bin/env python
"""Attribute an ncu capture's per-instruction counters to CUDA source lines.

  python scripts/hotlines.py <prof.ncu-rep> <object-or-so> <mangled kernel name> [top N]

ncu's `--page source --csv` is per SASS instruction; nvdisasm -g on the same binary gives the source
line of every instruction (needs -lineinfo).  Joined by instruction offset inside the function.
Prints, per source line: share of executed warp instructions, share of stall samples, dominant stall.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, binary, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
line_of = {}
for cb in cubins:
    dis = subprocess.run(["nvdisasm", "-g", cb], capture_output=True, text=True).stdout
    insec, cur = False, None
    for l in dis.splitlines():
        if l.startswith("\t.section"):
            insec = (".text." + kern) in l and l.split(",")[0].endswith(kern)
            continue
        if not insec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), 0.0])
base = None
tot_i = tot_s = 0.0
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        addr = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
    except ValueError:
        continue
    if base is None:
        base = addr
    off = addr - base
    inst = float(d["Instructions Executed"] or 0)
    samp = float(d["# Samples"] or 0)
    thr = float(d["Thread Instructions Executed"] or 0)
    key = line_of.get(off, (None, ""))[0]
    a = agg[key]
    a[0] += inst
    a[1] += samp
    a[3] += thr
    for c in stall_cols:
        v = float(d[c] or 0)
        if v:
            a[2][c] += v
    tot_i += inst
    tot_s += samp
print(f"kernel {kern}: {tot_i:.4g} warp instructions, {tot_s:.0f} stall samples")
src_cache = {}
def src(key):
    if not key:
        return ""
    f, n = key
    for root in ("alphazero-connect6_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            src_cache.setdefault(p, open(p).read().splitlines())
            L = src_cache[p]
            return L[n - 1].strip()[:100] if n - 1 < len(L) else ""
    return ""
print(f"{'inst%':>6} {'samp%':>6} {'thr/inst':>8}  top stalls / source")
for key, (i, s, st, thr) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    tops = ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in st.most_common(2))
    print(f"{100 * i / tot_i:6.2f} {100 * s / max(tot_s, 1):6.2f} {thr / max(i, 1):8.1f}  {key}  [{tops}]  {src(key)}")
