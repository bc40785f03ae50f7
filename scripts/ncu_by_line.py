#!/usr/bin/env python
"""Attribute an ncu capture of step_kernel to CUDA source lines (run where the .ncu-rep and the .so are).
usage: python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep 'step_kernelIfLi2ELi64' [top_n] [regions] [kernel-regex] [source-file]
       (regions = name:lo-hi,... line ranges of the source file; kernel-regex default step_kernel, source file default
       step_kernel.cuh -- e.g. warp_step_kernel warp_kernel.cuh for the one-warp-per-environment kernel)"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
kre = sys.argv[5] if len(sys.argv) > 5 else "step_kernel"
srcname = sys.argv[6] if len(sys.argv) > 6 else "step_kernel.cuh"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "dbsgym_b200", "csrc", "libdbsgym.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):      # one cubin per translation unit
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
    hits = [i for i, l in enumerate(dis) if l.startswith(".text.") and pat in l]
    if hits:
        break
start = hits[0]
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2).strip(), cur))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
h = rows[hi]; idx = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
ncu = []
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    try:
        ncu.append((r[idx["Source"]], int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]]),
                    {s: int(r[idx[s]] or 0) for s in stalls}))
    except ValueError:
        pass
if len(ncu) > len(seq) and len(ncu) % len(seq) == 0:
    ncu = ncu[:len(seq)]          # several captured launches of the same kernel: use the first
assert len(ncu) == len(seq), (len(ncu), len(seq))
ex, sm, ops, ops_s, st = defaultdict(int), defaultdict(int), Counter(), Counter(), Counter()
for (sass, loc), (s2, e, ns, sd) in zip(seq, ncu):
    ex[loc] += e; sm[loc] += ns
    op = (s2.split()[1] if s2.startswith("@") else s2.split()[0]).split(".")[0]
    ops[op] += e; ops_s[op] += ns
    for k, v in sd.items():
        st[k] += v
tot, ts = sum(ex.values()), sum(sm.values())
src = open(os.path.join(root, "dbsgym_b200", "csrc", srcname)).read().split("\n")
print(f"total warp instructions {tot}, samples {ts}")
print("-- stall reasons"); T = sum(st.values())
for k, v in st.most_common(8):
    print(f"   {k:26s} {100 * v / T:5.1f}%")
print("-- opcodes");
for op, v in ops.most_common(14):
    print(f"   {op:10s} instr {100 * v / tot:5.1f}%  samples {100 * ops_s[op] / ts:5.1f}%")
print("-- source lines")
for loc, v in sorted(sm.items(), key=lambda kv: -kv[1])[:top]:
    f, l = loc if loc else ("?", 0)
    text = src[l - 1].strip()[:90] if f == srcname and l > 0 else ""
    print(f"   {f}:{l:4d} instr {100 * ex[loc] / tot:5.2f}% samples {100 * v / ts:5.2f}% | {text}")

# ---- optional: samples grouped by code region of step_kernel.cuh (line ranges given as name:lo-hi,...)
if len(sys.argv) > 4 and sys.argv[4]:
    regions = []
    for item in sys.argv[4].split(","):
        name, rng = item.split(":"); lo, hi = rng.split("-"); regions.append((name, int(lo), int(hi)))
    agg_s, agg_i = Counter(), Counter()
    for loc, v in sm.items():
        f, l = loc if loc else ("?", 0)
        key = "other-file:" + f
        if f == srcname:
            key = "unassigned"
            for name, lo, hi in regions:
                if lo <= l <= hi:
                    key = name; break
        agg_s[key] += v; agg_i[key] += ex[loc]
    print("-- regions")
    for k, v in agg_s.most_common():
        print(f"   {k:28s} samples {100 * v / ts:5.1f}%  instr {100 * agg_i[k] / tot:5.1f}%")
