#!/usr/bin/env python
"""Turn one `ncu --set full --import-source on` capture of the step kernel into the small JSON bench.py reads
(profiles/r02_step_kernel_variant<V>_counts.json): DRAM bytes per launch, executed warp instructions and executed FP32 flop
(from the per-SASS-instruction execution counts of the source page: FFMA2 = 4, FFMA = 2, FADD2 / FMUL2 = 2, FADD / FMUL = 1
flop per lane, 32 lanes per warp instruction).

usage: python scripts/ncu_counts.py <capture.ncu-rep> <variant> <n_envs> [kernel-regex]"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import Counter

rep, variant, n_envs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kre = sys.argv[4] if len(sys.argv) > 4 else "step_kernel"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr = raw[0]
row = raw[2] if len(raw) > 2 else raw[1]          # row 1 holds the units


def metric(name):
    v = row[hdr.index(name)].replace(",", "")
    return float(v)


unit = {n: raw[1][i] for i, n in enumerate(hdr)}


def to_bytes(name):
    v, u = metric(name), unit[name].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


src = list(csv.reader(io.StringIO(ncu("source"))))
hi = [i for i, r in enumerate(src) if "Source" in r and "Instructions Executed" in r][0]
h = src[hi]
ops = Counter()
for r in src[hi + 1:]:
    if len(r) < len(h):
        continue
    try:
        n = int(r[h.index("Instructions Executed")])
    except ValueError:
        continue
    sass = r[h.index("Source")].strip()
    tok = sass.split()
    if not tok:
        continue
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    ops[op] += n
flop_per_lane = {"FFMA2": 4, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FADD": 1, "FMUL": 1}
flop = sum(ops[o] * f for o, f in flop_per_lane.items()) * 32
total_inst = sum(ops.values())
out = {
    "source": f"profiles/{os.path.basename(rep)} (ncu --set full --clock-control none, one launch at {n_envs} environments)",
    "variant": variant, "n_envs": n_envs,
    "kernel": row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else kre,
    "duration_us_under_ncu": metric("gpu__time_duration.sum") / (1e3 if unit["gpu__time_duration.sum"] in ("nsecond", "ns") else 1),
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "warp_inst_executed": total_inst, "warp_inst_per_env_step": total_inst / n_envs,
    "fp32_flop": flop, "fp32_flop_per_env_step": flop / n_envs,
    "opcodes_top": {o: n for o, n in ops.most_common(16)},
    "registers_per_thread": metric("launch__registers_per_thread"),
    "issue_slots_busy_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr else None,
}
for key, name in (("fma_pipe_cycles_active_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                  ("xu_pipe_inst_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                  ("lsu_pipe_inst_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                  ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                  ("shared_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                  ("shared_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                  ("local_load_requests", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum")):
    if name in hdr:
        out[key] = metric(name)
path = os.path.join(root, "profiles", f"r02_step_kernel_variant{variant}_counts.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
