#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics, stall mix, opcode mix, hottest SASS lines.
    python tools/ncu_summary.py report.ncu-rep [kernel_index]
    python tools/ncu_summary.py raw.csv sass.csv          (the `--page raw --csv` / `--page source --print-source sass --csv` exports)"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
from_csv = rep.endswith(".csv")
idx = 0 if from_csv else (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
raw = open(rep).read() if from_csv else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; units = rows[1]; r = rows[2 + idx]
def g(k):
    return (r[hdr.index(k)] + " " + units[hdr.index(k)]) if k in hdr else "n/a"
for k in ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]:
    print(f"{k:70s} {g(k)}")
print("-- stalls per issue")
st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for v, h in sorted(st, reverse=True)[:9]:
    print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.3f}")
if from_csv and len(sys.argv) < 3:
    sys.exit(0)
src = open(sys.argv[2]).read() if from_csv else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
ks, cur, shdr = [], None, None
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "Kernel Name": cur = []; ks.append(cur); continue
    if row and row[0] == "Address": shdr = row; continue
    if cur is not None and row: cur.append(row)
k = ks[idx]; ie = shdr.index("Instructions Executed"); ss = shdr.index("# Samples")
tot = sum(int(x[ie]) for x in k); ts = sum(int(x[ss]) for x in k)
print(f"-- SASS: {len(k)} instructions, {tot} warp-instr executed, {ts} samples")
ops = collections.Counter(); sm = collections.Counter()
for x in k:
    t = x[1].split(); o = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
    ops[o] += int(x[ie]); sm[o] += int(x[ss])
for o, n in ops.most_common(18): print(f"   {o:24s} {100*n/tot:5.1f}% exec  {100*sm[o]/max(ts,1):5.1f}% samples")
print("-- hottest lines (samples)")
stc = [i for i, c in enumerate(shdr) if c.startswith("stall_") and "Not Issued" not in c]
for x in sorted(k, key=lambda x: -int(x[ss]))[:22]:
    top = sorted(((int(x[i]), shdr[i][6:]) for i in stc), reverse=True)[:2]
    print(f"   {int(x[ss]):5d}  {x[1][:60]:60s} {top}")
