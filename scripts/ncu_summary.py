"""Summarise an .ncu-rep (raw page + per-SASS-segment instruction counts) -- used to write profiles/*."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
nprob = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, un, v = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
for h, u, x in zip(hdr, un, v):
    if h in keep or "issue_stalled" in h and "per_issue_active" in h:
        print(f"{h},{u},{x}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iI]) for r in data)
print(f"# total warp-instructions {tot}  per problem {tot / nprob:.0f}")
c, cs = Counter(), Counter()
for r in data:
    toks = r[iS].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[op.split(".")[0]] += int(r[iI])
    cs[op.split(".")[0]] += int(r[iSm])
ts = sum(cs.values())
for op, n in c.most_common(22):
    print(f"# {op:10s} inst/prob {n / nprob:9.0f} {100 * n / tot:5.1f}%   stall-samples {100 * cs[op] / ts:5.1f}%")
