"""Headline counters, stall reasons and the executed opcode mix of the first kernel in an .ncu-rep (ncu --set full)."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0   # work units per launch (e.g. warp-elements) for per-unit counts
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
d = {h: v for h, v in zip(raw[0], raw[2])}
for k in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"):
    print(f"{k:75s} {d.get(k)}")
for h, v in d.items():
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v.replace(",", "")) > 0.15:
        print(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()))
hdr = next(r for r in src if r and r[0] == "Address")
ie, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
ops, tot = collections.Counter(), 0
for r in src:
    if len(r) > ie and r[ie].isdigit():
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc].strip())
        op = ".".join(m.group(2).split(".")[:2]) if m else "?"
        ops[op] += int(r[ie]); tot += int(r[ie])
print(f"executed warp instructions {tot / 1e6:.1f} M = {tot / units:.2f} per unit")
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n / tot:6.3f}  {n / units:7.2f} per unit")
