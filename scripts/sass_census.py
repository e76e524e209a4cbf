"""Static SASS census of libmfgp.so: per kernel, instruction count and the mnemonics that show which hardware path it uses
(DMMA = FP64 tensor path; UTMALDG = TMA tensor loads (cp.async.bulk.tensor); SYNCS = mbarrier; UBLKCP = cp.async.bulk;
LDGSTS = cp.async).  Usage: python scripts/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi_fidelity_gpflow_b200", "libmfgp.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, stats = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        stats[cur]["n"] += 1
        stats[cur][m.group(1).split(".")[0]] += 1
rows = []
for k, c in stats.items():
    rows.append((c["n"], c["DMMA"], c["DFMA"], c["UTMALDG"], c["SYNCS"], c["UBLKCP"], c["LDGSTS"], demangle(k)[:150]))
print(f"{'instr':>6} {'DMMA':>5} {'DFMA':>5} {'UTMALDG':>7} {'SYNCS':>5} {'UBLKCP':>6} {'LDGSTS':>6}  kernel")
for r in sorted(rows, key=lambda r: r[-1]):
    print(f"{r[0]:6d} {r[1]:5d} {r[2]:5d} {r[3]:7d} {r[4]:5d} {r[5]:6d} {r[6]:6d}  {r[7]}")
