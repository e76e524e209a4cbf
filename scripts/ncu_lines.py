"""Per-CUDA-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
nprob = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
fname = ""
recs = []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > iI and r[0].isdigit() and r[iI].isdigit():
        recs.append((fname, int(r[0]), r[1].strip()[:90], int(r[iI]), int(r[iS]) if r[iS].isdigit() else 0))
ti, ts = sum(x[3] for x in recs), sum(x[4] for x in recs)
print(f"# total inst/prob {ti / nprob:.0f}  samples {ts}")
for f, ln, src, ni, ns in sorted(recs, key=lambda x: -x[4])[:top]:
    print(f"{f}:{ln:4d} inst/prob {ni / nprob:7.0f} ({100 * ni / ti:4.1f}%)  samples {100 * ns / ts:4.1f}%  | {src}")

if len(sys.argv) > 4:  # phase table: "name:lo-hi,name:lo-hi,..." on the main .cu file; everything else goes to "other"
    main = sys.argv[5] if len(sys.argv) > 5 else recs[0][0]
    ph = [(n, int(a), int(b)) for n, ab in (x.split(":") for x in sys.argv[4].split(",")) for a, b in [ab.split("-")]]
    agg = {}
    for f, ln, src, ni, ns in recs:
        key = "other:" + f
        if f == main:
            key = next((n for n, a, b in ph if a <= ln <= b), "unassigned")
        a = agg.setdefault(key, [0, 0])
        a[0] += ni
        a[1] += ns
    print("# phase table")
    for k, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} inst/prob {ni / nprob:7.0f} ({100 * ni / ti:4.1f}%)  samples {100 * ns / ts:4.1f}%")
