"""Summarise gpurun_out/dist_chol_timeline_w*_rank*.json (scripts/dist_chol_bench.py ... timeline): per step the chain period
(W_k+1 - W_k), how long after W_k the bulk panel is everywhere (pan_k - W_k), the duration of the main stream's step
(done_k - max(pan_k, done_k-1)) and the time the main stream waited for the panel (max(0, pan_k - done_k-1))."""
import glob, json, re, sys
files = sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dist_chol_timeline_*_rank*.json"))
for f in files:
    tl = json.load(open(f))
    n = max(int(re.findall(r"\d+", k)[0]) for k in tl if k.startswith("W")) + 1
    g = lambda lab, k, d=None: tl.get(f"{lab}{k}", d)
    chain = [g("W", k + 1) - g("W", k) for k in range(n - 1)]
    stall = work = 0.0
    rows = []
    for k in range(n):
        prev = g("done", k - 1, tl.get("assembled", 0.0))
        pan = g("pan", k, g("W", k))
        start = max(pan, prev)
        stall += max(0.0, pan - prev)
        work += g("done", k) - start
        rows.append((k, g("W", k), pan - g("W", k), g("done", k) - start, max(0.0, pan - prev)))
    print(f"{f}: total {g('done', n - 1):.1f} ms | main stream busy {work:.1f} ms, waiting for panels {stall:.1f} ms | "
          f"chain period mean {sum(chain) / len(chain):.2f} ms (first 8: {sum(chain[:8]) / 8:.2f}, last 8: {sum(chain[-8:]) / 8:.2f})")
    if "-v" in sys.argv:
        for r in rows:
            print("   step %2d  W at %6.2f  panel +%5.2f  main step %5.2f  main waited %5.2f" % r)
