"""The bench's Goku per-bin leg on its own and after the SVGP leg (to see whether its time depends on what ran before)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
cx = bench.Ctx(types.SimpleNamespace())
peak = 37.18e12
r = bench.leg_goku_per_bin(cx, peak); print("alone", round(r["ms"], 2), flush=True)
r = bench.leg_svgp_step(cx, peak); print("svgp", round(r["ms"], 3), round(r["per_call_ms"], 1), flush=True)
for i in range(3):
    r = bench.leg_goku_per_bin(cx, peak); print("after svgp", i, round(r["ms"], 2), flush=True)
# per-call spread
from multi_fidelity_gpflow_b200.data import PowerSpecs
ps = PowerSpecs().read_from_npz(os.path.join(bench.ROOT, "tests", "golden", "goku.npz"))
X, Y = ps.training_arrays()
P, d = Y.shape[1], X.shape[1] - 1
th = np.exp(0.2 * np.random.default_rng(7).standard_normal((P, 2 * d + 3))); nz = np.full(P, 1e-3)
h = cx.h; h.set_stream(None); h.set_async(False)
ts = []
for i in range(12):
    t0 = time.perf_counter(); h.gpr_batched_nlml_grad(X, Y, th, nz); ts.append((time.perf_counter() - t0) * 1e3)
print("per call ms", [round(t, 2) for t in ts])
