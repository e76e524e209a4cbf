"""Secondary measurements on the BASELINE.json configs (not the headline bench line): objective+gradient
evaluations per second through the C-ABI with host buffers, next to the CPU oracle (torch-fp64 autograd,
all host threads).  Writes gpurun_out/configs.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_fidelity_gpflow_b200 import _lib  # noqa: E402
from oracle import mfgp_oracle as onp  # noqa: E402
from oracle import mfgp_oracle_torch as otc  # noqa: E402


def timeit(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    h = _lib.Handle(0)
    out = {"host_cores": os.cpu_count(), "gpu": torch.cuda.get_device_name(0)}
    cpu_reps = int(os.environ.get("CPU_REPS", "1"))
    for name, d in (("hbs", 5), ("goku", 10)):
        ds = onp.load_dataset(name)
        X, Y = ds["X"], ds["Y"]
        N, P = Y.shape
        th = onp.default_theta(d)
        # C2s: shared-kernel exact GPR NLML+grad (the reference's own multi-bin model)
        g = timeit(lambda: h.gpr_nlml_grad(X, Y, th, 1e-3), 20, 3)
        c = timeit(lambda: otc.gpr_lml_value_and_grad(X, Y, th, 1e-3), cpu_reps * 3, 1)
        out[f"{name}_shared_gpr_nlml_grad"] = {"gpu_evals_per_s": 1 / g, "cpu_evals_per_s": 1 / c, "N": N, "P": P}
        # C2b/C2g: one GP per bin, batched
        ths = np.tile(th, (P, 1))
        nz = np.full(P, 1e-3)
        g = timeit(lambda: h.gpr_batched_nlml_grad(X, Y, ths, nz), 5 if N > 64 else 50, 2)
        c = timeit(lambda: otc.gpr_batched_value_and_grad(X, Y[:, :8], ths[:8], nz[:8]), cpu_reps, 0) * P / 8
        out[f"{name}_per_bin_gpr_nlml_grad"] = {"gpu_evals_per_s": 1 / g, "cpu_evals_per_s": 1 / c, "N": N, "bins": P,
                                                "cpu_note": "8 bins timed, scaled to all bins"}
        # C3: single-bin SVGP (SeparateIndependent), full batch
        M = 50 if name == "hbs" else 300
        Z = ds[f"Z_kmeans{M}"]
        q_mu, q_sqrt = np.zeros((M, P)), np.tile(0.1 * np.eye(M), (P, 1, 1))
        g = timeit(lambda: h.svgp_elbo_grad(X, Y, Z, ths, None, q_mu, q_sqrt, 1.0), 5, 2)
        Pc = min(P, 8)
        c = timeit(lambda: otc.svgp_value_and_grad(X, Y[:, :Pc], Z, ths[:Pc], q_mu[:, :Pc], q_sqrt[:Pc], 1.0), cpu_reps, 0) * P / Pc
        out[f"{name}_singlebin_svgp_elbo_grad"] = {"gpu_evals_per_s": 1 / g, "cpu_evals_per_s": 1 / c, "M": M, "L": P, "B": N,
                                                   "cpu_note": f"{Pc} latents timed, scaled"}
        # C4: latent SVGP (LinearCoregionalization)
        L = 10 if name == "hbs" else 15
        W = onp.initialize_W(P, L, 0.4, 0.2)
        thl = np.tile(th, (L, 1))
        q_mu, q_sqrt = np.zeros((M, L)), np.tile(np.eye(M), (L, 1, 1))
        g = timeit(lambda: h.svgp_elbo_grad(X, Y, Z, thl, W, q_mu, q_sqrt, 1.0, scale=1.0), 5, 2)
        c = timeit(lambda: otc.svgp_value_and_grad(X, Y, Z, thl, q_mu, q_sqrt, 1.0, W, N), cpu_reps, 0)
        out[f"{name}_latent_svgp_elbo_grad"] = {"gpu_evals_per_s": 1 / g, "cpu_evals_per_s": 1 / c, "M": M, "L": L, "B": N}
        print(json.dumps({k: v for k, v in out.items() if k.startswith(name)}, indent=1), flush=True)
    # C5: synthetic exact GPR
    for N in (4096, 8192, 16384, 32768):
        ds = onp.synthetic_exact_dataset(N)
        g = timeit(lambda: h.gpr_nlml_grad(ds["X"], ds["Y"], ds["theta"], ds["noise"]), 2, 1)
        flops = N**3 + 4 * N**2
        out[f"synthetic_exact_gpr_N{N}"] = {"gpu_evals_per_s": 1 / g, "sec": g, "alg_tflops": flops / g / 1e12}
        print(N, out[f"synthetic_exact_gpr_N{N}"], flush=True)
    # device-resident Adam training of 49 bins x R restarts (MultiBinMFGP)
    from multi_fidelity_gpflow_b200.multibin import MultiBinMFGP

    ds = onp.load_dataset("hbs")
    for R, steps in ((1, 200), (64, 200), (4096, 50)):
        mdl = MultiBinMFGP(ds["X"], ds["Y"], num_restarts=R, seed=0, handle=h)
        mdl.optimize(max_iters=5, learning_rate=0.01)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mdl.optimize(max_iters=steps, learning_rate=0.01)
        dt = time.perf_counter() - t0
        out[f"hbs_device_adam_R{R}"] = {"steps": steps, "sec": dt, "adam_steps_per_s": steps / dt,
                                        "bin_gp_evals_per_s": steps * R * 49 / dt, "us_per_step": dt / steps * 1e6}
        print(R, out[f"hbs_device_adam_R{R}"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
