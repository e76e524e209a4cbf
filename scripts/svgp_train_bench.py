"""SVGP training-step throughput: host optimize() loop (per-step H2D/D2H of every parameter) vs optimize_on_device()."""
import copy, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200.kernels import SquaredExponential
from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP
from multi_fidelity_gpflow_b200.singlebin_svgp import SingleBinSVGP
from oracle import mfgp_oracle as onp

out = {}
for name, d, M, Llat in (("hbs", 5, 50, 10), ("goku", 10, 300, 15)):
    ds = onp.load_dataset(name)
    X, Y = ds["X"], ds["Y"]
    P = Y.shape[1]
    k = lambda: (SquaredExponential(lengthscales=np.ones(d)), SquaredExponential(lengthscales=np.ones(d)))
    models = {"singlebin": SingleBinSVGP(X, Y, *k(), P, ds[f"Z_kmeans{M}"]),
              "latent": LatentMFCoregionalizationSVGP(X, Y, *k(), num_latents=Llat, num_inducing=M, num_outputs=P)}
    for mname, mdl in models.items():
        steps = 20 if name == "hbs" else 6
        a, b = copy.deepcopy(mdl), copy.deepcopy(mdl)
        a.optimize((X, Y), max_iters=2, initial_lr=0.005, verbose=False)
        t0 = time.perf_counter(); a.optimize((X, Y), max_iters=2 + steps, initial_lr=0.005, verbose=False) if mname == "latent" else a.optimize((X, Y), max_iters=steps, initial_lr=0.005, verbose=False); th = (time.perf_counter() - t0) / steps
        b.optimize_on_device((X, Y), max_iters=2, initial_lr=0.005)
        torch.cuda.synchronize(); t0 = time.perf_counter(); b.optimize_on_device((X, Y), max_iters=steps * 5, initial_lr=0.005); td = (time.perf_counter() - t0) / (steps * 5)
        out[f"{name}_{mname}"] = {"host_loop_ms_per_step": th * 1e3, "device_loop_ms_per_step": td * 1e3, "speedup": th / td}
        print(name, mname, out[f"{name}_{mname}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/svgp_train.json", "w"), indent=1)
