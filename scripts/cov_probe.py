import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream); h.set_async(True)
out = {}
for n, d in ((16384, 10), (16384, 5), (32768, 10)):
    X = torch.rand(n, d + 1, dtype=torch.float64, device=dev)
    X[:, -1] = (torch.arange(n, device=dev) >= n * 7 // 8).double()
    th = torch.ones(2 * d + 3, dtype=torch.float64, device=dev)
    K = torch.empty(n, n, dtype=torch.float64, device=dev)
    X2 = X.clone()
    for name, x2 in (("sym", None), ("rect", X2)):
        def run():
            assert _lib._lib.mfgp_cov(h._h, _lib._ptr(X), n, None if x2 is None else _lib._ptr(x2), n, d, _lib._ptr(th), _lib._ptr(K), n) == 0
        for _ in range(2): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): run()
        e1.record(s); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / 5
        gbs = (8.0 * n * n + 8.0 * 2 * n * (d + 1)) / t / 1e9
        out[f"cov_{name}_N{n}_d{d}_GBs"] = gbs
        print(f"N={n} d={d} {name}: {t*1e3:.3f} ms {gbs:.0f} GB/s ({gbs/6547.8*100:.1f}% of measured HBM)", flush=True)
    del K
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cov_probe.json", "w"), indent=1)
