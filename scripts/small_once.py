"""Run the batched small-GPR kernel (K6) a few times on the bench workload; used for quick timing and under ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multi_fidelity_gpflow_b200 import _lib  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
Xh, Yh = bench.load_hbs()
thh, nzh = bench.make_thetas(R, 1000)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
h = _lib.Handle(0)
h.set_stream(s.cuda_stream)
h.set_async(True)
X, Y = torch.from_numpy(Xh).to(dev), torch.from_numpy(np.ascontiguousarray(Yh)).to(dev)
th, nz = torch.from_numpy(thh).to(dev), torch.from_numpy(nzh).to(dev)
nprob = R * bench.NBINS
nlml = torch.empty(nprob, dtype=torch.float64, device=dev)
grad = torch.empty(nprob, 2 * bench.DIM + 4, dtype=torch.float64, device=dev)
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    h.gpr_batched_nlml_grad(X, Y, th, nz, nlml=nlml, grad=grad)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {ms:.3f} ms  {R / ms * 1e3:.0f} evals/s  {nprob * bench.ALG_FLOPS_PER_BIN / ms / 1e9:.3f} TFLOP/s alg", flush=True)
assert h.sync() == 0
print("checksum", float(nlml.sum()), float(grad.sum()))
