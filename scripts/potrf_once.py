"""Run mfgp_potrf once (after one warm-up) on an SPD matrix; used under ncu for the launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
x = torch.randn(n, 64, dtype=torch.float64, device=dev)
a = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev)
h = _lib.Handle(0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
h.set_stream(s.cuda_stream)
h.set_async(True)
w = a.clone()
for r in range(reps):
    w.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    h.potrf_device(w, n, n)
    e1.record(s)
    torch.cuda.synchronize()
    print(f"rep {r}: {e0.elapsed_time(e1):.3f} ms  {n**3 / 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12:.2f} TFLOP/s")
assert h.sync() == 0
