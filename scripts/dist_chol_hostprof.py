import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
from multi_fidelity_gpflow_b200 import _lib
import multi_fidelity_gpflow_b200.dist_chol as dc
from oracle import mfgp_oracle as onp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
h = _lib.Handle(rank)
ops = dc.GpuOps(h)
acc = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); acc[key] = acc.get(key, 0) + time.perf_counter() - t; acc[key + "_n"] = acc.get(key + "_n", 0) + 1; return r
    setattr(obj, name, g)
for n in ("gemm", "potrf_inv", "cov", "wait", "record", "use"): wrap(ops, n, n)
wrap(dist, "broadcast", "broadcast")
N, nb = int(sys.argv[1]), int(sys.argv[2])
ds = onp.synthetic_exact_dataset(N)
for rep in range(3):
    acc.clear(); ops.stats = {}
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    v = dc.distributed_gpr_nlml(ops, ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nb)
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"rep {rep}: total {dt*1e3:.1f} ms host-issue {ops.stats['host_issue_s']*1e3:.1f} ms", {k: (round(v*1e3, 2) if not k.endswith('_n') else v) for k, v in sorted(acc.items())}, flush=True)
dist.destroy_process_group()
