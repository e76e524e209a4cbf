"""torchrun script: time the distributed exact-GP NLML (2-D block-cyclic Cholesky with look-ahead over NCCL).

    torchrun ... scripts/dist_chol_bench.py N nb [P Q] [nolookahead] [profile]"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
from multi_fidelity_gpflow_b200.dist_chol import GpuOps, distributed_gpr_nlml
from oracle import mfgp_oracle as onp

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
h = _lib.Handle(rank)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nbd = int(sys.argv[2]) if len(sys.argv) > 2 else 512
grid = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 and sys.argv[3].isdigit() else None
la = "nolookahead" not in sys.argv
wg = "grad" in sys.argv
ds = onp.synthetic_exact_dataset(N)
res = {}
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    gops = globals().setdefault("gops", None) or GpuOps(h)
    globals()["gops"] = gops
    gops.stats = {}
    v = distributed_gpr_nlml(gops, ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nbd, grid=grid, lookahead=la, want_grad=wg)
    if wg:
        v, gvec = v
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"world={world} N={N} nbd={nbd} rep {rep}: {dt*1e3:.1f} ms  {N**3/3/dt/1e12:.2f} TFLOP/s (potrf flops)  nlml={v:.6f}  host-issue {gops.stats.get('host_issue_s', 0)*1e3:.1f} ms", flush=True)
        res = {"with_gradient": wg, "alg_tflops_nlml_grad": (N**3 + 4 * N**2) / dt / 1e12 if wg else None, "world": world, "grid": list(grid) if grid else "auto", "lookahead": la, "N": N, "nbd": nbd, "sec": dt, "potrf_tflops": N**3 / 3 / dt / 1e12, "nlml": v}
if "timeline" in sys.argv:  # device timestamps of the schedule's events on every rank (one extra call)
    gops.timeline = {}
    distributed_gpr_nlml(gops, ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nbd, grid=grid, lookahead=la)
    tl = gops.timeline_ms()
    gops.timeline = None
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(tl, open(f"gpurun_out/dist_chol_timeline_w{world}_N{N}_nb{nbd}_rank{rank}.json", "w"))
if "profile" in sys.argv:
    prof = {}
    distributed_gpr_nlml(h, ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nbd, grid=grid, lookahead=la, profile=prof)
    if rank == 0:
        print("phase seconds (serialised):", {k: round(v, 4) for k, v in prof.items()}, flush=True)
        res["profile"] = prof
if rank == 0:
    h.set_stream(None)
    if N <= 16384:
        t0 = time.perf_counter(); single = h.gpr_nlml(ds["X"], ds["Y"], ds["theta"], ds["noise"]); dt1 = time.perf_counter() - t0
        res["single_gpu_nlml"] = single; res["single_gpu_sec"] = dt1
        print("single-GPU nlml", single, f"{dt1*1e3:.1f} ms", "rel diff", abs(single - res["nlml"]) / abs(single))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/dist_chol2d_w{world}_N{N}_nb{nbd}" + ("" if la else "_nola") + ("_grad" if wg else "") + ".json", "w"))
dist.destroy_process_group()
