"""K2D (BASELINE config 5: 4096 horizontal nodes x 256 levels, Precipitation2M, Float32) on N ranks, strong scaling:
x-decomposition by whole elements, NCCL ring exchange of the GLL edge columns every SSPRK33 stage.
Run: python -m torch.distributed.run --nproc-per-node N scripts/bench_k2d.py [--weak]   (N = 1 works without torchrun)."""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
import torch
from kid_b200 import k2d, model as M
from kid_b200.lib import Handle

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=200); ap.add_argument("--weak", action="store_true")
ap.add_argument("--exchange", default="lib", choices=["lib", "torch"]); ap.add_argument("--ft", default="f32"); ap.add_argument("--tile-c", type=int, default=0); ap.add_argument("--nw", type=int, default=0); ap.add_argument("--legacy", type=int, default=0, help="1: two legacy horizontal launches per stage")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
FT = np.float32 if a.ft == "f32" else np.float64
nx = 1024 * (world if a.weak else 1)
kw = dict(nx=nx, nz=256, npoly=3, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
          domain_width=48000.0 * (world if a.weak else 1), domain_height=3000.0)
cs = k2d.k2d_case(FT, rank=rank, world=world, device=local, **kw)
if a.tile_c: cs.cfg.reserved[1] = a.tile_c
if a.nw: cs.cfg.reserved[0] = a.nw
if a.legacy: cs.cfg.reserved[5] = 1
h = M.make_handle(cs, Handle)
ring = ((k2d.LibraryRing if a.exchange == "lib" else k2d.DistributedRing)(h, dist, rank, world)) if world > 1 else None
def step(t0, n):
    if ring: ring.step(t0, n)
    else: h.step(t0, n)
step(0.0, 20); h.synchronize(); torch.cuda.synchronize()
if dist: dist.barrier()
t = time.perf_counter(); step(20.0, a.steps); h.synchronize(); torch.cuda.synchronize()
if dist: dist.barrier()
dt = time.perf_counter() - t
tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
if dist: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
Y = h.get_state()
if rank == 0:
    n_nodes = nx * 4
    print(json.dumps({"workload": "K2D NonEq+2M", "nodes": n_nodes, "levels": 256, "ft": a.ft, "n_gpus": world, "exchange": a.exchange if world > 1 else None, "scaling": "weak" if a.weak else "strong",
                      "ms_per_step": 1e3 * tt.item() / a.steps, "gp_steps_per_s": n_nodes * 256 * a.steps / tt.item(), "finite": bool(np.isfinite(Y).all())}))
if dist: dist.barrier(); dist.destroy_process_group()
