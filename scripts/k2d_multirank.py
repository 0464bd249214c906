"""Multi-process K2D check (run under torchrun, one rank per GPU): the NCCL ring exchange of the GLL edge
columns must reproduce the single-domain result.  Prints max |difference| per rank."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from kid_b200 import k2d, model as M
from kid_b200.lib import Handle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kw = dict(nx=16, nz=32, npoly=4, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
          t1=300.0, w1=2.5, domain_height=2400.0, domain_width=4800.0)
FT = np.float32
mine = M.make_handle(k2d.k2d_case(FT, rank=rank, world=world, device=local, **kw), Handle)
exchange = sys.argv[1] if len(sys.argv) > 1 else "lib"
ring = (k2d.LibraryRing if exchange == "lib" else k2d.DistributedRing)(mine, dist, rank, world)
ring.step(0.0, 120)
mine.synchronize()
Y = mine.get_state()
whole = M.make_handle(k2d.k2d_case(FT, device=local, **kw), Handle)
whole.step(0.0, 120)
Yw = whole.get_state()
off, cnt = k2d.partition_elements(16, world, rank)
ref = Yw[off * 5:(off + cnt) * 5]
print(f"[{exchange}] rank {rank}/{world}: bitwise_equal={np.array_equal(Y, ref)} max_abs_diff={np.abs(Y - ref).max():.3e} finite={np.isfinite(Y).all()}", flush=True)
dist.barrier(); dist.destroy_process_group()
