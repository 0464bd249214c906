"""Profiling driver for ncu: pre-roll the ensemble to a developed state in ONE launch, then issue
`--launches` launches of `--spl` steps each (the ones to capture with `ncu -k regex:k1d_tile -s 1`)."""
import argparse, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200.lib import Handle

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=65536)
ap.add_argument("--nz", type=int, default=128)
ap.add_argument("--spl", type=int, default=2)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--t-start", type=int, default=480)
ap.add_argument("--ft", default="f32")
ap.add_argument("--moisture", default="NonEquilibriumMoisture")
ap.add_argument("--precip", default="Precipitation2M")
ap.add_argument("--nw", type=int, default=0)
ap.add_argument("--tile-c", type=int, default=0)
ap.add_argument("--no-fixed", type=int, default=0)
ap.add_argument("--order", default="hash", choices=["hash", "sorted", "uniform", "morton"])
ap.add_argument("--ic", default="host", choices=["host", "device"])
ap.add_argument("--no-pair", type=int, default=0, help="1: scalar tile kernel instead of the packed pair kernel")
ap.add_argument("--reorder", type=int, default=0, help="store the members in model.warp_coherent_order")
a = ap.parse_args()
FT = np.float32 if a.ft == "f32" else np.float64
if a.order == "uniform":
    cs = M.k1d_case(FT, nc=a.columns, n_elem=a.nz, moisture_choice=a.moisture, precipitation_choice=a.precip)
else:
    cs = M.k1d_ensemble_case(FT, nc=a.columns, n_profiles=16, n_elem=a.nz, moisture_choice=a.moisture, precipitation_choice=a.precip,
                             order=a.order, ic=a.ic)
if a.reorder: cs = M.reorder_columns(cs, M.warp_coherent_order(cs))
if a.nw: cs.cfg.reserved[0] = a.nw
if a.tile_c: cs.cfg.reserved[1] = a.tile_c
if a.no_fixed: cs.cfg.reserved[2] = 1
if a.no_pair: cs.cfg.reserved[3] = 1
h = M.make_handle(cs, Handle)
h.step(0.0, a.t_start); h.synchronize()
t = float(a.t_start)
h.set_steps_per_launch(a.spl)
t0 = time.perf_counter()
for _ in range(a.launches):
    h.step(t, a.spl); t += a.spl
h.synchronize()
dt = time.perf_counter() - t0
Y = h.get_state()
print(f"order={a.order} columns={a.columns} nz={a.nz} {a.ft} {a.precip}: {a.launches} launches x {a.spl} steps: {dt*1e3/(a.launches*a.spl):.2f} ms/step, "
      f"{a.columns*a.nz*a.launches*a.spl/dt:.3e} gp-steps/s, finite={np.isfinite(Y).all()}")
