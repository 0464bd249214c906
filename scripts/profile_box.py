"""Profiling driver for the Box model (BASELINE configs[1]): 1 M boxes, Precipitation1M, `--launches` launches of `--spl` steps."""
import argparse, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200.lib import Handle
ap = argparse.ArgumentParser()
ap.add_argument("--ft", default="f64"); ap.add_argument("--precip", default="Precipitation1M"); ap.add_argument("--rf", default="CliMA_1M")
ap.add_argument("--boxes", type=int, default=1 << 20); ap.add_argument("--spl", type=int, default=100); ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
FT = np.float32 if a.ft == "f32" else np.float64
cs = M.box_ensemble_case(FT, nc=a.boxes, n_ic=32, precipitation_choice=a.precip, rain_formation_choice=a.rf)
h = M.make_handle(cs, Handle)
h.step(0.0, 20); h.synchronize()
t0 = time.perf_counter()
for i in range(a.launches): h.step(20.0 + i * a.spl, a.spl)
h.synchronize(); dt = time.perf_counter() - t0
print(f"box {a.precip} {a.ft} {a.boxes} boxes: {dt*1e3/(a.launches*a.spl):.4f} ms/step, {a.boxes*a.launches*a.spl/dt:.3e} box-steps/s, finite={np.isfinite(h.get_state()).all()}")
