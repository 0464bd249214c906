"""A/B of the packed two-levels-per-thread 2M kernel (k1d_pair_kernel) against the scalar tile kernel on the same
ensemble: state differences after the same steps, and ms per step of both (cfg.reserved[3] = 1 opts out of the pair kernel)."""
import argparse, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
from parity import errors

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=65536)
ap.add_argument("--t-start", type=int, default=480)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--spl", type=int, default=10)
ap.add_argument("--moisture", default="NonEquilibriumMoisture")
ap.add_argument("--order", default="sorted")
a = ap.parse_args()

def run(no_pair):
    cs = M.k1d_ensemble_case(np.float32, nc=a.columns, n_elem=128, moisture_choice=a.moisture, precipitation_choice="Precipitation2M",
                             order=a.order, ic="device")
    cs.cfg.reserved[3] = no_pair
    h = M.make_handle(cs, Handle)
    h.set_steps_per_launch(60)
    h.step(0.0, a.t_start); h.synchronize()
    Y0 = h.get_state()
    h.set_steps_per_launch(a.spl)
    t = float(a.t_start)
    h.step(t, a.spl); t += a.spl; h.synchronize()
    t0 = time.perf_counter()
    h.step(t, a.steps); h.synchronize()
    dt = time.perf_counter() - t0
    return cs, h, Y0, h.get_state(), dt

csA, hA, Y0A, YA, dA = run(1)
csB, hB, Y0B, YB, dB = run(0)
print(f"scalar tile kernel: {dA*1e3/a.steps:.3f} ms/step  {a.columns*128*a.steps/dA:.3e} gp-steps/s")
print(f"pair kernel       : {dB*1e3/a.steps:.3f} ms/step  {a.columns*128*a.steps/dB:.3e} gp-steps/s   finite={np.isfinite(YB).all()}")
print("pair vs scalar at t_start      :", {k: float('%.2e' % v) for k, v in errors(Y0B, Y0A, csA.var_names).items()})
print("pair vs scalar at t_start+steps:", {k: float('%.2e' % v) for k, v in errors(YB, YA, csA.var_names).items()})
