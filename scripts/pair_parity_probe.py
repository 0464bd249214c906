"""Pair kernel vs scalar tile kernel vs CPU oracle on the same sorted ensemble (error tables per variable)."""
import argparse, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
from parity import errors

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=2048)
ap.add_argument("--steps", type=int, default=520)
ap.add_argument("--moisture", default="NonEquilibriumMoisture")
a = ap.parse_args()
out = {}
for name, no_pair in (("scalar", 1), ("pair", 0)):
    cs = M.k1d_ensemble_case(np.float32, nc=a.columns, n_elem=128, moisture_choice=a.moisture, precipitation_choice="Precipitation2M",
                             order="sorted", ic="device")
    cs.cfg.reserved[3] = no_pair
    h = M.make_handle(cs, Handle)
    if name == "scalar":
        o = M.make_handle(M.materialise(cs, h), OracleHandle)
    h.set_steps_per_launch(40)
    h.step(0.0, a.steps)
    out[name] = h.get_state()
o.step(0.0, a.steps)
Yo = o.get_state()
f = lambda e: {k: float("%.2e" % v) for k, v in e.items()}
print("scalar vs oracle:", f(errors(out["scalar"], Yo, cs.var_names)))
print("pair   vs oracle:", f(errors(out["pair"], Yo, cs.var_names)))
print("pair   vs scalar:", f(errors(out["pair"], out["scalar"], cs.var_names)))
# where do the largest differences sit?
from kid_b200.lib import COL_W1, COL_PRESCRIBED_ND
for name in ("scalar", "pair"):
    print(name)
    for i, v in enumerate(cs.var_names):
        d = np.abs(out[name][:, i].astype(np.float64) - Yo[:, i])
        if d.max() == 0: continue
        c, k = np.unravel_index(np.argmax(d), d.shape)
        print(f"  {v}: max |d| {d.max():.3e} (field max {np.abs(Yo[:, i]).max():.3e}) at column {c} level {k}: w1={cs.col_scalars[COL_W1][c]:.3f} "
              f"Nd={cs.col_scalars[COL_PRESCRIBED_ND][c]:.3e} gpu={out[name][c, i, k]:.6e} oracle={Yo[c, i, k]:.6e}; columns with err > 1e-3 of max: "
              f"{int((d.max(axis=1) > 1e-3 * np.abs(Yo[:, i]).max()).sum())}")
