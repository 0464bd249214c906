"""One-step comparison pair kernel vs scalar tile kernel from the SAME developed state (column with the largest deviation)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
nc, T0 = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 250
hs = {}
for name, no_pair in (("scalar", 1), ("pair", 0)):
    cs = M.k1d_ensemble_case(np.float32, nc=nc, n_elem=128, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                             order="sorted", ic="device")
    cs.cfg.reserved[3] = no_pair
    hs[name] = M.make_handle(cs, Handle)
hs["scalar"].set_steps_per_launch(50)
hs["scalar"].step(0.0, T0)
Y0 = hs["scalar"].get_state()
hs["pair"].set_state(Y0)
for n in (1, 1, 8):
    out = {}
    for name in hs:
        hs[name].step(float(T0), n)
        out[name] = hs[name].get_state().astype(np.float64)
    dS, dP = out["scalar"] - Y0, out["pair"] - Y0
    print(f"--- {n} step(s) from t={T0}")
    for i, v in enumerate(cs.var_names):
        inc = np.abs(dS[:, i]).max()
        if inc == 0: continue
        d = np.abs(out["pair"][:, i] - out["scalar"][:, i])
        c, k = np.unravel_index(np.argmax(d), d.shape)
        print(f"  {v}: max|pair-scalar| {d.max():.3e} = {d.max()/inc:.2e} of the largest increment; at column {c} level {k}: "
              f"increments scalar {dS[c, i, k]:.6e} pair {dP[c, i, k]:.6e}")
    for name in hs: hs[name].set_state(Y0)
