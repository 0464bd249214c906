"""Step-by-step trace of one column: first step at which pair and scalar trajectories separate."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
nc, col = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 69
hs = {}
for name, no_pair in (("scalar", 1), ("pair", 0)):
    cs = M.k1d_ensemble_case(np.float32, nc=nc, n_elem=128, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                             order="sorted", ic="device")
    cs.cfg.reserved[3] = no_pair
    hs[name] = M.make_handle(cs, Handle)
iN = cs.var_names.index("N_liq")
prev = 0.0
for t in range(0, 100):
    for h in hs.values(): h.step(float(t), 1)
    Ys, Yp = hs["scalar"].get_state()[col].astype(np.float64), hs["pair"].get_state()[col].astype(np.float64)
    d = np.abs(Yp[iN] - Ys[iN]); k = int(np.argmax(d)); rel = d.max() / max(np.abs(Ys[iN]).max(), 1.0)
    if rel > 3 * prev + 1e-7 or t % 20 == 19:
        ks = np.nonzero(Ys[iN] > 1.0)[0]; kp = np.nonzero(Yp[iN] > 1.0)[0]
        print(f"t={t+1}: N_liq rel diff {rel:.2e} at k{k}: scalar {Ys[iN][k]:.5e} pair {Yp[iN][k]:.5e}; lowest level with N_liq>1: scalar {ks.min() if ks.size else -1} pair {kp.min() if kp.size else -1}")
    prev = max(prev, rel)
