"""One-step increments of ONE column: pair vs scalar from the same state, at several start times."""
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
t = 0
for T0 in (40, 50, 60, 70, 80, 90, 100):
    hs["scalar"].step(float(t), T0 - t); t = T0
    Y0 = hs["scalar"].get_state()
    hs["pair"].set_state(Y0)
    hs["pair"].step(float(T0), 1)
    hs["scalar"].step(float(T0), 1); t += 1
    Ys, Yp = hs["scalar"].get_state()[col].astype(np.float64), hs["pair"].get_state()[col].astype(np.float64)
    Y0c = Y0[col].astype(np.float64)
    line = f"t={T0}: "
    for i, v in enumerate(cs.var_names):
        inc = np.abs(Ys[i] - Y0c[i]).max()
        if inc == 0: continue
        d = np.abs(Yp[i] - Ys[i]); k = int(np.argmax(d))
        line += f"{v} {d.max()/inc:.1e}@k{k} "
    print(line)
