import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
nc, col = 1024, 69
hs = {}
for name, no_pair in (("scalar", 1), ("pair", 0)):
    cs = M.k1d_ensemble_case(np.float32, nc=nc, n_elem=128, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                             order="sorted", ic="device")
    cs.cfg.reserved[3] = no_pair
    hs[name] = M.make_handle(cs, Handle)
iN, iA, iL = cs.var_names.index("N_liq"), cs.var_names.index("N_aer"), cs.var_names.index("ρq_liq")
np.set_printoptions(linewidth=200, precision=4)
for t in range(0, 26):
    for h in hs.values(): h.step(float(t), 1)
    if t >= 19:
        for name, h in hs.items():
            Y = h.get_state()[col]
            print(f"t={t+1} {name:6s} N_liq[33:40] {Y[iN][33:40]}  q_liq[33:40] {Y[iL][33:40]}")
