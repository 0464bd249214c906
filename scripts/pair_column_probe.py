"""Follow one column: pair kernel vs scalar kernel vs oracle over time."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
nc, col = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 69
hs = {}
for name, no_pair in (("scalar", 1), ("pair", 0)):
    cs = M.k1d_ensemble_case(np.float32, nc=nc, n_elem=128, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                             order="sorted", ic="device")
    cs.cfg.reserved[3] = no_pair
    hs[name] = M.make_handle(cs, Handle)
    if name == "scalar":
        hs["oracle"] = M.make_handle(M.slice_columns(M.materialise(cs, hs[name]), col - col % 32, col - col % 32 + 32), OracleHandle)
t = 0
names = cs.var_names
for n in [100, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20]:
    for h in hs.values(): h.step(float(t), n)
    t += n
    Ys, Yp, Yo = hs["scalar"].get_state()[col].astype(np.float64), hs["pair"].get_state()[col].astype(np.float64), hs["oracle"].get_state()[col % 32].astype(np.float64)
    line = f"t={t:4d} "
    for v in ("ρq_liq", "ρq_rai", "N_liq", "N_rai"):
        i = names.index(v); m = max(np.abs(Yo[i]).max(), 1e-300)
        kp = int(np.argmax(np.abs(Yp[i] - Yo[i])))
        line += f"| {v}: sc {np.abs(Ys[i]-Yo[i]).max()/m:.1e} pr {np.abs(Yp[i]-Yo[i]).max()/m:.1e} @k{kp} "
    print(line)
