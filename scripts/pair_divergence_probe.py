"""Where do the pair kernel and the scalar kernel part ways?  Steps both one step at a time on the smoke() ensemble and
prints, for the worst column, the first steps at which any variable differs by more than `tol` of its field maximum."""
import copy, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200.lib import Handle

col = int(sys.argv[1]) if len(sys.argv) > 1 else 55
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cs = M.k1d_ensemble_case(np.float32, nc=64, n_elem=128, ic="device", order="morton",
                         moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
g = M.make_handle(cs, Handle)
cs2 = copy.copy(cs); cs2.cfg = copy.copy(cs.cfg); cs2.cfg.reserved[3] = 1
s = M.make_handle(cs2, Handle)
names = cs.var_names
shown = 0
prev = None
for n in range(nsteps):
    g.step(float(n), 1); s.step(float(n), 1)
    Yg, Ys = g.get_state()[col].astype(np.float64), s.get_state()[col].astype(np.float64)
    rel = np.abs(Yg - Ys).max(axis=1) / np.maximum(np.abs(Ys).max(axis=1), 1e-30)
    if rel.max() > 2e-5 and shown < 12:
        v = int(np.argmax(rel)); k = int(np.argmax(np.abs(Yg[v] - Ys[v])))
        print(f"step {n+1}: worst {names[v]} rel {rel[v]:.2e} at level {k}: pair {Yg[v,k]:.7e} scalar {Ys[v,k]:.7e}; all:",
              {nm: float('%.1e' % r) for nm, r in zip(names, rel)})
        lo, hi = max(0, k - 2), min(128, k + 3)
        for vv in range(len(names)):
            if rel[vv] > 1e-6:
                print("     ", names[vv], "pair", Yg[vv, lo:hi], "scalar", Ys[vv, lo:hi])
        shown += 1
print("final:", {nm: float('%.1e' % r) for nm, r in zip(names, rel)})
