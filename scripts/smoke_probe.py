"""Diagnose the smoke() configuration: pair kernel / scalar kernel / Float32 oracle, each against the Float64 oracle started
from the same device-built initial condition."""
import copy, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from kid_b200 import model as M
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
from parity import errors, column_integrated_number_error

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cs = M.k1d_ensemble_case(np.float32, nc=64, n_elem=128, ic="device", order="morton",
                         moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
g = M.make_handle(cs, Handle)
host = M.materialise(cs, g)
cs2 = copy.copy(cs); cs2.cfg = copy.copy(cs.cfg); cs2.cfg.reserved[3] = 1
g2 = M.make_handle(cs2, Handle)
o32 = M.make_handle(host, OracleHandle)
h64 = copy.copy(host); h64.cfg = copy.copy(host.cfg)
h64.cfg.float_type = M.float_enum(np.float64)
for a in ("ρ_dry", "T", "Y0"):
    setattr(h64, a, getattr(host, a).astype(np.float64))
h64.col_scalars = {k: v.astype(np.float64) for k, v in host.col_scalars.items()}
o64 = M.make_handle(h64, OracleHandle)
for h in (g, g2, o32, o64):
    h.step(0.0, nsteps)
Yp, Ys, Yo, Yt = g.get_state(), g2.get_state(), o32.get_state(), o64.get_state()
names = cs.var_names
fmt = lambda e: {k: float("%.2e" % v) for k, v in e.items()}
print("pair   vs f64:", fmt(errors(Yp, Yt, names)), column_integrated_number_error(Yp, Yt, names))
print("scalar vs f64:", fmt(errors(Ys, Yt, names)), column_integrated_number_error(Ys, Yt, names))
print("o32    vs f64:", fmt(errors(Yo, Yt, names)), column_integrated_number_error(Yo, Yt, names))
print("pair   vs o32:", fmt(errors(Yp, Yo, names)))
print("scalar vs o32:", fmt(errors(Ys, Yo, names)))
print("pair vs scalar:", fmt(errors(Yp, Ys, names)))
i = names.index("ρq_liq")
for tag, Y in (("pair", Yp), ("scalar", Ys), ("o32", Yo)):
    d = np.abs(Y[:, i].astype(np.float64) - Yt[:, i]) / np.abs(Yt[:, i]).max()
    c, k = np.unravel_index(np.argmax(d), d.shape)
    print(tag, "worst ρq_liq at column", c, "level", k, "err", d[c, k], "per-column max:", np.sort(d.max(axis=1))[-5:])
