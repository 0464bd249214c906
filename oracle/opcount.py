"""MEASUREMENT INFRASTRUCTURE (CPU only): algorithmic operation counts per grid point per rhs evaluation, SURVEY §8(d).

Builds the oracle with its operation-counting scalar (oracle/opcount.hpp), brings a default KiD column of each style to
a developed state with the regular Float64 oracle, evaluates ONE rhs in the counting build and prints the counts per
grid point: {add/sub, mul, div, sqrt, exp/log/pow/cbrt/sin/cos, erf/Γ, comparisons+min/max/abs}.
Run: python oracle/opcount.py  (writes profiles/r01_opcounts.json)."""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path = [p for p in sys.path if Path(p or ".").resolve() != ROOT / "oracle"]  # `oracle` is the package, not this dir
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M  # noqa: E402
from kid_b200.lib import KidConfig  # noqa: E402
from oracle.oracle import OracleHandle  # noqa: E402

subprocess.run(["make", "-C", str(ROOT / "oracle"), "libkid_oracle_opcount.so"], check=True, capture_output=True)
lib = C.CDLL(str(ROOT / "oracle" / "libkid_oracle_opcount.so"))
lib.kidoracle_create.restype = C.c_void_p
lib.kidoracle_create.argtypes = [C.POINTER(KidConfig)]
for f in ("set_params", "set_profiles", "set_column_scalar", "set_state", "rhs"):
    getattr(lib, "kidoracle_" + f).restype = C.c_int
P = C.c_void_p
NAMES = ["add_sub", "mul", "div", "sqrt", "exp_log_pow_cbrt_trig", "erf_gamma", "cmp_minmax_abs"]


def count_rhs(case, Y, t):
    h = P(lib.kidoracle_create(C.byref(case.cfg)))
    tab = np.ascontiguousarray(case.params, dtype=np.float64)
    lib.kidoracle_set_params(h, tab.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(tab.size))
    r, T = np.ascontiguousarray(case.ρ_dry, np.float64), np.ascontiguousarray(case.T, np.float64)
    lib.kidoracle_set_profiles(h, r.ctypes.data_as(P), T.ctypes.data_as(P), C.c_int(int(r.ndim == 2)))
    for which, v in case.col_scalars.items():
        vv = np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float64), (case.cfg.nc,)))
        lib.kidoracle_set_column_scalar(h, C.c_int(which), vv.ctypes.data_as(P))
    y = np.ascontiguousarray(Y, np.float64)
    lib.kidoracle_set_state(h, y.ctypes.data_as(P))
    dY = np.empty_like(y)
    lib.kidoracle_opcount_reset()
    lib.kidoracle_rhs(h, C.c_double(t), dY.ctypes.data_as(P))
    out = (C.c_double * 7)()
    lib.kidoracle_opcount_get(out)
    return np.array(out[:]) / (case.cfg.nc * case.cfg.nz), dY


rows = []
for moisture, precip in [("EquilibriumMoisture", "Precipitation0M"), ("NonEquilibriumMoisture", "Precipitation1M"),
                         ("EquilibriumMoisture", "Precipitation2M"), ("NonEquilibriumMoisture", "Precipitation2M")]:
    case = M.k1d_case(np.float64, nc=1, moisture_choice=moisture, precipitation_choice=precip)
    o = M.make_handle(case, OracleHandle)
    done = 0
    for t in (520, 1800):
        o.step(float(done), t - done)
        done = t
        Y = o.get_state()
        cnt, dY = count_rhs(case, Y, float(t))
        ref = o.rhs(float(t))
        assert np.allclose(dY, ref, rtol=1e-12, atol=0, equal_nan=True), "counting build disagrees with the oracle"
        flops = float(cnt[0] + cnt[1] + cnt[2] + cnt[3])  # add + mul + div + sqrt counted as one flop each
        rows.append(dict(style=f"{moisture} + {precip}", t=t, nvars=len(case.var_names),
                         per_gridpoint_per_rhs={n: round(float(c), 1) for n, c in zip(NAMES, cnt)},
                         arithmetic_flops_per_rhs=round(flops, 1),
                         # flops/gp-step = 3·count_rhs + 3·nvars·4 (three stage combines)
                         arithmetic_flops_per_gp_step=round(3 * flops + 12 * len(case.var_names), 1),
                         transcendentals_per_gp_step=round(3 * float(cnt[4] + cnt[5]), 1)))
        print(json.dumps(rows[-1]))
(ROOT / "profiles" / "r01_opcounts.json").write_text(json.dumps(rows, indent=1))
