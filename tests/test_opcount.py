"""The operation-counting build of the oracle (oracle/opcount.hpp) evaluates the same rhs as the oracle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from kid_b200 import model as M
from kid_b200.lib import KidConfig
from oracle.oracle import OracleHandle

ROOT = Path(__file__).resolve().parents[1]


def test_counting_scalar_reproduces_the_oracle_rhs_and_counts_operations():
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "libkid_oracle_opcount.so"], check=True, capture_output=True)
    lib = C.CDLL(str(ROOT / "oracle" / "libkid_oracle_opcount.so"))
    lib.kidoracle_create.restype = C.c_void_p
    lib.kidoracle_create.argtypes = [C.POINTER(KidConfig)]
    P = C.c_void_p
    case = M.k1d_case(np.float64, nc=1, n_elem=24, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    o = M.make_handle(case, OracleHandle)
    o.step(0.0, 400)
    Y = np.ascontiguousarray(o.get_state())
    h = P(lib.kidoracle_create(C.byref(case.cfg)))
    tab = np.ascontiguousarray(case.params, dtype=np.float64)
    assert lib.kidoracle_set_params(h, tab.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(tab.size)) == 0
    r, T = np.ascontiguousarray(case.ρ_dry), np.ascontiguousarray(case.T)
    lib.kidoracle_set_profiles(h, r.ctypes.data_as(P), T.ctypes.data_as(P), C.c_int(0))
    for which, v in case.col_scalars.items():
        vv = np.ascontiguousarray(np.asarray(v, np.float64))
        lib.kidoracle_set_column_scalar(h, C.c_int(which), vv.ctypes.data_as(P))
    lib.kidoracle_set_state(h, Y.ctypes.data_as(P))
    dY = np.empty_like(Y)
    lib.kidoracle_opcount_reset()
    assert lib.kidoracle_rhs(h, C.c_double(400.0), dY.ctypes.data_as(P)) == 0
    out = (C.c_double * 7)()
    lib.kidoracle_opcount_get(out)
    assert np.allclose(dY, o.rhs(400.0), rtol=1e-12, atol=0)
    per_point = np.array(out[:]) / 24
    assert per_point[0] > 100 and per_point[1] > 100 and per_point[4] > 5  # adds, muls, transcendentals per grid point
