"""Closes the oracle's parity gap when real reference output is available: compares the C++ oracle against the
states dumped by oracle/julia/dump_reference_states.jl (run on a machine with Julia + the reference's dependencies).
Skipped while tests/golden/julia/ holds no fixture — the oracle is then 'parity unpinned' (DESIGN.md §4)."""
import json
from pathlib import Path

import numpy as np
import pytest

from kid_b200 import model as M
from oracle.oracle import OracleHandle
from parity import errors

FIX = Path(__file__).parent / "golden" / "julia"
CASES = sorted(FIX.glob("*.json")) if FIX.exists() else []


@pytest.mark.skipif(not CASES, reason="no Julia reference fixtures (parity unpinned; see oracle/julia/dump_reference_states.jl)")
@pytest.mark.parametrize("meta_path", CASES, ids=lambda p: p.stem)
def test_oracle_matches_julia_reference(meta_path):
    meta = json.loads(meta_path.read_text())
    FT = np.float32 if meta["FT"] == "Float32" else np.float64
    toml = {k: v for k, v in meta["toml"].items()}
    cs = M.k1d_case(FT, nc=1, moisture_choice=meta["moisture"], precipitation_choice=meta["precipitation"],
                    rain_formation_scheme_choice=meta["rain_formation"], n_elem=meta["n_elem"], dt=meta["dt"],
                    toml_overrides={k: v for k, v in toml.items() if k in cs_known_names()})
    nz, nv = meta["n_elem"], len(meta["variables"])
    Yref = np.fromfile(meta_path.with_suffix("").with_suffix(".Y.bin"), dtype=FT).reshape(len(meta["t"]), nv, nz)
    cs.Y0[0] = Yref[0]
    cs.ρ_dry, cs.T = np.asarray(meta["rho_dry"], FT), np.asarray(meta["T"], FT)
    o = M.make_handle(cs, OracleHandle)
    dYref = np.fromfile(meta_path.with_suffix("").with_suffix(".dY.bin"), dtype=FT).reshape(nv, nz)
    e = errors(o.rhs(0.0), dYref[None], cs.var_names)
    assert max(e.values()) <= (1e-9 if FT == np.float64 else 1e-4), e
    t_prev = 0.0
    for it, t in enumerate(meta["t"][1:], start=1):
        o.step(t_prev, int(round((t - t_prev) / meta["dt"])))
        t_prev = t
        e = errors(o.get_state(), Yref[it][None], cs.var_names)
        assert max(e.values()) <= (1e-6 if FT == np.float64 else 5e-2), (t, e)


def cs_known_names():
    from kid_b200 import parameters as PR
    return set(PR.DEFAULTS)
