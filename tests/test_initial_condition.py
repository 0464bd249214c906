"""Golden-vector test: the initial profile against the PySDM arrays the reference ships
(reference test/initial_condition_tests/initial_profile_test.jl:88-114, data test/data_utils.jl;
fixture extracted by tests/golden/make_pysdm_fixture.py).  Same grid (222 cells, 0-3220 m),
same comparison heights (0:100:3220), same tolerances."""
import json
from pathlib import Path

import numpy as np
import pytest

from kid_b200 import initial_condition as IC
from kid_b200 import parameters as PR

GOLD = json.loads((Path(__file__).parent / "golden" / "pysdm_profiles.json").read_text())


def _line_interp(x, xp, fp):
    """Interpolations.linear_interpolation(..., extrapolation_bc = Line())"""
    x = np.asarray(x, dtype=float)
    y = np.interp(x, xp, fp)
    lo, hi = x < xp[0], x > xp[-1]
    y[lo] = fp[0] + (x[lo] - xp[0]) * (fp[1] - fp[0]) / (xp[1] - xp[0])
    y[hi] = fp[-1] + (x[hi] - xp[-1]) * (fp[-1] - fp[-2]) / (xp[-1] - xp[-2])
    return y


@pytest.mark.parametrize("dry", [True, False])
def test_initial_profiles_match_pysdm(dry):
    FT = np.float64
    td = PR.override_toml_dict(FT=FT)
    tp = PR.create_thermodynamics_parameters(td)
    cp = PR.create_common_parameters(td)
    kp = PR.create_kid_parameters(td)
    z_min, z_max, n_elem = 0.0, 3220.0, 222
    faces = np.linspace(z_min, z_max, n_elem + 1)
    zc = 0.5 * (faces[:-1] + faces[1:])
    ρ_profile = IC.ρ_ivp(FT, kp, tp, "ShipwayHill2012", dry=dry)
    ip = IC.initial_condition_1d(FT, cp, kp, tp, ρ_profile, zc, "ShipwayHill2012", dry=dry)
    sdm = {k: np.asarray(v) for k, v in GOLD["dry" if dry else "wet"].items()}
    sdm["rho_sdm"] = sdm["rhod_sdm"] * (1 + sdm["qv_sdm"])  # data_utils.jl:431
    z_test = np.arange(z_min, z_max + 1e-9, 100.0)
    km = lambda f: _line_interp(z_test, zc, f)  # noqa: E731
    sd = lambda k: _line_interp(z_test, sdm["z_sdm"], sdm[k])  # noqa: E731
    q_vap = ip["q_tot"] - ip["q_liq"] - ip["q_ice"]
    # The stored PySDM array is the water-vapour MIXING RATIO r_v (its first wet value is
    # rv_0 + (rv_1 - rv_0)/740*10 = 0.01498378), while the model carries specific humidity
    # q_v = r_v/(1+r_v); the reference's norm-wise isapprox(q_vap, qv_sdm, rtol=1e-4) can therefore
    # only hold after this conversion, which is applied here before using its tolerance.
    r_vap = q_vap / (1 - q_vap)
    np.testing.assert_allclose(km(r_vap), sd("qv_sdm"), rtol=1e-6 if dry else 1e-4, atol=0)
    np.testing.assert_allclose(km(ip["q_liq"]), sd("ql_sdm"), rtol=1e-6)
    np.testing.assert_allclose(km(ip["T"]), sd("T_sdm"), rtol=1e-2)
    np.testing.assert_allclose(km(ip["p"]), sd("P_sdm"), rtol=1e-2)
    np.testing.assert_allclose(km(ip["ρ"]), sd("rho_sdm"), rtol=1e-2)
    np.testing.assert_allclose(km(ip["θ_dry"]), sd("thetad_sdm"), rtol=4e-6)


def test_parameter_overwrites():
    """reference test/unit_tests/k1d_unit_test.jl:3-10"""
    td = PR.override_toml_dict(FT=np.float64)
    tp = PR.create_thermodynamics_parameters(td)
    assert tp.R_d == 8.314462618 / 0.02896998
    assert tp.R_v == 8.314462618 / 0.018015
    assert tp.MSLP == 100000.0


def test_box_initial_condition_partitions_water():
    td = PR.override_toml_dict(FT=np.float64)
    tp = PR.create_thermodynamics_parameters(td)
    ip = IC.initial_condition_0d(np.float64, tp, 1e-3, 1e8, 2.0, 1.22)
    assert ip["ρq_liq"] + ip["ρq_rai"] == pytest.approx(ip["ρq_tot"])
    assert ip["N_liq"] + ip["N_rai"] == pytest.approx(1e8)
    assert 0 <= ip["ρq_rai"] <= ip["ρq_tot"]
    assert ip["T"] == 300.0 and ip["N_aer"] == 0.0
