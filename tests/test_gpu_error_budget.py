"""How much of the GPU-Float32 error is the GPU's?  The Float32 tolerances of tests/test_gpu_parity.py are loose because
the reference formulation itself is ill-conditioned in Float32 (limiter cancellation, threshold branches).  Here the
GPU-Float32 result and the oracle-Float32 result are both measured against the oracle in Float64 — the truth of the
formulation — and the GPU error must stay within 2x the oracle's own Float32 error (plus a floor of a few ulp): the
MUFU transcendentals, rcp-based division and FTZ of the device build (-use_fast_math) do not add error beyond the noise
any Float32 evaluation of the reference has.  BASELINE configs 3 (NonEq + 1M) and 4 (NonEq + 2M) in small.

Also here: the pointwise relative metric (absolute floor 1e-3 of the field maximum) beside the field-maximum metric, and
conditioning-free assertions for Equilibrium + 2M, whose activation gate is rounding noise in the reference
(tests/test_gpu_parity.py): column-integrated number and cloud-base level."""
import numpy as np
import pytest

from kid_b200 import model as M
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
from parity import cloud_base_levels, column_integrated_number_error, errors, pointwise_errors

pytestmark = pytest.mark.gpu


def three_runs(precip, nsteps, moisture="NonEquilibriumMoisture", nc=32, n_elem=128):
    kw = dict(nc=nc, n_elem=n_elem, n_profiles=4, order="sorted", moisture_choice=moisture, precipitation_choice=precip)
    c64 = M.k1d_ensemble_case(np.float64, **kw)
    c32 = M.k1d_ensemble_case(np.float32, **kw)
    truth = M.make_handle(c64, OracleHandle)
    o32 = M.make_handle(c32, OracleHandle)
    g32 = M.make_handle(c32, Handle)
    for h in (truth, o32, g32):
        h.step(0.0, nsteps)
    return c32.var_names, truth.get_state(), o32.get_state(), g32.get_state()


@pytest.mark.parametrize("precip,nsteps", [("Precipitation1M", 300), ("Precipitation2M", 300)])
def test_gpu_float32_error_is_within_twice_the_oracles_float32_error(precip, nsteps):
    names, Yt, Yo, Yg = three_runs(precip, nsteps)
    assert np.isfinite(Yg).all()
    eo, eg = errors(Yo, Yt, names), errors(Yg, Yt, names)
    for v in names:
        if v.startswith("N_"):
            continue  # knife-edge of find_cloud_base!: judged by the column-integrated number below
        assert eg[v] <= 2.0 * eo[v] + 2e-5, (v, eg[v], eo[v], eg, eo)
    if precip == "Precipitation2M":
        co, cg = column_integrated_number_error(Yo, Yt, names), column_integrated_number_error(Yg, Yt, names)
        assert cg <= 2.0 * co + 1e-5, (cg, co)


def test_pointwise_relative_error_100_steps():
    """NonEq + 2M, Float32, GPU against the Float32 oracle, pointwise with a floor of 1e-3 of the field maximum."""
    names, _, Yo, Yg = three_runs("Precipitation2M", 100)
    pw = pointwise_errors(Yg, Yo, names, floor=1e-3)
    for v in names:
        if v.startswith("N_"):
            continue
        assert pw[v] <= 5e-2, (v, pw)


@pytest.mark.parametrize("FT", [np.float32, np.float64])
def test_equilibrium_2m_conditioning_free_observables(FT):
    """Eq + 2M: the supersaturation that gates activation is 0 up to rounding inside the cloud, so WHICH level activates is
    noise in the reference itself.  What hardly hinges on it: the column-integrated number N_liq + N_rai + N_aer (the
    activated number depends weakly on the level it is taken from: measured 2e-4 in Float64, 1.4e-3 in Float32 after 240
    steps, against 1e-1 .. 1 for the fields themselves) and the cloud-base level (within one level)."""
    kw = dict(nc=16, n_elem=64, n_profiles=2, order="sorted", moisture_choice="EquilibriumMoisture",
              precipitation_choice="Precipitation2M")
    cs = M.k1d_ensemble_case(FT, **kw)
    g, o = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    g.step(0.0, 240)
    o.step(0.0, 240)
    Yg, Yo = g.get_state(), o.get_state()
    assert np.isfinite(Yg).all()
    assert column_integrated_number_error(Yg, Yo, cs.var_names) <= (1e-3 if FT == np.float64 else 5e-3)
    bg, bo = cloud_base_levels(Yg, cs.var_names), cloud_base_levels(Yo, cs.var_names)
    assert np.all(np.abs(bg - bo) <= 1), (bg, bo)
