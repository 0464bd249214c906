"""GPU parity tests of the K2DModel path: single handle vs oracle, x-decomposed ring vs single handle."""
import numpy as np
import pytest

from kid_b200 import k2d
from kid_b200 import model as M
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
from parity import errors, rhs_errors

pytestmark = pytest.mark.gpu

STYLES = [("EquilibriumMoisture", "NoPrecipitation"), ("EquilibriumMoisture", "Precipitation1M"),
          ("NonEquilibriumMoisture", "Precipitation0M"), ("NonEquilibriumMoisture", "Precipitation1M"),
          ("NonEquilibriumMoisture", "Precipitation2M")]


def tol(FT, v, long=False):
    if FT == np.float64:
        return 1e-7 if long else 1e-8
    return 5e-2 if v.startswith("N_") else (1e-2 if long else 5e-3)


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("ms,ps", STYLES)
def test_k2d_parity_vs_oracle(FT, ms, ps):
    """5 x 4-node elements, 24 levels, strong and short updraft so that cloud, rain and the horizontal flow are all
    active inside 300 steps (run_kinematic2d_simulations.jl defaults otherwise)."""
    cs = k2d.k2d_case(FT, nx=5, nz=24, npoly=3, moisture_choice=ms, precipitation_choice=ps, t1=400.0, w1=2.5,
                      domain_height=2400.0, domain_width=2400.0)
    g, o = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    e = rhs_errors(g.rhs(30.0), o.rhs(30.0), cs.Y0, cs.var_names, cs.cfg.dt, FT)
    g, o = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)  # rhs() counts as a call for the stale aux
    for v, err in e.items():
        assert err <= (1e-8 if FT == np.float64 else 2e-2), (v, err, e)
    g.step(0.0, 300)
    o.step(0.0, 300)
    Yg, Yo = g.get_state(), o.get_state()
    assert np.all(np.isfinite(Yg))
    e = errors(Yg, Yo, cs.var_names)
    for v, err in e.items():
        assert err <= tol(FT, v, long=True), (v, err, e)
    Nq = 4
    assert np.array_equal(Yg[Nq - 1::Nq][:-1], Yg[Nq::Nq]) and np.array_equal(Yg[-1], Yg[0])  # DSS keeps continuity
    if ps != "NoPrecipitation" or ms == "NonEquilibriumMoisture":
        assert np.abs(Yo - Yo[0:1]).max() > 0


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("parts", [2, 3])
def test_k2d_decomposed_ring_matches_single_domain_bitwise(FT, parts):
    """x-decomposition by whole elements with the edge exchange (kid_k2d_stage_begin / exchange / stage_end) gives
    the same bits as one handle owning the periodic domain."""
    kw = dict(nx=6, nz=16, npoly=4, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
              t1=300.0, w1=2.5, domain_height=2400.0, domain_width=2400.0)
    whole = M.make_handle(k2d.k2d_case(FT, **kw), Handle)
    ring = k2d.LocalRing([M.make_handle(k2d.k2d_case(FT, rank=r, world=parts, **kw), Handle) for r in range(parts)])
    whole.step(0.0, 90)
    ring.step(0.0, 90)
    assert np.array_equal(ring.get_state(), whole.get_state())


def test_k2d_baseline_size_domain_properties():
    """BASELINE config 5: 4096 horizontal nodes x 256 levels, Precipitation2M (1024 elements of 4 GLL nodes):
    finite, continuous at every shared node, periodic, and x-symmetric fields stay bounded."""
    cs = k2d.k2d_case(np.float32, nx=1024, nz=256, npoly=3, moisture_choice="NonEquilibriumMoisture",
                      precipitation_choice="Precipitation2M", domain_width=48000.0, domain_height=3000.0)
    assert cs.cfg.nc == 4096
    g = M.make_handle(cs, Handle)
    g.step(0.0, 30)
    Y = g.get_state()
    assert np.all(np.isfinite(Y))
    assert np.array_equal(Y[3::4][:-1], Y[4::4]) and np.array_equal(Y[-1], Y[0])
    n = list(cs.var_names)
    assert Y[:, n.index("ρq_tot")].min() > 0
