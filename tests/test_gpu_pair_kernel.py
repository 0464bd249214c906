"""k1d_pair_kernel (the packed two-levels-per-thread Precipitation2M step kernel, the headline path) against the scalar tile
kernel and the CPU oracle.

The pair kernel serves K1DModel / col_sed, Precipitation2M (SB2006 sedimentation), Float32, 128 levels, one shared parameter
set, more than 16 columns; `cfg.reserved[3] = 1` forces the scalar tile kernel for the same handle configuration.

Tolerances.  From an identical state the two kernels evaluate the same formulas with different instruction selection
(packed FFMA2 / flux-form stencil / staged grid-point constants), so ONE step may differ by a few ulp of the state:
the difference is bounded by 2e-3 of the largest one-step increment of the variable plus 4 ulp of its largest value.  Over many steps the reference
formulation's knife-edges (find_cloud_base!: the sign of N_act - N_liq at the level where N_liq has just reached N_act)
amplify ulp differences in single columns, exactly as between either kernel and the oracle; there the Float32 tolerances of
tests/test_gpu_parity.py apply (q 5e-3, N 5e-2 up to 100 steps)."""
import numpy as np
import pytest

from kid_b200 import model as M
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
from parity import errors

pytestmark = pytest.mark.gpu

OPTS = dict(n_elem=128, precipitation_choice="Precipitation2M")


def ensemble(nc, no_pair, moisture="NonEquilibriumMoisture", **kw):
    cs = M.k1d_ensemble_case(np.float32, nc=nc, moisture_choice=moisture, order="sorted", ic="device", **OPTS, **kw)
    cs.cfg.reserved[3] = no_pair
    return cs


def one_step_bound(Yp, Ys, Y0, names, frac):
    for i, v in enumerate(names):
        inc = np.abs(Ys[:, i].astype(np.float64) - Y0[:, i]).max()
        d = np.abs(Yp[:, i].astype(np.float64) - Ys[:, i]).max()
        ulp = float(np.spacing(np.float32(np.abs(Ys[:, i]).max())))  # a state that hardly moves still rounds differently
        assert d <= frac * inc + 4 * ulp, (v, d, inc, ulp)


@pytest.mark.parametrize("moisture", ["NonEquilibriumMoisture", "EquilibriumMoisture"])
@pytest.mark.parametrize("t_dev", [60, 400])
def test_one_step_matches_the_scalar_kernel(moisture, t_dev):
    nc = 256
    hs = M.make_handle(ensemble(nc, 1, moisture), Handle)
    hp = M.make_handle(ensemble(nc, 0, moisture), Handle)
    hs.set_steps_per_launch(50)
    hs.step(0.0, t_dev)
    Y0 = hs.get_state()
    hp.set_state(Y0)
    hs.step(float(t_dev), 1)
    hp.step(float(t_dev), 1)
    assert hp.launch_count >= 1
    cs = ensemble(nc, 0, moisture)
    # Equilibrium + 2M: activation is gated by the sign of a supersaturation that is 0 up to rounding in cloud (see
    # tests/test_gpu_parity.py): the number fields may differ by one activation event, the masses may not
    names = cs.var_names
    Yp, Ys = hp.get_state(), hs.get_state()
    if moisture == "EquilibriumMoisture":
        keep = [i for i, v in enumerate(names) if not v.startswith("N_")]
        one_step_bound(Yp[:, keep], Ys[:, keep], Y0[:, keep], [names[i] for i in keep], 2e-3)
    else:
        one_step_bound(Yp, Ys, Y0, names, 2e-3)


def test_pair_kernel_against_the_oracle_100_steps():
    nc = 64
    cs = ensemble(nc, 0)
    g = M.make_handle(cs, Handle)
    o = M.make_handle(M.materialise(cs, g), OracleHandle)
    g.set_steps_per_launch(25)
    g.step(0.0, 100)
    o.step(0.0, 100)
    Yg, Yo = g.get_state(), o.get_state()
    assert np.isfinite(Yg).all()
    for v, e in errors(Yg, Yo, cs.var_names).items():
        assert e <= (5e-2 if v.startswith("N_") else 5e-3), (v, e)


def test_splitting_a_run_into_launches_is_bitwise_invariant():
    """The pseudo-stage at the start of a launch re-evaluates exactly what the previous launch's last stage would have
    handed over (terminal velocities, cloud base): 30 steps in one launch == 30 launches of one step."""
    nc = 64
    a = M.make_handle(ensemble(nc, 0), Handle)
    b = M.make_handle(ensemble(nc, 0), Handle)
    a.step(0.0, 30)
    for t in range(30):
        b.step(float(t), 1)
    assert np.array_equal(a.get_state(), b.get_state())


@pytest.mark.parametrize("switches", [dict(local_activation=True), dict(open_system_activation=True),
                                      dict(qtot_flux_correction=True), dict(rain_formation_scheme_choice="SB2006NL")])
def test_option_variants_match_the_scalar_kernel(switches):
    nc = 64
    hs = M.make_handle(ensemble(nc, 1, **switches), Handle)
    hp = M.make_handle(ensemble(nc, 0, **switches), Handle)
    hs.step(0.0, 300)
    Y0 = hs.get_state()
    hp.set_state(Y0)
    hs.step(300.0, 1)
    hp.step(300.0, 1)
    one_step_bound(hp.get_state(), hs.get_state(), Y0, ensemble(nc, 0, **switches).var_names, 2e-3)


def test_col_sed_model_on_the_pair_kernel():
    """make_rhs_function_col_sed (K1DModel/ode_utils.jl:56-73): collisions + sedimentation only, 128 levels, 32 columns."""
    nc = 32
    kw = dict(nc=nc, n_elem=128, qt=np.linspace(0.8e-3, 1.6e-3, nc), prescribed_Nd=np.geomspace(5e7, 3e8, nc))
    cs_s = M.col_sed_case(np.float32, **kw)
    cs_s.cfg.reserved[3] = 1
    cs_p = M.col_sed_case(np.float32, **kw)
    hs, hp = M.make_handle(cs_s, Handle), M.make_handle(cs_p, Handle)
    o = M.make_handle(cs_p, OracleHandle)
    for h in (hs, hp, o):
        h.step(0.0, 60)
    Yp, Ys, Yo = hp.get_state(), hs.get_state(), o.get_state()
    assert np.isfinite(Yp).all()
    ep, es = errors(Yp, Yo, cs_p.var_names), errors(Ys, Yo, cs_p.var_names)
    for v in cs_p.var_names:
        assert ep[v] <= (5e-2 if v.startswith("N_") else 5e-3), (v, ep[v], es[v])
