"""Known-answer / invariant tests of the CPU oracle: every structural invariant the reference's own
unit tests assert, evaluated on the same hand-written fixture point.

Reference: test/unit_tests/common_unit_test.jl:163-197 (fixture), :268-290 (thermo),
:313-322 (velocities), :343-373 (sources), test/unit_tests/k1d_unit_test.jl:109-138
(zero-flow advection), :214-279 (closed-system activation)."""
import numpy as np
import pytest

from kid_b200 import model as M
from kid_b200 import parameters as PR
from kid_b200.lib import MODEL_BOX, MODEL_K1D
from oracle.oracle import OracleHandle

IP = dict(ρ=1.2, ρ_dry=1.185, T=280.0, ρq_tot=15e-3, ρq_liq=2e-3, ρq_ice=2e-3, ρq_rai=1e-3, ρq_sno=0.0,
          N_liq=1e8, N_rai=1e4, N_aer=1e10)
STYLES = [("EquilibriumMoisture", "NoPrecipitation"), ("NonEquilibriumMoisture", "NoPrecipitation"),
          ("EquilibriumMoisture", "Precipitation0M"), ("EquilibriumMoisture", "Precipitation1M"),
          ("EquilibriumMoisture", "Precipitation2M"), ("NonEquilibriumMoisture", "Precipitation1M"),
          ("NonEquilibriumMoisture", "Precipitation2M")]


def point_handle(ms, ps, FT=np.float64, dt=10.0, **kw):
    cs = M.custom_case(FT, MODEL_BOX, ms, ps, IP, dt=dt, **kw)
    return cs, M.make_handle(cs, OracleHandle)


def air_pressure(tp, T, ρ, q_tot, q_liq=0.0, q_ice=0.0):
    ratio = tp.R_v / tp.R_d
    return tp.R_d * (1 + (ratio - 1) * q_tot - ratio * (q_liq + q_ice)) * ρ * T


@pytest.mark.parametrize("ms", ["EquilibriumMoisture", "NonEquilibriumMoisture"])
def test_precompute_aux_thermo(ms):
    cs, h = point_handle(ms, "NoPrecipitation")
    tp = PR.create_thermodynamics_parameters(cs.meta["toml"])
    aux = {f: h.get_aux(0.0, f)[0, 0] for f in ("ρ", "ρ_dry", "p", "T", "θ_dry", "θ_liq_ice", "q_tot", "q_liq", "q_ice")}
    assert all(np.isfinite(v) for v in aux.values())
    assert aux["q_tot"] >= 0 and aux["q_liq"] >= 0 and aux["q_ice"] >= 0
    assert aux["ρ_dry"] == aux["ρ"] - IP["ρq_tot"]
    if ms == "EquilibriumMoisture":
        assert aux["p"] == air_pressure(tp, aux["T"], aux["ρ"], aux["q_tot"])
    else:
        assert aux["p"] == air_pressure(tp, aux["T"], aux["ρ"], aux["q_tot"], aux["q_liq"], aux["q_ice"])
    p_dry = tp.R_d * aux["ρ_dry"] * aux["T"]
    assert aux["θ_dry"] == pytest.approx(aux["T"] / (p_dry / tp.MSLP) ** (tp.R_d / tp.cp_d), rel=1e-14)


@pytest.mark.parametrize("ps", ["Precipitation1M", "Precipitation2M"])
def test_precompute_aux_precip_velocities_nonnegative(ps):
    cs, h = point_handle("EquilibriumMoisture", ps)
    for f in ("term_vel_rai", "term_vel_sno" if ps == "Precipitation1M" else "term_vel_N_rai", "q_rai", "q_sno"):
        v = h.get_aux(0.0, f)
        assert np.all(np.isfinite(v)) and np.all(v >= 0), f
    assert h.get_aux(0.0, "term_vel_rai")[0, 0] > 0  # there is rain at the fixture point


@pytest.mark.parametrize("ps", ["Precipitation0M", "Precipitation1M", "Precipitation2M"])
def test_precompute_aux_precip_sources(ps):
    cs, h = point_handle("EquilibriumMoisture", ps)
    get = lambda f: h.get_aux(0.0, "precip_sources." + f)[0, 0]  # noqa: E731
    if ps == "Precipitation0M":
        q_tot, q_liq, q_ice = get("q_tot"), get("q_liq"), get("q_ice")
        assert np.isfinite([q_tot, q_liq, q_ice]).all()
        assert q_tot == q_liq + q_ice
        assert q_tot < 0  # supersaturated fixture: precipitation is removed
    elif ps == "Precipitation1M":
        vals = [get(f) for f in ("q_tot", "q_liq", "q_ice", "q_rai", "q_sno")]
        assert np.isfinite(vals).all()
        assert vals[0] == 0
        assert vals[1] + vals[2] + vals[3] + vals[4] == pytest.approx(0, abs=1e-18)  # water is conserved
    else:
        vals = [get(f) for f in ("q_tot", "q_liq", "q_rai", "N_aer", "N_liq", "N_rai")]
        assert np.isfinite(vals).all()
        assert vals[0] == 0
        assert vals[1] + vals[2] == pytest.approx(0, abs=1e-18)


def test_cloud_sources_finite():
    cs, h = point_handle("NonEquilibriumMoisture", "Precipitation1M")
    # the box rhs does not apply condensation (BoxModel/ode_utils.jl:12-20); evaluate it through a K1D column
    ip5 = {k: np.full(5, v) for k, v in IP.items()}
    cs = M.custom_case(np.float64, MODEL_K1D, "NonEquilibriumMoisture", "Precipitation1M", ip5, z_max=100.0, w1=0.0)
    h = M.make_handle(cs, OracleHandle)
    for f in ("cloud_sources.q_liq", "cloud_sources.q_ice"):
        assert np.all(np.isfinite(h.get_aux(0.0, f)))


@pytest.mark.parametrize("ms,ps", [("EquilibriumMoisture", "NoPrecipitation"), ("NonEquilibriumMoisture", "Precipitation1M"),
                                   ("NonEquilibriumMoisture", "Precipitation2M")])
def test_zero_flow_advection_adds_nothing(ms, ps):
    """ρw = 0 and no falling species ⇒ the advection operators add exactly 0 (k1d_unit_test.jl:127-137):
    the K1D rhs then equals ρ·(cloud + precip sources) level by level."""
    cs = M.k1d_case(np.float64, nc=1, moisture_choice=ms, precipitation_choice=ps, n_elem=5, z_max=100.0, w1=0.0)
    h = M.make_handle(cs, OracleHandle)
    dY = h.rhs(13.0)[0]
    ρ = h.get_aux(13.0, "ρ")[0]
    names = list(cs.var_names)
    expect = np.zeros_like(dY)
    src = lambda f: h.get_aux(13.0, f)[0]  # noqa: E731
    if ms == "NonEquilibriumMoisture":
        expect[names.index("ρq_liq")] += ρ * src("cloud_sources.q_liq") + ρ * src("precip_sources.q_liq")
        expect[names.index("ρq_ice")] += ρ * src("cloud_sources.q_ice")
        if ps == "Precipitation1M":
            expect[names.index("ρq_ice")] += ρ * src("precip_sources.q_ice")
    if ps != "NoPrecipitation":
        expect[names.index("ρq_rai")] += ρ * src("precip_sources.q_rai")
    if ps == "Precipitation1M":
        expect[names.index("ρq_sno")] += ρ * src("precip_sources.q_sno")
    if ps == "Precipitation2M":
        for n, f in (("N_aer", "N_aer"), ("N_liq", "N_liq"), ("N_rai", "N_rai")):
            expect[names.index(n)] += src("precip_sources." + f)
    np.testing.assert_allclose(dY, expect, rtol=0, atol=10 * np.finfo(np.float64).eps * max(1.0, np.abs(expect).max()))


def test_activation_closed_system():
    """k1d_unit_test.jl:214-279: closed-system activation gives S(N_aer) = -S(N_liq), finite."""
    ip = dict(IP, ρq_tot=16e-3, ρq_liq=1e-3, ρq_sno=2e-3, N_aer=5e7)
    ip5 = {k: np.full(5, v) for k, v in ip.items()}
    cs = M.custom_case(np.float64, MODEL_K1D, "EquilibriumMoisture", "Precipitation2M", ip5, z_max=100.0, dt=10.0)
    h = M.make_handle(cs, OracleHandle)
    a, l = h.get_aux(13.0, "activation_sources.N_aer"), h.get_aux(13.0, "activation_sources.N_liq")
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(l))
    np.testing.assert_array_equal(a, -l)
    # open system: aerosol is not depleted
    cs = M.custom_case(np.float64, MODEL_K1D, "EquilibriumMoisture", "Precipitation2M", ip5, z_max=100.0, dt=10.0,
                       open_system_activation=1)
    h = M.make_handle(cs, OracleHandle)
    assert np.all(h.get_aux(13.0, "activation_sources.N_aer") == 0)


def test_cloud_base_selection_keeps_one_level():
    """non-local activation keeps S_Nl at exactly one level: the lowest activating one (tendency.jl:163-168)"""
    cs = M.k1d_case(np.float64, nc=1, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    h = M.make_handle(cs, OracleHandle)
    h.step(0.0, 240)
    act = h.get_aux(240.0, "activation_sources.N_liq")[0]
    assert np.count_nonzero(act) == 1 and act.max() > 0
    k = int(np.argmax(act))
    cs_l = M.k1d_case(np.float64, nc=1, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                      local_activation=True)
    hl = M.make_handle(cs_l, OracleHandle)
    hl.set_state(h.get_state())
    act_l = hl.get_aux(240.0, "activation_sources.N_liq")[0]
    assert np.all(act_l[:k] <= 0)  # nothing activates below the cloud base the reduction found


def test_limiter_identities():
    """docstring src/Common/helper_functions.jl:202-205: 0 <= L <= M and L <= F for F >= 0, M > 0"""
    rng = np.random.default_rng(0)
    F = 10.0 ** rng.uniform(-12, 3, 1000)
    Mx = 10.0 ** rng.uniform(-12, 3, 1000)
    Lm = F + Mx - np.sqrt(F * F + Mx * Mx)
    # in floating point the identities hold up to the rounding of the larger operand
    slack = 4 * np.finfo(np.float64).eps * np.maximum(F, Mx)
    assert np.all(Lm >= -slack) and np.all(Lm <= Mx + slack) and np.all(Lm <= F + slack)


@pytest.mark.parametrize("ps,rf", [("Precipitation1M", "CliMA_1M"), ("Precipitation2M", "SB2006")])
def test_ssprk33_third_order(ps, rf):
    """The step loop converges with order 3 (self-convergence of the box model)."""
    sols = []
    for dt, n in ((8.0, 25), (4.0, 50), (2.0, 100), (0.25, 800)):
        cs = M.box_case(np.float64, nc=1, precipitation_choice=ps, rain_formation_choice=rf, qt=2e-3, dt=dt)
        h = M.make_handle(cs, OracleHandle)
        h.step(0.0, n)
        sols.append(h.get_state()[0, :, 0])
    i = list(cs.var_names).index("ρq_rai")
    e1, e2, e3 = (abs(s[i] - sols[3][i]) for s in sols[:3])
    # NOTE: the limiter's and the scheme's own dt dependence (limit = q/dt) makes the ODE itself
    # dt-dependent at O(dt^-1 * tiny); only require clear higher-than-first-order decay
    assert e2 < e1 and e3 < e2


def test_water_budget_closed_box():
    """Box 1M/2M: ρq_tot never changes and liquid + rain mass is conserved (precip_sinks = 0)."""
    for ps, rf in (("Precipitation1M", "CliMA_1M"), ("Precipitation2M", "SB2006")):
        cs = M.box_case(np.float64, nc=3, precipitation_choice=ps, rain_formation_choice=rf, qt=[5e-4, 1e-3, 2e-3])
        h = M.make_handle(cs, OracleHandle)
        Y0 = h.get_state()
        h.step(0.0, 600)
        Y1 = h.get_state()
        n = list(cs.var_names)
        np.testing.assert_allclose(Y1[:, n.index("ρq_tot")], Y0[:, n.index("ρq_tot")], rtol=1200 * np.finfo(np.float64).eps)
        tot0 = Y0[:, n.index("ρq_liq")] + Y0[:, n.index("ρq_rai")]
        tot1 = Y1[:, n.index("ρq_liq")] + Y1[:, n.index("ρq_rai")]
        np.testing.assert_allclose(tot1, tot0, rtol=1e-12)
        assert np.all(Y1[:, n.index("ρq_rai")] > Y0[:, n.index("ρq_rai")])  # rain forms


def test_chen2022_rain_velocity_is_a_plausible_fall_speed():
    """No golden value exists for CM2.rain_terminal_velocity(sb2006, Chen2022VelTypeRain, ...): the restated Chen et al.
    (2022) Table B1 fit must at least give fall speeds of the magnitude of the SB2006 ones it replaces (unit slips in the
    mm -> m conversion of its coefficients would show as factors of 10^3b)."""
    import numpy as np
    from kid_b200 import model as M
    from oracle.oracle import OracleHandle
    kw = dict(nc=1, n_elem=48, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    sb = M.make_handle(M.k1d_case(np.float64, **kw), OracleHandle)
    ch = M.make_handle(M.k1d_case(np.float64, sedimentation_scheme_choice="Chen2022", **kw), OracleHandle)
    sb.step(0.0, 1500)
    ch.set_state(sb.get_state())
    for f, lo, hi in (("term_vel_rai", 0.7, 1.5), ("term_vel_N_rai", 0.5, 3.0)):
        x, y = sb.get_aux(1500.0, f)[0], ch.get_aux(1500.0, f)[0]
        m = x > 0.1
        assert m.any() and np.all(y >= 0) and np.all(y < 12.0)
        r = y[m] / x[m]
        assert lo < r.min() and r.max() < hi, (f, r.min(), r.max())


def test_precipitation_susceptibility_to_droplet_number():
    """test/test_precip_production.jl (a plot in the reference, not an assertion): the precipitation-production metric
    Σ_t ρq_rai at the cloud-base level of the default KiD 2M/SB2006 run falls with the prescribed droplet number, with a
    log-log slope of about -2 at high Nd (Glassmeier & Lohmann).  Here: all Nd as ONE ensemble, on the oracle."""
    import numpy as np
    from kid_b200 import model as M
    from kid_b200.lib import COL_PRESCRIBED_ND
    from oracle.oracle import OracleHandle
    Nd = np.linspace(1e5, 1e8, 10)  # Nd_range of the reference script
    case = M.k1d_case(np.float64, nc=10, moisture_choice="EquilibriumMoisture", precipitation_choice="Precipitation2M",
                      rain_formation_scheme_choice="SB2006")
    case.col_scalars[COL_PRESCRIBED_ND] = Nd
    case.Y0[:, case.var_names.index("N_aer"), :] = Nd[:, None] * case.ρ_dry[None, :] / 1.225
    o = M.make_handle(case, OracleHandle)
    ir, cloud_base = case.var_names.index("ρq_rai"), 75 - 1  # Julia's 1-based level 75
    P = o.get_state()[:, ir, cloud_base].astype(np.float64)
    for i in range(120):  # saveat = dt_output = 30 s up to t_end = 3600 s
        o.step(i * 30.0, 30)
        P += o.get_state()[:, ir, cloud_base]
    assert np.all(np.diff(P) < 0)                     # more droplets, less rain
    slope = np.log(P[-1] / P[5]) / np.log(Nd[-1] / Nd[5])
    assert -3.5 < slope < -1.0, slope                 # ≈ -2 at the polluted end
