"""GPU parity tests: libkid_cuda.so (through its C ABI) against the CPU oracle on identical inputs.

Tolerances (relative, metric in tests/parity.py), stated per float type and per check:

  RHS / AUX from an identical developed state
      Float64  1e-8      (observed <= 4e-9; differences come from CUDA's pow/exp/cbrt/erf vs glibc's)
      Float32  see RHS_TOL_F32: the reference's limiter F + M - sqrt(F^2 + M^2) cancels
               catastrophically in Float32 when F << M (the result is quantised to ulp(M)); a
               1-ulp difference in M between two libm implementations therefore moves a source by
               ulp(M).  Mass tendencies are compared to 5e-3, number tendencies to 2e-2.
  STATE after K steps from an identical state
      Float64  1e-8 (K <= 60), 1e-6 (K = 600), except Equilibrium + 2M: 1e-1 — there the reference
               feeds activation with S = p_v/p_vs - 1 where q_vap == q_vs exactly in cloud, so the
               sign of S is rounding noise and the cloud-base level it selects is ill-conditioned
               in the reference itself.
      Float32  q variables 5e-3 (K <= 100) / 1e-2 (K = 600), N variables 5e-2 (same limiter quantisation,
               accumulated; threshold branches of the reference such as N_liq/N_aer > 1e-6 can flip on a
               1-ulp difference; the device evaluates pow/exp/log with the MUFU approximations).
               Equilibrium + 2M in Float32: q 5e-1, number fields only checked for boundedness."""
import numpy as np
import pytest

from kid_b200 import lib
from kid_b200 import model as M
from kid_b200.lib import AUX_FIELDS, Handle, MODEL_K1D_COL_SED
from oracle.oracle import OracleHandle
from parity import errors, field_error, rhs_errors

pytestmark = pytest.mark.gpu

STYLES = [("EquilibriumMoisture", "NoPrecipitation"), ("EquilibriumMoisture", "Precipitation0M"),
          ("EquilibriumMoisture", "Precipitation1M"), ("EquilibriumMoisture", "Precipitation2M"),
          ("NonEquilibriumMoisture", "NoPrecipitation"), ("NonEquilibriumMoisture", "Precipitation0M"),
          ("NonEquilibriumMoisture", "Precipitation1M"), ("NonEquilibriumMoisture", "Precipitation2M")]
FTS = [np.float64, np.float32]


def ill_conditioned(ms, ps):
    return ms == "EquilibriumMoisture" and ps == "Precipitation2M"


def state_tol(FT, var, nsteps, ms, ps):
    if ill_conditioned(ms, ps) and nsteps > 1:
        if FT == np.float64:
            return 1e-1
        return 1.0 if var.startswith("N_") else 5e-1  # Float32: only boundedness of the number fields is checked
    if FT == np.float64:
        return 1e-8 if nsteps <= 60 else 1e-6
    if nsteps > 100:
        return 5e-2 if var.startswith("N_") else 1e-2
    return 5e-2 if var.startswith("N_") else 5e-3


def rhs_tol(FT, var):
    if FT == np.float64:
        return 1e-8
    return 2e-2 if var.startswith("N_") else 5e-3


def developed_state(FT, t_dev=900, **opts):
    """Run the ORACLE to a developed state (cloud + rain below cloud base) and return (case, Y, t)."""
    cs = M.k1d_case(FT, nc=2, n_elem=48, **opts)
    o = M.make_handle(cs, OracleHandle)
    o.step(0.0, t_dev)
    return cs, o.get_state(), float(t_dev)


def pair(cs, Y=None):
    g, o = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    if Y is not None:
        g.set_state(Y)
        o.set_state(Y)
    return g, o


@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("ms,ps", STYLES)
def test_rhs_and_aux_parity_developed_state(FT, ms, ps):
    # t_dev inside the updraft window (t1 = 600 s) so that advection, activation, condensation and
    # the precipitation sources are all active in the single rhs evaluation
    cs, Y, t = developed_state(FT, 480, moisture_choice=ms, precipitation_choice=ps)
    g, o = pair(cs, Y)
    e = rhs_errors(g.rhs(t), o.rhs(t), Y, cs.var_names, cs.cfg.dt, FT)
    # Equilibrium + 2M: the level activation picks depends on the sign of a supersaturation that is pure
    # rounding noise in cloud (module docstring), so the activation-fed tendencies are not compared pointwise
    skip = {"N_liq", "N_aer"} if ill_conditioned(ms, ps) else set()
    for v, err in e.items():
        assert v in skip or err <= rhs_tol(FT, v), (v, err, e)
    aux_tol = 1e-8 if FT == np.float64 else 2e-2
    for f in AUX_FIELDS:
        if ill_conditioned(ms, ps) and f.startswith("activation_sources"):
            continue
        a, b = g.get_aux(t, f), o.get_aux(t, f)
        floor = 0.0
        if f.startswith(("precip_sources", "cloud_sources", "activation_sources")):
            # sources are differences of limited rates: measure them against the largest source of their kind
            kind = "N_" if ".N_" in f else "q_"
            floor = max(np.abs(o.get_aux(t, x)).max() for x in AUX_FIELDS
                        if x.split(".")[0] == f.split(".")[0] and (("." + kind) in x))
        assert field_error(a, b, floor * 1e-3) <= aux_tol, f


@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("ms,ps", STYLES)
def test_step_parity_from_developed_state(FT, ms, ps):
    """60 steps with rain falling and evaporating (t > t1: sedimentation, evaporation, breakup active)."""
    cs, Y, t = developed_state(FT, 900, moisture_choice=ms, precipitation_choice=ps)
    g, o = pair(cs, Y)
    g.step(t, 60)
    o.step(t, 60)
    Yg = g.get_state()
    assert np.all(np.isfinite(Yg))
    e = errors(Yg, o.get_state(), cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 60, ms, ps), (v, err, e)


@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("ms,ps", STYLES)
def test_step_parity_from_initial_condition(FT, ms, ps):
    """600 steps from the ShipwayHill2012 initial condition: the whole updraft phase."""
    cs = M.k1d_case(FT, nc=3, n_elem=64, moisture_choice=ms, precipitation_choice=ps)
    g, o = pair(cs)
    g.step(0.0, 600)
    o.step(0.0, 600)
    e = errors(g.get_state(), o.get_state(), cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 600, ms, ps), (v, err, e)


@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("opts", [
    dict(precipitation_choice="Precipitation1M", rain_formation_scheme_choice="KK2000"),
    dict(precipitation_choice="Precipitation1M", rain_formation_scheme_choice="B1994"),
    dict(precipitation_choice="Precipitation1M", rain_formation_scheme_choice="TC1980"),
    dict(precipitation_choice="Precipitation1M", rain_formation_scheme_choice="LD2004"),
    dict(precipitation_choice="Precipitation1M", rain_formation_scheme_choice="VarTimeScaleAcnv"),
    dict(precipitation_choice="Precipitation2M", rain_formation_scheme_choice="SB2006NL"),
    dict(precipitation_choice="Precipitation2M", local_activation=True),
    dict(precipitation_choice="Precipitation2M", open_system_activation=True),
    dict(precipitation_choice="Precipitation2M", qtot_flux_correction=True),
    dict(precipitation_choice="Precipitation1M", precip_sinks=False),
    dict(precipitation_choice="Precipitation2M", precip_sources=False),
], ids=lambda o: "-".join(f"{k[:6]}={v}" for k, v in o.items() if k != "precipitation_choice"))
def test_option_variants(FT, opts):
    """KiD_driver options that select other kernel branches (parse_commandline.jl:10-174)."""
    ms, ps = "NonEquilibriumMoisture", opts["precipitation_choice"]
    cs, Y, t = developed_state(FT, 700, moisture_choice=ms, **opts)
    g, o = pair(cs, Y)
    e = rhs_errors(g.rhs(t), o.rhs(t), Y, cs.var_names, cs.cfg.dt, FT)
    for v, err in e.items():
        assert err <= rhs_tol(FT, v), (v, err, e)
    g.step(t, 30)
    o.step(t, 30)
    e = errors(g.get_state(), o.get_state(), cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 30, ms, ps), (v, err, e)


@pytest.mark.parametrize("FT", FTS)
def test_cold_column_ice_and_snow_branches(FT):
    """A cold sounding (surface θ = 268 K) exercises λ(T) < 1, ice autoconversion, snow accretion, melting and
    deposition — branches the warm default never reaches."""
    cs = M.k1d_case(FT, nc=2, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                    precipitation_choice="Precipitation1M", tht_0=268.0, tht_1=268.0, tht_2=280.0,
                    rv_0=0.004, rv_1=0.0035, rv_2=0.001)
    assert cs.T.min() < 263 and cs.T.max() < 273.15
    g, o = pair(cs)
    g.step(0.0, 400)
    o.step(0.0, 400)
    Yo = o.get_state()
    n = list(cs.var_names)
    assert Yo[:, n.index("ρq_ice")].max() > 1e-6 and Yo[:, n.index("ρq_sno")].max() > 1e-9
    e = errors(g.get_state(), Yo, cs.var_names)
    for v, err in e.items():
        assert err <= (1e-6 if FT == np.float64 else 2e-3), (v, err, e)
    t = 400.0
    e = rhs_errors(g.rhs(t), o.rhs(t), Yo, cs.var_names, cs.cfg.dt, FT)
    for v, err in e.items():
        assert err <= (1e-6 if FT == np.float64 else 2e-2), (v, err, e)


@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("ps,rf", [("Precipitation1M", "CliMA_1M"), ("Precipitation1M", "KK2000"),
                                   ("Precipitation1M", "B1994"), ("Precipitation1M", "TC1980"),
                                   ("Precipitation1M", "LD2004"), ("Precipitation1M", "VarTimeScaleAcnv"),
                                   ("Precipitation2M", "SB2006"), ("Precipitation2M", "SB2006NL")])
def test_box_model_parity(FT, ps, rf):
    """run_box_simulation.jl defaults, an ensemble of 37 boxes (ragged: not a multiple of 32)."""
    nc = 37
    cs = M.box_case(FT, nc=nc, precipitation_choice=ps, rain_formation_choice=rf,
                    qt=np.linspace(4e-4, 2.5e-3, nc), Nd=np.geomspace(3e7, 3e8, nc))
    g, o = pair(cs)
    e = rhs_errors(g.rhs(0.0), o.rhs(0.0), cs.Y0, cs.var_names, cs.cfg.dt, FT)
    for v, err in e.items():
        assert err <= rhs_tol(FT, v), (v, err, e)
    g.step(0.0, 900)
    o.step(0.0, 900)
    Yg, Yo = g.get_state(), o.get_state()
    e = errors(Yg, Yo, cs.var_names)
    for v, err in e.items():
        assert err <= (1e-8 if FT == np.float64 else (5e-2 if v.startswith("N_") else 1e-3)), (v, err, e)
    n = list(cs.var_names)
    # the q_tot source is exactly 0; (u + 2u)/3 and (3u + u)/4 of the SSPRK33 combine may still round
    np.testing.assert_allclose(Yg[:, n.index("ρq_tot")], cs.Y0[:, n.index("ρq_tot")], rtol=2 * 900 * np.finfo(FT).eps)


@pytest.mark.parametrize("FT", FTS)
def test_box_precip_sinks_evaporation(FT):
    cs = M.box_case(FT, nc=5, precipitation_choice="Precipitation2M", qt=np.linspace(8e-4, 2e-3, 5), precip_sinks=1)
    g, o = pair(cs)
    g.step(0.0, 300)
    o.step(0.0, 300)
    e = errors(g.get_state(), o.get_state(), cs.var_names)
    for v, err in e.items():
        assert err <= (1e-8 if FT == np.float64 else (5e-2 if v.startswith("N_") else 5e-3)), (v, err, e)


@pytest.mark.parametrize("FT", FTS)
def test_calibration_style_ensemble_per_column_inputs(FT):
    """Per-column w1, prescribed Nd, q_surf and per-column ρ_dry/T profiles; 70 columns (ragged tile)."""
    cs = M.k1d_ensemble_case(FT, nc=70, n_profiles=4, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                             precipitation_choice="Precipitation2M")
    g, o = pair(cs)
    g.step(0.0, 240)
    o.step(0.0, 240)
    Yg, Yo = g.get_state(), o.get_state()
    e = errors(Yg, Yo, cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 240, "NonEquilibriumMoisture", "Precipitation2M"), (v, err, e)
    # members really differ
    n = list(cs.var_names)
    assert np.unique(Yo[:, n.index("ρq_liq")].max(axis=1)).size > 20


@pytest.mark.parametrize("FT", FTS)
def test_col_sed_model(FT):
    """make_rhs_function_col_sed (K1DModel/ode_utils.jl:56-73): collisions + sedimentation only."""
    cs0, Y, t = developed_state(FT, 900, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    cs = M.k1d_case(FT, nc=2, n_elem=48, model=MODEL_K1D_COL_SED, moisture_choice="NonEquilibriumMoisture",
                    precipitation_choice="Precipitation2M")
    g, o = pair(cs, Y)
    e = rhs_errors(g.rhs(t), o.rhs(t), Y, cs.var_names, cs.cfg.dt, FT)
    for v, err in e.items():
        assert err <= rhs_tol(FT, v), (v, err, e)
    g.step(t, 60)
    o.step(t, 60)
    e = errors(g.get_state(), o.get_state(), cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 60, "NonEquilibriumMoisture", "Precipitation2M"), (v, err, e)


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("nc", [1, 31, 32, 33, 65])
def test_state_roundtrip_is_bit_exact_for_ragged_column_counts(FT, nc):
    cs = M.k1d_case(FT, nc=nc, n_elem=7, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    rng = np.random.default_rng(nc)
    Y = rng.standard_normal(cs.Y0.shape).astype(FT)
    g = M.make_handle(cs, Handle)
    g.set_state(Y)
    assert np.array_equal(g.get_state(), Y)
    g.step(0.0, 0)  # zero steps is a no-op
    assert np.array_equal(g.get_state(), Y)


@pytest.mark.parametrize("FT", FTS)
@pytest.mark.parametrize("nz", [3, 4, 5, 11, 128, 200])
def test_small_and_tall_columns(FT, nz):
    cs = M.k1d_case(FT, nc=3, n_elem=nz, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M")
    g, o = pair(cs)
    g.step(0.0, 40)
    o.step(0.0, 40)
    e = errors(g.get_state(), o.get_state(), cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 40, "NonEquilibriumMoisture", "Precipitation1M"), (nz, v, err)


def test_splitting_a_run_into_launches_is_bitwise_invariant():
    """K steps fused in one launch == the same K steps in 1-step launches (state stays in shared memory vs
    round-trips through HBM)."""
    cs, Y, t = developed_state(np.float32, 600, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation2M")
    a, b = M.make_handle(cs, Handle), M.make_handle(cs, Handle)
    a.set_state(Y)
    b.set_state(Y)
    a.step(t, 24)
    b.set_steps_per_launch(1)
    b.step(t, 24)
    assert b.launch_count - a.launch_count == 23
    assert np.array_equal(a.get_state(), b.get_state())


def test_columns_do_not_interact():
    """Replicated columns give bitwise identical results whatever their position in a tile."""
    cs = M.k1d_case(np.float32, nc=100, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                    precipitation_choice="Precipitation2M")
    g = M.make_handle(cs, Handle)
    g.step(0.0, 300)
    Y = g.get_state()
    assert np.all(Y == Y[0:1])


@pytest.mark.parametrize("ms,ps", [("EquilibriumMoisture", "NoPrecipitation"), ("NonEquilibriumMoisture", "Precipitation1M")])
def test_zero_flow_advection_adds_nothing_on_gpu(ms, ps):
    """k1d_unit_test.jl:109-138 on the device: with ρw = 0 and no rain the rhs of ρq_tot is exactly 0."""
    cs = M.k1d_case(np.float64, nc=2, n_elem=5, z_max=100.0, w1=0.0, moisture_choice=ms, precipitation_choice=ps)
    g = M.make_handle(cs, Handle)
    dY = g.rhs(13.0)
    assert np.all(dY[:, 0] == 0)


def test_error_behaviour():
    cs = M.k1d_case(np.float32, nc=2, n_elem=8)
    g = Handle(cs.cfg)
    with pytest.raises(lib.KidCudaError, match="kid_set_params"):
        g.step(0.0, 1)
    bad = M.k1d_case(np.float32, nc=2, n_elem=8)
    bad.cfg.nz = 2
    with pytest.raises(lib.KidCudaError, match="at least 3 levels"):
        Handle(bad.cfg)
    bad.cfg.nz = 8
    bad.cfg.precip, bad.cfg.sedimentation = 1, 2   # Chen2022 sedimentation without a precipitating species (0M)
    with pytest.raises(lib.KidCudaError, match="Chen2022"):
        Handle(bad.cfg)
    with pytest.raises(lib.KidCudaError, match="KID_NPARAMS"):
        g.set_params(np.zeros(3))


# ------------------------------------------- BASELINE-size, size-independent properties
def test_full_size_box_ensemble_1M_boxes():
    """BASELINE config 2 (1,048,576 boxes, Precipitation1M): replicated ICs must reproduce the small case
    bitwise, q_tot must be untouched and liquid + rain conserved."""
    nc = 1 << 20
    small = M.box_ensemble_case(np.float64, nc=4096, n_ic=8, precipitation_choice="Precipitation1M",
                                rain_formation_choice="CliMA_1M")
    reps = nc // 4096
    big = M.Case(cfg=small.cfg, params=small.params, ρ_dry=np.tile(small.ρ_dry, (reps, 1)), T=np.tile(small.T, (reps, 1)),
                 Y0=np.tile(small.Y0, (reps, 1, 1)), var_names=small.var_names,
                 col_scalars={k: np.tile(v, reps) for k, v in small.col_scalars.items()})
    big.cfg.nc = nc
    gb = M.make_handle(big, Handle)
    gb.step(0.0, 200)
    Yb = gb.get_state()
    small.cfg.nc = 4096
    gs = M.make_handle(small, Handle)
    gs.step(0.0, 200)
    Ys = gs.get_state()
    assert np.array_equal(Yb.reshape(reps, 4096, -1), np.broadcast_to(Ys.reshape(1, 4096, -1), (reps, 4096, Ys.shape[1])))
    n = list(small.var_names)
    np.testing.assert_allclose(Yb[:, n.index("ρq_tot")], big.Y0[:, n.index("ρq_tot")], rtol=2 * 200 * np.finfo(np.float64).eps)
    np.testing.assert_allclose((Yb[:, n.index("ρq_liq")] + Yb[:, n.index("ρq_rai")]),
                               (big.Y0[:, n.index("ρq_liq")] + big.Y0[:, n.index("ρq_rai")]), rtol=1e-12)


def test_full_size_k1d_ensemble_65536_columns():
    """BASELINE config 3 (65,536 columns x 128 levels, NonEq + 1M): tile replication property + finite."""
    nc = 65536
    small = M.k1d_ensemble_case(np.float32, nc=256, n_profiles=4, moisture_choice="NonEquilibriumMoisture",
                                precipitation_choice="Precipitation1M")
    reps = nc // 256
    big = M.Case(cfg=small.cfg, params=small.params, ρ_dry=np.tile(small.ρ_dry, (reps, 1)), T=np.tile(small.T, (reps, 1)),
                 Y0=np.tile(small.Y0, (reps, 1, 1)), var_names=small.var_names,
                 col_scalars={k: np.tile(v, reps) for k, v in small.col_scalars.items()})
    big.cfg.nc = nc
    gb = M.make_handle(big, Handle)
    gb.step(0.0, 60)
    Yb = gb.get_state()
    small.cfg.nc = 256
    gs = M.make_handle(small, Handle)
    gs.step(0.0, 60)
    Ys = gs.get_state()
    assert np.all(np.isfinite(Yb))
    assert np.array_equal(Yb.reshape(reps, -1), np.broadcast_to(Ys.reshape(1, -1), (reps, Ys.size)))


# ------------------------------------------------------- per-column parameter sets (calibration ensembles)
def _run_multi_set(cs, t0, nsteps):
    g = M.make_handle(cs, Handle)
    g.step(t0, nsteps)
    Yg = g.get_state()
    Yo = np.empty_like(Yg)
    for cols, sub in M.split_by_parameter_set(cs):
        o = M.make_handle(sub, OracleHandle)
        o.step(t0, nsteps)
        Yo[cols] = o.get_state()
    return Yg, Yo


@pytest.mark.parametrize("FT", FTS)
def test_per_column_parameter_sets_k1d_1m(FT):
    """BASELINE config 3 style: NonEq + 1M, every member with its own (χv_rai, χa_rai, τ_acnv) as in
    calibration/examples/grid_search/grid_search.jl:10-13; 45 columns over 7 parameter sets."""
    base = M.k1d_case(FT, nc=45, n_elem=32, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M")
    ov = [{"rain_terminal_velocity_size_relation_coefficient_chiv": 0.2 + 0.3 * s,
           "rain_cross_section_size_relation_coefficient_chia": 0.5 + 0.7 * s,
           "rain_autoconversion_timescale": 500.0 + 200.0 * s} for s in range(7)]
    cs = M.with_parameter_sets(base, ov)
    Yg, Yo = _run_multi_set(cs, 0.0, 500)
    e = errors(Yg, Yo, cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 500, "NonEquilibriumMoisture", "Precipitation1M"), (v, err, e)
    n = list(cs.var_names)
    per_set = [Yo[cs.column_to_set == s][:, n.index("ρq_rai")].max() for s in range(7)]
    assert np.unique(np.round(per_set, 12)).size > 3  # the sets really change the physics


@pytest.mark.parametrize("FT", FTS)
def test_per_column_parameter_sets_k1d_2m(FT):
    """2M: kcc, kcr varied by ±25 % (priors of calibration/examples/Box/config.jl:25-30), one set per column."""
    base = M.k1d_case(FT, nc=9, n_elem=32, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    ov = [{"SB2006_collection_kernel_coeff_kcc": 4.44e9 * (0.75 + 0.0625 * s),
           "SB2006_collection_kernel_coeff_kcr": 5.25 * (1.25 - 0.0625 * s), "prescribed_flow_w1": 1.5 + 0.05 * s}
          for s in range(9)]
    cs = M.with_parameter_sets(base, ov)
    cs.col_scalars.pop(lib.COL_W1, None)  # w1 comes from each column's parameter set
    Yg, Yo = _run_multi_set(cs, 0.0, 500)
    e = errors(Yg, Yo, cs.var_names)
    for v, err in e.items():
        assert err <= state_tol(FT, v, 500, "NonEquilibriumMoisture", "Precipitation2M"), (v, err, e)


@pytest.mark.parametrize("FT", FTS)
def test_per_column_parameter_sets_box(FT):
    base = M.box_case(FT, nc=70, precipitation_choice="Precipitation2M", qt=np.linspace(6e-4, 2e-3, 70))
    ov = [{"SB2006_collection_kernel_coeff_kcc": 4.44e9 * (0.75 + 0.1 * s), "SB2006_collection_kernel_coeff_kcr": 5.25 * (0.8 + 0.1 * s)}
          for s in range(5)]
    cs = M.with_parameter_sets(base, ov)
    Yg, Yo = _run_multi_set(cs, 0.0, 600)
    e = errors(Yg, Yo, cs.var_names)
    for v, err in e.items():
        assert err <= (1e-8 if FT == np.float64 else (5e-2 if v.startswith("N_") else 5e-3)), (v, err, e)


def test_handles_with_different_parameters_do_not_interfere():
    """Distinct handles are independent (the reference's calibration runs one simulation per thread with separate aux,
    calibration/OptimizationUtils.jl:206): the parameter set travels with each launch, not in a global symbol."""
    a_case = M.box_case(np.float32, nc=33, precipitation_choice="Precipitation2M", qt=1.5e-3)
    b_case = M.box_case(np.float32, nc=33, precipitation_choice="Precipitation2M", qt=1.5e-3,
                        microphys_params={"SB2006_collection_kernel_coeff_kcc": 2 * 4.44e9})
    alone = M.make_handle(a_case, Handle)
    alone.step(0.0, 300)
    ref = alone.get_state()
    a, b = M.make_handle(a_case, Handle), M.make_handle(b_case, Handle)  # b created (and its params set) after a
    for s in range(0, 300, 50):  # interleaved launches
        a.step(float(s), 50)
        b.step(float(s), 50)
    assert np.array_equal(a.get_state(), ref)
    assert not np.array_equal(b.get_state(), ref)


def test_pipelined_ensemble_matches_single_handle():
    """Column chunks on several handles/streams with asynchronous pinned transfers give the same bits as one handle."""
    import torch
    from kid_b200.solve import PipelinedEnsemble
    cs = M.k1d_ensemble_case(np.float32, nc=200, n_profiles=3, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                             precipitation_choice="Precipitation2M")
    one = M.make_handle(cs, Handle)
    one.step(0.0, 120)
    ref = one.get_state()
    Yp = torch.from_numpy(cs.Y0.copy()).pin_memory().numpy()
    pipe = PipelinedEnsemble(cs, nchunks=3)
    pipe.solve(Yp, 0.0, 120)
    assert np.array_equal(Yp, ref)


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_chen2022_rain_sedimentation_parity(FT):
    """sedimentation_choice = "Chen2022" with Precipitation2M: terminal velocities, rhs and stepped state vs the oracle,
    and a different answer than the SB2006 velocities."""
    kw = dict(nc=8, n_profiles=2, n_elem=32, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    cs = M.k1d_ensemble_case(FT, sedimentation_scheme_choice="Chen2022", **kw)
    g, o = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    g.step(0.0, 900)
    o.set_state(g.get_state())
    for f in ("term_vel_rai", "term_vel_N_rai"):
        a, b = g.get_aux(900.0, f), o.get_aux(900.0, f)
        assert b.max() > 0.5 and np.abs(a - b).max() <= (1e-9 if FT is np.float64 else 2e-4) * b.max(), f
    sb = M.make_handle(M.k1d_ensemble_case(FT, **kw), Handle)
    sb.set_state(g.get_state())
    assert np.abs(sb.get_aux(900.0, "term_vel_rai") - g.get_aux(900.0, "term_vel_rai")).max() > 0.05
    g2, o2 = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    g2.step(0.0, 300)
    o2.step(0.0, 300)
    e = errors(g2.get_state(), o2.get_state(), cs.var_names)
    tol = 1e-6 if FT is np.float64 else 5e-2
    assert max(e.values()) < tol, e


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_chen2022_sedimentation_with_precipitation1m(FT):
    """sedimentation_choice = "Chen2022" with Precipitation1M (CMP.Chen2022VelType, helper_functions.jl:36-37): the rain
    velocity is Chen et al. (2022) eq. 20 over the Marshall-Palmer distribution (snow keeps the Blk1M law; warm columns
    carry none)."""
    kw = dict(nc=8, n_profiles=2, n_elem=32, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M")
    cs = M.k1d_ensemble_case(FT, sedimentation_scheme_choice="Chen2022", **kw)
    g, o = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    g.step(0.0, 900)
    o.set_state(g.get_state())
    a, b = g.get_aux(900.0, "term_vel_rai"), o.get_aux(900.0, "term_vel_rai")
    assert b.max() > 0.5 and np.abs(a - b).max() <= (1e-9 if FT is np.float64 else 2e-4) * b.max()
    blk = M.make_handle(M.k1d_ensemble_case(FT, **kw), Handle)
    blk.set_state(g.get_state())
    assert np.abs(blk.get_aux(900.0, "term_vel_rai") - a).max() > 0.05   # not the Blk1M law
    g2, o2 = M.make_handle(cs, Handle), M.make_handle(cs, OracleHandle)
    g2.step(0.0, 300)
    o2.step(0.0, 300)
    e = errors(g2.get_state(), o2.get_state(), cs.var_names)
    assert np.isfinite(g2.get_state()).all()
    assert max(e.values()) < (1e-6 if FT is np.float64 else 5e-3), e


def test_checkpoint_restart_is_bitwise():
    """kid_get_state + kid_set_state on a fresh handle resume a K1D run exactly (restart for free)."""
    cs = M.k1d_ensemble_case(np.float32, nc=40, n_profiles=3, n_elem=128, moisture_choice="NonEquilibriumMoisture",
                             precipitation_choice="Precipitation2M")
    a = M.make_handle(cs, Handle)
    a.step(0.0, 240)
    b = M.make_handle(cs, Handle)
    b.step(0.0, 100)
    ckpt = b.get_state()
    c = M.make_handle(cs, Handle)
    c.set_state(ckpt)
    c.step(100.0, 140)
    assert np.array_equal(a.get_state(), c.get_state())
