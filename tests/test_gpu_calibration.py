"""Device-side calibration observables (kid_obs_*) and the batched run_ensembles against the numpy restatement of
get_variable_data_from_ODE / ODEsolution2Gvector applied to downloaded states."""
import numpy as np
import pytest

from parity import on_backend

from kid_b200 import calibration as CAL
from kid_b200 import model as M
from kid_b200.lib import Handle, KidCudaError
from oracle import diagnostics as OD
from oracle.oracle import OracleHandle

pytestmark = pytest.mark.gpu

VARS_2M = ["rho", "qt", "ql", "qr", "qv", "qlr", "rt", "rl", "rr", "rv", "rlr", "Nl", "Nr", "Na",
           "rain averaged terminal velocity", "rainrate"]


def _reference_states(case, times, FT):
    """States at the save times from the device (stepping parity is covered elsewhere) and v_t from the C++ oracle."""
    h = M.make_handle(case, Handle)
    o = M.make_handle(case, OracleHandle)
    states, vts, done = [], [], 0
    for t in times:
        n = int(round(t / case.cfg.dt))
        if n > done:
            h.step(done * case.cfg.dt, n - done)
            done = n
        Y = h.get_state()
        o.set_state(Y)
        states.append(Y)
        vts.append(o.get_aux(t, "term_vel_rai"))
    return states, vts


def _column_dicts(case, states, c):
    return [{n: Y[c, i, :] for i, n in enumerate(case.var_names)} for Y in states]


@pytest.mark.parametrize("FT,rtol", [(np.float64, 1e-7), (np.float32, 2e-3)])
def test_gvector_unfiltered_matches_restated_reference(FT, rtol):
    case = M.k1d_ensemble_case(FT, nc=5, n_profiles=2, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation2M")
    times = [300.0, 600.0, 900.0]
    G = CAL.solve_gvector(case, VARS_2M, times)
    states, vts = _reference_states(case, times, FT)
    assert G.shape == (5, len(times) * len(VARS_2M) * 32)
    for c in range(5):
        rd = case.ρ_dry[c] if case.ρ_dry.ndim == 2 else case.ρ_dry
        ref = OD.gvector_unfiltered(_column_dicts(case, states, c), rd, VARS_2M, [v[c] for v in vts])
        # per variable block: relative to the block's scale (v_t differs from the oracle at the rhs tolerance)
        g, r = G[c].reshape(len(times), len(VARS_2M), 32), ref.reshape(len(times), len(VARS_2M), 32)
        for j, name in enumerate(VARS_2M):
            scale = np.abs(r[:, j]).max() + 1e-300
            tol = rtol if "rain" in name else (1e-12 if FT is np.float64 else 1e-6)
            assert np.abs(g[:, j] - r[:, j]).max() <= tol * scale, (name, c)


def test_gvector_filtered_matches_restated_reference():
    FT = np.float64
    case = M.k1d_ensemble_case(FT, nc=3, n_profiles=2, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation1M")
    variables = ["rho", "ql", "qr", "rainrate"]
    filt = CAL.make_filter_props([32] * 4, [0.0, 400.0, 800.0, 1200.0], apply=True, nz_per_filtered_cell=[1, 2, 4, 8],
                                 nt_per_filtered_cell=2)
    assert np.allclose(filt["saveat_t"], [100, 300, 500, 700, 900, 1100])
    G = CAL.solve_gvector(case, variables, None, filt)
    states, vts = _reference_states(case, filt["saveat_t"], FT)
    for c in range(3):
        rd = case.ρ_dry[c] if case.ρ_dry.ndim == 2 else case.ρ_dry
        ref = OD.gvector_filtered(_column_dicts(case, states, c), rd, variables, filt, [v[c] for v in vts])
        assert G[c].shape == ref.shape == (3 * (32 + 16 + 8 + 4),)
        assert np.abs(G[c] - ref).max() <= 1e-7 * np.abs(ref).max()


def test_observable_missing_in_style_is_an_error():
    case = M.k1d_case(np.float64, nc=1, n_elem=16, moisture_choice="EquilibriumMoisture", precipitation_choice="Precipitation0M")
    h = M.make_handle(case, Handle)
    with pytest.raises(KidCudaError):
        h.obs_configure(["ql"])
    with pytest.raises(KidCudaError):
        h.obs_configure(["rainrate"])
    assert h.obs_configure(["qt", "rho"]) == 32


def test_run_ensembles_batched_equals_member_by_member():
    """(member, case) -> column mapping: the batched ensemble returns what separate single-simulation runs return."""
    FT = np.float64
    config = {
        "model": dict(model="KiD", moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M",
                      n_elem=24, dt=2.0, t_end=600.0, t_calib=[0.0, 300.0, 600.0],
                      filter=dict(apply=True, nz_unfiltered=[24, 24], nz_per_filtered_cell=[2, 3], nt_per_filtered_cell=3)),
        "observations": dict(data_names=["rlr", "rainrate"],
                             cases=[dict(w1=2.0, p0=100000.0, Nd=1e8), dict(w1=1.5, p0=99000.0, Nd=5e7)]),
    }
    u_names = ["rain_terminal_velocity_size_relation_coefficient_chiv", "rain_cross_section_size_relation_coefficient_chia"]
    u = np.array([[1.0, 0.6, 1.4], [1.0, 2.0, 0.7]])
    G = CAL.run_ensembles(u, u_names, config, 3, FT=FT)
    assert len(G) == 3 and G[0].shape == (2 * 2 * (12 + 8),)
    assert not np.allclose(G[0], G[1]) and np.isfinite(np.stack(G)).all()
    for m in range(3):
        single = CAL.run_dyn_model(u[:, m], u_names, config, FT=FT)
        assert np.array_equal(single, G[m])
    # the first member/case is the plain default simulation with that case's scalars
    plain = M.k1d_case(FT, nc=1, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M",
                       n_elem=24, dt=2.0, t_end=600.0, w1=2.0, p0=100000.0, prescribed_Nd=1e8)
    filt = CAL.make_filter_props([24, 24], [0.0, 300.0, 600.0], apply=True, nz_per_filtered_cell=[2, 3], nt_per_filtered_cell=3)
    Gp = CAL.solve_gvector(plain, ["rlr", "rainrate"], None, filt)
    assert np.allclose(Gp[0], G[0][:Gp.shape[1]], rtol=1e-12, atol=0)


def test_reduce_max_on_device_matches_downloaded_field():
    case = M.k1d_ensemble_case(np.float32, nc=70, n_profiles=4, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation2M")
    h = M.make_handle(case, Handle)
    h.step(0.0, 400)
    for f in ("q_liq", "N_liq", "precip_sources.q_rai", "term_vel_rai"):
        assert h.reduce_max(400.0, f) == float(np.max(h.get_aux(400.0, f)))


def test_get_aux_many_equals_field_by_field():
    from kid_b200.lib import AUX_FIELDS
    case = M.k1d_ensemble_case(np.float32, nc=40, n_profiles=4, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation2M")
    h = M.make_handle(case, Handle)
    h.step(0.0, 500)
    l0 = h.launch_count
    A = h.get_aux_many(500.0, AUX_FIELDS)
    assert h.launch_count - l0 == 2  # one rhs evaluation + one layout transpose for all 29 fields
    assert A.shape == (40, len(AUX_FIELDS), 32)
    for i, f in enumerate(AUX_FIELDS):
        assert np.array_equal(A[:, i, :], h.get_aux(500.0, f), equal_nan=True), f


def test_output_pipeline_matches_synchronous_output():
    """Asynchronous double-buffered output (kid_output_enqueue) returns what simulation_output returns at the same times,
    while the integration continues between the enqueue and the wait."""
    from kid_b200.solve import OutputPipeline, simulation_output
    fields = ["q_liq", "q_rai", "N_liq", "term_vel_rai", "precip_sources.q_rai", "ρ", "p"]
    case = M.k1d_ensemble_case(np.float64, nc=9, n_profiles=3, n_elem=32, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation2M")
    times = [0.0, 100.0, 200.0, 300.0, 400.0, 500.0, 600.0]
    ha, hb = M.make_handle(case, Handle), M.make_handle(case, Handle)
    pipe = OutputPipeline(ha, fields)
    sync = []
    for i, t in enumerate(times):
        if i:
            ha.step(times[i - 1], 100)
            hb.step(times[i - 1], 100)
        pipe.enqueue(t)
        sync.append((t, simulation_output(hb, t, fields)))
    res = pipe.drain()
    assert [t for t, _ in res] == times
    for (t, a), (_, b) in zip(res, sync):
        for f in fields:
            assert np.array_equal(a[f], b[f]), (t, f)
        assert a["ql_max"] == b["ql_max"]


def test_box_calibration_ensemble_matches_the_oracle_backend():
    from oracle.oracle import OracleEnsembleHandle
    from test_calibration_host import BOX_CONFIG, BOX_NAMES
    u = np.array([[4.44e9, 6e9, 3e9], [5.25, 4.0, 6.5]])
    Gd = np.stack(CAL.run_ensembles(u, BOX_NAMES, BOX_CONFIG, 3, FT=np.float64))
    Go = np.stack(on_backend(OracleEnsembleHandle, CAL.run_ensembles, u, BOX_NAMES, BOX_CONFIG, 3, FT=np.float64))
    g, o = Gd.reshape(3, 2, 3, 4), Go.reshape(3, 2, 3, 4)  # [member][case][time][variable]
    for v in range(4):
        assert np.abs(g[..., v] - o[..., v]).max() <= 1e-7 * np.abs(o[..., v]).max(), v


def test_col_sed_driver_and_calibration_match_the_oracle_backend():
    from kid_b200.solve import run_KiD_col_sed_simulation
    from oracle.oracle import OracleEnsembleHandle, OracleHandle
    opts = dict(n_elem=24, z_max=2400.0, t_end=300.0, dt=2.0, dt_output=150.0, sedimentation_choice="Chen2022")
    a = run_KiD_col_sed_simulation(np.float64, opts, nc=2)
    b = on_backend(OracleHandle, run_KiD_col_sed_simulation, np.float64, opts, nc=2)
    for ua, ub in zip(a.u, b.u):
        assert np.abs(ua - ub).max() <= 1e-7 * np.abs(ub).max()
    config = {
        "model": dict(model="KiD_col_sed", precipitation_choice="Precipitation2M", rain_formation_choice="SB2006",
                      sedimentation_choice="SB2006", rhod=1.0, z_min=0.0, z_max=1200.0, n_elem=12, dt=2.0, t_end=240.0,
                      t_calib=[0.0, 120.0, 240.0], filter=dict(apply=False, nz_unfiltered=[12, 12])),
        "observations": dict(data_names=["qr", "Nr"], cases=[dict(qt=1e-3, Nd=1e8, k=2.0), dict(qt=2e-3, Nd=2e8, k=3.0)]),
    }
    u = np.array([[4.44e9, 6e9]])
    names = ["SB2006_collection_kernel_coeff_kcc"]
    Gd = np.stack(CAL.run_ensembles(u, names, config, 2))
    Go = np.stack(on_backend(OracleEnsembleHandle, CAL.run_ensembles, u, names, config, 2))
    g, o = Gd.reshape(2, 2, 3, 2, 12), Go.reshape(2, 2, 3, 2, 12)  # [member][case][time][variable][level]
    for v in range(2):
        assert np.abs(g[:, :, :, v] - o[:, :, :, v]).max() <= 1e-6 * np.abs(o[:, :, :, v]).max(), v


@pytest.mark.parametrize("FT,tol", [(np.float64, 1e-9), (np.float32, 2e-3)])
def test_z_top_radar_reflectivity_observable(FT, tol):
    """"Z_top" (helper_functions.jl:164-185): one value per column and save time, Precipitation1M only; against the numpy
    restatement applied to the downloaded state (absolute tolerance in dBZ)."""
    case = M.k1d_ensemble_case(FT, nc=6, n_profiles=2, n_elem=64, moisture_choice="NonEquilibriumMoisture",
                               precipitation_choice="Precipitation1M")
    times = [600.0, 1200.0, 1800.0]
    G = CAL.solve_gvector(case, ["Z_top", "qr"], times)
    assert G.shape == (6, len(times) * (1 + 64))
    td = case.meta["toml"]
    rain = dict(n0=td["rain_drop_size_distribution_coefficient_n0"], r0=td["rain_drop_length_scale"] if "rain_drop_length_scale" in td else 1e-3,
                m0=None, me=td["rain_mass_size_relation_coefficient_me"], Δm=td["rain_mass_size_relation_coefficient_delm"],
                χm=td["rain_mass_size_relation_coefficient_chim"])
    rain["m0"] = 4.0 / 3.0 * np.pi * td["density_liquid_water"] * rain["r0"] ** 3
    h = M.make_handle(case, Handle)
    done, saw_rain = 0, False
    for i, t in enumerate(times):
        h.step(float(done), int(t) - done)
        done = int(t)
        Y = h.get_state()
        for c in range(6):
            u = {n: Y[c, j].astype(np.float64) for j, n in enumerate(case.var_names)}
            rd = (case.ρ_dry[c] if case.ρ_dry.ndim == 2 else case.ρ_dry).astype(np.float64)
            ref = OD.z_top(u, rd, float(case.cfg.z_max - case.cfg.z_min), rain)
            if FT is np.float32 and ref < -300.0:
                ref = -300.0  # rain so thin that λ⁻¹ underflows in Float32: log10(0) = -Inf -> -300 (the reference's replace!)
            got = G[c].reshape(len(times), 65)[i, 0]
            assert abs(got - ref) <= tol * max(1.0, abs(ref)) * 50, (c, t, got, ref)
            saw_rain |= ref > -100.0
    assert saw_rain
    with pytest.raises(KidCudaError, match="radar reflectivity"):
        CAL.solve_gvector(M.k1d_ensemble_case(FT, nc=2, n_profiles=1, n_elem=16, moisture_choice="NonEquilibriumMoisture",
                                              precipitation_choice="Precipitation2M"), ["Z_top"], [10.0])
