"""Host logic of kid_b200.calibration (filter bookkeeping, (member, case) -> column mapping, save-time handling),
driven on the CPU through the oracle backend."""
import numpy as np
import pytest

from parity import on_backend

from kid_b200 import calibration as CAL
from kid_b200 import model as M
from oracle import diagnostics as OD
from oracle.oracle import OracleEnsembleHandle


def test_make_filter_props_matches_reference_bookkeeping():
    f = CAL.make_filter_props([128, 128], [0.0, 600.0, 1200.0], apply=True, nz_per_filtered_cell=[2, 4], nt_per_filtered_cell=3)
    assert list(f["nz_filtered"]) == [64, 32] and f["nt_filtered"] == 2
    assert np.allclose(f["saveat_t"], [100, 300, 500, 700, 900, 1100])  # midpoints of the sub-intervals
    with pytest.raises(AssertionError):
        CAL.make_filter_props([128], [0.0, 1.0], nz_per_filtered_cell=[3])
    g = CAL.make_filter_props([16], [0.0, 10.0])
    assert list(g["nz_per_filtered_cell"]) == [1] and np.allclose(g["saveat_t"], [5.0])


def test_filtered_gvector_with_unit_cells_is_the_unfiltered_one():
    rng = np.random.default_rng(1)
    nz = 8
    states = [{"ρq_tot": rng.uniform(1e-3, 1e-2, nz), "ρq_liq": rng.uniform(0, 1e-3, nz), "ρq_rai": rng.uniform(0, 1e-4, nz)}
              for _ in range(3)]
    rd = rng.uniform(0.9, 1.2, nz)
    variables = ["qt", "rl", "rlr", "rho"]
    filt = CAL.make_filter_props([nz] * 4, [0.0, 1.0, 2.0, 3.0], apply=True)
    a = OD.gvector_unfiltered(states, rd, variables)
    b = OD.gvector_filtered(states, rd, variables, filt)
    assert np.allclose(a, b, rtol=1e-15)


def test_save_times_off_the_step_grid_are_refused():
    case = M.k1d_case(np.float64, nc=1, n_elem=8, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M", dt=2.0)
    with pytest.raises(NotImplementedError):
        on_backend(OracleEnsembleHandle, CAL.solve_gvector, case, ["qt"], [3.0])


def test_run_ensembles_on_the_oracle_backend():
    config = {
        "model": dict(model="KiD", moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M",
                      n_elem=12, dt=5.0, t_end=200.0, t_calib=[0.0, 100.0, 200.0],
                      filter=dict(apply=True, nz_unfiltered=[12, 12], nz_per_filtered_cell=[2, 3], nt_per_filtered_cell=2)),
        "observations": dict(data_names=["rlr", "rainrate"],
                             cases=[dict(w1=2.0, p0=100000.0, Nd=1e8), dict(w1=1.5, p0=99000.0, Nd=5e7)]),
    }
    u_names = ["rain_terminal_velocity_size_relation_coefficient_chiv", "rain_cross_section_size_relation_coefficient_chia"]
    u = np.array([[1.0, 0.6], [1.0, 2.0]])
    G = on_backend(OracleEnsembleHandle, CAL.run_ensembles, u, u_names, config, 2)
    assert len(G) == 2 and G[0].shape == (2 * 2 * (6 + 4),) and np.isfinite(np.stack(G)).all()
    case = CAL.kid_calibration_case(np.float64, config["model"], config["observations"]["cases"], u, u_names)
    assert list(case.meta["col_member"]) == [0, 0, 1, 1] and list(case.meta["col_case"]) == [0, 1, 0, 1]
    assert case.params.shape[0] == 2 and case.ρ_dry.shape == (4, 12)
    assert not np.allclose(case.ρ_dry[0], case.ρ_dry[1])  # p0 enters the hydrostatic profile
    # member 0 / case 0 is the plain single-column run
    single = on_backend(OracleEnsembleHandle, CAL.run_dyn_model, u[:, 0], u_names, config)
    assert np.array_equal(single, G[0])


BOX_CONFIG = {
    "model": dict(model="Box", precipitation_choice="Precipitation2M", rain_formation_choice="SB2006", rhod=1.22, dt=1.0,
                  t_end=1200.0, t_calib=[0.0, 600.0, 1200.0], filter=dict(apply=False, nz_unfiltered=[1, 1, 1, 1])),
    "observations": dict(data_names=["ql", "qr", "Nl", "Nr"],
                         cases=[dict(qt=1e-3, Nd=1e8, k=2.0), dict(qt=1.5e-3, Nd=3e8, k=4.0)]),
}
BOX_NAMES = ["SB2006_collection_kernel_coeff_kcc", "SB2006_collection_kernel_coeff_kcr"]


def test_box_calibration_ensemble_on_the_oracle_backend():
    u = np.array([[4.44e9, 6e9], [5.25, 4.0]])
    G = on_backend(OracleEnsembleHandle, CAL.run_ensembles, u, BOX_NAMES, BOX_CONFIG, 2)
    assert len(G) == 2 and G[0].shape == (2 * 3 * 4,) and np.isfinite(np.stack(G)).all()
    case = CAL.box_calibration_case(np.float64, BOX_CONFIG["model"], BOX_CONFIG["observations"]["cases"], u, BOX_NAMES)
    # every (member, case) pair has its own parameter set: the case's k is SB2006's cloud ν
    assert case.params.shape[0] == 4 and list(case.column_to_set) == [0, 1, 2, 3]
    from kid_b200.parameters import param_names
    nu = case.params[:, param_names().index("sb_nu_c")]
    assert list(nu) == [2.0, 4.0, 2.0, 4.0]
    single = on_backend(OracleEnsembleHandle, CAL.run_dyn_model, u[:, 1], BOX_NAMES, BOX_CONFIG)
    assert np.array_equal(single, G[1])
    # rain forms from cloud water: ql decreases, qr increases over the run in every member/case
    g = G[0].reshape(2, 3, 4)   # [case][time][variable]
    assert np.all(g[:, -1, 0] < g[:, 0, 0]) and np.all(g[:, -1, 1] > g[:, 0, 1])


def test_col_sed_driver_and_calibration_on_the_oracle_backend():
    """KiD_col_sed: initial_condition_0d at every level, collisions + sedimentation only (run_KiD_col_sed_simulation.jl,
    calibration/KiDUtils.jl:155-236)."""
    from kid_b200.solve import run_KiD_col_sed_simulation
    from oracle.oracle import OracleHandle
    sol = on_backend(OracleHandle, run_KiD_col_sed_simulation, np.float64, dict(n_elem=12, z_max=1200.0, t_end=120.0, dt=2.0, dt_output=60.0))
    assert sol.t == [0.0, 60.0, 120.0]
    l0, r1 = sol.variable("ρq_liq")[0, 0], sol.variable("ρq_rai")[-1, 0]
    assert np.allclose(l0, l0[0]) and l0[0] > 0              # the same gamma-distribution state at every level
    assert r1.max() > 0 and np.isfinite(sol.u[-1]).all()     # collisions produce rain, which sediments
    config = {
        "model": dict(model="KiD_col_sed", precipitation_choice="Precipitation2M", rain_formation_choice="SB2006",
                      sedimentation_choice="SB2006", rhod=1.0, z_min=0.0, z_max=1200.0, n_elem=12, dt=2.0, t_end=120.0,
                      t_calib=[0.0, 60.0, 120.0], filter=dict(apply=False, nz_unfiltered=[12, 12])),
        "observations": dict(data_names=["qr", "Nr"], cases=[dict(qt=1e-3, Nd=1e8, k=2.0), dict(qt=2e-3, Nd=2e8, k=3.0)]),
    }
    u = np.array([[4.44e9, 6e9]])
    G = on_backend(OracleEnsembleHandle, CAL.run_ensembles, u, ["SB2006_collection_kernel_coeff_kcc"], config, 2)
    assert len(G) == 2 and G[0].shape == (2 * 3 * 2 * 12,) and np.isfinite(np.stack(G)).all()
    case = CAL.col_sed_calibration_case(np.float64, config["model"], config["observations"]["cases"], u,
                                        ["SB2006_collection_kernel_coeff_kcc"])
    assert case.cfg.model == M.MODEL_K1D_COL_SED and case.params.shape[0] == 4
    assert not np.allclose(case.Y0[0], case.Y0[1])  # the cases differ in qt, Nd and k


# ---- the reference's own calibration tests, with its test configuration (calibration/tests/config.jl:40-112)
def _reference_test_config():
    t_calib = list(np.arange(0.0, 1200.0 + 1e-9, 300.0))
    model = dict(model="KiD", moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M",
                 rain_formation_choice="CliMA_1M", sedimentation_choice="CliMA_1M", precip_sources=True, precip_sinks=True,
                 z_min=0.0, z_max=3000.0, n_elem=10, dt=10.0, t_ini=0.0, t_end=1200.0, t_calib=t_calib, t1=600.0,
                 qtot_flux_correction=False, open_system_activation=False, local_activation=False, r_dry=0.04e-6,
                 std_dry=1.4, κ=0.9, filter=dict(apply=False, nz_unfiltered=[10, 10]))
    obs = dict(data_names=["rl", "qr"], cases=[dict(w1=2.0, p0=99000.0, Nd=50e6), dict(w1=3.0, p0=99000.0, Nd=50e6)])
    return {"model": model, "observations": obs}


def test_reference_known_answer_make_filter():
    """calibration/tests/test_helper_funcs.jl:23-37 ("make filter")."""
    f = CAL.make_filter_props([10], [0, 100, 500], apply=True, nz_per_filtered_cell=[2], nt_per_filtered_cell=2)
    assert f["apply"] is True and list(f["nz_filtered"]) == [5] and f["nt_filtered"] == 2
    assert list(f["saveat_t"]) == [25.0, 75.0, 200.0, 400.0]


def test_reference_run_kid_and_gvector_shapes():
    """calibration/tests/test_KiD_utils.jl:41-78 ("Run KiD") and :80-112 (multiple cases / run_dyn_model): the lengths the
    reference asserts, and G of all cases = concatenation of the per-case G vectors."""
    cfg = _reference_test_config()
    u, u_names = np.array([1e-4]), ["cloud_liquid_water_specific_humidity_autoconversion_threshold"]
    n_elem, n_times, n_cases = 10, 5, 2
    one = dict(cfg, observations=dict(cfg["observations"], cases=cfg["observations"]["cases"][:1], data_names=["rlr", "rv"]))
    G = on_backend(OracleEnsembleHandle, CAL.run_dyn_model, u, u_names, one)
    assert G.shape == (2 * n_elem * n_times,)                       # length(G) == sum(n_heights) * n_times
    filt = dict(apply=True, nz_unfiltered=[n_elem, n_elem], nz_per_filtered_cell=[2, 5], nt_per_filtered_cell=5)
    onef = dict(one, model=dict(one["model"], filter=filt))
    Gf = on_backend(OracleEnsembleHandle, CAL.run_dyn_model, u, u_names, onef)
    assert Gf.shape == ((5 + 2) * (n_times - 1),)                    # sum(nz_filtered) * (n_times - 1)
    G_all = on_backend(OracleEnsembleHandle, CAL.run_dyn_model, u, u_names, cfg)
    assert G_all.shape == (n_elem * n_times * 2 * n_cases,)         # n_heights * n_times * n_vars * n_cases
    first = dict(cfg, observations=dict(cfg["observations"], cases=cfg["observations"]["cases"][:1]))
    G_1 = on_backend(OracleEnsembleHandle, CAL.run_dyn_model, u, u_names, first)
    assert np.array_equal(G_all[:G_1.size], G_1) and np.isfinite(G_all).all()


def test_reference_diagnostics_lwp_rwp_rainrate():
    """calibration/tests/test_diagnostics.jl:1-52, with the same random-vector construction."""
    cfg = _reference_test_config()
    cfg["observations"]["data_names"] = ["rho", "ql", "qr", "rainrate"]
    cfg["model"]["filter"] = CAL.make_filter_props([10] * 4, cfg["model"]["t_calib"])
    n = CAL.get_numbers_from_config(cfg)
    assert n["n_cases"] == 2 and n["n_heights"] == [10] * 4 and n["n_times"] == [5, 5]
    n_single = [sum(n["n_heights"]) * t for t in n["n_times"]]
    vec = np.random.default_rng(0).random(sum(n_single))
    dz = 3000.0 / 10
    lwp, rwp, rr = [], [], []
    for i in range(2):
        f = CAL.get_single_case_fields(CAL.get_case_i_vec(vec, i, n_single), n["n_heights"], n["n_times"][i])
        lwp.append(np.sum(f[1] * f[0], axis=0, keepdims=True) * dz)
        rwp.append(np.sum(f[2] * f[0], axis=0, keepdims=True) * dz)
        r0 = 1.5 * f[3][0, :] - 0.5 * f[3][1, :]
        rr.append(np.where(r0 < 0, 0.0, r0))
    for a, b in ((CAL.liquid_water_path(vec, cfg), lwp), (CAL.rainwater_path(vec, cfg), rwp), (CAL.rainrate(vec, cfg, 0.0), rr)):
        assert all(np.allclose(x, y, rtol=1e-15, atol=0) for x, y in zip(a, b))
    cfg["observations"]["data_names"] = ["rho", "ql", "qr"]
    with pytest.raises(AssertionError):
        CAL.rainrate(vec, cfg, 0.0)
    cfg["observations"]["data_names"] = ["ql", "qr", "rainrate"]
    with pytest.raises(AssertionError):
        CAL.liquid_water_path(vec, cfg)
    # the field layout is the one solve_gvector / kid_obs_get produce: [time][variable][level] per case
    g = np.arange(2 * (3 + 1), dtype=float)   # 2 times x (3 + 1) heights
    f = CAL.get_single_case_fields(g, [3, 1], 2)
    assert f[0].tolist() == [[0, 4], [1, 5], [2, 6]] and f[1].tolist() == [[3, 7]]
