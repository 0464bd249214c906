"""The ODE.solve replacement and the driver entry points (host logic, CPU): run with the oracle standing in as the
backend so that the save/callback bookkeeping is checked without a GPU."""
import numpy as np
import pytest

from parity import on_backend

from kid_b200 import model as M
from kid_b200 import solve as S
from oracle.oracle import OracleHandle


def test_solve_saves_like_ordinarydiffeq_saveat():
    case = M.box_case(np.float64, nc=2, precipitation_choice="Precipitation1M", rain_formation_choice="CliMA_1M",
                      qt=[1e-3, 2e-3], dt=1.0, dt_output=30.0, t_end=120.0)
    sol = on_backend(OracleHandle, S.solve, case, (0.0, 120.0), dt=1.0, saveat=30.0)
    assert sol.t == [0.0, 30.0, 60.0, 90.0, 120.0]
    assert len(sol.u) == 5 and sol.u[0].shape == (2, 5, 1)
    assert np.array_equal(sol.u[0], case.Y0)
    ref = M.make_handle(case, OracleHandle)
    ref.step(0.0, 120)
    assert np.array_equal(sol.u[-1], ref.get_state())
    q_rai = sol.variable("ρq_rai")
    assert q_rai.shape == (5, 2, 1) and np.all(np.diff(q_rai[:, :, 0], axis=0) > 0)


def test_solve_rejects_a_different_dt():
    case = M.box_case(np.float64, nc=1, dt=1.0, t_end=10.0)
    with pytest.raises(ValueError, match="TS.dt"):
        on_backend(OracleHandle, S.solve, case, (0.0, 10.0), dt=2.0)


def test_output_cadence_matches_condition_io():
    """netcdf_io.jl:176-185 with dt = 1, dt_output = 30, interval 1.0: the accumulator starts at 30 so the first
    step fires, then every 2nd step; t_max always fires."""
    TS = M.TimeStepping(1.0, 30.0, 10.0)
    cond = S.OutputCadence(TS)
    fired = [t for t in range(1, 11) if cond(float(t))]
    assert fired == [1, 3, 5, 7, 9, 10]


def test_run_KiD_simulation_driver_with_output():
    sol, outputs = on_backend(OracleHandle, S.run_KiD_simulation, np.float64, dict(moisture_choice="NonEquilibriumMoisture",
                                                         precipitation_choice="Precipitation1M", n_elem=16, t_end=20.0,
                                                         dt_output=10.0, FLOAT_TYPE="Float64"),
                                        output=True)
    assert sol.t == [0.0, 10.0, 20.0]
    times = [t for t, _ in outputs]
    assert times[0] == 0.0 and times[-1] == 20.0 and len(times) == 12
    snap = outputs[-1][1]
    assert {"ρ", "p", "T", "q_liq", "term_vel_rai", "precip_sources.q_rai", "ql_max"} <= set(snap)
    assert snap["ρ"].shape == (1, 16) and np.all(np.isfinite(snap["p"]))


def test_run_box_and_k2d_drivers():
    sol = on_backend(OracleHandle, S.run_box_simulation, np.float64, dict(t_end=60.0, dt_output=30.0))
    assert sol.t == [0.0, 30.0, 60.0] and sol.var_names[-1] == "N_aer"
    sol = on_backend(OracleHandle, S.run_K2D_simulation, np.float64, dict(nx=3, nz=6, npoly=2, t_end=4.0, dt_output=2.0))
    assert sol.t == [0.0, 2.0, 4.0] and sol.u[-1].shape == (9, 5, 6)
