"""Calibration forward map on the device: `run_dyn_model` / `run_ensembles` of the reference, batched.

Mirrors (reference):
  make_filter_props                         calibration/HelperFuncs.jl:21-47
  ODEsolution2Gvector (both methods)        calibration/KiDUtils.jl:329-376
  run_KiD / run_KiD_multiple_cases          calibration/KiDUtils.jl:45-151
  run_ensembles                             calibration/OptimizationUtils.jl:197-211

The reference maps one parameter vector to one G vector per call and threads over ensemble members on the CPU; here
every (ensemble member, case) pair is one COLUMN of a single device ensemble — members differ by their parameter set,
cases by (w1, p0, Nd) — and the G vectors are accumulated on the device (kid_obs_*), so one batch of launches serves a
whole EKI iteration and only the filtered observables are downloaded.
"""
from __future__ import annotations

import numpy as np

from . import initial_condition as IC
from . import model as M
from . import parameters as PR
from .lib import COL_PRESCRIBED_ND, COL_Q_SURF, COL_W1, Handle


def make_filter_props(n_z, t_calib, apply: bool = False, nz_per_filtered_cell=None, nt_per_filtered_cell: int = 1) -> dict:
    """calibration/HelperFuncs.jl:21-47."""
    n_z = np.asarray(n_z, dtype=int)
    per = np.ones_like(n_z) if nz_per_filtered_cell is None else np.asarray(nz_per_filtered_cell, dtype=int)
    if np.any(n_z % per != 0):
        raise AssertionError("n_z must be divisible by nz_per_filtered_cell")
    t_calib = np.asarray(t_calib, dtype=float)
    saveat = []
    for i in range(len(t_calib) - 1):
        dt = (t_calib[i + 1] - t_calib[i]) / nt_per_filtered_cell
        saveat.extend(np.linspace(t_calib[i] + dt / 2, t_calib[i + 1] - dt / 2, nt_per_filtered_cell))
    return {"apply": apply, "nz_per_filtered_cell": per, "nt_per_filtered_cell": int(nt_per_filtered_cell),
            "nz_unfiltered": n_z, "nz_filtered": n_z // per, "nt_filtered": len(t_calib) - 1,
            "saveat_t": np.asarray(saveat)}


def _step_index(t: float, t_ini: float, dt: float) -> int:
    n = (t - t_ini) / dt
    if abs(n - round(n)) > 1e-9 * max(1.0, abs(n)):
        raise NotImplementedError(
            f"save time {t} is not a multiple of dt={dt} after t_ini={t_ini}: the reference would interpolate the dense "
            "ODE output there; the device path saves at step boundaries only")
    return int(round(n))


def solve_gvector(case: M.Case, variables, t_calib, filter: dict | None = None, t_ini: float = 0.0,
                  handle=None) -> np.ndarray:
    """ODE.solve(..., saveat) + ODEsolution2Gvector for every column of `case`; returns [nc, n_out] Float64.

    filter["apply"] false (or filter None): the state is observed at the times t_calib (KiDUtils.jl:329-340).
    filter["apply"] true: at filter["saveat_t"], averaged over nt_per_filtered_cell snapshots and blocks of
    nz_per_filtered_cell levels (KiDUtils.jl:342-376)."""
    h = handle or M.make_handle(case)
    dt = float(case.cfg.dt)
    if filter is not None and filter["apply"]:
        times = np.asarray(filter["saveat_t"], dtype=float)
        nt_per = int(filter["nt_per_filtered_cell"])
        n_cells = int(filter["nt_filtered"])
        per = np.asarray(filter["nz_per_filtered_cell"], dtype=np.int32)
    else:
        times = np.asarray(t_calib, dtype=float)
        nt_per, n_cells, per = 1, len(times), None
    h.obs_configure(list(variables), per, n_cells, nt_per)
    done = 0
    for i, t in enumerate(times):
        n = _step_index(float(t), t_ini, dt)
        if n < done:
            raise ValueError("save times must be non-decreasing")
        if n > done:
            h.step(t_ini + done * dt, n - done)
            done = n
        h.obs_accumulate(i // nt_per)
    return h.obs_get()


def kid_calibration_case(FT, model_settings: dict, cases: list[dict], u: np.ndarray, u_names: list[str]) -> M.Case:
    """All (ensemble member, case) simulations of one `run_ensembles` call as the columns of ONE device ensemble.

    u: [n_params, n_ensemble] (the reference's layout); cases: dicts with w1, p0, Nd (config["observations"]["cases"]).
    Column index = member * n_cases + case.  Each case has its own hydrostatic profile (p0 enters ρ_ivp), each member
    its own parameter set (update_parameters!, KiDUtils.jl:378-382)."""
    u = np.atleast_2d(np.asarray(u, dtype=np.float64))
    if u.shape[0] != len(u_names):
        raise ValueError("u must be [n_params, n_ensemble]")
    n_ens, n_cases = u.shape[1], len(cases)
    ms = dict(model_settings)
    model = {"KiD": M.MODEL_K1D}[ms.get("model", "KiD")]
    kw = {k: ms[k] for k in ("moisture_choice", "precipitation_choice", "rain_formation_scheme_choice",
                             "sedimentation_scheme_choice", "z_min", "z_max", "n_elem", "dt", "t_end", "t1",
                             "precip_sources", "precip_sinks", "open_system_activation", "local_activation",
                             "qtot_flux_correction", "r_dry", "std_dry", "kappa", "toml_overrides") if k in ms}
    if "rain_formation_choice" in ms:
        kw["rain_formation_scheme_choice"] = ms["rain_formation_choice"]
    if "sedimentation_choice" in ms:
        kw["sedimentation_scheme_choice"] = ms["sedimentation_choice"]
    if "κ" in ms:
        kw["kappa"] = ms["κ"]
    per_case = [M.k1d_case(FT, nc=1, model=model, w1=c["w1"], p0=c["p0"], prescribed_Nd=c["Nd"], **kw) for c in cases]
    base = per_case[0]
    nc = n_ens * n_cases
    col_case = np.arange(nc) % n_cases
    col_member = np.arange(nc) // n_cases
    import copy
    cfg = copy.copy(base.cfg)
    cfg.nc = nc
    FTt = np.dtype(FT).type
    ens = M.Case(cfg=cfg, params=base.params,
                 ρ_dry=np.stack([per_case[i].ρ_dry for i in col_case]).astype(FTt),
                 T=np.stack([per_case[i].T for i in col_case]).astype(FTt),
                 Y0=np.concatenate([per_case[i].Y0 for i in col_case]).astype(FTt), var_names=base.var_names,
                 col_scalars={COL_W1: np.array([cases[i]["w1"] for i in col_case], dtype=FTt),
                              COL_PRESCRIBED_ND: np.array([cases[i]["Nd"] for i in col_case], dtype=FTt),
                              COL_Q_SURF: np.array([per_case[i].col_scalars[COL_Q_SURF][0] for i in col_case], dtype=FTt)},
                 z_centers=base.z_centers, meta=dict(base.meta, cases=cases, col_case=col_case, col_member=col_member))
    overrides = [{name: float(u[p, m]) for p, name in enumerate(u_names)} for m in range(n_ens)]
    return M.with_parameter_sets(ens, overrides, column_to_set=col_member)


def box_calibration_case(FT, model_settings: dict, cases: list[dict], u: np.ndarray, u_names: list[str]) -> M.Case:
    """run_box_multiple_cases (calibration/KiDUtils.jl:238-261) + run_box (:263-322) for a whole ensemble: column =
    member * n_cases + case.  A case is (qt, Nd, k); k is also the cloud gamma-distribution parameter ν of SB2006
    (`toml_dict["SB2006_cloud_gamma_distribution_coeff_nu"] = k`, KiDUtils.jl:267), so every (member, case) pair has its
    own parameter set."""
    u = np.atleast_2d(np.asarray(u, dtype=np.float64))
    if u.shape[0] != len(u_names):
        raise ValueError("u must be [n_params, n_ensemble]")
    ms = dict(model_settings)
    n_ens, n_cases = u.shape[1], len(cases)
    nc = n_ens * n_cases
    col_case, col_member = np.arange(nc) % n_cases, np.arange(nc) // n_cases
    kw = dict(precipitation_choice=ms.get("precipitation_choice", "Precipitation2M"),
              rain_formation_choice=ms.get("rain_formation_choice", "SB2006"), rhod=ms.get("rhod", 1.22),
              dt=ms.get("dt", 1.0), t_end=ms.get("t_end", 3600.0), precip_sinks=int(ms.get("precip_sinks", 0)),
              microphys_params=dict(ms.get("toml_overrides") or {}))
    per_case = [M.box_case(FT, nc=1, qt=c["qt"], Nd=c["Nd"], k=c["k"], **kw) for c in cases]
    base = per_case[0]
    import copy
    cfg = copy.copy(base.cfg)
    cfg.nc = nc
    FTt = np.dtype(FT).type
    ens = M.Case(cfg=cfg, params=base.params, ρ_dry=np.concatenate([per_case[i].ρ_dry for i in col_case]).astype(FTt),
                 T=np.concatenate([per_case[i].T for i in col_case]).astype(FTt),
                 Y0=np.concatenate([per_case[i].Y0 for i in col_case]).astype(FTt), var_names=base.var_names,
                 col_scalars={COL_PRESCRIBED_ND: np.array([cases[i]["Nd"] for i in col_case], dtype=FTt)},
                 z_centers=base.z_centers, meta=dict(base.meta, cases=cases, col_case=col_case, col_member=col_member))
    deps = ms.get("param_dependencies") or []  # apply_param_dependency! (KiDUtils.jl:384-391)
    overrides = []
    for col in range(nc):
        ov = {name: float(u[p, col_member[col]]) for p, name in enumerate(u_names)}
        for dep in deps:
            ov[dep["dependant"]] = ov.get(dep["base"], base.meta["toml"][dep["base"]]) * dep["ratio"]
        ov["SB2006_cloud_gamma_distribution_parameter"] = float(cases[col_case[col]]["k"])
        overrides.append(ov)
    return M.with_parameter_sets(ens, overrides, column_to_set=np.arange(nc))


def col_sed_calibration_case(FT, model_settings: dict, cases: list[dict], u: np.ndarray, u_names: list[str]) -> M.Case:
    """run_KiD_col_sed_multiple_cases + run_KiD_col_sed (calibration/KiDUtils.jl:155-236) for a whole ensemble: column =
    member * n_cases + case; a case is (qt, Nd, k), the initial condition is initial_condition_0d at every level and k is
    also SB2006's cloud ν, so every (member, case) pair has its own parameter set."""
    u = np.atleast_2d(np.asarray(u, dtype=np.float64))
    if u.shape[0] != len(u_names):
        raise ValueError("u must be [n_params, n_ensemble]")
    ms = dict(model_settings)
    n_ens, n_cases = u.shape[1], len(cases)
    nc = n_ens * n_cases
    col_case, col_member = np.arange(nc) % n_cases, np.arange(nc) // n_cases
    kw = {k: ms[k] for k in ("precipitation_choice", "rain_formation_choice", "sedimentation_choice", "rhod", "z_min",
                             "z_max", "n_elem", "dt", "t_end", "precip_sinks", "toml_overrides") if k in ms}
    base = M.col_sed_case(FT, nc=nc, qt=[cases[i]["qt"] for i in col_case], prescribed_Nd=[cases[i]["Nd"] for i in col_case],
                          k=cases[0]["k"], **kw)
    base.meta.update(cases=cases, col_case=col_case, col_member=col_member)
    deps = ms.get("param_dependencies") or []
    overrides = []
    for col in range(nc):
        ov = {name: float(u[p, col_member[col]]) for p, name in enumerate(u_names)}
        for dep in deps:
            ov[dep["dependant"]] = ov.get(dep["base"], base.meta["toml"][dep["base"]]) * dep["ratio"]
        ov["SB2006_cloud_gamma_distribution_parameter"] = float(cases[col_case[col]]["k"])
        overrides.append(ov)
    # the gamma initial condition depends on k as well: rebuild the state of the columns whose case has another k
    ks = np.array([cases[i]["k"] for i in col_case], dtype=np.float64)
    if np.unique(ks).size > 1:
        full = M.col_sed_case(FT, nc=nc, qt=[cases[i]["qt"] for i in col_case],
                              prescribed_Nd=[cases[i]["Nd"] for i in col_case], k=ks, **kw)
        base.Y0, base.ρ_dry, base.T = full.Y0, full.ρ_dry, full.T
    return M.with_parameter_sets(base, overrides, column_to_set=np.arange(nc))


def run_ensembles(u, u_names, config: dict, n_ensemble: int | None = None, FT=np.float64) -> list:
    """calibration/OptimizationUtils.jl:197-211 without the ReferenceStatistics transform (a host-side linear map):
    returns [G(u[:, i]) for i in 1:n_ensemble], each the concatenation over the cases (KiDUtils.jl:45-81)."""
    u = np.atleast_2d(np.asarray(u, dtype=np.float64))
    n_ensemble = u.shape[1] if n_ensemble is None else n_ensemble
    ms = config["model"]
    cases = config["observations"]["cases"]
    variables = config["observations"]["data_names"]
    fcfg = ms["filter"]
    t_calib = cases[0].get("t_cal", ms["t_calib"])
    if any(not np.array_equal(c.get("t_cal", ms["t_calib"]), t_calib) for c in cases):
        raise NotImplementedError("per-case t_cal: run the cases with different calibration times in separate calls")
    filt = make_filter_props(fcfg["nz_unfiltered"], t_calib, apply=fcfg["apply"],
                             nz_per_filtered_cell=fcfg.get("nz_per_filtered_cell"),
                             nt_per_filtered_cell=fcfg.get("nt_per_filtered_cell", 1))
    if ms.get("model", "KiD") == "Box":
        case = box_calibration_case(FT, ms, cases, u[:, :n_ensemble], list(u_names))
    elif ms.get("model", "KiD") == "KiD_col_sed":
        case = col_sed_calibration_case(FT, ms, cases, u[:, :n_ensemble], list(u_names))
    else:
        case = kid_calibration_case(FT, ms, cases, u[:, :n_ensemble], list(u_names))
    G = solve_gvector(case, variables, t_calib, filt, t_ini=float(ms.get("t_ini", 0.0)))
    n_cases = len(cases)
    return [G[m * n_cases:(m + 1) * n_cases].reshape(-1) for m in range(n_ensemble)]


def run_dyn_model(u, u_names, config: dict, FT=np.float64) -> np.ndarray:
    """calibration/KiDUtils.jl:10-43 with RS = nothing: the G vector of one parameter vector."""
    return run_ensembles(np.asarray(u, dtype=np.float64).reshape(-1, 1), u_names, config, 1, FT=FT)[0]


# ------------------------------------------------------------------ G-vector bookkeeping and diagnostics (host, small)
def get_numbers_from_config(config: dict) -> dict:
    """calibration/HelperFuncs.jl:49-68."""
    ms, cases = config["model"], config["observations"]["cases"]
    f = ms["filter"]
    n_heights = list(f["nz_filtered"] if f["apply"] else f["nz_unfiltered"])
    n_default = len(ms["t_calib"]) - 1 if f["apply"] else len(ms["t_calib"])
    n_times = [(len(c["t_cal"]) - 1 if f["apply"] else len(c["t_cal"])) if "t_cal" in c else n_default for c in cases]
    assert len(n_heights) == len(config["observations"]["data_names"])
    return dict(n_cases=len(cases), n_heights=n_heights, n_times=n_times)


def get_case_i_vec(vec: np.ndarray, i: int, n_single_case) -> np.ndarray:
    """calibration/HelperFuncs.jl:70-72 (i is 0-based here)."""
    a = int(np.sum(n_single_case[:i]))
    return np.asarray(vec)[a:a + int(n_single_case[i])]


def get_single_case_fields(vec_single_case: np.ndarray, n_heights, n_times: int) -> tuple:
    """calibration/HelperFuncs.jl:74-89: one [n_heights[j], n_times] array per variable."""
    n_single_time = int(np.sum(n_heights))
    if len(vec_single_case) != n_times * n_single_time:
        raise ValueError("Inconsistent vector size and provided numbers!")
    m = np.asarray(vec_single_case).reshape(n_times, n_single_time).T  # Julia's column-major reshape
    offs = np.concatenate([[0], np.cumsum(n_heights)])
    return tuple(m[offs[j]:offs[j + 1], :] for j in range(len(n_heights)))


def _path_integral(vec, config, name: str):
    names = config["observations"]["data_names"]
    assert {"rho", name} <= set(names)
    n = get_numbers_from_config(config)
    i_rho, i_q = names.index("rho"), names.index(name)
    dz = (config["model"]["z_max"] - config["model"]["z_min"]) / n["n_heights"][i_rho]
    n_single = [int(np.sum(n["n_heights"])) * t for t in n["n_times"]]
    out = []
    for i in range(n["n_cases"]):
        f = get_single_case_fields(get_case_i_vec(vec, i, n_single), n["n_heights"], n["n_times"][i])
        out.append(np.sum(f[i_q] * f[i_rho], axis=0, keepdims=True) * dz)
    return out


def liquid_water_path(vec, config):
    """calibration/Diagnostics.jl:11-32 [kg/m^2], one [1, n_times] array per case."""
    return _path_integral(vec, config, "ql")


def rainwater_path(vec, config):
    """calibration/Diagnostics.jl:44-65."""
    return _path_integral(vec, config, "qr")


def rainrate(vec, config, height: float = 0.0):
    """calibration/Diagnostics.jl:78-103 [mm/hr]: linear interpolation / extrapolation of the rainrate profile to `height`."""
    names = config["observations"]["data_names"]
    assert "rainrate" in names
    n = get_numbers_from_config(config)
    i_rr = names.index("rainrate")
    nh = n["n_heights"][i_rr]
    dz = (config["model"]["z_max"] - config["model"]["z_min"]) / nh
    ind_z1 = min(nh - 1, max(1, int(np.ceil((height - config["model"]["z_min"]) / dz - 0.5))))  # 1-based, as in Julia
    ind_z2 = ind_z1 + 1
    a1 = ((ind_z2 - 0.5) * dz - height) / dz
    a2 = (height - (ind_z1 - 0.5) * dz) / dz
    n_single = [int(np.sum(n["n_heights"])) * t for t in n["n_times"]]
    out = []
    for i in range(n["n_cases"]):
        f = get_single_case_fields(get_case_i_vec(vec, i, n_single), n["n_heights"], n["n_times"][i])
        r = a1 * f[i_rr][ind_z1 - 1, :] + a2 * f[i_rr][ind_z2 - 1, :]
        out.append(np.where(r < 0, 0.0, r))
    return out
