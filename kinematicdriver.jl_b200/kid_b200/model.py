"""Model assembly: the host-side mirror of src/{K1DModel,K2DModel,BoxModel}/ode_utils.jl and
src/Common/ode_utils.jl, on top of libkid_cuda.so.

A `Case` is everything the reference hands to `ODE.ODEProblem(ode_rhs!, Y, tspan, aux)`:
configuration (styles, grid, switches), the flattened parameter structs, the time-invariant
profiles (aux.thermo_variables.{ρ_dry,T}), per-column scalars and the initial state Y.
`apply_case` uploads a Case into any backend with the kid_b200.lib.Handle method surface
(the CUDA library, or — in tests only — the CPU oracle).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import equation_types as ET
from . import initial_condition as IC
from . import parameters as PR
from .lib import (COL_PRESCRIBED_ND, COL_Q_SURF, COL_W1, F32, F64, MODEL_BOX, MODEL_K1D, MODEL_K1D_COL_SED,
                  MODEL_K2D, Handle, KidConfig, make_config)


@dataclass
class TimeStepping:
    """src/Common/ode_utils.jl:12-16"""
    dt: float
    dt_io: float
    t_max: float


@dataclass
class Case:
    cfg: KidConfig
    params: np.ndarray            # KID_NPARAMS table (or nsets x KID_NPARAMS)
    ρ_dry: np.ndarray             # [nz] or [nc, nz]
    T: np.ndarray
    Y0: np.ndarray                # [nc, nvars, nz]
    var_names: tuple
    col_scalars: dict = field(default_factory=dict)   # which -> [nc] values
    column_to_set: np.ndarray | None = None
    z_centers: np.ndarray | None = None
    meta: dict = field(default_factory=dict)
    # initial condition built on the device (kid_init_shipway_hill): {"sounding": 9 values, "p0": [nc], "dry": bool};
    # ρ_dry, T, Y0 and q_surf are then None until `materialise` downloads them
    device_ic: dict | None = None

    @property
    def dtype(self):
        return np.float32 if self.cfg.float_type == F32 else np.float64


def float_enum(FT) -> int:
    return F32 if np.dtype(FT) == np.float32 else F64


def initialise_state(moisture, precip, ip: dict) -> np.ndarray:
    """src/Common/ode_utils.jl:23-54 — Y as a [nvars, nz] array in the reference's field order."""
    names = ET.state_variable_names(moisture, precip)
    return np.stack([np.asarray(ip[n]) for n in names], axis=0)


def make_function_space(FT, z_min, z_max, n_elem):
    """src/K1DModel/ode_utils.jl:9-21 — returns (center z, face z) of the uniform column."""
    faces = np.linspace(float(z_min), float(z_max), int(n_elem) + 1)
    centers = 0.5 * (faces[:-1] + faces[1:])
    return centers.astype(FT), faces.astype(FT)


_K1D_DEFAULTS = dict(
    moisture_choice="EquilibriumMoisture", precipitation_choice="Precipitation1M",
    rain_formation_scheme_choice=None, sedimentation_scheme_choice=None,
    prescribed_Nd=1e8, open_system_activation=False, local_activation=False,
    precip_sources=True, precip_sinks=True, qtot_flux_correction=False,
    z_min=0.0, z_max=2000.0, n_elem=128, dt=1.0, dt_output=30.0, t_ini=0.0, t_end=3600.0,
    w1=2.0, t1=600.0, p0=100000.0, r_dry=0.04e-6, std_dry=1.4, kappa=0.9,
    z_0=0.0, z_1=740.0, z_2=3260.0, rv_0=0.015, rv_1=0.0138, rv_2=0.0024,
    tht_0=297.9, tht_1=297.9, tht_2=312.66, init_sounding="ShipwayHill2012", velocity="ShipwayHill2012",
)


def kid_options(**overrides) -> dict:
    """Defaults of test/experiments/KiD_driver/parse_commandline.jl:10-174."""
    unknown = set(overrides) - set(_K1D_DEFAULTS) - {"FLOAT_TYPE", "toml_overrides"}
    if unknown:
        raise KeyError(f"unknown KiD option(s): {sorted(unknown)}")
    o = dict(_K1D_DEFAULTS)
    o.update(overrides)
    return o


def _toml_from_opts(FT, opts):
    td = PR.override_toml_dict(
        FT=FT, w1=opts["w1"], t1=opts["t1"], p0=opts["p0"], z_0=opts["z_0"], z_1=opts["z_1"], z_2=opts["z_2"],
        rv_0=opts["rv_0"], rv_1=opts["rv_1"], rv_2=opts["rv_2"], tht_0=opts["tht_0"], tht_1=opts["tht_1"],
        tht_2=opts["tht_2"], precip_sources=int(opts["precip_sources"]), precip_sinks=int(opts["precip_sinks"]),
        qtot_flux_correction=int(opts["qtot_flux_correction"]), prescribed_Nd=opts["prescribed_Nd"],
        open_system_activation=int(opts["open_system_activation"]), local_activation=int(opts["local_activation"]),
        r_dry=opts["r_dry"], std_dry=opts["std_dry"], kappa=opts["kappa"])
    if opts.get("rain_formation_scheme_choice") == "SB2006NL":  # run_KiD_simulation.jl:65-69
        td["SB2006_raindrops_terminal_velocity_coeff_aR"] = 6.0
        td["SB2006_raindrops_terminal_velocity_coeff_bR"] = 9.76
        td["SB2006_raindrops_terminal_velocity_coeff_cR"] = 1490.0
    for k, v in (opts.get("toml_overrides") or {}).items():
        if k not in td:
            raise KeyError(f"unknown ClimaParams name {k}")
        td[k] = v
    return td


def k1d_case(FT=np.float64, nc: int = 1, model: int = MODEL_K1D, device: int = 0, **options) -> Case:
    """Everything run_KiD_simulation.jl:13-157 builds before ODE.init, for nc identical columns."""
    opts = kid_options(**options)
    if opts["z_2"] < opts["z_max"]:
        raise ValueError("z_max cannot be higher than z_2")
    if opts["velocity"] != "ShipwayHill2012" or opts["init_sounding"] != "ShipwayHill2012":
        raise NotImplementedError("only the ShipwayHill2012 velocity/sounding is on the device hot path")
    FT = np.dtype(FT).type
    td = _toml_from_opts(FT, opts)
    common_params = PR.create_common_parameters(td)
    kid_params = PR.create_kid_parameters(td)
    thermo_params = PR.create_thermodynamics_parameters(td)
    moisture = ET.get_moisture_type(opts["moisture_choice"], td)
    precip = ET.get_precipitation_type(opts["precipitation_choice"], td,
                                       rain_formation_choice=opts["rain_formation_scheme_choice"],
                                       sedimentation_choice=opts["sedimentation_scheme_choice"])
    nz = int(opts["n_elem"])
    zc, zf = make_function_space(FT, opts["z_min"], opts["z_max"], nz)
    ρ_profile = IC.ρ_ivp(FT, kid_params, thermo_params, opts["init_sounding"])
    ip = IC.initial_condition_1d(FT, common_params, kid_params, thermo_params, ρ_profile, zc, opts["init_sounding"])
    q_surf = IC.init_profile(FT, kid_params, thermo_params, FT(0), opts["init_sounding"])["qv"]
    Y_col = initialise_state(moisture, precip, ip).astype(FT)
    cfg = make_config(
        model=model, moisture=moisture.kid_enum, precip=precip.kid_enum, rain_formation=precip.rain_formation,
        sedimentation=precip.sedimentation, float_type=float_enum(FT), nz=nz, nc=nc,
        precip_sources=int(common_params.precip_sources), precip_sinks=int(common_params.precip_sinks),
        open_system_activation=int(common_params.open_system_activation),
        local_activation=int(common_params.local_activation),
        qtot_flux_correction=int(kid_params.qtot_flux_correction), device=device,
        z_min=float(opts["z_min"]), z_max=float(opts["z_max"]), dt=float(opts["dt"]))
    table = PR.flatten_params(td, precip.rain_formation)
    Y0 = np.broadcast_to(Y_col, (nc,) + Y_col.shape).copy()
    return Case(cfg=cfg, params=table, ρ_dry=ip["ρ_dry"].astype(FT), T=ip["T"].astype(FT), Y0=Y0,
                var_names=ET.state_variable_names(moisture, precip),
                col_scalars={COL_Q_SURF: np.full(nc, q_surf, dtype=FT)}, z_centers=zc,
                meta=dict(opts=opts, toml=td, moisture=moisture, precip=precip, ip=ip,
                          TS=TimeStepping(opts["dt"], opts["dt_output"], opts["t_end"])))


def box_case(FT=np.float64, nc: int = 1, device: int = 0, precipitation_choice="Precipitation2M",
             rain_formation_choice="SB2006", qt=1e-3, Nd=1e8, k=2.0, rhod=1.22, dt=1.0, dt_output=30.0,
             t_ini=0.0, t_end=3600.0, microphys_params=None, precip_sinks=0, ic: str = "host") -> Case:
    """run_box_simulation.jl:12-81 for nc boxes.  qt, Nd may be arrays of length nc (ensemble).
    ic="device": initial_condition_0d of every box is evaluated on the device (kid_init_gamma_0d)."""
    FT = np.dtype(FT).type
    qt_a = np.broadcast_to(np.asarray(qt, dtype=np.float64), (nc,))
    Nd_a = np.broadcast_to(np.asarray(Nd, dtype=np.float64), (nc,))
    td = PR.override_toml_dict(FT=FT, precip_sources=1, precip_sinks=precip_sinks, prescribed_Nd=float(Nd_a[0]))
    for kk, v in (microphys_params or {}).items():
        if kk not in td:
            raise KeyError(f"unknown ClimaParams name {kk}")
        td[kk] = v
    common_params = PR.create_common_parameters(td)
    thermo_params = PR.create_thermodynamics_parameters(td)
    moisture = ET.get_moisture_type("NonEquilibriumMoisture", td)
    precip = ET.get_precipitation_type(precipitation_choice, td, rain_formation_choice=rain_formation_choice)
    names = ET.state_variable_names(moisture, precip)
    Y0 = ρ_dry = T = None
    if ic == "host":
        Y0 = np.zeros((nc, len(names), 1), dtype=FT)
        ρ_dry = np.zeros((nc, 1), dtype=FT)
        T = np.zeros((nc, 1), dtype=FT)
        cache = {}
        for c in range(nc):
            key = (qt_a[c], Nd_a[c])
            if key not in cache:
                cache[key] = IC.initial_condition_0d(FT, thermo_params, qt_a[c], Nd_a[c], k, rhod)
            ip = cache[key]
            Y0[c, :, 0] = [ip[n] for n in names]
            ρ_dry[c, 0], T[c, 0] = ip["ρ_dry"], ip["T"]
    elif ic != "device":
        raise ValueError("ic must be 'host' or 'device'")
    cfg = make_config(
        model=MODEL_BOX, moisture=moisture.kid_enum, precip=precip.kid_enum, rain_formation=precip.rain_formation,
        sedimentation=precip.sedimentation, float_type=float_enum(FT), nz=1, nc=nc,
        precip_sources=int(common_params.precip_sources), precip_sinks=int(common_params.precip_sinks),
        open_system_activation=0, local_activation=0, qtot_flux_correction=0, device=device,
        z_min=0.0, z_max=1.0, dt=float(dt))
    table = PR.flatten_params(td, precip.rain_formation)
    return Case(cfg=cfg, params=table, ρ_dry=ρ_dry, T=T, Y0=Y0, var_names=names,
                col_scalars={COL_PRESCRIBED_ND: Nd_a.astype(FT)}, z_centers=np.array([0.5], dtype=FT),
                meta=dict(toml=td, moisture=moisture, precip=precip, TS=TimeStepping(dt, dt_output, t_end)),
                device_ic=(dict(kind="gamma0d", qt=qt_a.astype(FT), Nd=Nd_a.astype(FT), k=np.full(nc, k, dtype=FT),
                                rhod=np.full(nc, rhod, dtype=FT)) if ic == "device" else None))


def col_sed_case(FT=np.float64, nc: int = 1, device: int = 0, precipitation_choice="Precipitation2M",
                 rain_formation_choice=None, sedimentation_choice=None, qt=1e-3, prescribed_Nd=1e8, k=2.0, rhod=1.0,
                 z_min=0.0, z_max=3000.0, n_elem=60, dt=1.0, dt_output=30.0, t_ini=0.0, t_end=3600.0,
                 precip_sinks=0, toml_overrides=None) -> Case:
    """test/experiments/KiD_col_sed_driver/run_KiD_col_sed_simulation.jl:13-128 (and calibration/KiDUtils.jl:181-236):
    the 1-D column with collisions + sedimentation only (make_rhs_function_col_sed); the initial condition is the box
    model's gamma-distribution state (initial_condition_0d) at EVERY level; qt, prescribed_Nd, k may be arrays of
    length nc (ensemble of cases).  k is also SB2006's cloud ν, so columns with different k get their own parameter set."""
    FT = np.dtype(FT).type
    qt_a = np.broadcast_to(np.asarray(qt, dtype=np.float64), (nc,))
    Nd_a = np.broadcast_to(np.asarray(prescribed_Nd, dtype=np.float64), (nc,))
    k_a = np.broadcast_to(np.asarray(k, dtype=np.float64), (nc,))
    td = PR.override_toml_dict(FT=FT, precip_sources=1, precip_sinks=precip_sinks, prescribed_Nd=float(Nd_a[0]))
    td["SB2006_cloud_gamma_distribution_parameter"] = float(k_a[0])
    for kk, v in (toml_overrides or {}).items():
        if kk not in td:
            raise KeyError(f"unknown ClimaParams name {kk}")
        td[kk] = v
    common_params = PR.create_common_parameters(td)
    kid_params = PR.create_kid_parameters(td)
    thermo_params = PR.create_thermodynamics_parameters(td)
    moisture = ET.get_moisture_type("NonEquilibriumMoisture", td)
    precip = ET.get_precipitation_type(precipitation_choice, td, rain_formation_choice=rain_formation_choice,
                                       sedimentation_choice=sedimentation_choice)
    names = ET.state_variable_names(moisture, precip)
    nz = int(n_elem)
    zc, _ = make_function_space(FT, z_min, z_max, nz)
    Y0 = np.zeros((nc, len(names), nz), dtype=FT)
    ρ_dry = np.zeros((nc, nz), dtype=FT)
    T = np.zeros((nc, nz), dtype=FT)
    cache = {}
    for c in range(nc):
        key = (qt_a[c], Nd_a[c], k_a[c])
        if key not in cache:
            cache[key] = IC.initial_condition_0d(FT, thermo_params, qt_a[c], Nd_a[c], k_a[c], rhod)
        ip = cache[key]
        Y0[c] = np.array([ip[n] for n in names], dtype=FT)[:, None]
        ρ_dry[c], T[c] = ip["ρ_dry"], ip["T"]
    cfg = make_config(
        model=MODEL_K1D_COL_SED, moisture=moisture.kid_enum, precip=precip.kid_enum, rain_formation=precip.rain_formation,
        sedimentation=precip.sedimentation, float_type=float_enum(FT), nz=nz, nc=nc,
        precip_sources=int(common_params.precip_sources), precip_sinks=int(common_params.precip_sinks),
        open_system_activation=0, local_activation=0, qtot_flux_correction=int(kid_params.qtot_flux_correction),
        device=device, z_min=float(z_min), z_max=float(z_max), dt=float(dt))
    case = Case(cfg=cfg, params=PR.flatten_params(td, precip.rain_formation), ρ_dry=ρ_dry, T=T, Y0=Y0, var_names=names,
                col_scalars={COL_PRESCRIBED_ND: Nd_a.astype(FT), COL_Q_SURF: np.zeros(nc, dtype=FT)}, z_centers=zc,
                meta=dict(toml=td, moisture=moisture, precip=precip, TS=TimeStepping(dt, dt_output, t_end),
                          opts=dict(t_ini=t_ini, t_end=t_end, dt=dt, dt_output=dt_output)))
    ks = np.unique(k_a)
    if ks.size > 1:  # one parameter set per distinct k
        case = with_parameter_sets(case, [{"SB2006_cloud_gamma_distribution_parameter": float(x)} for x in ks],
                                   column_to_set=np.searchsorted(ks, k_a))
    return case


def apply_case(handle, case: Case):
    """Upload a Case into a backend handle (order matters: params and profiles before the state)."""
    handle.set_params(case.params, case.column_to_set)
    if case.device_ic is not None and case.Y0 is None:
        if not hasattr(handle, "init_shipway_hill"):
            raise RuntimeError("this case builds its initial condition on the device: materialise(case) it first")
        for which, values in case.col_scalars.items():
            handle.set_column_scalar(which, values)
        if case.device_ic.get("kind") == "gamma0d":
            d = case.device_ic
            handle.init_gamma_0d(d["qt"], d["Nd"], d["k"], d["rhod"])
            return handle
        handle.init_shipway_hill(case.device_ic["sounding"], case.device_ic["p0"], case.device_ic.get("dry", False))
        return handle
    handle.set_profiles(case.ρ_dry, case.T)
    for which, values in case.col_scalars.items():
        handle.set_column_scalar(which, values)
    handle.set_state(case.Y0)
    return handle


def make_handle(case: Case, handle_type=None):
    """A handle of the CUDA library with the case uploaded.  (`handle_type`: tests hand in the CPU oracle's handle class to
    check results; the product never does — there is one backend.)"""
    return apply_case((handle_type or Handle)(case.cfg), case)


def slice_columns(case: Case, a: int, b: int) -> Case:
    """Columns [a, b) of an ensemble case as a case of their own (host arrays are views)."""
    import copy
    cfg = copy.copy(case.cfg)
    cfg.nc = b - a
    scal = {k: np.asarray(v)[a:b] for k, v in case.col_scalars.items()}
    c2s = None if case.column_to_set is None else case.column_to_set[a:b]
    if case.device_ic is not None and case.Y0 is None:
        per_col = {k: np.asarray(v)[a:b] for k, v in case.device_ic.items() if k in ("p0", "qt", "Nd", "k", "rhod")}
        return Case(cfg=cfg, params=case.params, ρ_dry=None, T=None, Y0=None, var_names=case.var_names, col_scalars=scal,
                    column_to_set=c2s, z_centers=case.z_centers, meta=case.meta, device_ic=dict(case.device_ic, **per_col))
    pc = np.ndim(case.ρ_dry) == 2
    return Case(cfg=cfg, params=case.params, ρ_dry=case.ρ_dry[a:b] if pc else case.ρ_dry, T=case.T[a:b] if pc else case.T,
                Y0=case.Y0[a:b], var_names=case.var_names, col_scalars=scal, column_to_set=c2s, z_centers=case.z_centers,
                meta=case.meta, device_ic=case.device_ic)


def take_columns(case: Case, idx: np.ndarray) -> Case:
    """The columns `idx` (any subset, in that order) of an ensemble case as a case of their own — what a rank of a sharded run
    owns when the shards are dealt out in blocks (bench.py: round-robin blocks of the sorted ensemble give every rank the same
    mix of members)."""
    import copy
    idx = np.asarray(idx)
    out = copy.copy(case)
    out.cfg = copy.copy(case.cfg)
    out.cfg.nc = int(idx.size)
    out.col_scalars = {k: np.asarray(v)[idx] for k, v in case.col_scalars.items()}
    out.column_to_set = None if case.column_to_set is None else np.asarray(case.column_to_set)[idx]
    if case.device_ic is not None:
        out.device_ic = dict(case.device_ic, **{k: np.asarray(v)[idx] for k, v in case.device_ic.items()
                                                if k in ("p0", "qt", "Nd", "k", "rhod")})
    if case.Y0 is not None:
        out.Y0 = case.Y0[idx]
        if np.ndim(case.ρ_dry) == 2:
            out.ρ_dry, out.T = case.ρ_dry[idx], case.T[idx]
    if "column_perm" in case.meta:
        out.meta = dict(case.meta, column_perm=np.asarray(case.meta["column_perm"])[idx])
    return out


def reorder_columns(case: Case, perm: np.ndarray) -> Case:
    """The same ensemble with column i of the result = column perm[i] of `case`.  Columns are independent, so any order
    gives the same per-column results; the order decides which members share a warp (32 neighbouring columns)."""
    import copy
    perm = np.asarray(perm)
    assert sorted(perm.tolist()) == list(range(case.cfg.nc)), "perm must be a permutation of the columns"
    out = copy.copy(case)
    out.col_scalars = {k: np.asarray(v)[perm] for k, v in case.col_scalars.items()}
    out.column_to_set = None if case.column_to_set is None else np.asarray(case.column_to_set)[perm]
    if case.device_ic is not None:
        out.device_ic = dict(case.device_ic, p0=np.asarray(case.device_ic["p0"])[perm])
    if case.Y0 is not None:
        out.Y0 = case.Y0[perm]
        if np.ndim(case.ρ_dry) == 2:
            out.ρ_dry, out.T = case.ρ_dry[perm], case.T[perm]
    out.meta = dict(case.meta, column_perm=perm)
    return out


def warp_coherent_order(case: Case) -> np.ndarray:
    """A permutation that stores similar members next to each other.  Ensembles that differ in the sounding (surface pressure
    or first-level dry density), the updraft w1 and the prescribed Nd are sorted into nested quantile bins, profile first
    (nested_bin_order: the measured ranking of what makes lanes diverge); anything else falls back to a Z-order curve through
    whatever distinguishes the columns.  A caller with members in arbitrary order runs `reorder_columns(case, perm)` and
    maps results back with `inverse_permutation(perm)`, or keeps its host arrays as they are and lets the layout kernels
    permute (`PipelinedEnsemble.for_member_order`, kid_set_column_map)."""
    nc = case.cfg.nc
    named = {}
    if COL_W1 in case.col_scalars:
        named["w1"] = np.asarray(case.col_scalars[COL_W1], dtype=np.float64)
    if COL_PRESCRIBED_ND in case.col_scalars:
        named["Nd"] = np.log10(np.maximum(np.asarray(case.col_scalars[COL_PRESCRIBED_ND], dtype=np.float64), 1e-300))
    if case.device_ic is not None and "p0" in case.device_ic:
        named["profile"] = np.asarray(case.device_ic["p0"], dtype=np.float64)
    elif case.ρ_dry is not None and np.ndim(case.ρ_dry) == 2:
        named["profile"] = np.asarray(case.ρ_dry[:, 0], dtype=np.float64)
    if case.column_to_set is not None:
        named["set"] = np.asarray(case.column_to_set, dtype=np.float64)
    named = {k: v for k, v in named.items() if v.shape == (nc,) and np.ptp(v) > 0}
    if not named:
        return np.arange(nc)
    if "set" not in named and "profile" in named and "w1" in named:
        return nested_bin_order(named["profile"], named["w1"], named.get("Nd"))
    keys = list(named.values())
    return morton_order(*keys[:3], bits=10) if len(keys) <= 3 else np.lexsort(tuple(reversed(keys)))


def inverse_permutation(perm: np.ndarray) -> np.ndarray:
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return inv


def materialise(case: Case, handle=None) -> Case:
    """Host copy of a device-built initial condition: downloads ρ_dry, T, the initial state and q_surf so that the case
    can be handed to another backend (the test oracle) or sliced on the host."""
    if case.Y0 is not None:
        return case
    import copy
    h = handle or make_handle(case)
    out = copy.copy(case)
    out.ρ_dry, out.T = h.get_profiles()
    out.Y0 = h.get_state()
    out.col_scalars = dict(case.col_scalars)
    if "q_surf" in case.device_ic:
        out.col_scalars[COL_Q_SURF] = np.full(case.cfg.nc, case.device_ic["q_surf"], dtype=case.dtype)
    return out


def _host_member_ic(job):
    FT, opts = job
    cs = k1d_case(FT, nc=1, **opts)
    return cs.ρ_dry, cs.T, cs.Y0[0], cs.col_scalars[COL_Q_SURF][0]


def materialise_on_host(case: Case) -> Case:
    """The same, without a GPU: the hydrostatic profile and initial state of every member from the host restatement
    (initial_condition.py: one ρ_ivp solve per member — small samples only; the CPU arm of bench.py)."""
    if case.Y0 is not None:
        return case
    import copy
    opts = {k: v for k, v in case.meta["opts"].items() if k in _K1D_DEFAULTS}
    FT = case.dtype
    p0 = np.asarray(case.device_ic["p0"], dtype=np.float64)
    Nd = np.asarray(case.col_scalars[COL_PRESCRIBED_ND], dtype=np.float64)
    jobs = [(FT, {**opts, "p0": float(p0[c])}) for c in range(case.cfg.nc)]
    workers = min(len(jobs), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 1)
    if workers > 1 and len(jobs) >= 16:
        import concurrent.futures as cf
        import multiprocessing as mp
        with cf.ProcessPoolExecutor(workers, mp_context=mp.get_context("fork")) as ex:
            res = list(ex.map(_host_member_ic, jobs, chunksize=max(1, len(jobs) // (4 * workers))))
    else:
        res = [_host_member_ic(j) for j in jobs]
    ρ, T, Y, q = map(list, zip(*res))
    out = copy.copy(case)
    out.ρ_dry, out.T, out.Y0 = np.asarray(ρ, dtype=FT), np.asarray(T, dtype=FT), np.asarray(Y, dtype=FT)
    if "N_aer" in case.var_names:  # IC N_aer = prescribed_Nd·ρ_dry/1.225 (initial_condition.jl:171)
        out.Y0[:, case.var_names.index("N_aer"), :] = (Nd[:, None] * out.ρ_dry.astype(np.float64) / 1.225).astype(FT)
    out.col_scalars = dict(case.col_scalars)
    out.col_scalars[COL_Q_SURF] = np.asarray(q, dtype=FT)
    return out


def morton_order(*keys: np.ndarray, bits: int = 10) -> np.ndarray:
    """Permutation that sorts members along a Z-order curve through the unit cube of their (normalised) keys:
    neighbouring members are close in EVERY key, not only in the first one of a lexicographic sort."""
    code = np.zeros(keys[0].shape, dtype=np.uint64)
    q = []
    for k in keys:
        k = np.asarray(k, dtype=np.float64)
        span = k.max() - k.min()
        q.append(np.minimum(((k - k.min()) / (span if span > 0 else 1.0) * (1 << bits)).astype(np.uint64), (1 << bits) - 1))
    n = len(keys)
    for b in range(bits):
        for d, qd in enumerate(q):
            code |= ((qd >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * n + (n - 1 - d))
    return np.argsort(code, kind="stable")


def nested_bin_order(primary: np.ndarray, secondary: np.ndarray, tertiary: np.ndarray | None = None) -> np.ndarray:
    """Permutation that sorts members into quantile bins of `primary`, inside a bin into quantile bins of `secondary`, and
    inside those by `tertiary`.  Measured on B200 with the KiD ensemble of SURVEY 8(d) (scripts/order_probe.py): the
    surface pressure (it shifts the whole saturation profile, hence the cloud-base LEVEL) matters most for which branch a lane
    takes, then the updraft w1, then the droplet number: binning p0 finely beats the Z-order curve through all three keys by
    5 % at 1 M members (10.32 vs 10.86 ms/step) and by 9 % for a 32,768-member pipeline chunk.  About 64 members per
    (primary, secondary) cell, four times as many primary as secondary bins."""
    n = primary.size
    if n < 2048:  # too few members for nested bins to resolve the secondary keys: Z-order curve through all of them
        return morton_order(*[k for k in (secondary, tertiary, primary) if k is not None])
    cells = max(1, n // 64)
    b1 = max(1, int(round(np.sqrt(4.0 * cells))))
    b2 = max(1, cells // b1)

    def bins(x, nb):
        r = np.argsort(np.argsort(x, kind="stable"), kind="stable")
        return np.minimum(r * nb // max(n, 1), nb - 1)
    last = tertiary if tertiary is not None else np.zeros(n)
    return np.lexsort((last, bins(secondary, b2), bins(primary, b1)))


# ------------------------------------------------------------------ ensembles
def splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based hash used for the deterministic synthetic ensembles."""
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, c: np.ndarray, k: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) ^ (np.asarray(c, dtype=np.uint64) * np.uint64(8) + np.uint64(k)))
    return (h >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def k1d_ensemble_case(FT=np.float32, nc: int = 1024, seed: int = 20240601, n_profiles: int = 16,
                      column_offset: int = 0, device: int = 0, order: str = "hash", ic: str = "host", **options) -> Case:
    """Synthetic calibration-style ensemble of SURVEY §8(d): member c perturbs w1, prescribed Nd and the
    surface pressure p0.  ic="host": p0 is quantised to n_profiles distinct soundings so the host-side IVP stays cheap;
    ic="device": every member has its own p0 = 1e5·(0.98 + 0.04·u_2) exactly as §8(d) defines it, and the hydrostatic
    profile + initial state of each column are built on the device (kid_init_shipway_hill).
    order: "hash" (member index order), "sorted" (lexicographic by sounding, w1, Nd) or "morton" (Z-order through
    (w1, Nd, p0): neighbouring columns are similar in all three)."""
    FT = np.dtype(FT).type
    c = np.arange(column_offset, column_offset + nc)
    if ic == "device":
        return _k1d_ensemble_case_device_ic(FT, c, seed, device, order, options)
    if ic != "host":
        raise ValueError("ic must be 'host' or 'device'")
    # SURVEY §8(d) proposes w1 in [1, 3]; the reference formulation itself (closed lid: top SetValue(0) flux for
    # N_liq/N_rai/ρq_rai under an updraft) grows without bound at the domain top for w1 >~ 2.5 and overflows
    # Float32, so the synthetic ensemble keeps w1 in [1, 2.2]
    w1 = 2.0 * (0.5 + 0.6 * uniform01(seed, c, 0))
    Nd = 10.0 ** (7.0 + 1.5 * uniform01(seed, c, 1))
    u2 = uniform01(seed, c, 2)
    if n_profiles is None:  # every member its own p0 (one host IVP per member: small ensembles only)
        p0_exact = 1e5 * (0.98 + 0.04 * u2)
        if order in ("sorted", "morton"):
            perm = morton_order(w1, np.log10(Nd), p0_exact)
            w1, Nd, p0_exact = w1[perm], Nd[perm], p0_exact[perm]
        elif order != "hash":
            raise ValueError("order must be 'hash', 'sorted' or 'morton'")
        p0_levels, pidx, n_profiles, order = p0_exact, np.arange(nc), nc, "done"
    else:
        pidx = np.minimum((u2 * n_profiles).astype(np.int64), n_profiles - 1)
        p0_levels = 1e5 * (0.98 + 0.04 * (np.arange(n_profiles) + 0.5) / n_profiles)
    if order == "sorted":
        # the order of ensemble members is arbitrary: placing similar members (same sounding, similar updraft) in
        # neighbouring columns keeps the 32 lanes of a warp on the same microphysics branches
        perm = np.lexsort((Nd, w1, pidx))
        w1, Nd, pidx = w1[perm], Nd[perm], pidx[perm]
    elif order not in ("hash", "done"):
        raise ValueError("order must be 'hash' or 'sorted'")
    base = None
    bank_ρ, bank_T, bank_Y, bank_q = [], [], [], []
    for p0 in p0_levels:
        cs = k1d_case(FT, nc=1, device=device, **{**options, "p0": float(p0)})
        base = base or cs
        bank_ρ.append(cs.ρ_dry); bank_T.append(cs.T); bank_Y.append(cs.Y0[0]); bank_q.append(cs.col_scalars[COL_Q_SURF][0])
    bank_ρ, bank_T, bank_Y, bank_q = map(np.asarray, (bank_ρ, bank_T, bank_Y, bank_q))
    Y0 = bank_Y[pidx].astype(FT)
    names = base.var_names
    if "N_aer" in names:  # IC N_aer = prescribed_Nd·ρ_dry/1.225 (initial_condition.jl:171)
        Y0[:, names.index("N_aer"), :] = (Nd[:, None] * bank_ρ[pidx].astype(np.float64) / 1.225).astype(FT)
    cfg = base.cfg
    cfg.nc = nc
    return Case(cfg=cfg, params=base.params, ρ_dry=bank_ρ[pidx].astype(FT), T=bank_T[pidx].astype(FT), Y0=Y0,
                var_names=names,
                col_scalars={COL_Q_SURF: bank_q[pidx].astype(FT), COL_W1: w1.astype(FT),
                             COL_PRESCRIBED_ND: Nd.astype(FT)},
                z_centers=base.z_centers, meta=dict(base.meta, seed=seed, n_profiles=n_profiles, order=order))


def _k1d_ensemble_case_device_ic(FT, c, seed, device, order, options) -> Case:
    nc = c.size
    w1 = 2.0 * (0.5 + 0.6 * uniform01(seed, c, 0))
    Nd = 10.0 ** (7.0 + 1.5 * uniform01(seed, c, 1))
    p0 = 1e5 * (0.98 + 0.04 * uniform01(seed, c, 2))
    perm = np.arange(nc)
    if order in ("sorted", "morton"):
        perm = nested_bin_order(p0, w1, np.log10(Nd)) if order == "sorted" else morton_order(w1, np.log10(Nd), p0)
        w1, Nd, p0 = w1[perm], Nd[perm], p0[perm]
    elif order != "hash":
        raise ValueError("order must be 'hash', 'sorted' or 'morton'")
    base = k1d_case(FT, nc=1, device=device, **options)  # parameters, grid, q_surf (p0-independent)
    o = base.meta["opts"]
    import copy
    cfg = copy.copy(base.cfg)
    cfg.nc = nc
    sounding = [o[k] for k in ("z_0", "z_1", "z_2", "rv_0", "rv_1", "rv_2", "tht_0", "tht_1", "tht_2")]
    return Case(cfg=cfg, params=base.params, ρ_dry=None, T=None, Y0=None, var_names=base.var_names,
                col_scalars={COL_W1: w1.astype(FT), COL_PRESCRIBED_ND: Nd.astype(FT)}, z_centers=base.z_centers,
                meta=dict(base.meta, seed=seed, order=order, p0=p0, column_perm=perm),  # column i = member c[column_perm[i]]
                device_ic=dict(sounding=sounding, p0=p0.astype(FT), dry=False,
                               q_surf=float(base.col_scalars[COL_Q_SURF][0])))


def box_ensemble_case(FT=np.float64, nc: int = 1024, seed: int = 20240601, n_ic: int = 64, column_offset: int = 0,
                      device: int = 0, **options) -> Case:
    """Synthetic box ensemble of SURVEY §8(d): qt = 1e-3·(0.5+1.5u0), Nd = 10^(7.5+u1), quantised to
    n_ic x n_ic distinct initial conditions."""
    c = np.arange(column_offset, column_offset + nc)
    iq = np.minimum((uniform01(seed, c, 0) * n_ic).astype(np.int64), n_ic - 1)
    iN = np.minimum((uniform01(seed, c, 1) * n_ic).astype(np.int64), n_ic - 1)
    qt = 1e-3 * (0.5 + 1.5 * (iq + 0.5) / n_ic)
    Nd = 10.0 ** (7.5 + (iN + 0.5) / n_ic)
    return box_case(FT, nc=nc, device=device, qt=qt, Nd=Nd, **options)


def custom_case(FT, model: int, moisture_choice: str, precipitation_choice: str, ip: dict, *, z_min=0.0, z_max=1.0,
                dt=1.0, nc: int = 1, rain_formation_choice=None, sedimentation_choice=None, q_surf=0.0, device: int = 0,
                **toml_kw) -> Case:
    """A Case from a hand-written initial-profile dict `ip` of per-level arrays (or scalars for one
    level) — the `_ip` NamedTuple fixtures of the reference's unit tests
    (test/unit_tests/common_unit_test.jl:163-197)."""
    FT = np.dtype(FT).type
    td = PR.override_toml_dict(FT=FT, **toml_kw)
    common_params = PR.create_common_parameters(td)
    kid_params = PR.create_kid_parameters(td)
    moisture = ET.get_moisture_type(moisture_choice, td)
    precip = ET.get_precipitation_type(precipitation_choice, td, rain_formation_choice=rain_formation_choice,
                                       sedimentation_choice=sedimentation_choice)
    ipa = {k: np.atleast_1d(np.asarray(v, dtype=FT)) for k, v in ip.items()}
    nz = ipa["ρq_tot"].size
    names = ET.state_variable_names(moisture, precip)
    Y_col = np.stack([np.broadcast_to(ipa.get(n, np.zeros(nz, FT)), (nz,)) for n in names]).astype(FT)
    cfg = make_config(
        model=model, moisture=moisture.kid_enum, precip=precip.kid_enum, rain_formation=precip.rain_formation,
        sedimentation=precip.sedimentation, float_type=float_enum(FT), nz=nz, nc=nc,
        precip_sources=int(common_params.precip_sources), precip_sinks=int(common_params.precip_sinks),
        open_system_activation=int(common_params.open_system_activation),
        local_activation=int(common_params.local_activation),
        qtot_flux_correction=int(kid_params.qtot_flux_correction), device=device,
        z_min=float(z_min), z_max=float(z_max), dt=float(dt))
    zc, _ = make_function_space(FT, z_min, z_max, nz)
    return Case(cfg=cfg, params=PR.flatten_params(td, precip.rain_formation),
                ρ_dry=np.broadcast_to(ipa["ρ_dry"], (nz,)).astype(FT), T=np.broadcast_to(ipa["T"], (nz,)).astype(FT),
                Y0=np.broadcast_to(Y_col, (nc,) + Y_col.shape).copy(), var_names=names,
                col_scalars={COL_Q_SURF: np.full(nc, q_surf, dtype=FT)}, z_centers=zc,
                meta=dict(toml=td, moisture=moisture, precip=precip))


def with_parameter_sets(case: Case, overrides: list[dict], column_to_set=None) -> Case:
    """Calibration-style ensemble: `overrides[s]` maps ClimaParams names to the values of parameter set s
    (what calibration/KiDUtils.jl:378-382 `update_parameters!` writes into the toml dict per ensemble member);
    column c uses set column_to_set[c] (default: round-robin)."""
    td0 = case.meta["toml"]
    rf = case.meta["precip"].rain_formation
    tables = []
    for ov in overrides:
        td = dict(td0)
        for k, v in ov.items():
            if k not in td:
                raise KeyError(f"unknown ClimaParams name {k}")
            td[k] = v
        tables.append(PR.flatten_params(td, rf))
    nc = case.cfg.nc
    idx = np.arange(nc) % len(overrides) if column_to_set is None else np.asarray(column_to_set)
    new = Case(cfg=case.cfg, params=np.stack(tables), ρ_dry=case.ρ_dry, T=case.T, Y0=case.Y0, var_names=case.var_names,
               col_scalars=dict(case.col_scalars), column_to_set=idx.astype(np.int32), z_centers=case.z_centers,
               meta=dict(case.meta, overrides=overrides))
    return new


def split_by_parameter_set(case: Case):
    """Yield (column indices, single-set Case) per parameter set — lets a backend that handles one set per handle
    (the test oracle) run a multi-set ensemble."""
    import copy
    for s in range(case.params.shape[0]):
        cols = np.where(case.column_to_set == s)[0]
        if cols.size == 0:
            continue
        cfg = copy.copy(case.cfg)
        cfg.nc = int(cols.size)
        pc = case.ρ_dry.ndim == 2
        yield cols, Case(cfg=cfg, params=case.params[s], ρ_dry=case.ρ_dry[cols] if pc else case.ρ_dry,
                         T=case.T[cols] if pc else case.T, Y0=case.Y0[cols], var_names=case.var_names,
                         col_scalars={k: np.asarray(v)[cols] for k, v in case.col_scalars.items()},
                         z_centers=case.z_centers, meta=case.meta)
