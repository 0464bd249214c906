"""Initial conditions (host side, runs once per simulation).

Mirrors src/Common/initial_condition.jl:50-84 (ShipwayHill2012 sounding), :100-149
(hydrostatic ρ(z) IVP), :155-231 (initial_condition_1d), :237-308 (initial_condition_0d)
and src/Common/pysdm_functions.jl:8-63.  Pinned by tests/test_initial_condition.py against
the PySDM golden profiles the reference ships in test/data_utils.jl.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.integrate import solve_ivp
from scipy.special import gammaincc


# ------------------------------------------------------------- pysdm_functions.jl
def SDM_ρ_dry_of_ρ(ρ, q_vap):
    r_vap = q_vap / (1 - q_vap)
    return ρ / (1 + r_vap)


def SDM_ρ_of_ρ_dry(ρ_dry, q_vap):
    r_vap = q_vap / (1 - q_vap)
    return ρ_dry * (1 + r_vap)


def SDM_ρ_dry(tp, p, q_vap, θ_std):
    r_vap = q_vap / (1 - q_vap)
    molmass_ratio = tp.Rv_over_Rd
    return p * (1 - 1 / (1 + 1 / molmass_ratio / r_vap)) / ((p / tp.MSLP) ** (tp.R_d / tp.cp_d) * tp.R_d * θ_std)


def SDM_θ_dry(tp, θ, q_vap):
    r_vap = q_vap / (1 - q_vap)
    return θ * (1 + r_vap * tp.Rv_over_Rd) ** (tp.R_d / tp.cp_d)


def SDM_T(tp, θ_dry, ρ_dry):
    return θ_dry * (ρ_dry * θ_dry / tp.MSLP * tp.R_d) ** (tp.R_d / tp.cp_d / (1 - tp.R_d / tp.cp_d))


def SDM_p(tp, ρ_dry, T, q_vap):
    r_vap = q_vap / (1 - q_vap)
    return ρ_dry * (1 + r_vap) * (tp.R_v / (1 / r_vap + 1) + tp.R_d / (1 + r_vap)) * T


# --------------------------------------------------------- initial_condition.jl
def init_profile_shipwayhill(FT, kid_params, tp, z, dry=False):
    """src/Common/initial_condition.jl:50-84"""
    kp = kid_params
    z_0, z_1, z_2 = FT(kp.z_0), FT(kp.z_1), FT(kp.z_2)
    rv_0, rv_1, rv_2 = FT(kp.rv_0), FT(kp.rv_1), FT(kp.rv_2)
    θ_0, θ_1, θ_2 = FT(kp.tht_0), FT(kp.tht_1), FT(kp.tht_2)
    if dry:
        rv_0 = rv_1 = rv_2 = FT(0)
    z = FT(z)
    rv = rv_0 + (rv_1 - rv_0) / (z_1 - z_0) * (z - z_0) if z < z_1 else rv_1 + (rv_2 - rv_1) / (z_2 - z_1) * (z - z_1)
    qv = FT(rv / (1 + rv))
    qv_0 = FT(rv_0 / (1 + rv_0))
    θ_std = θ_0 if z < z_1 else θ_1 + (θ_2 - θ_1) / (z_2 - z_1) * (z - z_1)
    p_0 = kp.p0
    if qv_0 > 0:
        SDM_ρ_dry_0 = SDM_ρ_dry(tp, p_0, qv_0, θ_0)
    else:  # limit r_vap -> 0 of the expression in SDM_ρ_dry (Julia evaluates 1/0 = Inf there)
        SDM_ρ_dry_0 = p_0 / ((p_0 / tp.MSLP) ** (tp.R_d / tp.cp_d) * tp.R_d * θ_0)
    SDM_ρ_0 = SDM_ρ_of_ρ_dry(SDM_ρ_dry_0, qv_0)
    return dict(qv=qv, θ_std=FT(θ_std), ρ_0=FT(SDM_ρ_0), z_0=z_0, z_2=z_2)


def init_profile(FT, kid_params, tp, z, init_sounding="ShipwayHill2012", dry=False):
    if init_sounding == "ShipwayHill2012":
        return init_profile_shipwayhill(FT, kid_params, tp, z, dry=dry)
    if init_sounding == "Jouan2020":
        raise NotImplementedError("Jouan2020 sounding is outside the device hot path (SURVEY §2 row 13)")
    raise ValueError("Unrecognized initial condition sounding option. Supported options are: Jouan2020 or ShipwayHill2012")


def dρ_dz(ρ, kid_params, tp, z, dry=False):
    """src/Common/initial_condition.jl:100-130 (evaluated in Float64)"""
    init = init_profile_shipwayhill(np.float64, kid_params, tp, z, dry=dry)
    θ_std, q_vap = float(init["θ_std"]), float(init["qv"])
    g, R_d, R_v, cp_d, cp_v = float(tp.grav), float(tp.R_d), float(tp.R_v), float(tp.cp_d), float(tp.cp_v)
    θ_dry = SDM_θ_dry(tp, θ_std, q_vap)
    ρ_dry = SDM_ρ_dry_of_ρ(ρ, q_vap)
    T = SDM_T(tp, θ_dry, ρ_dry)
    if q_vap > 0:
        p = SDM_p(tp, ρ_dry, T, q_vap)
        r_vap = q_vap / (1 - q_vap)
        R_m = R_v / (1 / r_vap + 1) + R_d / (1 + r_vap)
        cp_m = cp_v / (1 / r_vap + 1) + cp_d / (1 + r_vap)
    else:
        R_m, cp_m = R_d, cp_d
        p = ρ_dry * R_d * T
    ρ_SDM = p / R_m / T
    return g / T * ρ_SDM * (R_m / cp_m - 1) / R_m


class DensityProfile:
    """ρ_profile(z): the dense ODE solution returned by ρ_ivp (initial_condition.jl:135-149)."""

    def __init__(self, sol, z_0, z_max):
        self._sol, self.z_0, self.z_max = sol, z_0, z_max

    def __call__(self, z):
        return self._sol.sol(np.asarray(z, dtype=np.float64))[0]


def ρ_ivp(FT, kid_params, tp, init_sounding="ShipwayHill2012", dry=False):
    init_surface = init_profile(np.float64, kid_params, tp, 0.0, init_sounding, dry=dry)
    ρ_0, z_0, z_max = float(init_surface["ρ_0"]), float(init_surface["z_0"]), float(init_surface["z_2"])
    sol = solve_ivp(lambda z, ρ: [dρ_dz(ρ[0], kid_params, tp, z, dry=dry)], (z_0, z_max), [ρ_0],
                    method="DOP853", rtol=1e-10, atol=1e-10, dense_output=True)
    return DensityProfile(sol, z_0, z_max)


def air_pressure(tp, T, ρ, q_tot, q_liq=0.0, q_ice=0.0):
    ratio = tp.R_v / tp.R_d
    return tp.R_d * (1 + (ratio - 1) * q_tot - ratio * (q_liq + q_ice)) * ρ * T


def initial_condition_1d(FT, common_params, kid_params, tp, ρ_profile, z, init_sounding="ShipwayHill2012", dry=False):
    """src/Common/initial_condition.jl:155-231 — returns a dict of per-level arrays (z may be an array)."""
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    out = {k: np.zeros(z.shape, dtype=FT) for k in (
        "ρ", "ρ_dry", "T", "p", "θ_liq_ice", "θ_dry", "ρq_tot", "ρq_liq", "ρq_ice", "ρq_rai", "ρq_sno", "ρq_vap",
        "q_tot", "q_liq", "q_ice", "q_rai", "q_sno", "N_liq", "N_ice", "N_rai", "N_aer", "zero")}
    for i, zi in enumerate(z):
        prof = init_profile(FT, kid_params, tp, zi, init_sounding, dry=dry)
        q_vap, θ_std = FT(prof["qv"]), FT(prof["θ_std"])
        ρ = FT(ρ_profile(zi))
        ρ_dry = FT(SDM_ρ_dry_of_ρ(ρ, q_vap))
        θ_dry = FT(SDM_θ_dry(tp, θ_std, q_vap))
        T = FT(SDM_T(tp, θ_dry, ρ_dry))
        q_tot = q_vap
        p = FT(air_pressure(tp, T, ρ, q_tot))
        out["ρ"][i], out["ρ_dry"][i], out["T"][i], out["p"][i] = ρ, ρ_dry, T, p
        out["θ_liq_ice"][i], out["θ_dry"][i] = θ_std, θ_dry
        out["ρq_tot"][i] = FT(q_tot * ρ)
        out["ρq_vap"][i] = FT(q_vap * ρ)
        out["q_tot"][i] = q_tot
        out["N_aer"][i] = FT(common_params.prescribed_Nd * ρ_dry / FT(1.225))
    return out


def initial_condition_0d(FT, tp, qt, Nd, k, ρ_dry):
    """src/Common/initial_condition.jl:237-308 — gamma-distribution IC of the box model."""
    qt, Nd, k, ρ_dry = FT(qt), FT(Nd), FT(k), FT(ρ_dry)
    Lt = FT(ρ_dry * qt / (1 - qt))
    rhow = FT(1000)
    radius_th = FT(40 * 1e-6)
    thrshld = FT(FT(4 / 3) * FT(math.pi) * radius_th ** 3 * rhow / (Lt / Nd / k))
    # SpecialFunctions.gamma_inc(a, x) returns (P, Q); the reference calls gamma_inc(thrshld, k + 1)[2],
    # i.e. a = thrshld, x = k + 1 and takes Q(a, x) — restated literally:
    mass_ratio = FT(gammaincc(float(thrshld), float(k) + 1.0))
    ρq_liq = FT(mass_ratio * Lt)
    ρq_rai = FT(Lt - ρq_liq)
    ρ = FT(ρ_dry + Lt)
    num_ratio = FT(gammaincc(float(thrshld), float(k)))
    N_liq = FT(num_ratio * Nd)
    N_rai = FT(Nd - N_liq)
    T = FT(300)
    q_liq, q_rai = FT(ρq_liq / ρ), FT(ρq_rai / ρ)
    return dict(ρ=ρ, ρ_dry=ρ_dry, T=T, p=FT(air_pressure(tp, T, ρ, qt, q_liq, FT(0))),
                ρq_tot=Lt, ρq_liq=ρq_liq, ρq_ice=FT(0), ρq_rai=ρq_rai, ρq_sno=FT(0), ρq_vap=FT(0),
                q_tot=qt, q_liq=q_liq, q_ice=FT(0), q_rai=q_rai, q_sno=FT(0),
                N_liq=N_liq, N_rai=N_rai, N_aer=FT(0), zero=FT(0))
