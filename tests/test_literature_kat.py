"""Literature known-answer tests for the upstream formulas the oracle restates (CloudMicrophysics / Thermodynamics are not
vendored in the reference checkout, so there is no Julia value to compare with — DESIGN.md §4).  Each test pins a formula
or a constant table to a number that exists independently of this repository:

  * raindrop fall speeds: Gunn & Kinzer (1949) measured terminal velocities, which both the Seifert & Beheng (2006) law
    v(D) = a_R - b_R exp(-c_R D) and the Chen et al. (2022) Table B1 fit were constructed to reproduce;
  * the SB2006 bulk (number- and mass-weighted) fall speeds are the size-distribution averages of that single-drop law:
    checked by numerical quadrature, independent of the closed form the code evaluates;
  * SB2006 autoconversion at a hand-evaluated point of the published rate (SB2006 eq. 4 with the Long kernel constant);
  * Khairoutdinov & Kogan (2000) autoconversion 1350 q_c^2.47 N_c^-1.79 (N_c per cm^3) at q_c = 1 g/kg, N_c = 100 cm^-3;
  * saturation vapour pressure over liquid water at 0, 10, 20, 30 degC (Wexler 1976 / WMO tables);
  * ARG2000 activation: limits and monotonicity stated in Abdul-Razzak & Ghan (2000), and an independent numpy restatement
    of the published equations.

All on the CPU oracle (the GPU kernels are tied to the oracle by tests/test_gpu_parity.py)."""
import math

import numpy as np
import pytest
from scipy import integrate, special

from kid_b200 import model as M
from kid_b200 import parameters as PR
from kid_b200.lib import MODEL_BOX, MODEL_K1D
from oracle.oracle import OracleHandle

# Gunn & Kinzer (1949), terminal velocity of water drops in stagnant air at 1013 hPa, 20 degC: D [mm] -> v [m/s]
GUNN_KINZER = {0.5: 2.06, 1.0: 4.03, 2.0: 6.49, 3.0: 8.06, 4.0: 8.83, 5.0: 9.09}


def toml():
    return PR.override_toml_dict(FT=np.float64)


def test_sb2006_single_drop_law_matches_gunn_kinzer():
    td = toml()
    aR, bR, cR = (td["SB2006_raindrops_terminal_velocity_coeff_" + k] for k in ("aR", "bR", "cR"))
    for D_mm, v_gk in GUNN_KINZER.items():
        v = aR - bR * math.exp(-cR * D_mm * 1e-3)
        assert v == pytest.approx(v_gk, rel=0.03), (D_mm, v, v_gk)


def test_chen2022_table_b1_matches_gunn_kinzer():
    """Chen et al. (2022) eq. 19 / Table B1 (rain): v(D) = q(ρ) Σ_i a_i D^b_i exp(-c_i D), D in mm, at ρ = 1.2 kg/m3."""
    td = toml()
    g = lambda k: td["Chen2022_table_B1_" + k + "_coeff"]  # noqa: E731
    rho = 1.2
    q = math.exp(g("q") * rho)
    a = [g("a1"), g("a2"), g("a3") * rho ** g("a3_pow")]
    b = [g("b1") - g("b_rho") * rho, g("b2") - g("b_rho") * rho, g("b3") - g("b_rho") * rho]
    c = [g("c1"), g("c2"), g("c3")]
    for D_mm, v_gk in GUNN_KINZER.items():
        v = q * sum(ai * D_mm ** bi * math.exp(-ci * D_mm) for ai, bi, ci in zip(a, b, c))
        assert v == pytest.approx(v_gk, rel=0.05), (D_mm, v, v_gk)


def rain_point(rho=1.2, rhoq_rai=1e-3, N_rai=1e4, **kw):
    ip = dict(ρ=rho, ρ_dry=rho - 15e-3, T=285.0, ρq_tot=15e-3, ρq_liq=1e-3, ρq_ice=0.0, ρq_rai=rhoq_rai, ρq_sno=0.0,
              N_liq=1e8, N_rai=N_rai, N_aer=1e9)
    cs = M.custom_case(np.float64, MODEL_BOX, "NonEquilibriumMoisture", "Precipitation2M", ip, dt=1.0, **kw)
    return cs, M.make_handle(cs, OracleHandle)


@pytest.mark.parametrize("rhoq_rai,N_rai", [(1e-3, 1e4), (2e-4, 5e3), (5e-3, 2e4)])
def test_sb2006_bulk_fall_speeds_are_size_distribution_averages(rhoq_rai, N_rai):
    """CM2.rain_terminal_velocity (SB2006, eq. 28 of the paper): the number- and mass-weighted means of
    max(0, a_R - b_R exp(-c_R D)) over the exponential distribution n(D) ~ exp(-λ_r D), times sqrt(ρ0/ρ)."""
    rho = 1.2
    cs, h = rain_point(rho, rhoq_rai, N_rai, rain_formation_choice="SB2006NL")  # unlimited distribution: λ_r from x_r only
    td = cs.meta["toml"]
    aR, bR, cR = (td["SB2006_raindrops_terminal_velocity_coeff_" + k] for k in ("aR", "bR", "cR"))
    rho0, rho_w = td["SB2006_reference_air_density"], td["density_liquid_water"]
    x_r = min(max(rhoq_rai / N_rai, td["SB2006_raindrops_min_mass"]), td["SB2006_raindrops_max_mass"])
    lam = (math.pi * rho_w / x_r) ** (1.0 / 3.0)
    v = lambda D: max(0.0, aR - bR * math.exp(-cR * D))  # noqa: E731
    expect = []
    for k in (0, 3):  # number, mass (D^3) weighting
        num = integrate.quad(lambda D: v(D) * D ** k * math.exp(-lam * D), 0, 60 / lam, epsabs=0, epsrel=1e-11, limit=400)[0]
        den = integrate.quad(lambda D: D ** k * math.exp(-lam * D), 0, 60 / lam, epsabs=0, epsrel=1e-11, limit=400)[0]
        expect.append(math.sqrt(rho0 / rho) * num / den)
    got_N, got_q = h.get_aux(0.0, "term_vel_N_rai")[0, 0], h.get_aux(0.0, "term_vel_rai")[0, 0]
    assert got_N == pytest.approx(expect[0], rel=2e-3)
    # the mass-weighted value drops the (tiny) sub-cutoff part of the D^3-weighted integral analytically
    assert got_q == pytest.approx(expect[1], rel=2e-3)
    assert got_q > got_N > 0


def test_chen2022_bulk_rain_velocity_of_the_1m_scheme_is_the_marshall_palmer_average():
    """CM1.terminal_velocity(rain, Chen2022VelTypeRain, ρ, q): mass-weighted mean of the Table B1 single-drop law over
    n(x) ~ exp(-λ x), λ = 1/lambda_inverse(rain) (the fit is applied to the distribution's own size variable x)."""
    rho, q_rai = 1.1, 1e-3
    ip = dict(ρ=rho, ρ_dry=rho * (1 - 12e-3), T=288.0, ρq_tot=12e-3 * rho, ρq_liq=5e-4 * rho, ρq_ice=0.0, ρq_rai=q_rai * rho, ρq_sno=0.0)
    cs = M.custom_case(np.float64, MODEL_BOX, "NonEquilibriumMoisture", "Precipitation1M", ip, sedimentation_choice="Chen2022")
    td = cs.meta["toml"]
    g = lambda k: td["Chen2022_table_B1_" + k + "_coeff"]  # noqa: E731
    qf = math.exp(g("q") * rho)
    a = [g("a1") * qf, g("a2") * qf, g("a3") * qf * rho ** g("a3_pow")]
    b = [g(k) - g("b_rho") * rho for k in ("b1", "b2", "b3")]
    c = [g("c1") * 1e3, g("c2") * 1e3, g("c3") * 1e3]
    a = [ai * 1e3 ** bi for ai, bi in zip(a, b)]  # mm -> m
    # Marshall-Palmer λ of the 1-moment scheme: ρ q = χm m0 n0 Γ(me+Δm+1) λ^-(me+Δm+1) / r0^(me+Δm)
    r0, m0 = td["rain_drop_radius"] if "rain_drop_radius" in td else 1e-3, None
    h = M.make_handle(cs, OracleHandle)
    v_code = h.get_aux(0.0, "term_vel_rai")[0, 0]
    # λ from the Blk1M law of the same distribution: v_blk = χv v0 (λ r0)^-(ve+Δv) Γ(me+ve+Δm+Δv+1)/Γ(me+Δm+1)
    blk = M.make_handle(M.custom_case(np.float64, MODEL_BOX, "NonEquilibriumMoisture", "Precipitation1M", ip), OracleHandle)
    v_blk = blk.get_aux(0.0, "term_vel_rai")[0, 0]
    me, ve, r0 = 3.0, 0.5, 1e-3
    v0 = math.sqrt(8.0 / 3.0 / 0.55 * (1000.0 / rho - 1.0) * td["gravitational_acceleration"] * r0)
    lam = (v0 * special.gamma(me + ve + 1) / special.gamma(me + 1) / v_blk) ** (1.0 / ve) / r0
    v = lambda x: sum(ai * x ** bi * math.exp(-ci * x) for ai, bi, ci in zip(a, b, c))  # noqa: E731
    num = integrate.quad(lambda x: v(x) * x ** 3 * math.exp(-lam * x), 0, 60 / lam, epsabs=0, epsrel=1e-11, limit=400)[0]
    den = integrate.quad(lambda x: x ** 3 * math.exp(-lam * x), 0, 60 / lam, epsabs=0, epsrel=1e-11, limit=400)[0]
    assert v_code == pytest.approx(num / den, rel=1e-6)
    assert 1.0 < v_code < 10.0


def test_sb2006_autoconversion_at_a_hand_evaluated_point():
    """SB2006 eq. 4: dL_r/dt = k_cc/(20 x*) (ν+2)(ν+4)/(ν+1)^2 L_c^2 x_c^2 [1 + Φ_au(τ)/(1-τ)^2] ρ0/ρ with k_cc = 4.44e9,
    x* = 2.6e-10 kg, ν = 2: at L_c = 1 g/m3, N_c = 100 cm^-3, no rain (τ = 0, Φ_au = 0), ρ = ρ0:
    8.538e17 · 8/3 · 1e-6 · 1e-22 = 2.277e-10 kg m^-3 s^-1."""
    rho = 1.225
    ip = dict(ρ=rho, ρ_dry=rho - 10e-3, T=285.0, ρq_tot=10e-3, ρq_liq=1e-3, ρq_ice=0.0, ρq_rai=0.0, ρq_sno=0.0,
              N_liq=1e8, N_rai=0.0, N_aer=1e9)
    cs = M.custom_case(np.float64, MODEL_BOX, "NonEquilibriumMoisture", "Precipitation2M", ip, dt=1.0, precip_sinks=0)
    h = M.make_handle(cs, OracleHandle)
    dL = rho * h.get_aux(0.0, "precip_sources.q_rai")[0, 0]
    assert dL == pytest.approx(2.2769e-10, rel=1e-3)
    # the matching number rates: dN_r = dL/x*, dN_c = -2 dN_r plus cloud self-collection -k_cc (ν+2)/(ν+1) L_c^2 - dN_c,au
    dNr = h.get_aux(0.0, "precip_sources.N_rai")[0, 0]
    assert dNr == pytest.approx(2.2769e-10 / 2.6e-10, rel=1e-3)
    sc = 4.44e9 * 4.0 / 3.0 * 1e-6  # k_cc (ν+2)/(ν+1) L_c^2 at ρ = ρ0: total cloud-number sink of collisions
    dNl = h.get_aux(0.0, "precip_sources.N_liq")[0, 0]
    assert dNl == pytest.approx(-sc, rel=2e-3)


def test_kk2000_autoconversion_reference_point():
    """Khairoutdinov & Kogan (2000) eq. 29: dq_r/dt = 1350 q_c^2.47 N_c^-1.79 (N_c in cm^-3) = 1.38e-8 s^-1 at
    q_c = 1 g/kg, N_c = 100 cm^-3 (ρ = 1 kg/m3)."""
    rho = 1.0
    q_c = 1e-3
    ip = dict(ρ=rho, ρ_dry=rho * (1 - 10e-3), T=285.0, ρq_tot=10e-3 * rho, ρq_liq=q_c * rho, ρq_ice=0.0, ρq_rai=0.0, ρq_sno=0.0)
    cs = M.custom_case(np.float64, MODEL_BOX, "NonEquilibriumMoisture", "Precipitation1M", ip, dt=1.0, rain_formation_choice="KK2000",
                       prescribed_Nd=1e8, precip_sinks=0)
    h = M.make_handle(cs, OracleHandle)
    rate = h.get_aux(0.0, "precip_sources.q_rai")[0, 0]
    assert rate == pytest.approx(1350.0 * q_c ** 2.47 * 100.0 ** -1.79, rel=5e-3)


@pytest.mark.parametrize("T,es", [(273.16, 611.657), (283.15, 1228.1), (293.15, 2338.8), (303.15, 4246.0)])
def test_saturation_vapour_pressure_over_liquid(T, es):
    """TD.saturation_vapor_pressure (Rankine-Kirchhoff form) against tabulated e_s(T) (Wexler 1976; 0.3 % band: the
    constant-cp approximation).  Read off the oracle through the equilibrium condensate: q_vs = q_tot - q_liq."""
    rho, q_tot = 1.2, 40e-3
    ip = dict(ρ=rho, ρ_dry=rho * (1 - q_tot), T=T, ρq_tot=rho * q_tot)
    cs = M.custom_case(np.float64, MODEL_BOX, "EquilibriumMoisture", "NoPrecipitation", ip)
    h = M.make_handle(cs, OracleHandle)
    q_vs = q_tot - h.get_aux(0.0, "q_liq")[0, 0]
    R_v = cs.meta["toml"]["gas_constant_vapor"]
    assert q_vs * rho * R_v * T == pytest.approx(es, rel=3e-3)


# ---------------------------------------------------------------- ARG2000
def arg2000_numpy(td, T, p, q_tot, q_liq, w, N_aer, r_dry, sigma, kappa):
    """Abdul-Razzak & Ghan (2000) eqs. 1-13 for one lognormal κ-Köhler mode, written from the paper: returns (S_max, N_act)."""
    R_d, R_v = td["gas_constant_dry_air"], td["gas_constant_vapor"]
    cp_d, cp_v, cp_l = td["isobaric_specific_heat_dry_air"] if "isobaric_specific_heat_dry_air" in td else R_d / td["adiabatic_exponent_dry_air"], \
        td["isobaric_specific_heat_vapor"], td["isobaric_specific_heat_liquid"]
    g, rho_w = td["gravitational_acceleration"], td["density_liquid_water"]
    M_w, R, sig = td["molar_mass_water"], td["universal_gas_constant"], td["surface_tension_water"]
    K, D = td["thermal_conductivity_of_air"], td["diffusivity_of_water_vapor"]
    T0, LH0 = td["temperature_reference"], td["latent_heat_vaporization_at_reference"]
    T_tr, p_tr = td["temperature_triple_point"], td["pressure_triple_point"]
    L = LH0 + (cp_v - cp_l) * (T - T0)
    dcp = cp_v - cp_l
    p_vs = p_tr * (T / T_tr) ** (dcp / R_v) * math.exp((LH0 - dcp * T0) / R_v * (1 / T_tr - 1 / T))
    eps = R_d / R_v
    R_m = R_d * (1 + (1 / eps - 1) * q_tot - q_liq / eps)
    cp_m = cp_d + (cp_v - cp_d) * q_tot + (cp_l - cp_v) * q_liq
    G = 1.0 / (L / K / T * (L / R_v / T - 1) + R_v * T / D / p_vs) / rho_w
    alpha = L * g * eps / R_m / cp_m / T ** 2 - g / R_m / T
    gamma = R_m * T / eps / p_vs + eps * L ** 2 / cp_m / T / p
    A = 2 * sig * M_w / rho_w / R / T
    S_m = 2 / math.sqrt(kappa) * (A / 3 / r_dry) ** 1.5
    zeta = 2 * A / 3 * math.sqrt(alpha * w / G)
    eta = (alpha * w / G) ** 1.5 / (2 * math.pi * rho_w * gamma * N_aer)
    lns = math.log(sigma)
    f = td["ARG2000_f_coeff_1"] * math.exp(td["ARG2000_f_coeff_2"] * lns ** 2)
    gg = td["ARG2000_g_coeff_1"] + td["ARG2000_g_coeff_2"] * lns
    S_max = 1 / math.sqrt(1 / S_m ** 2 * (f * (zeta / eta) ** td["ARG2000_pow_1"] + gg * (S_m ** 2 / (eta + 3 * zeta)) ** td["ARG2000_pow_2"]))
    u = 2 * math.log(S_m / S_max) / 3 / math.sqrt(2) / lns
    return S_max, N_aer * 0.5 * (1 - special.erf(u))


def activation_rate(w, N_aer=1e8, T=288.0, rho=1.1, dt=1.0, supersat=0.002, **kw):
    """One K1D column whose middle level is slightly supersaturated with no liquid: the oracle's activation source there."""
    td = PR.override_toml_dict(FT=np.float64, **kw)
    tp = PR.create_thermodynamics_parameters(td)
    dcp = tp.cp_v - tp.cp_l
    p_vs = tp.press_triple * (T / tp.T_triple) ** (dcp / tp.R_v) * math.exp((tp.LH_v0 - dcp * tp.T_0) / tp.R_v * (1 / tp.T_triple - 1 / T))
    q_vs = p_vs / (rho * tp.R_v * T)
    q_tot = q_vs * (1 + supersat)
    nz = 5
    ip = dict(ρ=np.full(nz, rho), ρ_dry=np.full(nz, rho * (1 - q_tot)), T=np.full(nz, T), ρq_tot=np.full(nz, rho * q_tot),
              ρq_liq=np.zeros(nz), ρq_ice=np.zeros(nz), ρq_rai=np.zeros(nz), ρq_sno=np.zeros(nz), N_liq=np.zeros(nz),
              N_rai=np.zeros(nz), N_aer=np.full(nz, N_aer))
    cs = M.custom_case(np.float64, MODEL_K1D, "NonEquilibriumMoisture", "Precipitation2M", ip, z_max=100.0, dt=dt, w1=w, t1=600.0,
                       local_activation=1, **kw)
    h = M.make_handle(cs, OracleHandle)
    t = 300.0  # sin(π t / t1) = 1: ρw = w1
    S = h.get_aux(t, "activation_sources.N_liq")[0]
    p = h.get_aux(t, "p")[0, 2]
    return S[2] * dt, td, dict(T=T, p=p, q_tot=q_tot, rho=rho)


def test_arg2000_against_an_independent_restatement_of_the_paper():
    for w_rho in (0.5, 2.0):
        for N_aer in (5e7, 5e8):
            N_act, td, st = activation_rate(w_rho, N_aer)
            _, expect = arg2000_numpy(td, st["T"], st["p"], st["q_tot"], 0.0, w_rho / st["rho"], N_aer, td["r_dry"], td["std_dry"], td["kappa"])
            assert N_act == pytest.approx(expect, rel=1e-6), (w_rho, N_aer)


def test_arg2000_limits_and_monotonicity():
    """ARG2000 §3 / Fig. 1-3: the activated fraction rises monotonically with the updraught, falls with the aerosol number
    and approaches 1 for strong updraughts / small numbers."""
    # (local activation only acts where S_max exceeds the ambient supersaturation: start from a barely saturated parcel)
    fr = [activation_rate(w, 1e8, supersat=1e-6)[0] / 1e8 for w in (0.05, 0.2, 1.0, 5.0)]
    assert all(b > a for a, b in zip(fr, fr[1:]))
    assert 0.0 < fr[0] < fr[-1] <= 1.0
    assert fr[-1] > 0.9
    fN = [activation_rate(0.5, N, supersat=1e-6)[0] / N for N in (1e7, 1e8, 1e9, 1e10)]
    assert all(b < a for a, b in zip(fN, fN[1:]))
