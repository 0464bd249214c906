// ORACLE — test infrastructure only.  PARITY UNPINNED against real Julia output.
//
// kid_physics.hpp: scalar restatement (CPU, C++17, templated on the float type)
// of every third-party formula the reference's hot path calls.  None of these
// packages is vendored under /root/reference and there is no Manifest.toml, so
// the formulas below restate the PUBLISHED algorithms of
//   Thermodynamics.jl   0.15 / 1.0   (Project.toml:35)
//   CloudMicrophysics.jl 0.29 - 0.32 (Project.toml:25)
// anchored on the reference's own call sites, which are cited per function
// (paths relative to the reference checkout).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use this code; the
// product (libkid_cuda.so) never includes or links it.
//
// Every function evaluates in FT exactly like a Julia method specialised on FT
// would (no implicit promotion to double), except the special functions
// Γ(a), Γ(a,x), erf, which SpecialFunctions.jl evaluates in Float64 and rounds.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

#include "../include/kid_cuda.h"

// math shims: <cmath> for float/double; oracle/opcount.hpp adds overloads for its operation-counting scalar (it must be
// included BEFORE this header so that the templates below see them)
namespace kidm {
using std::pow; using std::max; using std::sqrt; using std::exp; using std::cbrt; using std::fabs; using std::min;
using std::log; using std::sin; using std::cos; using std::tgamma; using std::erf; using std::isnan;
}  // namespace kidm

namespace kidoracle {

// ---------------------------------------------------------------- parameters
template <class FT>
struct Prm {
#define KID_X_FIELD(name) FT name;
    KID_PARAM_LIST(KID_X_FIELD)
#undef KID_X_FIELD
    void load(const double* table) {
        int i = 0;
#define KID_X_LOAD(name) name = FT(table[i++]);
        KID_PARAM_LIST(KID_X_LOAD)
#undef KID_X_LOAD
    }
};

struct Switches {
    int moisture, precip, rain_formation, sedimentation;
    bool precip_sources, precip_sinks, open_system_activation, local_activation, qtot_flux_correction;
};

// ------------------------------------------------ special functions (double)
// SpecialFunctions.gamma(a, x): upper incomplete gamma, also for a <= 0
// (used at src/Common/tendency.jl:643-644 through CM2.rain_evaporation).
inline double expint_E1(double x) {
    const double euler = 0.57721566490153286061;
    if (x <= 1.0) {
        double sum = 0.0, term = 1.0;
        for (int k = 1; k < 200; ++k) {
            term *= -x / k;
            double add = -term / k;
            sum += add;
            if (kidm::fabs(add) < 1e-17 * kidm::fabs(sum)) break;
        }
        return -euler - kidm::log(x) + sum;
    }
    // modified Lentz continued fraction
    double b = x + 1.0, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 500; ++i) {
        double an = -double(i) * i;
        b += 2.0;
        d = 1.0 / (an * d + b);
        c = b + an / c;
        double del = c * d;
        h *= del;
        if (kidm::fabs(del - 1.0) < 1e-16) break;
    }
    return h * kidm::exp(-x);
}
inline double gamma_upper_pos(double a, double x) {  // a > 0
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 1000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (kidm::fabs(del) < kidm::fabs(sum) * 1e-17) break;
        }
        double lower = sum * kidm::exp(-x + a * kidm::log(x));
        return kidm::tgamma(a) - lower;
    }
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 1000; ++i) {
        double an = -i * (i - a);
        b += 2.0;
        d = an * d + b;
        if (kidm::fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c;
        if (kidm::fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (kidm::fabs(del - 1.0) < 1e-16) break;
    }
    return kidm::exp(-x + a * kidm::log(x)) * h;
}
inline double gamma_upper(double a, double x) {
    if (a > 0.0) return gamma_upper_pos(a, x);
    if (a == 0.0) return expint_E1(x);
    // Γ(a,x) = (Γ(a+1,x) − x^a e^{−x}) / a
    return (gamma_upper(a + 1.0, x) - kidm::pow(x, a) * kidm::exp(-x)) / a;
}
template <class FT>
inline FT sf_gamma(FT a) { return FT(kidm::tgamma(double(a))); }
template <class FT>
inline FT sf_gamma_inc_upper(FT a, FT x) { return FT(gamma_upper(double(a), double(x))); }
template <class FT>
inline FT sf_erf(FT x) { return FT(kidm::erf(double(x))); }

// ------------------------------------------------------------ Thermodynamics
template <class FT>
struct Thermo {
    const Prm<FT>& P;
    explicit Thermo(const Prm<FT>& p) : P(p) {}
    FT eps() const { return std::numeric_limits<FT>::epsilon(); }

    // TD.gas_constant_air / TD.air_pressure (call sites: Common/tendency.jl:42,62,116,132)
    FT gas_constant_air(FT q_tot, FT q_liq, FT q_ice) const {
        FT ratio = P.td_R_v / P.td_R_d;
        return P.td_R_d * (FT(1) + (ratio - FT(1)) * q_tot - ratio * (q_liq + q_ice));
    }
    FT air_pressure(FT T, FT rho, FT q_tot, FT q_liq = 0, FT q_ice = 0) const {
        return gas_constant_air(q_tot, q_liq, q_ice) * rho * T;
    }
    FT cp_m(FT q_tot, FT q_liq, FT q_ice) const {
        return P.td_cp_d + (P.td_cp_v - P.td_cp_d) * q_tot + (P.td_cp_l - P.td_cp_v) * q_liq +
               (P.td_cp_i - P.td_cp_v) * q_ice;
    }
    FT latent_heat_vapor(FT T) const { return P.td_LH_v0 + (P.td_cp_v - P.td_cp_l) * (T - P.td_T_0); }
    FT latent_heat_sublim(FT T) const { return P.td_LH_s0 + (P.td_cp_v - P.td_cp_i) * (T - P.td_T_0); }
    // TD.latent_heat_fusion (Common/tendency.jl:6,484)
    FT latent_heat_fusion(FT T) const {
        return (P.td_LH_s0 - P.td_LH_v0) + (P.td_cp_l - P.td_cp_i) * (T - P.td_T_0);
    }
    // Clausius-Clapeyron with constant Δcp (TD.saturation_vapor_pressure)
    FT psat_generic(FT T, FT LH_0, FT dcp) const {
        return P.td_p_triple * kidm::pow(T / P.td_T_triple, dcp / P.td_R_v) *
               kidm::exp((LH_0 - dcp * P.td_T_0) / P.td_R_v * (FT(1) / P.td_T_triple - FT(1) / T));
    }
    FT psat_liquid(FT T) const { return psat_generic(T, P.td_LH_v0, P.td_cp_v - P.td_cp_l); }
    FT psat_ice(FT T) const { return psat_generic(T, P.td_LH_s0, P.td_cp_v - P.td_cp_i); }
    // TD.liquid_fraction(param_set, T) for the equilibrium partition
    FT liquid_fraction_T(FT T) const {
        if (T > P.td_T_freeze) return FT(1);
        if (T > P.td_T_icenuc)
            return kidm::pow((T - P.td_T_icenuc) / (P.td_T_freeze - P.td_T_icenuc), P.td_pow_icenuc);
        return FT(0);
    }
    FT psat_mixed(FT T) const {
        FT lam = liquid_fraction_T(T);
        FT LH_0 = lam * P.td_LH_v0 + (FT(1) - lam) * P.td_LH_s0;
        FT dcp = lam * (P.td_cp_v - P.td_cp_l) + (FT(1) - lam) * (P.td_cp_v - P.td_cp_i);
        return psat_generic(T, LH_0, dcp);
    }
    // TD.q_vap_saturation(thermo_params, T, ρ) (Common/tendency.jl:396)
    FT q_vap_saturation(FT T, FT rho) const { return psat_mixed(T) / (rho * P.td_R_v * T); }
    // TD.condensate_partition(thermo_params, T, ρ, q_tot) -> (q_liq, q_ice)
    // (Common/tendency.jl:44-45,64-65,329,334,359,371)
    void condensate_partition(FT T, FT rho, FT q_tot, FT& q_liq, FT& q_ice) const {
        FT q_c = kidm::max(FT(0), q_tot - q_vap_saturation(T, rho));
        FT lam = liquid_fraction_T(T);
        q_liq = lam * q_c;
        q_ice = (FT(1) - lam) * q_c;
    }
    // TD.liquid_fraction(thermo_params, T, q_liq, q_ice) (Common/tendency.jl:392)
    FT liquid_fraction(FT T, FT q_liq, FT q_ice) const {
        FT q_c = q_liq + q_ice;
        return (q_c > eps()) ? q_liq / q_c : liquid_fraction_T(T);
    }
    // TD.potential_temperature(thermo_params, T, ρ_dry) (Common/tendency.jl:47)
    FT potential_temperature_dry(FT T, FT rho_dry) const {
        FT p = P.td_R_d * rho_dry * T;
        return T / kidm::pow(p / P.td_MSLP, P.td_R_d / P.td_cp_d);
    }
    // TD.liquid_ice_pottemp(thermo_params, T, ρ, q_tot, q_liq, q_ice) (Common/tendency.jl:48)
    FT liquid_ice_pottemp(FT T, FT rho, FT q_tot, FT q_liq, FT q_ice) const {
        FT R_m = gas_constant_air(q_tot, q_liq, q_ice);
        FT cpm = cp_m(q_tot, q_liq, q_ice);
        FT p = R_m * rho * T;
        FT theta = T / kidm::pow(p / P.td_MSLP, R_m / cpm);
        return theta * (FT(1) - (P.td_LH_v0 * q_liq + P.td_LH_s0 * q_ice) / (cpm * T));
    }
    // CM.ThermodynamicsInterface.supersaturation_over_liquid/ice (K1DModel/tendency.jl:55)
    FT supersaturation_liquid(FT q_tot, FT q_liq, FT q_ice, FT rho, FT T) const {
        FT p_v = (q_tot - q_liq - q_ice) * (rho * P.td_R_v * T);
        return p_v / psat_liquid(T) - FT(1);
    }
    FT supersaturation_ice(FT q_tot, FT q_liq, FT q_ice, FT rho, FT T) const {
        FT p_v = (q_tot - q_liq - q_ice) * (rho * P.td_R_v * T);
        return p_v / psat_ice(T) - FT(1);
    }
    // CM.Common.G_func_liquid / G_func_ice
    FT G_func_liquid(FT T) const {
        FT L = latent_heat_vapor(T);
        FT p_vs = psat_liquid(T);
        return FT(1) / (L / P.air_K_therm / T * (L / P.td_R_v / T - FT(1)) + P.td_R_v * T / P.air_D_vapor / p_vs);
    }
    FT G_func_ice(FT T) const {
        FT L = latent_heat_sublim(T);
        FT p_vs = psat_ice(T);
        return FT(1) / (L / P.air_K_therm / T * (L / P.td_R_v / T - FT(1)) + P.td_R_v * T / P.air_D_vapor / p_vs);
    }
};

// -------------------------------------------- limiter (reference, not upstream)
// src/Common/helper_functions.jl:222-224
template <class FT>
inline FT triangle_inequality_limiter(FT force, FT limit) {
    return force + limit - kidm::sqrt(force * force + limit * limit);
}
// src/Common/tendency.jl:11-14
template <class FT>
inline FT limit(FT q, FT dt, int n = 1) { return kidm::max(FT(0), q) / dt / FT(n); }
// src/Common/tendency.jl:17-19
template <class FT>
inline FT q_(FT rhoq, FT rho) { return kidm::max(FT(0), rhoq / rho); }

// ------------------------------------------------- CloudMicrophysics: common
template <class FT>
inline FT logistic_function(FT x, FT x0, FT k) {
    const FT eps = std::numeric_limits<FT>::epsilon();
    x = kidm::max(FT(0), x);
    if (kidm::fabs(x) < eps) return FT(0);
    return FT(1) / (FT(1) + kidm::exp(-k * (x / x0 - x0 / x)));
}
template <class FT>
inline FT logistic_function_integral(FT x, FT x0, FT k) {
    const FT eps = std::numeric_limits<FT>::epsilon();
    x = kidm::max(FT(0), x);
    if (kidm::fabs(x) < eps) return FT(0);
    FT trnslt = -kidm::log(FT(1) - kidm::exp(-k)) / k;
    FT kt = k * (x / x0 - FT(1) + trnslt);
    return (kt > FT(40)) ? x - x0 : (kidm::log(FT(1) + kidm::exp(kt)) / k - trnslt) * x0;
}

// ------------------------------------------------- CloudMicrophysics: 0M, NonEq
// CM0.remove_precipitation(params, q_liq, q_ice, q_vap_sat) (Common/tendency.jl:396)
template <class FT>
inline FT cm0_remove_precipitation(const Prm<FT>& P, FT q_liq, FT q_ice, FT q_vap_sat) {
    return -kidm::max(FT(0), q_liq + q_ice - P.p0m_S_0 * q_vap_sat) / P.p0m_tau_precip;
}
// CMNe.conv_q_vap_to_q_lcl_icl (Common/tendency.jl:327,332,357,369)
template <class FT>
inline FT cmne_conv_q_vap_to_q_lcl_icl(FT tau_relax, FT q_eq, FT q) { return (q_eq - q) / tau_relax; }

// ------------------------------------------------- CloudMicrophysics: 1-moment
template <class FT>
struct Species1M {  // the (mass, area, vent, pdf, vel) fields the 1M formulas read
    FT r0, m0, me, dm, chim, a0, ae, da, chia, a_vent, b_vent, chiv, ve, dv;
    bool is_snow;
};
template <class FT>
struct CM1 {
    const Prm<FT>& P;
    Thermo<FT> td;
    bool chen2022 = false;  // sedimentation_choice = "Chen2022": CMP.Chen2022VelType (helper_functions.jl:36-37)
    explicit CM1(const Prm<FT>& p, bool chen = false) : P(p), td(p), chen2022(chen) {}
    Species1M<FT> rain() const {
        return {P.rain_r0, P.rain_m0, P.rain_me, P.rain_dm, P.rain_chim, P.rain_a0, P.rain_ae, P.rain_da,
                P.rain_chia, P.rain_a_vent, P.rain_b_vent, P.rain_chiv, P.rain_ve, P.rain_dv, false};
    }
    Species1M<FT> snow() const {
        return {P.snow_r0, P.snow_m0, P.snow_me, P.snow_dm, P.snow_chim, P.snow_a0, P.snow_ae, P.snow_da,
                P.snow_chia, P.snow_a_vent, P.snow_b_vent, P.snow_chiv, P.snow_ve, P.snow_dv, true};
    }
    // Marshall-Palmer intercept: rain/ice constant, snow n0 = μ (ρ q / ρ0)^ν with ρ0 = 1 kg/m3
    FT n0_rain() const { return P.rain_n0; }
    FT n0_ice() const { return P.ice_n0; }
    FT n0_snow(FT q_sno, FT rho) const { return P.snow_mu * kidm::pow(rho * kidm::max(FT(0), q_sno), P.snow_nu); }
    FT n0(const Species1M<FT>& s, FT q, FT rho) const { return s.is_snow ? n0_snow(q, rho) : n0_rain(); }
    // terminal-velocity prefactor: rain from drag balance, snow a parameter
    FT v0(const Species1M<FT>& s, FT rho) const {
        if (s.is_snow) return P.snow_v0;
        return kidm::sqrt(FT(8) / FT(3) / P.rain_C_drag * (P.rain_rho_w / rho - FT(1)) * P.td_grav * P.rain_r0);
    }
    // CM1.lambda_inverse(pdf, mass, q, ρ)
    FT lambda_inverse(FT n0v, FT r0, FT m0, FT me, FT dm, FT chim, FT q, FT rho) const {
        if (q > FT(0) && rho > FT(0)) {
            return kidm::pow(rho * q * kidm::pow(r0, me + dm) / (chim * m0 * n0v * sf_gamma(me + dm + FT(1))),
                            FT(1) / (me + dm + FT(1)));
        }
        return FT(0);
    }
    FT lambda_inverse(const Species1M<FT>& s, FT q, FT rho) const {
        return lambda_inverse(n0(s, q, rho), s.r0, s.m0, s.me, s.dm, s.chim, q, rho);
    }
    FT lambda_inverse_ice(FT q, FT rho) const {
        return lambda_inverse(n0_ice(), P.ice_r0, P.ice_m0, P.ice_me, P.ice_dm, P.ice_chim, q, rho);
    }
    // CM1.terminal_velocity(precip, vel, ρ, q) (Common/tendency.jl:191-192)
    FT terminal_velocity(const Species1M<FT>& s, FT rho, FT q) const {
        if (chen2022 && !s.is_snow) return terminal_velocity_chen2022_rain(rho, q);  // snow: see below
        if (q > FT(0)) {
            FT li = lambda_inverse(s, q, rho);
            return s.chiv * v0(s, rho) * kidm::pow(li / s.r0, s.ve + s.dv) *
                   sf_gamma(s.me + s.ve + s.dm + s.dv + FT(1)) / sf_gamma(s.me + s.dm + FT(1));
        }
        return FT(0);
    }
    // CM1.terminal_velocity(rain, vel::Chen2022VelTypeRain, ρ, q): Chen et al. (2022) eq. 20, mass-weighted (k = 3), over the
    // 1-moment Marshall-Palmer distribution: Σ_i a_i λ^4 Γ(b_i+4) / (λ+c_i)^(b_i+4) / Γ(4) with λ = 1/lambda_inverse(rain) and
    // the Table B1 coefficients as in rain_terminal_velocity_chen2022 below (the fit is applied to the distribution's own
    // size variable, as the upstream call does).
    FT terminal_velocity_chen2022_rain(FT rho, FT q) const {
        if (!(q > FT(0))) return FT(0);
        FT lam = FT(1) / lambda_inverse(rain(), q, rho);
        FT qf = kidm::exp(P.ch_rho0 * rho);
        FT ai[3] = {P.ch_a1 * qf, P.ch_a2 * qf, P.ch_a3 * qf * kidm::pow(rho, P.ch_a3_pow)};
        FT bi[3] = {P.ch_b1 - P.ch_b_rho * rho, P.ch_b2 - P.ch_b_rho * rho, P.ch_b3 - P.ch_b_rho * rho};
        FT ci[3] = {P.ch_c1 * FT(1000), P.ch_c2 * FT(1000), P.ch_c3 * FT(1000)};
        FT s3 = FT(0);
        for (int i = 0; i < 3; ++i) {
            FT aiu = ai[i] * kidm::pow(FT(1000), bi[i]);
            s3 = s3 + aiu * kidm::pow(lam, FT(4)) * FT(std::tgamma(double(bi[i] + FT(4)))) / kidm::pow(lam + ci[i], bi[i] + FT(4)) / FT(6);
        }
        return kidm::max(FT(0), s3);
    }
    // Chen2022VelTypeSnowIce needs the Table B2-B4 constants of Chen et al. (2022), which the parameter table does not carry
    // (they cannot be reproduced here): snow keeps the Blk1M fall-speed law.  Warm columns — every KiD default — carry no
    // snow beyond round-off noise, so only a cold-cloud Chen2022 + 1M run deviates from the reference, in the snow fall
    // speed alone.
    // CM1.conv_q_lcl_to_q_rai(acnv1M, q_liq, smooth_transition = true) (Common/tendency.jl:430)
    FT conv_q_lcl_to_q_rai(FT q_liq) const {
        return logistic_function_integral(q_liq, P.rain_acnv_q_thr, P.rain_acnv_k) / P.rain_acnv_tau;
    }
    // CM1.conv_q_icl_to_q_sno(ice, aps, tps, q_tot, q_liq, q_ice, q_rai, q_sno, ρ, T) (Common/tendency.jl:451)
    FT conv_q_icl_to_q_sno(FT q_tot, FT q_liq, FT q_ice, FT q_rai, FT q_sno, FT rho, FT T) const {
        FT rate = FT(0);
        FT S = td.supersaturation_ice(q_tot, q_liq + q_rai, q_ice + q_sno, rho, T);
        if (q_ice > FT(0) && S > FT(0)) {
            FT G = td.G_func_ice(T);
            FT li = lambda_inverse_ice(q_ice, rho);
            FT r_is = P.ice_r_ice_snow;
            rate = FT(4) * FT(M_PI) * S * G * n0_ice() / rho * kidm::exp(-r_is / li) *
                   (r_is * r_is / (P.ice_me + P.ice_dm) + (r_is * li + li * li));
        }
        return rate;
    }
    // CM1.accretion(cloud, precip, vel, ce, q_clo, q_pre, ρ) (Common/tendency.jl:459,473,481,490)
    FT accretion(FT E, const Species1M<FT>& s, FT q_clo, FT q_pre, FT rho) const {
        FT rate = FT(0);
        if (q_clo > FT(0) && q_pre > FT(0)) {
            FT li = lambda_inverse(s, q_pre, rho);
            FT ex = s.ae + s.ve + s.da + s.dv;
            rate = q_clo * E * n0(s, q_pre, rho) * s.a0 * v0(s, rho) * s.chia * s.chiv * li *
                   sf_gamma(ex + FT(1)) / kidm::pow(s.r0 / li, ex);
        }
        return rate;
    }
    // CM1.accretion_rain_sink(rain, ice, vel, ce, q_ice, q_rai, ρ) (Common/tendency.jl:498)
    FT accretion_rain_sink(FT q_ice, FT q_rai, FT rho) const {
        FT rate = FT(0);
        if (q_ice > FT(0) && q_rai > FT(0)) {
            Species1M<FT> r = rain();
            FT li_ice = lambda_inverse_ice(q_ice, rho);
            FT li = lambda_inverse(r, q_rai, rho);
            FT ex = r.me + r.ae + r.ve + r.dm + r.da + r.dv;
            rate = P.ce_ice_rai / rho * n0_rain() * n0_ice() * r.m0 * r.a0 * v0(r, rho) * r.chim * r.chia * r.chiv *
                   li_ice * li * sf_gamma(ex + FT(1)) / kidm::pow(r.r0 / li, ex);
        }
        return rate;
    }
    // CM1.accretion_snow_rain(type_i, type_j, vel_i, vel_j, ce, q_i, q_j, ρ) (Common/tendency.jl:507,520)
    FT accretion_snow_rain(const Species1M<FT>& si, const Species1M<FT>& sj, FT q_i, FT q_j, FT rho) const {
        FT rate = FT(0);
        if (q_i > FT(0) && q_j > FT(0)) {
            FT n0_i = n0(si, q_i, rho), n0_j = n0(sj, q_j, rho);
            FT li_i = lambda_inverse(si, q_i, rho), li_j = lambda_inverse(sj, q_j, rho);
            FT v_ti = terminal_velocity(si, rho, q_i), v_tj = terminal_velocity(sj, rho, q_j);
            FT mj = sj.me + sj.dm;
            rate = FT(M_PI) / rho * n0_i * n0_j * sj.m0 * sj.chim * P.ce_rai_sno * kidm::fabs(v_ti - v_tj) /
                   kidm::pow(sj.r0, mj) *
                   (FT(2) * sf_gamma(mj + FT(1)) * kidm::pow(li_i, FT(3)) * kidm::pow(li_j, mj + FT(1)) +
                    FT(2) * sf_gamma(mj + FT(2)) * kidm::pow(li_i, FT(2)) * kidm::pow(li_j, mj + FT(2)) +
                    sf_gamma(mj + FT(3)) * li_i * kidm::pow(li_j, mj + FT(3)));
        }
        return rate;
    }
    FT ventilation(const Species1M<FT>& s, FT li, FT rho) const {
        return s.a_vent + s.b_vent * kidm::cbrt(P.air_nu_air / P.air_D_vapor) /
                              kidm::pow(s.r0 / li, (s.ve + s.dv) / FT(2)) *
                              kidm::sqrt(FT(2) * v0(s, rho) * s.chiv / P.air_nu_air * li) *
                              sf_gamma((s.ve + s.dv + FT(5)) / FT(2));
    }
    // CM1.evaporation_sublimation(rain, vel, aps, tps, q_tot, q_liq, q_ice, q_rai, q_sno, ρ, T) (Common/tendency.jl:540)
    FT evaporation_rain(FT q_tot, FT q_liq, FT q_ice, FT q_rai, FT q_sno, FT rho, FT T) const {
        FT rate = FT(0);
        FT S = td.supersaturation_liquid(q_tot, q_liq + q_rai, q_ice + q_sno, rho, T);
        if (q_rai > FT(0) && S < FT(0)) {
            Species1M<FT> r = rain();
            FT G = td.G_func_liquid(T);
            FT li = lambda_inverse(r, q_rai, rho);
            rate = FT(4) * FT(M_PI) * n0_rain() / rho * S * G * li * li * ventilation(r, li, rho);
        }
        return kidm::min(FT(0), rate);
    }
    // ... same method for snow (Common/tendency.jl:563): deposition or sublimation
    FT sublimation_snow(FT q_tot, FT q_liq, FT q_ice, FT q_rai, FT q_sno, FT rho, FT T) const {
        FT rate = FT(0);
        if (q_sno > FT(0)) {
            Species1M<FT> s = snow();
            FT S = td.supersaturation_ice(q_tot, q_liq + q_rai, q_ice + q_sno, rho, T);
            FT G = td.G_func_ice(T);
            FT li = lambda_inverse(s, q_sno, rho);
            rate = FT(4) * FT(M_PI) * n0_snow(q_sno, rho) / rho * S * G * li * li * ventilation(s, li, rho);
        }
        return rate;
    }
    // CM1.snow_melt(snow, vel, aps, tps, q_sno, ρ, T) (Common/tendency.jl:557)
    FT snow_melt(FT q_sno, FT rho, FT T) const {
        FT rate = FT(0);
        if (q_sno > FT(0) && T > P.td_T_freeze) {
            Species1M<FT> s = snow();
            FT L = td.latent_heat_fusion(T);
            FT li = lambda_inverse(s, q_sno, rho);
            rate = FT(4) * FT(M_PI) * n0_snow(q_sno, rho) / rho * P.air_K_therm / L * (T - P.td_T_freeze) * li * li *
                   ventilation(s, li, rho);
        }
        return rate;
    }
};

// ------------------- CloudMicrophysics: 2-moment-style rain formation used by 1M
// CM2.conv_q_lcl_to_q_rai(scheme, q_liq, ρ, N_d) (Common/tendency.jl:434,440)
template <class FT>
inline FT cm2_conv_q_lcl_to_q_rai(const Prm<FT>& P, int rf, FT q_liq, FT rho, FT N_d) {
    const FT eps = std::numeric_limits<FT>::epsilon();
    switch (rf) {
        case KID_RF_KK2000: {
            q_liq = kidm::max(FT(0), q_liq);
            return P.rf_acnv0 * kidm::pow(q_liq, P.rf_acnv1) * kidm::pow(N_d, P.rf_acnv2) * kidm::pow(rho, P.rf_acnv3);
        }
        case KID_RF_B1994: {
            q_liq = kidm::max(FT(0), q_liq);
            FT d = (N_d >= P.rf_acnv4) ? P.rf_acnv5 : P.rf_acnv6;
            return P.rf_acnv0 * kidm::pow(d, P.rf_acnv1) * kidm::pow(q_liq * rho, P.rf_acnv2) *
                   kidm::pow(N_d, P.rf_acnv3) / rho;
        }
        case KID_RF_TC1980: {
            q_liq = kidm::max(FT(0), q_liq);
            FT q_thr = P.rf_acnv5 * N_d / rho * kidm::pow(P.rf_acnv3, P.rf_acnv4);
            FT heav = (q_liq - q_thr > FT(0)) ? FT(1) : FT(0);
            return P.rf_acnv2 * kidm::pow(q_liq, P.rf_acnv0) * kidm::pow(N_d, P.rf_acnv1) * heav;
        }
        case KID_RF_LD2004: {
            if (q_liq <= eps) return FT(0);
            FT L = q_liq * rho;
            FT r_vol = kidm::cbrt(FT(3) * L / FT(4) / FT(M_PI) / P.rf_acnv2 / N_d) * FT(1e6);
            FT beta_6 = kidm::cbrt((r_vol + FT(3)) / r_vol);
            FT E = P.rf_acnv1 * kidm::pow(beta_6, FT(6));
            FT R_6 = beta_6 * r_vol;
            FT R_6C = P.rf_acnv0 / kidm::pow(L, FT(1) / FT(6)) / kidm::sqrt(R_6);
            FT heav = (R_6 - R_6C > FT(0)) ? FT(1) : FT(0);
            return E * L * L * L / N_d / rho * heav;
        }
        case KID_RF_VARTIMESCALE: {
            return kidm::max(FT(0), q_liq) / (P.rf_acnv0 * kidm::pow(N_d / FT(1e8), P.rf_acnv1));
        }
    }
    return FT(0);
}
// CM2.accretion(scheme, q_liq, q_rai[, ρ]) (Common/tendency.jl:463,465)
template <class FT>
inline FT cm2_accretion_1m(const Prm<FT>& P, int rf, FT q_liq, FT q_rai, FT rho) {
    q_liq = kidm::max(FT(0), q_liq);
    q_rai = kidm::max(FT(0), q_rai);
    switch (rf) {
        case KID_RF_KK2000: return P.rf_accr0 * kidm::pow(q_liq * q_rai, P.rf_accr1) * kidm::pow(rho, P.rf_accr2);
        case KID_RF_B1994: return P.rf_accr0 * q_liq * rho * q_rai;
        case KID_RF_TC1980: return P.rf_accr0 * q_liq * q_rai;
    }
    return FT(0);
}

// ------------------------------------------------ CloudMicrophysics: SB2006 (2M)
template <class FT>
struct SB2006 {
    const Prm<FT>& P;
    Thermo<FT> td;
    bool limited;  // CMP.SB2006(toml) vs CMP.SB2006(toml, false) (helper_functions.jl:74-78)
    SB2006(const Prm<FT>& p, bool lim) : P(p), td(p), limited(lim) {}
    FT eps() const { return std::numeric_limits<FT>::epsilon(); }

    struct RainPdf { FT lam_r, x_r; };
    // CM2.pdf_rain_parameters(pdf_r, q_rai, ρ, N_rai): λr [1/m] of n(D)=N0 exp(-λr D), mean mass xr
    RainPdf pdf_rain(FT q_rai, FT rho, FT N_rai) const {
        FT L = rho * q_rai;
        RainPdf o;
        if (limited) {
            FT xr_hat = kidm::max(P.sb_xr_min, kidm::min(P.sb_xr_max, L / N_rai));
            FT N0 = kidm::max(P.sb_N0_min, kidm::min(P.sb_N0_max, N_rai * kidm::cbrt(FT(M_PI) * P.sb_rho_w / xr_hat)));
            o.lam_r = kidm::max(P.sb_lam_min,
                               kidm::min(P.sb_lam_max, kidm::sqrt(kidm::sqrt(FT(M_PI) * P.sb_rho_w * N0 / L))));
            o.x_r = kidm::max(P.sb_xr_min, kidm::min(P.sb_xr_max, L * o.lam_r / N0));
        } else {
            o.x_r = kidm::max(P.sb_xr_min, kidm::min(P.sb_xr_max, L / N_rai));
            o.lam_r = kidm::cbrt(FT(M_PI) * P.sb_rho_w / o.x_r);
        }
        return o;
    }
    // CM2.rain_terminal_velocity(sb2006, vel, q_rai, ρ, N_rai) -> (vt_N, vt_q) (Common/tendency.jl:208-209)
    void rain_terminal_velocity(FT q_rai, FT rho, FT N_rai, FT& vt0, FT& vt1) const {
        FT lam = pdf_rain(q_rai, rho, N_rai).lam_r;
        FT aR = P.sb_aR, bR = P.sb_bR, cR = P.sb_cR;
        FT rc = -FT(1) / (FT(2) * cR) * kidm::log(aR / bR);
        auto G1 = [](FT t) { return kidm::exp(-t); };
        auto G4 = [](FT t) { return (t * t * t + FT(3) * t * t + FT(6) * t + FT(6)) * kidm::exp(-t); };
        FT pa0 = G1(FT(2) * rc * lam), pb0 = G1(FT(2) * rc * (lam + cR));
        FT pa1 = G4(FT(2) * rc * lam) / FT(6), pb1 = G4(FT(2) * rc * (lam + cR)) / FT(6);
        FT sr = kidm::sqrt(P.sb_vel_rho0 / rho);
        FT r1 = FT(1) + cR / lam;
        vt0 = (N_rai < eps()) ? FT(0) : kidm::max(FT(0), sr * (aR * pa0 - bR * pb0 / r1));
        vt1 = (q_rai < eps()) ? FT(0) : kidm::max(FT(0), sr * (aR * pa1 - bR * pb1 / (r1 * r1 * r1 * r1)));
    }
    // CM2.rain_terminal_velocity(sb2006, vel::Chen2022VelTypeRain, q_rai, ρ, N_rai) -> (vt_N, vt_q)
    // (sedimentation_choice = "Chen2022", helper_functions.jl:69-70).  Chen et al. (2022) eq. 20 with the Table B1
    // coefficients: v_k = Σ_i a_i λ^δ Γ(b_i+δ) / (λ+c_i)^(b_i+δ) / Γ(δ), δ = k+1, k = 0 (number) and 3 (mass);
    // a_i = a_i0·exp(ρ0 ρ) (a_3 also ·ρ^a3_pow), b_i = b_i0 - b_ρ ρ, converted from mm to m (a·1000^b, c·1000).
    void rain_terminal_velocity_chen2022(FT q_rai, FT rho, FT N_rai, FT& vt0, FT& vt3) const {
        FT lam = pdf_rain(q_rai, rho, N_rai).lam_r;
        FT q = kidm::exp(P.ch_rho0 * rho);
        FT ai[3] = {P.ch_a1 * q, P.ch_a2 * q, P.ch_a3 * q * kidm::pow(rho, P.ch_a3_pow)};
        FT bi[3] = {P.ch_b1 - P.ch_b_rho * rho, P.ch_b2 - P.ch_b_rho * rho, P.ch_b3 - P.ch_b_rho * rho};
        FT ci[3] = {P.ch_c1 * FT(1000), P.ch_c2 * FT(1000), P.ch_c3 * FT(1000)};
        auto add = [&](FT a, FT b, FT c, int k) {
            FT d = FT(k + 1);
            return a * kidm::pow(lam, d) * FT(std::tgamma(double(b + d))) / kidm::pow(lam + c, b + d) / FT(std::tgamma(double(d)));
        };
        FT s0 = FT(0), s3 = FT(0);
        for (int i = 0; i < 3; ++i) {
            FT aiu = ai[i] * kidm::pow(FT(1000), bi[i]);
            s0 = s0 + add(aiu, bi[i], ci[i], 0);
            s3 = s3 + add(aiu, bi[i], ci[i], 3);
        }
        vt0 = (N_rai < eps()) ? FT(0) : kidm::max(FT(0), s0);
        vt3 = (q_rai < eps()) ? FT(0) : kidm::max(FT(0), s3);
    }
    // CM2.autoconversion(acnv, pdf_c, q_liq, q_rai, ρ, N_liq) (Common/tendency.jl:600,605)
    void autoconversion(FT q_liq, FT q_rai, FT rho, FT N_liq, FT& dq_rai_dt, FT& dN_rai_dt) const {
        dq_rai_dt = dN_rai_dt = FT(0);
        if (q_liq < eps() || N_liq < eps()) return;
        FT L_liq = rho * q_liq;
        FT x_liq = kidm::min(P.sb_x_star, L_liq / N_liq);
        q_rai = kidm::max(FT(0), q_rai);
        FT tau = FT(1) - q_liq / (q_liq + q_rai);
        FT phi_au = P.sb_A * kidm::pow(tau, P.sb_a) * kidm::pow(FT(1) - kidm::pow(tau, P.sb_a), P.sb_b);
        FT nu = P.sb_nu_c;
        FT dL = P.sb_kcc / FT(20) / P.sb_x_star * (nu + FT(2)) * (nu + FT(4)) / ((nu + FT(1)) * (nu + FT(1))) *
                (L_liq * L_liq) * (x_liq * x_liq) * (FT(1) + phi_au / ((FT(1) - tau) * (FT(1) - tau))) * P.sb_rho0 / rho;
        dN_rai_dt = dL / P.sb_x_star;
        dq_rai_dt = dL / rho;
    }
    // CM2.accretion(sb2006, q_liq, q_rai, ρ, N_liq) (Common/tendency.jl:621-622)
    void accretion(FT q_liq, FT q_rai, FT rho, FT N_liq, FT& dq_rai_dt, FT& dN_liq_dt) const {
        dq_rai_dt = dN_liq_dt = FT(0);
        if (q_liq < eps() || q_rai < eps() || N_liq < eps()) return;
        FT L_liq = rho * q_liq, L_rai = rho * q_rai;
        FT x_liq = L_liq / N_liq;
        FT tau = FT(1) - q_liq / (q_liq + q_rai);
        FT phi_ac = kidm::pow(tau / (tau + P.sb_tau0), P.sb_c);
        FT dL = P.sb_kcr * L_liq * L_rai * phi_ac * kidm::sqrt(P.sb_rho0 / rho);
        dN_liq_dt = -dL / x_liq;
        dq_rai_dt = dL / rho;
    }
    // CM2.cloud_liquid_self_collection(acnv, pdf_c, q_liq, ρ, dN_liq_dt_au) (Common/tendency.jl:610)
    FT cloud_liquid_self_collection(FT q_liq, FT rho, FT dN_liq_dt_au) const {
        if (q_liq < eps()) return FT(0);
        FT L = rho * q_liq;
        return -P.sb_kcc * (P.sb_nu_c + FT(2)) / (P.sb_nu_c + FT(1)) * (P.sb_rho0 / rho) * L * L - dN_liq_dt_au;
    }
    // CM2.rain_self_collection(pdf_r, self, q_rai, ρ, N_rai) (Common/tendency.jl:613)
    FT rain_self_collection(FT q_rai, FT rho, FT N_rai) const {
        if (q_rai < eps() || N_rai < eps()) return FT(0);
        FT L = rho * q_rai;
        // λr converted from diameter space [1/m] to mass space [kg^(-1/3)]
        FT lam_x = pdf_rain(q_rai, rho, N_rai).lam_r * kidm::cbrt(FT(6) / FT(M_PI) / P.sb_rho_w);
        return -P.sb_krr * N_rai * L * kidm::sqrt(P.sb_rho0 / rho) * kidm::pow(FT(1) + P.sb_kappa_rr / lam_x, P.sb_d);
    }
    // CM2.rain_breakup(pdf_r, brek, q_rai, ρ, N_rai, dN_rai_dt_sc) (Common/tendency.jl:617)
    FT rain_breakup(FT q_rai, FT rho, FT N_rai, FT dN_rai_dt_sc) const {
        if (q_rai < eps() || N_rai < eps()) return FT(0);
        FT xr = pdf_rain(q_rai, rho, N_rai).x_r;
        FT Dr = kidm::cbrt(xr * FT(6) / FT(M_PI) / P.sb_rho_w);
        FT dD = Dr - P.sb_Deq;
        FT phi_br = (Dr < P.sb_Dr_th) ? FT(-1) : ((dD <= FT(0)) ? P.sb_kbr * dD : FT(2) * (kidm::exp(P.sb_kappa_br * dD) - FT(1)));
        return -(phi_br + FT(1)) * dN_rai_dt_sc;
    }
    // CM2.number_increase/decrease_for_mass_limit(numadj, x, q, ρ, N) (Common/tendency.jl:628-635)
    FT number_increase_for_mass_limit(FT x_max, FT q, FT rho, FT N) const {
        return kidm::max(FT(0), rho * q / x_max - N) / P.sb_numadj_tau;
    }
    FT number_decrease_for_mass_limit(FT x_min, FT q, FT rho, FT N) const {
        return kidm::min(FT(0), rho * q / x_min - N) / P.sb_numadj_tau;
    }
    // CM2.rain_evaporation(sb2006, aps, tps, q_tot, q_liq, q_ice, q_rai, q_sno, ρ, N_rai, T)
    // (Common/tendency.jl:642-644)
    void rain_evaporation(FT q_tot, FT q_liq, FT q_ice, FT q_rai, FT q_sno, FT rho, FT N_rai, FT T, FT& evap0,
                          FT& evap1) const {
        evap0 = evap1 = FT(0);
        FT S = td.supersaturation_liquid(q_tot, q_liq + q_rai, q_ice + q_sno, rho, T);
        if (q_rai > eps() && S < FT(0)) {
            FT x_star = P.sb_xr_min;
            FT G = td.G_func_liquid(T);
            FT xr = pdf_rain(q_rai, rho, N_rai).x_r;
            FT Dr = kidm::cbrt(FT(6) / FT(M_PI) / P.sb_rho_w) * kidm::cbrt(xr);
            FT t_star = kidm::cbrt(FT(6) * x_star / xr);
            FT beta = P.sb_beta;
            FT a_vent_0 = P.sb_av * sf_gamma_inc_upper(FT(-1), t_star) / kidm::pow(FT(6), FT(-2) / FT(3));
            FT b_vent_0 = P.sb_bv * sf_gamma_inc_upper(FT(-0.5) + FT(1.5) * beta, t_star) /
                          kidm::pow(FT(6), beta / FT(2) - FT(0.5));
            FT a_vent_1 = P.sb_av * sf_gamma(FT(2)) / kidm::cbrt(FT(6));
            FT b_vent_1 = P.sb_bv * sf_gamma(FT(2.5) + FT(1.5) * beta) / kidm::pow(FT(6), beta / FT(2) + FT(0.5));
            FT N_Re = P.sb_alpha * kidm::pow(xr, beta) * kidm::sqrt(P.sb_rho0 / rho) * Dr / P.air_nu_air;
            FT sc3 = kidm::cbrt(P.air_nu_air / P.air_D_vapor);
            FT Fv0 = a_vent_0 + b_vent_0 * sc3 * kidm::sqrt(N_Re);
            FT Fv1 = a_vent_1 + b_vent_1 * sc3 * kidm::sqrt(N_Re);
            evap0 = kidm::min(FT(0), FT(2) * FT(M_PI) * G * S * N_rai * Dr * Fv0 / xr);
            evap1 = kidm::min(FT(0), FT(2) * FT(M_PI) * G * S * N_rai * Dr * Fv1 / rho);
            evap0 = (N_rai < eps()) ? FT(0) : evap0;
            evap1 = (q_rai < eps()) ? FT(0) : evap1;
        }
    }
};

// ------------------------------------- CloudMicrophysics: ARG2000 activation
// CMAA.max_supersaturation / total_N_activated for ONE κ-Köhler lognormal mode
// (K1DModel/tendency.jl:69-87).  N_ice = 0 at the only call site.
template <class FT>
struct ARG {
    const Prm<FT>& P;
    Thermo<FT> td;
    explicit ARG(const Prm<FT>& p) : P(p), td(p) {}
    FT coeff_of_curvature(FT T) const { return FT(2) * P.arg_sigma * P.arg_M_w / P.arg_rho_w / P.arg_R / T; }
    FT critical_supersaturation(FT T, FT r_dry, FT kappa) const {
        FT A = coeff_of_curvature(T);
        return FT(2) / kidm::sqrt(kappa) * kidm::pow(A / FT(3) / r_dry, FT(3) / FT(2));
    }
    FT max_supersaturation(FT r_dry, FT stdev, FT N_mode, FT kappa, FT T, FT p, FT w, FT q_tot, FT q_liq, FT q_ice,
                           FT N_liq) const {
        FT eps_m = P.td_R_d / P.td_R_v;  // 1 / molmass_ratio
        FT R_m = td.gas_constant_air(q_tot, q_liq, q_ice);
        FT cpm = td.cp_m(q_tot, q_liq, q_ice);
        FT L = td.latent_heat_vapor(T);
        FT p_vs = td.psat_liquid(T);
        FT G = td.G_func_liquid(T) / P.arg_rho_w;
        FT g = P.arg_g;
        FT alpha = L * g * eps_m / R_m / cpm / (T * T) - g / R_m / T;
        FT gamma = R_m * T / eps_m / p_vs + eps_m * L * L / cpm / T / p;
        FT A = coeff_of_curvature(T);
        FT zeta = FT(2) * A / FT(3) * kidm::sqrt(alpha * w / G);
        FT Sm = critical_supersaturation(T, r_dry, kappa);
        FT lns = kidm::log(stdev);
        FT f = P.arg_f1 * kidm::exp(P.arg_f2 * lns * lns);
        FT gg = P.arg_g1 + P.arg_g2 * lns;
        FT eta = kidm::pow(alpha * w / G, FT(3) / FT(2)) / (FT(2) * FT(M_PI) * P.arg_rho_w * gamma * N_mode);
        FT tmp = FT(1) / (Sm * Sm) *
                 (f * kidm::pow(zeta / eta, P.arg_p1) + gg * kidm::pow(Sm * Sm / (eta + FT(3) * zeta), P.arg_p2));
        FT S_max_ARG = FT(1) / kidm::sqrt(tmp);
        // pre-existing liquid droplets (N_ice = 0 ⇒ the ice terms vanish)
        FT rho_air = p / (R_m * T);
        FT r_liq = (N_liq > std::numeric_limits<FT>::epsilon())
                       ? kidm::cbrt(rho_air * q_liq / N_liq / P.arg_rho_w / (FT(4) / FT(3) * FT(M_PI)))
                       : FT(0);
        FT K_l = FT(4) * FT(M_PI) * P.arg_rho_w * N_liq * r_liq * G * gamma;
        return S_max_ARG * (alpha * w) / (alpha * w + K_l * S_max_ARG);
    }
    FT total_N_activated(FT S_max, FT r_dry, FT stdev, FT N_mode, FT kappa, FT T) const {
        FT Sm = critical_supersaturation(T, r_dry, kappa);
        FT u = FT(2) * kidm::log(Sm / S_max) / FT(3) / kidm::sqrt(FT(2)) / kidm::log(stdev);
        return N_mode * (FT(1) / FT(2)) * (FT(1) - sf_erf(u));
    }
};

}  // namespace kidoracle
