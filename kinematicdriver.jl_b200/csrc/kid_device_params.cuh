// Device-side parameter block of libkid_cuda.so.
//
// DevParams<FT> = the flat table of include/kid_params.h converted to FLOAT_TYPE, plus
// parameter-only derived constants (Γ-function values, reciprocal exponents) that the
// reference recomputes inside every broadcast (CloudMicrophysics evaluates SF.gamma(...) of
// parameter sums per grid point).  When all columns share a parameter set the instance travels
// as a __grid_constant__ kernel argument (operands come straight from the constant bank); for
// per-column sets the tile kernel stages one instance per column in shared memory.
#pragma once
#include <cmath>

#include "../../include/kid_cuda.h"

namespace kid {

constexpr int KID_GSER_TERMS = 26;

#define KID_DERIVED_LIST(X)                                                         \
    X(rain_gam_m)      /* Γ(me+Δm+1)                  CM1.lambda_inverse        */  \
    X(rain_gam_v)      /* Γ(me+ve+Δm+Δv+1)            CM1.terminal_velocity     */  \
    X(rain_gam_accr)   /* Γ(ae+ve+Δa+Δv+1)            CM1.accretion             */  \
    X(rain_gam_vent)   /* Γ((ve+Δv+5)/2)              CM1.evaporation_sublimation */\
    X(rain_gam_sink)   /* Γ(me+ae+ve+Δm+Δa+Δv+1)      CM1.accretion_rain_sink   */  \
    X(rain_gam_m1) X(rain_gam_m2) X(rain_gam_m3) /* Γ(me+Δm+1..3) accretion_snow_rain (j = rain) */ \
    X(snow_gam_m) X(snow_gam_v) X(snow_gam_accr) X(snow_gam_vent)                   \
    X(snow_gam_m1) X(snow_gam_m2) X(snow_gam_m3)                                    \
    X(ice_gam_m)                                                                    \
    X(sb_gam_vent1)    /* Γ(5/2 + 3β/2)               CM2.rain_evaporation      */  \
    X(sb_gam_a1)       /* Γ(a+1), a = -1/2 + 3β/2     (upper incomplete Γ recurrence) */  \
    /* parameter-only sub-expressions hoisted out of the per-point formulas */      \
    X(sb_rc)           /* -1/(2cR) log(aR/bR)         CM2.rain_terminal_velocity */ \
    X(sb_c6pr)         /* (6/π/ρw)^(1/3)                                          */ \
    X(sc3)             /* (ν_air/D_vapor)^(1/3)       ventilation factors         */ \
    X(sb_av0) X(sb_bv0) X(sb_av1) X(sb_bv1) /* evaporation ventilation prefactors */ \
    X(sb_acnv_pref)    /* kcc/20/x*·(ν+2)(ν+4)/(ν+1)²                             */ \
    X(sb_sc_pref)      /* kcc(ν+2)/(ν+1)                                          */ \
    X(arg_lns) X(arg_f) X(arg_gg) X(arg_Sm_pref) /* ARG2000 mode constants        */ \
    X(rain_r0_pow) X(snow_r0_pow) X(ice_r0_pow) /* r0^(me+Δm)                     */ \
    X(rain_li_exp) X(snow_li_exp) X(ice_li_exp) /* 1/(me+Δm+1)                    */ \
    /* reciprocals of parameters that only ever divide (one FMUL instead of rcp + mul per grid point) */ \
    X(inv_tau_relax_liq) X(inv_tau_relax_ice) X(inv_sb_x_star) X(inv_sb_numadj_tau)    \
    X(inv_sb_xc_max) X(inv_sb_xc_min) X(inv_sb_xr_max) X(inv_sb_xr_min)

template <class FT>
struct DevParamsBody {
#define KID_X_FIELD(name) FT name;
    KID_PARAM_LIST(KID_X_FIELD)
    KID_DERIVED_LIST(KID_X_FIELD)
#undef KID_X_FIELD
    // series coefficients 1/(s(s+1)...(s+n)), s = 1/2 + 3β/2, of the lower incomplete γ(s, t*) in
    // CM2.rain_evaporation: t* <= 6^(1/3) < s + 1, so a fixed-length Horner sum replaces the adaptive loop
    FT sb_gser[KID_GSER_TERMS];
};
template <class FT>
struct DevParams : DevParamsBody<FT> {
    // keeps sizeof/sizeof(FT) odd: conflict-free when staged per column in shared memory
    FT pad_to_odd[(sizeof(DevParamsBody<FT>) / sizeof(FT)) % 2 == 0 ? 1 : 2];
};

static_assert((sizeof(DevParams<float>) / sizeof(float)) % 2 == 1, "DevParams word count must be odd");

// host: fill a DevParams<FT> from the double table exactly like the FT-typed constructors would
template <class FT>
inline void fill_dev_params(DevParams<FT>& P, const double* table) {
    int i = 0;
#define KID_X_LOAD(name) P.name = FT(table[i++]);
    KID_PARAM_LIST(KID_X_LOAD)
#undef KID_X_LOAD
    auto G = [](FT x) { return FT(std::tgamma(double(x))); };
    P.rain_gam_m = G(P.rain_me + P.rain_dm + FT(1));
    P.rain_gam_v = G(P.rain_me + P.rain_ve + P.rain_dm + P.rain_dv + FT(1));
    P.rain_gam_accr = G(P.rain_ae + P.rain_ve + P.rain_da + P.rain_dv + FT(1));
    P.rain_gam_vent = G((P.rain_ve + P.rain_dv + FT(5)) / FT(2));
    P.rain_gam_sink = G(P.rain_me + P.rain_ae + P.rain_ve + P.rain_dm + P.rain_da + P.rain_dv + FT(1));
    P.rain_gam_m1 = G(P.rain_me + P.rain_dm + FT(1));
    P.rain_gam_m2 = G(P.rain_me + P.rain_dm + FT(2));
    P.rain_gam_m3 = G(P.rain_me + P.rain_dm + FT(3));
    P.snow_gam_m = G(P.snow_me + P.snow_dm + FT(1));
    P.snow_gam_v = G(P.snow_me + P.snow_ve + P.snow_dm + P.snow_dv + FT(1));
    P.snow_gam_accr = G(P.snow_ae + P.snow_ve + P.snow_da + P.snow_dv + FT(1));
    P.snow_gam_vent = G((P.snow_ve + P.snow_dv + FT(5)) / FT(2));
    P.snow_gam_m1 = G(P.snow_me + P.snow_dm + FT(1));
    P.snow_gam_m2 = G(P.snow_me + P.snow_dm + FT(2));
    P.snow_gam_m3 = G(P.snow_me + P.snow_dm + FT(3));
    P.ice_gam_m = G(P.ice_me + P.ice_dm + FT(1));
    P.sb_gam_vent1 = G(FT(2.5) + FT(1.5) * P.sb_beta);
    P.sb_gam_a1 = G(FT(-0.5) + FT(1.5) * P.sb_beta + FT(1));
    P.sb_rc = -FT(1) / (FT(2) * P.sb_cR) * FT(std::log(double(P.sb_aR / P.sb_bR)));
    P.sb_c6pr = FT(std::cbrt(double(FT(6) / FT(M_PI) / P.sb_rho_w)));
    P.sc3 = FT(std::cbrt(double(P.air_nu_air / P.air_D_vapor)));
    P.sb_av0 = P.sb_av / FT(std::pow(6.0, double(FT(-2) / FT(3))));
    P.sb_bv0 = P.sb_bv / FT(std::pow(6.0, double(P.sb_beta / FT(2) - FT(0.5))));
    P.sb_av1 = P.sb_av * FT(1) / FT(std::cbrt(6.0));
    P.sb_bv1 = P.sb_bv * P.sb_gam_vent1 / FT(std::pow(6.0, double(P.sb_beta / FT(2) + FT(0.5))));
    P.sb_acnv_pref = P.sb_kcc / FT(20) / P.sb_x_star * (P.sb_nu_c + FT(2)) * (P.sb_nu_c + FT(4)) /
                     ((P.sb_nu_c + FT(1)) * (P.sb_nu_c + FT(1)));
    P.sb_sc_pref = P.sb_kcc * (P.sb_nu_c + FT(2)) / (P.sb_nu_c + FT(1));
    P.arg_lns = FT(std::log(double(P.kid_std_dry)));
    P.arg_f = P.arg_f1 * FT(std::exp(double(P.arg_f2 * P.arg_lns * P.arg_lns)));
    P.arg_gg = P.arg_g1 + P.arg_g2 * P.arg_lns;
    P.arg_Sm_pref = FT(2) / FT(std::sqrt(double(P.kid_kappa))) / FT(std::pow(double(FT(3) * P.kid_r_dry), 1.5));
    P.rain_r0_pow = FT(std::pow(double(P.rain_r0), double(P.rain_me + P.rain_dm)));
    P.snow_r0_pow = FT(std::pow(double(P.snow_r0), double(P.snow_me + P.snow_dm)));
    P.ice_r0_pow = FT(std::pow(double(P.ice_r0), double(P.ice_me + P.ice_dm)));
    P.rain_li_exp = FT(1) / (P.rain_me + P.rain_dm + FT(1));
    P.snow_li_exp = FT(1) / (P.snow_me + P.snow_dm + FT(1));
    P.ice_li_exp = FT(1) / (P.ice_me + P.ice_dm + FT(1));
    P.inv_tau_relax_liq = FT(1) / P.ne_tau_relax_liq;
    P.inv_tau_relax_ice = FT(1) / P.ne_tau_relax_ice;
    P.inv_sb_x_star = FT(1) / P.sb_x_star;
    P.inv_sb_numadj_tau = FT(1) / P.sb_numadj_tau;
    P.inv_sb_xc_max = FT(1) / P.sb_xc_max;
    P.inv_sb_xc_min = FT(1) / P.sb_xc_min;
    P.inv_sb_xr_max = FT(1) / P.sb_xr_max;
    P.inv_sb_xr_min = FT(1) / P.sb_xr_min;
    {
        double sgam = double(FT(-0.5) + FT(1.5) * P.sb_beta + FT(1)), d = 1.0;
        for (int n = 0; n < KID_GSER_TERMS; ++n) { d /= (sgam + n); P.sb_gser[n] = FT(d); }
    }
    for (FT& x : P.pad_to_odd) x = FT(0);
}

// run-time switches shared by every kernel (uniform across the grid)
struct DevSwitches {
    int rain_formation;
    int sedimentation;  // enum kid_sedimentation
    int sb_limited;  // CMP.SB2006(toml) vs CMP.SB2006(toml, false)
    int precip_sources, precip_sinks, open_system_activation, local_activation, qtot_flux_correction;
};

}  // namespace kid
