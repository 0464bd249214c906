// Device-side parameter block of libkid_cuda.so.
//
// DevParams<FT> = the flat table of include/kid_params.h converted to FLOAT_TYPE, plus
// parameter-only derived constants (Γ-function values, reciprocal exponents) that the
// reference recomputes inside every broadcast (CloudMicrophysics evaluates SF.gamma(...) of
// parameter sums per grid point).  One instance lives in __constant__ memory when all
// columns share a parameter set (operands come straight from the constant bank); for
// per-column sets the tile kernel stages one instance per column in shared memory.
#pragma once
#include <cmath>

#include "../../include/kid_cuda.h"

namespace kid {

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
    X(sb_gam_a1)       /* Γ(a+1), a = -1/2 + 3β/2     (upper incomplete Γ recurrence) */

template <class FT>
struct DevParams {
#define KID_X_FIELD(name) FT name;
    KID_PARAM_LIST(KID_X_FIELD)
    KID_DERIVED_LIST(KID_X_FIELD)
#undef KID_X_FIELD
    FT pad_to_odd[2];  // keeps sizeof/sizeof(FT) odd: conflict-free when staged per column in shared memory
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
    P.pad_to_odd[0] = P.pad_to_odd[1] = FT(0);
}

// run-time switches shared by every kernel (uniform across the grid)
struct DevSwitches {
    int rain_formation;
    int sb_limited;  // CMP.SB2006(toml) vs CMP.SB2006(toml, false)
    int precip_sources, precip_sinks, open_system_activation, local_activation, qtot_flux_correction;
};

}  // namespace kid
