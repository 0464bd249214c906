// ORACLE — test infrastructure only.  PARITY UNPINNED against real Julia output
// (no Julia toolchain and no vendored CloudMicrophysics/Thermodynamics/ClimaCore in
// this container; see DESIGN.md "Oracle").  What IS pinned: the initial profile
// against the reference's PySDM golden arrays and the reference's unit-test
// invariants (tests/test_oracle_*.py).
//
// kid_oracle.cpp: CPU restatement of the reference's rhs! + SSPRK33 loop, written
// the way the reference is written — one pass over the column per broadcast
// statement, aux fields as arrays — so that it can be read side by side with
//   src/Common/tendency.jl, src/K1DModel/tendency.jl, src/K2DModel/tendency.jl,
//   src/{K1DModel,K2DModel,BoxModel}/ode_utils.jl.
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
// reference) may load the resulting liboracle; libkid_cuda.so never does.
//
// Build: see oracle/Makefile (portable: g++ -O2 -fopenmp -ffp-contract=off; `make native`: -O3 -march=native, built on the
// machine that times it — the CPU arm of bench.py).
#include <cstdint>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <string>
#include <vector>

#ifdef KID_OPCOUNT
#include "opcount.hpp"  // operation-counting scalar: libkid_oracle_opcount.so (oracle/Makefile), measurement only
#endif
#include "kid_physics.hpp"

namespace kidoracle {

// state layout per (moisture, precip): src/Common/ode_utils.jl:23-54
struct Layout {
    int nvars = 0;
    int tot = -1, liq = -1, ice = -1, rai = -1, sno = -1, Nl = -1, Nr = -1, Na = -1;
};
static bool make_layout(int moisture, int precip, Layout& L) {
    L = Layout();
    int n = 0;
    bool neq = moisture == KID_MOIST_NONEQ;
    if (moisture != KID_MOIST_EQ && !neq) return false;
    L.tot = n++;
    if (neq) { L.liq = n++; L.ice = n++; }
    if (precip == KID_PRECIP_1M || precip == KID_PRECIP_2M) { L.rai = n++; L.sno = n++; }
    if (precip == KID_PRECIP_2M) { L.Nl = n++; L.Nr = n++; L.Na = n++; }
    if (precip < KID_PRECIP_NONE || precip > KID_PRECIP_2M) return false;
    L.nvars = n;
    return true;
}

template <class FT>
struct ColAux {  // aux of ONE column: src/Common/ode_utils.jl:110-175
    int nz;
    std::vector<FT> rho, rho_dry, p, T, th_li, th_dry;                         // thermo_variables
    std::vector<FT> q_tot, q_liq, q_ice, q_rai, q_sno, N_liq, N_rai, N_aer;     // microph_variables
    std::vector<FT> vt_rai, vt_sno, vt_N_rai;                                   // velocities
    std::vector<FT> s_q_tot, s_q_liq, s_q_ice, s_q_rai, s_q_sno, s_N_aer, s_N_liq, s_N_rai;  // precip_sources
    std::vector<FT> act_N_aer, act_N_liq;                                       // activation_sources
    std::vector<FT> cs_q_liq, cs_q_ice;                                         // cloud_sources
    std::vector<FT> tmp, tmp2;                                                  // scratch
    std::vector<FT> rho_w;                                                      // prescribed_velocity.ρw (faces, nz+1)
    std::vector<FT> rho_u;                                                      // K2D prescribed_velocity.ρu (centers)
    FT rho_w0 = 0;                                                              // prescribed_velocity.ρw0
    FT q_surf = 0, w1 = 0, prescribed_Nd = 0;
    void resize(int n) {
        nz = n;
        for (auto* v : {&rho, &rho_dry, &p, &T, &th_li, &th_dry, &q_tot, &q_liq, &q_ice, &q_rai, &q_sno, &N_liq,
                        &N_rai, &N_aer, &vt_rai, &vt_sno, &vt_N_rai, &s_q_tot, &s_q_liq, &s_q_ice, &s_q_rai, &s_q_sno,
                        &s_N_aer, &s_N_liq, &s_N_rai, &act_N_aer, &act_N_liq, &cs_q_liq, &cs_q_ice, &tmp, &tmp2, &rho_u})
            v->assign(n, FT(0));
        rho_w.assign(n + 1, FT(0));
    }
};

template <class FT>
struct Model {
    kid_config_t cfg;
    Prm<FT> P;
    Switches sw;
    Layout L;
    int nz, nc;
    FT dz, dt;
    std::vector<ColAux<FT>> aux;      // one per column
    std::vector<FT> Y;                // [column][var][level]
    std::vector<FT> zc;               // center heights
    // K2D horizontal geometry
    int Nq = 0, helem = 0;
    std::vector<double> gll_x, gll_w, gll_D;  // nodes, weights, derivative matrix D[i][j] = l_j'(ξ_i)
    std::vector<FT> xc;                       // node x coordinate per column

    FT* col(std::vector<FT>& a, int c) { return a.data() + size_t(c) * L.nvars * nz; }
    FT* var(FT* base, int v) { return base + size_t(v) * nz; }
    const FT* var(const FT* base, int v) const { return base + size_t(v) * nz; }

    // ------------------------------------------------------------------ rhs
    // src/Common/tendency.jl:24-26
    void zero_tendencies(FT* dY) { std::fill(dY, dY + size_t(L.nvars) * nz, FT(0)); }

    // K1D: src/K1DModel/tendency.jl:208-210, 240-248
    void precompute_aux_prescribed_velocity_k1d(ColAux<FT>& a, FT t) {
        FT rw = FT((t < P.kid_t1) ? double(a.w1) * kidm::sin(M_PI * double(t) / double(P.kid_t1)) : 0.0);
        std::fill(a.rho_w.begin(), a.rho_w.end(), rw);
        a.rho_w0 = rw;
    }
    // K2D: src/K2DModel/tendency.jl:4-46
    void precompute_aux_prescribed_velocity_k2d(ColAux<FT>& a, FT x, FT t) {
        double W = cfg.domain_width, H = cfg.domain_height;
        double ts = (t < P.kid_t1) ? kidm::sin(M_PI * double(t) / double(P.kid_t1)) : 0.0;
        for (int k = 0; k < nz; ++k) {
            double z = double(zc[k]);
            a.rho_u[k] = FT(0.5 * double(a.w1) * W / H * kidm::cos(M_PI / H * z) * kidm::cos(2 * M_PI / W * double(x)) * ts);
        }
        for (int k = 0; k <= nz; ++k) {
            double z = cfg.z_min + double(dz) * k;
            a.rho_w[k] = FT(double(a.w1) * kidm::sin(M_PI / H * z) * kidm::sin(2 * M_PI / W * double(x)) * ts);
        }
        a.rho_w0 = FT(0);
    }

    // src/Common/tendency.jl:34-73 (Eq) and :104-135 (NonEq)
    void precompute_aux_thermo(const FT* Yc, ColAux<FT>& a) {
        Thermo<FT> td(P);
        const FT* r_tot = var(Yc, L.tot);
        bool has_pr = (sw.precip == KID_PRECIP_1M || sw.precip == KID_PRECIP_2M);
        for (int k = 0; k < nz; ++k) a.rho[k] = a.rho_dry[k] + r_tot[k];
        for (int k = 0; k < nz; ++k) a.q_tot[k] = q_(r_tot[k], a.rho[k]);
        if (sw.moisture == KID_MOIST_EQ) {
            for (int k = 0; k < nz; ++k) a.p[k] = td.air_pressure(a.T[k], a.rho[k], a.q_tot[k]);
            for (int k = 0; k < nz; ++k) {
                FT ql, qi;
                td.condensate_partition(a.T[k], a.rho[k], a.q_tot[k], ql, qi);
                if (has_pr) {  // :64-65 — NOT clamped, may go negative
                    a.q_liq[k] = ql - q_(var(Yc, L.rai)[k], a.rho[k]);
                    a.q_ice[k] = qi - q_(var(Yc, L.sno)[k], a.rho[k]);
                } else {
                    a.q_liq[k] = ql;
                    a.q_ice[k] = qi;
                }
                a.th_li[k] = td.liquid_ice_pottemp(a.T[k], a.rho[k], a.q_tot[k], ql, qi);
            }
            for (int k = 0; k < nz; ++k) a.th_dry[k] = td.potential_temperature_dry(a.T[k], a.rho_dry[k]);
        } else {
            for (int k = 0; k < nz; ++k) a.q_liq[k] = q_(var(Yc, L.liq)[k], a.rho[k]);
            for (int k = 0; k < nz; ++k) a.q_ice[k] = q_(var(Yc, L.ice)[k], a.rho[k]);
            for (int k = 0; k < nz; ++k) {
                FT ql = a.q_liq[k], qi = a.q_ice[k];
                if (has_pr) {
                    ql += q_(var(Yc, L.rai)[k], a.rho[k]);
                    qi += q_(var(Yc, L.sno)[k], a.rho[k]);
                }
                a.p[k] = td.air_pressure(a.T[k], a.rho[k], a.q_tot[k], ql, qi);
                a.th_dry[k] = td.potential_temperature_dry(a.T[k], a.rho_dry[k]);
                a.th_li[k] = td.liquid_ice_pottemp(a.T[k], a.rho[k], a.q_tot[k], ql, qi);
            }
        }
    }

    // src/Common/tendency.jl:179-210
    void precompute_aux_precip(const FT* Yc, ColAux<FT>& a) {
        if (sw.precip == KID_PRECIP_1M) {
            CM1<FT> cm1(P, sw.sedimentation == KID_SED_CHEN2022);
            for (int k = 0; k < nz; ++k) a.q_rai[k] = q_(var(Yc, L.rai)[k], a.rho[k]);
            for (int k = 0; k < nz; ++k) a.q_sno[k] = q_(var(Yc, L.sno)[k], a.rho[k]);
            for (int k = 0; k < nz; ++k) a.vt_rai[k] = cm1.terminal_velocity(cm1.rain(), a.rho[k], a.q_rai[k]);
            for (int k = 0; k < nz; ++k) a.vt_sno[k] = cm1.terminal_velocity(cm1.snow(), a.rho[k], a.q_sno[k]);
        } else if (sw.precip == KID_PRECIP_2M) {
            SB2006<FT> sb(P, sw.rain_formation != KID_RF_SB2006NL);
            for (int k = 0; k < nz; ++k) a.q_rai[k] = q_(var(Yc, L.rai)[k], a.rho[k]);
            for (int k = 0; k < nz; ++k) a.q_sno[k] = q_(var(Yc, L.sno)[k], a.rho[k]);
            for (int k = 0; k < nz; ++k) a.N_rai[k] = kidm::max(FT(0), var(Yc, L.Nr)[k]);
            for (int k = 0; k < nz; ++k) a.N_liq[k] = kidm::max(FT(0), var(Yc, L.Nl)[k]);
            for (int k = 0; k < nz; ++k)
                if (sw.sedimentation == KID_SED_CHEN2022)
                    sb.rain_terminal_velocity_chen2022(a.q_rai[k], a.rho[k], a.N_rai[k], a.vt_N_rai[k], a.vt_rai[k]);
                else
                    sb.rain_terminal_velocity(a.q_rai[k], a.rho[k], a.N_rai[k], a.vt_N_rai[k], a.vt_rai[k]);
        }
    }

    // src/K1DModel/tendency.jl:35-96
    FT get_aerosol_activation_rate(FT q_tot, FT q_liq, FT q_ice, FT N_liq, FT N_aer, FT T, FT p, FT rho, FT rho_w) {
        const FT eps = std::numeric_limits<FT>::epsilon();
        Thermo<FT> td(P);
        ARG<FT> arg(P);
        FT S_Nl = FT(0);
        FT S = td.supersaturation_liquid(q_tot, q_liq, q_ice, rho, T);
        if (S < FT(0) || N_aer < eps) return S_Nl;
        FT w = rho_w / rho;
        FT budget = N_aer + (sw.open_system_activation ? FT(0) : N_liq);
        FT preexisting = sw.local_activation ? N_liq : FT(0);
        FT S_max = arg.max_supersaturation(P.kid_r_dry, P.kid_std_dry, budget, P.kid_kappa, T, p, w, q_tot, q_liq,
                                           q_ice, preexisting);
        FT N_act = arg.total_N_activated(S_max, P.kid_r_dry, P.kid_std_dry, budget, P.kid_kappa, T);
        bool zero = kidm::isnan(N_act) || (sw.local_activation && (S_max < S || N_act < N_liq));
        return zero ? FT(0) : (N_act - N_liq) / dt;
    }
    // src/K1DModel/tendency.jl:133-177 (+ find_cloud_base! :21-32)
    void precompute_aux_activation(const FT* Yc, ColAux<FT>& a) {
        if (sw.precip != KID_PRECIP_2M) return;
        std::vector<FT>& S_Nl = a.tmp;
        for (int k = 0; k < nz; ++k) a.N_aer[k] = var(Yc, L.Na)[k];
        for (int k = 0; k < nz; ++k) {
            FT rw_c = (a.rho_w[k + 1] + a.rho_w[k]) / FT(2);  // InterpolateF2C
            S_Nl[k] = get_aerosol_activation_rate(a.q_tot[k], a.q_liq[k] + a.q_rai[k], a.q_ice[k],
                                                  a.N_liq[k] + a.N_rai[k], a.N_aer[k], a.T[k], a.p[k], a.rho[k], rw_c);
        }
        if (!sw.local_activation) {
            // column_reduce!: accumulator starts at level 1, sequential bottom→top
            FT acc_S = S_Nl[0];
            int acc_k = 0;
            for (int k = 1; k < nz; ++k) {
                if (acc_S == FT(0) && S_Nl[k] > FT(0)) { acc_S = S_Nl[k]; acc_k = k; }
            }
            for (int k = 0; k < nz; ++k) S_Nl[k] = (k == acc_k) ? S_Nl[k] : FT(0);
        }
        for (int k = 0; k < nz; ++k) {
            a.act_N_aer[k] = FT(-1) * FT(!sw.open_system_activation) * S_Nl[k];
            a.act_N_liq[k] = S_Nl[k];
        }
    }

    // src/Common/tendency.jl:310-376 + :737-744
    void cloud_sources_tendency(FT* dY, ColAux<FT>& a) {
        if (sw.moisture != KID_MOIST_NONEQ) return;
        Thermo<FT> td(P);
        bool has_pr = (sw.precip == KID_PRECIP_1M || sw.precip == KID_PRECIP_2M);
        for (int k = 0; k < nz; ++k) {
            FT ql, qi;
            td.condensate_partition(a.T[k], a.rho[k], a.q_tot[k], ql, qi);
            FT S_liq = cmne_conv_q_vap_to_q_lcl_icl(P.ne_tau_relax_liq, has_pr ? ql - a.q_rai[k] : ql, a.q_liq[k]);
            if (sw.precip == KID_PRECIP_2M) {
                S_liq = (a.N_liq[k] / a.N_aer[k] > FT(1e-6)) ? S_liq : FT(0);
            }
            FT S_ice = cmne_conv_q_vap_to_q_lcl_icl(P.ne_tau_relax_ice, has_pr ? qi - a.q_sno[k] : qi, a.q_ice[k]);
            a.cs_q_liq[k] = S_liq;
            a.cs_q_ice[k] = S_ice;
        }
        FT* d_liq = var(dY, L.liq);
        FT* d_ice = var(dY, L.ice);
        for (int k = 0; k < nz; ++k) d_liq[k] += a.rho[k] * a.cs_q_liq[k];
        for (int k = 0; k < nz; ++k) d_ice[k] += a.rho[k] * a.cs_q_ice[k];
    }

    // src/Common/tendency.jl:382-404
    void precompute_aux_precip_sources_0m(ColAux<FT>& a) {
        Thermo<FT> td(P);
        for (int k = 0; k < nz; ++k) {
            FT lam = td.liquid_fraction(a.T[k], a.q_liq[k], a.q_ice[k]);
            FT S_qt = -triangle_inequality_limiter(
                -cm0_remove_precipitation(P, a.q_liq[k], a.q_ice[k], td.q_vap_saturation(a.T[k], a.rho[k])),
                limit(a.q_liq[k] + a.q_ice[k], dt));
            a.s_q_tot[k] = S_qt;
            a.s_q_liq[k] = S_qt * lam;
            a.s_q_ice[k] = S_qt * (FT(1) - lam);
        }
    }
    // src/Common/tendency.jl:405-580
    void precompute_aux_precip_sources_1m(ColAux<FT>& a) {
        Thermo<FT> td(P);
        CM1<FT> cm1(P, sw.sedimentation == KID_SED_CHEN2022);
        const FT T_fr = P.td_T_freeze;
        const FT c_vl = P.td_cp_l;  // TD.Parameters.cv_l == cp_l
        const int rf = sw.rain_formation;
        Species1M<FT> rain = cm1.rain(), snow = cm1.snow();
        for (int k = 0; k < nz; ++k) {
            FT q_tot = a.q_tot[k], q_liq = a.q_liq[k], q_ice = a.q_ice[k], q_rai = a.q_rai[k], q_sno = a.q_sno[k];
            FT rho = a.rho[k], T = a.T[k];
            FT s_tot = 0, s_liq = 0, s_ice = 0, s_rai = 0, s_sno = 0, S;
            if (sw.precip_sources) {
                // autoconversion liquid -> rain
                if (rf == KID_RF_CLIMA_1M) {
                    S = -triangle_inequality_limiter(cm1.conv_q_lcl_to_q_rai(q_liq), limit(q_liq, dt));
                } else {
                    S = -triangle_inequality_limiter(cm2_conv_q_lcl_to_q_rai(P, rf, q_liq, rho, a.prescribed_Nd),
                                                     limit(q_liq, dt));
                }
                s_liq += S; s_rai += -S;
                // autoconversion ice -> snow
                S = -triangle_inequality_limiter(cm1.conv_q_icl_to_q_sno(q_tot, q_liq, q_ice, q_rai, q_sno, rho, T),
                                                 limit(q_ice, dt));
                s_ice += S; s_sno += -S;
                // accretion cloud water + rain
                if (rf == KID_RF_CLIMA_1M || rf == KID_RF_LD2004 || rf == KID_RF_VARTIMESCALE) {
                    S = triangle_inequality_limiter(cm1.accretion(P.ce_liq_rai, rain, q_liq, q_rai, rho), limit(q_liq, dt));
                } else {
                    S = triangle_inequality_limiter(cm2_accretion_1m(P, rf, q_liq, q_rai, rho), limit(q_liq, dt));
                }
                s_liq += -S; s_rai += S;
                // accretion cloud ice + snow
                S = triangle_inequality_limiter(cm1.accretion(P.ce_ice_sno, snow, q_ice, q_sno, rho), limit(q_ice, dt));
                s_ice += -S; s_sno += S;
                // sink of cloud water via accretion cloud water + snow
                S = -triangle_inequality_limiter(cm1.accretion(P.ce_liq_sno, snow, q_liq, q_sno, rho), limit(q_liq, dt));
                FT alpha = c_vl / td.latent_heat_fusion(T) * (T - T_fr);
                if (T < T_fr) { s_liq += S; s_sno += -S; }
                else { s_liq += S; s_rai += -S * (FT(1) + alpha); s_sno += S * alpha; }
                // sink of cloud ice via accretion cloud ice - rain
                S = -triangle_inequality_limiter(cm1.accretion(P.ce_ice_rai, rain, q_ice, q_rai, rho), limit(q_ice, dt));
                s_ice += S; s_sno += -S;
                // sink of rain via accretion cloud ice - rain
                S = -triangle_inequality_limiter(cm1.accretion_rain_sink(q_ice, q_rai, rho), limit(q_rai, dt));
                s_rai += S; s_sno += -S;
                // accretion rain - snow
                if (T < T_fr) {
                    S = triangle_inequality_limiter(cm1.accretion_snow_rain(snow, rain, q_sno, q_rai, rho), limit(q_rai, dt));
                } else {
                    S = -triangle_inequality_limiter(cm1.accretion_snow_rain(rain, snow, q_rai, q_sno, rho), limit(q_sno, dt));
                }
                s_rai += -S; s_sno += S;
            }
            if (sw.precip_sinks) {
                // evaporation
                S = -triangle_inequality_limiter(-cm1.evaporation_rain(q_tot, q_liq, q_ice, q_rai, q_sno, rho, T),
                                                 limit(q_rai, dt));
                s_rai += S;
                // melting
                S = -triangle_inequality_limiter(cm1.snow_melt(q_sno, rho, T), limit(q_sno, dt));
                s_rai += -S; s_sno += S;
                // deposition, sublimation
                S = cm1.sublimation_snow(q_tot, q_liq, q_ice, q_rai, q_sno, rho, T);
                S = (S > FT(0)) ? triangle_inequality_limiter(S, limit(q_tot - q_liq - q_ice, dt))
                                : -triangle_inequality_limiter(-S, limit(q_sno, dt));
                s_sno += S;
            }
            a.s_q_tot[k] = s_tot; a.s_q_liq[k] = s_liq; a.s_q_ice[k] = s_ice; a.s_q_rai[k] = s_rai; a.s_q_sno[k] = s_sno;
        }
    }
    // src/Common/tendency.jl:581-648
    void precompute_aux_precip_sources_2m(ColAux<FT>& a) {
        SB2006<FT> sb(P, sw.rain_formation != KID_RF_SB2006NL);
        auto triangle = [&](FT S_x, FT x) { return triangle_inequality_limiter(S_x, limit(x, dt)); };
        const FT closed = FT(-1) * FT(!sw.open_system_activation);
        for (int k = 0; k < nz; ++k) {
            FT q_tot = a.q_tot[k], q_liq = a.q_liq[k], q_ice = a.q_ice[k], q_rai = a.q_rai[k], q_sno = a.q_sno[k];
            FT N_liq = a.N_liq[k], N_rai = a.N_rai[k], N_aer = a.N_aer[k], rho = a.rho[k], T = a.T[k];
            FT s_tot = 0, s_liq = 0, s_rai = 0, s_Na = 0, s_Nl = 0, s_Nr = 0, S1, S2, dq, dN;
            if (sw.precip_sources) {
                sb.autoconversion(q_liq, q_rai, rho, N_liq, dq, dN);
                S1 = triangle(dq, q_liq);
                s_liq += -S1; s_rai += S1;
                S2 = triangle(dN, N_liq / FT(2));
                s_Nl += -FT(2) * S2; s_Nr += S2;
                // liquid self-collection
                S1 = -triangle(-sb.cloud_liquid_self_collection(q_liq, rho, -FT(2) * S2), N_liq);
                s_Nl += S1;
                // rain self-collection
                S1 = -triangle(-sb.rain_self_collection(q_rai, rho, N_rai), N_rai);
                s_Nr += S1;
                // rain breakup (takes the LIMITED self-collection value)
                S1 = triangle(sb.rain_breakup(q_rai, rho, N_rai, S1), N_rai);
                s_Nr += S1;
                // accretion cloud water + rain
                sb.accretion(q_liq, q_rai, rho, N_liq, dq, dN);
                S1 = triangle(dq, q_liq);
                S2 = -triangle(-dN, N_liq);
                s_liq += -S1; s_rai += S1; s_Nl += S2;
                // cloud liquid number adjustment for mass limits
                S1 = triangle(sb.number_increase_for_mass_limit(P.sb_xc_max, q_liq, rho, N_liq), N_aer);
                S2 = -triangle(-sb.number_decrease_for_mass_limit(P.sb_xc_min, q_liq, rho, N_liq), N_liq);
                s_Na += closed * (S1 + S2);
                s_Nl += S1 + S2;
                // rain number adjustment for mass limits
                S1 = triangle(sb.number_increase_for_mass_limit(P.sb_xr_max, q_rai, rho, N_rai), N_aer);
                S2 = -triangle(-sb.number_decrease_for_mass_limit(P.sb_xr_min, q_rai, rho, N_rai), N_rai);
                s_Na += closed * (S1 + S2);
                s_Nr += S1 + S2;
            }
            if (sw.precip_sinks) {
                FT e0, e1;
                sb.rain_evaporation(q_tot, q_liq, q_ice, q_rai, q_sno, rho, N_rai, T, e0, e1);
                S1 = -triangle(-e0, N_rai);
                S2 = -triangle(-e1, q_rai);
                s_rai += S2;
                s_Nr += S1;
            }
            a.s_q_tot[k] = s_tot; a.s_q_liq[k] = s_liq; a.s_q_rai[k] = s_rai;
            a.s_N_aer[k] = s_Na; a.s_N_liq[k] = s_Nl; a.s_N_rai[k] = s_Nr;
        }
    }
    // src/Common/tendency.jl:758-805
    void precip_sources_tendency(FT* dY, ColAux<FT>& a) {
        bool neq = sw.moisture == KID_MOIST_NONEQ;
        if (sw.precip == KID_PRECIP_0M) {
            precompute_aux_precip_sources_0m(a);
            for (int k = 0; k < nz; ++k) var(dY, L.tot)[k] += a.rho[k] * a.s_q_tot[k];
            if (neq) {
                for (int k = 0; k < nz; ++k) var(dY, L.liq)[k] += a.rho[k] * a.s_q_liq[k];
                for (int k = 0; k < nz; ++k) var(dY, L.ice)[k] += a.rho[k] * a.s_q_ice[k];
            }
        } else if (sw.precip == KID_PRECIP_1M) {
            precompute_aux_precip_sources_1m(a);
            for (int k = 0; k < nz; ++k) var(dY, L.tot)[k] += a.rho[k] * a.s_q_tot[k];
            for (int k = 0; k < nz; ++k) var(dY, L.rai)[k] += a.rho[k] * a.s_q_rai[k];
            for (int k = 0; k < nz; ++k) var(dY, L.sno)[k] += a.rho[k] * a.s_q_sno[k];
            if (neq) {
                for (int k = 0; k < nz; ++k) var(dY, L.liq)[k] += a.rho[k] * a.s_q_liq[k];
                for (int k = 0; k < nz; ++k) var(dY, L.ice)[k] += a.rho[k] * a.s_q_ice[k];
            }
        } else if (sw.precip == KID_PRECIP_2M) {
            precompute_aux_precip_sources_2m(a);
            for (int k = 0; k < nz; ++k) var(dY, L.tot)[k] += a.rho[k] * a.s_q_tot[k];
            for (int k = 0; k < nz; ++k) var(dY, L.rai)[k] += a.rho[k] * a.s_q_rai[k];
            for (int k = 0; k < nz; ++k) var(dY, L.Na)[k] += a.s_N_aer[k];
            for (int k = 0; k < nz; ++k) var(dY, L.Nl)[k] += a.s_N_liq[k];
            for (int k = 0; k < nz; ++k) var(dY, L.Nr)[k] += a.s_N_rai[k];
            if (neq) {
                for (int k = 0; k < nz; ++k) var(dY, L.liq)[k] += a.rho[k] * a.s_q_liq[k];
            }
            for (int k = 0; k < nz; ++k) var(dY, L.Nl)[k] += a.act_N_liq[k];
            for (int k = 0; k < nz; ++k) var(dY, L.Na)[k] += a.act_N_aer[k];
        }
    }

    // ------------------------------------------ ClimaCore FD operators (Appendix A)
    enum BC { SET_VALUE, EXTRAPOLATE };
    // face velocity w̃ = ρw / If(ρ) − If(vt) on interior faces 1..nz-1 (vt may be null)
    void face_velocity(const ColAux<FT>& a, const FT* vt, std::vector<FT>& wf) {
        wf.assign(nz + 1, FT(0));
        for (int f = 1; f < nz; ++f) {
            FT If_rho = (a.rho[f] + a.rho[f - 1]) / FT(2);
            FT w = a.rho_w[f] / If_rho;
            if (vt) w = w - (vt[f] + vt[f - 1]) / FT(2);
            wf[f] = w;
        }
    }
    // out[k] += −∂(w̃ · If(x)), ∂ = DivergenceF2C(bottom, top)
    void add_minus_div(const std::vector<FT>& wf, const FT* x, BC bot, FT bot_val, BC top, FT top_val, FT* out) {
        static thread_local std::vector<FT> F, D;  // scratch: no allocation inside the column loop
        F.assign(nz + 1, FT(0)); D.assign(nz, FT(0));
        for (int f = 1; f < nz; ++f) F[f] = wf[f] * ((x[f] + x[f - 1]) / FT(2));
        for (int k = 1; k < nz - 1; ++k) D[k] = (F[k + 1] - F[k]) / dz;
        if (bot == SET_VALUE) D[0] = (F[1] - bot_val) / dz; else D[0] = D[1];
        if (top == SET_VALUE) D[nz - 1] = (top_val - F[nz - 1]) / dz; else D[nz - 1] = D[nz - 2];
        for (int k = 0; k < nz; ++k) out[k] += -D[k];
    }
    // out[k] += fcc(w̃, x), FluxCorrectionC2C(bottom = Extrapolate, top = Extrapolate)
    void add_fcc(const std::vector<FT>& wf, const FT* x, FT* out) {
        static thread_local std::vector<FT> C;
        C.assign(nz, FT(0));
        for (int k = 1; k < nz - 1; ++k)
            C[k] = P.num_fcc_scale * (kidm::fabs(wf[k + 1]) * (x[k + 1] - x[k]) - kidm::fabs(wf[k]) * (x[k] - x[k - 1])) / dz;
        C[0] = C[1];
        C[nz - 1] = C[nz - 2];
        for (int k = 0; k < nz; ++k) out[k] += C[k];
    }

    // K1D src/K1DModel/tendency.jl:274-289 (Eq), :306-332 (NonEq);
    // K2D src/K2DModel/tendency.jl:57-112 (vertical part; fcc on q_tot always)
    void advection_tendency_moisture(FT* dY, const FT* Yc, ColAux<FT>& a, bool k2d) {
        static thread_local std::vector<FT> wf;
        face_velocity(a, nullptr, wf);
        bool fc_tot = k2d ? true : sw.qtot_flux_correction;
        add_minus_div(wf, var(Yc, L.tot), SET_VALUE, a.rho_w0 * a.q_surf, EXTRAPOLATE, 0, var(dY, L.tot));
        if (sw.moisture == KID_MOIST_NONEQ) {
            add_minus_div(wf, var(Yc, L.liq), SET_VALUE, a.rho_w0 * FT(0), EXTRAPOLATE, 0, var(dY, L.liq));
            add_minus_div(wf, var(Yc, L.ice), SET_VALUE, a.rho_w0 * FT(0), EXTRAPOLATE, 0, var(dY, L.ice));
        }
        if (fc_tot) add_fcc(wf, var(Yc, L.tot), var(dY, L.tot));
        if (sw.moisture == KID_MOIST_NONEQ) {
            add_fcc(wf, var(Yc, L.liq), var(dY, L.liq));
            add_fcc(wf, var(Yc, L.ice), var(dY, L.ice));
        }
    }
    // K1D src/K1DModel/tendency.jl:353-435; K2D src/K2DModel/tendency.jl:114-213 (vertical part).
    // Returns the S fields (precip tendencies that are also added to ρq_tot) in S1/S2 so
    // that the K2D caller can add the horizontal part and DSS them before accumulation.
    void advection_tendency_precip(FT* dY, const FT* Yc, ColAux<FT>& a, bool k2d, std::vector<FT>& S1,
                                   std::vector<FT>& S2) {
        S1.assign(nz, FT(0));
        S2.assign(nz, FT(0));
        static thread_local std::vector<FT> wf;
        BC top = k2d ? EXTRAPOLATE : SET_VALUE;
        if (sw.precip == KID_PRECIP_1M) {
            face_velocity(a, a.vt_rai.data(), wf);
            add_minus_div(wf, var(Yc, L.rai), EXTRAPOLATE, 0, top, 0, S1.data());
            add_fcc(wf, var(Yc, L.rai), S1.data());
            face_velocity(a, a.vt_sno.data(), wf);
            add_minus_div(wf, var(Yc, L.sno), EXTRAPOLATE, 0, top, 0, S2.data());
            add_fcc(wf, var(Yc, L.sno), S2.data());
        } else if (sw.precip == KID_PRECIP_2M) {
            if (!k2d) {
                FT surf_aero_flux = a.rho_w0 * a.prescribed_Nd * (FT(1) - a.q_surf) / FT(1.225);
                face_velocity(a, nullptr, wf);
                add_minus_div(wf, var(Yc, L.Na), SET_VALUE, surf_aero_flux, EXTRAPOLATE, 0, var(dY, L.Na));
                add_minus_div(wf, var(Yc, L.Nl), EXTRAPOLATE, 0, SET_VALUE, 0, var(dY, L.Nl));
                add_fcc(wf, var(Yc, L.Na), var(dY, L.Na));
                add_fcc(wf, var(Yc, L.Nl), var(dY, L.Nl));
            }
            face_velocity(a, a.vt_N_rai.data(), wf);
            add_minus_div(wf, var(Yc, L.Nr), EXTRAPOLATE, 0, top, 0, S2.data());
            add_fcc(wf, var(Yc, L.Nr), S2.data());
            face_velocity(a, a.vt_rai.data(), wf);
            add_minus_div(wf, var(Yc, L.rai), EXTRAPOLATE, 0, top, 0, S1.data());
            add_fcc(wf, var(Yc, L.rai), S1.data());
        }
    }
    void accumulate_precip_advection(FT* dY, const std::vector<FT>& S1, const std::vector<FT>& S2) {
        if (sw.precip == KID_PRECIP_1M) {
            for (int k = 0; k < nz; ++k) var(dY, L.tot)[k] += S1[k] + S2[k];
            for (int k = 0; k < nz; ++k) var(dY, L.rai)[k] += S1[k];
            for (int k = 0; k < nz; ++k) var(dY, L.sno)[k] += S2[k];
        } else if (sw.precip == KID_PRECIP_2M) {
            for (int k = 0; k < nz; ++k) var(dY, L.Nr)[k] += S2[k];
            for (int k = 0; k < nz; ++k) var(dY, L.tot)[k] += S1[k];
            for (int k = 0; k < nz; ++k) var(dY, L.rai)[k] += S1[k];
        }
    }

    // rhs! of one column for BOX / K1D / K1D_COL_SED
    // (BoxModel/ode_utils.jl:12-20, K1DModel/ode_utils.jl:30-47, :58-70)
    void rhs_column(FT* dY, const FT* Yc, ColAux<FT>& a, FT t) {
        zero_tendencies(dY);
        if (cfg.model == KID_MODEL_K1D) precompute_aux_prescribed_velocity_k1d(a, t);
        precompute_aux_thermo(Yc, a);
        precompute_aux_precip(Yc, a);
        if (cfg.model == KID_MODEL_K1D) {
            precompute_aux_activation(Yc, a);
            cloud_sources_tendency(dY, a);
        }
        precip_sources_tendency(dY, a);
        if (cfg.model == KID_MODEL_K1D || cfg.model == KID_MODEL_K1D_COL_SED) {
            static thread_local std::vector<FT> S1, S2;
            advection_tendency_moisture(dY, Yc, a, false);
            advection_tendency_precip(dY, Yc, a, false, S1, S2);
            accumulate_precip_advection(dY, S1, S2);
        }
    }

    // ------------------------------------------------------------------ K2D
    // horizontal WeakDivergence of the UVector flux f = ρu/ρ·x on one level, per element,
    // followed by weighted_dss! (src/K2DModel/tendency.jl:68-70,100-108,152-157,202-207).
    // ClimaCore weak divergence (1-D element, uniform J): (wdiv f)_i = −(1/(w_i)) Σ_j D_ji w_j f_j · (2/Δx_e)
    void k2d_hdiv(const std::vector<FT>& flux /*[nc]*/, std::vector<FT>& out /*[nc]*/) {
        FT dxe = FT(cfg.domain_width / double(cfg.helem_global));
        for (int e = 0; e < helem; ++e) {
            for (int i = 0; i < Nq; ++i) {
                FT acc = FT(0);
                for (int j = 0; j < Nq; ++j)
                    acc += FT(gll_D[j * Nq + i]) * (FT(gll_w[j]) * flux[e * Nq + j]);
                out[e * Nq + i] = -acc / FT(gll_w[i]) * (FT(2) / dxe);
            }
        }
    }
    // weighted_dss! on one level: average the duplicated node shared by neighbouring
    // elements (periodic); weights J·w are equal on a uniform mesh.
    void k2d_dss(std::vector<FT>& v /*[nc]*/) {
        for (int e = 0; e < helem; ++e) {
            int en = (e + 1) % helem;
            int r = e * Nq + (Nq - 1), l = en * Nq;
            FT avg = (v[r] + v[l]) / FT(2);
            v[r] = avg;
            v[l] = avg;
        }
    }
    // rhs! of the whole 2-D domain (src/K2DModel/ode_utils.jl:37-52)
    void rhs_k2d(std::vector<FT>& dYall, const std::vector<FT>& Yall, FT t) {
        std::vector<std::vector<FT>> S1(nc), S2(nc);
        for (int c = 0; c < nc; ++c) {
            ColAux<FT>& a = aux[c];
            FT* dY = col(dYall, c);
            const FT* Yc = Yall.data() + size_t(c) * L.nvars * nz;
            zero_tendencies(dY);
            precompute_aux_prescribed_velocity_k2d(a, xc[c], t);
            precompute_aux_thermo(Yc, a);
            precompute_aux_activation(Yc, a);  // BEFORE precip: sees previous call's q_rai, N_liq, N_rai
            precompute_aux_precip(Yc, a);
            cloud_sources_tendency(dY, a);
            precip_sources_tendency(dY, a);
            advection_tendency_moisture(dY, Yc, a, true);
        }
        // horizontal part + DSS of the moisture variables (whole dY field is DSS'd)
        auto hadv_dss = [&](int v, bool into_dY, std::vector<std::vector<FT>>* S) {
            std::vector<FT> flux(nc), hd(nc), fld(nc);
            for (int k = 0; k < nz; ++k) {
                for (int c = 0; c < nc; ++c) {
                    const FT* Yc = Yall.data() + size_t(c) * L.nvars * nz;
                    flux[c] = aux[c].rho_u[k] / aux[c].rho[k] * var(Yc, v)[k];
                }
                k2d_hdiv(flux, hd);
                for (int c = 0; c < nc; ++c) {
                    FT& tgt = into_dY ? var(col(dYall, c), v)[k] : (*S)[c][k];
                    tgt += -hd[c];
                    fld[c] = tgt;
                }
                k2d_dss(fld);
                for (int c = 0; c < nc; ++c) {
                    FT& tgt = into_dY ? var(col(dYall, c), v)[k] : (*S)[c][k];
                    tgt = fld[c];
                }
            }
        };
        hadv_dss(L.tot, true, nullptr);
        if (sw.moisture == KID_MOIST_NONEQ) {
            hadv_dss(L.liq, true, nullptr);
            hadv_dss(L.ice, true, nullptr);
        }
        if (sw.precip == KID_PRECIP_1M || sw.precip == KID_PRECIP_2M) {
            for (int c = 0; c < nc; ++c) {
                const FT* Yc = Yall.data() + size_t(c) * L.nvars * nz;
                advection_tendency_precip(col(dYall, c), Yc, aux[c], true, S1[c], S2[c]);
            }
            if (sw.precip == KID_PRECIP_1M) {
                hadv_dss(L.rai, false, &S1);
                hadv_dss(L.sno, false, &S2);
            } else {
                // 2M: dY.N_rai is DSS'd as a whole field (sources included), S only for ρq_rai
                for (int c = 0; c < nc; ++c) {
                    for (int k = 0; k < nz; ++k) { var(col(dYall, c), L.Nr)[k] += S2[c][k]; S2[c][k] = 0; }
                }
                hadv_dss(L.Nr, true, nullptr);
                hadv_dss(L.rai, false, &S1);
            }
            for (int c = 0; c < nc; ++c) accumulate_precip_advection(col(dYall, c), S1[c], S2[c]);
        }
    }

    // ------------------------------------------------------------ stepping
    void rhs_all(std::vector<FT>& dYall, const std::vector<FT>& Yall, FT t) {
        if (cfg.model == KID_MODEL_K2D) { rhs_k2d(dYall, Yall, t); return; }
#pragma omp parallel for schedule(static)
        for (int c = 0; c < nc; ++c)
            rhs_column(col(dYall, c), Yall.data() + size_t(c) * L.nvars * nz, aux[c], t);
    }
    // OrdinaryDiffEq SSPRK33, fixed dt (call sites: run_KiD_simulation.jl:159-166,
    // run_box_simulation.jl:84-85).  Shu-Osher form exactly as the perform_step! writes it.
    void step(double t0, int nsteps) {
        size_t n = Y.size();
        std::vector<FT> k(n), tmp(n);
        for (int s = 0; s < nsteps; ++s) {
            FT t = FT(t0 + double(s) * double(dt));
            rhs_all(k, Y, t);
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) tmp[i] = Y[i] + dt * k[i];
            rhs_all(k, tmp, t + dt);
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) tmp[i] = (FT(3) * Y[i] + tmp[i] + dt * k[i]) / FT(4);
            rhs_all(k, tmp, t + dt / FT(2));
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) Y[i] = (Y[i] + FT(2) * tmp[i] + FT(2) * dt * k[i]) / FT(3);
        }
    }
};

// ------------------------------------------------------------------ GLL setup
static void gll_setup(int Nq, std::vector<double>& x, std::vector<double>& w, std::vector<double>& D) {
    // Gauss-Lobatto-Legendre nodes/weights by Newton iteration on (1-ξ²)P'_N, N = Nq-1
    int N = Nq - 1;
    x.assign(Nq, 0); w.assign(Nq, 0); D.assign(size_t(Nq) * Nq, 0);
    auto legendre = [&](double xi, double& PN, double& PNm1) {
        double p0 = 1, p1 = xi;
        if (N == 0) { PN = 1; PNm1 = 0; return; }
        for (int n = 2; n <= N; ++n) { double p2 = ((2 * n - 1) * xi * p1 - (n - 1) * p0) / n; p0 = p1; p1 = p2; }
        PN = p1; PNm1 = p0;
    };
    for (int i = 0; i < Nq; ++i) {
        double xi = -kidm::cos(M_PI * i / N);
        for (int it = 0; it < 100; ++it) {
            double PN, PNm1;
            legendre(xi, PN, PNm1);
            // f = (1-ξ²)P'_N = N (P_{N-1} − ξ P_N);  f' = −N(N+1) P_N
            double f = N * (PNm1 - xi * PN), fp = -double(N) * (N + 1) * PN;
            double dx = f / fp;
            if (i == 0 || i == N) break;
            xi -= dx;
            if (kidm::fabs(dx) < 1e-16) break;
        }
        if (i == 0) xi = -1;
        if (i == N) xi = 1;
        double PN, PNm1;
        legendre(xi, PN, PNm1);
        x[i] = xi;
        w[i] = 2.0 / (N * (N + 1) * PN * PN);
    }
    // differentiation matrix D[i][j] = l_j'(ξ_i) via barycentric weights
    std::vector<double> bw(Nq, 1.0);
    for (int j = 0; j < Nq; ++j)
        for (int m = 0; m < Nq; ++m)
            if (m != j) bw[j] /= (x[j] - x[m]);
    for (int i = 0; i < Nq; ++i) {
        double diag = 0;
        for (int j = 0; j < Nq; ++j) {
            if (i == j) continue;
            D[i * Nq + j] = bw[j] / bw[i] / (x[i] - x[j]);
            diag -= D[i * Nq + j];
        }
        D[i * Nq + i] = diag;
    }
}

// --------------------------------------------------------------- type-erased API
struct OracleBase {
    virtual ~OracleBase() {}
    virtual int set_params(const double* table) = 0;
    virtual int set_profiles(const void* rho_dry, const void* T, int per_column) = 0;
    virtual int set_column_scalar(int which, const void* values) = 0;
    virtual int set_state(const void* Y) = 0;
    virtual int get_state(void* Y) = 0;
    virtual int step(double t0, int nsteps) = 0;
    virtual int rhs(double t, void* dY) = 0;
    virtual int get_aux(double t, int field, void* out) = 0;
    virtual int nvars() = 0;
};

template <class FT>
struct Oracle : OracleBase {
    Model<FT> m;
    explicit Oracle(const kid_config_t& cfg) {
        m.cfg = cfg;
        make_layout(cfg.moisture, cfg.precip, m.L);
        m.sw = {cfg.moisture, cfg.precip, cfg.rain_formation, cfg.sedimentation, cfg.precip_sources != 0,
                cfg.precip_sinks != 0, cfg.open_system_activation != 0, cfg.local_activation != 0,
                cfg.qtot_flux_correction != 0};
        m.nz = cfg.nz;
        m.nc = cfg.nc;
        m.dz = FT((cfg.z_max - cfg.z_min) / cfg.nz);
        m.dt = FT(cfg.dt);
        m.aux.resize(m.nc);
        for (auto& a : m.aux) a.resize(m.nz);
        m.Y.assign(size_t(m.nc) * m.L.nvars * m.nz, FT(0));
        m.zc.resize(m.nz);
        for (int k = 0; k < m.nz; ++k) m.zc[k] = FT(cfg.z_min + (k + 0.5) * (cfg.z_max - cfg.z_min) / cfg.nz);
        if (cfg.model == KID_MODEL_K2D) {
            m.Nq = cfg.npoly + 1;
            m.helem = cfg.helem;
            gll_setup(m.Nq, m.gll_x, m.gll_w, m.gll_D);
            m.xc.resize(m.nc);
            double dxe = cfg.domain_width / cfg.helem_global;
            for (int e = 0; e < m.helem; ++e)
                for (int i = 0; i < m.Nq; ++i)
                    m.xc[e * m.Nq + i] = FT((e + cfg.helem_offset) * dxe + (m.gll_x[i] + 1.0) * 0.5 * dxe);
        }
    }
    int nvars() override { return m.L.nvars; }
    int set_params(const double* table) override {
        m.P.load(table);
        for (auto& a : m.aux) { a.w1 = m.P.kid_w1; a.prescribed_Nd = m.P.com_prescribed_Nd; }
        return 0;
    }
    int set_profiles(const void* rho_dry, const void* T, int per_column) override {
        const FT* r = static_cast<const FT*>(rho_dry);
        const FT* t = static_cast<const FT*>(T);
        for (int c = 0; c < m.nc; ++c)
            for (int k = 0; k < m.nz; ++k) {
                size_t i = per_column ? size_t(c) * m.nz + k : size_t(k);
                m.aux[c].rho_dry[k] = r[i];
                m.aux[c].T[k] = t[i];
            }
        return 0;
    }
    int set_column_scalar(int which, const void* values) override {
        const FT* v = static_cast<const FT*>(values);
        for (int c = 0; c < m.nc; ++c) {
            if (which == KID_COL_W1) m.aux[c].w1 = v ? v[c] : m.P.kid_w1;
            else if (which == KID_COL_Q_SURF) m.aux[c].q_surf = v ? v[c] : FT(0);
            else if (which == KID_COL_PRESCRIBED_ND) m.aux[c].prescribed_Nd = v ? v[c] : m.P.com_prescribed_Nd;
            else return 1;
        }
        return 0;
    }
    int set_state(const void* Y) override {
        std::memcpy(m.Y.data(), Y, m.Y.size() * sizeof(FT));
        // aux microphysics fields come from the initial profile `ip` (Common/ode_utils.jl:110-175):
        // the reference never refreshes aux.N_aer in Box / col_sed, and K2D reads the
        // previous call's q_rai, N_liq, N_rai in activation.
        for (int c = 0; c < m.nc; ++c) {
            ColAux<FT>& a = m.aux[c];
            const FT* Yc = m.Y.data() + size_t(c) * m.L.nvars * m.nz;
            for (int k = 0; k < m.nz; ++k) {
                FT rho = a.rho_dry[k] + m.var(Yc, m.L.tot)[k];
                if (m.L.Na >= 0) a.N_aer[k] = m.var(Yc, m.L.Na)[k];
                if (m.L.Nl >= 0) a.N_liq[k] = m.var(Yc, m.L.Nl)[k];
                if (m.L.Nr >= 0) a.N_rai[k] = m.var(Yc, m.L.Nr)[k];
                if (m.L.rai >= 0) a.q_rai[k] = m.var(Yc, m.L.rai)[k] / rho;
                if (m.L.sno >= 0) a.q_sno[k] = m.var(Yc, m.L.sno)[k] / rho;
            }
        }
        return 0;
    }
    int get_state(void* Y) override {
        std::memcpy(Y, m.Y.data(), m.Y.size() * sizeof(FT));
        return 0;
    }
    int step(double t0, int nsteps) override { m.step(t0, nsteps); return 0; }
    int rhs(double t, void* dY) override {
        std::vector<FT> k(m.Y.size());
        m.rhs_all(k, m.Y, FT(t));
        std::memcpy(dY, k.data(), k.size() * sizeof(FT));
        return 0;
    }
    int get_aux(double t, int field, void* out) override {
        // aux as the FSAL rhs!(u, t) call leaves it; evaluated on a copy so that the
        // "previous call" aux of the next real rhs is not disturbed
        std::vector<ColAux<FT>> saved = m.aux;
        std::vector<FT> k(m.Y.size());
        m.rhs_all(k, m.Y, FT(t));
        FT* o = static_cast<FT*>(out);
        for (int c = 0; c < m.nc; ++c) {
            ColAux<FT>& a = m.aux[c];
            const std::vector<FT>* src = nullptr;
            switch (field) {
                case KID_AUX_RHO: src = &a.rho; break;
                case KID_AUX_RHO_DRY: src = &a.rho_dry; break;
                case KID_AUX_P: src = &a.p; break;
                case KID_AUX_T: src = &a.T; break;
                case KID_AUX_THETA_LIQ_ICE: src = &a.th_li; break;
                case KID_AUX_THETA_DRY: src = &a.th_dry; break;
                case KID_AUX_Q_TOT: src = &a.q_tot; break;
                case KID_AUX_Q_LIQ: src = &a.q_liq; break;
                case KID_AUX_Q_ICE: src = &a.q_ice; break;
                case KID_AUX_Q_RAI: src = &a.q_rai; break;
                case KID_AUX_Q_SNO: src = &a.q_sno; break;
                case KID_AUX_N_LIQ: src = &a.N_liq; break;
                case KID_AUX_N_RAI: src = &a.N_rai; break;
                case KID_AUX_N_AER: src = &a.N_aer; break;
                case KID_AUX_TERM_VEL_RAI: src = &a.vt_rai; break;
                case KID_AUX_TERM_VEL_SNO: src = &a.vt_sno; break;
                case KID_AUX_TERM_VEL_N_RAI: src = &a.vt_N_rai; break;
                case KID_AUX_SRC_Q_TOT: src = &a.s_q_tot; break;
                case KID_AUX_SRC_Q_LIQ: src = &a.s_q_liq; break;
                case KID_AUX_SRC_Q_ICE: src = &a.s_q_ice; break;
                case KID_AUX_SRC_Q_RAI: src = &a.s_q_rai; break;
                case KID_AUX_SRC_Q_SNO: src = &a.s_q_sno; break;
                case KID_AUX_SRC_N_AER: src = &a.s_N_aer; break;
                case KID_AUX_SRC_N_LIQ: src = &a.s_N_liq; break;
                case KID_AUX_SRC_N_RAI: src = &a.s_N_rai; break;
                case KID_AUX_ACT_N_AER: src = &a.act_N_aer; break;
                case KID_AUX_ACT_N_LIQ: src = &a.act_N_liq; break;
                case KID_AUX_CLOUD_SRC_Q_LIQ: src = &a.cs_q_liq; break;
                case KID_AUX_CLOUD_SRC_Q_ICE: src = &a.cs_q_ice; break;
                default: m.aux = saved; return 1;
            }
            std::memcpy(o + size_t(c) * m.nz, src->data(), m.nz * sizeof(FT));
        }
        m.aux = saved;
        return 0;
    }
};

}  // namespace kidoracle

using namespace kidoracle;

extern "C" {
// OpenMP threads of the column loops.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: bench.py sets the
// count explicitly so that the CPU arm uses all host cores under torchrun too.  Returns the count in effect.
int kidoracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
void* kidoracle_create(const kid_config_t* cfg) {
    Layout L;
    if (!make_layout(cfg->moisture, cfg->precip, L)) return nullptr;
#ifndef KID_OPCOUNT
    if (cfg->float_type == KID_F32) return new Oracle<float>(*cfg);
#endif
#ifdef KID_OPCOUNT
    // the counting build runs everything in its counting scalar (a wrapped double): host arrays are Float64
    if (cfg->float_type == KID_F64) return new Oracle<kidm::Cnt>(*cfg);
    return nullptr;
#else
    if (cfg->float_type == KID_F64) return new Oracle<double>(*cfg);
    return nullptr;
#endif
}
#ifdef KID_OPCOUNT
// counters since the last reset: {add, mul, div, sqrt, exp/log/pow/cbrt/sin/cos, erf/Γ, comparisons+min/max/abs}
void kidoracle_opcount_reset(void) { kidm::counters() = kidm::OpCounters(); }
void kidoracle_opcount_get(double* out) {
    const kidm::OpCounters& c = kidm::counters();
    out[0] = c.add; out[1] = c.mul; out[2] = c.div; out[3] = c.sqrt; out[4] = c.transc; out[5] = c.special; out[6] = c.cmp;
}
#endif
void kidoracle_destroy(void* h) { delete static_cast<OracleBase*>(h); }
int kidoracle_nvars(void* h) { return static_cast<OracleBase*>(h)->nvars(); }
int kidoracle_set_params(void* h, const double* table, int nparams) {
    if (nparams != KID_NPARAMS) return 1;
    return static_cast<OracleBase*>(h)->set_params(table);
}
int kidoracle_set_profiles(void* h, const void* rho_dry, const void* T, int per_column) {
    return static_cast<OracleBase*>(h)->set_profiles(rho_dry, T, per_column);
}
int kidoracle_set_column_scalar(void* h, int which, const void* v) {
    return static_cast<OracleBase*>(h)->set_column_scalar(which, v);
}
int kidoracle_set_state(void* h, const void* Y) { return static_cast<OracleBase*>(h)->set_state(Y); }
int kidoracle_get_state(void* h, void* Y) { return static_cast<OracleBase*>(h)->get_state(Y); }
int kidoracle_step(void* h, double t0, int nsteps) { return static_cast<OracleBase*>(h)->step(t0, nsteps); }
int kidoracle_rhs(void* h, double t, void* dY) { return static_cast<OracleBase*>(h)->rhs(t, dY); }
int kidoracle_get_aux(void* h, double t, int field, void* out) {
    return static_cast<OracleBase*>(h)->get_aux(t, field, out);
}
int kidoracle_nparams(void) { return KID_NPARAMS; }
// special functions exposed for pinning against scipy in tests/test_oracle_special.py
double kidoracle_gamma_upper(double a, double x) { return gamma_upper(a, x); }
double kidoracle_expint_E1(double x) { return expint_E1(x); }
// GLL nodes/weights/derivative matrix for pinning against numpy's Legendre module
void kidoracle_gll(int Nq, double* x, double* w, double* D) {
    std::vector<double> vx, vw, vD;
    gll_setup(Nq, vx, vw, vD);
    std::memcpy(x, vx.data(), Nq * sizeof(double));
    std::memcpy(w, vw.data(), Nq * sizeof(double));
    std::memcpy(D, vD.data(), size_t(Nq) * Nq * sizeof(double));
}
}
