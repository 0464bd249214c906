// Point physics of the Precipitation2M (SB2006) Float32 path for TWO vertically adjacent grid points at once, in the
// packed Float32 instructions of sm_100a (kid_f2.cuh).  Same formulas, operand order and documented deviations as the
// scalar functions of kid_physics.cuh (which every other style, Float64 and the diagnostics keep using); the
// reference call sites are cited there and repeated per function here.  A branch of the scalar code becomes
// "if any of the two points takes it: evaluate for both, keep the result only where the condition holds" — adjacent
// levels are almost always in the same microphysical regime, so little is evaluated in vain.
#pragma once
#include "kid_f2.cuh"
#include "kid_physics.cuh"

namespace kid {

// aux.thermo_variables + aux.microph_variables of the pair
struct PointX2 {
    F2 rho, inv_rho, q_tot, q_liq, q_ice, q_rai, q_sno, N_liq, N_rai, N_aer;
};
// Time-invariant values per grid point, staged per round (grid_consts_kernel): T never changes
// (src/Common/tendency.jl:34-135 only read it).
//   T, ps_mix = p_sat,mixed(T)   q_vap_saturation(T, ρ) = ps_mix / (ρ R_v T)      (the scalar kernel's values and operation
//   inv_psl = 1 / p_sat,liq(T)   supersaturation = p_v / p_sat,liq - 1             order: q_tot - q_vs and S cancel, so the
//   lam     = liquid_fraction(T) condensate_partition                              two kernels must round alike)
//   G       = G_func_liquid(T)   CM2.rain_evaporation (read from global memory inside that branch)
struct GridConst2 {
    F2 T, ps_mix, inv_psl, lam;  // lam is not staged: liquid_fraction_x2(T) (1 wherever T > T_freeze)
    const float *G0, *G1;  // G(T) of the two points: read from global memory where rain evaporates
};
// TD<float>::liquid_fraction_T for the pair; one warp-uniform branch in a warm run
template <class PRM>
KID_DEV F2 liquid_fraction_x2(const PRM& P, F2 T) {
    F2 lam = f2(1.f);
    if (any(!(T > P.td_T_freeze))) {
        TD<float, PRM> td(P);
        lam = f2(td.liquid_fraction_T(T.x), td.liquid_fraction_T(T.y));
    }
    return lam;
}
// ρ R_v T as TD<float>::condensate_partition / supersaturation form it: (ρ R_v) T
KID_DEV F2 rho_RvT_x2(float R_v, F2 rho, const GridConst2& g) { return (rho * R_v) * g.T; }
// TD.q_vap_saturation: ps_mix / (ρ R_v T)
KID_DEV F2 q_vs_x2(float R_v, F2 rho, const GridConst2& g) { return g.ps_mix / rho_RvT_x2(R_v, rho, g); }
// TD.supersaturation over liquid: (q_tot - q_liq - q_ice) (ρ R_v T) / p_sat,liq - 1, with q_liq, q_ice the condensate totals
KID_DEV F2 supersaturation_x2(float R_v, F2 rho, F2 q_tot, F2 q_l, F2 q_i, const GridConst2& g) {
    return fma2(((q_tot - q_l) - q_i) * rho_RvT_x2(R_v, rho, g), g.inv_psl, -1.f);  // p_v / p_vs - 1 contracts to one FFMA in the scalar build
}

// precompute_aux_thermo! + the q/N part of precompute_aux_precip! (src/Common/tendency.jl:34-73, :104-135, :203-206)
template <int MOIST>
KID_DEV void thermo_point_x2(float R_v, const F2* y, F2 rho_dry, const GridConst2& g, PointX2& a) {
    using L = Lay<MOIST, KID_PRECIP_2M>;
    a.rho = rho_dry + y[L::tot];
    a.inv_rho = vrcp(a.rho);
    a.q_tot = vmax(y[L::tot] * a.inv_rho, 0.f);
    a.q_rai = vmax(y[L::rai] * a.inv_rho, 0.f);
    a.q_sno = vmax(y[L::sno] * a.inv_rho, 0.f);
    a.N_rai = vmax(y[L::Nr], 0.f);
    a.N_liq = vmax(y[L::Nl], 0.f);
    a.N_aer = y[L::Na];
    if constexpr (L::neq) {
        a.q_liq = vmax(y[L::liq] * a.inv_rho, 0.f);
        a.q_ice = vmax(y[L::ice] * a.inv_rho, 0.f);
    } else {
        const F2 q_c = vmax(a.q_tot - q_vs_x2(R_v, a.rho, g), 0.f);
        const F2 ql_eq = g.lam * q_c;
        a.q_liq = ql_eq - a.q_rai;            // src/Common/tendency.jl:64-65: unclamped
        a.q_ice = (q_c - ql_eq) - a.q_sno;    // (1 - λ) q_c
    }
}

// triangle_inequality_limiter(force, limit(x, dt)) with limiter(0, M) = 0 (src/Common/helper_functions.jl:222-224)
KID_DEV F2 limit_x2(F2 q, float inv_dt) { return vmax(q, 0.f) * inv_dt; }
KID_DEV F2 tri_x2(F2 force, F2 lim) {
    const F2 r = (force + lim) - vsqrt_lim(fma2(force, force, lim * lim));
    return sel(veq(force, 0.f), 0.f, r);
}

// CM2 rain size distribution (limited: CMP.SB2006(toml); unlimited: CMP.SB2006(toml, false))
template <class PRM>
KID_DEV void sb_pdf_rain_x2(const PRM& P, bool limited, F2 L, F2 N_rai, F2& lam_r, F2& x_r) {
    const float pirw = float(M_PI) * P.sb_rho_w;
    if (limited) {
        const F2 xr_hat = vmax(vmin(L / N_rai, P.sb_xr_max), P.sb_xr_min);
        const F2 N0 = vmax(vmin(N_rai * vcbrt(pirw / xr_hat), P.sb_N0_max), P.sb_N0_min);
        lam_r = vmax(vmin(vsqrt_fast(vsqrt_fast((pirw * N0) / L)), P.sb_lam_max), P.sb_lam_min);
        x_r = vmax(vmin((L * lam_r) / N0, P.sb_xr_max), P.sb_xr_min);
    } else {
        x_r = vmax(vmin(L / N_rai, P.sb_xr_max), P.sb_xr_min);
        lam_r = vcbrt(pirw / x_r);
    }
}

// CM2.rain_terminal_velocity(SB2006, SB2006VelType) -> (vt_N, vt_q) (src/Common/tendency.jl:208-209)
template <class PRM>
KID_DEV void terminal_velocities_x2(const PRM& P, const DevSwitches& sw, const PointX2& a, F2& vt_q, F2& vt_N) {
    const float eps = 1.1920929e-07f;
    vt_q = vt_N = f2(0.f);
    const B2 no_N = a.N_rai < eps, no_q = a.q_rai < eps;
    if (all(no_N && no_q)) return;
    F2 lam, x_r;
    sb_pdf_rain_x2(P, sw.sb_limited != 0, a.rho * a.q_rai, a.N_rai, lam, x_r);
    const float two_rc = 2.f * P.sb_rc;
    const F2 ta = two_rc * lam, tb = two_rc * (lam + P.sb_cR);
    const F2 ea = vexp(-ta), eb = vexp(-tb);
    const F2 pa1 = fma2(fma2(ta + 3.f, ta, 6.f), ta, 6.f) * ea * (1.f / 6.f);
    const F2 pb1 = fma2(fma2(tb + 3.f, tb, 6.f), tb, 6.f) * eb * (1.f / 6.f);
    const F2 sr = vsqrt_fast(P.sb_vel_rho0 * a.inv_rho);
    const F2 ir1 = vrcp(1.f + P.sb_cR / lam);   // 1 / (1 + cR/λ)
    const F2 ir2 = ir1 * ir1;
    const F2 v0 = vmax(sr * (P.sb_aR * ea - P.sb_bR * eb * ir1), 0.f);
    const F2 v1 = vmax(sr * (P.sb_aR * pa1 - P.sb_bR * pb1 * (ir2 * ir2)), 0.f);
    vt_N = sel(no_N, 0.f, v0);
    vt_q = sel(no_q, 0.f, v1);
}

// out-of-line scalar fallbacks (never taken for CM2.rain_evaporation's argument range; kept out of the hot code)
__device__ __noinline__ float expint_E1_slow(float x) { return expint_E1<float>(x); }
__device__ __noinline__ float gamma_upper_neg_slow(float a, float gam_a1, const float* gser, float x) {
    return gamma_upper_neg<float>(a, gam_a1, gser, x);
}

// E1(x) on x <= 2 (Horner form of kid_physics.cuh::expint_E1); larger arguments fall back to the scalar function
KID_DEV F2 expint_E1_x2(F2 x) {
    constexpr float c[14] = {1.0f, -0.25f, 0.05555555555555555f, -0.010416666666666666f, 0.0016666666666666668f,
        -0.0002314814814814815f, 2.834467120181406e-05f, -3.1001984126984127e-06f, 3.0619243582206544e-07f,
        -2.755731922398589e-08f, 2.2773222086638208e-09f, -1.7397600205070854e-10f, 1.2352142155282914e-11f,
        -8.19283959278968e-13f};
    F2 sum = f2(c[13]);
#pragma unroll
    for (int k = 12; k >= 0; --k) sum = fma2(sum, x, c[k]);
    F2 r = fma2(sum, x, -0.57721566490153286061f) - vlog(x);
    if (any(x > 2.f)) {  // never for CM2.rain_evaporation's t* <= 6^(1/3)
        if (x.x > 2.f) r.x = expint_E1_slow(x.x);
        if (x.y > 2.f) r.y = expint_E1_slow(x.y);
    }
    return r;
}
// Γ(a, x), a in (-1, 0), series branch of kid_physics.cuh::gamma_upper_neg (x < a + 2)
KID_DEV F2 gamma_upper_neg_x2(float a, float gam_a1, const float* gser, F2 x) {
    const float s = a + 1.f;
    const F2 ex = vexp(-x);
    const F2 xa = vpow(x, a);
    F2 sum = f2(gser[14]);
#pragma unroll
    for (int n = 13; n >= 0; --n) sum = fma2(sum, x, gser[n]);
    const F2 upper = gam_a1 - sum * ex * (xa * x);
    F2 r = (upper - xa * ex) * (1.f / a);
    if (any(!(x < s + 1.f))) {
        if (!(x.x < s + 1.f)) r.x = gamma_upper_neg_slow(a, gam_a1, gser, x.x);
        if (!(x.y < s + 1.f)) r.y = gamma_upper_neg_slow(a, gam_a1, gser, x.y);
    }
    return r;
}

// precompute_aux_moisture_sources! (src/Common/tendency.jl:310-376), NonEquilibriumMoisture + 2M
KID_DEV void cloud_sources_x2(float R_v, float inv_tau_liq, float inv_tau_ice, const PointX2& a, const GridConst2& g, F2& cs_liq, F2& cs_ice) {
    const F2 q_c = vmax(a.q_tot - q_vs_x2(R_v, a.rho, g), 0.f);
    const F2 ql_eq = g.lam * q_c;
    const F2 ql = ql_eq - a.q_rai, qi = (q_c - ql_eq) - a.q_sno;
    const F2 S_liq = (ql - a.q_liq) * inv_tau_liq;
    cs_liq = sel(a.N_liq / a.N_aer > 1e-6f, S_liq, 0.f);
    cs_ice = (qi - a.q_ice) * inv_tau_ice;
}

struct Src2M { F2 s_liq, s_rai, s_Na, s_Nl, s_Nr; };

// precompute_aux_precip_sources! for 2M / SB2006 (src/Common/tendency.jl:581-648): kid_physics.cuh::precip_sources_2m
// for two points.  The two mass-limit number adjustments of a species (increase limited by N_aer, decrease limited by
// the species' own number) exclude each other — L/x_max <= L/x_min — so ONE limiter evaluation serves both.
template <class PRM>
KID_DEV void precip_sources_2m_x2(const PRM& P, const DevSwitches& sw, const PointX2& a, const GridConst2& g, float inv_dt,
                                  Src2M& s) {
    const float eps = 1.1920929e-07f;
    const F2 z = f2(0.f);
    const F2 q_liq = a.q_liq, q_rai = a.q_rai, rho = a.rho, N_liq = a.N_liq, N_rai = a.N_rai, N_aer = a.N_aer;
    s.s_liq = s.s_rai = s.s_Na = s.s_Nl = s.s_Nr = z;
    const B2 has_liq = !(q_liq < eps), has_Nl = !(N_liq < eps), has_rai = !(q_rai < eps), has_Nr = !(N_rai < eps);
    // clear air (no cloud water, no rain, no drops): skipped, as in the scalar kernel
    const B2 live = has_liq || has_Nl || has_rai || has_Nr;
    if (!any(live)) return;
    const F2 L_liq = rho * q_liq, L_rai = rho * q_rai;
    // Few, large basic blocks: everything that needs cloud water in one, everything that needs rain in another, so that
    // the independent dependency chains of the processes (pow = lg2 -> mul -> ex2, limiter = rsqrt -> Newton step) overlap;
    // the per-process conditions of the scalar code are applied as selects on the results.
    F2 lam_r = z, x_r = z, sq_rho0 = z;
    const bool do_rai = any(has_rai);
    if (do_rai) {
        sb_pdf_rain_x2(P, sw.sb_limited != 0, L_rai, N_rai, lam_r, x_r);
        sq_rho0 = vsqrt_fast(P.sb_rho0 * a.inv_rho);
    }
    F2 s_liq = z, s_rai = z, s_Na = z, s_Nl = z, s_Nr = z;
    if (sw.precip_sources) {
        const float closed = sw.open_system_activation ? 0.f : -1.f;
        // autoconversion (mass, number): CM2.autoconversion
        F2 S2au = z;
        const B2 m_au = has_liq && has_Nl;
        if (any(has_liq)) {
            const F2 x_liq = vmin(L_liq / N_liq, P.sb_x_star);
            const F2 tau = vmax(1.f - q_liq / (q_liq + vmax(q_rai, 0.f)), 0.f);
            const F2 tau_a = vpow(tau, P.sb_a);
            const F2 phi_au = P.sb_A * tau_a * vpow_small_int(vmax(1.f - tau_a, 0.f), P.sb_b);
            const F2 omt = 1.f - tau;
            const F2 dL = P.sb_acnv_pref * (L_liq * L_liq) * (x_liq * x_liq) * (1.f + phi_au / (omt * omt)) * (P.sb_rho0 * a.inv_rho);
            const F2 S1 = sel(m_au, tri_x2(dL * a.inv_rho, limit_x2(q_liq, inv_dt)), 0.f);
            S2au = sel(m_au, tri_x2(dL * P.inv_sb_x_star, limit_x2(N_liq * 0.5f, inv_dt)), 0.f);
            s_liq -= S1; s_rai += S1;
            s_Nl = fma2(S2au, -2.f, s_Nl); s_Nr += S2au;
            // cloud liquid self-collection
            const F2 sc_l = fma2(S2au, 2.f, (-P.sb_sc_pref * (P.sb_rho0 * a.inv_rho)) * L_liq * L_liq);
            s_Nl -= sel(has_liq, tri_x2(-sc_l, limit_x2(N_liq, inv_dt)), 0.f);
        }
        // rain self-collection and breakup
        const B2 m_r = has_rai && has_Nr;
        if (do_rai) {
            const F2 lam_x = lam_r * P.sb_c6pr;
            const F2 sc_r = (-P.sb_krr * N_rai) * L_rai * sq_rho0 * vpow_small_int(1.f + P.sb_kappa_rr / lam_x, P.sb_d);
            const F2 S1 = -sel(m_r, tri_x2(-sc_r, limit_x2(N_rai, inv_dt)), 0.f);
            const F2 Dr = P.sb_c6pr * vcbrt(x_r);
            const F2 dD = Dr - P.sb_Deq;
            const F2 phi_br = sel(Dr < P.sb_Dr_th, -1.f, sel(dD <= 0.f, P.sb_kbr * dD, 2.f * (vexp(P.sb_kappa_br * dD) - 1.f)));
            const F2 br = -(phi_br + 1.f) * S1;  // takes the LIMITED self-collection value
            s_Nr += S1 + sel(m_r, tri_x2(br, limit_x2(N_rai, inv_dt)), 0.f);
            // accretion
            const B2 m_ac = has_liq && has_rai && has_Nl;
            const F2 tau = vmax(1.f - q_liq / (q_liq + q_rai), 0.f);
            const F2 phi_ac = vpow_small_int(tau / (tau + P.sb_tau0), P.sb_c);
            const F2 dL = P.sb_kcr * L_liq * L_rai * phi_ac * sq_rho0;
            const F2 dN = dL * N_liq / L_liq;   // dL / x_liq, x_liq = L_liq / N_liq  (= -dN_ac)
            const F2 A1 = sel(m_ac, tri_x2(dL * a.inv_rho, limit_x2(q_liq, inv_dt)), 0.f);
            const F2 A2 = sel(m_ac, tri_x2(dN, limit_x2(N_liq, inv_dt)), 0.f);
            s_liq -= A1; s_rai += A1; s_Nl -= A2;
        }
        // number adjustment for mass limits, cloud liquid then rain
        {
            const F2 up = vmax(fma2(L_liq, P.inv_sb_xc_max, -N_liq), 0.f) * P.inv_sb_numadj_tau;
            const F2 dn = vmin(fma2(L_liq, P.inv_sb_xc_min, -N_liq), 0.f) * (-P.inv_sb_numadj_tau);
            const B2 inc = up > 0.f;
            const F2 T = tri_x2(up + dn, limit_x2(sel(inc, N_aer, N_liq), inv_dt));
            const F2 S12 = sel(inc, T, -T);
            s_Na = fma2(S12, closed, s_Na); s_Nl += S12;
        }
        {
            const F2 up = vmax(fma2(L_rai, P.inv_sb_xr_max, -N_rai), 0.f) * P.inv_sb_numadj_tau;
            const F2 dn = vmin(fma2(L_rai, P.inv_sb_xr_min, -N_rai), 0.f) * (-P.inv_sb_numadj_tau);
            const B2 inc = up > 0.f;
            const F2 T = tri_x2(up + dn, limit_x2(sel(inc, N_aer, N_rai), inv_dt));
            const F2 S12 = sel(inc, T, -T);
            s_Na = fma2(S12, closed, s_Na); s_Nr += S12;
        }
    }
    if (sw.precip_sinks) {
        // CM2.rain_evaporation -> (evap_rate_0 -> N_rai, evap_rate_1 -> q_rai)
        const B2 m_q = q_rai > eps;
        if (any(m_q)) {
            const F2 S = supersaturation_x2(P.td_R_v, rho, a.q_tot, q_liq + q_rai, a.q_ice + a.q_sno, g);
            const B2 m_ev = m_q && (S < 0.f);
            if (any(m_ev)) {
                const F2 xr = x_r;
                const F2 Dr = P.sb_c6pr * vcbrt(xr);
                const F2 t_star = vcbrt((6.f * P.sb_xr_min) / xr);
                const float beta = P.sb_beta;
                const F2 ga0 = vexp(-t_star) / t_star - expint_E1_x2(t_star);  // Γ(-1, t*)
                const F2 gb0 = gamma_upper_neg_x2(-0.5f + 1.5f * beta, P.sb_gam_a1, P.sb_gser, t_star);
                const F2 N_Re = P.sb_alpha * vpow(xr, beta) * sq_rho0 * Dr / P.air_nu_air;
                const F2 sre = P.sc3 * vsqrt_fast(N_Re);
                const F2 Fv0 = fma2(P.sb_bv0 * gb0, sre, P.sb_av0 * ga0);
                const F2 Fv1 = fma2(sre, P.sb_bv1, P.sb_av1);
                const F2 common = (2.f * float(M_PI)) * f2(__ldg(g.G0), __ldg(g.G1)) * S * N_rai * Dr;
                F2 e0 = vmin(common * Fv0 / xr, 0.f);
                const F2 e1 = vmin(common * Fv1 * a.inv_rho, 0.f);
                e0 = sel(N_rai < eps, 0.f, e0);
                s_Nr -= sel(m_ev, tri_x2(-e0, limit_x2(N_rai, inv_dt)), 0.f);
                s_rai -= sel(m_ev, tri_x2(-e1, limit_x2(q_rai, inv_dt)), 0.f);
            }
        }
    }
    s.s_liq = sel(live, s_liq, 0.f); s.s_rai = sel(live, s_rai, 0.f); s.s_Na = sel(live, s_Na, 0.f);
    s.s_Nl = sel(live, s_Nl, 0.f); s.s_Nr = sel(live, s_Nr, 0.f);
}

}  // namespace kid
