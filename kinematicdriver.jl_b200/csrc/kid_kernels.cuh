// Kernels of the hot path (sm_100a): BoxModel and K1DModel rhs + SSPRK33.
//
// Device state layout (structure of arrays over the batch of columns):
//     Y[var][level][column], column fastest, row pitch ncp (columns padded to 32)
// so that the 32 lanes of a warp always touch 32 neighbouring columns of one (var, level)
// row: every global access is one fully used 128-byte line (f32) / two lines (f64).
//
// k1d_tile_kernel: one CTA owns a TILE of C columns x all nz levels and keeps the tile's
// state in shared memory for the whole launch (all SSPRK33 stages of `nsteps` steps), so
// HBM sees one read and one write of the state per LAUNCH while the rhs — thermodynamics,
// terminal velocities, ARG activation + cloud-base search, condensation, 0M/1M/2M process
// rates, the flux-corrected advection/sedimentation stencil and the SSPRK33 stage combine
// — is evaluated from shared memory/registers.  Thread (c, j) serves column c; in round r it
// works on level r·NW + j, so a warp is 32 columns at the SAME level (the microphysics
// branches are warp-coherent across ensemble members) and every warp gets levels from all
// parts of the column (balanced load, no serial sweep).  Per stage: pass 1 (terminal
// velocities, activation, cloud base) -> barrier -> per round: READ phase (old neighbours ->
// advective tendencies) -> barrier -> WRITE phase (sources + stage combine, in place).
// DESIGN.md §3.1 has the full description.
//
// Replaces (reference): rhs! closures src/K1DModel/ode_utils.jl:29-73,
// src/BoxModel/ode_utils.jl:10-22, the SSPRK33 loop of ODE.solve
// (test/experiments/KiD_driver/run_KiD_simulation.jl:159-166).
#pragma once
#include <climits>
// threads per tile CTA: the Float32 kernel fits in 64 registers (1024 threads = 32 warps per SM);
// Float64 needs the full 128 (512 threads)
#ifndef KID_TILE_MAXT_F32
#define KID_TILE_MAXT_F32 1024
#endif
#ifndef KID_TILE_MAXT_F64
#define KID_TILE_MAXT_F64 512
#endif
namespace kid {
template <class FT> constexpr int tile_max_threads() { return sizeof(FT) == 4 ? KID_TILE_MAXT_F32 : KID_TILE_MAXT_F64; }
}
#include <type_traits>

#include "kid_f2.cuh"
#include "kid_physics.cuh"

// -DKID_DEBUG_BOUNDS (make debug): device-side asserts on the shared-memory rows the tile kernel addresses — the
// stand-in for compute-sanitizer's memcheck on the hand-computed tile/stash indices (scripts/all_kernels_smoke.py runs
// every kernel family against such a build).
#ifdef KID_DEBUG_BOUNDS
#include <cassert>
#define KID_ASSERT_ROW(row, nrows) assert((row) >= 0 && (row) < (nrows))
#else
#define KID_ASSERT_ROW(row, nrows) ((void)0)
#endif

namespace kid {

// compile-time loop: f(std::integral_constant<int, I>) for I in [0, N)
template <int I, int N, class F>
KID_DEV void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <class FT>
struct KArgs {
    FT* Y;                  // [NV][nz][ncp] device state (in/out)
    const FT* rho_dry;      // [nz][ncp] or [nz]
    const FT* T;            // same shape
    const FT* w1;           // [ncp] kid_params.w1 per column
    const FT* q_surf;       // [ncp] aux.q_surf per column
    const FT* Nd;           // [ncp] common_params.prescribed_Nd per column
    const FT* aux_Naer;     // [nz][ncp] aux.microph_variables.N_aer snapshot (Box / col_sed), may be null
    const FT* ps_liq;       // saturation vapour pressure over liquid / of the mixed phase at the (time-invariant) T,
    const FT* ps_mix;       // same shape as T; filled by profile_consts_kernel when params or profiles change
    const DevParams<FT>* prm_sets;  // per-column parameter sets (calibration ensembles), null if one set is shared
    const int* col_to_set;          // [ncp] column -> parameter set
    FT* out;                // DIAG: dY [NV][nz][ncp] (mode 1) or the requested aux fields [nfields][nz][ncp] (mode 2)
    int nz, nc, ncp;
    int prof_per_column;    // 0: rho_dry/T are [nz]; 1: [nz][ncp]
    int model;              // enum kid_model
    int nsteps;
    int C, NW;              // tile: columns per CTA, level chunks per column
    int no_fixed_shape;     // tuning: never pick the compile-time-shape instantiation
    int diag_mode, field;   // 0 = step, 1 = rhs only, 2 = aux fields (fields[0..nfields) -> out[f][nz][ncp])
    int nfields;
    int fields[KID_NAUX];
    FT dt, dz;
    double t0;
    DevSwitches sw;
    // K2D only (src/K2DModel): node x coordinate per column, previous-call aux (q_rai, N_liq, N_rai) that
    // activation reads before precompute_aux_precip! refreshes them, separate output of the precipitation
    // advection fields S1/S2 (they are DSS'd on their own before being added to ρq_tot)
    const FT* xc;           // [ncp]
    const double* zsin;     // [nz + 1] sin(π z_face / H): the vertical factor of the stream function per face
    FT* prev_aux;           // [3][nz][ncp]
    FT* outS;               // [2][nz][ncp]
    int update_prev;
    double width, height, z_min;
};

// ------------------------------------------------------------ advected variables
enum { VEL_AIR = 0, VEL_VT1 = 1, VEL_VT2 = 2 };
enum { BC_SET = 0, BC_EXTRAP = 1 };
enum { BV_ZERO = 0, BV_QSURF = 1, BV_AERO = 2 };
enum { FCC_ALWAYS = 0, FCC_QTOT_FLAG = 1 };
struct AdvVar {
    int idx, vel, bot, botval, top, fcc, to_tot;
};
// K1D advection tables: src/K1DModel/tendency.jl:274-289 (Eq), :306-332 (NonEq),
// :353-400 (1M), :401-435 (2M)
template <int MOIST, int PRECIP>
__host__ __device__ constexpr int adv_count() {
    using L = Lay<MOIST, PRECIP>;
    return 1 + (L::neq ? 2 : 0) + (PRECIP == KID_PRECIP_1M ? 2 : 0) + (PRECIP == KID_PRECIP_2M ? 4 : 0);
}
template <int MOIST, int PRECIP>
__host__ __device__ constexpr AdvVar adv_entry(int i) {
    using L = Lay<MOIST, PRECIP>;
    if (i == 0) return {L::tot, VEL_AIR, BC_SET, BV_QSURF, BC_EXTRAP, FCC_QTOT_FLAG, 0};
    int b = 1;
    if (L::neq) {
        if (i == 1) return {L::liq, VEL_AIR, BC_SET, BV_ZERO, BC_EXTRAP, FCC_ALWAYS, 0};
        if (i == 2) return {L::ice, VEL_AIR, BC_SET, BV_ZERO, BC_EXTRAP, FCC_ALWAYS, 0};
        b = 3;
    }
    if (PRECIP == KID_PRECIP_1M) {
        if (i == b) return {L::rai, VEL_VT1, BC_EXTRAP, BV_ZERO, BC_SET, FCC_ALWAYS, 1};
        return {L::sno, VEL_VT2, BC_EXTRAP, BV_ZERO, BC_SET, FCC_ALWAYS, 1};
    }
    // 2M
    if (i == b) return {L::Na, VEL_AIR, BC_SET, BV_AERO, BC_EXTRAP, FCC_ALWAYS, 0};
    if (i == b + 1) return {L::Nl, VEL_AIR, BC_EXTRAP, BV_ZERO, BC_SET, FCC_ALWAYS, 0};
    if (i == b + 2) return {L::Nr, VEL_VT2, BC_EXTRAP, BV_ZERO, BC_SET, FCC_ALWAYS, 0};
    return {L::rai, VEL_VT1, BC_EXTRAP, BV_ZERO, BC_SET, FCC_ALWAYS, 1};
}

// prescribed momentum flux ρw(t), src/K1DModel/tendency.jl:208-210 (evaluated in double, rounded to FT)
KID_DEV double rho_w_of_t(double t, double w1, double t1) { return (t < t1) ? w1 * sin(M_PI * t / t1) : 0.0; }
// Float32: w1·sin(π t/t1) with sinpif (< 2 ulp) instead of a double-precision sine per thread and stage
KID_DEV float rho_w_of_t(float t, float w1, float t1) { return (t < t1) ? w1 * sinpif(t / t1) : 0.0f; }

// SSPRK33 stage combine exactly as OrdinaryDiffEq's perform_step! writes it
template <class FT>
KID_DEV FT ssprk33_combine(int stage, FT u, FT y, FT dt, FT k) {
    if (stage == 0) return y + dt * k;
    if (stage == 1) return (FT(3) * u + y + dt * k) / FT(4);
    return (u + FT(2) * y + FT(2) * dt * k) / FT(3);
}
template <class FT>
KID_DEV FT stage_time(int stage, FT t, FT dt) {
    return stage == 0 ? t : (stage == 1 ? t + dt : t + dt / FT(2));
}

// requested aux field of one point (kid_get_aux; schema src/Common/ode_utils.jl:110-175)
template <class FT, int MOIST, int PRECIP, class PRM>
KID_DEV FT aux_field_value(const PRM& P, int field, const Point<FT>& a, const Sources<FT>& s, FT v1, FT v2, FT act_Nl,
                           bool open_system) {
    using L = Lay<MOIST, PRECIP>;
    TD<FT, PRM> td(P);
    switch (field) {
        case KID_AUX_RHO: return a.rho;
        case KID_AUX_RHO_DRY: return a.rho_dry;
        case KID_AUX_P: return pressure_point<FT, MOIST, PRECIP>(P, a);
        case KID_AUX_T: return a.T;
        case KID_AUX_THETA_DRY: return td.potential_temperature_dry(a.T, a.rho_dry);
        case KID_AUX_THETA_LIQ_ICE:
            if constexpr (!L::neq) return td.liquid_ice_pottemp(a.T, a.rho, a.q_tot, a.ql_eq, a.qi_eq);
            else if constexpr (L::has_pr)
                return td.liquid_ice_pottemp(a.T, a.rho, a.q_tot, a.q_liq + a.q_rai, a.q_ice + a.q_sno);
            else return td.liquid_ice_pottemp(a.T, a.rho, a.q_tot, a.q_liq, a.q_ice);
        case KID_AUX_Q_TOT: return a.q_tot;
        case KID_AUX_Q_LIQ: return a.q_liq;
        case KID_AUX_Q_ICE: return a.q_ice;
        case KID_AUX_Q_RAI: return a.q_rai;
        case KID_AUX_Q_SNO: return a.q_sno;
        case KID_AUX_N_LIQ: return a.N_liq;
        case KID_AUX_N_RAI: return a.N_rai;
        case KID_AUX_N_AER: return a.N_aer;
        case KID_AUX_TERM_VEL_RAI: return v1;
        case KID_AUX_TERM_VEL_SNO: return PRECIP == KID_PRECIP_1M ? v2 : FT(0);
        case KID_AUX_TERM_VEL_N_RAI: return PRECIP == KID_PRECIP_2M ? v2 : FT(0);
        case KID_AUX_SRC_Q_TOT: return s.s_tot;
        case KID_AUX_SRC_Q_LIQ: return s.s_liq;
        case KID_AUX_SRC_Q_ICE: return s.s_ice;
        case KID_AUX_SRC_Q_RAI: return s.s_rai;
        case KID_AUX_SRC_Q_SNO: return s.s_sno;
        case KID_AUX_SRC_N_AER: return s.s_Na;
        case KID_AUX_SRC_N_LIQ: return s.s_Nl;
        case KID_AUX_SRC_N_RAI: return s.s_Nr;
        case KID_AUX_ACT_N_AER: return (open_system ? FT(0) : FT(-1)) * act_Nl;
        case KID_AUX_ACT_N_LIQ: return act_Nl;
        case KID_AUX_CLOUD_SRC_Q_LIQ: return s.cs_liq;
        case KID_AUX_CLOUD_SRC_Q_ICE: return s.cs_ice;
    }
    return FT(0);
}

// shared-memory footprint of one tile (bytes); must match the carve-up in k1d_tile_kernel
template <class FT, int MOIST, int PRECIP>
inline size_t k1d_tile_smem_bytes(int nz, int C, bool per_column_params = false) {
    using L = Lay<MOIST, PRECIP>;
    constexpr int NADV = adv_count<MOIST, PRECIP>();
    if (per_column_params) return k1d_tile_smem_bytes<FT, MOIST, PRECIP>(nz, C, false) + size_t(C) * sizeof(DevParams<FT>) + 16;
    size_t nfl = size_t(L::NV) * (nz + 4) * C;  // Ys: nz levels + 4 stash rows [2 parities][2 levels] per variable
    nfl += size_t(2) * nz * C;                 // rho_dry, T
    if (L::has_pr) nfl += size_t(2) * nz * C;  // vt1, vt2
    if (L::is2m) nfl += size_t(nz) * C;        // activation rate
    return nfl * sizeof(FT) + size_t(2) * C * sizeof(int);
}

// FIXNZ > 0: the tile shape is a compile-time constant (nz = FIXNZ levels, C = tile_fixed_columns<FT>() columns,
// NW = tile_max_threads / C level-warps), so every shared-memory offset folds into an immediate; FIXNZ = 0: any shape.
template <class FT> constexpr int tile_fixed_columns() { return sizeof(FT) == 4 ? 32 : 16; }
template <class FT, int MOIST, int PRECIP, bool DIAG, class PRMSRC, int FIXNZ = 0>
__global__ void __launch_bounds__(tile_max_threads<FT>(), 1)
k1d_tile_kernel(const KArgs<FT> A, const __grid_constant__ DevParams<FT> Pshared) {
    using L = Lay<MOIST, PRECIP>;
    using M = Mth<FT>;
    constexpr int NV = L::NV;
    constexpr int NADV = adv_count<MOIST, PRECIP>();
    const int nz = FIXNZ ? FIXNZ : A.nz, C = FIXNZ ? tile_fixed_columns<FT>() : A.C;
    const int NW = FIXNZ ? tile_max_threads<FT>() / tile_fixed_columns<FT>() : A.NW, ncp = A.ncp;
    const int c = threadIdx.x, j = threadIdx.y;
    const int gc = blockIdx.x * C + c;
    const bool active = gc < A.nc;
    // level assignment: in round r thread (c, j) works on level r·NW + j.  A round is a block of NW consecutive
    // levels handled by NW different warps at the same time, so every warp carries one level of the cloudy part of
    // the column and one of the clear part: the work is balanced across warps and there is no serial sweep.
    const int R = (nz + NW - 1) / NW;
    const size_t plane = size_t(nz) * ncp;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [NV][nz + 4][C]: rows nz..nz+3 of every variable plane are the round-boundary stash (old values of the last two
    // levels of the previous block, two parities), so a neighbour in the stash and a neighbour in the tile are reached
    // with the SAME variable stride
    FT* Ys = reinterpret_cast<FT*>(smem_raw);
    FT* rds = Ys + size_t(NV) * (nz + 4) * C;  // [nz][C]
    FT* Ts = rds + size_t(nz) * C;
    FT* nxt = Ts + size_t(nz) * C;
    FT* vt1s = nullptr; FT* vt2s = nullptr; FT* acts = nullptr;
    if constexpr (L::has_pr) { vt1s = nxt; vt2s = nxt + size_t(nz) * C; nxt += size_t(2) * nz * C; }
    if constexpr (L::is2m) { acts = nxt; nxt += size_t(nz) * C; }
    int* cb = reinterpret_cast<int*>(nxt);     // [2][C]
    // parameters: the shared set (kernel argument, constant bank), or this column's own set staged in shared memory
    // as [column][word] with an odd word count (conflict-free when the 32 lanes read the same field)
    const DevParams<FT>* Pp;
    if constexpr (PRMSRC::per_column) {
        size_t off = (reinterpret_cast<size_t>(cb + 2 * C) + 15) & ~size_t(15);
        DevParams<FT>* sp = reinterpret_cast<DevParams<FT>*>(off);
        if (gc < A.nc) {
            const FT* src = reinterpret_cast<const FT*>(A.prm_sets + A.col_to_set[gc]);
            FT* dst = reinterpret_cast<FT*>(sp + c);
            for (int w = j; w < int(sizeof(DevParams<FT>) / sizeof(FT)); w += NW) dst[w] = src[w];
        }
        Pp = sp + c;
    } else {
        // the shared set travels as a kernel argument: it sits in the constant bank (FFMA operands come straight from
        // it) and, unlike a __constant__ symbol, belongs to this launch only — handles never see each other's values
        Pp = &Pshared;
    }
    const DevParams<FT>& P = *Pp;
    // shared-memory indices stay 32-bit: (variable, level) plane stride sV, level stride C
    const int sV = (nz + 4) * C;
    auto sidx = [&](int v, int k) { return v * sV + k * C + c; };
    auto kidx = [&](int k) { return k * C + c; };

    // ---- load the tile (coalesced: lanes = neighbouring columns)
    FT w1c = 0, qsurf = 0, Ndc = 0;
    if (active) {
        for (int k = j; k < nz; k += NW) {
#pragma unroll
            for (int v = 0; v < NV; ++v) Ys[sidx(v, k)] = A.Y[v * plane + size_t(k) * ncp + gc];
            size_t pi = A.prof_per_column ? size_t(k) * ncp + gc : size_t(k);
            rds[kidx(k)] = A.rho_dry[pi];
            Ts[kidx(k)] = A.T[pi];
        }
        w1c = A.w1[gc]; qsurf = A.q_surf[gc]; Ndc = A.Nd[gc];
    }
    if (j == 0) { cb[c] = INT_MAX; cb[C + c] = INT_MAX; }
    __syncthreads();

    const bool k2d = DIAG && (A.model == KID_MODEL_K2D);  // K2D steps through the rhs-only instantiation
    const bool k1d = (A.model == KID_MODEL_K1D) || k2d;   // velocity, activation, condensation switched on
    const bool with_adv = (A.model == KID_MODEL_K1D || A.model == KID_MODEL_K1D_COL_SED) || k2d;
    const bool stale_Naer = (A.model == KID_MODEL_BOX || A.model == KID_MODEL_K1D_COL_SED) && A.aux_Naer != nullptr;
    const FT* naer_p = stale_Naer ? A.aux_Naer + gc : nullptr;  // Box / col_sed: aux N_aer is a snapshot
    const FT dt = A.dt, dz = A.dz;
    const FT inv_dt = FT(1) / dt, inv_dz = FT(1) / dz;
    const int nstage_total = DIAG ? 1 : 3 * A.nsteps;

#pragma unroll 1
    for (int it = 0; it < nstage_total; ++it) {
        const int step = it / 3, stage = DIAG ? 0 : it % 3;
        const int par = it & 1;
        const FT tn = FT(A.t0 + double(step) * double(dt));
        const FT ts = stage_time(stage, tn, dt);
        const FT rk_b = stage == 0 ? FT(1) : (stage == 1 ? FT(0.25) : FT(2) / FT(3));
        const FT rk_a = FT(1) - rk_b, rk_c = rk_b * dt;
        // K1D: ρw(t) uniform over the column (K1DModel/tendency.jl:240-248).  K2D: ρw(x, z_face, t) from the
        // stream function (K2DModel/tendency.jl:4-9), evaluated in double and rounded like the reference's map
        FT rw = (k1d && !k2d) ? rho_w_of_t(ts, w1c, FT(P.kid_t1)) : FT(0);
        double wxd = 0.0;
        if (k2d && active)
            wxd = (ts < FT(P.kid_t1)) ? double(w1c) * sin(2.0 * M_PI / A.width * double(A.xc[gc])) *
                                            sin(M_PI * double(ts) / double(FT(P.kid_t1)))
                                      : 0.0;
        auto rwf = [&](int kface) -> FT {  // ρw at face kface (between cells kface-1 and kface)
            if (!k2d) return rw;
            return FT(wxd * A.zsin[kface]);
        };

        // ---- pass 1 (pointwise): terminal velocities and activation rate of this thread's levels
        if constexpr (L::has_pr) {
            if (active) {
                int first = INT_MAX;
                // Single-evaluation (DIAG) launches — every K2D stage is one — are bound by the latency of the global
                // loads in this loop, not by issue slots: they request the NEXT level's values before evaluating this
                // one.  (The fused step kernel is issue-bound; there the extra registers cost more than they hide.)
                FT nx_ps_l = FT(0), nx_ps_m = FT(0), nx_prev[3] = {FT(0), FT(0), FT(0)};
                auto fetch = [&](int kk) {
                    const size_t ppi = A.prof_per_column ? size_t(kk) * ncp + gc : size_t(kk);
                    nx_ps_l = A.ps_liq[ppi];
                    nx_ps_m = A.ps_mix[ppi];
                    if constexpr (L::is2m) {
                        if (k2d) {
                            const size_t pi = size_t(kk) * ncp + gc;
#pragma unroll
                            for (int q = 0; q < 3; ++q) nx_prev[q] = A.prev_aux[q * plane + pi];
                        }
                    }
                };
                if constexpr (DIAG) { if (j < nz) fetch(j); }
#pragma unroll 1
                for (int k = j; k < nz; k += NW) {
                    FT y[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) y[v] = Ys[sidx(v, k)];
                    Point<FT> a;
                    FT prev0 = FT(0), prev1 = FT(0), prev2 = FT(0);
                    if constexpr (DIAG) {
                        const FT ps_l = nx_ps_l, ps_m = nx_ps_m;
                        prev0 = nx_prev[0]; prev1 = nx_prev[1]; prev2 = nx_prev[2];
                        if (k + NW < nz) fetch(k + NW);
                        thermo_point<FT, MOIST, PRECIP>(P, y, rds[kidx(k)], Ts[kidx(k)], ps_l, ps_m, a);
                    } else {
                        const size_t ppi = A.prof_per_column ? size_t(k) * ncp + gc : size_t(k);
                        thermo_point<FT, MOIST, PRECIP>(P, y, rds[kidx(k)], Ts[kidx(k)], A.ps_liq[ppi], A.ps_mix[ppi], a);
                    }
                    FT v1, v2;
                    terminal_velocities<FT, PRECIP>(P, A.sw, a, v1, v2);
                    vt1s[kidx(k)] = v1;
                    vt2s[kidx(k)] = v2;
                    if constexpr (L::is2m) {
                        FT S_Nl = FT(0);
                        FT q_rai_act = a.q_rai, N_liq_act = a.N_liq, N_rai_act = a.N_rai;
                        if (k2d) {
                            // K2D calls activation BEFORE precompute_aux_precip! (K2DModel/ode_utils.jl:43-44): it
                            // sees the q_rai, N_liq, N_rai the previous rhs! call left in aux
                            const size_t pi = size_t(k) * ncp + gc;
                            q_rai_act = prev0;
                            N_liq_act = prev1;
                            N_rai_act = prev2;
                            if (A.update_prev) {
                                A.prev_aux[pi] = a.q_rai;
                                A.prev_aux[plane + pi] = a.N_liq;
                                A.prev_aux[2 * plane + pi] = a.N_rai;
                            }
                        }
                        // non-local activation keeps only the LOWEST activating level, so once this thread has a
                        // candidate its higher levels can never be selected: skip their ARG evaluation
                        if (k1d && (A.sw.local_activation || first == INT_MAX || k2d)) {
                            FT p = pressure_point<FT, MOIST, PRECIP>(P, a);
                            FT rw_c = k2d ? (rwf(k + 1) + rwf(k)) / FT(2) : rw;  // InterpolateF2C
                            S_Nl = aerosol_activation_rate<FT>(P, A.sw, a.q_tot, a.q_liq + q_rai_act, a.q_ice,
                                                               N_liq_act + N_rai_act, a.N_aer, a.T, a.ps_liq, p, a.rho, rw_c, dt);
                            // find_cloud_base! (src/K1DModel/tendency.jl:21-32): the accumulator starts at
                            // level 1 and is replaced only while its S is 0 and the candidate's S is > 0
                            bool cand = (k == 0) ? (S_Nl != FT(0)) : (S_Nl > FT(0));
                            if (cand && first == INT_MAX) first = k;
                        }
                        acts[kidx(k)] = S_Nl;
                    }
                }
                if constexpr (L::is2m) {
                    if (k1d && !A.sw.local_activation && first != INT_MAX) atomicMin(&cb[par * C + c], first);
                }
            }
            __syncthreads();
        }
        int cbk = INT_MAX;
        if constexpr (L::is2m) {
            cbk = cb[par * C + c];
            if (j == 0) cb[(par ^ 1) * C + c] = INT_MAX;  // next stage's slot; its last readers passed the barrier above
        }
        const FT surf_flux[3] = {FT(0), rw * qsurf, rw * Ndc * (FT(1) - qsurf) / FT(1.225)};  // ρw0 = 0 in K2D

        // ---- pass 2: rounds over blocks of NW levels.  READ phase: every thread gathers the OLD neighbours of its
        // level and turns them into the advective tendency; barrier; WRITE phase: sources + stage combine, in place.
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const int k = r * NW + j;
            const int kb = r * NW;  // first level of this block: lower levels were overwritten in earlier rounds
            const bool valid = active && k < nz;
            FT adv[NADV];
#pragma unroll
            for (int i = 0; i < NADV; ++i) adv[i] = FT(0);
            // u^n of the stage combine and the grid-point constants come from global memory / L2: issue the loads now,
            // the whole READ phase and the barrier hide their latency
            FT un[DIAG ? 1 : NV], ps_l = FT(0), ps_m = FT(0);
            if (valid) {
                if constexpr (!DIAG) {
                    if (stage != 0) {
                        const FT* gU = A.Y + size_t(k) * ncp + gc;
#pragma unroll
                        for (int v = 0; v < NV; ++v) un[v] = gU[v * plane];
                    }
                }
                const size_t ppi = A.prof_per_column ? size_t(k) * ncp + gc : size_t(k);
                ps_l = A.ps_liq[ppi];
                ps_m = A.ps_mix[ppi];
            }
            if (valid && with_adv) {
                // The boundary cells take the stencil of their inner neighbour (Extrapolate: the tendency of cell 0 is
                // that of cell 1, of cell nz-1 that of cell nz-2; the flux correction always extrapolates), so every
                // thread evaluates ONE stencil, centred on kc; a SetValue boundary only swaps the two face fluxes of D.
                const bool is_bot = (k == 0), is_top = (k == nz - 1);
                const int kc = is_bot ? 1 : (is_top ? nz - 2 : k);
                // OLD values: this block and everything above it are still untouched; the two levels below the block
                // were overwritten by the previous round and sit in the stash rows nz + 2·(r & 1) + {0: kb-1, 1: kb-2}
                int row_c = k, row_l = k - 1;
                if (j < 2 || is_top) {  // warp-uniform: the block's first warps (stash below) and the boundary cells
                    auto row = [&](int kk) { return (kk >= kb) ? kk : nz + 2 * (r & 1) + (kb - 1 - kk); };
                    row_c = row(kc);
                    row_l = row(kc - 1);
                }
                KID_ASSERT_ROW(row_c, nz + 4); KID_ASSERT_ROW(row_l, nz + 4); KID_ASSERT_ROW(kc + 1, nz);
                KID_ASSERT_ROW(kc - 1, nz); KID_ASSERT_ROW(c, C);
                const FT* c_p = Ys + row_c * C + c;      // + variable · sV
                const FT* l_p = Ys + row_l * C + c;
                const FT* u_p = Ys + (kc + 1) * C + c;
                // face velocities (air, air - vt1, air - vt2) of faces kc (lower) and kc+1 (upper), pre-multiplied:
                // wh = w / (2 dz) for the centred flux, aw = fcc_scale |w| / dz for the first-order correction
                FT wh_lo[3], wh_up[3], aw_lo[3], aw_up[3];
                {
                    const FT rho_c = rds[kidx(kc)] + c_p[L::tot * sV];
                    const FT rho_l = rds[kidx(kc - 1)] + l_p[L::tot * sV];
                    const FT rho_u = rds[kidx(kc + 1)] + u_p[L::tot * sV];
                    const FT half_inv_dz = FT(0.5) * inv_dz, fs = P.num_fcc_scale * inv_dz;
                    auto face_w = [&](int f, FT rho_a, FT rho_b, FT* wh, FT* aw) {
                        const FT wa = rwf(f) / ((rho_a + rho_b) * FT(0.5));
                        FT wv[3] = {wa, wa, wa};
                        if constexpr (L::has_pr) {
                            wv[1] = wa - (vt1s[kidx(f)] + vt1s[kidx(f - 1)]) * FT(0.5);
                            wv[2] = wa - (vt2s[kidx(f)] + vt2s[kidx(f - 1)]) * FT(0.5);
                        }
#pragma unroll
                        for (int q = 0; q < 3; ++q) { wh[q] = wv[q] * half_inv_dz; aw[q] = M::abs(wv[q]) * fs; }
                    };
                    face_w(kc + 1, rho_u, rho_c, wh_up, aw_up);
                    face_w(kc, rho_c, rho_l, wh_lo, aw_lo);
                }
                // one advected variable: centred flux divergence + first-order correction from the pre-scaled face velocities
                auto stencil_one = [&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr AdvVar e = adv_entry<MOIST, PRECIP>(i);
                    // K2D (K2DModel/tendency.jl:114-213): N_liq and N_aer are not advected, the precipitation
                    // divergence extrapolates at BOTH ends, q_tot always gets the flux correction
                    if (k2d && (e.idx == L::Nl || e.idx == L::Na)) return;
                    const FT yk = c_p[e.idx * sV], yu = u_p[e.idx * sV], yl = l_p[e.idx * sV];
                    const FT Fup = wh_up[e.vel] * (yu + yk), Flo = wh_lo[e.vel] * (yk + yl);  // face fluxes / dz
                    const FT fc = aw_up[e.vel] * (yu - yk) - aw_lo[e.vel] * (yk - yl);
                    const bool fcc_on = (e.fcc == FCC_ALWAYS) || A.sw.qtot_flux_correction || k2d;
                    FT a_i = Flo - Fup;
                    if (fcc_on) a_i += fc;
                    adv[i] = a_i;
                };
                if constexpr (sizeof(FT) == 4) {
                    // Float32: two variables per packed instruction (FADD2 / FFMA2, kid_f2.cuh) — 4 adds + 4 fmas for two
                    // stencils; the per-variable face coefficients are paired (the flux correction switched off = coefficient 0)
                    static_for<0, NADV / 2>([&](auto I) {
                        constexpr int i = 2 * decltype(I)::value, i2 = i + 1;
                        constexpr AdvVar e = adv_entry<MOIST, PRECIP>(i), e2 = adv_entry<MOIST, PRECIP>(i2);
                        const bool f1 = (e.fcc == FCC_ALWAYS) || A.sw.qtot_flux_correction || k2d;
                        const bool f2on = (e2.fcc == FCC_ALWAYS) || A.sw.qtot_flux_correction || k2d;
                        const F2 yk = f2(c_p[e.idx * sV], c_p[e2.idx * sV]), yu = f2(u_p[e.idx * sV], u_p[e2.idx * sV]);
                        const F2 yl = f2(l_p[e.idx * sV], l_p[e2.idx * sV]);
                        const F2 whu = f2(wh_up[e.vel], wh_up[e2.vel]), whl = f2(wh_lo[e.vel], wh_lo[e2.vel]);
                        const F2 awu = f2(f1 ? aw_up[e.vel] : 0.f, f2on ? aw_up[e2.vel] : 0.f);
                        const F2 awl = f2(f1 ? aw_lo[e.vel] : 0.f, f2on ? aw_lo[e2.vel] : 0.f);
                        const F2 r = fma2(whl, yk + yl, fma2(-whu, yu + yk, fma2(awu, yu - yk, -(awl * (yk - yl)))));
                        adv[i] = (k2d && (e.idx == L::Nl || e.idx == L::Na)) ? 0.f : r.x;
                        adv[i2] = (k2d && (e2.idx == L::Nl || e2.idx == L::Na)) ? 0.f : r.y;
                    });
                    if constexpr (NADV % 2 == 1) stencil_one(std::integral_constant<int, NADV - 1>{});
                } else {
                    static_for<0, NADV>(stencil_one);
                }
                // SetValue boundaries (two warps of the tile, warp-uniform branch): one face flux of D is prescribed
                if (is_bot || is_top) {
                    static_for<0, NADV>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr AdvVar e = adv_entry<MOIST, PRECIP>(i);
                        if (k2d && (e.idx == L::Nl || e.idx == L::Na)) return;
                        const int top_bc = (k2d && e.vel != VEL_AIR) ? BC_EXTRAP : e.top;
                        const bool set_bot = is_bot && e.bot == BC_SET, set_top = is_top && top_bc == BC_SET;
                        if (!set_bot && !set_top) return;
                        const FT yk = c_p[e.idx * sV], yu = u_p[e.idx * sV], yl = l_p[e.idx * sV];
                        const FT Fup = wh_up[e.vel] * (yu + yk), Flo = wh_lo[e.vel] * (yk + yl);
                        // bottom cell: (flux at face 1 - surface flux) / dz; top cell: (0 - flux at face nz-1) / dz
                        const FT D = set_bot ? Flo - surf_flux[e.botval] * inv_dz : FT(0) - Fup;
                        adv[i] += (Fup - Flo) - D;  // replace the divergence, keep the flux correction
                    });
                }
                if (k2d) {  // S is DSS'd on its own before it is added to ρq_tot and its own variable
                    static_for<0, NADV>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr AdvVar e = adv_entry<MOIST, PRECIP>(i);
                        if (e.to_tot) A.outS[size_t(e.vel == VEL_VT1 ? 0 : 1) * plane + size_t(k) * ncp + gc] = adv[i];
                    });
                }
                // the next round needs the OLD values of this block's last two levels
                if (kb + NW < nz && j >= NW - 2) {
                    static_for<0, NADV>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr int vi = adv_entry<MOIST, PRECIP>(i).idx;
                        KID_ASSERT_ROW(nz + 2 * ((r + 1) & 1) + (NW - 1 - j), nz + 4);
                        KID_ASSERT_ROW(nz + 2 * ((r + 1) & 1) + (NW - 1 - j) - nz, 4);
                        Ys[vi * sV + (nz + 2 * ((r + 1) & 1) + (NW - 1 - j)) * C + c] = Ys[sidx(vi, k)];
                    });
                }
            }
            __syncthreads();  // every thread of the tile has read its old neighbours

            if (valid) {
                FT y[NV], tend[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) y[v] = Ys[sidx(v, k)];
                Point<FT> a;
                thermo_point<FT, MOIST, PRECIP>(P, y, rds[kidx(k)], Ts[kidx(k)], ps_l, ps_m, a);
                if (naer_p) a.N_aer = naer_p[size_t(k) * ncp];
                FT act_Nl = FT(0);
                if constexpr (L::is2m) {
                    if (k1d) act_Nl = (A.sw.local_activation || k == cbk) ? acts[kidx(k)] : FT(0);
                }
                Sources<FT> src;
                point_source_tendencies<FT, MOIST, PRECIP>(P, A.sw, a, Ndc, inv_dt, k1d, act_Nl, src, tend);
                if (with_adv) {
                    FT S_tot = FT(0);
                    bool any_S = false;
                    static_for<0, NADV>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr AdvVar e = adv_entry<MOIST, PRECIP>(i);
                        if (e.to_tot) {
                            if (k2d) return;
                            tend[e.idx] += adv[i];
                            S_tot = any_S ? S_tot + adv[i] : adv[i];
                            any_S = true;
                        } else {
                            tend[e.idx] += adv[i];
                        }
                    });
                    if (any_S) tend[L::tot] += S_tot;
                }

                if constexpr (DIAG) {
                    if (A.diag_mode == 1) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) A.out[v * plane + size_t(k) * ncp + gc] = tend[v];
                    } else {
                        FT v1 = FT(0), v2 = FT(0);
                        if constexpr (L::has_pr) { v1 = vt1s[kidx(k)]; v2 = vt2s[kidx(k)]; }
                        for (int f = 0; f < A.nfields; ++f)  // one rhs evaluation serves every requested field
                            A.out[size_t(f) * plane + size_t(k) * ncp + gc] = aux_field_value<FT, MOIST, PRECIP>(
                                P, A.fields[f], a, src, v1, v2, act_Nl, A.sw.open_system_activation != 0);
                    }
                } else {
                    // SSPRK33 as a convex combination  u_new = a·u^n + b·(y + dt·f(y)),  (a, b) = (0, 1), (3/4, 1/4),
                    // (1/3, 2/3): OrdinaryDiffEq's (3u + y + dt k)/4 and (u + 2y + 2dt k)/3 with the division folded in
                    FT* gY = A.Y + size_t(k) * ncp + gc;
                    FT yn[NV];
                    if constexpr (sizeof(FT) == 4) {  // two variables per packed instruction
#pragma unroll
                        for (int v = 0; v + 1 < NV; v += 2) {
                            const F2 u2 = (stage == 0) ? f2(y[v], y[v + 1]) : f2(un[v], un[v + 1]);
                            const F2 r = fma2(u2, rk_a, fma2(f2(y[v], y[v + 1]), rk_b, f2(tend[v], tend[v + 1]) * rk_c));
                            yn[v] = r.x; yn[v + 1] = r.y;
                        }
                        if constexpr (NV % 2 == 1) yn[NV - 1] = rk_a * ((stage == 0) ? y[NV - 1] : un[NV - 1]) + (rk_b * y[NV - 1] + rk_c * tend[NV - 1]);
                    } else {
#pragma unroll
                        for (int v = 0; v < NV; ++v) yn[v] = rk_a * ((stage == 0) ? y[v] : un[v]) + (rk_b * y[v] + rk_c * tend[v]);
                    }
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        KID_ASSERT_ROW(k, nz);
                        Ys[sidx(v, k)] = yn[v];
                        if (stage == 2) gY[v * plane] = yn[v];
                    }
                }
            }
        }
        // the next stage's pass 1 only reads this thread's own levels and writes vt/act entries whose readers have all
        // passed the last round's barrier; without a pass 1 the next READ phase needs the new neighbours: barrier
        if constexpr (!L::has_pr) __syncthreads();
    }
}

// ------------------------------------------------------------------ Box model
// One thread per box, the whole state in registers for all nsteps x 3 stages
// (src/BoxModel/ode_utils.jl:10-22: zero -> thermo -> precip -> precip sources).
// Float64: 4 resident blocks (128 registers, 16 warps per SM) instead of the 194 registers / 8 warps the compiler picks on its
// own: the kernel is bound by the latency of dependent double-precision chains (1 M boxes, Precipitation1M: 4.35e9 -> 7.28e9
// box-steps/s with 316 B of spills; Precipitation2M unchanged).  Float32 fits in 90 registers either way.
template <class FT> constexpr int box_min_blocks() { return sizeof(FT) == 8 ? 4 : 1; }
template <class FT, int MOIST, int PRECIP, class PRMSRC>
__global__ void __launch_bounds__(128, box_min_blocks<FT>()) box_step_kernel(const KArgs<FT> A, const __grid_constant__ DevParams<FT> Pshared) {
    using L = Lay<MOIST, PRECIP>;
    constexpr int NV = L::NV;
    const int gc = blockIdx.x * blockDim.x + threadIdx.x;
    extern __shared__ __align__(16) unsigned char box_smem[];
    const DevParams<FT>* Pp;
    if constexpr (PRMSRC::per_column) {
        DevParams<FT>* sp = reinterpret_cast<DevParams<FT>*>(box_smem);
        if (gc < A.nc) sp[threadIdx.x] = A.prm_sets[A.col_to_set[gc]];
        Pp = sp + threadIdx.x;
    } else {
        Pp = &Pshared;
    }
    const DevParams<FT>& P = *Pp;
    if (gc >= A.nc) return;
    const size_t plane = size_t(A.ncp);  // nz == 1
    FT u[NV], y[NV], tend[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) u[v] = A.Y[v * plane + gc];
    const FT rho_dry = A.rho_dry[A.prof_per_column ? gc : 0];
    const FT T = A.T[A.prof_per_column ? gc : 0];
    const FT psl = A.ps_liq[A.prof_per_column ? gc : 0], psm = A.ps_mix[A.prof_per_column ? gc : 0];
    const FT Ndc = A.Nd[gc];
    const FT Naer_aux = A.aux_Naer ? A.aux_Naer[gc] : FT(0);
    const FT dt = A.dt;
#pragma unroll 1
    for (int s = 0; s < A.nsteps; ++s) {
#pragma unroll
        for (int v = 0; v < NV; ++v) y[v] = u[v];
#pragma unroll 1
        for (int stage = 0; stage < 3; ++stage) {
            Point<FT> a;
            thermo_point<FT, MOIST, PRECIP>(P, y, rho_dry, T, psl, psm, a);
            if (A.aux_Naer) a.N_aer = Naer_aux;
            Sources<FT> src;
            point_source_tendencies<FT, MOIST, PRECIP>(P, A.sw, a, Ndc, FT(1) / dt, false, FT(0), src, tend);
#pragma unroll
            for (int v = 0; v < NV; ++v) y[v] = ssprk33_combine(stage, u[v], y[v], dt, tend[v]);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) u[v] = y[v];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) A.Y[v * plane + gc] = u[v];
}

// --------------------------------------------------- time-invariant profile constants
// p_sat,liq(T) and p_sat,mixed(T) per grid point (full-accuracy pow/exp; evaluated once per upload)
template <class FT, class PRMSRC>
__global__ void profile_consts_kernel(const FT* __restrict__ T, FT* __restrict__ ps_liq, FT* __restrict__ ps_mix, int nz, int nc,
                                      int ncp, int per_column, const DevParams<FT>* prm_sets, const int* col_to_set,
                                      const __grid_constant__ DevParams<FT> Pshared) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (c >= (per_column ? nc : 1)) return;
    const DevParams<FT>* Pp;
    if constexpr (PRMSRC::per_column) Pp = prm_sets + col_to_set[c];
    else Pp = &Pshared;
    TD<FT, DevParams<FT>> td(*Pp);
    const size_t i = per_column ? size_t(k) * ncp + c : size_t(k);
    const FT t = T[i];
    ps_liq[i] = td.psat_liquid(t);
    ps_mix[i] = td.psat_mixed(t, td.liquid_fraction_T(t));
}

// ------------------------------------------------------------ layout transposes
// host layout [column][var][level] <-> device layout [var][level][column(ncp)].
// `perm` (device array, may be null): device column i holds host column perm[i] — callers with members in arbitrary order
// get the warp-coherent storage order (model.warp_coherent_order) without gathering on the host.
// The host-layout side may be a device staging buffer or page-locked HOST memory (unified addressing): the kernels then
// move the state over PCIe themselves, every column's (var, level) record in pieces of 128 B (or 512 B: *_wide).
template <class FT>
__global__ void host_to_device_layout(const FT* __restrict__ src, FT* __restrict__ dst, const int* __restrict__ perm, int nc, int ncp,
                                      int nv, int nz) {
    __shared__ FT tile[32][33];
    // rows = (v, k) flattened (R = nv*nz), cols = columns
    const int R = nv * nz;
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int cc = c0 + i, r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (cc < nc && r < R) ? src[size_t(perm ? perm[cc] : cc) * R + r] : FT(0);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, cc = c0 + threadIdx.x;
        if (r < R && cc < ncp) dst[size_t(r) * ncp + cc] = tile[threadIdx.x][i];
    }
}
template <class FT>
__global__ void device_to_host_layout(const FT* __restrict__ src, FT* __restrict__ dst, const int* __restrict__ perm, int nc, int ncp,
                                      int nv, int nz) {
    __shared__ FT tile[32][33];
    const int R = nv * nz;
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, cc = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && cc < ncp) ? src[size_t(r) * ncp + cc] : FT(0);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int cc = c0 + i, r = r0 + threadIdx.x;
        if (cc < nc && r < R) dst[size_t(perm ? perm[cc] : cc) * R + r] = tile[threadIdx.x][i];
    }
}
// Float32, R a multiple of 128: a warp moves 512 contiguous bytes of one column's record (float4 per lane) — the request
// size that matters when the host-layout side is page-locked host memory reached over PCIe.  Tile: 32 columns x 128 rows.
static __global__ void __launch_bounds__(256) host_to_device_layout_wide(const float* __restrict__ src, float* __restrict__ dst,
                                                                  const int* __restrict__ perm, int nc, int ncp, int R) {
    __shared__ float tile[128][33];
    const int r0 = blockIdx.y * 128, c0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = w; i < 32; i += 8) {
        const int cc = c0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cc < nc) v = *reinterpret_cast<const float4*>(src + size_t(perm ? perm[cc] : cc) * R + r0 + 4 * lane);
        tile[4 * lane][i] = v.x; tile[4 * lane + 1][i] = v.y; tile[4 * lane + 2][i] = v.z; tile[4 * lane + 3][i] = v.w;
    }
    __syncthreads();
    for (int r = w; r < 128; r += 8)
        if (c0 + lane < ncp) dst[size_t(r0 + r) * ncp + c0 + lane] = tile[r][lane];
}
static __global__ void __launch_bounds__(256) device_to_host_layout_wide(const float* __restrict__ src, float* __restrict__ dst,
                                                                  const int* __restrict__ perm, int nc, int ncp, int R) {
    __shared__ float tile[128][33];
    const int r0 = blockIdx.y * 128, c0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int r = w; r < 128; r += 8) tile[r][lane] = (c0 + lane < ncp) ? src[size_t(r0 + r) * ncp + c0 + lane] : 0.f;
    __syncthreads();
    for (int i = w; i < 32; i += 8) {
        const int cc = c0 + i;
        if (cc < nc)
            *reinterpret_cast<float4*>(dst + size_t(perm ? perm[cc] : cc) * R + r0 + 4 * lane) =
                make_float4(tile[4 * lane][i], tile[4 * lane + 1][i], tile[4 * lane + 2][i], tile[4 * lane + 3][i]);
    }
}

}  // namespace kid
