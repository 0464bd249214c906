// k1d_pair_kernel — the headline kernel: K1DModel / col_sed, Precipitation2M (SB2006), Float32, 128 levels, one
// parameter set shared by the ensemble (BASELINE configs[3]).  DESIGN.md §3.1a; measurements in profiles/README.md.
//
// Same tile idea as k1d_tile_kernel (kid_kernels.cuh: one CTA owns a tile of columns x all levels in shared memory for all
// SSPRK33 stages of a launch; lanes = neighbouring columns; rounds over blocks of levels with a READ phase that turns the
// OLD neighbours into advective tendencies and a WRITE phase that applies sources + stage combine in place), built for
// fewer issued instructions per grid point and fewer stalls:
//
//  * every thread works on TWO vertically adjacent levels (k0, k0+1) and keeps their arithmetic in the packed
//    Float32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2, kid_f2.cuh, kid_physics_x2.cuh): one issue slot per two
//    results, and all per-thread overhead — addressing, parameter loads, branches, loop control — is paid once per
//    two grid points.  The two cells share their middle face, so the stencil needs 3 face fluxes and 4 loads per
//    variable for 2 cells instead of 4 and 6.  A tile is 16 columns x 128 levels: 256 threads x 128 registers (16
//    level-pairs x 16 columns per round, 4 rounds per stage), two tiles per SM.
//  * all HBM traffic of the stage loop is asynchronous and never passes through registers: u^n of the stage combine is one
//    TMA box per WARP and round (cp.async.bulk.tensor.3d, 16 columns x 4 levels x NV variables, completion on the warp's
//    own mbarrier), requested a round ahead; the new state of a step leaves through the same staging rows as one TMA STORE
//    per warp and round (loads and stores both in the async proxy: no proxy fence on global memory); the grid-point constants
//    come by cp.async during the READ phase.
//  * the round barrier is split: a warp ARRIVES on an mbarrier after its stencil reads and WAITS just before its stores.
//  * there is no separate pass 1: the WRITE phase of stage s evaluates the terminal velocities and the cloud-base search
//    (ARG rate of the candidates) of the NEW state it has just formed, i.e. the aux fields stage s+1 needs
//    (precompute_aux_precip!, precompute_aux_activation!; K1DModel/ode_utils.jl:30-47 evaluates them at the start of
//    the next rhs! call from the same state).  The terminal velocities are two more planes of the shared-memory tile
//    with their own round-boundary stash row; the cloud-base slot carries (level, rate), so the activation source needs
//    no second ARG evaluation.  A launch starts with a pseudo-stage that only does this evaluation.
//
// Time-invariant values per grid point (p_sat,mixed(T), 1 / p_sat,liq(T), T; λ(T) is 1 above freezing and recomputed below;
// G(T) in an array of its own) come from arrays filled at upload (grid_consts_kernel).
#pragma once
#include <cuda.h>  // CUtensorMap

#include "kid_kernels.cuh"
#include "kid_physics_x2.cuh"

namespace kid {


constexpr int PAIR_LPR = 2 * PAIR_TJ;  // levels per round

// shared memory of one tile: state + terminal-velocity planes (NZ levels + 2 stash rows each), ρ_dry, the staging buffers
// of one round (u^n of the stage combine / new state of a step: one TMA box per warp; grid-point constants: cp.async),
// cloud-base slots, the round mbarrier and one mbarrier per warp
template <int MOIST>
inline size_t k1d_pair_smem_bytes() {
    using L = Lay<MOIST, KID_PRECIP_2M>;
    return (size_t(L::NV + 2) * (PAIR_NZ + 2) * PAIR_C + size_t(PAIR_NZ) * PAIR_C + size_t(L::NV) * PAIR_LPR * PAIR_C) * sizeof(float) +
           size_t(PAIR_LPR) * PAIR_C * (sizeof(float2) + sizeof(float)) + size_t(2) * PAIR_C * sizeof(unsigned long long) +
           size_t(1 + PAIR_C * PAIR_TJ / 32) * sizeof(unsigned long long)
#ifdef KID_EXP_SMEM_PAD  // experiment: fewer resident tiles per SM
           + KID_EXP_SMEM_PAD
#endif
        ;
}

// ---- sm_90+/sm_100 asynchronous-copy plumbing (PTX): mbarrier, bulk copy global -> shared (UBLKCP), cp.async (LDGSTS)
namespace pairasync {
KID_DEV unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
KID_DEV void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
KID_DEV void mbar_arrive_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
KID_DEV void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
KID_DEV void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KID_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KID_MBAR_DONE;\n"
        "bra KID_MBAR_WAIT;\n"
        "KID_MBAR_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA: one 3-D box (columns x levels x variables) global -> shared; completion is signalled on the mbarrier
KID_DEV void tma_load_3d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, void* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst_smem)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
KID_DEV void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
KID_DEV void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
KID_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
KID_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// TMA store: one 3-D box shared -> global (bulk async-group completion)
KID_DEV void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(src_smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
KID_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
KID_DEV void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // sources may be reused
KID_DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }            // stores are complete
// generic-proxy writes to SHARED memory before a bulk copy (async proxy) reads them
KID_DEV void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace pairasync

// ARG activation rate of one grid point (get_aerosol_activation_rate, K1DModel/tendency.jl:35-96).  Rare (cloud-base
// candidates only) and large: kept out of line so that the hot loop stays small.
template <bool NEQ>
__device__ __noinline__ float pair_activation_rate(const DevParams<float>& P, const DevSwitches& sw, float T, float ps_l, float rho,
                                                   float q_tot, float q_liq, float q_ice, float q_rai, float q_sno, float N_liq,
                                                   float N_rai, float N_aer, float rw, float dt) {
    TD<float, DevParams<float>> td(P);
    const float p = NEQ ? td.air_pressure(T, rho, q_tot, q_liq + q_rai, q_ice + q_sno) : td.air_pressure(T, rho, q_tot, 0.f, 0.f);
    return aerosol_activation_rate<float>(P, sw, q_tot, q_liq + q_rai, q_ice, N_liq + N_rai, N_aer, T, ps_l, p, rho, rw, dt);
}

// per-grid-point constants (see GridConst2) from the profile constants of the scalar kernels (profile_consts_kernel: the
// same p_sat values), λ(T) as TD<float> evaluates it, and G(T) in double
template <class PRMSRC>
__global__ void grid_consts_kernel(const float* __restrict__ T, const float* __restrict__ ps_liq, const float* __restrict__ ps_mix,
                                   float4* __restrict__ out, float* __restrict__ outG, int nz, int nc, int ncp, int per_column,
                                   const __grid_constant__ DevParams<float> P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (c >= (per_column ? nc : 1)) return;
    const size_t i = per_column ? size_t(k) * ncp + c : size_t(k);
    const float t = T[i];
    TD<float, DevParams<float>> tdf(P);
    TD<double, DevParams<float>> td(P);
    out[i] = make_float4(ps_mix[i], rcp1(ps_liq[i]), t, tdf.liquid_fraction_T(t));  // (ps_mix, inv_psl, T, lam); lam is not staged
    const double td_ = double(t);
    outG[i] = float(td.G_func(td_, td.latent_heat_vapor(td_), td.psat_liquid(td_)));
}

#ifdef KID_EXP_TIMING
// experiment build (scripts/build_variant.sh timing -DKID_EXP_TIMING): cycles the warps spend between phases, per round
__device__ unsigned long long g_pair_wait[8][8];    // [warp][round]: cycles in the round-barrier wait
__device__ unsigned long long g_pair_work[8][8];    // [warp][round]: cycles outside the round-barrier wait
__device__ unsigned long long g_pair_cycles[8][8];  // [barrier wait | READ | sources | act + tendencies | u^n wait | combine + stores | prefetch issue | post][round]
#endif

template <int MOIST>
__global__ void __launch_bounds__(PAIR_C * PAIR_TJ, PAIR_CTAS_PER_SM)
k1d_pair_kernel(const KArgs<float> A, const float4* __restrict__ gconst, const float* __restrict__ gG, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ DevParams<float> P) {
    using L = Lay<MOIST, KID_PRECIP_2M>;
    using namespace pairasync;
    constexpr int NV = L::NV, NZ = PAIR_NZ, C = PAIR_C, TJ = PAIR_TJ;
    constexpr int LPR = PAIR_LPR, R = NZ / LPR, NR = NZ + 2;  // levels per round, rounds per stage, rows per plane (2 stash rows)
    constexpr int VT1 = NV, VT2 = NV + 1, NP = NV + 2;         // planes: state variables, vt of ρq_rai, vt of N_rai
    constexpr int sV = NR * C;
    constexpr int HO = (NZ / 2) * C;  // a plane stores its even levels first, then the odd ones: offset of level k0 + 1 from level k0
    constexpr int NADV = adv_count<MOIST, KID_PRECIP_2M>();
    constexpr int NWARP = C * TJ / 32, WROWS = 32 / C, LW = 2 * WROWS;  // warps per tile; pair-rows and levels of one warp per round
    constexpr unsigned UN_BYTES = NV * LW * C * sizeof(float);           // u^n box of one warp
    static_assert(32 % C == 0 && TJ % WROWS == 0, "a warp is a whole number of pair-rows");
    static_assert(R >= 2 && NZ % LPR == 0, "a stage is R >= 2 whole rounds");
    // R >= 3: a block's WRITE phase and its next readers (the READ phase of the block below, one stage later) are always
    // separated by a round barrier; with R = 2 the last block is read right at the start of the next stage: extra barrier
    constexpr bool STAGE_BARRIER = R < 3;
    const int c = threadIdx.x, j = threadIdx.y;
    const int tid = j * C + c;
    const int lane = tid & 31, warp = tid >> 5, jj = j % WROWS;  // jj: this thread's pair-row inside its warp
    const int gc0 = blockIdx.x * C, gc = gc0 + c;
    const bool active = gc < A.nc;
    const int ncp = A.ncp;
    const size_t plane = size_t(NZ) * ncp;  // (tile load)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Ys = reinterpret_cast<float*>(smem_raw);    // [NP][NR][C]; rows NZ, NZ+1: old values of the level below a block (2 parities)
    float* rds = Ys + NP * sV;                         // [NZ][C] ρ_dry
    float* uns = rds + NZ * C + warp * (NV * LW * C);  // [NWARP][NV][LW][C] u^n of this round's levels: one TMA box per warp
    // staging slots of a thread: row j (lower level of its pair) and row TJ + j (upper level): the two pair-rows of a warp are
    // neighbouring rows, so a warp reads 32 consecutive slots
    float2* gca = reinterpret_cast<float2*>(rds + NZ * C + NV * LPR * C);  // [LPR][C] (ps_mix, 1/ps_liq) of this round's levels (cp.async)
    float* gcb = reinterpret_cast<float*>(gca + LPR * C);        // [LPR][C] T
    // [2][C] cloud base of the column, two stages in flight: (lowest activating level << 32) | bits of its ARG rate, so that an
    // atomicMin keeps the lowest level together with ITS rate and the activation source needs no second ARG evaluation
    unsigned long long* cb = reinterpret_cast<unsigned long long*>(gcb + LPR * C);
    unsigned long long* mbar_round = cb + 2 * C;  // round barrier: one arrival per warp
    unsigned long long* mbar = mbar_round + 1 + warp;                                     // completion of this warp's u^n box

    // row of level k inside a plane: even levels first, then the odd ones.  The pair rows (k0, k0 + 1) of consecutive
    // threadIdx.y are then contiguous, so a warp that spans two pair-rows (C = 16) still reads conflict-free.
    auto lrow = [](int k) { return (k & 1) * (NZ / 2) + (k >> 1); };
    float w1c = 0, qsurf = 0, Ndc = 0;
    if (active) {
        for (int k = j; k < NZ; k += TJ) {
#pragma unroll
            for (int v = 0; v < NV; ++v) Ys[v * sV + lrow(k) * C + c] = A.Y[v * plane + size_t(k) * ncp + gc];
            rds[lrow(k) * C + c] = A.rho_dry[A.prof_per_column ? size_t(k) * ncp + gc : size_t(k)];
        }
        w1c = A.w1[gc]; qsurf = A.q_surf[gc]; Ndc = A.Nd[gc];
    }
    constexpr unsigned long long CB_NONE = ~0ull;
    if (j == 0) { cb[c] = CB_NONE; cb[C + c] = CB_NONE; }
    if (tid == 0) {
        mbar_init(mbar_round, NWARP);
        for (int w = 0; w < NWARP; ++w) mbar_init(mbar_round + 1 + w, 1);
    }
    __syncthreads();

    const DevSwitches sw = A.sw;
    const bool k1d = (A.model == KID_MODEL_K1D);  // velocity, activation, condensation on; col_sed: sedimentation only
    const bool loc = sw.local_activation != 0;
    const bool stale_Naer = (A.model == KID_MODEL_K1D_COL_SED) && A.aux_Naer != nullptr;
    const float dt = A.dt, inv_dt = 1.f / dt, inv_dz = 1.f / A.dz;
    const float half_inv_dz = 0.5f * inv_dz, fs = P.num_fcc_scale * inv_dz;
    const float closed = sw.open_system_activation ? 0.f : -1.f;
    // The ARG evaluation is gated by the supersaturation.  The cheap estimate from the staged constants only screens (with a
    // margin); the sign that decides — and with it the cloud-base level — is taken by aerosol_activation_rate from the same
    // formula as everywhere else (TD.supersaturation with T and p_sat(T)).
    constexpr float S_SCREEN = -1e-4f;
    const int nst = 3 * A.nsteps;
    int first = INT_MAX;      // lowest activation candidate among this thread's levels, for the stage being prepared
    unsigned un_phase = 0;    // parity of the mbarrier phase this warp's next u^n fetch completes
    unsigned round_phase = 0; // parity of the round barrier's current phase

    // ARG rate of component e of the pair (scalar, rare)
    auto arg_rate = [&](const PointX2& a, int e, int k, float rwx) -> float {
        const size_t pi = A.prof_per_column ? size_t(k) * ncp + gc : size_t(k);
        return pair_activation_rate<L::neq>(P, sw, A.T[pi], A.ps_liq[pi], e ? a.rho.y : a.rho.x, e ? a.q_tot.y : a.q_tot.x,
                                            e ? a.q_liq.y : a.q_liq.x, e ? a.q_ice.y : a.q_ice.x, e ? a.q_rai.y : a.q_rai.x,
                                            e ? a.q_sno.y : a.q_sno.x, e ? a.N_liq.y : a.N_liq.x, e ? a.N_rai.y : a.N_rai.x,
                                            e ? a.N_aer.y : a.N_aer.x, rwx, dt);
    };

#pragma unroll 1
    for (int it = -1; it < nst; ++it) {
        const bool real = it >= 0;                 // it = -1: only the aux fields of the first stage are evaluated
        const int stage = real ? it % 3 : 0;
        const int itn = it + 1, stage_n = itn % 3;
        const bool do_post = itn < nst;
        const bool fetch_un = real && stage != 0;
        const int parn = itn & 1, par = parn ^ 1;
        const float tn = float(A.t0 + double(real ? it / 3 : 0) * double(dt));
        const float ts = stage_time(stage, tn, dt);
        const float tsn = stage_time(stage_n, float(A.t0 + double(itn / 3) * double(dt)), dt);
        const float rk_b = stage == 0 ? 1.f : (stage == 1 ? 0.25f : 2.f / 3.f);
        const float rk_a = 1.f - rk_b, rk_c = rk_b * dt;
        const float rw = k1d ? rho_w_of_t(ts, w1c, P.kid_t1) : 0.f;
        const float rw_n = k1d ? rho_w_of_t(tsn, w1c, P.kid_t1) : 0.f;
        const float rw2 = 2.f * rw;
        const float surf_flux[3] = {0.f, rw * qsurf, rw * Ndc * (1.f - qsurf) / 1.225f};
        const bool fcc_tot = sw.qtot_flux_correction != 0;
        unsigned cbk = ~0u;   // cloud-base level of this stage (none: all ones)
        float cb_rate = 0.f;  // its ARG rate

#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const int k0 = r * LPR + 2 * j;  // this thread's pair: levels k0, k0 + 1
            float* pa = Ys + (k0 >> 1) * C + c;  // level k0 (even); level k0 + 1 at + HO; + plane · sV
            const float* rda = rds + (k0 >> 1) * C + c;
            F2 adv[NADV];
#ifdef KID_EXP_TIMING
            const long long tc0 = clock64();
#endif
            if (active) {
                {   // grid-point constants of this thread's pair -> its own two staging slots (consumed after the barrier)
                    const size_t i0 = A.prof_per_column ? size_t(k0) * ncp + gc : size_t(k0);
                    const float* g0 = reinterpret_cast<const float*>(gconst + i0);
                    const float* g1 = reinterpret_cast<const float*>(gconst + i0 + (A.prof_per_column ? size_t(ncp) : size_t(1)));
                    cp_async8(gca + j * C + c, g0);
                    cp_async8(gca + (TJ + j) * C + c, g1);
                    cp_async4(gcb + j * C + c, g0 + 2);
                    cp_async4(gcb + (TJ + j) * C + c, g1 + 2);
                    cp_async_commit();
                }
                if (real) {
                    // ---- READ phase.  OLD values of levels k0-1 .. k0+2: the level below a block was overwritten by the
                    // previous round and sits in the stash row; at the domain boundaries the row is clamped (the value is
                    // not used: the boundary cells are fixed up below).
                    const bool bot = (k0 == 0), top = (k0 == NZ - 2);
                    const float* pm = (j == 0) ? (r == 0 ? pa : Ys + (NZ + (r & 1)) * C + c) : pa + HO - C;  // level k0 - 1 (odd)
                    const float* pu = top ? pa + HO : pa + C;                                               // level k0 + 2 (even)
                    const float* rdm = bot ? rda : rda + HO - C;
                    const float* rdu = top ? rda + HO : rda + C;
                    // faces: f0 = (k0-1 | k0), f1 = (k0 | k0+1), f2 = (k0+1 | k0+2); f0 and f2 packed, f1 scalar.
                    // Face velocities (air, air - vt_q, air - vt_N) pre-multiplied: wh = w / (2 dz) for the centred flux,
                    // aw = fcc_scale |w| / dz for the first-order correction; Q_f = wh (y_hi + y_lo) - aw (y_hi - y_lo) is
                    // the flux through face f (over dz), and the tendency of a cell is Q_lower - Q_upper.
                    F2 wh02[3], aw02[3];
                    float wh1[3], aw1[3];
                    {
                        const float rho_m = rdm[0] + pm[L::tot * sV], rho_a = rda[0] + pa[L::tot * sV];
                        const float rho_b = rda[HO] + pa[L::tot * sV + HO], rho_u = rdu[0] + pu[L::tot * sV];
                        const F2 wa02 = rw2 * vrcp(f2(rho_m, rho_b) + f2(rho_a, rho_u));
                        const float wa1 = rw2 * rcp1(rho_a + rho_b);
                        const F2 v1lo = f2(pm[VT1 * sV], pa[VT1 * sV + HO]), v1hi = f2(pa[VT1 * sV], pu[VT1 * sV]);
                        const F2 v2lo = f2(pm[VT2 * sV], pa[VT2 * sV + HO]), v2hi = f2(pa[VT2 * sV], pu[VT2 * sV]);
                        const F2 wv02[3] = {wa02, fma2(v1lo + v1hi, -0.5f, wa02), fma2(v2lo + v2hi, -0.5f, wa02)};
                        const float wv1[3] = {wa1, fmaf(v1hi.x + v1lo.y, -0.5f, wa1), fmaf(v2hi.x + v2lo.y, -0.5f, wa1)};
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            wh02[q] = wv02[q] * half_inv_dz; aw02[q] = vabs(wv02[q]) * fs;
                            wh1[q] = wv1[q] * half_inv_dz; aw1[q] = fabsf(wv1[q]) * fs;
                        }
                    }
                    static_for<0, NADV>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr AdvVar e = adv_entry<MOIST, KID_PRECIP_2M>(i);
                        const bool fcc_on = (e.fcc == FCC_ALWAYS) || fcc_tot;
                        const float ym = pm[e.idx * sV], ya = pa[e.idx * sV], yb = pa[e.idx * sV + HO], yu = pu[e.idx * sV];
                        const F2 lo = f2(ym, yb), hi = f2(ya, yu);
                        const F2 G02 = fcc_on ? aw02[e.vel] * (hi - lo) : f2(0.f);
                        const float G1 = fcc_on ? aw1[e.vel] * (yb - ya) : 0.f;
                        const F2 F02 = wh02[e.vel] * (hi + lo);
                        const float F1 = wh1[e.vel] * (ya + yb);
                        const F2 Q02 = F02 - G02;
                        const float Q1 = F1 - G1;
                        adv[i] = f2(Q02.x - Q1, Q1 - Q02.y);
                        // domain boundaries (one warp each, warp-uniform).  Extrapolate: the boundary cell takes the tendency of
                        // its inner neighbour; SetValue: the boundary face flux of the divergence is prescribed, the flux
                        // correction still extrapolates (DivergenceF2C / FluxCorrectionC2C, K1DModel/tendency.jl:274-435).
                        if (bot) adv[i].x = (e.bot == BC_SET) ? (surf_flux[e.botval] * inv_dz - F1) + (G02.y - G1) : adv[i].y;
                        if (top) adv[i].y = (e.top == BC_SET) ? F1 + (G1 - G02.x) : adv[i].x;
                    });
                    // the next round needs the OLD values of this block's last level (state and terminal velocities)
                    if (j == TJ - 1 && r + 1 < R) {
#pragma unroll
                        for (int p = 0; p < NP; ++p) Ys[p * sV + (NZ + ((r + 1) & 1)) * C + c] = pa[p * sV + HO];
                    }
                }
            }
            // Split round barrier.  No thread may overwrite its rows before every thread of the tile has read its old
            // neighbours — but the rows are only written at the END of the WRITE phase, after the sources.  So a warp ARRIVES
            // here (non-blocking) and WAITS just before its stores: by then the other warps have long arrived, and nobody idles
            // at a barrier while the slowest warp of the round finishes.
            __syncwarp();
            if (lane == 0) mbar_arrive(mbar_round);
#ifdef KID_EXP_TIMING
            const long long tc1 = clock64();
#endif

            // ---- WRITE phase: sources + SSPRK33 stage combine in place, then the aux fields of the next stage
            F2 y[NV];
            F2 rd = f2(1.f);
            GridConst2 g;
            if (active) {
#pragma unroll
                for (int v = 0; v < NV; ++v) y[v] = f2(pa[v * sV], pa[v * sV + HO]);
                rd = f2(rda[0], rda[HO]);
                cp_async_wait_all();
                const float2 a0 = gca[j * C + c], a1 = gca[(TJ + j) * C + c];
                g.ps_mix = f2(a0.x, a1.x); g.inv_psl = f2(a0.y, a1.y);
                g.T = f2(gcb[j * C + c], gcb[(TJ + j) * C + c]);
                g.lam = liquid_fraction_x2(P, g.T);
                // G(T): only rain evaporation reads it; straight from global memory inside that branch
                const size_t i0 = A.prof_per_column ? size_t(k0) * ncp + gc : size_t(k0);
                g.G0 = gG + i0;
                g.G1 = gG + i0 + (A.prof_per_column ? size_t(ncp) : size_t(1));
            }
            PointX2 a;
            Src2M s;
            F2 cs_liq = f2(0.f), cs_ice = f2(0.f);
            if (real && active) {
                thermo_point_x2<MOIST>(P.td_R_v, y, rd, g, a);
                if (stale_Naer) a.N_aer = f2(A.aux_Naer[size_t(k0) * ncp + gc], A.aux_Naer[size_t(k0 + 1) * ncp + gc]);
                if constexpr (L::neq) {
                    if (k1d) cloud_sources_x2(P.td_R_v, P.inv_tau_relax_liq, P.inv_tau_relax_ice, a, g, cs_liq, cs_ice);
                }
                precip_sources_2m_x2(P, sw, a, g, inv_dt, s);
            }
            // every warp of the tile has read its old neighbours (arrived above): the rows may be overwritten now
#ifdef KID_EXP_TIMING
            const long long tc2 = clock64();
#endif
            mbar_wait(mbar_round, round_phase);
#ifdef KID_EXP_TIMING
            const long long tc3 = clock64();
            long long tcA = tc3, tcB = tc3, tcC = tc3, tcD = tc3;
#endif
            round_phase ^= 1u;
            if (r == 0) {  // complete: its last atomicMin precedes the owner's arrival at this round's barrier
                const unsigned long long v = cb[par * C + c];
                cbk = unsigned(v >> 32);
                cb_rate = __uint_as_float(unsigned(v));
            }
            // The slot just consumed serves the NEXT stage's candidates: every thread read it after the wait of round 0, i.e.
            // before it arrived at round 1's barrier, whose wait this thread has passed; the next writers (atomicMin in the
            // next stage) have passed the wait of this stage's LAST round, for which this warp must have arrived (R >= 3;
            // R = 2: the stage barrier)
            if (r == 1 && j == 0) cb[par * C + c] = CB_NONE;
            if (real && active) {
                // activation source: the rate at the cloud-base level found one stage ago (or at every level, local_activation)
                F2 act = f2(0.f);
                if (k1d) {
                    if (loc) {
                        const F2 S = supersaturation_x2(P.td_R_v, a.rho, a.q_tot, a.q_liq + a.q_rai, a.q_ice, g);
                        if (!(S.x < S_SCREEN) && !(a.N_aer.x < 1.1920929e-07f)) act.x = arg_rate(a, 0, k0, rw);
                        if (!(S.y < S_SCREEN) && !(a.N_aer.y < 1.1920929e-07f)) act.y = arg_rate(a, 1, k0 + 1, rw);
                    } else if ((cbk >> 1) == unsigned(k0 >> 1)) {
                        // the rate the cloud-base search of the previous stage's post phase found for this level, from the
                        // same state and the same ρw: no second evaluation
                        if (cbk & 1u) act.y = cb_rate; else act.x = cb_rate;
                    }
                }
                F2 tend[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) tend[v] = f2(0.f);
                if constexpr (L::neq) {
                    if (k1d) {
                        tend[L::liq] = a.rho * cs_liq;
                        tend[L::ice] = a.rho * cs_ice;
                    }
                }
                // precip_sources_tendency! (src/Common/tendency.jl:787-805)
                tend[L::rai] = a.rho * s.s_rai;
                tend[L::Na] = fma2(act, closed, s.s_Na);
                tend[L::Nl] = s.s_Nl + act;
                tend[L::Nr] = s.s_Nr;
                if constexpr (L::neq) tend[L::liq] = fma2(a.rho, s.s_liq, tend[L::liq]);
                static_for<0, NADV>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr AdvVar e = adv_entry<MOIST, KID_PRECIP_2M>(i);
                    tend[e.idx] += adv[i];
                    if (e.to_tot) tend[L::tot] += adv[i];  // precipitation also moves ρq_tot (K1DModel/tendency.jl:431-432)
                });
                // SSPRK33 as a convex combination  u_new = a·u^n + b·(y + dt·f(y))
#ifdef KID_EXP_TIMING
                tcA = clock64();
#endif
                if (fetch_un) { mbar_wait(mbar, un_phase); un_phase ^= 1u; }  // one phase of the warp's mbarrier per fetched box
#ifdef KID_EXP_TIMING
                tcB = clock64();
#endif
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float* un0 = uns + (v * LW + 2 * jj) * C + c;  // this thread's two slots of the warp's box
                    const F2 yb = fma2(y[v], rk_b, tend[v] * rk_c);
                    if (stage == 0) y[v] = yb;
                    else y[v] = fma2(f2(un0[0], un0[C]), rk_a, yb);
                    pa[v * sV] = y[v].x;
                    pa[v * sV + HO] = y[v].y;
                    // end of a step: the new state leaves through the same box (TMA store below)
                    if (stage == 2) { un0[0] = y[v].x; un0[C] = y[v].y; }
                }
            } else if (fetch_un) {
                un_phase ^= 1u;  // columns beyond the ensemble: keep the phase count of the warp's fetches
                if (stage == 2) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) { uns[(v * LW + 2 * jj) * C + c] = 0.f; uns[(v * LW + 2 * jj + 1) * C + c] = 0.f; }
                }
            }
            // End of a step: the warp's (C columns x LW levels x NV variables) box of the new state -> device state with ONE TMA
            // store.  State stores and u^n loads both go through the async proxy, so no generic <-> async proxy fence on global
            // memory is needed (fence.proxy.async after the stores cost 10 % of the kernel); only the shared-memory writes
            // above are fenced for the bulk copy.
            if (real && stage == 2) {
                fence_proxy_async_shared();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tmY, gc0, r * LPR + warp * LW, 0, uns);
                    bulk_commit();
                }
            }
#ifdef KID_EXP_TIMING
            tcC = clock64();
#endif
            if (do_post && active) {
                // aux of the state just formed, as the next rhs! call starts with (precompute_aux_thermo!,
                // precompute_aux_precip!, precompute_aux_activation!: K1DModel/ode_utils.jl:33-38)
                thermo_point_x2<MOIST>(P.td_R_v, y, rd, g, a);
                F2 vq, vN;
                terminal_velocities_x2(P, sw, a, vq, vN);
                pa[VT1 * sV] = vq.x; pa[VT1 * sV + HO] = vq.y;
                pa[VT2 * sV] = vN.x; pa[VT2 * sV + HO] = vN.y;
                if (k1d && !loc) {
                    // find_cloud_base! (K1DModel/tendency.jl:21-32): the lowest level with a positive ARG rate; the
                    // accumulator starts at level 1 and is replaced only while its S is 0 and the candidate's S is > 0 (so a
                    // NEGATIVE rate at level 1 sticks).  Once this thread has a candidate its higher levels are skipped.
                    // A candidate is published at once, and a level above an already published candidate of the column (any
                    // thread's; the rounds climb) can never be the lowest: its ARG evaluation is skipped.  A stale read of the
                    // slot only costs an evaluation.
                    const unsigned cb_seen = reinterpret_cast<volatile unsigned*>(&cb[parn * C + c])[1];  // level (high word)
                    if (first == INT_MAX && unsigned(k0) < cb_seen) {
                        const F2 S = supersaturation_x2(P.td_R_v, a.rho, a.q_tot, a.q_liq + a.q_rai, a.q_ice, g);
                        float rate = 0.f;
                        if (!(S.x < S_SCREEN) && !(a.N_aer.x < 1.1920929e-07f)) {
                            rate = arg_rate(a, 0, k0, rw_n);
                            if ((k0 == 0) ? (rate != 0.f) : (rate > 0.f)) first = k0;
                        }
                        if (first == INT_MAX && !(S.y < S_SCREEN) && !(a.N_aer.y < 1.1920929e-07f)) {
                            rate = arg_rate(a, 1, k0 + 1, rw_n);
                            if (rate > 0.f) first = k0 + 1;
                        }
                        if (first != INT_MAX)
                            atomicMin(&cb[parn * C + c], (static_cast<unsigned long long>(first) << 32) | __float_as_uint(rate));
                    }
                    if (r == R - 1) first = INT_MAX;
                }
            }
#ifdef KID_EXP_TIMING
            tcD = clock64();
#endif
            // u^n of this warp's levels in the NEXT round that combines with it (stages 2 and 3 of a step) — a (C columns x LW
            // levels x NV variables) box of the device state — straight into the warp's staging rows with ONE TMA instruction:
            // it lands while the next round's stencil and sources are evaluated (requested after the post phase, so that a step's
            // store of the same box has long been read).  The box is private to the warp, so the only
            // ordering needed is that all its lanes are done with the previous box.
            {
                const bool last = (r + 1 == R);
                const bool next_fetch = last ? (do_post && stage_n != 0) : fetch_un;
                if (next_fetch) {
                    __syncwarp();
                    if (lane == 0) {
                        // the box is the source of this round's store (stage 3): wait until the store has read it; the
                        // step's first u^n box waits for the previous step's stores of this warp to be complete
                        if (stage == 2) bulk_wait_read_all();
                        if (last && stage_n == 1) bulk_wait_all();
                        mbar_arrive_expect_tx(mbar, UN_BYTES);
                        tma_load_3d(uns, &tmY, gc0, (last ? 0 : r + 1) * LPR + warp * LW, 0, mbar);
                    }
                }
            }
#ifdef KID_EXP_TIMING
            if (lane == 0 && real) {
                const long long tc4 = clock64();
                atomicAdd(&g_pair_cycles[0][r], (unsigned long long)(tc3 - tc2));
                atomicAdd(&g_pair_cycles[1][r], (unsigned long long)(tc1 - tc0));
                atomicAdd(&g_pair_cycles[2][r], (unsigned long long)(tc2 - tc1));
                atomicAdd(&g_pair_cycles[3][r], (unsigned long long)(tcA - tc3));
                atomicAdd(&g_pair_cycles[4][r], (unsigned long long)(tcB - tcA));
                atomicAdd(&g_pair_cycles[5][r], (unsigned long long)(tcC - tcB));
                atomicAdd(&g_pair_cycles[7][r], (unsigned long long)(tcD - tcC));
                atomicAdd(&g_pair_cycles[6][r], (unsigned long long)(tc4 - tcD));
                atomicAdd(&g_pair_wait[warp & 7][r], (unsigned long long)(tc3 - tc2));
                atomicAdd(&g_pair_work[warp & 7][r], (unsigned long long)((tc2 - tc0) + (tc4 - tc3)));
            }
#endif
        }
        if constexpr (STAGE_BARRIER) __syncthreads();
    }
    if (lane == 0) bulk_wait_all();  // the last step's stores read this CTA's shared memory
}

}  // namespace kid
