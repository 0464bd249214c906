// Kernel instantiations for ONE precipitation style (compile with -DKID_PRECIP=0..3).
#ifndef KID_PRECIP
#error "compile with -DKID_PRECIP=<enum kid_precip value>"
#endif
#include "kid_launch.h"
#if KID_PRECIP == 3
#include "kid_pair_kernel.cuh"
#endif

namespace kid {
namespace {

// parameter source tags: one set shared by all columns (kernel argument) or per-column sets (shared memory)
template <class FT> struct ConstPrm { static constexpr bool per_column = false; };
template <class FT> struct ColumnPrm { static constexpr bool per_column = true; };

template <class FT, int MOIST, bool DIAG, class PRM>
cudaError_t launch_tile_p(const KArgs<FT>& a, const DevParams<FT>& P, cudaStream_t s) {
    // the default KiD grid (128 levels, full tile) runs the instantiation with compile-time tile dimensions
    constexpr int FZ = 128;
    const bool fixed = !DIAG && !PRM::per_column && !a.no_fixed_shape && a.nz == FZ && a.C == tile_fixed_columns<FT>() &&
                       a.NW == tile_max_threads<FT>() / tile_fixed_columns<FT>();
    auto kern = k1d_tile_kernel<FT, MOIST, KID_PRECIP, DIAG, PRM, 0>;
    if constexpr (!DIAG && !PRM::per_column) {
        if (fixed) kern = k1d_tile_kernel<FT, MOIST, KID_PRECIP, DIAG, PRM, FZ>;
    }
    size_t smem = k1d_tile_smem_bytes<FT, MOIST, KID_PRECIP>(a.nz, a.C, PRM::per_column);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    dim3 block(a.C, a.NW), grid((a.nc + a.C - 1) / a.C);
    kern<<<grid, block, smem, s>>>(a, P);
    return cudaGetLastError();
}
template <class FT, int MOIST, bool DIAG>
cudaError_t launch_tile(const KArgs<FT>& a, const DevParams<FT>& P, cudaStream_t s) {
    return a.prm_sets ? launch_tile_p<FT, MOIST, DIAG, ColumnPrm<FT>>(a, P, s) : launch_tile_p<FT, MOIST, DIAG, ConstPrm<FT>>(a, P, s);
}
template <class FT>
cudaError_t k1d(int moisture, bool diag, const KArgs<FT>& a, const DevParams<FT>& P, cudaStream_t s) {
    if (moisture == KID_MOIST_EQ) return diag ? launch_tile<FT, KID_MOIST_EQ, true>(a, P, s) : launch_tile<FT, KID_MOIST_EQ, false>(a, P, s);
    return diag ? launch_tile<FT, KID_MOIST_NONEQ, true>(a, P, s) : launch_tile<FT, KID_MOIST_NONEQ, false>(a, P, s);
}
template <class FT, int MOIST, class PRM>
cudaError_t box_p(const KArgs<FT>& a, const DevParams<FT>& P, cudaStream_t s) {
    const int bs = PRM::per_column ? 64 : 128;
    const size_t smem = PRM::per_column ? size_t(bs) * sizeof(DevParams<FT>) : 0;
    auto kern = box_step_kernel<FT, MOIST, KID_PRECIP, PRM>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
    }
    dim3 block(bs), grid((a.nc + bs - 1) / bs);
    kern<<<grid, block, smem, s>>>(a, P);
    return cudaGetLastError();
}
template <class FT>
cudaError_t box(int moisture, const KArgs<FT>& a, const DevParams<FT>& P, cudaStream_t s) {
    if (moisture == KID_MOIST_EQ)
        return a.prm_sets ? box_p<FT, KID_MOIST_EQ, ColumnPrm<FT>>(a, P, s) : box_p<FT, KID_MOIST_EQ, ConstPrm<FT>>(a, P, s);
    return a.prm_sets ? box_p<FT, KID_MOIST_NONEQ, ColumnPrm<FT>>(a, P, s) : box_p<FT, KID_MOIST_NONEQ, ConstPrm<FT>>(a, P, s);
}
template <class FT>
cudaError_t consts(const KArgs<FT>& a, const DevParams<FT>& P, FT* ps_liq, FT* ps_mix, cudaStream_t s) {
    const int ncol = a.prof_per_column ? a.nc : 1;
    dim3 block(128), grid((ncol + 127) / 128, a.nz);
    if (a.prm_sets && a.prof_per_column)
        profile_consts_kernel<FT, ColumnPrm<FT>><<<grid, block, 0, s>>>(a.T, ps_liq, ps_mix, a.nz, a.nc, a.ncp, 1, a.prm_sets, a.col_to_set, P);
    else
        profile_consts_kernel<FT, ConstPrm<FT>><<<grid, block, 0, s>>>(a.T, ps_liq, ps_mix, a.nz, a.nc, a.ncp, a.prof_per_column, nullptr, nullptr, P);
    return cudaGetLastError();
}
template <class FT>
cudaError_t obs(int moisture, const ObsArgs<FT>& a, const DevParams<FT>& P, cudaStream_t s) {
    dim3 block(128), grid((a.nc + 127) / 128, a.n_out);
    if (moisture == KID_MOIST_EQ) obs_accumulate_kernel<FT, KID_MOIST_EQ, KID_PRECIP><<<grid, block, 0, s>>>(a, P);
    else obs_accumulate_kernel<FT, KID_MOIST_NONEQ, KID_PRECIP><<<grid, block, 0, s>>>(a, P);
    return cudaGetLastError();
}
template <class FT>
size_t smem_bytes(int moisture, int nz, int C, bool pcp) {
    return moisture == KID_MOIST_EQ ? k1d_tile_smem_bytes<FT, KID_MOIST_EQ, KID_PRECIP>(nz, C, pcp)
                                    : k1d_tile_smem_bytes<FT, KID_MOIST_NONEQ, KID_PRECIP>(nz, C, pcp);
}

#if KID_PRECIP == 3
template <int MOIST>
cudaError_t pair_m(const KArgs<float>& a, const float4* gconst, const float* gG, const CUtensorMap& tmY, const DevParams<float>& P, cudaStream_t s) {
    auto kern = k1d_pair_kernel<MOIST>;
    const size_t smem = k1d_pair_smem_bytes<MOIST>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    dim3 block(PAIR_C, PAIR_TJ), grid((a.nc + PAIR_C - 1) / PAIR_C);
    kern<<<grid, block, smem, s>>>(a, gconst, gG, tmY, P);
    return cudaGetLastError();
}
cudaError_t pair(int moisture, const KArgs<float>& a, const float4* gconst, const float* gG, const CUtensorMap& tmY, const DevParams<float>& P, cudaStream_t s) {
    return moisture == KID_MOIST_EQ ? pair_m<KID_MOIST_EQ>(a, gconst, gG, tmY, P, s) : pair_m<KID_MOIST_NONEQ>(a, gconst, gG, tmY, P, s);
}
cudaError_t gconsts(const KArgs<float>& a, const DevParams<float>& P, float4* out, float* outG, cudaStream_t s) {
    const int ncol = a.prof_per_column ? a.nc : 1;
    dim3 block(128), grid((ncol + 127) / 128, a.nz);
    grid_consts_kernel<ConstPrm<float>><<<grid, block, 0, s>>>(a.T, a.ps_liq, a.ps_mix, out, outG, a.nz, a.nc, a.ncp, a.prof_per_column, P);
    return cudaGetLastError();
}
#ifdef KID_EXP_TIMING
}  // namespace
}  // namespace kid
extern "C" int kid_exp_pair_cycles(unsigned long long* out, int reset) {
    if (cudaMemcpyFromSymbol(out, kid::g_pair_cycles, sizeof(kid::g_pair_cycles)) != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(out + 64, kid::g_pair_work, sizeof(kid::g_pair_work)) != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(out + 128, kid::g_pair_wait, sizeof(kid::g_pair_wait)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[8][8] = {}; cudaMemcpyToSymbol(kid::g_pair_cycles, z, sizeof(z)); cudaMemcpyToSymbol(kid::g_pair_work, z, sizeof(z)); cudaMemcpyToSymbol(kid::g_pair_wait, z, sizeof(z)); }
    return 0;
}
namespace kid {
namespace {
#endif
#define KID_PAIR_ENTRIES pair, gconsts
#else
#define KID_PAIR_ENTRIES nullptr, nullptr
#endif

const LaunchTable table = {k1d<float>, k1d<double>, box<float>, box<double>,
                           consts<float>, consts<double>, obs<float>, obs<double>, smem_bytes<float>, smem_bytes<double>,
                           KID_PAIR_ENTRIES};
}  // namespace

#define KID_CAT2(a, b) a##b
#define KID_CAT(a, b) KID_CAT2(a, b)
const LaunchTable* KID_CAT(launch_table_precip, KID_PRECIP)() { return &table; }

}  // namespace kid
