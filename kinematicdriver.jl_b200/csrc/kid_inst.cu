// Kernel instantiations for ONE precipitation style (compile with -DKID_PRECIP=0..3).
#ifndef KID_PRECIP
#error "compile with -DKID_PRECIP=<enum kid_precip value>"
#endif
#include "kid_launch.h"

namespace kid {
namespace {

__constant__ DevParams<float> c_prm_f;
__constant__ DevParams<double> c_prm_d;

template <class FT> struct ConstPrm;
template <> struct ConstPrm<float> {
    static constexpr bool per_column = false;
    static __device__ __forceinline__ const DevParams<float>& get() { return c_prm_f; }
};
template <> struct ConstPrm<double> {
    static constexpr bool per_column = false;
    static __device__ __forceinline__ const DevParams<double>& get() { return c_prm_d; }
};
// per-column parameter sets (calibration ensembles): staged in shared memory by the kernels
template <class FT> struct ColumnPrm {
    static constexpr bool per_column = true;
    static __device__ __forceinline__ const DevParams<FT>& get();  // never called
};

cudaError_t upload_f(const DevParams<float>& p, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_prm_f, &p, sizeof(p), 0, cudaMemcpyHostToDevice, s);
}
cudaError_t upload_d(const DevParams<double>& p, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_prm_d, &p, sizeof(p), 0, cudaMemcpyHostToDevice, s);
}

template <class FT, int MOIST, bool DIAG, class PRM>
cudaError_t launch_tile_p(const KArgs<FT>& a, cudaStream_t s) {
    auto kern = k1d_tile_kernel<FT, MOIST, KID_PRECIP, DIAG, PRM>;
    size_t smem = k1d_tile_smem_bytes<FT, MOIST, KID_PRECIP>(a.nz, a.C, PRM::per_column);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    dim3 block(a.C, a.NW), grid((a.nc + a.C - 1) / a.C);
    kern<<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}
template <class FT, int MOIST, bool DIAG>
cudaError_t launch_tile(const KArgs<FT>& a, cudaStream_t s) {
    return a.prm_sets ? launch_tile_p<FT, MOIST, DIAG, ColumnPrm<FT>>(a, s) : launch_tile_p<FT, MOIST, DIAG, ConstPrm<FT>>(a, s);
}
template <class FT>
cudaError_t k1d(int moisture, bool diag, const KArgs<FT>& a, cudaStream_t s) {
    if (moisture == KID_MOIST_EQ) return diag ? launch_tile<FT, KID_MOIST_EQ, true>(a, s) : launch_tile<FT, KID_MOIST_EQ, false>(a, s);
    return diag ? launch_tile<FT, KID_MOIST_NONEQ, true>(a, s) : launch_tile<FT, KID_MOIST_NONEQ, false>(a, s);
}
template <class FT, int MOIST, class PRM>
cudaError_t box_p(const KArgs<FT>& a, cudaStream_t s) {
    const int bs = PRM::per_column ? 64 : 128;
    const size_t smem = PRM::per_column ? size_t(bs) * sizeof(DevParams<FT>) : 0;
    auto kern = box_step_kernel<FT, MOIST, KID_PRECIP, PRM>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
    }
    dim3 block(bs), grid((a.nc + bs - 1) / bs);
    kern<<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}
template <class FT>
cudaError_t box(int moisture, const KArgs<FT>& a, cudaStream_t s) {
    if (moisture == KID_MOIST_EQ)
        return a.prm_sets ? box_p<FT, KID_MOIST_EQ, ColumnPrm<FT>>(a, s) : box_p<FT, KID_MOIST_EQ, ConstPrm<FT>>(a, s);
    return a.prm_sets ? box_p<FT, KID_MOIST_NONEQ, ColumnPrm<FT>>(a, s) : box_p<FT, KID_MOIST_NONEQ, ConstPrm<FT>>(a, s);
}
template <class FT>
cudaError_t consts(const KArgs<FT>& a, FT* ps_liq, FT* ps_mix, cudaStream_t s) {
    const int ncol = a.prof_per_column ? a.nc : 1;
    dim3 block(128), grid((ncol + 127) / 128, a.nz);
    if (a.prm_sets && a.prof_per_column)
        profile_consts_kernel<FT, ColumnPrm<FT>><<<grid, block, 0, s>>>(a.T, ps_liq, ps_mix, a.nz, a.nc, a.ncp, 1, a.prm_sets, a.col_to_set);
    else
        profile_consts_kernel<FT, ConstPrm<FT>><<<grid, block, 0, s>>>(a.T, ps_liq, ps_mix, a.nz, a.nc, a.ncp, a.prof_per_column, nullptr, nullptr);
    return cudaGetLastError();
}
template <class FT>
size_t smem_bytes(int moisture, int nz, int C, bool pcp) {
    return moisture == KID_MOIST_EQ ? k1d_tile_smem_bytes<FT, KID_MOIST_EQ, KID_PRECIP>(nz, C, pcp)
                                    : k1d_tile_smem_bytes<FT, KID_MOIST_NONEQ, KID_PRECIP>(nz, C, pcp);
}

const LaunchTable table = {upload_f, upload_d, k1d<float>, k1d<double>, box<float>, box<double>,
                           consts<float>, consts<double>, smem_bytes<float>, smem_bytes<double>};
}  // namespace

#define KID_CAT2(a, b) a##b
#define KID_CAT(a, b) KID_CAT2(a, b)
const LaunchTable* KID_CAT(launch_table_precip, KID_PRECIP)() { return &table; }

}  // namespace kid
