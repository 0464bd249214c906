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
    static __device__ __forceinline__ const DevParams<float>& get() { return c_prm_f; }
};
template <> struct ConstPrm<double> {
    static __device__ __forceinline__ const DevParams<double>& get() { return c_prm_d; }
};

cudaError_t upload_f(const DevParams<float>& p, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_prm_f, &p, sizeof(p), 0, cudaMemcpyHostToDevice, s);
}
cudaError_t upload_d(const DevParams<double>& p, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_prm_d, &p, sizeof(p), 0, cudaMemcpyHostToDevice, s);
}

template <class FT, int MOIST, bool DIAG>
cudaError_t launch_tile(const KArgs<FT>& a, cudaStream_t s) {
    auto kern = k1d_tile_kernel<FT, MOIST, KID_PRECIP, DIAG, ConstPrm<FT>>;
    size_t smem = k1d_tile_smem_bytes<FT, MOIST, KID_PRECIP>(a.nz, a.C);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    dim3 block(a.C, a.NW), grid((a.nc + a.C - 1) / a.C);
    kern<<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}
template <class FT>
cudaError_t k1d(int moisture, bool diag, const KArgs<FT>& a, cudaStream_t s) {
    if (moisture == KID_MOIST_EQ) return diag ? launch_tile<FT, KID_MOIST_EQ, true>(a, s) : launch_tile<FT, KID_MOIST_EQ, false>(a, s);
    return diag ? launch_tile<FT, KID_MOIST_NONEQ, true>(a, s) : launch_tile<FT, KID_MOIST_NONEQ, false>(a, s);
}
template <class FT>
cudaError_t box(int moisture, const KArgs<FT>& a, cudaStream_t s) {
    dim3 block(128), grid((a.nc + 127) / 128);
    if (moisture == KID_MOIST_EQ) box_step_kernel<FT, KID_MOIST_EQ, KID_PRECIP, ConstPrm<FT>><<<grid, block, 0, s>>>(a);
    else box_step_kernel<FT, KID_MOIST_NONEQ, KID_PRECIP, ConstPrm<FT>><<<grid, block, 0, s>>>(a);
    return cudaGetLastError();
}
template <class FT>
size_t smem_bytes(int moisture, int nz, int C) {
    return moisture == KID_MOIST_EQ ? k1d_tile_smem_bytes<FT, KID_MOIST_EQ, KID_PRECIP>(nz, C)
                                    : k1d_tile_smem_bytes<FT, KID_MOIST_NONEQ, KID_PRECIP>(nz, C);
}

const LaunchTable table = {upload_f, upload_d, k1d<float>, k1d<double>, box<float>, box<double>,
                           smem_bytes<float>, smem_bytes<double>};
}  // namespace

#define KID_CAT2(a, b) a##b
#define KID_CAT(a, b) KID_CAT2(a, b)
const LaunchTable* KID_CAT(launch_table_precip, KID_PRECIP)() { return &table; }

}  // namespace kid
