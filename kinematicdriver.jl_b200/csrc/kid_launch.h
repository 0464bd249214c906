// Host-side launch interface between the C ABI (kid_capi.cu) and the per-precipitation-style
// kernel translation units (kid_inst.cu compiled once per KID_PRECIP value, so the four
// heavy instantiation sets build in parallel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "kid_diag.cuh"
#include "kid_kernels.cuh"

namespace kid {

// tile of k1d_pair_kernel (kid_pair_kernel.cuh): 128 levels x 16 columns, 16 level-pair rows (256 threads, 128 registers):
// two tiles are resident per SM, so the barrier waits of one overlap with the work of the other
#ifndef KID_PAIR_C
#define KID_PAIR_C 16
#endif
#ifndef KID_PAIR_TJ
#define KID_PAIR_TJ 16
#endif
#ifndef KID_PAIR_CTAS
#define KID_PAIR_CTAS 2
#endif
constexpr int PAIR_NZ = 128, PAIR_C = KID_PAIR_C, PAIR_TJ = KID_PAIR_TJ, PAIR_CTAS_PER_SM = KID_PAIR_CTAS;

struct LaunchTable {
    // K1D / col_sed / (diag of) Box: tile kernel.  diag = false: step; true: rhs / aux mode.
    // The shared parameter set is a kernel argument; per-column sets (args.prm_sets != null) are staged in shared memory.
    cudaError_t (*k1d_f)(int moisture, bool diag, const KArgs<float>&, const DevParams<float>&, cudaStream_t);
    cudaError_t (*k1d_d)(int moisture, bool diag, const KArgs<double>&, const DevParams<double>&, cudaStream_t);
    cudaError_t (*box_f)(int moisture, const KArgs<float>&, const DevParams<float>&, cudaStream_t);
    cudaError_t (*box_d)(int moisture, const KArgs<double>&, const DevParams<double>&, cudaStream_t);
    cudaError_t (*consts_f)(const KArgs<float>&, const DevParams<float>&, float* ps_liq, float* ps_mix, cudaStream_t);
    cudaError_t (*consts_d)(const KArgs<double>&, const DevParams<double>&, double* ps_liq, double* ps_mix, cudaStream_t);
    cudaError_t (*obs_f)(int moisture, const ObsArgs<float>&, const DevParams<float>&, cudaStream_t);
    cudaError_t (*obs_d)(int moisture, const ObsArgs<double>&, const DevParams<double>&, cudaStream_t);
    size_t (*k1d_smem_f)(int moisture, int nz, int C, bool per_column_params);
    size_t (*k1d_smem_d)(int moisture, int nz, int C, bool per_column_params);
    // Precipitation2M, Float32, 128 levels, shared parameter set: the packed two-levels-per-thread step kernel
    // (kid_pair_kernel.cuh) and its per-grid-point constants; null for the other precipitation styles
    cudaError_t (*pair_f)(int moisture, const KArgs<float>&, const float4* gconst, const float* gG, const CUtensorMap& state_map, const DevParams<float>&, cudaStream_t);
    cudaError_t (*gconsts_f)(const KArgs<float>&, const DevParams<float>&, float4* out, float* outG, cudaStream_t);
};

const LaunchTable* launch_table_precip0();
const LaunchTable* launch_table_precip1();
const LaunchTable* launch_table_precip2();
const LaunchTable* launch_table_precip3();

}  // namespace kid
