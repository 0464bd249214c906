// Device-side calibration observables (sm_100a): the variables of `get_variable_data_from_ODE`
// (src/Common/helper_functions.jl:98-191) and the space/time filtering of `ODEsolution2Gvector`
// (calibration/KiDUtils.jl:329-376), accumulated into a per-column G vector on the device so that a calibration
// ensemble never downloads its full state, plus the device-side maximum behind kid_reduce_max
// (`ql_max`, src/Common/netcdf_io.jl:166-169).
#pragma once
#include <cuda_runtime.h>

#include "kid_physics.cuh"

namespace kid {

constexpr int KID_MAX_OBS = 17;

template <class FT>
struct ObsArgs {
    const FT* Y;        // [NV][nz][ncp]
    const FT* rho_dry;  // [nz][ncp] or [nz]
    double* G;          // [nc][nt_filtered][n_out]
    const DevParams<FT>* prm_sets;  // per-column parameter sets, or null: the shared set
    const int* col_to_set;
    int nz, nc, ncp, prof_per_column;
    int nvars, n_out, nt_filtered, time_cell;
    int var_id[KID_MAX_OBS];      // enum kid_obs_var
    int nz_per_cell[KID_MAX_OBS]; // levels averaged into one filtered cell
    int out_offset[KID_MAX_OBS + 1];  // prefix sum of nz / nz_per_cell over the variables
    double weight;                // 1 / nt_per_filtered_cell
    double z_span;                // z_max - z_min (Z_top: the 500 m averaging depth in levels)
    DevSwitches sw;
};

// one variable of get_variable_data_from_ODE at one grid point.  The specific humidities are ρq/ρ WITHOUT the
// max(0, ·) of the rhs (helper_functions.jl:103-127 divides the raw state).
template <class FT, int MOIST, int PRECIP, class PRM>
KID_DEV FT obs_value(const PRM& P, const DevSwitches& sw, int var, const FT* y, FT rho_dry) {
    using L = Lay<MOIST, PRECIP>;
    const FT rho = rho_dry + y[L::tot];
    const FT qt = y[L::tot] / rho;
    FT rq_liq = FT(0), rq_rai = FT(0);
    if constexpr (L::neq) rq_liq = y[L::liq];
    if constexpr (L::has_pr) rq_rai = y[L::rai];
    switch (var) {
        case KID_OBS_QT: return qt;
        case KID_OBS_QL: return rq_liq / rho;
        case KID_OBS_QR: return rq_rai / rho;
        case KID_OBS_QV: return (y[L::tot] - rq_liq) / rho;
        case KID_OBS_QLR: return (rq_liq + rq_rai) / rho;
        case KID_OBS_RT: return qt / (FT(1) - qt);
        case KID_OBS_RL: return rq_liq / rho / (FT(1) - qt);
        case KID_OBS_RR: return rq_rai / rho / (FT(1) - qt);
        case KID_OBS_RV: return (y[L::tot] - rq_liq) / rho / (FT(1) - qt);
        case KID_OBS_RLR: return (rq_liq + rq_rai) / rho / (FT(1) - qt);
        case KID_OBS_RHO: return rho;
        default: break;
    }
    if constexpr (L::is2m) {
        if (var == KID_OBS_NL) return y[L::Nl];
        if (var == KID_OBS_NR) return y[L::Nr];
        if (var == KID_OBS_NA) return y[L::Na];
    }
    if constexpr (L::has_pr) {
        if (var == KID_OBS_RAIN_VT || var == KID_OBS_RAINRATE) {
            // CM1.terminal_velocity(rain, vel, ρ, qr) / CM2.rain_terminal_velocity(sb2006, vel, qr, ρ, Nr)[2]
            Point<FT> a{};
            a.rho = rho;
            a.q_rai = rq_rai / rho;
            a.q_sno = FT(0);
            if constexpr (L::is2m) a.N_rai = y[L::Nr];
            FT v1, v2;
            terminal_velocities<FT, PRECIP>(P, sw, a, v1, v2);
            return var == KID_OBS_RAIN_VT ? v1 : a.q_rai * rho * v1 * FT(3600);
        }
    }
    return FT(0);
}

// "Z_top" of get_variable_data_from_ODE (src/Common/helper_functions.jl:164-185): cloud top = highest level with
// q_liq + q_rai >= 1e-3 (level 1 if none), q_rai and ρ averaged over the `steps` levels below it — steps =
// round(500 m / dz) with dz = (z_max - z_min) / (nz + 1), the reference divides by the number of FACES — then
// CMD.radar_reflectivity_1M(rain, q_rai, ρ) = 10 log10(720 n0 λ⁻⁷ / 1e-18 m³) of the Marshall-Palmer distribution, with the
// reference's replacements -Inf -> -300, Inf -> 300, NaN -> 0.
template <class FT, int MOIST, int PRECIP, class PRM>
KID_DEV double obs_z_top(const PRM& P, const ObsArgs<FT>& A, int c) {
    using L = Lay<MOIST, PRECIP>;
    if constexpr (L::neq && PRECIP == KID_PRECIP_1M) {
        const size_t plane = size_t(A.nz) * A.ncp;
        auto rho_at = [&](int k) { return A.rho_dry[A.prof_per_column ? size_t(k) * A.ncp + c : size_t(k)] + A.Y[L::tot * plane + size_t(k) * A.ncp + c]; };
        int z_top = A.nz - 1;  // 0-based
        while (z_top > 0) {
            const FT rho = rho_at(z_top);
            const FT qlr = A.Y[L::liq * plane + size_t(z_top) * A.ncp + c] / rho + A.Y[L::rai * plane + size_t(z_top) * A.ncp + c] / rho;
            if (!(qlr < FT(1e-3))) break;
            --z_top;
        }
        const double dz = A.z_span / double(A.nz + 1);
        const int steps = int(rint(500.0 / dz));
        const int lo = max(z_top - steps + 1, 0);
        FT qr = FT(0), rho_m = FT(0);
        for (int k = lo; k <= z_top; ++k) {
            const FT rho = rho_at(k);
            qr += A.Y[L::rai * plane + size_t(k) * A.ncp + c] / rho;
            rho_m += rho;
        }
        qr /= FT(z_top - lo + 1);
        rho_m /= FT(z_top - lo + 1);
        CM1<FT, PRM> cm1(P);
        const double li = double(cm1.lambda_inverse(P.rain_n0, P.rain_r0_pow, P.rain_m0, P.rain_li_exp, P.rain_chim, P.rain_gam_m, qr, rho_m));
        double Z = 10.0 * (log10(720.0 * double(P.rain_n0)) + 7.0 * log10(li) + 18.0);
        if (Z != Z) Z = 0.0;
        else if (Z == -INFINITY) Z = -300.0;
        else if (Z == INFINITY) Z = 300.0;
        return Z;
    }
    return 0.0;
}

// one thread per (column, filtered output cell): mean over the cell's levels, added to the time cell with weight
// 1 / nt_per_filtered_cell (the Float64 accumulation of ODEsolution2Gvector)
template <class FT, int MOIST, int PRECIP>
__global__ void obs_accumulate_kernel(const ObsArgs<FT> A, const __grid_constant__ DevParams<FT> Pshared) {
    using L = Lay<MOIST, PRECIP>;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int o = blockIdx.y;
    if (c >= A.nc) return;
    int j = 0;
    while (j + 1 < A.nvars && o >= A.out_offset[j + 1]) ++j;
    const int cell = o - A.out_offset[j], per = A.nz_per_cell[j];
    const DevParams<FT>& P = A.prm_sets ? A.prm_sets[A.col_to_set[c]] : Pshared;
    const size_t plane = size_t(A.nz) * A.ncp;
    if (A.var_id[j] == KID_OBS_Z_TOP) {  // one value per column
        A.G[(size_t(c) * A.nt_filtered + A.time_cell) * A.n_out + o] += obs_z_top<FT, MOIST, PRECIP>(P, A, c) * A.weight;
        return;
    }
    double sum = 0.0;
    for (int l = 0; l < per; ++l) {
        const int k = cell * per + l;
        FT y[L::NV];
#pragma unroll
        for (int v = 0; v < L::NV; ++v) y[v] = A.Y[v * plane + size_t(k) * A.ncp + c];
        const FT rd = A.rho_dry[A.prof_per_column ? size_t(k) * A.ncp + c : size_t(k)];
        sum += double(obs_value<FT, MOIST, PRECIP>(P, A.sw, A.var_id[j], y, rd));
    }
    double* g = A.G + (size_t(c) * A.nt_filtered + A.time_cell) * A.n_out + o;
    *g += sum / double(per) * A.weight;
}

// maximum of a [nz][ncp] device field over the nc real columns: block reduction + one atomic per block on the
// order-preserving integer image of the double value
__device__ inline unsigned long long ordered_bits(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
inline double from_ordered_bits(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double x;
    memcpy(&x, &u, 8);
    return x;
}
template <class FT>
__global__ void field_max_kernel(const FT* __restrict__ f, int nz, int nc, int ncp, unsigned long long* out) {
    double m = -1.7976931348623157e308;
    const size_t n = size_t(nz) * ncp;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        if (int(i % ncp) < nc) m = fmax(m, double(f[i]));  // fmax drops NaN, like the host loop it replaces
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    __shared__ double wm[32];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : -1.7976931348623157e308;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
        if (threadIdx.x == 0) atomicMax(out, ordered_bits(m));
    }
}

}  // namespace kid
