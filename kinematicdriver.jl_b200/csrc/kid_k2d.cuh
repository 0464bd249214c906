// K2DModel horizontal operators (sm_100a): spectral-element weak divergence per element, direct stiffness
// summation (DSS) at the shared GLL edge nodes, SSPRK33 stage combine.
//
// Replaces (reference): the `-hdiv(ρu/ρ·x)` + `CC.Spaces.weighted_dss!` statements of
// src/K2DModel/tendency.jl:68-70,100-108,152-157,202-207 and the stage axpys of ODE.solve.
// The vertical/pointwise part of the K2D rhs runs in k1d_tile_kernel (K2D mode) and leaves
//     dY[var][level][column]   pointwise sources + vertical advection
//     S[0..1][level][column]   precipitation advection fields (ρq_rai; 1M also ρq_sno)
// Columns are the GLL nodes, c = element*Nq + node; a rank owns whole elements (x-decomposition).
#pragma once
#include <cuda_runtime.h>

#include "kid_physics.cuh"

namespace kid {

constexpr int KID_MAX_NQ = 8;
constexpr int KID_MAX_DSS = 5;

struct K2DField {
    int adv_var;  // state variable whose horizontal flux is taken
    int is_S;     // 0: target is dY plane `plane`; 1: target is S plane `plane`
    int plane;
};

template <class FT>
struct K2DArgs {
    FT* Y;          // [NV][nz][ncp] stage state (in/out)
    FT* Yout;       // k2d_horizontal_kernel: the new stage state goes to a second buffer (other blocks still read Y); host swaps
    FT* U;          // [NV][nz][ncp] u^n of the current step
    FT* dY;         // [NV][nz][ncp]
    FT* S;          // [2][nz][ncp]
    FT* out;        // rhs-only mode: final tendencies [NV][nz][ncp]
    const FT* rho_dry;
    const FT* w1;   // [ncp]
    const FT* xc;   // [ncp]
    const double* cosx;  // [ncp]    cos(2π x_c / W)      horizontal factor of ρu (src/K2DModel/tendency.jl:4-9)
    const double* cosz;  // [nz]     cos(π z_center / H)  vertical factor of ρu
    double ts;           // sin(π t / t1) for t < t1, else 0
    FT* send_left; FT* send_right;              // [nfields][nz]
    const FT* recv_left; const FT* recv_right;  // [nfields][nz]
    int nz, nc, ncp, nv, helem, Nq, prof_per_column;
    int nfields;
    K2DField fields[KID_MAX_DSS];
    int i_tot, i_rai, i_sno, precip;  // for the S accumulation
    int stage, use_recv, rhs_only;
    FT dt, t, t1, z_min, dz;
    double width, height, dxe;
    FT D[KID_MAX_NQ * KID_MAX_NQ];  // D[i*Nq+j] = l_j'(ξ_i)
    FT w[KID_MAX_NQ];               // GLL weights
};

// one thread per (element, level): weak divergence of every horizontally advected field at the element's Nq nodes,
// then pack the rank-edge nodes for the neighbour exchange.  ρu = ½ w1 (W/H) cos(πz/H) cos(2πx/W) ts is evaluated in
// double from host-built factor tables and rounded like the reference's map.  (A thread-per-node variant with warp
// shuffles was measured slower: 45 vs 34 µs — the element-local loop keeps the fluxes in registers.)
template <class FT>
__global__ void k2d_hdiv_kernel(const K2DArgs<FT> A) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (e >= A.helem) return;
    const int Nq = A.Nq;
    const size_t plane = size_t(A.nz) * A.ncp;
    const size_t row = size_t(k) * A.ncp;
    FT u_over_rho[KID_MAX_NQ];
    const double czk = A.cosz[k];
    for (int j = 0; j < Nq; ++j) {
        const int c = e * Nq + j;
        const FT rd = A.rho_dry[A.prof_per_column ? row + c : size_t(k)];
        const FT rho = rd + A.Y[size_t(A.i_tot) * plane + row + c];
        const FT rho_u = FT(0.5 * double(A.w1[c]) * A.width / A.height * czk * A.cosx[c] * A.ts);
        u_over_rho[j] = rho_u / rho;
    }
    const FT scale = FT(2) / FT(A.dxe);
    for (int f = 0; f < A.nfields; ++f) {
        const K2DField fd = A.fields[f];
        FT wflux[KID_MAX_NQ];
        for (int j = 0; j < Nq; ++j) wflux[j] = A.w[j] * (u_over_rho[j] * A.Y[size_t(fd.adv_var) * plane + row + e * Nq + j]);
        FT* tgt = (fd.is_S ? A.S : A.dY) + size_t(fd.plane) * plane + row;
        for (int i = 0; i < Nq; ++i) {
            FT acc = FT(0);
            for (int j = 0; j < Nq; ++j) acc += A.D[j * Nq + i] * wflux[j];
            const FT hd = -acc / A.w[i] * scale;
            tgt[e * Nq + i] += -hd;
        }
        if (e == 0) A.send_left[size_t(f) * A.nz + k] = tgt[0];
        if (e == A.helem - 1) A.send_right[size_t(f) * A.nz + k] = tgt[(A.helem - 1) * Nq + Nq - 1];
    }
}

// one thread per (column, level): DSS the fields, assemble the tendencies, SSPRK33 stage combine
template <class FT>
__global__ void k2d_dss_update_kernel(const K2DArgs<FT> A) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (c >= A.nc) return;
    const int Nq = A.Nq, e = c / Nq, i = c % Nq;
    const size_t plane = size_t(A.nz) * A.ncp;
    const size_t row = size_t(k) * A.ncp;
    constexpr int MAXV = 8;  // largest state (NonEq + 2M)
    // ---- all global loads first: they are independent, so their latencies overlap
    // weighted_dss!: the duplicated node of two neighbouring elements gets the average of both copies (uniform mesh:
    // equal weights).  Periodic in x; across a rank boundary the partner value was received.
    const bool left_edge = (i == 0), right_edge = (i == Nq - 1);
    int partner_col = -1;  // column of the duplicate node inside this rank, or -1: none / received
    bool from_recv = false;
    if (left_edge) {
        if (e > 0) partner_col = (e - 1) * Nq + Nq - 1;
        else if (A.use_recv) from_recv = true;
        else partner_col = (A.helem - 1) * Nq + Nq - 1;
    } else if (right_edge) {
        if (e < A.helem - 1) partner_col = (e + 1) * Nq;
        else if (A.use_recv) from_recv = true;
        else partner_col = 0;
    }
    FT own[KID_MAX_DSS], par[KID_MAX_DSS];
#pragma unroll
    for (int f = 0; f < KID_MAX_DSS; ++f) {
        own[f] = par[f] = FT(0);
        if (f < A.nfields) {
            const K2DField fd = A.fields[f];
            const FT* src = (fd.is_S ? A.S : A.dY) + size_t(fd.plane) * plane + row;
            own[f] = src[c];
            if (partner_col >= 0) par[f] = src[partner_col];
            else if (from_recv) par[f] = (left_edge ? A.recv_left : A.recv_right)[size_t(f) * A.nz + k];
        }
    }
    FT dy[MAXV], yv[MAXV], uv[MAXV];
#pragma unroll
    for (int var = 0; var < MAXV; ++var) {
        dy[var] = yv[var] = uv[var] = FT(0);
        if (var < A.nv) {
            const size_t gi = size_t(var) * plane + row + c;
            dy[var] = A.dY[gi];
            if (!A.rhs_only) {
                yv[var] = A.Y[gi];
                if (A.stage != 0) uv[var] = A.U[gi];
            }
        }
    }
    // ---- DSS
    FT Sval[2] = {FT(0), FT(0)};
#pragma unroll
    for (int f = 0; f < KID_MAX_DSS; ++f) {
        if (f < A.nfields) {
            FT v = own[f];
            if (left_edge && (partner_col >= 0 || from_recv)) v = (par[f] + v) / FT(2);
            else if (right_edge && (partner_col >= 0 || from_recv)) v = (v + par[f]) / FT(2);
            own[f] = v;
            if (A.fields[f].is_S) { if (A.fields[f].plane == 0) Sval[0] = v; else Sval[1] = v; }
        }
    }
    // ---- assemble the tendencies in the reference's order, SSPRK33 stage combine
#pragma unroll
    for (int var = 0; var < MAXV; ++var) {
        if (var >= A.nv) break;
        const size_t gi = size_t(var) * plane + row + c;
        FT tend = dy[var];
#pragma unroll
        for (int f = 0; f < KID_MAX_DSS; ++f)
            if (f < A.nfields && !A.fields[f].is_S && A.fields[f].plane == var) tend = own[f];
        if (A.precip == KID_PRECIP_1M) {
            if (var == A.i_tot) tend += Sval[0] + Sval[1];
            if (var == A.i_rai) tend += Sval[0];
            if (var == A.i_sno) tend += Sval[1];
        } else if (A.precip == KID_PRECIP_2M) {
            if (var == A.i_tot) tend += Sval[0];
            if (var == A.i_rai) tend += Sval[0];
        }
        if (A.rhs_only) { A.out[gi] = tend; continue; }
        const FT y = yv[var];
        const FT u = (A.stage == 0) ? y : uv[var];
        if (A.stage == 0) A.U[gi] = y;
        FT yn;
        if (A.stage == 0) yn = y + A.dt * tend;
        else if (A.stage == 1) yn = (FT(3) * u + y + A.dt * tend) / FT(4);
        else yn = (u + FT(2) * y + FT(2) * A.dt * tend) / FT(3);
        A.Y[gi] = yn;
    }
}

// ---------------------------------------------------------------- one launch for the horizontal part of a stage
// horizontally advected / DSS'd fields per style, in the order of kid_capi.cu::k2d_setup (src/K2DModel/tendency.jl:57-213):
// {state variable whose flux is taken, target is an S plane, target plane}
template <int MOIST, int PRECIP>
__host__ __device__ constexpr int k2d_field_count() {
    return 1 + (MOIST == KID_MOIST_NONEQ ? 2 : 0) + (PRECIP >= KID_PRECIP_1M ? 2 : 0);
}
template <int MOIST, int PRECIP>
__host__ __device__ constexpr K2DField k2d_field(int f) {
    using L = Lay<MOIST, PRECIP>;
    if (f == 0) return {L::tot, 0, L::tot};
    int b = 1;
    if (MOIST == KID_MOIST_NONEQ) {
        if (f == 1) return {L::liq, 0, L::liq};
        if (f == 2) return {L::ice, 0, L::ice};
        b = 3;
    }
    if (PRECIP == KID_PRECIP_1M) return f == b ? K2DField{L::rai, 1, 0} : K2DField{L::sno, 1, 1};
    return f == b ? K2DField{L::Nr, 0, L::Nr} : K2DField{L::rai, 1, 0};  // 2M
}

// u/ρ at (column c, level k) exactly as k2d_hdiv_kernel forms it
template <class FT>
KID_DEV FT k2d_u_over_rho(const K2DArgs<FT>& A, int c, int k, FT y_tot) {
    const FT rd = A.rho_dry[A.prof_per_column ? size_t(k) * A.ncp + c : size_t(k)];
    const FT rho = rd + y_tot;
    const FT rho_u = FT(0.5 * double(A.w1[c]) * A.width / A.height * A.cosz[k] * A.cosx[c] * A.ts);
    return rho_u / rho;
}
// The not yet DSS'd tendency of field f at an arbitrary column (slow path: the duplicate of a node that lies in another
// thread block, and the rank-edge pack): vertical part from dY / S plus the horizontal weak divergence of the column's element.
template <class FT, int MOIST, int PRECIP>
KID_DEV FT k2d_field_tendency_at(const K2DArgs<FT>& A, K2DField fd, int c, int k) {
    using L = Lay<MOIST, PRECIP>;
    const int Nq = A.Nq, e = c / Nq, i = c % Nq;
    const size_t plane = size_t(A.nz) * A.ncp, row = size_t(k) * A.ncp;
    FT acc = FT(0);
    for (int q = 0; q < Nq; ++q) {
        const int cq = e * Nq + q;
        const FT uor = k2d_u_over_rho(A, cq, k, A.Y[size_t(L::tot) * plane + row + cq]);
        acc += A.D[q * Nq + i] * (A.w[q] * (uor * A.Y[size_t(fd.adv_var) * plane + row + cq]));
    }
    const FT hd = -acc / A.w[i] * (FT(2) / FT(A.dxe));
    FT t = ((fd.is_S ? A.S : A.dY) + size_t(fd.plane) * plane + row)[c];
    t += -hd;
    return t;
}
// rank-edge pack of a decomposed domain: the tendencies of the first and last column, for the neighbour ranks' DSS
template <class FT, int MOIST, int PRECIP>
__global__ void k2d_edge_pack_kernel(const K2DArgs<FT> A) {
    // one thread per (level, side, field): the pack sits on the critical path of every stage of a decomposed run
    constexpr int NF = k2d_field_count<MOIST, PRECIP>();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = t % A.nz, side = (t / A.nz) % 2, f = t / (2 * A.nz);
    if (f >= NF) return;
    K2DField fd = k2d_field<MOIST, PRECIP>(0);
#pragma unroll
    for (int g = 1; g < NF; ++g)
        if (g == f) fd = k2d_field<MOIST, PRECIP>(g);
    const FT v = k2d_field_tendency_at<FT, MOIST, PRECIP>(A, fd, side ? A.nc - 1 : 0, k);
    (side ? A.send_right : A.send_left)[size_t(f) * A.nz + k] = v;
}

// One thread per (column, level).  A block is 128 neighbouring columns = whole elements (the Nq nodes of an element are
// neighbouring lanes of one warp, Nq in {2, 4, 8}); its first and last ELEMENT are halo: every thread forms the not yet DSS'd
// tendencies of its column — horizontal weak divergence with the D-matrix row applied across the lanes (coalesced loads; the
// arithmetic of k2d_hdiv_kernel) — and the 128 - 2 Nq interior columns then take their duplicate node's value from shared
// memory (weighted_dss!; periodic wrap, or the received rank edge), assemble in the reference's order and do the SSPRK33 stage
// combine (the arithmetic of k2d_dss_point).  Replaces k2d_hdiv_kernel + k2d_dss_update_kernel for kid_step.
template <class FT, int MOIST, int PRECIP>
__global__ void __launch_bounds__(128) k2d_horizontal_kernel(const K2DArgs<FT> A) {
    using L = Lay<MOIST, PRECIP>;
    constexpr int NV = L::NV, NF = k2d_field_count<MOIST, PRECIP>();
    __shared__ FT tvs[NF][128];
    __shared__ FT Ds[KID_MAX_NQ * KID_MAX_NQ], ws[KID_MAX_NQ];
    const int tid = threadIdx.x, k = blockIdx.y, Nq = A.Nq;
    const int c_raw = int(blockIdx.x) * (128 - 2 * Nq) - Nq + tid;   // first / last Nq threads: halo elements
    // periodic domain: the halo wraps around; decomposed domain: beyond the rank's columns there is nothing (the rank edges
    // take the received values)
    const bool outside = c_raw < 0 || c_raw >= A.nc;
    const int c = outside ? (c_raw < 0 ? c_raw + A.nc : c_raw - A.nc) : c_raw;
    const bool valid = !(outside && A.use_recv) && c >= 0 && c < A.nc;
    const bool interior = valid && !outside && tid >= Nq && tid < 128 - Nq;
    const int ni = tid % Nq, lane = tid & 31, lane0 = lane - ni;
    const size_t plane = size_t(A.nz) * A.ncp, row = size_t(k) * A.ncp;
    if (tid < Nq * Nq) Ds[tid] = A.D[tid];
    if (tid < Nq) ws[tid] = A.w[tid];
    // ---- all global loads first
    FT y[NV], tend[NV], u[NV], Sv[2] = {FT(0), FT(0)};
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        y[v] = tend[v] = u[v] = FT(0);
        if (valid) {
            const size_t gi = size_t(v) * plane + row + c;
            y[v] = A.Y[gi];
            tend[v] = A.dY[gi];
            if (A.stage != 0 && interior) u[v] = A.U[gi];
        }
    }
    if constexpr (L::has_pr) {
        if (valid) {
            Sv[0] = A.S[row + c];
            if constexpr (PRECIP == KID_PRECIP_1M) Sv[1] = A.S[plane + row + c];
        }
    }
    const FT uor = valid ? k2d_u_over_rho(A, c, k, y[L::tot]) : FT(0);
    __syncthreads();
    // ---- horizontal weak divergence: -(1/w_i) Σ_q D_qi w_q f_q · 2/Δx_e, f = (ρu/ρ) x
    const FT w_own = ws[ni], scale = FT(2) / FT(A.dxe);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const K2DField fd = k2d_field<MOIST, PRECIP>(f);
        const FT wflux = w_own * (uor * y[fd.adv_var]);
        FT acc = FT(0);
        for (int q = 0; q < Nq; ++q) acc += Ds[q * Nq + ni] * __shfl_sync(0xffffffffu, wflux, lane0 + q);
        const FT hd = -acc / w_own * scale;
        FT& val = fd.is_S ? Sv[fd.plane] : tend[fd.plane];
        val += -hd;
        tvs[f][tid] = val;
    }
    __syncthreads();
    if (!interior) return;
    // ---- weighted_dss!: mean of the two copies of an element-edge node (the neighbouring thread)
    const bool left_edge = (ni == 0), right_edge = (ni == Nq - 1);
    if (left_edge || right_edge) {
        const bool from_recv = A.use_recv && ((left_edge && c == 0) || (right_edge && c == A.nc - 1));
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const K2DField fd = k2d_field<MOIST, PRECIP>(f);
            const FT pv = from_recv ? (left_edge ? A.recv_left : A.recv_right)[size_t(f) * A.nz + k]
                                    : tvs[f][left_edge ? tid - 1 : tid + 1];
            FT& val = fd.is_S ? Sv[fd.plane] : tend[fd.plane];
            val = left_edge ? (pv + val) / FT(2) : (val + pv) / FT(2);
        }
    }
    // ---- assembly in the reference's order, SSPRK33 stage combine
    if constexpr (PRECIP == KID_PRECIP_1M) {
        tend[L::tot] += Sv[0] + Sv[1]; tend[L::rai] += Sv[0]; tend[L::sno] += Sv[1];
    } else if constexpr (PRECIP == KID_PRECIP_2M) {
        tend[L::tot] += Sv[0]; tend[L::rai] += Sv[0];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const size_t gi = size_t(v) * plane + row + c;
        const FT yv = y[v];
        const FT uu = (A.stage == 0) ? yv : u[v];
        if (A.stage == 0) A.U[gi] = yv;
        FT yn;
        if (A.stage == 0) yn = yv + A.dt * tend[v];
        else if (A.stage == 1) yn = (FT(3) * uu + yv + A.dt * tend[v]) / FT(4);
        else yn = (uu + FT(2) * yv + FT(2) * A.dt * tend[v]) / FT(3);
        A.Yout[gi] = yn;
    }
}

// aux.microph_variables.{q_rai, N_liq, N_rai} as initialise_aux builds them from the initial profile
template <class FT>
__global__ void k2d_init_prev_kernel(const FT* Y, const FT* rho_dry, FT* prev, int nz, int nc, int ncp, int prof_per_column,
                                     int i_tot, int i_rai, int i_Nl, int i_Nr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (c >= nc) return;
    const size_t plane = size_t(nz) * ncp, pi = size_t(k) * ncp + c;
    const FT rho = rho_dry[prof_per_column ? pi : size_t(k)] + Y[size_t(i_tot) * plane + pi];
    prev[pi] = i_rai >= 0 ? Y[size_t(i_rai) * plane + pi] / rho : FT(0);
    prev[plane + pi] = i_Nl >= 0 ? Y[size_t(i_Nl) * plane + pi] : FT(0);
    prev[2 * plane + pi] = i_Nr >= 0 ? Y[size_t(i_Nr) * plane + pi] : FT(0);
}

// Gauss-Lobatto-Legendre nodes, weights and differentiation matrix D[i][j] = l_j'(ξ_i) (host, double)
inline void gll_rule(int Nq, double* x, double* w, double* D) {
    const int N = Nq - 1;
    auto legendre = [&](double xi, double& PN, double& PNm1) {
        double p0 = 1, p1 = xi;
        for (int n = 2; n <= N; ++n) { double p2 = ((2 * n - 1) * xi * p1 - (n - 1) * p0) / n; p0 = p1; p1 = p2; }
        PN = (N == 0) ? 1 : p1; PNm1 = (N == 0) ? 0 : p0;
    };
    for (int i = 0; i < Nq; ++i) {
        double xi = -std::cos(M_PI * i / N);
        if (i > 0 && i < N) {
            for (int it = 0; it < 100; ++it) {
                double PN, PNm1;
                legendre(xi, PN, PNm1);
                double dx = (N * (PNm1 - xi * PN)) / (-double(N) * (N + 1) * PN);
                xi -= dx;
                if (std::fabs(dx) < 1e-16) break;
            }
        } else xi = (i == 0) ? -1.0 : 1.0;
        double PN, PNm1;
        legendre(xi, PN, PNm1);
        x[i] = xi;
        w[i] = 2.0 / (N * (N + 1) * PN * PN);
    }
    double bw[KID_MAX_NQ];
    for (int j = 0; j < Nq; ++j) { bw[j] = 1.0; for (int m = 0; m < Nq; ++m) if (m != j) bw[j] /= (x[j] - x[m]); }
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

}  // namespace kid
