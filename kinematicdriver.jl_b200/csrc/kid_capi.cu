// libkid_cuda.so — implementation of the C ABI declared in include/kid_cuda.h.
// Host-side plumbing only (device memory, layout conversion, launch geometry); all
// arithmetic of the hot path lives in kid_kernels.cuh / kid_physics.cuh.  There is no CPU
// fallback: every entry point that computes fails when no CUDA device is usable.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kid_k2d.cuh"
#include "kid_launch.h"
#include "kid_init.cuh"

using namespace kid;

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }
#define KID_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e__) + " (libkid_cuda has no CPU fallback)"); \
    } while (0)

// NCCL is bound at run time (dlopen): libkid_cuda.so has no link-time dependency on it, and inside a process that already
// loaded a NCCL (PyTorch's bundled libnccl.so.2) the same library instance is reused.  Only the point-to-point subset the
// K2D edge exchange needs.  ncclUniqueId is 128 bytes, ncclDataType_t: ncclFloat32 = 7, ncclFloat64 = 8.
struct NcclUniqueIdBytes { char internal[128]; };  // layout of ncclUniqueId (passed BY VALUE to ncclCommInitRank)
struct NcclApi {
    typedef struct ncclComm* comm_t;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(comm_t*, int, NcclUniqueIdBytes, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
const NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return a;
        a.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(lib, "ncclGetUniqueId"));
        a.CommDestroy = reinterpret_cast<int (*)(NcclApi::comm_t)>(dlsym(lib, "ncclCommDestroy"));
        a.GroupStart = reinterpret_cast<int (*)()>(dlsym(lib, "ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<int (*)()>(dlsym(lib, "ncclGroupEnd"));
        a.Send = reinterpret_cast<int (*)(const void*, size_t, int, int, NcclApi::comm_t, cudaStream_t)>(dlsym(lib, "ncclSend"));
        a.Recv = reinterpret_cast<int (*)(void*, size_t, int, int, NcclApi::comm_t, cudaStream_t)>(dlsym(lib, "ncclRecv"));
        a.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(lib, "ncclGetErrorString"));
        a.CommInitRank = reinterpret_cast<int (*)(NcclApi::comm_t*, int, NcclUniqueIdBytes, int)>(dlsym(lib, "ncclCommInitRank"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv;
        return a;
    }();
    return api;
}

// Every entry point runs on the handle's device and leaves the caller's current device untouched: several handles on
// different GPUs may be driven from one thread (kid_k2d_exchange_local, k2d.LocalRing).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int nvars_of(int moisture, int precip) {
    if (moisture != KID_MOIST_EQ && moisture != KID_MOIST_NONEQ) return -1;
    if (precip < KID_PRECIP_NONE || precip > KID_PRECIP_2M) return -1;
    int n = 1 + (moisture == KID_MOIST_NONEQ ? 2 : 0);
    if (precip == KID_PRECIP_1M || precip == KID_PRECIP_2M) n += 2;
    if (precip == KID_PRECIP_2M) n += 3;
    return n;
}
int na_index(int moisture, int precip) {  // index of N_aer in the state, -1 if absent
    return precip == KID_PRECIP_2M ? nvars_of(moisture, precip) - 1 : -1;
}
const LaunchTable* table_of(int precip) {
    switch (precip) {
        case KID_PRECIP_NONE: return launch_table_precip0();
        case KID_PRECIP_0M: return launch_table_precip1();
        case KID_PRECIP_1M: return launch_table_precip2();
        case KID_PRECIP_2M: return launch_table_precip3();
    }
    return nullptr;
}
}  // namespace

struct kid_handle {
    kid_config_t cfg{};
    int nv = 0, ncp = 0;
    bool f64 = false;
    size_t fsz = 4;
    cudaStream_t stream = nullptr;
    void *Y = nullptr, *rho_dry = nullptr, *T = nullptr, *w1 = nullptr, *q_surf = nullptr, *Nd = nullptr;
    void *aux_Naer = nullptr, *out = nullptr, *staging = nullptr;
    void *prm_sets = nullptr, *col_to_set = nullptr;  // per-column parameter sets (calibration ensembles)
    void* colmap = nullptr;                            // kid_set_column_map: host column of every device column
    // the map is a permutation of ONE contiguous block of host rows [colmap_base, colmap_base + nc): the block travels through
    // the copy engine + staging and the layout kernel permutes on the device (colmap_local = map - base); -1: scattered rows
    void* colmap_local = nullptr;
    long long colmap_base = -1;
    void *ps_liq = nullptr, *ps_mix = nullptr;        // p_sat(T) per grid point (profile constants)
    float* gconst_G = nullptr;                         // G(T) per grid point (inside the gconst allocation)
    void* gconst = nullptr;                            // float4 per grid point: constants of the packed 2M kernel (kid_pair_kernel.cuh)
    CUtensorMap state_map{};                           // TMA descriptor of the device state [var][level][column] (box 32 x 32 x nv)
    bool state_map_ok = false;
    bool consts_dirty = true;
    size_t staging_bytes = 0, out_bytes = 0;
    int prof_per_column = 0;
    DevSwitches sw{};
    const LaunchTable* lt = nullptr;
    int steps_per_launch = 0;
    int64_t launches = 0;
    std::vector<double> params;
    DevParams<float> prm_f{};   // the shared parameter set, passed to every launch as a kernel argument
    DevParams<double> prm_d{};
    bool params_set = false, profiles_set = false, state_set = false;
    int C = 32, NW = 1;
    // K2D
    void *U = nullptr, *Y3 = nullptr, *S = nullptr, *prev_aux = nullptr, *xc = nullptr, *out2 = nullptr;
    void* Y2 = nullptr;        // second state buffer of the single-launch horizontal kernel (swapped with Y after every stage)
    double k2d_t_stage = 0.0;  // stage time of the last k2d_begin (the horizontal kernel of k2d_end needs ρu(t))
    // in-library neighbour exchange of a decomposed domain (kid_k2d_comm_init): NCCL communicator of the x-ring
    struct ncclComm* comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    double *zsin = nullptr, *cosx = nullptr, *cosz = nullptr;
    void *send_left = nullptr, *send_right = nullptr, *recv_left = nullptr, *recv_right = nullptr;
    int Nq = 0, nfields = 0;
    K2DField fields[KID_MAX_DSS];
    double gll_x[KID_MAX_NQ], gll_w[KID_MAX_NQ], gll_D[KID_MAX_NQ * KID_MAX_NQ];
    // aux output: fields of the next diag launch; asynchronous double-buffered download (kid_output_*)
    int diag_nfields = 1, diag_fields[KID_NAUX];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    void* ostage[2] = {nullptr, nullptr};
    size_t ostage_bytes[2] = {0, 0};
    bool ev_used[2] = {false, false};
    int obuf = 0;
    // calibration observables (kid_obs_*)
    double* obsG = nullptr;
    unsigned long long* red = nullptr;
    int obs_nvars = 0, obs_nt_filtered = 0, obs_nt_per = 1;
    int64_t obs_n_out = 0;
    int obs_var[KID_MAX_OBS], obs_per[KID_MAX_OBS], obs_off[KID_MAX_OBS + 1];
    int i_tot = 0, i_liq = -1, i_ice = -1, i_rai = -1, i_sno = -1, i_Nl = -1, i_Nr = -1, i_Na = -1;
};

#define KID_ON_DEVICE(h) DeviceGuard kid_device_guard__((h) ? (h)->cfg.device : 0)

namespace {

size_t state_elems(const kid_handle* h) { return size_t(h->nv) * h->cfg.nz * h->ncp; }

int ensure_staging(kid_handle* h, size_t bytes) {
    if (h->staging_bytes >= bytes) return 0;
    if (h->staging) cudaFree(h->staging);
    h->staging = nullptr;
    h->staging_bytes = 0;
    KID_CUDA(cudaMalloc(&h->staging, bytes));
    h->staging_bytes = bytes;
    return 0;
}
int ensure_out(kid_handle* h, size_t bytes) {
    if (h->out_bytes >= bytes) return 0;
    if (h->out) cudaFree(h->out);
    h->out = nullptr;
    h->out_bytes = 0;
    KID_CUDA(cudaMalloc(&h->out, bytes));
    h->out_bytes = bytes;
    return 0;
}

template <class FT>
void fill_const(std::vector<FT>& v, size_t n, double x) { v.assign(n, FT(x)); }

// upload a host vector of nc per-column values into a padded [ncp] device array
int upload_columns(kid_handle* h, void* dst, const void* src_ft, double fallback) {
    const int nc = h->cfg.nc, ncp = h->ncp;
    if (h->f64) {
        std::vector<double> v(ncp, fallback);
        if (src_ft) std::memcpy(v.data(), src_ft, size_t(nc) * 8);
        for (int i = nc; i < ncp; ++i) v[i] = v[nc - 1];
        KID_CUDA(cudaMemcpyAsync(dst, v.data(), size_t(ncp) * 8, cudaMemcpyHostToDevice, h->stream));
        KID_CUDA(cudaStreamSynchronize(h->stream));
    } else {
        std::vector<float> v(ncp, float(fallback));
        if (src_ft) std::memcpy(v.data(), src_ft, size_t(nc) * 4);
        for (int i = nc; i < ncp; ++i) v[i] = v[nc - 1];
        KID_CUDA(cudaMemcpyAsync(dst, v.data(), size_t(ncp) * 4, cudaMemcpyHostToDevice, h->stream));
        KID_CUDA(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

// Is `p` page-locked host memory the device can address (cudaHostAlloc / kid_host_alloc / cudaHostRegister)?  Then the layout
// kernels CAN read / write it in place over PCIe.  Measured on B200 (profiles/README.md, round 2): the copy engine + a
// transpose from device staging moves a 4 GiB state at 97 % of the resident stepping rate end to end, in-place access from
// the kernels (512-byte requests) at 73 % — so in-place access is used where it is needed (column maps: scattered host rows)
// or asked for (cfg.reserved[4] = 2), and plain transfers stage through device memory.
const void* device_view_of_pinned(const kid_handle* h, const void* p, bool needed) {
    if (!needed && h->cfg.reserved[4] != 2) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (a.type == cudaMemoryTypeHost && a.devicePointer) ? a.devicePointer : nullptr;
}
template <bool TO_DEVICE>
int launch_layout(kid_handle* h, const void* hostside, void* devside, int nv, const int* perm) {
    const int nc = h->cfg.nc, nz = h->cfg.nz, R = nv * nz;
    if (!h->f64 && R % 128 == 0 && (reinterpret_cast<size_t>(hostside) % 16) == 0) {
        dim3 grid((h->ncp + 31) / 32, R / 128);
        if (TO_DEVICE) host_to_device_layout_wide<<<grid, 256, 0, h->stream>>>((const float*)hostside, (float*)devside, perm, nc, h->ncp, R);
        else device_to_host_layout_wide<<<grid, 256, 0, h->stream>>>((const float*)devside, (float*)hostside, perm, nc, h->ncp, R);
    } else {
        dim3 block(32, 8), grid((h->ncp + 31) / 32, (R + 31) / 32);
        if (h->f64) {
            if (TO_DEVICE) host_to_device_layout<double><<<grid, block, 0, h->stream>>>((const double*)hostside, (double*)devside, perm, nc, h->ncp, nv, nz);
            else device_to_host_layout<double><<<grid, block, 0, h->stream>>>((const double*)devside, (double*)hostside, perm, nc, h->ncp, nv, nz);
        } else {
            if (TO_DEVICE) host_to_device_layout<float><<<grid, block, 0, h->stream>>>((const float*)hostside, (float*)devside, perm, nc, h->ncp, nv, nz);
            else device_to_host_layout<float><<<grid, block, 0, h->stream>>>((const float*)devside, (float*)hostside, perm, nc, h->ncp, nv, nz);
        }
    }
    KID_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}
// host [nc][nv][nz] -> device [nv][nz][ncp]; with a column map: device column i <- host column map[i]
int to_device_layout(kid_handle* h, const void* host, void* dev, int nv, bool use_map = false) {
    const int nc = h->cfg.nc, nz = h->cfg.nz;
    const int* perm = use_map ? (const int*)h->colmap : nullptr;
    size_t bytes = size_t(nc) * nv * nz * h->fsz;
    if (perm && h->colmap_base >= 0 && h->cfg.reserved[4] != 2) {
        // block-local map: the contiguous block of host rows goes through the copy engine, the permutation happens on the device
        if (ensure_staging(h, bytes)) return 1;
        const char* src = static_cast<const char*>(host) + size_t(h->colmap_base) * nv * nz * h->fsz;
        KID_CUDA(cudaMemcpyAsync(h->staging, src, bytes, cudaMemcpyHostToDevice, h->stream));
        return launch_layout<true>(h, h->staging, dev, nv, (const int*)h->colmap_local);
    }
    if (const void* hv = device_view_of_pinned(h, host, perm != nullptr)) return launch_layout<true>(h, hv, dev, nv, perm);
    if (perm) return fail("a scattered column map (kid_set_column_map) needs page-locked host buffers: allocate them with kid_host_alloc");
    if (ensure_staging(h, bytes)) return 1;
    KID_CUDA(cudaMemcpyAsync(h->staging, host, bytes, cudaMemcpyHostToDevice, h->stream));
    return launch_layout<true>(h, h->staging, dev, nv, nullptr);
}
// device [nv][nz][ncp] -> host [nc][nv][nz]
int to_host_layout(kid_handle* h, const void* dev, void* host, int nv, bool sync = true, bool use_map = false) {
    const int nc = h->cfg.nc, nz = h->cfg.nz;
    const int* perm = use_map ? (const int*)h->colmap : nullptr;
    size_t bytes = size_t(nc) * nv * nz * h->fsz;
    if (perm && h->colmap_base >= 0 && h->cfg.reserved[4] != 2) {
        if (ensure_staging(h, bytes)) return 1;
        if (launch_layout<false>(h, h->staging, const_cast<void*>(dev), nv, (const int*)h->colmap_local)) return 1;
        char* dst = static_cast<char*>(host) + size_t(h->colmap_base) * nv * nz * h->fsz;
        KID_CUDA(cudaMemcpyAsync(dst, h->staging, bytes, cudaMemcpyDeviceToHost, h->stream));
        if (sync) KID_CUDA(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (const void* hv = device_view_of_pinned(h, host, perm != nullptr)) {
        if (launch_layout<false>(h, hv, const_cast<void*>(dev), nv, perm)) return 1;
        if (sync) KID_CUDA(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (perm) return fail("a scattered column map (kid_set_column_map) needs page-locked host buffers: allocate them with kid_host_alloc");
    if (ensure_staging(h, bytes)) return 1;
    if (launch_layout<false>(h, h->staging, const_cast<void*>(dev), nv, nullptr)) return 1;
    KID_CUDA(cudaMemcpyAsync(host, h->staging, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (sync) KID_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

template <class FT>
KArgs<FT> make_args(kid_handle* h, double t0, int nsteps, int diag_mode, int field) {
    KArgs<FT> a{};
    a.Y = (FT*)h->Y;
    a.rho_dry = (const FT*)h->rho_dry;
    a.T = (const FT*)h->T;
    a.w1 = (const FT*)h->w1;
    a.q_surf = (const FT*)h->q_surf;
    a.Nd = (const FT*)h->Nd;
    a.aux_Naer = (const FT*)h->aux_Naer;
    a.ps_liq = (const FT*)h->ps_liq;
    a.ps_mix = (const FT*)h->ps_mix;
    a.prm_sets = (const DevParams<FT>*)h->prm_sets;
    a.col_to_set = (const int*)h->col_to_set;
    a.out = (FT*)h->out;
    a.nz = h->cfg.nz;
    a.nc = h->cfg.nc;
    a.ncp = h->ncp;
    a.prof_per_column = h->prof_per_column;
    a.model = h->cfg.model;
    a.nsteps = nsteps;
    a.C = h->C;
    a.NW = h->NW;
    a.diag_mode = diag_mode;
    a.no_fixed_shape = h->cfg.reserved[2];
    a.nfields = (diag_mode == 2) ? h->diag_nfields : 0;
    for (int f = 0; f < a.nfields; ++f) a.fields[f] = h->diag_fields[f];
    a.field = field;
    a.dt = FT(h->cfg.dt);
    a.dz = FT((h->cfg.z_max - h->cfg.z_min) / h->cfg.nz);
    a.t0 = t0;
    a.sw = h->sw;
    a.xc = (const FT*)h->xc;
    a.zsin = h->zsin;
    a.prev_aux = (FT*)h->prev_aux;
    a.outS = (FT*)h->S;
    a.update_prev = 0;
    a.width = h->cfg.domain_width;
    a.height = h->cfg.domain_height;
    a.z_min = h->cfg.z_min;
    return a;
}

template <class FT>
K2DArgs<FT> make_k2d_args(kid_handle* h, double t_stage, int stage, int use_recv, int rhs_only) {
    K2DArgs<FT> a{};
    a.Y = (FT*)h->Y; a.Yout = (FT*)h->Y2; a.U = (FT*)h->U; a.dY = (FT*)h->out; a.S = (FT*)h->S; a.out = (FT*)h->out2;
    a.rho_dry = (const FT*)h->rho_dry; a.w1 = (const FT*)h->w1; a.xc = (const FT*)h->xc;
    a.send_left = (FT*)h->send_left; a.send_right = (FT*)h->send_right;
    a.recv_left = (const FT*)h->recv_left; a.recv_right = (const FT*)h->recv_right;
    a.nz = h->cfg.nz; a.nc = h->cfg.nc; a.ncp = h->ncp; a.nv = h->nv; a.helem = h->cfg.helem; a.Nq = h->Nq;
    a.prof_per_column = h->prof_per_column;
    a.nfields = h->nfields;
    for (int f = 0; f < h->nfields; ++f) a.fields[f] = h->fields[f];
    a.i_tot = h->i_tot; a.i_rai = h->i_rai; a.i_sno = h->i_sno; a.precip = h->cfg.precip;
    a.stage = stage; a.use_recv = use_recv; a.rhs_only = rhs_only;
    a.dt = FT(h->cfg.dt); a.t = FT(t_stage); a.t1 = FT(h->params[KID_P_kid_t1]);
    a.cosx = h->cosx; a.cosz = h->cosz;
    a.ts = (a.t < a.t1) ? std::sin(M_PI * double(a.t) / double(a.t1)) : 0.0;
    a.z_min = FT(h->cfg.z_min); a.dz = FT((h->cfg.z_max - h->cfg.z_min) / h->cfg.nz);
    a.width = h->cfg.domain_width; a.height = h->cfg.domain_height;
    a.dxe = h->cfg.domain_width / h->cfg.helem_global;
    for (int i = 0; i < h->Nq * h->Nq; ++i) a.D[i] = FT(h->gll_D[i]);
    for (int i = 0; i < h->Nq; ++i) a.w[i] = FT(h->gll_w[i]);
    return a;
}

// The packed two-levels-per-thread step kernel serves the headline configuration: K1DModel / col_sed, Precipitation2M with
// SB2006 sedimentation, Float32, the default 128-level grid, one shared parameter set, full 32-column tiles.
// cfg.reserved[3] = 1 opts out (A/B measurements against the scalar tile kernel).
bool pair_capable(const kid_handle* h) {
    return !h->f64 && h->cfg.precip == KID_PRECIP_2M && h->cfg.nz == 128 && h->prm_sets == nullptr && h->lt->pair_f != nullptr &&
           (h->cfg.model == KID_MODEL_K1D || h->cfg.model == KID_MODEL_K1D_COL_SED) && h->cfg.sedimentation != KID_SED_CHEN2022 &&
           h->C == 32 && h->cfg.reserved[3] == 0;
}

// TMA descriptor of the device state for k1d_pair_kernel: 3-D tensor (columns, levels, variables), Float32, box = one
// warp's share of a round of one tile (PAIR_C columns x 2·32/PAIR_C levels x all variables).  cuTensorMapEncodeTiled is a driver entry point: fetched
// through the runtime so that the library keeps linking against the static cudart only.
int make_state_map(kid_handle* h) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    KID_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[3] = {cuuint64_t(h->ncp), cuuint64_t(h->cfg.nz), cuuint64_t(h->nv)};
    const cuuint64_t gstride[2] = {cuuint64_t(h->ncp) * 4, cuuint64_t(h->ncp) * h->cfg.nz * 4};
    const cuuint32_t box[3] = {cuuint32_t(PAIR_C), 2 * cuuint32_t(32 / PAIR_C), cuuint32_t(h->nv)};  // one warp's levels of a round
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&h->state_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->Y, gdim, gstride, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
    h->state_map_ok = true;
    return 0;
}

int check_ready(kid_handle* h) {
    if (!h) return fail("null handle");
    if (!h->params_set) return fail("kid_set_params has not been called");
    if (!h->profiles_set) return fail("kid_set_profiles has not been called");
    if (!h->state_set) return fail("kid_set_state has not been called");
    if (h->consts_dirty) {
        // (re)evaluate the time-invariant profile constants after params or profiles changed; a shared profile
        // with per-column thermodynamic parameter sets is not supported (the constants would differ per column)
        cudaError_t e;
        if (h->f64) e = h->lt->consts_d(make_args<double>(h, 0.0, 0, 0, 0), h->prm_d, (double*)h->ps_liq, (double*)h->ps_mix, h->stream);
        else e = h->lt->consts_f(make_args<float>(h, 0.0, 0, 0, 0), h->prm_f, (float*)h->ps_liq, (float*)h->ps_mix, h->stream);
        if (e != cudaSuccess) return fail(std::string("profile_consts_kernel launch: ") + cudaGetErrorString(e));
        h->launches++;
        if (pair_capable(h)) {
            if (!h->state_map_ok && make_state_map(h)) return 1;
            const size_t n = h->prof_per_column ? size_t(h->cfg.nz) * h->ncp : size_t(h->cfg.nz);
            if (h->gconst) cudaFree(h->gconst);
            h->gconst = nullptr;
            // float4 (ps_mix, 1/ps_liq, λ, T) per grid point, followed by the G(T) array
            KID_CUDA(cudaMalloc(&h->gconst, n * (sizeof(float4) + sizeof(float))));
            KID_CUDA(cudaMemsetAsync(h->gconst, 0, n * (sizeof(float4) + sizeof(float)), h->stream));
            h->gconst_G = (float*)((float4*)h->gconst + n);
            e = h->lt->gconsts_f(make_args<float>(h, 0.0, 0, 0, 0), h->prm_f, (float4*)h->gconst, h->gconst_G, h->stream);
            if (e != cudaSuccess) return fail(std::string("grid_consts_kernel launch: ") + cudaGetErrorString(e));
            h->launches++;
        }
        h->consts_dirty = false;
    }
    return 0;
}

// tile geometry: the largest power-of-two number of columns per CTA whose tile fits in shared memory, and as
// many levels per round (threads per column) as the thread budget allows
int choose_tile(kid_handle* h) {
    int dev = 0, max_smem = 0;
    KID_CUDA(cudaGetDevice(&dev));
    KID_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int nz = h->cfg.nz, moist = h->cfg.moisture;
    const bool pcp = h->prm_sets != nullptr;
    int C = 32;
    // few columns (the drivers' single-column runs): a narrower tile puts more level-warps on each column, so a stage
    // needs fewer rounds (barriers) — one column x 128 levels is one round of 128 threads
    while (C / 2 >= h->cfg.nc && C > 1) C /= 2;
    if (h->cfg.reserved[1] > 0) C = std::max(1, std::min(32, int(h->cfg.reserved[1])));  // tuning override
    while (C > 1) {
        size_t s = h->f64 ? h->lt->k1d_smem_d(moist, nz, C, pcp) : h->lt->k1d_smem_f(moist, nz, C, pcp);
        if (s <= size_t(max_smem)) break;
        C /= 2;
    }
    size_t s = h->f64 ? h->lt->k1d_smem_d(moist, nz, C, pcp) : h->lt->k1d_smem_f(moist, nz, C, pcp);
    if (s > size_t(max_smem)) return fail("column too tall for one shared-memory tile (nz too large)");
    int maxt = h->f64 ? tile_max_threads<double>() : tile_max_threads<float>();
    int NW = std::max(2, std::min(maxt / C, nz));
    if (h->cfg.reserved[0] > 0) NW = std::max(2, std::min(NW, int(h->cfg.reserved[0])));  // tuning override (the round-boundary stash needs NW >= 2)
    h->C = C;
    h->NW = NW;
    return 0;
}

int launch_tile(kid_handle* h, double t0, int nsteps, int diag_mode, int field, int update_prev = 0, void* Y_stage = nullptr) {
    if (diag_mode == 2 && field >= 0) { h->diag_nfields = 1; h->diag_fields[0] = field; }  // field < 0: h->diag_fields preset
    cudaError_t e;
    if (diag_mode == 0 && pair_capable(h) && h->gconst && h->state_map_ok) {
        e = h->lt->pair_f(h->cfg.moisture, make_args<float>(h, t0, nsteps, 0, 0), (const float4*)h->gconst, h->gconst_G, h->state_map, h->prm_f, h->stream);
        if (e != cudaSuccess) return fail(std::string("k1d_pair_kernel launch: ") + cudaGetErrorString(e));
        h->launches++;
        return 0;
    }
    if (h->f64) {
        auto a = make_args<double>(h, t0, nsteps, diag_mode, field);
        a.update_prev = update_prev;
        if (Y_stage) a.Y = (double*)Y_stage;
        e = h->lt->k1d_d(h->cfg.moisture, diag_mode != 0, a, h->prm_d, h->stream);
    } else {
        auto a = make_args<float>(h, t0, nsteps, diag_mode, field);
        a.update_prev = update_prev;
        if (Y_stage) a.Y = (float*)Y_stage;
        e = h->lt->k1d_f(h->cfg.moisture, diag_mode != 0, a, h->prm_f, h->stream);
    }
    if (e != cudaSuccess) return fail(std::string("k1d_tile_kernel launch: ") + cudaGetErrorString(e));
    h->launches++;
    return 0;
}
int launch_box(kid_handle* h, double t0, int nsteps) {
    cudaError_t e;
    if (h->f64) e = h->lt->box_d(h->cfg.moisture, make_args<double>(h, t0, nsteps, 0, 0), h->prm_d, h->stream);
    else e = h->lt->box_f(h->cfg.moisture, make_args<float>(h, t0, nsteps, 0, 0), h->prm_f, h->stream);
    if (e != cudaSuccess) return fail(std::string("box_step_kernel launch: ") + cudaGetErrorString(e));
    h->launches++;
    return 0;
}

// ---- K2D: one SSPRK33 stage = vertical/pointwise rhs (tile kernel) -> horizontal weak divergence + edge pack
//      -> [neighbour exchange by the caller] -> DSS + stage combine
double k2d_stage_time(kid_handle* h, double t_step, int stage) {
    // the stage times as OrdinaryDiffEq forms them in FLOAT_TYPE arithmetic
    if (h->f64) { double t = t_step, dt = h->cfg.dt; return stage == 0 ? t : (stage == 1 ? t + dt : t + dt / 2.0); }
    float t = float(t_step), dt = float(h->cfg.dt);
    return double(stage == 0 ? t : (stage == 1 ? t + dt : t + dt / 2.0f));
}
// style dispatch of the templated K2D kernels
#define KID_K2D_STYLE_CASE(M, P, ...) if (h->cfg.moisture == M && h->cfg.precip == P) { constexpr int MOIST = M, PRECIP = P; __VA_ARGS__; }
#define KID_K2D_DISPATCH(...)                                                                   \
    KID_K2D_STYLE_CASE(KID_MOIST_EQ, KID_PRECIP_NONE, __VA_ARGS__)                              \
    KID_K2D_STYLE_CASE(KID_MOIST_EQ, KID_PRECIP_0M, __VA_ARGS__)                                \
    KID_K2D_STYLE_CASE(KID_MOIST_EQ, KID_PRECIP_1M, __VA_ARGS__)                                \
    KID_K2D_STYLE_CASE(KID_MOIST_EQ, KID_PRECIP_2M, __VA_ARGS__)                                \
    KID_K2D_STYLE_CASE(KID_MOIST_NONEQ, KID_PRECIP_NONE, __VA_ARGS__)                           \
    KID_K2D_STYLE_CASE(KID_MOIST_NONEQ, KID_PRECIP_0M, __VA_ARGS__)                             \
    KID_K2D_STYLE_CASE(KID_MOIST_NONEQ, KID_PRECIP_1M, __VA_ARGS__)                             \
    KID_K2D_STYLE_CASE(KID_MOIST_NONEQ, KID_PRECIP_2M, __VA_ARGS__)

// Single-launch horizontal part (k2d_horizontal_kernel): needs the Nq nodes of an element in one warp.  cfg.reserved[5] = 1
// keeps the two legacy launches (k2d_hdiv_kernel + k2d_dss_update_kernel), which kid_rhs always uses.
bool k2d_single_launch(const kid_handle* h) { return (32 % h->Nq == 0) && h->cfg.reserved[5] == 0; }

int k2d_begin(kid_handle* h, double t_stage, int /*stage*/, bool pack_edges, bool legacy) {
    if (ensure_out(h, state_elems(h) * h->fsz)) return 1;
    if (launch_tile(h, t_stage, 1, 1, 0, 1)) return 1;
    h->k2d_t_stage = t_stage;
    if (!legacy && k2d_single_launch(h)) {
        if (pack_edges) {  // decomposed domain: the rank-edge tendencies for the neighbours' DSS
            dim3 block(128), grid((2 * KID_MAX_DSS * h->cfg.nz + 127) / 128);
            if (h->f64) { KID_K2D_DISPATCH((k2d_edge_pack_kernel<double, MOIST, PRECIP><<<grid, block, 0, h->stream>>>(make_k2d_args<double>(h, t_stage, 0, 0, 0)))) }
            else { KID_K2D_DISPATCH((k2d_edge_pack_kernel<float, MOIST, PRECIP><<<grid, block, 0, h->stream>>>(make_k2d_args<float>(h, t_stage, 0, 0, 0)))) }
            KID_CUDA(cudaGetLastError());
            h->launches++;
        }
        return 0;
    }
    dim3 block(64), grid((h->cfg.helem + 63) / 64, h->cfg.nz);
    if (h->f64) k2d_hdiv_kernel<double><<<grid, block, 0, h->stream>>>(make_k2d_args<double>(h, t_stage, 0, 0, 0));
    else k2d_hdiv_kernel<float><<<grid, block, 0, h->stream>>>(make_k2d_args<float>(h, t_stage, 0, 0, 0));
    KID_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}
int k2d_end(kid_handle* h, int stage, int use_recv, int rhs_only) {
    if (!rhs_only && k2d_single_launch(h)) {
        const int per_block = 128 - 2 * h->Nq;  // the first and last element of a block are halo
        dim3 block(128), grid((h->cfg.nc + per_block - 1) / per_block, h->cfg.nz);
        if (h->f64) { KID_K2D_DISPATCH((k2d_horizontal_kernel<double, MOIST, PRECIP><<<grid, block, 0, h->stream>>>(make_k2d_args<double>(h, h->k2d_t_stage, stage, use_recv, 0)))) }
        else { KID_K2D_DISPATCH((k2d_horizontal_kernel<float, MOIST, PRECIP><<<grid, block, 0, h->stream>>>(make_k2d_args<float>(h, h->k2d_t_stage, stage, use_recv, 0)))) }
        KID_CUDA(cudaGetLastError());
        h->launches++;
        std::swap(h->Y, h->Y2);  // the kernel wrote the new stage state into the second buffer
        return 0;
    }
    dim3 block(128), grid((h->cfg.nc + 127) / 128, h->cfg.nz);
    if (h->f64) k2d_dss_update_kernel<double><<<grid, block, 0, h->stream>>>(make_k2d_args<double>(h, 0.0, stage, use_recv, rhs_only));
    else k2d_dss_update_kernel<float><<<grid, block, 0, h->stream>>>(make_k2d_args<float>(h, 0.0, stage, use_recv, rhs_only));
    KID_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}
int k2d_setup(kid_handle* h) {
    const kid_config_t& c = h->cfg;
    if (c.npoly < 1 || c.npoly + 1 > KID_MAX_NQ) return fail("K2D: npoly must be in [1, 7]");
    if (c.helem < 1 || c.helem_global < c.helem || c.helem_offset < 0 || c.helem_offset + c.helem > c.helem_global)
        return fail("K2D: inconsistent helem / helem_global / helem_offset");
    if (c.nc != c.helem * (c.npoly + 1)) return fail("K2D: nc must equal helem * (npoly + 1)");
    if (!(c.domain_width > 0) || !(c.domain_height > 0)) return fail("K2D: domain_width and domain_height must be positive");
    h->Nq = c.npoly + 1;
    gll_rule(h->Nq, h->gll_x, h->gll_w, h->gll_D);
    // state variable indices (src/Common/ode_utils.jl:23-54)
    int n = 1;
    if (c.moisture == KID_MOIST_NONEQ) { h->i_liq = n++; h->i_ice = n++; }
    if (c.precip >= KID_PRECIP_1M) { h->i_rai = n++; h->i_sno = n++; }
    if (c.precip == KID_PRECIP_2M) { h->i_Nl = n++; h->i_Nr = n++; h->i_Na = n++; }
    // DSS'd / horizontally advected fields (src/K2DModel/tendency.jl:57-213)
    int f = 0;
    h->fields[f++] = {h->i_tot, 0, h->i_tot};
    if (c.moisture == KID_MOIST_NONEQ) { h->fields[f++] = {h->i_liq, 0, h->i_liq}; h->fields[f++] = {h->i_ice, 0, h->i_ice}; }
    if (c.precip == KID_PRECIP_1M) { h->fields[f++] = {h->i_rai, 1, 0}; h->fields[f++] = {h->i_sno, 1, 1}; }
    if (c.precip == KID_PRECIP_2M) { h->fields[f++] = {h->i_Nr, 0, h->i_Nr}; h->fields[f++] = {h->i_rai, 1, 0}; }
    h->nfields = f;
    size_t st = state_elems(h) * h->fsz, plane = size_t(c.nz) * h->ncp * h->fsz, edge = size_t(KID_MAX_DSS) * c.nz * h->fsz;
    KID_CUDA(cudaMalloc(&h->U, st));
    KID_CUDA(cudaMalloc(&h->out2, st));
    KID_CUDA(cudaMalloc(&h->Y2, st));
    KID_CUDA(cudaMemsetAsync(h->Y2, 0, st, h->stream));
    KID_CUDA(cudaMalloc(&h->S, 2 * plane));
    KID_CUDA(cudaMalloc(&h->prev_aux, 3 * plane));
    KID_CUDA(cudaMemsetAsync(h->S, 0, 2 * plane, h->stream));
    KID_CUDA(cudaMemsetAsync(h->prev_aux, 0, 3 * plane, h->stream));
    for (void** p : {&h->send_left, &h->send_right, &h->recv_left, &h->recv_right}) {
        KID_CUDA(cudaMalloc(p, edge));
        KID_CUDA(cudaMemsetAsync(*p, 0, edge, h->stream));
    }
    KID_CUDA(cudaMalloc(&h->xc, size_t(h->ncp) * h->fsz));
    const double dxe = c.domain_width / c.helem_global;
    std::vector<double> x(h->ncp, 0.0);
    for (int e = 0; e < c.helem; ++e)
        for (int i = 0; i < h->Nq; ++i) x[e * h->Nq + i] = (e + c.helem_offset) * dxe + (h->gll_x[i] + 1.0) * 0.5 * dxe;
    if (h->f64) {
        KID_CUDA(cudaMemcpyAsync(h->xc, x.data(), size_t(h->ncp) * 8, cudaMemcpyHostToDevice, h->stream));
    } else {
        std::vector<float> xf(x.begin(), x.end());
        KID_CUDA(cudaMemcpyAsync(h->xc, xf.data(), size_t(h->ncp) * 4, cudaMemcpyHostToDevice, h->stream));
    }
    // vertical factor of the stream function at the faces (src/K2DModel/tendency.jl:4-9): sin(π z_f / H)
    std::vector<double> zs(c.nz + 1);
    const double dz = (c.z_max - c.z_min) / c.nz;
    for (int k = 0; k <= c.nz; ++k) zs[k] = std::sin(M_PI / c.domain_height * (c.z_min + double(h->f64 ? dz : double(float(dz))) * k));
    // factors of ρu at the cell centres / nodes: cos(π z_c / H), cos(2π x_c / W) with x_c as the kernels see it (FT)
    std::vector<double> cz(c.nz), cx(h->ncp, 0.0);
    const double dzf = h->f64 ? dz : double(float(dz));
    const double zmin_f = h->f64 ? c.z_min : double(float(c.z_min));
    for (int k = 0; k < c.nz; ++k) cz[k] = std::cos(M_PI / c.domain_height * (zmin_f + (k + 0.5) * dzf));
    for (int i = 0; i < h->ncp; ++i) cx[i] = std::cos(2.0 * M_PI / c.domain_width * (h->f64 ? x[i] : double(float(x[i]))));
    KID_CUDA(cudaMalloc(&h->cosz, cz.size() * sizeof(double)));
    KID_CUDA(cudaMalloc(&h->cosx, cx.size() * sizeof(double)));
    KID_CUDA(cudaMemcpyAsync(h->cosz, cz.data(), cz.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KID_CUDA(cudaMemcpyAsync(h->cosx, cx.data(), cx.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KID_CUDA(cudaMalloc(&h->zsin, zs.size() * sizeof(double)));
    KID_CUDA(cudaMemcpyAsync(h->zsin, zs.data(), zs.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KID_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// FFMA issue-rate probe (kid_measure_fp32_peak): 8 independent accumulators per thread, 2 CTAs of 1024 threads per SM
__global__ void __launch_bounds__(1024) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class FT>
int obs_launch(kid_handle_t* h, int time_cell) {
    ObsArgs<FT> a{};
    a.Y = (const FT*)h->Y;
    a.rho_dry = (const FT*)h->rho_dry;
    a.G = h->obsG;
    a.prm_sets = (const DevParams<FT>*)h->prm_sets;
    a.col_to_set = (const int*)h->col_to_set;
    a.nz = h->cfg.nz; a.nc = h->cfg.nc; a.ncp = h->ncp; a.prof_per_column = h->prof_per_column;
    a.nvars = h->obs_nvars; a.n_out = int(h->obs_n_out); a.nt_filtered = h->obs_nt_filtered; a.time_cell = time_cell;
    for (int j = 0; j < h->obs_nvars; ++j) { a.var_id[j] = h->obs_var[j]; a.nz_per_cell[j] = h->obs_per[j]; a.out_offset[j] = h->obs_off[j]; }
    a.out_offset[h->obs_nvars] = h->obs_off[h->obs_nvars];
    a.weight = 1.0 / double(h->obs_nt_per);
    a.z_span = h->cfg.z_max - h->cfg.z_min;
    a.sw = h->sw;
    cudaError_t e;
    if constexpr (sizeof(FT) == 8) e = h->lt->obs_d(h->cfg.moisture, a, h->prm_d, h->stream);
    else e = h->lt->obs_f(h->cfg.moisture, a, h->prm_f, h->stream);
    if (e != cudaSuccess) return fail(std::string("obs_accumulate_kernel: ") + cudaGetErrorString(e));
    h->launches++;
    return 0;
}

}  // namespace

extern "C" {

static int k2d_exchange_nccl(kid_handle_t* h);

int kid_nvars(int moisture, int precip) { return nvars_of(moisture, precip); }

int kid_var_names(int moisture, int precip, char* buf, size_t buflen) {
    int n = nvars_of(moisture, precip);
    if (n < 0) return fail("unsupported (moisture, precip) pair");
    std::string s = std::string("rhoq_tot") + '\0';
    if (moisture == KID_MOIST_NONEQ) s += std::string("rhoq_liq") + '\0' + std::string("rhoq_ice") + '\0';
    if (precip >= KID_PRECIP_1M) s += std::string("rhoq_rai") + '\0' + std::string("rhoq_sno") + '\0';
    if (precip == KID_PRECIP_2M) s += std::string("N_liq") + '\0' + std::string("N_rai") + '\0' + std::string("N_aer") + '\0';
    if (s.size() > buflen) return fail("buffer too small");
    std::memcpy(buf, s.data(), s.size());
    return 0;
}

int kid_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int kid_measure_fp32_peak(int32_t device, double* flops_per_s, int32_t* sm_count, int32_t* sm_clock_khz) {
    if (!flops_per_s) return fail("null argument");
    if (kid_device_count() < 1) return fail("no CUDA device visible: libkid_cuda has no CPU fallback");
    int prev = 0;
    KID_CUDA(cudaGetDevice(&prev));
    KID_CUDA(cudaSetDevice(device));
    int sms = 0, khz = 0;
    KID_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    KID_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    const int blocks = sms * 2, threads = 1024, iters = 40000;
    float* out = nullptr;
    KID_CUDA(cudaMalloc(&out, size_t(blocks) * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    KID_CUDA(cudaEventCreate(&e0));
    KID_CUDA(cudaEventCreate(&e1));
    fp32_peak_kernel<<<blocks, threads>>>(out, iters / 8, 1.0001f, 1e-7f);  // warm-up
    KID_CUDA(cudaEventRecord(e0));
    fp32_peak_kernel<<<blocks, threads>>>(out, iters, 1.0001f, 1e-7f);
    KID_CUDA(cudaEventRecord(e1));
    KID_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    KID_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    cudaSetDevice(prev);
    *flops_per_s = double(blocks) * threads * double(iters) * 8 * 2 / (double(ms) * 1e-3);
    if (sm_count) *sm_count = sms;
    if (sm_clock_khz) *sm_clock_khz = khz;
    return 0;
}

int kid_create(const kid_config_t* cfg, kid_handle_t** out) {
    if (!cfg || !out) return fail("null argument");
    *out = nullptr;
    if (cfg->abi_version != KID_ABI_VERSION) return fail("kid_config_t.abi_version mismatch");
    if (cfg->params_version != KID_PARAMS_VERSION) return fail("kid_config_t.params_version mismatch");
    int nv = nvars_of(cfg->moisture, cfg->precip);
    if (nv < 0) return fail("unsupported (moisture, precip) pair: only Equilibrium/NonEquilibrium moisture and No/0M/1M/2M precipitation are on the device path");
    if (cfg->model < KID_MODEL_BOX || cfg->model > KID_MODEL_K2D) return fail("invalid model");
    if (cfg->float_type != KID_F32 && cfg->float_type != KID_F64) return fail("invalid float_type");
    if (cfg->nz < 1 || cfg->nc < 1) return fail("nz and nc must be positive");
    if (cfg->model == KID_MODEL_BOX && cfg->nz != 1) return fail("BoxModel needs nz == 1");
    if ((cfg->model == KID_MODEL_K1D || cfg->model == KID_MODEL_K1D_COL_SED || cfg->model == KID_MODEL_K2D) && cfg->nz < 3)
        return fail("K1DModel needs at least 3 levels (Extrapolate boundary stencils)");
    if (!(cfg->dt > 0)) return fail("dt must be positive");
    if (cfg->sedimentation == KID_SED_CHEN2022 && cfg->precip != KID_PRECIP_2M && cfg->precip != KID_PRECIP_1M)
        return fail("sedimentation Chen2022 needs Precipitation1M or Precipitation2M");
    if (cfg->precip == KID_PRECIP_1M && (cfg->rain_formation < KID_RF_CLIMA_1M || cfg->rain_formation > KID_RF_VARTIMESCALE))
        return fail("invalid rain_formation for Precipitation1M");
    if (cfg->precip == KID_PRECIP_2M && cfg->rain_formation != KID_RF_SB2006 && cfg->rain_formation != KID_RF_SB2006NL)
        return fail("invalid rain_formation for Precipitation2M");
    if (kid_device_count() < 1) return fail("no CUDA device visible: libkid_cuda has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= kid_device_count()) return fail("kid_config_t.device out of range");
    DeviceGuard guard(cfg->device);  // the caller's current device is restored on return

    kid_handle* h = new kid_handle();
    h->cfg = *cfg;
    h->nv = nv;
    h->ncp = (cfg->nc + 31) / 32 * 32;
    h->f64 = cfg->float_type == KID_F64;
    h->fsz = h->f64 ? 8 : 4;
    h->lt = table_of(cfg->precip);
    h->sw.rain_formation = cfg->rain_formation;
    h->sw.sedimentation = cfg->sedimentation;
    h->sw.sb_limited = cfg->rain_formation != KID_RF_SB2006NL;
    h->sw.precip_sources = cfg->precip_sources;
    h->sw.precip_sinks = cfg->precip_sinks;
    h->sw.open_system_activation = cfg->open_system_activation;
    h->sw.local_activation = cfg->local_activation;
    h->sw.qtot_flux_correction = cfg->qtot_flux_correction;
    auto cleanup = [&](int rc) { kid_destroy(h); return rc; };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail("cudaStreamCreate failed"));
    size_t colb = size_t(h->ncp) * h->fsz;
    if (cudaMalloc(&h->Y, state_elems(h) * h->fsz) != cudaSuccess) return cleanup(fail("cudaMalloc(state) failed"));
    if (cudaMemsetAsync(h->Y, 0, state_elems(h) * h->fsz, h->stream) != cudaSuccess) return cleanup(fail("cudaMemset failed"));
    if (cudaMalloc(&h->w1, colb) != cudaSuccess || cudaMalloc(&h->q_surf, colb) != cudaSuccess ||
        cudaMalloc(&h->Nd, colb) != cudaSuccess)
        return cleanup(fail("cudaMalloc(column scalars) failed"));
    if (cudaMemsetAsync(h->q_surf, 0, colb, h->stream) != cudaSuccess) return cleanup(fail("cudaMemset failed"));
    if (choose_tile(h)) return cleanup(1);
    if (cfg->model == KID_MODEL_K2D && k2d_setup(h)) return cleanup(1);
    *out = h;
    return 0;
}

void kid_destroy(kid_handle_t* h) {
    KID_ON_DEVICE(h);
    if (!h) return;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && nccl_api().ok) nccl_api().CommDestroy(h->comm);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (int b = 0; b < 2; ++b) {
        if (h->ostage[b]) cudaFree(h->ostage[b]);
        if (h->ev_ready[b]) cudaEventDestroy(h->ev_ready[b]);
        if (h->ev_done[b]) cudaEventDestroy(h->ev_done[b]);
    }
    for (void* p : {h->colmap, h->colmap_local, h->gconst, h->ps_liq, h->ps_mix, h->prm_sets, h->col_to_set, h->Y, h->rho_dry, h->T, h->w1, h->q_surf, h->Nd, h->aux_Naer,
                    h->out, h->staging, h->U, h->Y3, h->Y2, h->S, h->prev_aux, h->xc, h->out2, h->send_left, h->send_right, h->recv_left,
                    h->recv_right, (void*)h->obsG, (void*)h->red, (void*)h->zsin, (void*)h->cosx, (void*)h->cosz})
        if (p) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int kid_set_params(kid_handle_t* h, const double* params, int32_t nparams, int32_t nsets, const int32_t* column_to_set) {
    KID_ON_DEVICE(h);
    if (!h || !params) return fail("null argument");
    if (nparams != KID_NPARAMS) return fail("kid_set_params: nparams != KID_NPARAMS (header/library version skew)");
    if (nsets < 1) return fail("kid_set_params: nsets must be >= 1");
    if (nsets > 1 && !column_to_set) return fail("kid_set_params: nsets > 1 needs column_to_set");
    if (nsets > 1 && h->cfg.model == KID_MODEL_K2D) return fail("kid_set_params: per-column parameter sets are not supported for K2DModel");
    h->params.assign(params, params + KID_NPARAMS);
    h->consts_dirty = true;
    for (void** p : {&h->prm_sets, &h->col_to_set}) { if (*p) cudaFree(*p); *p = nullptr; }
    if (nsets > 1) {
        const int nc = h->cfg.nc;
        std::vector<int> idx(h->ncp, 0);
        for (int c = 0; c < nc; ++c) {
            if (column_to_set[c] < 0 || column_to_set[c] >= nsets) return fail("kid_set_params: column_to_set entry out of range");
            idx[c] = column_to_set[c];
        }
        KID_CUDA(cudaMalloc(&h->col_to_set, size_t(h->ncp) * sizeof(int)));
        KID_CUDA(cudaMemcpyAsync(h->col_to_set, idx.data(), size_t(h->ncp) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        std::vector<double> w1(nc), nd(nc);
        for (int c = 0; c < nc; ++c) {
            w1[c] = params[size_t(idx[c]) * KID_NPARAMS + KID_P_kid_w1];
            nd[c] = params[size_t(idx[c]) * KID_NPARAMS + KID_P_com_prescribed_Nd];
        }
        // set 0 also becomes the shared kernel-argument set: it serves the profile constants of a SHARED profile
        // (thermodynamic parameters must then be common to all sets)
        if (h->f64) {
            std::vector<DevParams<double>> sets(nsets);
            for (int i = 0; i < nsets; ++i) fill_dev_params(sets[i], params + size_t(i) * KID_NPARAMS);
            h->prm_d = sets[0];
            KID_CUDA(cudaMalloc(&h->prm_sets, sets.size() * sizeof(sets[0])));
            KID_CUDA(cudaMemcpyAsync(h->prm_sets, sets.data(), sets.size() * sizeof(sets[0]), cudaMemcpyHostToDevice, h->stream));
            KID_CUDA(cudaStreamSynchronize(h->stream));
            if (upload_columns(h, h->w1, w1.data(), 0)) return 1;
            if (upload_columns(h, h->Nd, nd.data(), 0)) return 1;
        } else {
            std::vector<DevParams<float>> sets(nsets);
            for (int i = 0; i < nsets; ++i) fill_dev_params(sets[i], params + size_t(i) * KID_NPARAMS);
            h->prm_f = sets[0];
            KID_CUDA(cudaMalloc(&h->prm_sets, sets.size() * sizeof(sets[0])));
            KID_CUDA(cudaMemcpyAsync(h->prm_sets, sets.data(), sets.size() * sizeof(sets[0]), cudaMemcpyHostToDevice, h->stream));
            KID_CUDA(cudaStreamSynchronize(h->stream));
            std::vector<float> w1f(w1.begin(), w1.end()), ndf(nd.begin(), nd.end());
            if (upload_columns(h, h->w1, w1f.data(), 0)) return 1;
            if (upload_columns(h, h->Nd, ndf.data(), 0)) return 1;
        }
        if (choose_tile(h)) return 1;
        h->params_set = true;
        return 0;
    }
    if (choose_tile(h)) return 1;
    if (h->f64) fill_dev_params(h->prm_d, params);
    else fill_dev_params(h->prm_f, params);
    // per-column scalars default to the table values
    if (upload_columns(h, h->w1, nullptr, params[KID_P_kid_w1])) return 1;
    if (upload_columns(h, h->Nd, nullptr, params[KID_P_com_prescribed_Nd])) return 1;
    h->params_set = true;
    return 0;
}

int kid_set_profiles(kid_handle_t* h, const void* rho_dry, const void* T, int32_t per_column) {
    KID_ON_DEVICE(h);
    if (!h || !rho_dry || !T) return fail("null argument");
    const int nz = h->cfg.nz;
    size_t n = per_column ? size_t(nz) * h->ncp : size_t(nz);
    h->consts_dirty = true;
    h->profiles_set = false;  // until the new profiles are on the device
    for (void** p : {&h->rho_dry, &h->T, &h->ps_liq, &h->ps_mix}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
        KID_CUDA(cudaMalloc(p, n * h->fsz));
        KID_CUDA(cudaMemsetAsync(*p, 0, n * h->fsz, h->stream));
    }
    h->prof_per_column = per_column ? 1 : 0;
    if (per_column) {
        if (to_device_layout(h, rho_dry, h->rho_dry, 1)) return 1;
        KID_CUDA(cudaStreamSynchronize(h->stream));
        if (to_device_layout(h, T, h->T, 1)) return 1;
    } else {
        KID_CUDA(cudaMemcpyAsync(h->rho_dry, rho_dry, n * h->fsz, cudaMemcpyHostToDevice, h->stream));
        KID_CUDA(cudaMemcpyAsync(h->T, T, n * h->fsz, cudaMemcpyHostToDevice, h->stream));
    }
    KID_CUDA(cudaStreamSynchronize(h->stream));
    h->profiles_set = true;
    return 0;
}

int kid_set_column_scalar(kid_handle_t* h, int32_t which, const void* values) {
    KID_ON_DEVICE(h);
    if (!h) return fail("null handle");
    if (!h->params_set) return fail("kid_set_params must be called before kid_set_column_scalar");
    switch (which) {
        case KID_COL_W1: return upload_columns(h, h->w1, values, h->params[KID_P_kid_w1]);
        case KID_COL_Q_SURF: return upload_columns(h, h->q_surf, values, 0.0);
        case KID_COL_PRESCRIBED_ND: return upload_columns(h, h->Nd, values, h->params[KID_P_com_prescribed_Nd]);
    }
    return fail("invalid column scalar id");
}

static int set_state_impl(kid_handle_t* h, const void* Y, bool sync);
int kid_set_state(kid_handle_t* h, const void* Y) {
    KID_ON_DEVICE(h); return set_state_impl(h, Y, true); }
int kid_set_state_async(kid_handle_t* h, const void* Y) {
    KID_ON_DEVICE(h); return set_state_impl(h, Y, false); }
int kid_get_state_async(kid_handle_t* h, void* Y) {
    KID_ON_DEVICE(h);
    if (!h || !Y) return fail("null argument");
    if (!h->state_set) return fail("kid_set_state has not been called");
    return to_host_layout(h, h->Y, Y, h->nv, false, true);
}
static int set_state_impl(kid_handle_t* h, const void* Y, bool sync) {
    if (!h || !Y) return fail("null argument");
    if (to_device_layout(h, Y, h->Y, h->nv, true)) return 1;
    // aux.microph_variables.N_aer is built from the initial profile and never refreshed by
    // the Box / col_sed rhs (no precompute_aux_activation! call): snapshot it
    int na = na_index(h->cfg.moisture, h->cfg.precip);
    if (na >= 0 && (h->cfg.model == KID_MODEL_BOX || h->cfg.model == KID_MODEL_K1D_COL_SED)) {
        size_t plane = size_t(h->cfg.nz) * h->ncp * h->fsz;
        if (!h->aux_Naer) KID_CUDA(cudaMalloc(&h->aux_Naer, plane));
        KID_CUDA(cudaMemcpyAsync(h->aux_Naer, (const char*)h->Y + size_t(na) * plane, plane, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (h->cfg.model == KID_MODEL_K2D) {
        if (!h->profiles_set) return fail("kid_set_profiles must be called before kid_set_state for K2DModel");
        dim3 block(128), grid((h->cfg.nc + 127) / 128, h->cfg.nz);
        if (h->f64)
            k2d_init_prev_kernel<double><<<grid, block, 0, h->stream>>>((const double*)h->Y, (const double*)h->rho_dry, (double*)h->prev_aux,
                h->cfg.nz, h->cfg.nc, h->ncp, h->prof_per_column, h->i_tot, h->i_rai, h->i_Nl, h->i_Nr);
        else
            k2d_init_prev_kernel<float><<<grid, block, 0, h->stream>>>((const float*)h->Y, (const float*)h->rho_dry, (float*)h->prev_aux,
                h->cfg.nz, h->cfg.nc, h->ncp, h->prof_per_column, h->i_tot, h->i_rai, h->i_Nl, h->i_Nr);
        KID_CUDA(cudaGetLastError());
        h->launches++;
    }
    if (sync) KID_CUDA(cudaStreamSynchronize(h->stream));
    h->state_set = true;
    return 0;
}

int kid_get_state(kid_handle_t* h, void* Y) {
    KID_ON_DEVICE(h);
    if (!h || !Y) return fail("null argument");
    if (!h->state_set) return fail("kid_set_state has not been called");
    return to_host_layout(h, h->Y, Y, h->nv, true, true);
}

int kid_set_column_map(kid_handle_t* h, const int32_t* host_column_of) {
    KID_ON_DEVICE(h);
    if (!h) return fail("null handle");
    if (h->colmap) { cudaFree(h->colmap); h->colmap = nullptr; }
    if (h->colmap_local) { cudaFree(h->colmap_local); h->colmap_local = nullptr; }
    h->colmap_base = -1;
    if (!host_column_of) return 0;
    const int nc = h->cfg.nc;
    int lo = INT_MAX, hi = -1;
    for (int c = 0; c < nc; ++c) {
        if (host_column_of[c] < 0) return fail("kid_set_column_map: negative host column");
        lo = std::min(lo, host_column_of[c]);
        hi = std::max(hi, host_column_of[c]);
    }
    KID_CUDA(cudaMalloc(&h->colmap, size_t(nc) * sizeof(int)));
    KID_CUDA(cudaMemcpyAsync(h->colmap, host_column_of, size_t(nc) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    // a permutation of one contiguous block of host rows?  (max - min + 1 == nc and every row hit once)
    if (hi - lo + 1 == nc) {
        std::vector<int> local(nc);
        std::vector<char> seen(nc, 0);
        bool perm_ok = true;
        for (int c = 0; c < nc && perm_ok; ++c) {
            local[c] = host_column_of[c] - lo;
            perm_ok = !seen[local[c]];
            seen[local[c]] = 1;
        }
        if (perm_ok) {
            KID_CUDA(cudaMalloc(&h->colmap_local, size_t(nc) * sizeof(int)));
            KID_CUDA(cudaMemcpy(h->colmap_local, local.data(), size_t(nc) * sizeof(int), cudaMemcpyHostToDevice));
            h->colmap_base = lo;
        }
    }
    KID_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int kid_set_steps_per_launch(kid_handle_t* h, int32_t n) {
    KID_ON_DEVICE(h);
    if (!h) return fail("null handle");
    if (n < 0) return fail("steps_per_launch must be >= 0");
    h->steps_per_launch = n;
    return 0;
}

int kid_step(kid_handle_t* h, double t0, int32_t nsteps) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (nsteps < 0) return fail("nsteps must be >= 0");
    if (h->cfg.model == KID_MODEL_K2D) {
        const bool decomposed = h->cfg.helem != h->cfg.helem_global;
        if (decomposed && !h->comm)
            return fail("kid_step on a decomposed K2D handle: call kid_k2d_comm_init first (in-library NCCL edge exchange), or drive the "
                        "stages with kid_k2d_stage_begin/_end and exchange the edges yourself");
        for (int s = 0; s < nsteps; ++s) {
            double ts = h->f64 ? t0 + double(s) * h->cfg.dt : double(float(t0 + double(s) * h->cfg.dt));
            for (int stage = 0; stage < 3; ++stage) {
                if (k2d_begin(h, k2d_stage_time(h, ts, stage), stage, decomposed, false)) return 1;
                if (decomposed && k2d_exchange_nccl(h)) return 1;
                if (k2d_end(h, stage, decomposed ? 1 : 0, 0)) return 1;
            }
        }
        return 0;
    }
    int spl = h->steps_per_launch > 0 ? h->steps_per_launch : nsteps;
    int done = 0;
    while (done < nsteps) {
        int n = std::min(spl, nsteps - done);
        double t = t0 + double(done) * h->cfg.dt;
        int rc = (h->cfg.model == KID_MODEL_BOX) ? launch_box(h, t, n) : launch_tile(h, t, n, 0, 0);
        if (rc) return rc;
        done += n;
    }
    return 0;
}

int kid_rhs(kid_handle_t* h, double t, void* dY_out) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (!dY_out) return fail("null argument");
    if (ensure_out(h, state_elems(h) * h->fsz)) return 1;
    KID_CUDA(cudaMemsetAsync(h->out, 0, state_elems(h) * h->fsz, h->stream));
    if (h->cfg.model == KID_MODEL_K2D) {
        if (h->cfg.helem != h->cfg.helem_global) return fail("kid_rhs is only available on an undecomposed K2D handle");
        if (k2d_begin(h, t, 0, false, true)) return 1;
        if (k2d_end(h, 0, 0, 1)) return 1;
        return to_host_layout(h, h->out2, dY_out, h->nv);
    }
    if (launch_tile(h, t, 1, 1, 0)) return 1;
    return to_host_layout(h, h->out, dY_out, h->nv);
}

int kid_get_aux(kid_handle_t* h, double t, int32_t field, void* out) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (!out) return fail("null argument");
    if (field < 0 || field >= KID_NAUX) return fail("invalid aux field id");
    if (ensure_out(h, state_elems(h) * h->fsz)) return 1;
    if (launch_tile(h, t, 1, 2, field)) return 1;
    return to_host_layout(h, h->out, out, 1);
}

static int set_diag_fields(kid_handle_t* h, int32_t nfields, const int32_t* fields) {
    if (nfields < 1 || nfields > KID_NAUX || !fields) return fail("aux output: 1..KID_NAUX field ids");
    for (int f = 0; f < nfields; ++f) {
        if (fields[f] < 0 || fields[f] >= KID_NAUX) return fail("invalid aux field id");
        h->diag_fields[f] = fields[f];
    }
    h->diag_nfields = nfields;
    return 0;
}

int kid_get_aux_many(kid_handle_t* h, double t, int32_t nfields, const int32_t* fields, void* out) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (!out) return fail("null argument");
    if (set_diag_fields(h, nfields, fields)) return 1;
    if (ensure_out(h, std::max(state_elems(h), size_t(nfields) * h->cfg.nz * h->ncp) * h->fsz)) return 1;
    if (launch_tile(h, t, 1, 2, -1)) return 1;
    return to_host_layout(h, h->out, out, nfields);
}

int kid_output_enqueue(kid_handle_t* h, double t, int32_t nfields, const int32_t* fields, void* out_pinned) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (!out_pinned) return fail("null argument");
    if (set_diag_fields(h, nfields, fields)) return 1;
    const int nc = h->cfg.nc, nz = h->cfg.nz;
    if (!h->copy_stream) {
        KID_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            KID_CUDA(cudaEventCreateWithFlags(&h->ev_ready[b], cudaEventDisableTiming));
            KID_CUDA(cudaEventCreateWithFlags(&h->ev_done[b], cudaEventDisableTiming));
        }
    }
    const int b = h->obuf;
    h->obuf ^= 1;
    const size_t bytes = size_t(nc) * nfields * nz * h->fsz;
    // the staging buffer is reused every second output: the compute stream waits for its previous download
    if (h->ev_used[b]) KID_CUDA(cudaStreamWaitEvent(h->stream, h->ev_done[b], 0));
    if (h->ostage_bytes[b] < bytes) {
        if (h->ostage[b]) { KID_CUDA(cudaStreamSynchronize(h->copy_stream)); cudaFree(h->ostage[b]); }
        h->ostage[b] = nullptr; h->ostage_bytes[b] = 0;
        KID_CUDA(cudaMalloc(&h->ostage[b], bytes));
        h->ostage_bytes[b] = bytes;
    }
    if (ensure_out(h, std::max(state_elems(h), size_t(nfields) * nz * h->ncp) * h->fsz)) return 1;
    if (launch_tile(h, t, 1, 2, -1)) return 1;
    dim3 block(32, 8), grid((h->ncp + 31) / 32, (nfields * nz + 31) / 32);
    if (h->f64)
        device_to_host_layout<double><<<grid, block, 0, h->stream>>>((const double*)h->out, (double*)h->ostage[b], nullptr, nc, h->ncp, nfields, nz);
    else
        device_to_host_layout<float><<<grid, block, 0, h->stream>>>((const float*)h->out, (float*)h->ostage[b], nullptr, nc, h->ncp, nfields, nz);
    KID_CUDA(cudaGetLastError());
    h->launches++;
    KID_CUDA(cudaEventRecord(h->ev_ready[b], h->stream));
    KID_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_ready[b], 0));
    KID_CUDA(cudaMemcpyAsync(out_pinned, h->ostage[b], bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    KID_CUDA(cudaEventRecord(h->ev_done[b], h->copy_stream));
    h->ev_used[b] = true;
    return 0;
}

int kid_init_shipway_hill(kid_handle_t* h, const double* sounding, const void* p0, int32_t dry) {
    KID_ON_DEVICE(h);
    if (!h || !sounding) return fail("null argument");
    if (!h->params_set) return fail("kid_set_params must be called before kid_init_shipway_hill");
    if (h->cfg.model != KID_MODEL_K1D && h->cfg.model != KID_MODEL_K1D_COL_SED)
        return fail("kid_init_shipway_hill: K1DModel handles only");
    const int nz = h->cfg.nz, nc = h->cfg.nc;
    const size_t n = size_t(nz) * h->ncp;
    // arguments are checked before any device state is touched
    if (!(sounding[2] >= h->cfg.z_max)) return fail("z_max cannot be higher than z_2");
    if (!(sounding[1] > sounding[0]) || !(sounding[2] > sounding[1])) return fail("kid_init_shipway_hill: the sounding needs z_0 < z_1 < z_2");
    void* p0_dev = nullptr;
    KID_CUDA(cudaMalloc(&p0_dev, size_t(h->ncp) * h->fsz));
    struct FreeOnExit { void* p; ~FreeOnExit() { if (p) cudaFree(p); } } p0_owner{p0_dev};
    if (upload_columns(h, p0_dev, p0, 100000.0)) return 1;
    // from here on the previous profiles / state are gone: the handle is not ready until the kernel has been launched
    h->profiles_set = false;
    h->state_set = false;
    for (void** p : {&h->rho_dry, &h->T, &h->ps_liq, &h->ps_mix}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
        KID_CUDA(cudaMalloc(p, n * h->fsz));
        KID_CUDA(cudaMemsetAsync(*p, 0, n * h->fsz, h->stream));
    }
    h->prof_per_column = 1;
    h->consts_dirty = true;
    KID_CUDA(cudaMemsetAsync(h->Y, 0, state_elems(h) * h->fsz, h->stream));
    auto fill = [&](auto& a) {
        a.z_0 = sounding[0]; a.z_1 = sounding[1]; a.z_2 = sounding[2];
        a.rv_0 = sounding[3]; a.rv_1 = sounding[4]; a.rv_2 = sounding[5];
        a.th_0 = sounding[6]; a.th_1 = sounding[7]; a.th_2 = sounding[8];
        a.R_d = h->params[KID_P_td_R_d]; a.R_v = h->params[KID_P_td_R_v]; a.cp_d = h->params[KID_P_td_cp_d];
        a.cp_v = h->params[KID_P_td_cp_v]; a.grav = h->params[KID_P_td_grav]; a.MSLP = h->params[KID_P_td_MSLP];
        a.z_min = h->cfg.z_min; a.dz = (h->cfg.z_max - h->cfg.z_min) / nz;
        a.nz = nz; a.nc = nc; a.ncp = h->ncp; a.nv = h->nv; a.nsub = 4; a.dry = dry;
        a.i_tot = 0; a.i_Na = na_index(h->cfg.moisture, h->cfg.precip);
    };
    if (h->f64) {
        ICArgs<double> a{}; fill(a);
        a.p0 = (const double*)p0_dev; a.Nd = (const double*)h->Nd; a.rho_dry = (double*)h->rho_dry; a.T = (double*)h->T;
        a.Y = (double*)h->Y; a.q_surf = (double*)h->q_surf;
        init_shipway_hill_kernel<double><<<(nc + 63) / 64, 64, 0, h->stream>>>(a);
    } else {
        ICArgs<float> a{}; fill(a);
        a.p0 = (const float*)p0_dev; a.Nd = (const float*)h->Nd; a.rho_dry = (float*)h->rho_dry; a.T = (float*)h->T;
        a.Y = (float*)h->Y; a.q_surf = (float*)h->q_surf;
        init_shipway_hill_kernel<float><<<(nc + 63) / 64, 64, 0, h->stream>>>(a);
    }
    KID_CUDA(cudaGetLastError());
    h->launches++;
    const int na = na_index(h->cfg.moisture, h->cfg.precip);
    if (na >= 0 && h->cfg.model == KID_MODEL_K1D_COL_SED) {  // aux N_aer snapshot, as in kid_set_state
        const size_t plane = n * h->fsz;
        if (!h->aux_Naer) KID_CUDA(cudaMalloc(&h->aux_Naer, plane));
        KID_CUDA(cudaMemcpyAsync(h->aux_Naer, (const char*)h->Y + size_t(na) * plane, plane, cudaMemcpyDeviceToDevice, h->stream));
    }
    KID_CUDA(cudaStreamSynchronize(h->stream));
    h->profiles_set = true;
    h->state_set = true;
    return 0;
}

int kid_init_gamma_0d(kid_handle_t* h, const void* qt, const void* Nd, const void* k, const void* rho_dry) {
    KID_ON_DEVICE(h);
    if (!h || !qt || !Nd || !k || !rho_dry) return fail("null argument");
    if (!h->params_set) return fail("kid_set_params must be called before kid_init_gamma_0d");
    if (h->cfg.model != KID_MODEL_BOX && h->cfg.model != KID_MODEL_K1D_COL_SED)
        return fail("kid_init_gamma_0d: BoxModel / K1D col_sed handles only");
    if (h->cfg.moisture != KID_MOIST_NONEQ) return fail("kid_init_gamma_0d: the box initial condition has ρq_liq (NonEquilibriumMoisture)");
    const int nz = h->cfg.nz, nc = h->cfg.nc;
    const size_t n = size_t(nz) * h->ncp, colb = size_t(h->ncp) * h->fsz;
    void* in[4] = {nullptr, nullptr, nullptr, nullptr};
    struct FreeAll { void** p; ~FreeAll() { for (int i = 0; i < 4; ++i) if (p[i]) cudaFree(p[i]); } } owner{in};
    const void* src[4] = {qt, Nd, k, rho_dry};
    for (int i = 0; i < 4; ++i) {
        KID_CUDA(cudaMalloc(&in[i], colb));
        if (upload_columns(h, in[i], src[i], 0.0)) return 1;
    }
    h->profiles_set = false;
    h->state_set = false;
    for (void** p : {&h->rho_dry, &h->T, &h->ps_liq, &h->ps_mix}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
        KID_CUDA(cudaMalloc(p, n * h->fsz));
        KID_CUDA(cudaMemsetAsync(*p, 0, n * h->fsz, h->stream));
    }
    h->prof_per_column = 1;
    h->consts_dirty = true;
    KID_CUDA(cudaMemsetAsync(h->Y, 0, state_elems(h) * h->fsz, h->stream));
    int idx = 1, i_liq = idx++, i_ice = idx++;
    (void)i_ice;
    int i_rai = -1, i_Nl = -1, i_Nr = -1;
    if (h->cfg.precip >= KID_PRECIP_1M) { i_rai = idx; idx += 2; }
    if (h->cfg.precip == KID_PRECIP_2M) { i_Nl = idx++; i_Nr = idx++; }
    auto fill = [&](auto& a) {
        a.nz = nz; a.nc = nc; a.ncp = h->ncp;
        a.i_tot = 0; a.i_liq = i_liq; a.i_rai = i_rai; a.i_Nl = i_Nl; a.i_Nr = i_Nr;
    };
    if (h->f64) {
        IC0dArgs<double> a{}; fill(a);
        a.qt = (const double*)in[0]; a.Nd = (const double*)in[1]; a.k = (const double*)in[2]; a.rho_dry = (const double*)in[3];
        a.rho_dry_out = (double*)h->rho_dry; a.T_out = (double*)h->T; a.Y = (double*)h->Y;
        init_gamma_0d_kernel<double><<<(nc + 63) / 64, 64, 0, h->stream>>>(a);
    } else {
        IC0dArgs<float> a{}; fill(a);
        a.qt = (const float*)in[0]; a.Nd = (const float*)in[1]; a.k = (const float*)in[2]; a.rho_dry = (const float*)in[3];
        a.rho_dry_out = (float*)h->rho_dry; a.T_out = (float*)h->T; a.Y = (float*)h->Y;
        init_gamma_0d_kernel<float><<<(nc + 63) / 64, 64, 0, h->stream>>>(a);
    }
    KID_CUDA(cudaGetLastError());
    h->launches++;
    const int na = na_index(h->cfg.moisture, h->cfg.precip);
    if (na >= 0) {  // aux N_aer snapshot (= 0), as in kid_set_state
        const size_t plane = n * h->fsz;
        if (!h->aux_Naer) KID_CUDA(cudaMalloc(&h->aux_Naer, plane));
        KID_CUDA(cudaMemcpyAsync(h->aux_Naer, (const char*)h->Y + size_t(na) * plane, plane, cudaMemcpyDeviceToDevice, h->stream));
    }
    KID_CUDA(cudaStreamSynchronize(h->stream));
    h->profiles_set = true;
    h->state_set = true;
    return 0;
}

int kid_get_profiles(kid_handle_t* h, void* rho_dry, void* T) {
    KID_ON_DEVICE(h);
    if (!h || !rho_dry || !T) return fail("null argument");
    if (!h->profiles_set) return fail("no profiles on the device yet");
    if (h->prof_per_column) {
        if (to_host_layout(h, h->rho_dry, rho_dry, 1)) return 1;
        return to_host_layout(h, h->T, T, 1);
    }
    KID_CUDA(cudaMemcpyAsync(rho_dry, h->rho_dry, size_t(h->cfg.nz) * h->fsz, cudaMemcpyDeviceToHost, h->stream));
    KID_CUDA(cudaMemcpyAsync(T, h->T, size_t(h->cfg.nz) * h->fsz, cudaMemcpyDeviceToHost, h->stream));
    KID_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

void* kid_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
    return p;
}
void kid_host_free(void* p) { if (p) cudaFreeHost(p); }

int kid_output_wait(kid_handle_t* h) {
    KID_ON_DEVICE(h);
    if (!h) return fail("null handle");
    if (h->copy_stream) KID_CUDA(cudaStreamSynchronize(h->copy_stream));
    return 0;
}

int kid_reduce_max(kid_handle_t* h, double t, int32_t field, double* out) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (!out) return fail("null argument");
    if (field < 0 || field >= KID_NAUX) return fail("invalid aux field id");
    if (ensure_out(h, state_elems(h) * h->fsz)) return 1;
    if (launch_tile(h, t, 1, 2, field)) return 1;  // the field, [nz][ncp] on the device
    if (!h->red) KID_CUDA(cudaMalloc(&h->red, sizeof(unsigned long long)));
    KID_CUDA(cudaMemsetAsync(h->red, 0, sizeof(unsigned long long), h->stream));  // 0 orders below every double
    const int nz = h->cfg.nz, nc = h->cfg.nc;
    const int blocks = int(std::min<size_t>(1184, (size_t(nz) * h->ncp + 255) / 256));
    if (h->f64) field_max_kernel<double><<<blocks, 256, 0, h->stream>>>((const double*)h->out, nz, nc, h->ncp, h->red);
    else field_max_kernel<float><<<blocks, 256, 0, h->stream>>>((const float*)h->out, nz, nc, h->ncp, h->red);
    KID_CUDA(cudaGetLastError());
    h->launches++;
    unsigned long long bits = 0;
    KID_CUDA(cudaMemcpyAsync(&bits, h->red, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
    KID_CUDA(cudaStreamSynchronize(h->stream));
    *out = from_ordered_bits(bits);
    return 0;
}

int kid_obs_configure(kid_handle_t* h, int32_t nvars, const int32_t* var_ids, const int32_t* nz_per_filtered_cell,
                      int32_t nt_filtered, int32_t nt_per_filtered_cell, int64_t* n_out) {
    KID_ON_DEVICE(h);
    if (!h || !var_ids) return fail("null argument");
    if (h->cfg.model == KID_MODEL_K2D) return fail("kid_obs_*: K1D / Box handles only");
    if (nvars < 1 || nvars > KID_MAX_OBS) return fail("kid_obs_configure: 1..17 observables");
    if (nt_filtered < 1 || nt_per_filtered_cell < 1) return fail("kid_obs_configure: nt_filtered, nt_per_filtered_cell >= 1");
    const int nz = h->cfg.nz;
    const bool neq = h->cfg.moisture == KID_MOIST_NONEQ;
    const bool has_pr = h->cfg.precip == KID_PRECIP_1M || h->cfg.precip == KID_PRECIP_2M;
    const bool is2m = h->cfg.precip == KID_PRECIP_2M;
    int off = 0;
    for (int j = 0; j < nvars; ++j) {
        const int v = var_ids[j], per = nz_per_filtered_cell ? nz_per_filtered_cell[j] : 1;
        if (v < 0 || v >= KID_NOBS) return fail("kid_obs_configure: unknown observable id");
        if (per < 1 || nz % per != 0) return fail("kid_obs_configure: nz_per_filtered_cell must divide nz");
        // the reference reads u.ρq_liq / u.ρq_rai / u.N_*: a style without that field is an error there too
        const bool needs_liq = v == KID_OBS_QL || v == KID_OBS_QV || v == KID_OBS_QLR || v == KID_OBS_RL || v == KID_OBS_RV || v == KID_OBS_RLR;
        const bool needs_rai = v == KID_OBS_QR || v == KID_OBS_QLR || v == KID_OBS_RR || v == KID_OBS_RLR || v == KID_OBS_RAIN_VT || v == KID_OBS_RAINRATE;
        const bool needs_N = v == KID_OBS_NL || v == KID_OBS_NR || v == KID_OBS_NA;
        if (needs_liq && !neq) return fail("observable needs ρq_liq: NonEquilibriumMoisture only");
        if (needs_rai && !has_pr) return fail("observable needs ρq_rai: Precipitation1M / Precipitation2M only");
        if (needs_N && !is2m) return fail("observable needs N_liq/N_rai/N_aer: Precipitation2M only");
        if (v == KID_OBS_Z_TOP && !(neq && h->cfg.precip == KID_PRECIP_1M))
            return fail("Computing radar reflectivity for the given precipitation style is invalid!! (Z_top: NonEquilibriumMoisture + Precipitation1M)");
        h->obs_var[j] = v;
        h->obs_per[j] = v == KID_OBS_Z_TOP ? nz : per;  // Z_top: one value per column
        h->obs_off[j] = off;
        off += v == KID_OBS_Z_TOP ? 1 : nz / per;
    }
    h->obs_off[nvars] = off;
    h->obs_nvars = nvars;
    h->obs_nt_filtered = nt_filtered;
    h->obs_nt_per = nt_per_filtered_cell;
    h->obs_n_out = int64_t(off);
    if (h->obsG) cudaFree(h->obsG);
    h->obsG = nullptr;
    const size_t bytes = size_t(h->cfg.nc) * nt_filtered * off * sizeof(double);
    KID_CUDA(cudaMalloc(&h->obsG, bytes));
    KID_CUDA(cudaMemsetAsync(h->obsG, 0, bytes, h->stream));
    if (n_out) *n_out = int64_t(nt_filtered) * off;
    return 0;
}

int kid_obs_accumulate(kid_handle_t* h, int32_t time_cell) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (!h->obsG) return fail("kid_obs_accumulate: call kid_obs_configure first");
    if (time_cell < 0 || time_cell >= h->obs_nt_filtered) return fail("kid_obs_accumulate: time cell out of range");
    return h->f64 ? obs_launch<double>(h, time_cell) : obs_launch<float>(h, time_cell);
}

int kid_obs_get(kid_handle_t* h, double* G) {
    KID_ON_DEVICE(h);
    if (!h || !G) return fail("null argument");
    if (!h->obsG) return fail("kid_obs_get: call kid_obs_configure first");
    const size_t bytes = size_t(h->cfg.nc) * h->obs_nt_filtered * h->obs_n_out * sizeof(double);
    KID_CUDA(cudaMemcpyAsync(G, h->obsG, bytes, cudaMemcpyDeviceToHost, h->stream));
    KID_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int kid_k2d_stage_begin(kid_handle_t* h, double t0, int32_t stage) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (h->cfg.model != KID_MODEL_K2D) return fail("not a K2D handle");
    if (stage < 0 || stage > 2) return fail("stage must be 0, 1 or 2");
    double ts = h->f64 ? t0 : double(float(t0));
    return k2d_begin(h, k2d_stage_time(h, ts, stage), stage, true, false);
}
int kid_k2d_edge_buffers(kid_handle_t* h, void** send_left, void** send_right, void** recv_left, void** recv_right,
                         size_t* edge_count) {
    KID_ON_DEVICE(h);
    if (!h || h->cfg.model != KID_MODEL_K2D) return fail("not a K2D handle");
    if (send_left) *send_left = h->send_left;
    if (send_right) *send_right = h->send_right;
    if (recv_left) *recv_left = h->recv_left;
    if (recv_right) *recv_right = h->recv_right;
    if (edge_count) *edge_count = size_t(h->nfields) * h->cfg.nz;
    return 0;
}
int kid_k2d_stage_end(kid_handle_t* h, int32_t stage) {
    KID_ON_DEVICE(h);
    if (check_ready(h)) return 1;
    if (h->cfg.model != KID_MODEL_K2D) return fail("not a K2D handle");
    if (stage < 0 || stage > 2) return fail("stage must be 0, 1 or 2");
    return k2d_end(h, stage, 1, 0);
}

int kid_nccl_unique_id(void* id128) {
    if (!id128) return fail("null argument");
    const NcclApi& n = nccl_api();
    if (!n.ok) return fail("libnccl.so.2 could not be loaded: the in-library K2D exchange needs NCCL");
    int rc = n.GetUniqueId(id128);
    if (rc != 0) return fail(std::string("ncclGetUniqueId: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"));
    return 0;
}

int kid_k2d_comm_init(kid_handle_t* h, int32_t rank, int32_t world, const void* id128) {
    if (!h || !id128) return fail("null argument");
    KID_ON_DEVICE(h);
    if (h->cfg.model != KID_MODEL_K2D) return fail("not a K2D handle");
    if (world < 1 || rank < 0 || rank >= world) return fail("kid_k2d_comm_init: invalid rank / world");
    const NcclApi& n = nccl_api();
    if (!n.ok) return fail("libnccl.so.2 could not be loaded: the in-library K2D exchange needs NCCL");
    if (h->comm) { n.CommDestroy(h->comm); h->comm = nullptr; }
    NcclUniqueIdBytes id;
    std::memcpy(&id, id128, sizeof(id));
    int rc = n.CommInitRank(&h->comm, world, id, rank);
    if (rc != 0) { h->comm = nullptr; return fail(std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error")); }
    h->comm_rank = rank;
    h->comm_world = world;
    return 0;
}

// periodic ring exchange of the packed rank-edge columns on the handle's stream (after the pack of k2d_begin, before the
// DSS of k2d_end): recv_left <- left neighbour's send_right, recv_right <- right neighbour's send_left
static int k2d_exchange_nccl(kid_handle_t* h) {
    const NcclApi& n = nccl_api();
    const size_t count = size_t(h->nfields) * h->cfg.nz;
    const int dtype = h->f64 ? 8 : 7;  // ncclFloat64 : ncclFloat32
    if (h->comm_world == 1) {
        KID_CUDA(cudaMemcpyAsync(h->recv_left, h->send_right, count * h->fsz, cudaMemcpyDeviceToDevice, h->stream));
        KID_CUDA(cudaMemcpyAsync(h->recv_right, h->send_left, count * h->fsz, cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    }
    const int left = (h->comm_rank + h->comm_world - 1) % h->comm_world, right = (h->comm_rank + 1) % h->comm_world;
    int rc = n.GroupStart();
    // world == 2: both neighbours are the same peer; NCCL matches sends and receives to a peer in issue order, so the
    // receive order mirrors the peer's send order (its send_left is my recv_right)
    if (!rc) rc = n.Send(h->send_left, count, dtype, left, h->comm, h->stream);
    if (!rc) rc = n.Send(h->send_right, count, dtype, right, h->comm, h->stream);
    if (!rc) rc = n.Recv(h->recv_right, count, dtype, right, h->comm, h->stream);
    if (!rc) rc = n.Recv(h->recv_left, count, dtype, left, h->comm, h->stream);
    int rc2 = n.GroupEnd();
    if (rc || rc2) return fail(std::string("NCCL edge exchange: ") + (n.GetErrorString ? n.GetErrorString(rc ? rc : rc2) : "error"));
    return 0;
}

int kid_k2d_exchange_local(kid_handle_t* left, kid_handle_t* right) {
    if (!left || !right) return fail("null handle");
    if (left->cfg.model != KID_MODEL_K2D || right->cfg.model != KID_MODEL_K2D) return fail("not a K2D handle");
    if (left->nfields != right->nfields || left->cfg.nz != right->cfg.nz || left->fsz != right->fsz)
        return fail("kid_k2d_exchange_local: handles of different shape");
    size_t bytes = size_t(left->nfields) * left->cfg.nz * left->fsz;
    // each copy runs on the RECEIVER's stream after the sender's pack kernel has finished; an event belongs to the device
    // of the stream it is recorded on
    cudaEvent_t el = nullptr, er = nullptr;
    {
        DeviceGuard g(left->cfg.device);
        KID_CUDA(cudaEventCreateWithFlags(&el, cudaEventDisableTiming));
        KID_CUDA(cudaEventRecord(el, left->stream));
    }
    {
        DeviceGuard g(right->cfg.device);
        KID_CUDA(cudaEventCreateWithFlags(&er, cudaEventDisableTiming));
        KID_CUDA(cudaEventRecord(er, right->stream));
        KID_CUDA(cudaStreamWaitEvent(right->stream, el, 0));
        KID_CUDA(cudaMemcpyPeerAsync(right->recv_left, right->cfg.device, left->send_right, left->cfg.device, bytes, right->stream));
    }
    {
        DeviceGuard g(left->cfg.device);
        KID_CUDA(cudaStreamWaitEvent(left->stream, er, 0));
        KID_CUDA(cudaMemcpyPeerAsync(left->recv_right, left->cfg.device, right->send_left, right->cfg.device, bytes, left->stream));
        KID_CUDA(cudaEventDestroy(el));
    }
    {
        DeviceGuard g(right->cfg.device);
        KID_CUDA(cudaEventDestroy(er));
    }
    return 0;
}

int kid_synchronize(kid_handle_t* h) {
    KID_ON_DEVICE(h);
    if (!h) return fail("null handle");
    KID_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
void* kid_stream(kid_handle_t* h) {
    KID_ON_DEVICE(h); return h ? (void*)h->stream : nullptr; }
void* kid_state_device_ptr(kid_handle_t* h) {
    KID_ON_DEVICE(h); return h ? h->Y : nullptr; }
int64_t kid_launch_count(kid_handle_t* h) {
    KID_ON_DEVICE(h); return h ? h->launches : 0; }
const char* kid_last_error(void) { return g_err.c_str(); }
const char* kid_version(void) { return "libkid_cuda 0.1 (sm_100a)"; }

}  // extern "C"
