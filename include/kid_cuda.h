/*
 * kid_cuda.h — C ABI of libkid_cuda.so, the B200 (sm_100a) implementation of
 * KinematicDriver.jl's hot path: the ODE right-hand side + SSPRK33 step loop of
 * K1DModel, K2DModel and BoxModel.
 *
 * The reference has no FFI seam: its drivers hand `rhs!(dY, Y, aux, t)` to
 * OrdinaryDiffEq (`ODE.solve(problem, ODE.SSPRK33(); dt, saveat, callback)`).
 * The entry points below are what a Julia `ccall` shim binds in place of that
 * hand-off.  Each comment cites the reference interface it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; the message is
 *    available from kid_last_error() (thread-local).
 *  - `void*` host buffers hold FLOAT_TYPE elements (float if float_type ==
 *    KID_F32, double if KID_F64); they are borrowed for the duration of the
 *    call only.  No torch / CUDA types appear in any signature.
 *  - host state layout is `[column][variable][level]`, level fastest: exactly
 *    `parent(Y)` of the reference's ClimaCore column FieldVector, column after
 *    column (src/Common/ode_utils.jl:23-54 gives the variable order).
 *  - a handle is not thread-safe; distinct handles may be used concurrently
 *    (the reference's calibration loop runs one simulation per thread:
 *    calibration/OptimizationUtils.jl:206).
 *  - there is NO CPU fallback: every compute entry point fails if no CUDA
 *    device is present.
 */
#ifndef KID_CUDA_H
#define KID_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "kid_params.h"

#ifdef __cplusplus
extern "C" {
#endif

#define KID_ABI_VERSION 1

/* src/{BoxModel,K1DModel,K2DModel}/ode_utils.jl: which make_rhs_function */
enum kid_model {
    KID_MODEL_BOX = 0,         /* src/BoxModel/ode_utils.jl:10-22 */
    KID_MODEL_K1D = 1,         /* src/K1DModel/ode_utils.jl:29-49 */
    KID_MODEL_K1D_COL_SED = 2, /* src/K1DModel/ode_utils.jl:56-73 */
    KID_MODEL_K2D = 3          /* src/K2DModel/ode_utils.jl:36-54 */
};
/* src/Common/equation_types.jl:31-61 / helper_functions.jl:4-93 */
enum kid_moisture { KID_MOIST_EQ = 0, KID_MOIST_NONEQ = 1 };
enum kid_precip { KID_PRECIP_NONE = 0, KID_PRECIP_0M = 1, KID_PRECIP_1M = 2, KID_PRECIP_2M = 3 };
enum kid_rain_formation {
    KID_RF_CLIMA_1M = 0, KID_RF_KK2000 = 1, KID_RF_B1994 = 2, KID_RF_TC1980 = 3,
    KID_RF_LD2004 = 4, KID_RF_VARTIMESCALE = 5,       /* Precipitation1M */
    KID_RF_SB2006 = 6, KID_RF_SB2006NL = 7            /* Precipitation2M */
};
enum kid_sedimentation { KID_SED_CLIMA_1M = 0, KID_SED_SB2006 = 1, KID_SED_CHEN2022 = 2 };
enum kid_float { KID_F32 = 0, KID_F64 = 1 }; /* FLOAT_TYPE (parse_commandline.jl:10-14) */

/* per-column scalars that the reference keeps in kid_params / common_params /
 * aux and that differ between ensemble members (calibration/KiDUtils.jl:45-81) */
enum kid_column_scalar {
    KID_COL_W1 = 0,            /* kid_params.w1   (K1DModel/tendency.jl:243) */
    KID_COL_Q_SURF = 1,        /* aux.q_surf      (K1DModel/ode_utils.jl:96) */
    KID_COL_PRESCRIBED_ND = 2  /* common_params.prescribed_Nd */
};

/* aux fields readable by kid_get_aux: everything simulation_output() reads
 * (src/Common/netcdf_io.jl:144-170; schema src/Common/ode_utils.jl:110-175) */
enum kid_aux_field {
    KID_AUX_RHO = 0, KID_AUX_RHO_DRY, KID_AUX_P, KID_AUX_T, KID_AUX_THETA_LIQ_ICE, KID_AUX_THETA_DRY,
    KID_AUX_Q_TOT, KID_AUX_Q_LIQ, KID_AUX_Q_ICE, KID_AUX_Q_RAI, KID_AUX_Q_SNO,
    KID_AUX_N_LIQ, KID_AUX_N_RAI, KID_AUX_N_AER,
    KID_AUX_TERM_VEL_RAI, KID_AUX_TERM_VEL_SNO, KID_AUX_TERM_VEL_N_RAI,
    KID_AUX_SRC_Q_TOT, KID_AUX_SRC_Q_LIQ, KID_AUX_SRC_Q_ICE, KID_AUX_SRC_Q_RAI, KID_AUX_SRC_Q_SNO,
    KID_AUX_SRC_N_AER, KID_AUX_SRC_N_LIQ, KID_AUX_SRC_N_RAI,
    KID_AUX_ACT_N_AER, KID_AUX_ACT_N_LIQ,
    KID_AUX_CLOUD_SRC_Q_LIQ, KID_AUX_CLOUD_SRC_Q_ICE,
    KID_NAUX
};

/* observables of CO.get_variable_data_from_ODE (src/Common/helper_functions.jl:98-191), in the order of its
 * if-chain: "qt","ql","qr","qv","qlr","rt","rl","rr","rv","rlr","Nl","Nr","Na","rho",
 * "rain averaged terminal velocity","rainrate".  "Z_top" needs CloudMicrophysics' radar reflectivity: not available. */
enum kid_obs_var {
    KID_OBS_QT = 0, KID_OBS_QL, KID_OBS_QR, KID_OBS_QV, KID_OBS_QLR,
    KID_OBS_RT, KID_OBS_RL, KID_OBS_RR, KID_OBS_RV, KID_OBS_RLR,
    KID_OBS_NL, KID_OBS_NR, KID_OBS_NA, KID_OBS_RHO, KID_OBS_RAIN_VT, KID_OBS_RAINRATE,
    KID_OBS_Z_TOP, /* radar reflectivity of the rain below cloud top: ONE value per column (Precipitation1M only) */
    KID_NOBS
};
#define KID_MAX_OBS_VARS 16

typedef struct kid_config {
    int32_t abi_version;     /* = KID_ABI_VERSION */
    int32_t params_version;  /* = KID_PARAMS_VERSION */
    int32_t model;           /* enum kid_model */
    int32_t moisture;        /* enum kid_moisture */
    int32_t precip;          /* enum kid_precip */
    int32_t rain_formation;  /* enum kid_rain_formation */
    int32_t sedimentation;   /* enum kid_sedimentation */
    int32_t float_type;      /* enum kid_float */
    int32_t nz;              /* levels per column (n_elem; 1 for BoxModel) */
    int32_t nc;              /* columns owned by this handle (K2D: helem_local*(npoly+1)) */
    int32_t helem;           /* K2D: horizontal elements owned by this handle */
    int32_t npoly;           /* K2D: polynomial degree (Nq = npoly+1 GLL nodes) */
    int32_t helem_global;    /* K2D: elements of the whole periodic domain */
    int32_t helem_offset;    /* K2D: global index of this handle's first element */
    /* CommonParameters / kid_params switches (src/Common/parameters.jl:18-41) */
    int32_t precip_sources;
    int32_t precip_sinks;
    int32_t open_system_activation;
    int32_t local_activation;
    int32_t qtot_flux_correction;
    int32_t device;          /* CUDA device ordinal */
    int32_t reserved[8];     /* reserved[0..5]: tuning overrides (0 = library default): max level-warps per tile (>= 2), columns per tile, 1 = never use the compile-time-shape kernel, 1 = never use the packed two-levels-per-thread 2M kernel, 2 = layout kernels access page-locked host buffers in place even without a column map, 1 = K2D steps with the two legacy horizontal launches per stage */
    double z_min, z_max;     /* make_function_space (K1DModel/ode_utils.jl:9-21) */
    double dt;               /* TimeStepping.dt (src/Common/ode_utils.jl:12-16) */
    double domain_width;     /* K2D aux.domain_width  */
    double domain_height;    /* K2D aux.domain_height */
} kid_config_t;

typedef struct kid_handle kid_handle_t;

/* number of prognostic variables of a (moisture, precip) pair and their order:
 * initialise_state, src/Common/ode_utils.jl:23-54.  Returns -1 if unsupported. */
int kid_nvars(int moisture, int precip);
/* writes the NUL-separated variable names ("rhoq_tot\0rhoq_liq\0...") */
int kid_var_names(int moisture, int precip, char* buf, size_t buflen);

/* replaces: aux = initialise_aux(...) + Y = initialise_state(...) allocation
 * (src/Common/ode_utils.jl:61-175, K1DModel/ode_utils.jl:80-117) */
int kid_create(const kid_config_t* cfg, kid_handle_t** out);
void kid_destroy(kid_handle_t* h);

/* replaces: the parameter structs captured in `aux` and in the style types.
 * `params` is nsets x KID_NPARAMS doubles.  nsets == 1: one set for every
 * column.  nsets > 1: column c uses set column_to_set[c] (calibration
 * ensembles: calibration/KiDUtils.jl:378-382 update_parameters!). */
int kid_set_params(kid_handle_t* h, const double* params, int32_t nparams, int32_t nsets,
                   const int32_t* column_to_set);

/* replaces: aux.thermo_variables.{ρ_dry,T} from the initial profile (time
 * invariant: src/Common/tendency.jl:34-135).  per_column == 0: one profile of
 * nz values shared by all columns; else nc x nz values, level fastest. */
int kid_set_profiles(kid_handle_t* h, const void* rho_dry, const void* T, int32_t per_column);

/* per-column scalars; values == NULL resets to the value in the param table */
int kid_set_column_scalar(kid_handle_t* h, int32_t which, const void* values);

/* replaces: Y = initialise_state(...) / reading integrator.u.
 * kid_set_state also snapshots the aux microphysics fields that the reference
 * initialises from the initial profile and never refreshes on some paths
 * (aux N_aer in BoxModel / col_sed; previous-call q_rai,N_liq,N_rai in K2D). */
int kid_set_state(kid_handle_t* h, const void* Y);
int kid_get_state(kid_handle_t* h, void* Y);
/* asynchronous variants for PINNED host buffers: the copy is only enqueued on the handle's stream, so uploads and
 * downloads of one handle overlap the kernels of other handles (column chunks of a large ensemble pipelined through
 * several handles).  The buffer must stay valid and untouched until kid_synchronize(h) returns. */
int kid_set_state_async(kid_handle_t* h, const void* Y_pinned);
int kid_get_state_async(kid_handle_t* h, void* Y_pinned);
/* Column map (optional): device column i holds HOST column host_column_of[i] (nc entries; NULL = identity).  The order of
 * ensemble members is arbitrary for the caller (OptimizationUtils.jl:197-211 loops over members), but it decides which
 * members share a warp: with a map the caller keeps its own order in the host arrays of kid_set_state* / kid_get_state*
 * while the device stores similar members next to each other (kid_b200.model.warp_coherent_order).  Host entries may point
 * anywhere into a larger page-locked array (column chunks of a pipelined ensemble).  Profiles, per-column scalars and aux
 * output stay in device column order.  A map that permutes ONE contiguous block of host rows (a chunk of the caller's array
 * stored in its own warp-coherent order) moves the block with the copy engine and permutes on the device — as fast as an
 * unmapped transfer.  A scattered map needs page-locked host buffers (kid_host_alloc, cudaHostAlloc, cudaHostRegister): the
 * layout kernels then gather / scatter the host rows in place over PCIe (slower; cfg.reserved[4] = 2 forces in-place access). */
int kid_set_column_map(kid_handle_t* h, const int32_t* host_column_of);

/* replaces: ρ_ivp + initial_condition_1d + initialise_state and the ρ_dry, T, q_surf fields of initialise_aux
 * (src/Common/initial_condition.jl:50-231, src/Common/ode_utils.jl:23-54, src/K1DModel/ode_utils.jl:80-117) for the
 * ShipwayHill2012 sounding: the hydrostatic profile and the initial state of every column are built ON the device from
 * sounding = {z_0, z_1, z_2, rv_0, rv_1, rv_2, tht_0, tht_1, tht_2} and one surface pressure per column
 * (p0: nc FLOAT_TYPE values, NULL = the drivers' default 100000 Pa), so each ensemble member can have its own p0 without a host ODE
 * solve.  Needs kid_set_params (and kid_set_column_scalar(PRESCRIBED_ND) if members differ in Nd) first; afterwards
 * the handle is ready to step (profiles, q_surf and state are set).  K1DModel handles only. */
int kid_init_shipway_hill(kid_handle_t* h, const double* sounding, const void* p0, int32_t dry);
/* replaces: initial_condition_0d + initialise_state for BoxModel and the col_sed column (src/Common/initial_condition.jl:
 * 237-308; run_box_simulation.jl:60-76, run_KiD_col_sed_simulation.jl:80-103): the gamma-distribution state (liquid / rain
 * split at a 40 um radius) of every column from its total water qt, number Nd, shape k and dry density, built on the device
 * (T = 300 K, every level alike) — calibration ensembles with per-case (qt, Nd, k) need no host loop over members.
 * Each array: nc values of FLOAT_TYPE. */
int kid_init_gamma_0d(kid_handle_t* h, const void* qt, const void* Nd, const void* k, const void* rho_dry);

/* download the time-invariant profiles (the layout kid_set_profiles takes: nc x nz, or nz if one profile is shared) */
int kid_get_profiles(kid_handle_t* h, void* rho_dry, void* T);

/* replaces: ODE.solve!(integrator) between two save/callback times: advances
 * the device-resident state by nsteps SSPRK33 steps of size dt starting at
 * time t0 (3 rhs evaluations per step).  Asynchronous w.r.t. the host. */
int kid_step(kid_handle_t* h, double t0, int32_t nsteps);
/* how many model steps one kernel launch fuses (0 = library default) */
int kid_set_steps_per_launch(kid_handle_t* h, int32_t n);

/* replaces: one rhs!(dY, Y, aux, t) call (testing / parity of a single
 * evaluation).  dY_out has the layout of kid_get_state. */
int kid_rhs(kid_handle_t* h, double t, void* dY_out);

/* replaces: reading aux.* in simulation_output (netcdf_io.jl:144-170); the
 * field is evaluated from the current state at time t (what the FSAL rhs call
 * leaves in aux).  out: nc x nz values, level fastest. */
int kid_get_aux(kid_handle_t* h, double t, int32_t field, void* out);
/* several fields from ONE rhs evaluation.  out: nc x nfields x nz values ([column][field][level]). */
int kid_get_aux_many(kid_handle_t* h, double t, int32_t nfields, const int32_t* fields, void* out);
/* replaces: the DiscreteCallback(condition_io, affect_io!) output path (netcdf_io.jl:176-194) at its default cadence
 * (every 2nd step): the fields are evaluated on the compute stream and downloaded by a second stream from
 * double-buffered staging, so the following kid_step launches overlap the PCIe transfer.  out_pinned (page-locked,
 * layout of kid_get_aux_many) is valid after kid_output_wait; at most two outputs are in flight. */
int kid_output_enqueue(kid_handle_t* h, double t, int32_t nfields, const int32_t* fields, void* out_pinned);
int kid_output_wait(kid_handle_t* h);
/* page-locked host memory for the *_async / kid_output_enqueue calls (cudaHostAlloc / cudaFreeHost) */
void* kid_host_alloc(size_t bytes);
void kid_host_free(void* p);
/* replaces: maximum(aux.microph_variables.q_liq) (netcdf_io.jl:166) */
int kid_reduce_max(kid_handle_t* h, double t, int32_t field, double* out);

/* replaces: ODEsolution2Gvector (calibration/KiDUtils.jl:329-376) fed by get_variable_data_from_ODE: the calibration
 * G vector is accumulated on the device, one row per column (ensemble member x case), so only the filtered
 * observables ever cross PCIe.
 *   kid_obs_configure: nvars observables (enum kid_obs_var), variable j averaged over blocks of
 *       nz_per_filtered_cell[j] levels (must divide nz; 1 = unfiltered) and over nt_per_filtered_cell snapshots per
 *       time cell; nt_filtered time cells.  Row length n_out = nt_filtered * sum_j nz / nz_per_filtered_cell[j],
 *       ordered [time cell][variable][filtered level] like the reference's outputs vector.  Zeroes the accumulator.
 *   kid_obs_accumulate: add the observables of the CURRENT device state to time cell `time_cell`
 *       (call it at every saveat time of make_filter_props: nt_per_filtered_cell times per cell).
 *   kid_obs_get: download the accumulator, nc rows of n_out Float64 values (the reference's outputs are Float64).
 * Without filtering (apply = false): nz_per = 1, nt_per = 1 and one time cell per saved time. */
int kid_obs_configure(kid_handle_t* h, int32_t nvars, const int32_t* var_ids, const int32_t* nz_per_filtered_cell,
                      int32_t nt_filtered, int32_t nt_per_filtered_cell, int64_t* n_out);
int kid_obs_accumulate(kid_handle_t* h, int32_t time_cell);
int kid_obs_get(kid_handle_t* h, double* G);

/* K2D domain decomposition: the caller (one process per GPU) exchanges the
 * packed edge-node tendencies of its first/last element with its periodic
 * neighbours between the two halves of every SSPRK33 stage.
 *   kid_k2d_stage_begin: rhs for stage s (0,1,2) of the step starting at t0,
 *       leaves the un-DSS'd tendencies on the device and packs the edges.
 *   kid_k2d_edge_buffers: device pointers of the send (left,right) and recv
 *       (left,right) buffers, each edge_count FLOAT_TYPE elements.
 *   kid_k2d_stage_end: DSS average with the received edges + stage update.
 * kid_step() on a K2D handle with helem == helem_global does all of this
 * internally (single GPU, periodic wrap inside the kernel). */
int kid_k2d_stage_begin(kid_handle_t* h, double t0, int32_t stage);
int kid_k2d_edge_buffers(kid_handle_t* h, void** send_left, void** send_right, void** recv_left,
                         void** recv_right, size_t* edge_count);
int kid_k2d_stage_end(kid_handle_t* h, int32_t stage);
/* neighbour exchange between two handles driven by the SAME process (one process steering several
 * sub-domains / GPUs): right-edge of `left` -> recv_left of `right`, left-edge of `right` -> recv_right of
 * `left`, ordered on both handles' streams (peer copies over NVLink when the handles sit on different GPUs).
 * With separate processes the caller exchanges the buffers of kid_k2d_edge_buffers itself (NCCL send/recv). */
int kid_k2d_exchange_local(kid_handle_t* left, kid_handle_t* right);
/* In-library exchange for one process per GPU (replaces the cross-rank part of CC.Spaces.weighted_dss!,
 * src/K2DModel/tendency.jl:70,106-108,156-157,206-207, for a host that has no device-side communication of its own):
 * rank 0 obtains a 128-byte id with kid_nccl_unique_id and distributes it by any means (Julia: Distributed / MPI.jl; the
 * Python mirror: torch.distributed.broadcast_object_list); every rank of the periodic x-ring then calls kid_k2d_comm_init.
 * Afterwards kid_step on the decomposed handle runs all SSPRK33 stages itself, with one grouped ncclSend/ncclRecv of the packed
 * edge columns per stage on the handle's stream.  NCCL is bound at run time (dlopen of libnccl.so.2). */
int kid_nccl_unique_id(void* id128);
int kid_k2d_comm_init(kid_handle_t* h, int32_t rank, int32_t world, const void* id128);

/* plumbing */
int kid_synchronize(kid_handle_t* h);
void* kid_stream(kid_handle_t* h);            /* the cudaStream_t kernels are launched on */
void* kid_state_device_ptr(kid_handle_t* h);  /* device SoA state, [var][level][column] */
int64_t kid_launch_count(kid_handle_t* h);    /* kernels launched by this handle so far */
const char* kid_last_error(void);
const char* kid_version(void);
int kid_device_count(void);
/* Measurement support for the roofline of bench.py ("measure with a FMA micro-benchmark on the box"): runs a
 * register-resident FFMA loop on `device` for a few milliseconds and returns the achieved Float32 rate in flop/s (2 per
 * FFMA), with the SM count and the SM clock (kHz) the driver reports.  No handle needed. */
int kid_measure_fp32_peak(int32_t device, double* flops_per_s, int32_t* sm_count, int32_t* sm_clock_khz);

#ifdef __cplusplus
}
#endif
#endif /* KID_CUDA_H */
