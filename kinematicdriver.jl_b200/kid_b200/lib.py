"""ctypes binding of libkid_cuda.so (include/kid_cuda.h) — the only way the host side reaches the GPU.

There is deliberately no fallback: if the shared library is missing or no CUDA device is
visible, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

KID_ABI_VERSION = 1
KID_PARAMS_VERSION = 2

# enum kid_model
MODEL_BOX, MODEL_K1D, MODEL_K1D_COL_SED, MODEL_K2D = 0, 1, 2, 3
# enum kid_float
F32, F64 = 0, 1
# enum kid_column_scalar
COL_W1, COL_Q_SURF, COL_PRESCRIBED_ND = 0, 1, 2

AUX_FIELDS = [
    "ρ", "ρ_dry", "p", "T", "θ_liq_ice", "θ_dry",
    "q_tot", "q_liq", "q_ice", "q_rai", "q_sno", "N_liq", "N_rai", "N_aer",
    "term_vel_rai", "term_vel_sno", "term_vel_N_rai",
    "precip_sources.q_tot", "precip_sources.q_liq", "precip_sources.q_ice", "precip_sources.q_rai",
    "precip_sources.q_sno", "precip_sources.N_aer", "precip_sources.N_liq", "precip_sources.N_rai",
    "activation_sources.N_aer", "activation_sources.N_liq",
    "cloud_sources.q_liq", "cloud_sources.q_ice",
]
AUX_ID = {n: i for i, n in enumerate(AUX_FIELDS)}
# enum kid_obs_var: the names get_variable_data_from_ODE accepts (src/Common/helper_functions.jl:98-191)
OBS_VARS = ["qt", "ql", "qr", "qv", "qlr", "rt", "rl", "rr", "rv", "rlr", "Nl", "Nr", "Na", "rho",
            "rain averaged terminal velocity", "rainrate", "Z_top"]
OBS_ID = {n: i for i, n in enumerate(OBS_VARS)}


class KidConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("params_version", C.c_int32),
        ("model", C.c_int32), ("moisture", C.c_int32), ("precip", C.c_int32),
        ("rain_formation", C.c_int32), ("sedimentation", C.c_int32), ("float_type", C.c_int32),
        ("nz", C.c_int32), ("nc", C.c_int32), ("helem", C.c_int32), ("npoly", C.c_int32),
        ("helem_global", C.c_int32), ("helem_offset", C.c_int32),
        ("precip_sources", C.c_int32), ("precip_sinks", C.c_int32),
        ("open_system_activation", C.c_int32), ("local_activation", C.c_int32),
        ("qtot_flux_correction", C.c_int32), ("device", C.c_int32),
        ("reserved", C.c_int32 * 8),
        ("z_min", C.c_double), ("z_max", C.c_double), ("dt", C.c_double),
        ("domain_width", C.c_double), ("domain_height", C.c_double),
    ]


def make_config(**kw) -> KidConfig:
    cfg = KidConfig()
    cfg.abi_version, cfg.params_version = KID_ABI_VERSION, KID_PARAMS_VERSION
    cfg.helem = cfg.helem_global = 0
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def dtype_of(float_type: int):
    return np.float32 if float_type == F32 else np.float64


_LIB = None
LIB_PATH = Path(os.environ.get("KID_CUDA_LIB") or Path(__file__).resolve().parents[1] / "csrc" / "libkid_cuda.so")


class KidCudaError(RuntimeError):
    pass


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen libkid_cuda.so and declare every symbol of include/kid_cuda.h."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise KidCudaError(f"{p} not found: build it with `python __graft_entry__.py build` "
                           "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(str(p))
    H = C.c_void_p
    sig = {
        "kid_nvars": (C.c_int, [C.c_int, C.c_int]),
        "kid_var_names": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
        "kid_create": (C.c_int, [C.POINTER(KidConfig), C.POINTER(H)]),
        "kid_destroy": (None, [H]),
        "kid_set_params": (C.c_int, [H, C.POINTER(C.c_double), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
        "kid_set_profiles": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_int32]),
        "kid_set_column_scalar": (C.c_int, [H, C.c_int32, C.c_void_p]),
        "kid_set_state": (C.c_int, [H, C.c_void_p]),
        "kid_get_state": (C.c_int, [H, C.c_void_p]),
        "kid_set_state_async": (C.c_int, [H, C.c_void_p]),
        "kid_set_column_map": (C.c_int, [H, C.POINTER(C.c_int32)]),
        "kid_get_state_async": (C.c_int, [H, C.c_void_p]),
        "kid_init_shipway_hill": (C.c_int, [H, C.POINTER(C.c_double), C.c_void_p, C.c_int32]),
        "kid_get_profiles": (C.c_int, [H, C.c_void_p, C.c_void_p]),
        "kid_init_gamma_0d": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "kid_step": (C.c_int, [H, C.c_double, C.c_int32]),
        "kid_set_steps_per_launch": (C.c_int, [H, C.c_int32]),
        "kid_rhs": (C.c_int, [H, C.c_double, C.c_void_p]),
        "kid_get_aux": (C.c_int, [H, C.c_double, C.c_int32, C.c_void_p]),
        "kid_reduce_max": (C.c_int, [H, C.c_double, C.c_int32, C.POINTER(C.c_double)]),
        "kid_get_aux_many": (C.c_int, [H, C.c_double, C.c_int32, C.POINTER(C.c_int32), C.c_void_p]),
        "kid_output_enqueue": (C.c_int, [H, C.c_double, C.c_int32, C.POINTER(C.c_int32), C.c_void_p]),
        "kid_output_wait": (C.c_int, [H]),
        "kid_host_alloc": (C.c_void_p, [C.c_size_t]),
        "kid_host_free": (None, [C.c_void_p]),
        "kid_obs_configure": (C.c_int, [H, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                        C.POINTER(C.c_int64)]),
        "kid_obs_accumulate": (C.c_int, [H, C.c_int32]),
        "kid_obs_get": (C.c_int, [H, C.POINTER(C.c_double)]),
        "kid_k2d_stage_begin": (C.c_int, [H, C.c_double, C.c_int32]),
        "kid_k2d_edge_buffers": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        "kid_k2d_stage_end": (C.c_int, [H, C.c_int32]),
        "kid_k2d_exchange_local": (C.c_int, [H, H]),
        "kid_nccl_unique_id": (C.c_int, [C.c_void_p]),
        "kid_k2d_comm_init": (C.c_int, [H, C.c_int32, C.c_int32, C.c_void_p]),
        "kid_synchronize": (C.c_int, [H]),
        "kid_stream": (C.c_void_p, [H]),
        "kid_state_device_ptr": (C.c_void_p, [H]),
        "kid_launch_count": (C.c_int64, [H]),
        "kid_last_error": (C.c_char_p, []),
        "kid_version": (C.c_char_p, []),
        "kid_device_count": (C.c_int, []),
        "kid_measure_fp32_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    lib._kid_symbols = tuple(sig)
    if path is None:
        _LIB = lib
    return lib


class _PinnedOwner:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.kid_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype, lib: C.CDLL | None = None) -> np.ndarray:
    """numpy array in page-locked host memory (kid_host_alloc), for the asynchronous transfers."""
    lib = lib or load_library()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.kid_host_alloc(max(nbytes, 1))
    if not ptr:
        raise KidCudaError(lib.kid_last_error().decode())
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr)
    buf._owner = _PinnedOwner(lib, ptr)  # freed when the last array view dies
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def bind_to_gpu_numa_node(pci_bus_id: str) -> dict:
    """One process per GPU: run this process (and first-touch its page-locked staging buffers) on the CPUs of the NUMA node
    the GPU hangs off, so that host<->device copies of several ranks do not cross the socket interconnect.
    Returns what was done ({} if the topology is not exposed in sysfs)."""
    try:
        dom_bus = pci_bus_id.lower()
        if len(dom_bus.split(":")[0]) == 8:  # CUDA prints 8 hex digits for the PCI domain, sysfs uses 4
            dom_bus = dom_bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{dom_bus}/numa_node").read_text())
        if node < 0:
            return {}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except (OSError, ValueError):
        return {}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def nccl_unique_id(lib: C.CDLL | None = None) -> bytes:
    """128-byte ncclUniqueId (kid_nccl_unique_id): rank 0 creates it, every rank passes it to Handle.k2d_comm_init."""
    lib = lib or load_library()
    buf = C.create_string_buffer(128)
    if lib.kid_nccl_unique_id(buf) != 0:
        raise KidCudaError(lib.kid_last_error().decode())
    return buf.raw


def measure_fp32_peak(device: int = 0, lib: C.CDLL | None = None) -> dict:
    """FFMA micro-benchmark on the device (kid_measure_fp32_peak): the Float32 roofline denominator, measured in the run."""
    lib = lib or load_library()
    f, sms, khz = C.c_double(), C.c_int32(), C.c_int32()
    if lib.kid_measure_fp32_peak(int(device), C.byref(f), C.byref(sms), C.byref(khz)) != 0:
        raise KidCudaError(lib.kid_last_error().decode())
    return {"flops_per_s": f.value, "sm_count": sms.value, "sm_clock_khz": khz.value}


class Handle:
    """Thin OO wrapper of a kid_handle_t*; array arguments are numpy arrays of FLOAT_TYPE."""

    def __init__(self, cfg: KidConfig, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.cfg = cfg
        self.dtype = dtype_of(cfg.float_type)
        self.nvars = self.lib.kid_nvars(cfg.moisture, cfg.precip)
        if self.nvars < 0:
            raise KidCudaError("unsupported (moisture, precip) pair")
        self._h = C.c_void_p()
        self._check(self.lib.kid_create(C.byref(cfg), C.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise KidCudaError(self.lib.kid_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.kid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs
    def set_params(self, table: np.ndarray, column_to_set: np.ndarray | None = None):
        t = np.ascontiguousarray(table, dtype=np.float64)
        nsets = 1 if t.ndim == 1 else t.shape[0]
        nparams = t.shape[-1]
        idx = None
        if column_to_set is not None:
            idx = np.ascontiguousarray(column_to_set, dtype=np.int32)
        self._check(self.lib.kid_set_params(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), nparams, nsets,
                                            idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None))

    def set_profiles(self, ρ_dry: np.ndarray, T: np.ndarray):
        r = np.ascontiguousarray(ρ_dry, dtype=self.dtype)
        t = np.ascontiguousarray(T, dtype=self.dtype)
        per_column = int(r.ndim == 2)
        self._check(self.lib.kid_set_profiles(self._h, _ptr(r), _ptr(t), per_column))
        self._profiles_per_column = bool(per_column)

    def set_column_scalar(self, which: int, values: np.ndarray | None):
        if values is None:
            self._check(self.lib.kid_set_column_scalar(self._h, which, None))
            return
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=self.dtype), (self.cfg.nc,)))
        self._check(self.lib.kid_set_column_scalar(self._h, which, _ptr(v)))

    def init_shipway_hill(self, sounding, p0: np.ndarray | None = None, dry: bool = False):
        """Profiles, q_surf and initial state built on the device (ρ_ivp + initial_condition_1d per column).
        sounding: (z_0, z_1, z_2, rv_0, rv_1, rv_2, tht_0, tht_1, tht_2); p0: one surface pressure per column."""
        snd = np.ascontiguousarray(sounding, dtype=np.float64)
        assert snd.size == 9
        p = None
        if p0 is not None:
            p = np.ascontiguousarray(np.broadcast_to(np.asarray(p0, dtype=self.dtype), (self.cfg.nc,)))
        self._check(self.lib.kid_init_shipway_hill(self._h, snd.ctypes.data_as(C.POINTER(C.c_double)),
                                                   None if p is None else _ptr(p), int(dry)))
        self._profiles_per_column = True

    def init_gamma_0d(self, qt, Nd, k, rho_dry):
        """Box / col_sed initial condition built on the device (initial_condition_0d per column; scalars broadcast)."""
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=self.dtype), (self.cfg.nc,))) for x in (qt, Nd, k, rho_dry)]
        self._check(self.lib.kid_init_gamma_0d(self._h, *[_ptr(a) for a in arrs]))
        self._profiles_per_column = True

    def get_profiles(self):
        """(ρ_dry, T) as uploaded / built on the device: [nc, nz] each (or [nz] for one shared profile)."""
        shape = (self.cfg.nc, self.cfg.nz) if getattr(self, "_profiles_per_column", True) else (self.cfg.nz,)
        r, t = np.empty(shape, dtype=self.dtype), np.empty(shape, dtype=self.dtype)
        self._check(self.lib.kid_get_profiles(self._h, _ptr(r), _ptr(t)))
        return r, t

    def set_state(self, Y: np.ndarray):
        y = np.ascontiguousarray(Y, dtype=self.dtype)
        assert y.size == self.cfg.nc * self.nvars * self.cfg.nz, (y.shape, self.cfg.nc, self.nvars, self.cfg.nz)
        self._check(self.lib.kid_set_state(self._h, _ptr(y)))

    # ---- outputs
    def get_state(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.cfg.nc, self.nvars, self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kid_get_state(self._h, _ptr(out)))
        return out

    def set_state_async(self, Y_pinned: np.ndarray):
        """Enqueue the upload only; Y_pinned must be page-locked, C-contiguous FLOAT_TYPE and stay untouched until synchronize()."""
        assert Y_pinned.dtype == self.dtype and Y_pinned.flags.c_contiguous
        self._check(self.lib.kid_set_state_async(self._h, _ptr(Y_pinned)))

    def set_column_map(self, host_column_of: np.ndarray | None):
        """Device column i <-> host column host_column_of[i] in set_state* / get_state* (page-locked host arrays)."""
        if host_column_of is None:
            self._check(self.lib.kid_set_column_map(self._h, None))
            return
        m = np.ascontiguousarray(host_column_of, dtype=np.int32)
        assert m.size == self.cfg.nc
        self._check(self.lib.kid_set_column_map(self._h, m.ctypes.data_as(C.POINTER(C.c_int32))))

    def get_state_async(self, out_pinned: np.ndarray):
        assert out_pinned.dtype == self.dtype and out_pinned.flags.c_contiguous
        self._check(self.lib.kid_get_state_async(self._h, _ptr(out_pinned)))

    def rhs(self, t: float) -> np.ndarray:
        out = np.empty((self.cfg.nc, self.nvars, self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kid_rhs(self._h, float(t), _ptr(out)))
        return out

    def get_aux(self, t: float, field) -> np.ndarray:
        fid = AUX_ID[field] if isinstance(field, str) else int(field)
        out = np.empty((self.cfg.nc, self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kid_get_aux(self._h, float(t), fid, _ptr(out)))
        return out

    @staticmethod
    def _field_ids(fields) -> np.ndarray:
        return np.asarray([AUX_ID[f] if isinstance(f, str) else int(f) for f in fields], dtype=np.int32)

    def get_aux_many(self, t: float, fields, out: np.ndarray | None = None) -> np.ndarray:
        """[nc, nfields, nz]: all requested aux fields from ONE rhs evaluation (one kernel launch)."""
        ids = self._field_ids(fields)
        if out is None:
            out = np.empty((self.cfg.nc, len(ids), self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kid_get_aux_many(self._h, float(t), len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(out)))
        return out

    def output_enqueue(self, t: float, fields, out_pinned: np.ndarray):
        """Asynchronous get_aux_many into page-locked memory (see pinned_empty); valid after output_wait()."""
        ids = self._field_ids(fields)
        assert out_pinned.flags["C_CONTIGUOUS"] and out_pinned.dtype == self.dtype
        assert out_pinned.size == self.cfg.nc * len(ids) * self.cfg.nz
        self._check(self.lib.kid_output_enqueue(self._h, float(t), len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                _ptr(out_pinned)))

    def output_wait(self):
        self._check(self.lib.kid_output_wait(self._h))

    def reduce_max(self, t: float, field) -> float:
        fid = AUX_ID[field] if isinstance(field, str) else int(field)
        v = C.c_double()
        self._check(self.lib.kid_reduce_max(self._h, float(t), fid, C.byref(v)))
        return v.value

    # ---- calibration observables accumulated on the device (ODEsolution2Gvector)
    def obs_configure(self, variables, nz_per_filtered_cell=None, nt_filtered: int = 1, nt_per_filtered_cell: int = 1) -> int:
        ids = np.asarray([OBS_ID[v] if isinstance(v, str) else int(v) for v in variables], dtype=np.int32)
        per = np.ones(len(ids), dtype=np.int32) if nz_per_filtered_cell is None else np.asarray(nz_per_filtered_cell, dtype=np.int32)
        if per.shape != ids.shape:
            raise ValueError("nz_per_filtered_cell needs one entry per variable")
        n_out = C.c_int64()
        self._check(self.lib.kid_obs_configure(self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                               per.ctypes.data_as(C.POINTER(C.c_int32)), int(nt_filtered),
                                               int(nt_per_filtered_cell), C.byref(n_out)))
        self._obs_n_out = int(n_out.value)
        return self._obs_n_out

    def obs_accumulate(self, time_cell: int):
        self._check(self.lib.kid_obs_accumulate(self._h, int(time_cell)))

    def obs_get(self) -> np.ndarray:
        """[nc, n_out] Float64: row c is the G vector of column c, ordered [time cell][variable][filtered level]."""
        out = np.empty((self.cfg.nc, self._obs_n_out), dtype=np.float64)
        self._check(self.lib.kid_obs_get(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    # ---- stepping
    def step(self, t0: float, nsteps: int):
        self._check(self.lib.kid_step(self._h, float(t0), int(nsteps)))

    def set_steps_per_launch(self, n: int):
        self._check(self.lib.kid_set_steps_per_launch(self._h, int(n)))

    def synchronize(self):
        self._check(self.lib.kid_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(self.lib.kid_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.kid_launch_count(self._h))

    # ---- K2D split-stage API (domain decomposition)
    def k2d_stage_begin(self, t0: float, stage: int):
        self._check(self.lib.kid_k2d_stage_begin(self._h, float(t0), int(stage)))

    def k2d_stage_end(self, stage: int):
        self._check(self.lib.kid_k2d_stage_end(self._h, int(stage)))

    def k2d_exchange_local(self, right: "Handle"):
        self._check(self.lib.kid_k2d_exchange_local(self._h, right._h))

    def k2d_comm_init(self, rank: int, world: int, unique_id: bytes):
        """In-library NCCL ring of a decomposed K2D domain: afterwards step() exchanges the edge columns itself."""
        assert len(unique_id) == 128
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.kid_k2d_comm_init(self._h, int(rank), int(world), buf))

    def k2d_edge_buffers(self):
        p = [C.c_void_p() for _ in range(4)]
        n = C.c_size_t()
        self._check(self.lib.kid_k2d_edge_buffers(self._h, *[C.byref(x) for x in p], C.byref(n)))
        return [x.value for x in p], n.value
