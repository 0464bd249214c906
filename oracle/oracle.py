"""ORACLE — test infrastructure only (PARITY UNPINNED vs. real Julia output, see kid_oracle.cpp).

ctypes wrapper of oracle/libkid_oracle.so exposing the same method surface as
kid_b200.lib.Handle, so that parity tests drive both with identical inputs.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_ROOT = _HERE.parent
sys.path.insert(0, str(_ROOT / "kinematicdriver.jl_b200"))
from kid_b200.lib import AUX_ID, KidConfig, dtype_of  # noqa: E402  (interface definitions only)

_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libkid_oracle.so"
    srcs = [_HERE / "kid_oracle.cpp", _HERE / "kid_physics.hpp", _ROOT / "include" / "kid_cuda.h",
            _ROOT / "include" / "kid_params.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return so


def load() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "libkid_oracle.so"
        if not so.exists():
            build()
        lib = C.CDLL(str(so))
        H = C.c_void_p
        lib.kidoracle_create.restype = H
        lib.kidoracle_create.argtypes = [C.POINTER(KidConfig)]
        lib.kidoracle_destroy.argtypes = [H]
        lib.kidoracle_nvars.argtypes = [H]
        lib.kidoracle_set_params.argtypes = [H, C.POINTER(C.c_double), C.c_int]
        lib.kidoracle_set_profiles.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int]
        lib.kidoracle_set_column_scalar.argtypes = [H, C.c_int, C.c_void_p]
        lib.kidoracle_set_state.argtypes = [H, C.c_void_p]
        lib.kidoracle_get_state.argtypes = [H, C.c_void_p]
        lib.kidoracle_step.argtypes = [H, C.c_double, C.c_int]
        lib.kidoracle_rhs.argtypes = [H, C.c_double, C.c_void_p]
        lib.kidoracle_get_aux.argtypes = [H, C.c_double, C.c_int, C.c_void_p]
        lib.kidoracle_gamma_upper.restype = C.c_double
        lib.kidoracle_gamma_upper.argtypes = [C.c_double, C.c_double]
        lib.kidoracle_expint_E1.restype = C.c_double
        lib.kidoracle_expint_E1.argtypes = [C.c_double]
        lib.kidoracle_gll.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleHandle:
    def __init__(self, cfg: KidConfig):
        self.lib = load()
        self.cfg = cfg
        self.dtype = dtype_of(cfg.float_type)
        self._h = C.c_void_p(self.lib.kidoracle_create(C.byref(cfg)))
        if not self._h.value:
            raise RuntimeError("oracle: unsupported configuration")
        self.nvars = self.lib.kidoracle_nvars(self._h)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.kidoracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle call failed rc={rc}")

    def set_params(self, table, column_to_set=None):
        t = np.ascontiguousarray(table, dtype=np.float64)
        if t.ndim != 1:
            raise NotImplementedError("oracle handles one parameter set per handle; run one handle per set")
        self._check(self.lib.kidoracle_set_params(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), t.size))

    def set_profiles(self, ρ_dry, T):
        r = np.ascontiguousarray(ρ_dry, dtype=self.dtype)
        t = np.ascontiguousarray(T, dtype=self.dtype)
        self._check(self.lib.kidoracle_set_profiles(self._h, _ptr(r), _ptr(t), int(r.ndim == 2)))

    def set_column_scalar(self, which, values):
        if values is None:
            self._check(self.lib.kidoracle_set_column_scalar(self._h, which, None))
            return
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=self.dtype), (self.cfg.nc,)))
        self._check(self.lib.kidoracle_set_column_scalar(self._h, which, _ptr(v)))

    def set_state(self, Y):
        y = np.ascontiguousarray(Y, dtype=self.dtype)
        assert y.size == self.cfg.nc * self.nvars * self.cfg.nz
        self._check(self.lib.kidoracle_set_state(self._h, _ptr(y)))

    def get_state(self, out=None):
        if out is None:
            out = np.empty((self.cfg.nc, self.nvars, self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kidoracle_get_state(self._h, _ptr(out)))
        return out

    def rhs(self, t):
        out = np.empty((self.cfg.nc, self.nvars, self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kidoracle_rhs(self._h, float(t), _ptr(out)))
        return out

    def get_aux(self, t, field):
        fid = AUX_ID[field] if isinstance(field, str) else int(field)
        out = np.empty((self.cfg.nc, self.cfg.nz), dtype=self.dtype)
        self._check(self.lib.kidoracle_get_aux(self._h, float(t), fid, _ptr(out)))
        return out

    def reduce_max(self, t, field):
        return float(self.get_aux(t, field).max())

    def step(self, t0, nsteps):
        self._check(self.lib.kidoracle_step(self._h, float(t0), int(nsteps)))

    def synchronize(self):
        pass


def gamma_upper(a, x):
    return load().kidoracle_gamma_upper(float(a), float(x))


def expint_E1(x):
    return load().kidoracle_expint_E1(float(x))


def gll(Nq):
    x = np.empty(Nq); w = np.empty(Nq); D = np.empty((Nq, Nq))
    load().kidoracle_gll(Nq, _ptr(x), _ptr(w), _ptr(D))
    return x, w, D
