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


def _bind(so: Path) -> C.CDLL:
    if True:
        if True:
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
        lib.kidoracle_set_threads.argtypes = [C.c_int]
        lib.kidoracle_set_threads.restype = C.c_int
    return lib


def load() -> C.CDLL:
    """The portable build (g++ -O2): the parity checker of the test-suite."""
    global _LIB
    if _LIB is None:
        so = _HERE / "libkid_oracle.so"
        if not so.exists():
            build()
        _LIB = _bind(so)
    return _LIB


_NATIVE = None


def load_native():
    """(library, description): `make native` (g++ -O3 -march=native) built ON THIS MACHINE — the CPU-baseline build of
    bench.py; falls back to the portable build when the host has no compiler."""
    global _NATIVE
    if _NATIVE is None:
        so = _HERE / "_native" / "libkid_oracle.so"
        try:
            subprocess.run(["make", "-C", str(_HERE), "native"], check=True, capture_output=True)
            _NATIVE = (_bind(so), "g++ -O3 -march=native -fopenmp -ffp-contract=off, built on this host")
        except Exception:  # no compiler on the box: time the portable build and say so
            _NATIVE = (load(), "g++ -O2 -fopenmp -ffp-contract=off (portable build)")
    return _NATIVE


def set_threads(n: int, lib=None) -> int:
    """OpenMP threads of the column loops (torchrun exports OMP_NUM_THREADS=1 to its workers); returns the count in effect."""
    return int((lib or load()).kidoracle_set_threads(int(n)))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleHandle:
    def __init__(self, cfg: KidConfig, lib=None):
        self.lib = lib or load()
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


class OracleEnsembleHandle:
    """Handle surface of lib.Handle for a multi-parameter-set ensemble, on top of one OracleHandle per set, plus the
    obs_* calls (numpy restatement in oracle/diagnostics.py).  Lets the CPU tests drive kid_b200.calibration."""

    def __init__(self, cfg: KidConfig):
        self.cfg, self.dtype = cfg, dtype_of(cfg.float_type)
        self._subs, self._cols = [], []
        self._pending = {}

    # configuration is recorded and replayed into the per-set handles once the parameter table is known
    def set_params(self, table, column_to_set=None):
        import copy
        t = np.atleast_2d(np.asarray(table, dtype=np.float64))
        c2s = np.zeros(self.cfg.nc, dtype=int) if column_to_set is None else np.asarray(column_to_set)
        for s in range(t.shape[0]):
            cols = np.where(c2s == s)[0]
            if cols.size == 0:
                continue
            cfg = copy.copy(self.cfg)
            cfg.nc = int(cols.size)
            h = OracleHandle(cfg)
            h.set_params(t[s])
            self._subs.append(h)
            self._cols.append(cols)
        self.nvars = self._subs[0].nvars

    def set_profiles(self, ρ_dry, T):
        self._rho_dry = np.asarray(ρ_dry, dtype=self.dtype)
        for h, cols in zip(self._subs, self._cols):
            h.set_profiles(ρ_dry[cols] if np.ndim(ρ_dry) == 2 else ρ_dry, T[cols] if np.ndim(T) == 2 else T)

    def set_column_scalar(self, which, values):
        for h, cols in zip(self._subs, self._cols):
            h.set_column_scalar(which, None if values is None else np.broadcast_to(np.asarray(values), (self.cfg.nc,))[cols])

    def set_state(self, Y):
        for h, cols in zip(self._subs, self._cols):
            h.set_state(np.asarray(Y)[cols])

    def get_state(self, out=None):
        out = np.empty((self.cfg.nc, self.nvars, self.cfg.nz), dtype=self.dtype) if out is None else out
        for h, cols in zip(self._subs, self._cols):
            out[cols] = h.get_state()
        return out

    def get_aux(self, t, field):
        out = np.empty((self.cfg.nc, self.cfg.nz), dtype=self.dtype)
        for h, cols in zip(self._subs, self._cols):
            out[cols] = h.get_aux(t, field)
        return out

    def step(self, t0, nsteps):
        for h in self._subs:
            h.step(t0, nsteps)
        self._t = t0 + nsteps * self.cfg.dt

    def synchronize(self):
        pass

    def obs_configure(self, variables, nz_per_filtered_cell=None, nt_filtered=1, nt_per_filtered_cell=1):
        from kid_b200 import equation_types as ET  # names of the state variables of this style
        self._var_names = ET.state_variable_names_by_enum(self.cfg.moisture, self.cfg.precip)
        self._obs_vars = list(variables)
        nz = self.cfg.nz
        per = np.ones(len(variables), dtype=int) if nz_per_filtered_cell is None else np.asarray(nz_per_filtered_cell)
        self._obs_per, self._obs_ntp = per, nt_per_filtered_cell
        self._obs_off = np.concatenate([[0], np.cumsum(nz // per)])
        self._obs_G = np.zeros((self.cfg.nc, nt_filtered, int(self._obs_off[-1])))
        self._t = getattr(self, "_t", 0.0)
        return nt_filtered * int(self._obs_off[-1])

    def obs_accumulate(self, time_cell):
        from . import diagnostics as OD
        Y = self.get_state()
        need_vt = any("rain" in v for v in self._obs_vars)
        vt = self.get_aux(self._t, "term_vel_rai") if need_vt else None
        for c in range(self.cfg.nc):
            u = {n: Y[c, i] for i, n in enumerate(self._var_names)}
            rd = self._rho_dry[c] if self._rho_dry.ndim == 2 else self._rho_dry
            for j, var in enumerate(self._obs_vars):
                x = np.asarray(OD.get_variable_data_from_ODE(u, rd, var, None if vt is None else vt[c]), dtype=np.float64)
                cell = x.reshape(-1, self._obs_per[j]).mean(axis=1)
                self._obs_G[c, time_cell, self._obs_off[j]:self._obs_off[j + 1]] += cell / self._obs_ntp

    def obs_get(self):
        return self._obs_G.reshape(self.cfg.nc, -1).copy()
