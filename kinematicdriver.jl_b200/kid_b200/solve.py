"""`ODE.solve` replacement and the driver entry points of the reference, on top of libkid_cuda.so.

Mirrors (reference):
  ODE.solve(problem, ODE.SSPRK33(); dt, saveat[, callback])   run_KiD_simulation.jl:159-166, run_box_simulation.jl:84-85
  CO.simulation_output(aux, t) / condition_io / affect_io!      src/Common/netcdf_io.jl:144-194
  run_KiD_simulation / run_box_simulation / run_K2D_simulation  test/experiments/*/

The state stays on the device between save times; at every save time the state (and, if a callback asks for them, the
aux fields) is downloaded once.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import k2d as K2
from . import model as M
from .lib import AUX_FIELDS, Handle


@dataclass
class ODESolution:
    """What the reference's consumers read from the OrdinaryDiffEq solution: `.t` and `.u`
    (run_box_simulation.jl:88-94, calibration/KiDUtils.jl:332-357)."""
    t: list = field(default_factory=list)
    u: list = field(default_factory=list)   # each [nc, nvars, nz]
    var_names: tuple = ()
    handle: object = None

    def variable(self, name: str) -> np.ndarray:
        """[n_saved, nc, nz] series of one prognostic variable."""
        i = self.var_names.index(name)
        return np.stack([u[:, i, :] for u in self.u])


class OutputCadence:
    """condition_io of src/Common/netcdf_io.jl:176-185, including its self-mutating accumulator: dt_io starts at
    TS.dt_io, grows by dt every step and fires (and resets to 0) when it exceeds `output_interval`, or at t≈t_max."""

    def __init__(self, TS: M.TimeStepping, output_interval: float = 1.0):
        self.TS, self.interval = TS, output_interval
        self.dt_io = TS.dt_io

    def __call__(self, t: float) -> bool:
        self.dt_io += self.TS.dt
        if self.dt_io > self.interval:
            self.dt_io = 0.0
            return True
        return abs(t - self.TS.t_max) < 1e-14 * max(1.0, abs(self.TS.t_max))


def simulation_output(handle, t: float, fields=None) -> dict:
    """The aux fields `CO.simulation_output(aux, t)` appends to the NetCDF file (netcdf_io.jl:144-170) plus the
    `ql_max` time series entry, read from the device at time t."""
    names = list(fields or AUX_FIELDS)
    if hasattr(handle, "get_aux_many"):
        A = handle.get_aux_many(t, names)  # one rhs evaluation / kernel launch for all fields
        out = {f: A[:, i, :] for i, f in enumerate(names)}
    else:
        out = {f: handle.get_aux(t, f) for f in names}
    out["ql_max"] = float(np.max(out["q_liq"])) if "q_liq" in out else handle.reduce_max(t, "q_liq")
    return out


def solve(case: M.Case, tspan=None, dt: float | None = None, saveat: float | None = None, callback=None,
          condition=None, handle=None) -> ODESolution:
    """ODE.solve(ODEProblem(rhs!, Y, tspan, aux), SSPRK33(); dt, saveat, callback).

    saveat: save interval (the state at t0, every multiple of saveat and t_end is returned, as OrdinaryDiffEq does).
    callback(handle, t): called after a step when condition(t) is true (DiscreteCallback(condition_io, affect_io!));
    with a callback the device advances step by step between its firings.
    """
    TS = case.meta.get("TS")
    dt = float(dt if dt is not None else case.cfg.dt)
    if abs(dt - case.cfg.dt) > 0:
        raise ValueError("dt differs from the dt the case was built with (it enters every limiter: TS.dt)")
    t0, t_end = tspan if tspan is not None else (0.0, TS.t_max)
    saveat = float(saveat if saveat is not None else (TS.dt_io if TS else t_end - t0))
    h = handle or M.make_handle(case)
    sol = ODESolution(var_names=case.var_names, handle=h)
    nsteps = int(round((t_end - t0) / dt))
    save_every = max(1, int(round(saveat / dt)))
    sol.t.append(t0)
    sol.u.append(h.get_state())
    done = 0
    while done < nsteps:
        nxt = min(nsteps, (done // save_every + 1) * save_every)
        if callback is None:
            h.step(t0 + done * dt, nxt - done)
            done = nxt
        else:
            while done < nxt:
                h.step(t0 + done * dt, 1)
                done += 1
                t = t0 + done * dt
                if condition is None or condition(t):
                    callback(h, t)
        sol.t.append(t0 + done * dt)
        sol.u.append(h.get_state())
    return sol


class OutputPipeline:
    """The output callback of the drivers (DiscreteCallback(condition_io, affect_io!), netcdf_io.jl:176-194) without
    stalling the integration: the aux fields of an output time are evaluated on the device and downloaded by a second
    stream into a ring of page-locked buffers while the next steps run; `sink(t, {field: [nc, nz]})` is called when a
    buffer has landed (one output late).  The arrays handed to a user sink are VIEWS of the page-locked ring, valid
    during the call only (write them to disk or copy what is kept); the default sink keeps copies in `.results`."""

    def __init__(self, handle, fields=None, sink=None, depth: int = 2):
        from .lib import pinned_empty
        self.h, self.fields = handle, list(fields or AUX_FIELDS)
        self.results = []
        self.sink = sink or (lambda t, out: self.results.append(
            (t, {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in out.items()})))
        shape = (handle.cfg.nc, len(self.fields), handle.cfg.nz)
        self.ring = [pinned_empty(shape, handle.dtype, handle.lib) for _ in range(depth)]
        self.pending = []  # (slot, t) in flight, oldest first
        self.n = 0

    def _retire(self, count: int):
        if not count:
            return
        self.h.output_wait()  # the copy stream is in order: everything enqueued so far has landed
        for slot, t in self.pending[:count]:
            A = self.ring[slot]
            self.sink(t, dict({f: A[:, i, :] for i, f in enumerate(self.fields)},
                              ql_max=float(A[:, self.fields.index("q_liq")].max()) if "q_liq" in self.fields else None))
        del self.pending[:count]

    def enqueue(self, t: float):
        slot = self.n % len(self.ring)
        if len(self.pending) == len(self.ring):  # the slot about to be reused must be consumed first
            self._retire(1)
        self.h.output_enqueue(t, self.fields, self.ring[slot])
        self.pending.append((slot, t))
        self.n += 1

    def drain(self):
        self._retire(len(self.pending))
        return self.results


# ------------------------------------------------------------------ drivers
def run_KiD_simulation(FT, opts: dict | None = None, nc: int = 1, output: bool = False):
    """test/experiments/KiD_driver/run_KiD_simulation.jl:13-166 (minus NetCDF/plots): returns (solution, outputs)."""
    opts = dict(opts or {})
    FT = {"Float64": np.float64, "Float32": np.float32}.get(opts.pop("FLOAT_TYPE", None), FT)
    opts.pop("plotting_flag", None)
    case = M.k1d_case(FT, nc=nc, **opts)
    TS = case.meta["TS"]
    outputs = []
    cb = cond = pipe = None
    h = M.make_handle(case)
    if output:
        cond = OutputCadence(TS)
        if hasattr(h, "output_enqueue"):  # asynchronous, double-buffered download of the output fields
            pipe = OutputPipeline(h)
            outputs = pipe.results
            cb = lambda hh, t: pipe.enqueue(t)  # noqa: E731
            pipe.enqueue(0.0)
        else:
            cb = lambda hh, t: outputs.append((t, simulation_output(hh, t)))  # noqa: E731
            outputs.append((0.0, simulation_output(h, 0.0)))
    o = case.meta["opts"]
    sol = solve(case, (o["t_ini"], o["t_end"]), dt=TS.dt, saveat=TS.dt_io, callback=cb, condition=cond, handle=h)
    if pipe is not None:
        pipe.drain()
    return sol, outputs


def run_box_simulation(FT, opts: dict | None = None, nc: int = 1):
    """test/experiments/box_driver/run_box_simulation.jl:12-124 (minus plots)."""
    o = dict(qt=1e-3, Nd=1e8, k=2.0, rhod=1.22, precipitation_choice="Precipitation2M", rain_formation_choice="SB2006",
             dt=1.0, dt_output=30.0, t_ini=0.0, t_end=3600.0, microphys_params={})
    o.update(opts or {})
    case = M.box_case(FT, nc=nc, precipitation_choice=o["precipitation_choice"],
                      rain_formation_choice=o["rain_formation_choice"], qt=o["qt"], Nd=o["Nd"], k=o["k"], rhod=o["rhod"],
                      dt=o["dt"], dt_output=o["dt_output"], t_ini=o["t_ini"], t_end=o["t_end"],
                      microphys_params=o["microphys_params"])
    return solve(case, (o["t_ini"], o["t_end"]), dt=o["dt"], saveat=o["dt_output"])


def run_KiD_col_sed_simulation(FT, opts: dict | None = None, nc: int = 1):
    """test/experiments/KiD_col_sed_driver/run_KiD_col_sed_simulation.jl:13-146 (minus NetCDF/plots), with the script's
    option names (qt, prescribed_Nd, k, rhod, precipitation_choice, rain_formation_choice, sedimentation_choice, z_min,
    z_max, n_elem, dt, dt_output, t_ini, t_end)."""
    o = dict(precipitation_choice="Precipitation2M", rain_formation_choice=None, sedimentation_choice=None)
    o.update(opts or {})
    if o["precipitation_choice"] == "CloudyPrecip":
        raise NotImplementedError("CloudyPrecip is outside the device hot path of libkid_cuda.so and there is no CPU fallback")
    case = M.col_sed_case(FT, nc=nc, **o)
    oo = case.meta["opts"]
    return solve(case, (oo["t_ini"], oo["t_end"]), dt=oo["dt"], saveat=oo["dt_output"])


def run_K2D_simulation(FT, opts: dict | None = None):
    """test/experiments/Ki2D_driver/run_kinematic2d_simulations.jl:15-171 (single-domain, minus NetCDF/plots)."""
    o = dict(opts or {})
    case = K2.k2d_case(FT, **o)
    oo = case.meta["opts"]
    return solve(case, (oo["t_ini"], oo["t_end"]), dt=oo["dt"], saveat=oo["dt_output"])


class PipelinedEnsemble:
    """A large ensemble split into column chunks, one handle (and CUDA stream) per chunk: the upload of chunk i+1 and
    the download of chunk i-1 overlap the stepping of chunk i.  Host arrays must be page-locked (pinned)."""

    def __init__(self, case: M.Case, nchunks: int = 4, host_column_of: np.ndarray | None = None):
        """host_column_of: column i of `case` (the device storage order, e.g. model.warp_coherent_order) is row
        host_column_of[i] of the host arrays handed to `solve` — the caller keeps its own member order, the layout kernels
        gather / scatter straight from / to the page-locked host array (kid_set_column_map)."""
        import copy
        nc = case.cfg.nc
        self.host_column_of = None if host_column_of is None else np.ascontiguousarray(host_column_of, dtype=np.int32)
        edges = np.linspace(0, nc, nchunks + 1).astype(int)
        edges = (edges // 32) * 32
        edges[-1] = nc
        self.ranges = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.handles = []
        for a, b in self.ranges:
            self.handles.append(M.make_handle(M.slice_columns(case, a, b)))
            if self.host_column_of is not None:
                self.handles[-1].set_column_map(self.host_column_of[a:b])

    @classmethod
    def for_member_order(cls, case: M.Case, nchunks: int = 4):
        """`case` and the host arrays handed to `solve` are in the CALLER's member order (any order).  Every chunk is a
        contiguous block of host rows whose members are stored on the device in their own warp-coherent order
        (model.warp_coherent_order of the chunk): the block travels through the copy engine and the layout kernel permutes on
        the device (block-local column map), so an unordered caller pays neither a host gather nor scattered PCIe reads."""
        self = cls.__new__(cls)
        nc = case.cfg.nc
        edges = (np.linspace(0, nc, nchunks + 1).astype(int) // 32) * 32
        edges[-1] = nc
        self.ranges = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.handles, maps = [], []
        for a, b in self.ranges:
            sub = M.slice_columns(case, a, b)
            perm = M.warp_coherent_order(sub)
            h = M.make_handle(M.reorder_columns(sub, perm))
            h.set_column_map(a + perm)
            self.handles.append(h)
            maps.append(a + perm)
        self.host_column_of = np.concatenate(maps).astype(np.int32)
        return self

    def set_steps_per_launch(self, n: int):
        for h in self.handles:
            h.set_steps_per_launch(n)

    def solve(self, Y_pinned: np.ndarray, t0: float, nsteps: int):
        """In place: Y_pinned[column, variable, level] holds the state at t0 on entry and at t0 + nsteps·dt on return."""
        for (a, b), h in zip(self.ranges, self.handles):
            view = Y_pinned if self.host_column_of is not None else Y_pinned[a:b]  # with a map: rows are addressed through it
            h.set_state_async(view)
            h.step(t0, nsteps)
            h.get_state_async(view)
        for h in self.handles:
            h.synchronize()
        return Y_pinned

    @property
    def launch_count(self):
        return sum(h.launch_count for h in self.handles)
