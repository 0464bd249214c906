#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric: grid-point-steps/s).

Workload: BASELINE.json configs[3] — KiD 1-D NonEquilibriumMoisture + Precipitation2M (SB2006) ensemble of
1,048,576 columns x 128 levels, Float32, the synthetic calibration-style ensemble of SURVEY §8(d) (per-column w1,
prescribed Nd, surface pressure -> per-column ρ_dry/T profiles built on the device).
--gpus N shards the SAME ensemble over the N ranks (strong scaling, `--columns` is the total; blocks of 4096 columns dealt
round-robin, no data-path collective).  --weak keeps `--columns` per GPU instead (round-1 mode).

One "step" = one SSPRK33 step (3 rhs evaluations) of every column of the ensemble.

    python bench.py --gpus N --steps K --warmup W          # ours  (libkid_cuda.so, sm_100a)
    python bench.py --impl reference ...                   # the reference's CPU path on the host cores

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the stream the kernels are launched on, state
resident in HBM; `e2e` goes through the public host API with pinned HOST buffers (upload + K steps + download inside the
timed region).  Both arms time the same model-time window [t_start + W, t_start + W + K) of the same member ordering; the
CPU arm steps the first columns of that ensemble (a bounded sample), and the GPU arm checks its own result on those
columns against the CPU arm's (`parity_sample`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200"))
sys.path.insert(0, str(ROOT))

NZ = 128
BYTES_PER_GP_STEP = 2 * 8 * 4 + 2 * 4  # NonEq+2M f32: read+write 8 prognostics, read per-column ρ_dry, T
METRIC = "grid-point-steps/s"
PREROLL_SPL = 60  # the untimed pre-roll runs in launches of 60 fused steps
WORKLOAD = "KiD 1D NonEq+Precipitation2M(SB2006) ensemble, 128 levels, Float32 (BASELINE configs[3])"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_profile():
    """ncu-derived constants of the dominant kernel (profiles/traffic.json): DRAM bytes per launch and executed
    thread-instructions per grid point per rhs.  Re-captured whenever the kernel changes (profiles/README.md)."""
    tp = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(tp.read_text())
    except Exception:
        return {}


def flops_per_gp_step():
    """Arithmetic operations of the REFERENCE formulation per grid-point-step (oracle/opcount.py)."""
    try:
        d = json.loads((ROOT / "profiles" / "r01_opcounts.json").read_text())
        vals = [v["arithmetic_flops_per_gp_step"] for v in d if v["style"].startswith("NonEquilibrium") and "2M" in v["style"]]
        if vals:
            return float(np.mean(vals)), "profiles/r01_opcounts.json (operation-counting oracle, NonEq+2M)"
    except Exception:
        pass
    return 1965.0, "DESIGN.md §3.4 (operation-counting oracle, NonEq+2M: 1939-1991)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ workload shared by both arms
def total_columns(args, world: int) -> int:
    return args.columns * world if args.weak else args.columns


def global_case(args, world: int, device: int = 0):
    """The whole SURVEY §8(d) ensemble (numpy only: member c has w1, Nd, p0 from the counter hash; hydrostatic profile and
    initial state are built on the device).  Member order: "sorted" stores the members in nested quantile bins of p0, w1 and
    Nd (model.nested_bin_order), so that neighbouring columns — the 32 lanes of a warp — follow the same microphysics branches;
    "hash" keeps the member-index order of §8(d)."""
    from kid_b200 import model as M
    return M.k1d_ensemble_case(np.float32, nc=total_columns(args, world), ic="device", order=args.order, device=device, n_elem=NZ,
                               moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")


SHARD_BLOCK = 4096  # columns per dealt block (a multiple of the 32-column row pitch and of the 16-column tile)


def shard_columns(total: int, rank: int, world: int) -> np.ndarray:
    """Columns of the (sorted) global ensemble that rank `rank` owns: blocks of SHARD_BLOCK columns dealt round-robin, so every
    rank steps the same mix of members.  (Contiguous halves of the sorted ensemble differ in cost — the surface pressure sets the
    cloud depth — and cost 16 % of the strong-scaling efficiency at N = 2.)  Columns never interact: no collective either way."""
    c = np.arange(total)
    return c if world == 1 else c[(c // SHARD_BLOCK) % world == rank]


def workload_config(args, world: int) -> dict:
    """The `config` object of the JSON line — identical in both arms (same ensemble, same member order, same window)."""
    total = total_columns(args, world)
    return {"workload": WORKLOAD, "columns": total, "levels": NZ,
            "window_model_time_s": [args.t_start + args.warmup, args.t_start + args.warmup + args.steps],
            "member_order": ("hash (SURVEY 8d counter hash)" if args.order == "hash"
                             else "sorted into nested quantile bins (p0, then w1, then Nd): model.nested_bin_order"),
            "initial_condition": "every member its own p0 (SURVEY 8d), hydrostatic profile + initial state per column",
            "l2": "no flush between timed steps: the state per GPU (%.2f GiB) is far larger than the 126 MB L2" % (total / world * NZ * 8 * 4 / 2**30),
            "parallelism": f"columns sharded x{world}: blocks of {SHARD_BLOCK} columns dealt round-robin, no collective"}


def sample_errors(a, b, var_names) -> dict:
    """max |a - b| per prognostic variable, relative to the field maximum (floor: 1e-6 of the largest field of the same
    kind, mass or number densities) — the metric of tests/parity.py."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    kind = lambda v: "N" if v.startswith("N_") else "q"
    scale = {}
    for i, v in enumerate(var_names):
        scale[kind(v)] = max(scale.get(kind(v), 0.0), float(np.max(np.abs(b[:, i]))))
    out = {}
    for i, v in enumerate(var_names):
        den = max(float(np.max(np.abs(b[:, i]))), 1e-6 * scale[kind(v)], 1e-300)
        out[v] = float("%.3e" % (np.max(np.abs(a[:, i] - b[:, i])) / den))
    return out


def knife_edge_measures(a, b, names, tol_q=5e-3, tol_n=5e-2) -> dict:
    """The number fields carry the reference formulation's knife-edge (find_cloud_base!: the sign of N_act - N_liq at the level
    where N_liq has just reached N_act decides which level activates): a flip moves the activated number by one level in a
    single column.  Two measures that do not hinge on it: how many columns exceed the tolerance, and the error of the
    column-integrated aerosol + droplet number (conserved by activation in the closed system)."""
    a64, b64 = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    over = {}
    for i, v in enumerate(names):
        tol = tol_n if v.startswith("N_") else tol_q
        den = max(np.abs(b64[:, i]).max(), 1e-300)
        over[v] = int((np.abs(a64[:, i] - b64[:, i]).max(axis=1) / den > tol).sum())
    iN = [names.index(v) for v in ("N_liq", "N_rai", "N_aer") if v in names]
    col_int = None
    if iN:
        sa, sb = a64[:, iN].sum(axis=(1, 2)), b64[:, iN].sum(axis=(1, 2))
        col_int = float("%.3e" % np.max(np.abs(sa - sb) / np.maximum(np.abs(sb), 1e-300)))
    return {"columns_over_tolerance": over, "tolerance": {"q": tol_q, "N": tol_n}, "column_integrated_number_rel_err": col_int}


def cpu_arm_threads(lib) -> int:
    """All host cores, explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers."""
    from oracle import oracle as O
    return O.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), lib)


def cpu_sample_columns(total: int) -> int:
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return int(min(total, 32 * cores))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Julia and the reference's un-vendored
    dependencies are absent (no oracle/_ref can be built), so this arm times the restated CPU path (oracle/, built
    -O3 -march=native on this host) with all host threads on a bounded sample of the arm's workload: the first columns of
    the same ensemble in the same member order, over the same model-time window."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    from kid_b200 import model as M
    from oracle import oracle as O
    lib, build = O.load_native()
    threads = cpu_arm_threads(lib)
    total = total_columns(args, world)
    ncpu = cpu_sample_columns(total)
    case = M.materialise_on_host(M.slice_columns(global_case(args, world), 0, ncpu))
    o = M.apply_case(O.OracleHandle(case.cfg, lib), case)
    # bring the sample to the developed state the device arm times (untimed), then the same warm-up steps
    o.step(0.0, int(args.t_start))
    t = float(args.t_start)
    o.step(t, args.warmup)
    t += args.warmup
    t0 = time.perf_counter()
    o.step(t, args.steps)
    dt = time.perf_counter() - t0
    value = ncpu * NZ * args.steps / dt
    sample = (f"first {ncpu} columns of the ensemble x {NZ} levels x {args.steps} SSPRK33 steps over the same model-time window; "
              f"restated CPU path (oracle/, {build}), OpenMP over columns on {threads} threads; "
              "Julia + CloudMicrophysics/ClimaCore are not installable here")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def cpu_check_and_baseline(case_sample, Y_start, Y_gpu_end, t_start: float, nsteps: int, target_s: float):
    """Rank 0 of the device arm: step the sampled columns on the CPU oracle from the state the timed region started
    from, compare with what the GPU produced for them (parity_sample), and time the CPU path on the same columns."""
    from kid_b200 import model as M
    from oracle import oracle as O
    import copy
    lib, build = O.load_native()
    threads = cpu_arm_threads(lib)
    sub = copy.copy(case_sample)
    sub.Y0 = Y_start
    o = M.apply_case(O.OracleHandle(sub.cfg, lib), sub)
    ncpu = Y_start.shape[0]
    t0 = time.perf_counter()
    o.step(t_start, nsteps)
    per_step = (time.perf_counter() - t0) / nsteps
    Yo = o.get_state()
    parity = sample_errors(Y_gpu_end, Yo, case_sample.var_names)
    parity = {"max_err_rel_field_max": parity, **knife_edge_measures(Y_gpu_end, Yo, case_sample.var_names)}
    cpu = None
    if target_s > 0:
        n2 = int(max(8, min(4000, target_s / max(per_step, 1e-6))))
        t0 = time.perf_counter()
        o.step(t_start + nsteps, n2)
        dt = time.perf_counter() - t0
        cpu = {"value": ncpu * NZ * n2 / dt, "unit": METRIC, "cores": threads, "kind": "port",
               "sample": f"first {ncpu} columns x {NZ} levels x {n2} SSPRK33 steps continuing the same developed ensemble state "
                         f"(t={t_start + nsteps:g}s); restated CPU path (oracle/, {build}), OpenMP over columns on {threads} threads; not Julia"}
    return parity, cpu


K2D_PARITY_STEPS = 3  # the K2D oracle is single-threaded: about a second per step of the 4096 x 256 domain


def k2d_parity_sample(case, Y_start, Y_gpu_end, t_start: float, nsteps: int) -> dict:
    """The whole K2D domain stepped `nsteps` steps on the CPU oracle from the state the GPU started from; errors per variable
    relative to the field maximum (the metric of tests/parity.py)."""
    from kid_b200 import model as M
    from oracle import oracle as O
    o = M.make_handle(case, O.OracleHandle)
    o.set_state(np.ascontiguousarray(Y_start))
    o.step(t_start, nsteps)
    Yo = o.get_state()
    return {"columns": int(case.cfg.nc), "levels": int(case.cfg.nz), "steps": nsteps, "vs": "CPU oracle from the same start state",
            "max_err_rel_field_max": sample_errors(Y_gpu_end, Yo, case.var_names), **knife_edge_measures(Y_gpu_end, Yo, case.var_names)}


def run_k2d(args, rank, world, local, dist):
    """BASELINE configs[4]: K2DModel, 1024 spectral elements x npoly 3 = 4096 GLL node columns x 256 levels, NonEq +
    Precipitation2M, Float32, decomposed in x by whole elements over the ranks; every SSPRK33 stage exchanges the packed
    GLL edge columns inside the library (kid_k2d_comm_init: one grouped ncclSend/ncclRecv per stage).  Strong scaling."""
    import torch
    from kid_b200 import k2d, model as M
    from kid_b200.lib import Handle
    kw = dict(nx=1024, nz=256, npoly=3, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
              domain_width=48000.0, domain_height=3000.0)
    cs = k2d.k2d_case(np.float32, rank=rank, world=world, device=local, **kw)
    h = M.make_handle(cs, Handle)
    if world > 1:
        k2d.LibraryRing(h, dist, rank, world)
    stream = torch.cuda.ExternalStream(h.stream, device=local)
    h.step(0.0, 20)
    h.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = h.launch_count
    e0.record(stream)
    h.step(20.0, args.k2d_steps)
    e1.record(stream)
    h.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    finite = bool(np.isfinite(h.get_state()).all())
    launches_per_step = (h.launch_count - l0) / args.k2d_steps
    parity = None
    if world == 1 and args.parity:
        # parity at BASELINE size: the whole 4096 x 256 domain stepped K2D_PARITY_STEPS steps on the CPU oracle from the state the
        # timed window ended in.  (Not later: with dt = 1 s this synthetic domain — u ~ w1 W/H = 32 m/s over 13 m node spacing —
        # violates the horizontal CFL limit and goes non-finite by t = 600 s in the oracle and on the GPU alike.)
        try:
            t_dev = 20.0 + args.k2d_steps
            Y1 = h.get_state()
            h.set_state(Y1)                      # both sides restart the "previous call's aux" of K2D activation from Y1
            h.step(t_dev, K2D_PARITY_STEPS)
            parity = k2d_parity_sample(cs, Y1, h.get_state(), t_dev, K2D_PARITY_STEPS)
            parity["model_time_s"] = t_dev
        except Exception as ex:
            parity = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    n_nodes = 4096
    return {"workload": "K2D 4096 GLL node columns x 256 levels, NonEq+Precipitation2M, Float32 (BASELINE configs[4])",
            "value": n_nodes * 256 * args.k2d_steps / (float(ms.item()) * 1e-3), "unit": METRIC, "n_gpus": world,
            "ms_per_step": float(ms.item()) / args.k2d_steps, "steps": args.k2d_steps, "scaling": "strong",
            "exchange": "in-library NCCL send/recv of the GLL edge columns, one group per SSPRK33 stage" if world > 1 else None,
            "launches_per_step": launches_per_step, "state_finite": finite, "parity_sample": parity}


def run_ours(args):
    import torch
    from kid_b200 import model as M
    from kid_b200 import lib as KL
    from kid_b200.lib import Handle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    numa = {}
    if world > 1:
        # several ranks share the host: keep each rank and its pinned staging buffers on its GPU's NUMA node
        pr = torch.cuda.get_device_properties(local)
        numa = KL.bind_to_gpu_numa_node(f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        import torch.distributed as dist
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # WARN still prints the version line: not on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    total = total_columns(args, world)
    mine = shard_columns(total, rank, world)
    nc = int(mine.size)
    gcase = global_case(args, world, local)
    case = M.take_columns(gcase, mine)
    if args.kernel == "scalar":
        case.cfg.reserved[3] = 1  # A/B: the scalar tile kernel instead of the packed two-levels-per-thread kernel
    if args.direct_host:
        case.cfg.reserved[4] = 2  # A/B: layout kernels access the page-locked host arrays in place instead of copy engine + staging
    h = M.make_handle(case, Handle)
    spl = args.steps_per_launch
    stream = torch.cuda.ExternalStream(h.stream, device=local)

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # untimed pre-roll into the developed state (updraft on: activation, condensation, collisions, advection), then W warm-up steps
    h.set_steps_per_launch(PREROLL_SPL)
    h.step(0.0, int(args.t_start))
    h.set_steps_per_launch(spl)
    t = float(args.t_start)
    h.step(t, args.warmup)
    t += args.warmup
    barrier()

    state_shape = (nc, len(case.var_names), NZ)
    Yh = torch.empty(state_shape, dtype=torch.float32).pin_memory()
    Yn = Yh.numpy()
    ncpu = cpu_sample_columns(total) if rank == 0 else 0
    ncpu = min(ncpu, nc)
    Y_start = None
    if rank == 0 and (args.cpu_seconds > 0 or args.parity):
        h.get_state(Yn)
        Y_start = Yn[:ncpu].copy()  # what the CPU arm starts from

    # ---------------- device-resident timing (value)
    fp32_peak = KL.measure_fp32_peak(local)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(1.0)  # let nvidia-smi finish initialising (it briefly holds driver locks) before the timed region
    n_launch = (args.steps + spl - 1) // spl
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_launch + 1)]
    l0 = h.launch_count
    barrier()
    t_timed0 = t
    ev[0].record(stream)
    done = 0
    for i in range(n_launch):
        n = min(spl, args.steps - done)
        h.step(t, n)
        t += n
        done += n
        ev[i + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = h.launch_count - l0
    ms_total = ev[0].elapsed_time(ev[-1])
    launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n_launch)]
    tt = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max = float(tt.item())
    gp_steps_total = total * NZ * args.steps
    value = gp_steps_total / (ms_max * 1e-3)

    # ---------------- end-to-end through the host API with pinned host buffers (e2e)
    # The user-facing call for a large ensemble: the columns go through `nchunks` handles (one CUDA stream each), so the
    # H2D upload, the K steps and the D2H download of different chunks overlap (kid_set_state_async / kid_step /
    # kid_get_state_async).  Everything — copies included — is inside the timed region.
    from kid_b200.solve import PipelinedEnsemble
    h.get_state(Yn)  # developed state as the "user's" host input
    Y_gpu_end = Yn[:ncpu].copy() if Y_start is not None else None
    finite = bool(np.isfinite(Yn[:: max(1, nc // 4096)]).all())
    pipe = PipelinedEnsemble(case, nchunks=args.e2e_chunks)
    pipe.set_steps_per_launch(spl)
    e2e_steps = args.steps
    Yw = torch.empty(state_shape, dtype=torch.float32).pin_memory().numpy()
    Yw[...] = Yn
    pipe.solve(Yw, t, spl)  # warm the chunk handles (first launch of each, pinned-page first touch)
    del Yw
    barrier()
    w0 = time.perf_counter()
    pipe.solve(Yn, t, e2e_steps)
    wall = time.perf_counter() - w0
    barrier()
    e2e_ms = wall * 1e3  # several streams: the host wall clock around the synchronised call is the measure
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = gp_steps_total / (float(te.item()) * 1e-3)
    # The same call for a caller whose host arrays are in MEMBER-INDEX order (SURVEY 8d hash order): every pipeline chunk is a
    # contiguous block of the caller's rows, stored on the device in the chunk's own warp-coherent order; the block travels
    # through the copy engine and the layout kernel permutes on the device (block-local kid_set_column_map): no host gather.
    e2e_user = None
    if args.order == "sorted":
        del pipe
        members = np.asarray(gcase.meta["column_perm"])[mine]       # member index of every device column of this rank
        by_member = np.argsort(members)                              # this rank's columns in member-index order
        case_m = M.reorder_columns(case, by_member)
        pipe = PipelinedEnsemble.for_member_order(case_m, nchunks=args.e2e_chunks)
        pipe.set_steps_per_launch(spl)
        Yu = torch.empty(state_shape, dtype=torch.float32).pin_memory().numpy()
        Yu[...] = Yn[by_member]
        pipe.solve(Yu, t, spl)  # warm-up
        Yu[...] = Yn[by_member]
        barrier()
        w0 = time.perf_counter()
        pipe.solve(Yu, t + e2e_steps, e2e_steps)
        wall = time.perf_counter() - w0
        barrier()
        tu = torch.tensor([wall * 1e3], dtype=torch.float64, device=f"cuda:{local}")
        if dist is not None:
            dist.all_reduce(tu, op=dist.ReduceOp.MAX)
        e2e_user = gp_steps_total / (float(tu.item()) * 1e-3)
        del Yu
    state_bytes = Yn.nbytes

    # ---------------- secondary: BASELINE configs[4], K2D 4096 nodes x 256 levels, x-decomposed over the same ranks
    k2d_line = None
    if args.k2d_steps > 0:
        try:
            del pipe
            k2d_line = run_k2d(args, rank, world, local, dist)
        except Exception as ex:  # never lose the headline line over the secondary measurement
            k2d_line = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank == 0:
        peak, peak_src = measured_peaks()
        avg_launch_ms = statistics.mean(launch_ms)
        steps_per = args.steps / n_launch
        achieved = BYTES_PER_GP_STEP * nc * NZ * steps_per / (avg_launch_ms * 1e-3) / 1e9
        parity, cpu = None, None
        if Y_start is not None:
            sample_case = M.materialise(M.slice_columns(case, 0, ncpu))  # device-built profiles of the sample -> host
            parity, cpu = cpu_check_and_baseline(sample_case, Y_start, Y_gpu_end, t_timed0, args.steps, args.cpu_seconds)
        prof = kernel_profile()
        traffic = prof.get("dram_bytes_per_gp_step")
        traffic = traffic * nc * NZ * steps_per if traffic else None
        # Float32 roofline: the reference formulation's arithmetic per grid-point-step (operation-counting oracle) over
        # the FFMA rate measured in this run; issue roofline: executed warp-instructions (ncu count of the committed
        # capture) over 4 issue slots per cycle per SM at the SM clock sampled during the timed region
        fl, fl_src = flops_per_gp_step()
        per_gpu_rate = nc * NZ * steps_per / (avg_launch_ms * 1e-3)  # grid-point-steps/s of this GPU's kernel
        fp32_frac = fl * per_gpu_rate / fp32_peak["flops_per_s"]
        inst = prof.get("thread_inst_per_gp_rhs")
        sm_mhz = clocks.get("sm_mhz") or fp32_peak["sm_clock_khz"] / 1e3
        issue = None
        if inst:
            wipc = inst * 3 * per_gpu_rate / 32 / (fp32_peak["sm_count"] * sm_mhz * 1e6)
            issue = {"warp_inst_per_clk_per_sm": wipc, "peak": 4.0, "frac": wipc / 4.0, "thread_inst_per_gp_rhs": inst,
                     "source": prof.get("source")}
        fracs = {"hbm": achieved / peak, "fp32": fp32_frac, "issue": issue["frac"] if issue else 0.0}
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "run": {"columns_per_gpu": nc, "steps_per_launch": spl, "state_finite": finite,
                    "l2": "state per GPU (%.2f GiB) is far larger than the 126 MB L2; no flush needed" % (state_bytes / 2**30),
                    **({"host_numa_binding": numa} if numa else {})},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": world * state_bytes / e2e_steps,
                    "d2h_bytes_per_step": world * state_bytes / e2e_steps, "resident_fraction": e2e_value / value,
                    # what the end-to-end leg moved between host and devices, all ranks together (the ranks share one host
                    # memory system: at N = 8 this, not the GPUs, bounds the leg)
                    "host_transfer_gbs_all_ranks": 2 * world * state_bytes / (float(te.item()) * 1e-3) / 1e9,
                    "member_index_order_host_arrays": ({"value": e2e_user, "of_sorted_host_arrays": e2e_user / e2e_value,
                                                        "how": "every pipeline chunk = a contiguous block of the caller's rows stored in its own warp-coherent order "
                                                               "(block-local kid_set_column_map: copy engine + permutation on the device)"}
                                                       if e2e_user else None),
                    "call": f"PipelinedEnsemble.solve: {args.e2e_chunks} column chunks x [set_state_async(pinned) + step({e2e_steps}) + get_state_async(pinned)], per rank"},
            "gpu_launches": int(launches),
            "roofline": {"bound": max(fracs, key=fracs.get), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_gp_step": BYTES_PER_GP_STEP,
                         "kernel": prof.get("kernel", "k1d_pair_kernel<NonEq>"), "avg_launch_ms": avg_launch_ms,
                         "launch_ms": [round(x, 3) for x in launch_ms], "timed_region_ms": ms_total,
                         "hbm": {"achieved_gbs": achieved, "peak_gbs": peak, "frac": achieved / peak},
                         "fp32": {"achieved_tflops": fl * per_gpu_rate / 1e12, "peak_tflops": fp32_peak["flops_per_s"] / 1e12,
                                  "frac": fp32_frac, "flops_per_gp_step": fl, "flops_source": fl_src,
                                  "peak_source": "FFMA micro-benchmark in this run (kid_measure_fp32_peak)"},
                         "issue": issue,
                         "note": "top-level achieved/peak/frac are the HBM figures the contract asks for (algorithmic bytes, not DRAM traffic: "
                                 "several steps are fused per launch); `bound` names the largest of the hbm / fp32 / issue fractions"},
            "parity_sample": ({"columns": ncpu, "steps": args.steps, "vs": "CPU oracle from the same start state", **parity}
                              if parity else None),
            "cpu_baseline": cpu,
            "secondary": {"k2d": k2d_line},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--columns", type=int, default=1 << 20, help="columns of the ensemble (total; per GPU with --weak)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --columns per GPU instead of in total")
    ap.add_argument("--steps-per-launch", type=int, default=10)
    ap.add_argument("--t-start", type=int, default=480, help="model time [s] the pre-roll ends at (the timed window starts W steps later)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample length (0 = skip)")
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the in-bench parity sample")
    ap.add_argument("--e2e-chunks", type=int, default=32, help="column chunks (handles/streams) of the end-to-end pipelined call")
    ap.add_argument("--k2d-steps", type=int, default=200, help="steps of the secondary K2D measurement (0 = skip)")
    ap.add_argument("--kernel", default="pair", choices=["pair", "scalar"], help="step kernel of the resident leg (A/B measurements)")
    ap.add_argument("--direct-host", action="store_true", help="A/B: in-place PCIe access to the page-locked host arrays instead of copy engine + staging")
    ap.add_argument("--order", default="sorted", choices=["sorted", "hash"], help="storage order of the ensemble members")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
