#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric: grid-point-steps/s).

Workload (N = 1): BASELINE.json configs[3] — KiD 1-D NonEquilibriumMoisture + Precipitation2M (SB2006)
ensemble of 1,048,576 columns x 128 levels, Float32, synthetic calibration-style ensemble of
SURVEY §8(d) (per-column w1, prescribed Nd, surface pressure -> per-column ρ_dry/T profiles).
N > 1: every rank steps its own 1,048,576-column shard (weak scaling, no data-path collective).

One "step" = one SSPRK33 step (3 rhs evaluations) of every column of the shard.

    python bench.py --gpus N --steps K --warmup W          # ours  (libkid_cuda.so, sm_100a)
    python bench.py --impl reference ...                   # the reference's CPU path on the host cores

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the stream the kernels are
launched on, state resident in HBM; `e2e` goes through the public host API with pinned HOST
buffers (upload + K steps + download inside the timed region).
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
ORDER = "sorted"
IC = "device"
PREROLL_SPL = 60  # the untimed pre-roll runs in launches of 60 fused steps


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def build_case(nc: int, column_offset: int, device: int, host_ic: bool = False):
    """SURVEY §8(d) ensemble: member c has w1, Nd, p0 from the counter hash.  --ic device (default): every member its own
    p0, hydrostatic profile and initial state built on the device (kid_init_shipway_hill); --ic host: p0 quantised to 16
    soundings solved on the host.  Member order: "sorted" stores the members along a Z-order curve through
    (w1, log Nd, p0) (--ic host: lexicographically by sounding, w1, Nd), so that neighbouring columns — the 32 lanes of a
    warp — follow the same microphysics branches; --order hash keeps the hash order of §8(d)."""
    from kid_b200 import model as M
    kw = dict(column_offset=column_offset, device=device, n_elem=NZ, order=ORDER,
              moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    if IC == "device" and not host_ic:
        return M.k1d_ensemble_case(np.float32, nc=nc, ic="device", **kw)
    # host IC; the CPU arm of a device-IC run keeps the exact per-member p0 (one host IVP per member of its small sample)
    return M.k1d_ensemble_case(np.float32, nc=nc, n_profiles=None if IC == "device" else 16, **kw)


def cpu_reference_throughput(case, Y_dev_sample, t_start: float, target_s: float = 15.0):
    """The reference's CPU path (restated: oracle/kid_oracle.cpp, OpenMP over columns) on a bounded
    sample of the same workload: the first `ncpu` columns of the developed ensemble state."""
    from kid_b200 import model as M
    from oracle.oracle import OracleHandle
    cores = os.cpu_count() or 1
    ncpu = Y_dev_sample.shape[0]
    import copy
    sub = copy.copy(M.materialise(M.slice_columns(case, 0, ncpu)))  # device-built profiles of the sample -> host
    sub.Y0 = Y_dev_sample
    o = M.make_handle(sub, OracleHandle)
    t0 = time.perf_counter()
    o.step(t_start, 4)
    per_step = (time.perf_counter() - t0) / 4
    nsteps = int(max(8, min(2000, target_s / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    o.step(t_start + 4, nsteps)
    dt = time.perf_counter() - t0
    return {"value": ncpu * NZ * nsteps / dt, "unit": METRIC, "cores": cores, "kind": "port",
            "sample": f"{ncpu} columns x {NZ} levels x {nsteps} SSPRK33 steps of the same developed ensemble state "
                      f"(t={t_start:g}s), OpenMP over columns, g++ -O2, Float32; restated CPU path, not Julia"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Julia and the reference's
    un-vendored dependencies are absent (no oracle/_ref can be built), so this arm times the restated CPU
    path (oracle/) with all host threads on a bounded sample of the arm's workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from kid_b200 import model as M
    from oracle.oracle import OracleHandle
    cores = os.cpu_count() or 1
    ncpu = 32 * cores
    case = build_case(ncpu, 0, 0, host_ic=True)
    o = M.make_handle(case, OracleHandle)
    t_start = float(args.t_start)
    # bring the sample to the same developed state the device arm times (untimed)
    o.step(0.0, int(t_start))
    for _ in range(args.warmup):
        o.step(t_start, 1)
        t_start += 1
    t0 = time.perf_counter()
    o.step(t_start, args.steps)
    dt = time.perf_counter() - t0
    value = ncpu * NZ * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "KiD 1D NonEq+Precipitation2M(SB2006) ensemble, 128 levels, Float32 (BASELINE configs[3])",
                   "columns_per_step": ncpu, "t_start_s": args.t_start},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "port",
                         "sample": f"{ncpu} columns x {NZ} levels x {args.steps} steps; restated CPU path (oracle/), "
                                   "OpenMP over columns; Julia + CloudMicrophysics/ClimaCore are not installable here"},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    from kid_b200 import model as M
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
        from kid_b200.lib import bind_to_gpu_numa_node
        pr = torch.cuda.get_device_properties(local)
        numa = bind_to_gpu_numa_node(f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        import torch.distributed as dist
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nc = args.columns
    case = build_case(nc, rank * nc, local)
    h = M.make_handle(case, Handle)
    spl = args.steps_per_launch
    stream = torch.cuda.ExternalStream(h.stream, device=local)

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # untimed pre-roll into the developed state (updraft on: activation, condensation, collisions, advection)
    t = 0.0
    h.set_steps_per_launch(PREROLL_SPL)
    h.step(t, int(args.t_start))
    h.set_steps_per_launch(spl)
    t = float(args.t_start)
    for _ in range(args.warmup):
        h.step(t, spl)
        t += spl
    barrier()

    # ---------------- device-resident timing (value)
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
    gp_steps_total = world * nc * NZ * args.steps
    value = gp_steps_total / (ms_max * 1e-3)

    # ---------------- end-to-end through the host API with pinned host buffers (e2e)
    # The user-facing call for a large ensemble: the columns go through `nchunks` handles (one CUDA stream each), so the
    # H2D upload, the K steps and the D2H download of different chunks overlap (kid_set_state_async / kid_step /
    # kid_get_state_async).  Everything — copies included — is inside the timed region.
    from kid_b200.solve import PipelinedEnsemble
    state_shape = (nc, len(case.var_names), NZ)
    Yh = torch.empty(state_shape, dtype=torch.float32).pin_memory()
    Yn = Yh.numpy()
    h.get_state(Yn)  # developed state as the "user's" host input
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
    state_bytes = Yn.nbytes

    finite = bool(np.isfinite(Yn[:: max(1, nc // 4096)]).all())

    if rank == 0:
        peak, peak_src = measured_peaks()
        avg_launch_ms = statistics.mean(launch_ms)
        steps_per = args.steps / n_launch
        achieved = BYTES_PER_GP_STEP * nc * NZ * steps_per / (avg_launch_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 or rank == 0:
            ncpu = min(nc, 32 * (os.cpu_count() or 1))
            cpu = cpu_reference_throughput(case, Yn[:ncpu].copy(), t, target_s=args.cpu_seconds) if args.cpu_seconds > 0 else None
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch_per_step")
                traffic = traffic * steps_per if traffic else None
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "KiD 1D NonEq+Precipitation2M(SB2006) ensemble, 128 levels, Float32 (BASELINE configs[3])",
                       "columns_per_gpu": nc, "levels": NZ, "steps_per_launch": spl, "t_start_s": t_timed0,
                       "member_order": ("hash (SURVEY 8d counter hash)" if ORDER == "hash" else
                                        "sorted along a Z-order curve through (w1, log Nd, p0)" if IC == "device" else
                                        "sorted by (sounding, w1, Nd)"),
                       "initial_condition": ("built on the device, every member its own p0 (SURVEY 8d)" if IC == "device" else
                                             "host, p0 quantised to 16 soundings"),
                       "l2": "state per GPU (%.2f GiB) is far larger than the 126 MB L2; no flush needed" % (state_bytes / 2**30),
                       "parallelism": f"columns sharded x{world}, no collective", "state_finite": finite,
                       **({"host_numa_binding": numa} if numa else {})},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": state_bytes / e2e_steps,
                    "d2h_bytes_per_step": state_bytes / e2e_steps,
                    "call": f"PipelinedEnsemble.solve: {args.e2e_chunks} column chunks x [set_state_async(pinned) + step({e2e_steps}) + get_state_async(pinned)], per rank"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_gp_step": BYTES_PER_GP_STEP,
                         "kernel": "k1d_tile_kernel<float,NonEq,2M>", "avg_launch_ms": avg_launch_ms,
                         "launch_ms": [round(x, 3) for x in launch_ms], "timed_region_ms": ms_total,
                         "note": "the fused kernel is FP32/MUFU-pipe bound, not HBM bound: see profiles/ for pipe utilisation"},
            "cpu_baseline": cpu,
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
    ap.add_argument("--columns", type=int, default=1 << 20, help="columns per GPU")
    ap.add_argument("--steps-per-launch", type=int, default=10)
    ap.add_argument("--t-start", type=int, default=480, help="model time [s] at which the timed window starts")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample length (0 = skip)")
    ap.add_argument("--e2e-chunks", type=int, default=32, help="column chunks (handles/streams) of the end-to-end pipelined call")
    ap.add_argument("--ic", default="device", choices=["device", "host"], help="where the initial condition is built")
    ap.add_argument("--order", default="sorted", choices=["sorted", "hash"], help="storage order of the ensemble members")
    args = ap.parse_args()
    global ORDER, IC
    ORDER = args.order
    IC = args.ic
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
