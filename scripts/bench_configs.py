"""Throughput of every BASELINE.json config on one B200 (device-resident state, wall clock around
step + synchronize, after a warm-up), with the CPU oracle beside it on a bounded sample.
Writes gpurun_out/configs.json."""
import json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M, k2d
from kid_b200.lib import Handle
from oracle.oracle import OracleHandle
import copy

def gpu_rate(case, t0, nsteps, warm=20, spl=0):
    h = M.make_handle(case, Handle)
    if spl: h.set_steps_per_launch(spl)
    h.step(0.0, int(t0) + warm); h.synchronize()
    t = time.perf_counter(); h.step(t0 + warm, nsteps); h.synchronize(); dt = time.perf_counter() - t
    Y = h.get_state()
    return case.cfg.nc * case.cfg.nz * nsteps / dt, dt / nsteps * 1e3, bool(np.isfinite(Y).all())

def cpu_rate(case, ncpu, t0, nsteps):
    sub = copy.copy(case); sub.cfg = copy.copy(case.cfg); sub.cfg.nc = ncpu
    pc = case.ρ_dry.ndim == 2
    sub.ρ_dry = case.ρ_dry[:ncpu] if pc else case.ρ_dry; sub.T = case.T[:ncpu] if pc else case.T
    sub.Y0 = case.Y0[:ncpu]; sub.col_scalars = {k: np.asarray(v)[:ncpu] for k, v in case.col_scalars.items()}
    o = M.make_handle(sub, OracleHandle)
    o.step(0.0, int(t0))
    t = time.perf_counter(); o.step(t0, nsteps); dt = time.perf_counter() - t
    return ncpu * case.cfg.nz * nsteps / dt

out = []
cores = os.cpu_count()
def rec(name, case, t0, nsteps, ncpu, cpu_steps, **kw):
    g, ms, fin = gpu_rate(case, t0, nsteps, **kw)
    c = cpu_rate(case, ncpu, t0, cpu_steps)
    r = dict(config=name, gp_steps_per_s=g, ms_per_step=ms, finite=fin, cpu_gp_steps_per_s=c, cpu_cores=cores if ncpu > 1 else 1,
             cpu_sample=f"{ncpu} columns x {cpu_steps} steps", speedup=g / c, columns=case.cfg.nc, levels=case.cfg.nz)
    out.append(r); print(json.dumps(r), flush=True)

# 1. KiD 1D Eq + 0M, Float64, default grid, ONE column (the reference's own CPU-runnable case)
cs = M.k1d_case(np.float64, nc=1, moisture_choice="EquilibriumMoisture", precipitation_choice="Precipitation0M")
rec("1: KiD 1D Eq+0M f64 128 levels, 1 column", cs, 0, 3580, 1, 3600)
# 2. Box ensemble of 1M boxes, Precipitation1M
for FT, tag in ((np.float64, "f64"), (np.float32, "f32")):
    cs = M.box_ensemble_case(FT, nc=1 << 20, n_ic=32, precipitation_choice="Precipitation1M", rain_formation_choice="CliMA_1M")
    rec(f"2: Box 1M boxes Precipitation1M {tag}", cs, 0, 600, 4096 * 4, 600)
cs = M.box_ensemble_case(np.float64, nc=1 << 20, n_ic=32, precipitation_choice="Precipitation2M", rain_formation_choice="SB2006")
rec("2b: Box 1M boxes Precipitation2M f64", cs, 0, 600, 4096 * 4, 600)
# 3. KiD NonEq + 1M, 65,536 columns
for FT, tag in ((np.float64, "f64"), (np.float32, "f32")):
    cs = M.k1d_ensemble_case(FT, nc=65536, n_profiles=16, order="sorted", moisture_choice="NonEquilibriumMoisture",
                             precipitation_choice="Precipitation1M", w1=1.5)
    rec(f"3: KiD NonEq+1M 65536 columns x 128 {tag}", cs, 300, 40, 32 * cores, 40, spl=4)
# 4b. the headline style in Float64 for reference
cs = M.k1d_ensemble_case(np.float64, nc=65536, n_profiles=16, order="sorted", moisture_choice="NonEquilibriumMoisture",
                         precipitation_choice="Precipitation2M")
rec("4b: KiD NonEq+2M 65536 columns x 128 f64", cs, 480, 20, 32 * cores, 20, spl=4)
# 5. K2D 4096 x 256, Precipitation2M
for FT, tag in ((np.float32, "f32"), (np.float64, "f64")):
    cs = k2d.k2d_case(FT, nx=1024, nz=256, npoly=3, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                      domain_width=48000.0, domain_height=3000.0)
    rec(f"5: K2D 4096 nodes x 256 levels 2M {tag}", cs, 0, 100, 4096, 6, warm=20)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "configs.json", "w"), indent=1)
