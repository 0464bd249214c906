"""Experiment build (-DKID_EXP_TIMING): where do the warps of k1d_pair_kernel spend their cycles, per round of a stage?"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200 import lib as KL
from kid_b200.lib import Handle
t_start = int(sys.argv[1]) if len(sys.argv) > 1 else 480   # 0: nothing but clear air in every column (identical work in every warp)
cs = M.k1d_ensemble_case(np.float32, nc=262144, n_elem=128, ic="device", order="sorted", moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
h = M.make_handle(cs, Handle)
h.step(0.0, max(t_start, 1)); h.synchronize()
lib = KL.load_library()
buf = (C.c_ulonglong * 192)()
lib.kid_exp_pair_cycles(buf, 1)
h.set_steps_per_launch(10); h.step(float(max(t_start, 1)), 10); h.synchronize()
lib.kid_exp_pair_cycles(buf, 0)
a = np.array(list(buf)[:64], dtype=np.float64).reshape(8, 8)[:, :4]
w = np.array(list(buf)[64:128], dtype=np.float64).reshape(8, 8)[:, :4]
wt = np.array(list(buf)[128:], dtype=np.float64).reshape(8, 8)[:, :4]
tot = a.sum()
for name, row in zip(("wait at round barrier", "READ (stencil)", "sources", "activation + tendencies", "wait for the u^n box", "combine + stores", "u^n prefetch issue", "post (vt, cloud base)"), a):
    print(f"{name:24s}", " ".join(f"r{r}: {100*v/tot:5.1f}%" for r, v in enumerate(row)), f" | {100*row.sum()/tot:5.1f}%   mean cycles per warp and round: {row.sum() / (cs.cfg.nc / 16 * 8 * 30 * 4):7.0f}")
print("work (cycles outside the barrier wait) per warp and round, relative to the mean of the round:")
for i, row in enumerate(w):
    print(f"warp {i}:", " ".join(f"r{r}: {v / w[:, r].mean():5.2f}" for r, v in enumerate(row)), f" | stage total {row.sum() / w.sum(axis=1).mean():5.2f}")
print("round maximum / round mean:", " ".join(f"r{r}: {w[:, r].max() / w[:, r].mean():5.2f}" for r in range(4)), "| sum of round maxima / mean stage total:", f"{w.max(axis=0).sum() / w.sum(axis=1).mean():5.2f}")
n_wr = cs.cfg.nc / 16 * 30
print("mean cycles in the round-barrier wait per warp and round:")
for i, row in enumerate(wt):
    print(f"warp {i}:", " ".join(f"r{r}: {v / n_wr:6.0f}" for r, v in enumerate(row)), "  work:", " ".join(f"r{r}: {v / n_wr:6.0f}" for r, v in enumerate(w[i])))
