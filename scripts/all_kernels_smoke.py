"""Small instance of every kernel family (quick functional check on a GPU box): K1D tile (fixed and generic shape, f32/f64, per-column
parameter sets), Box, K2D, observables, multi-field output, asynchronous transfers."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M, k2d as K2, calibration as CAL
from kid_b200.lib import Handle, AUX_FIELDS, pinned_empty
from kid_b200.solve import OutputPipeline, PipelinedEnsemble

def run(case, steps=6, t0=0.0):
    h = M.make_handle(case, Handle); h.step(t0, steps); Y = h.get_state(); assert np.isfinite(Y).all(); return h

kw = dict(moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
h = run(M.k1d_ensemble_case(np.float32, nc=70, n_profiles=3, n_elem=128, **kw))            # fixed-shape tile, ragged last tile
h.get_aux_many(6.0, AUX_FIELDS); h.rhs(6.0); h.reduce_max(6.0, "q_liq")
run(M.k1d_ensemble_case(np.float32, nc=40, n_profiles=3, n_elem=50, **kw))                   # generic shape, partial last round
run(M.k1d_ensemble_case(np.float64, nc=20, n_profiles=2, n_elem=128, **kw))
run(M.k1d_ensemble_case(np.float32, nc=33, n_profiles=2, n_elem=32, moisture_choice="EquilibriumMoisture", precipitation_choice="Precipitation1M"))
run(M.k1d_case(np.float64, nc=3, n_elem=16, moisture_choice="EquilibriumMoisture", precipitation_choice="Precipitation0M"))
cs = M.k1d_ensemble_case(np.float32, nc=48, n_profiles=2, n_elem=32, **kw)
run(M.with_parameter_sets(cs, [{"SB2006_collection_kernel_coeff_kcc": 4.44e9 * f} for f in (0.8, 1.0, 1.2)]))
run(M.box_ensemble_case(np.float64, nc=100))
run(M.box_ensemble_case(np.float32, nc=100))
c2 = K2.k2d_case(np.float32, nx=6, nz=16, npoly=3, **kw)
run(c2, steps=4)
G = CAL.solve_gvector(M.k1d_ensemble_case(np.float32, nc=9, n_profiles=2, n_elem=32, **kw), ["rho", "ql", "rainrate"], [4.0, 8.0])
assert np.isfinite(G).all()
pipe = OutputPipeline(h, ["q_liq", "q_rai"]); pipe.enqueue(6.0); h.step(6.0, 2); pipe.enqueue(8.0); pipe.drain()
cs = M.k1d_ensemble_case(np.float32, nc=96, n_profiles=2, n_elem=32, **kw)
Yp = pinned_empty(cs.Y0.shape, np.float32); Yp[...] = cs.Y0
PipelinedEnsemble(cs, nchunks=3).solve(Yp, 0.0, 4); assert np.isfinite(Yp).all()
print("all kernel families ok")
