"""Which member ordering keeps the 32 lanes of a warp on the same branches?  Kernel rate of one pipeline chunk (hash-ordered
members, 32,768 columns by default) stored in different orders."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200.lib import Handle
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cs = M.k1d_ensemble_case(np.float32, nc=nc, n_elem=128, ic="device", order="hash", moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
w1 = np.asarray(cs.col_scalars[M.COL_W1], dtype=np.float64); nd = np.log10(np.asarray(cs.col_scalars[M.COL_PRESCRIBED_ND], dtype=np.float64)); p0 = np.asarray(cs.device_ic["p0"], dtype=np.float64)
def rank01(x): return np.argsort(np.argsort(x)) / x.size
def bins(x, n): return np.minimum((rank01(x) * n).astype(int), n - 1)
orders = {
    "hash": np.arange(nc),
    "morton(w1,Nd,p0)": M.morton_order(w1, nd, p0),
    "p0 x8 | morton(w1,Nd)": np.lexsort((np.argsort(np.argsort(M.morton_order(w1, nd))) * 0 + M.inverse_permutation(M.morton_order(w1, nd)), bins(p0, 8))),
    "p0 x32 | morton(w1,Nd)": np.lexsort((M.inverse_permutation(M.morton_order(w1, nd)), bins(p0, 32))),
    "w1 x32 | morton(Nd,p0)": np.lexsort((M.inverse_permutation(M.morton_order(nd, p0)), bins(w1, 32))),
    "Nd x32 | morton(w1,p0)": np.lexsort((M.inverse_permutation(M.morton_order(w1, p0)), bins(nd, 32))),
    "lex(p0 x32, w1 x32, Nd)": np.lexsort((nd, bins(w1, 32), bins(p0, 32))),
    "lex(w1 x32, p0 x32, Nd)": np.lexsort((nd, bins(p0, 32), bins(w1, 32))),
    "morton(p0,p0',w1,Nd)": M.morton_order(w1, nd, p0, bits=7),
}
if nc >= (1 << 19):
    orders = {"morton(w1,Nd,p0)": orders["morton(w1,Nd,p0)"]}
    for b1, b2 in ((32, 32), (64, 64), (128, 32), (32, 128), (256, 64)):
        orders[f"lex(p0 x{b1}, w1 x{b2}, Nd)"] = np.lexsort((nd, bins(w1, b2), bins(p0, b1)))
    orders["p0 x64 | morton(w1,Nd)"] = np.lexsort((M.inverse_permutation(M.morton_order(w1, nd)), bins(p0, 64)))
    orders["p0 x256 | morton(w1,Nd)"] = np.lexsort((M.inverse_permutation(M.morton_order(w1, nd)), bins(p0, 256)))
for name, perm in orders.items():
    h = M.make_handle(M.reorder_columns(cs, np.asarray(perm)), Handle)
    h.step(0.0, 480); h.synchronize(); h.set_steps_per_launch(10)
    t0 = time.perf_counter(); h.step(480.0, 30); h.synchronize(); dt = time.perf_counter() - t0
    print(f"{name:28s} {dt*1e3/30:.3f} ms/step  {nc*128*30/dt:.3e}", flush=True)
    del h
