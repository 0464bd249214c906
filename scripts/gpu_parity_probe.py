"""Development probe (GPU box): CUDA library vs CPU oracle on small cases; prints error tables."""
import sys, time, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200.lib import Handle, AUX_FIELDS
from oracle.oracle import OracleHandle

def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b)) if b.size else 0.0
    if scale == 0: return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / scale)

def run(name, cs, steps, t0=0.0, aux=True):
    g = M.make_handle(cs, Handle); o = M.make_handle(cs, OracleHandle)
    out = {"case": name}
    t = t0
    for n in steps:
        r_g = g.rhs(t); r_o = o.rhs(t)
        out[f"rhs@{t:g}"] = {v: relerr(r_g[:, i], r_o[:, i]) for i, v in enumerate(cs.var_names)}
        g.step(t, n); o.step(t, n); t += n * cs.cfg.dt
        Yg = g.get_state(); Yo = o.get_state()
        out[f"Y@{t:g}"] = {v: relerr(Yg[:, i], Yo[:, i]) for i, v in enumerate(cs.var_names)}
        if not np.all(np.isfinite(Yg)): out["nonfinite"] = True
    if aux:
        bad = {}
        for f in AUX_FIELDS:
            e = relerr(g.get_aux(t, f), o.get_aux(t, f))
            if e > 1e-5: bad[f] = e
        out["aux_bad"] = bad
    print(json.dumps(out, indent=None, default=float), flush=True)
    g.close(); o.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    for FT in (np.float64, np.float32):
        tag = "f64" if FT == np.float64 else "f32"
        for ms in ("EquilibriumMoisture", "NonEquilibriumMoisture"):
            for ps in ("NoPrecipitation", "Precipitation0M", "Precipitation1M", "Precipitation2M"):
                cs = M.k1d_case(FT, nc=3, moisture_choice=ms, precipitation_choice=ps, n_elem=64)
                t0 = time.time()
                run(f"k1d {tag} {ms[:5]} {ps[-2:]}", cs, [50, 250, 300])
                print("   time", time.time() - t0, flush=True)
        for ps, rf in (("Precipitation1M", "CliMA_1M"), ("Precipitation1M", "KK2000"), ("Precipitation2M", "SB2006")):
            cs = M.box_case(FT, nc=5, precipitation_choice=ps, rain_formation_choice=rf, qt=np.linspace(5e-4, 2e-3, 5), Nd=1e8)
            run(f"box {tag} {ps[-2:]} {rf}", cs, [100, 500])
