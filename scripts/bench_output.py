"""Cost of the output path at the drivers' default cadence (every 2nd step, condition_io of netcdf_io.jl:176-185):
no output / one launch per field (kid_get_aux) / one launch for all fields (kid_get_aux_many) / asynchronous
double-buffered download (kid_output_enqueue).  Prints one JSON line per mode."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kinematicdriver.jl_b200")); sys.path.insert(0, str(ROOT))
from kid_b200 import model as M
from kid_b200.lib import Handle, AUX_FIELDS
from kid_b200.solve import OutputPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=16384)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--every", type=int, default=2)
a = ap.parse_args()
fields = [f for f in AUX_FIELDS]
cs = M.k1d_ensemble_case(np.float32, nc=a.columns, n_profiles=16, n_elem=128, moisture_choice="NonEquilibriumMoisture",
                         precipitation_choice="Precipitation2M", order="sorted")
from kid_b200.lib import pinned_empty
buf = pinned_empty((a.columns, len(fields), 128), np.float32)
for mode in ("none", "per_field", "many", "many_pinned", "pipeline"):
    h = M.make_handle(cs, Handle)
    h.step(0.0, 480); h.synchronize()
    pipe = OutputPipeline(h, fields, sink=lambda t, out: None) if mode == "pipeline" else None
    t = 480.0
    t0 = time.perf_counter()
    for i in range(a.steps // a.every):
        h.step(t, a.every); t += a.every
        if mode == "per_field":
            for f in fields: h.get_aux(t, f)
        elif mode == "many":
            h.get_aux_many(t, fields)
        elif mode == "many_pinned":
            h.get_aux_many(t, fields, out=buf)
        elif mode == "pipeline":
            pipe.enqueue(t)
    if pipe: pipe.drain()
    h.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"mode": mode, "columns": a.columns, "levels": 128, "fields": len(fields), "output_every_steps": a.every,
                      "steps": a.steps, "ms_per_step": dt * 1e3 / a.steps,
                      "output_bytes_per_output": a.columns * 128 * len(fields) * 4}))
