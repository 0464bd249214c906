import sys, time
sys.path.insert(0, "kinematicdriver.jl_b200"); sys.path.insert(0, ".")
import numpy as np
from kid_b200 import model as M
from kid_b200.lib import Handle
for FT in (np.float64, np.float32):
    for mo, pr in (("EquilibriumMoisture", "Precipitation0M"), ("NonEquilibriumMoisture", "Precipitation2M")):
        for nc in (1, 4):
            cs = M.k1d_case(FT, nc=nc, moisture_choice=mo, precipitation_choice=pr)
            h = M.make_handle(cs, Handle); h.step(0.0, 200); h.synchronize()
            t = time.perf_counter(); h.step(200.0, 3400); h.synchronize(); dt = time.perf_counter() - t
            print(FT.__name__, mo[:5], pr[-2:], "nc", nc, f"{dt/3400*1e6:.1f} us/step", np.isfinite(h.get_state()).all())
