"""Extracts the PySDM golden initial profiles the reference ships as Julia array literals
(/root/reference/test/data_utils.jl, the only golden vectors in the reference: SURVEY §4)
into tests/golden/pysdm_profiles.json.  Run in the build container (the reference checkout
does not exist on the GPU box); the JSON is committed."""
import json
import re
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/data_utils.jl").read_text()
out = {}
# split into the four `case == "..."` branches
parts = re.split(r'case\s*==\s*"(\w+)"', src)
for name, body in zip(parts[1::2], parts[2::2]):
    case = {}
    for var, arr in re.findall(r"(\w+_sdm)\s*=\s*\[(.*?)\]", body, flags=re.S):
        vals = [float(x) for x in re.split(r"[,\s]+", arr.strip()) if x]
        case[var] = vals
    out[name] = case
    print(name, {k: len(v) for k, v in case.items()})
dst = Path(__file__).resolve().parent / "pysdm_profiles.json"
dst.write_text(json.dumps(out))
print("wrote", dst, dst.stat().st_size, "bytes")
