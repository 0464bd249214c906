"""Aggregate an ncu report's executed (warp) instructions per CUDA source line: `python scripts/ncu_lines.py rep.ncu-rep [N]`."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def f(x):
    try: return float(x)
    except ValueError: return 0.0
cur = hdr = line = None; agg = {}; ops = {}
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); continue
    if hdr and len(r) > ie:
        if r[0]: line = (cur, int(r[0]), r[1])
        if r[2]:
            agg[line] = agg.get(line, 0) + f(r[ie])
            op = r[3].split()[0] if r[3].split() else "?"
            if op.startswith("@"): op = r[3].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + f(r[ie])
tot = sum(agg.values())
print("total warp instructions", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v/tot*100:5.2f}% {k[0]}:{k[1]:>5} {k[2].strip()[:130]}")
print("---- by opcode")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v/tot*100:5.2f}% {k}")
import pickle; pickle.dump(agg, open("/tmp/agg.pkl", "wb"))
