"""SASS opcode histogram, per-source-line instruction counts and warp-stall samples of one ncu report (captured with
--import-source on, compiled with -lineinfo):  python scripts/ncu_sass_hist.py rep.ncu-rep [top] > profiles/....txt
The source page of ncu lists every SASS instruction under the CUDA line it was generated from; with inlining an instruction is
listed under every line of its inline stack, so the per-line tables are INCLUSIVE; the opcode histogram and the stall totals
count every SASS address once."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


hdr = None
cur_file = None
cur_line = None
by_op = collections.Counter()
by_line = collections.Counter()
samp_line = collections.Counter()
samp_op = collections.Counter()
stall_tot = collections.Counter()
stall_line = collections.defaultdict(collections.Counter)
func = None
seen = set()
by_func = collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "Function Name" and len(r) > 1:
        func = r[1][:60]
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = r.index("Instructions Executed")
        isamp = r.index("# Samples")
        stall_cols = [(i, n) for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or len(r) <= ie:
        continue
    if r[0]:
        cur_line = (cur_file, int(r[0]), r[1].strip())
    if not r[2] or r[2] in ("-", "..."):
        continue  # CUDA line row (aggregate of its SASS rows) / elision marker
    first = r[2] not in seen  # with inlining ncu lists an instruction under every line of its inline stack
    seen.add(r[2])
    sass = r[3].split()
    op = sass[1] if sass and sass[0].startswith("@") and len(sass) > 1 else (sass[0] if sass else "?")
    op = op.split(".")[0]
    n = num(r[ie])
    s = num(r[isamp])
    by_line[cur_line] += n      # inclusive: a call-site line carries the instructions inlined into it
    samp_line[cur_line] += s
    for i, name in stall_cols:
        v = num(r[i])
        if v:
            stall_line[cur_line][name] += v
            if first:
                stall_tot[name] += v
    if first:
        by_op[op] += n
        samp_op[op] += s
        by_func[func] += n

tot = sum(by_op.values())
stot = sum(samp_op.values())
print(f"report: {rep}\ntotal warp instructions executed: {tot:.4g}; stall samples: {stot:.0f}")
print("\n== SASS opcode histogram (share of executed warp instructions | share of stall samples)")
for op, n in by_op.most_common(45):
    print(f"{100*n/tot:6.2f}%  {100*samp_op[op]/max(stot,1):6.2f}%  {op}")
print("\n== stall reasons (share of samples)")
st = sum(stall_tot.values())
for k, v in stall_tot.most_common():
    print(f"{100*v/max(st,1):6.2f}%  {k}")
print(f"\n== top {top} source lines by stall samples (samples% | instructions% | top stall reasons)")
for k, v in samp_line.most_common(top):
    rs = ", ".join(f"{n[6:]} {100*c/max(v,1):.0f}%" for n, c in stall_line[k].most_common(3))
    print(f"{100*v/max(stot,1):5.2f}% {100*by_line[k]/tot:5.2f}%  {k[0]}:{k[1]:<4} {k[2][:90]}   [{rs}]")
print(f"\n== top {top} source lines by executed instructions")
for k, v in by_line.most_common(top):
    print(f"{100*v/tot:5.2f}%  {k[0]}:{k[1]:<4} {k[2][:110]}")
