"""Per-instruction stall samples of one kernel from an .ncu-rep (source page).  Usage: ncu_source.py rep [kernel-regex] [launch#] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else "k_xdot"; nth = sys.argv[3] if len(sys.argv) > 3 else "1"; ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{nth}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if r and r[0] in ('Address', 'Kernel Name'): break
    if len(r) == len(hdr): data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k[6:]: sum(int(r[ix[k]]) for r in data) for k in stalls}
print("total samples", tot, "instructions", len(data))
print("stall totals", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:ntop]
for i in sorted(top):
    r = data[i]
    s = {k[6:]: int(r[ix[k]]) for k in stalls if int(r[ix[k]])}
    print(f"{i:5d} {r[ix['Source']].strip()[:64]:64s} smp {r[ix['# Samples']]:>5s} exe {r[ix['Instructions Executed']]:>8s}", s)
# opcode histogram weighted by executions
ops = {}
for r in data:
    op = r[ix['Source']].strip().split()
    op = (op[1] if op and op[0].startswith('@') else op[0]) if op else '?'
    op = op.split('.')[0]
    ops[op] = ops.get(op, 0) + int(r[ix['Instructions Executed']])
te = sum(ops.values())
print("executed warp instructions", te, {k: round(v / te, 4) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]})
