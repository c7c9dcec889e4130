"""Prints the hottest SASS lines (stall samples) of an ncu report: python scripts/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci, cs, ce = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
names = ["stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_branch_resolving", "stall_dispatch", "stall_no_inst"]
cols = [hdr.index(n) for n in names]
cw = hdr.index("L1 Wavefronts Shared Excessive")
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((float(r[cs]), r[ci].strip(), int(r[ce]), [int(r[c]) for c in cols], i, int(r[cw])))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("samples", tot, "warp-instr", sum(d[2] for d in data), "sass lines", len(data))
agg = [sum(d[3][j] for d in data) for j in range(len(names))]
print("stall mix:", ", ".join(f"{n[6:]}={a / tot * 100:.1f}%" for n, a in zip(names, agg)))
print("excessive smem wavefronts:", sum(d[5] for d in data))
for d in sorted(data, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = max(range(len(names)), key=lambda j: d[3][j])
    print(f"{d[0] / tot * 100:5.1f}% inst={d[2]:>10d} {names[top][6:]:>14s} exc={d[5]:8d} #{d[4]:4d} {d[1][:90]}")
