"""Summarise an ncu report's source page: stall-reason totals and the hottest SASS instructions per kernel.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [kernel-regex] [top-n]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kre:
    cmd += ["--kernel-name-base", "demangled", "--kernel-name", f"regex:{kre}"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "header": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["header"] is None:
        cur["header"] = row
    elif cur is not None:
        cur["rows"].append(row)
for b in blocks:
    h = b["header"]
    print("=" * 100)
    print(b["name"][:140])
    si = h.index("# Samples")
    src = h.index("Source")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print("total samples", tot, " instructions", len(b["rows"]))
    sums = {h[i]: sum(int(r[i] or 0) for r in b["rows"]) for i in stall_cols}
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(sums.items(), key=lambda kv: -kv[1]) if v))
    ranked = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si] or 0))[:topn]
    for i in sorted(ranked):
        r = b["rows"][i]
        st = sorted(((int(r[c] or 0), h[c][6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {int(r[si] or 0):6d} {r[src].strip()[:70]:70s} " + " ".join(f"{n}:{v}" for v, n in st if v))
