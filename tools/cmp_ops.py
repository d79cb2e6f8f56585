"""Compare per-op tables written by `bench.py --ops` (columns: avg microseconds per launch and TFLOP/s)."""
import csv
import json
import sys

names = sys.argv[1:]
tabs = [{(r["op"], r["tag"]): r for r in csv.DictReader(open(f"gpurun_out/ops_{n}.csv"))} for n in names]
for n in names:
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(f"{n:12s} ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['ms_per_step']:.2f}  value {d['value']:.0f}  clocks {d['clocks']}")
    except Exception as e:  # noqa: BLE001
        print(n, "no bench json", e)
tot = [0.0] * len(names)
keys = list(tabs[0])
for t in tabs[1:]:
    keys += [k for k in t if k not in keys]
for k in keys:
    cells = []
    for i, t in enumerate(tabs):
        r = t.get(k)
        if r:
            tot[i] += float(r["total_ms"])
            cells.append(f"{float(r['avg_ms']) * 1e3:8.1f} {float(r['tflops']):7.1f}")
        else:
            cells.append(" " * 16)
    n = tabs[0].get(k, {}).get("launches") or next(t[k]["launches"] for t in tabs if k in t)
    print(f"{k[0][4:]:22s} {k[1][:34]:36s} {n:>3s} " + " | ".join(cells))
print("total ms", [round(x, 2) for x in tot])
