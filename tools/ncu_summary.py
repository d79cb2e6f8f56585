"""Condense ncu reports (--set full) into one CSV row per profiled launch with the columns DESIGN.md quotes.

    python tools/ncu_summary.py "note for the header line" gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/x.csv
"""
import csv
import io
import os
import subprocess
import sys

COLS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "smsp__issue_active.avg.per_cycle_active",
]


def rows_of(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    head, units, body = rd[0], rd[1], rd[2:]
    ki = head.index("Kernel Name")
    idx = [head.index(c) if c in head else None for c in COLS]
    tag = os.path.splitext(os.path.basename(rep))[0]
    for r in body:
        vals = []
        for i in idx:
            if i is None:
                vals.append("")
            else:
                u = units[i]
                vals.append(f"{r[i]} {u}".strip())
        yield [tag, r[ki][:60]] + vals


def main():
    note, reps = sys.argv[1], sys.argv[2:]
    w = csv.writer(sys.stdout, quoting=csv.QUOTE_MINIMAL)
    print("# " + note)
    w.writerow(["report", "kernel"] + COLS)
    for rep in reps:
        for row in rows_of(rep):
            w.writerow(row)


if __name__ == "__main__":
    main()
