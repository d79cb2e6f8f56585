#!/bin/bash
# DRAM bytes of the launches of one call IN SITU (ncu --cache-control none: the L2 contents a launch inherits from its
# predecessors are kept), config-2 shape, NFE 2: what each launch really reads from / writes to HBM inside the step
mkdir -p gpurun_out
export SRB_GRAPHS=0
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -s 620 -c 700 --csv --log-file gpurun_out/r02_ncu_insitu.csv python tools/ncu_hbm_kernels.py 2 > gpurun_out/r02_ncu_insitu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_ncu_insitu.csv")) if len(r) > 10 and r[0].isdigit()]
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r[0], {"name": r[4][:70], "grid": r[8]})
    d[r[12]] = (float(r[14].replace(",", "")), r[13])
print(len(byid), "launches")
def mb(v):
    val, unit = v
    return val * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
for i, d in list(byid.items())[:140]:
    print(i, d["name"][:64].ljust(64), d["grid"].rjust(8), "rd %8.1f MB  wr %8.1f MB  %8.1f us" % (
        mb(d["dram__bytes_read.sum"]), mb(d["dram__bytes_write.sum"]), d["gpu__time_duration.sum"][0] * {"us": 1, "ms": 1e3, "ns": 1e-3}[d["gpu__time_duration.sum"][1]]))
PY
