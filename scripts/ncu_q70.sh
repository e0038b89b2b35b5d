#!/bin/bash
# launch list (durations) of the 70-query search over a 125k-row shard and over 1M rows
set -e
python scripts/q70_probe.py 3 125000,1000000 only > gpurun_out/q70_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q70_launches.csv \
    python scripts/q70_probe.py 3 125000,1000000 only > gpurun_out/q70_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/q70_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
for r in rows[1:]:
    if "cir::" in r[ki] and "pack" not in r[ki]:
        print(r[ki].split("(")[0][:50], r[gi], r[vi])
PY
