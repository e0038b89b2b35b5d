#!/bin/bash
# launch list of the full ranking at 70 x 1M (scripts/bench_round2.py rank)
python scripts/bench_round2.py rank > gpurun_out/rank_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rank_launches.csv \
    python scripts/bench_round2.py rank > gpurun_out/rank_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/rank_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
seen=0
for r in rows[1:]:
    if "cir::radix" in r[ki] or "sort" in r[ki]:
        print(r[ki].split("(")[0][:60], r[gi], r[vi]); seen+=1
        if seen>=40: break
PY
