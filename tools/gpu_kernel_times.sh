#!/bin/bash
# Cheap per-kernel timing of ONE kernel family inside a bench run: ncu only instruments the launches whose name matches.
# bash tools/gpu_kernel_times.sh <tag> <kernel-regex> [max-launches]
TAG=$1; K=$2; N=${3:-200}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K -c $N --csv --log-file $OUT/times_$K.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_times_$K.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/times_$K.csv")) if len(r) > 10]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ki][:60], []).append(float(r[vi].replace(",", "")))
for k, v in per.items():
    print(f"{k:60s} n={len(v):3d} last={v[-1] / 1e3:9.1f} us  min={min(v) / 1e3:9.1f} us")
PY
