#!/bin/bash
# ncu --set full capture of one kernel inside a bench run.  bash tools/gpu_ncu.sh <tag> <kernel-regex> [launch-skip]
TAG=$1; K=$2; SKIP=${3:-3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP -c 1 -f -o $OUT/full_$K \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$K.log 2>&1
ls -la $OUT
