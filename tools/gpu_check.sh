#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, per-op timings, ncu launch list + full captures of the top kernels.
# Usage (from the repo root, on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 1500 $OUT/bench_ref.json
timeout 600 python tools/op_times.py > $OUT/op_times.txt 2>&1; tail -40 $OUT/op_times.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
for K in decode_lattice_kernel conv_tc_kernel decode_tc_kernel linear_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 -c 1 -f -o $OUT/full_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$K.log 2>&1
done
ls -la $OUT
