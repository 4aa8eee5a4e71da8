#!/bin/bash
# Profiles of one round: bench (both arms), the ncu launch list of ONE timed pass (NVTX range gnb.timed_step), and
# `ncu --set full` captures of the kernels the review names.  bash tools/gpu_profile.sh [tag]
TAG=${1:-profile}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 900 ncu --nvtx --nvtx-include "gnb.timed_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
# full captures, one launch per kernel family inside the timed pass (the largest launch of each family comes first in a pass
# for conv_tc (E0.c1) and linear_tc is taken as the family's first edge-MLP-free launch)
# KERNELS="a b" restricts the full captures (each costs one bench start-up, ~30 s of GPU time)
KERNELS=${KERNELS:-decode_lattice_kernel decode_query_kernel sa_mlp_kernel conv_tc_kernel conv_tc_dx_kernel linear_tc_kernel fps_kernel \
         mc_classify_kernel mc_compact_kernel mc_vertices_kernel mc_faces_kernel ball_query_kernel ggm_fused_kernel}
for K in $KERNELS; do
  timeout 600 ncu --nvtx --nvtx-include "gnb.timed_step/" --set full --clock-control none --import-source on -k regex:$K -c 2 -f \
      -o $OUT/full_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$K.log 2>&1
done
ls -la $OUT | head -40
