#!/bin/bash
# compute-sanitizer over the small-shape GPU tests.  bash tools/gpu_sanitizer.sh [tag]
TAG=${1:-sanitizer}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SMALL="tests/test_pointops.py tests/test_dense_ops.py tests/test_mesh_cleanup.py tests/test_mesh_ops.py tests/test_metrics.py tests/test_mc33.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest $SMALL tests/test_marching_cubes.py -m gpu -x -q -k "not 128 and not batch32" > $OUT/memcheck.log 2>&1
echo "memcheck exit $?" >> $OUT/memcheck.log
tail -8 $OUT/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_pointops.py tests/test_mesh_cleanup.py tests/test_mc33.py -m gpu -x -q > $OUT/racecheck.log 2>&1
echo "racecheck exit $?" >> $OUT/racecheck.log
tail -8 $OUT/racecheck.log
# the tensor-core kernels (tcgen05 / TMA / mbarrier pipelines): memcheck only, smallest decoder + linear tests
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_linear_tc.py tests/test_conv_tc.py tests/test_decode_tc.py -m gpu -x -q -k "not 128 and not lattice" > $OUT/memcheck_tc.log 2>&1
echo "memcheck_tc exit $?" >> $OUT/memcheck_tc.log
tail -8 $OUT/memcheck_tc.log
