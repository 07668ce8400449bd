#!/bin/bash
# compute-sanitizer on the kernels rewritten in round 2: the select path (classify queue, last-block folds, per-tile
# counts in the fine histogram), the prologue (last-block fold), the compositor, the member reduce.
#   gpurun --timeout 2400 -- 'bash tools/sanitize_r2.sh'
out=gpurun_out
mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_select.py -q -m gpu \
    -k "ragged or arbitrary or special or randomized or small_segments or background" > $out/r2_san_memcheck_select.log 2>&1
echo "memcheck select rc=$?" ; tail -3 $out/r2_san_memcheck_select.log
timeout 900 $S --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_select.py -q -m gpu \
    -k "arbitrary or ragged or small_segments or few_distinct" > $out/r2_san_racecheck_select.log 2>&1
echo "racecheck select rc=$?" ; tail -3 $out/r2_san_racecheck_select.log
timeout 900 $S --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_reduce.py tests/test_gpu_composite.py -q -m gpu \
    > $out/r2_san_memcheck_rest.log 2>&1
echo "memcheck scoring/reduce/composite rc=$?" ; tail -3 $out/r2_san_memcheck_rest.log
timeout 900 $S --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_scoring.py -q -m gpu -k "auce or prologue or score_rgb_batch" \
    > $out/r2_san_racecheck_scoring.log 2>&1
echo "racecheck scoring rc=$?" ; tail -3 $out/r2_san_racecheck_scoring.log
