#!/bin/bash
# GPU iteration loop for the scoring kernels: parity tests of the select / scoring path, then the launch list of one
# batched scoring call (16 x 800x800) and of a single 1297x840 view, then the timed micro-benchmarks.
#   gpurun --timeout 900 -- 'bash tools/score_iter.sh tag'
tag=${1:-it}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_scoring.py tests/test_gpu_full_size_configs.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1
tail -3 $out/${tag}_pytest.log
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none --csv --log-file $out/${tag}_score16.csv python tools/profile_score.py > $out/${tag}_score16.log 2>&1
UB_PROFILE_VIEWS=1 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none --csv --log-file $out/${tag}_score1.csv python tools/profile_score.py > $out/${tag}_score1.log 2>&1
timeout 200 python tools/perf_select.py 16 800 800 > $out/${tag}_perf_select.jsonl 2>&1
timeout 200 python tools/perf_select.py 1 840 1297 >> $out/${tag}_perf_select.jsonl 2>&1
cat $out/${tag}_perf_select.jsonl
