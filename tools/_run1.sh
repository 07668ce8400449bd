timeout 150 python -m pytest tests/test_gpu_select.py -x -q 2>&1 | tail -15 > gpurun_out/sel_tests.log
cat gpurun_out/sel_tests.log
grep -q passed gpurun_out/sel_tests.log || exit 1
rm -f gpurun_out/sel_perf.jsonl
timeout 60 python tools/perf_select.py 16 800 800 > gpurun_out/sel_perf.jsonl 2> gpurun_out/sel_perf.err
UB_PERF_STD_FLOOR=0 timeout 60 python tools/perf_select.py 16 800 800 >> gpurun_out/sel_perf.jsonl 2>> gpurun_out/sel_perf.err
timeout 60 python tools/perf_select.py 1 840 1297 >> gpurun_out/sel_perf.jsonl 2>> gpurun_out/sel_perf.err
cat gpurun_out/sel_perf.jsonl; tail -5 gpurun_out/sel_perf.err
timeout 100 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/sel_launches.csv python tools/profile_score.py > gpurun_out/sel_ncu.log 2>&1
