#!/bin/bash
# Runs on the GPU box (through gpurun): regenerates the raw material of profiles/ into gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/refresh_profiles.sh r1'
# then locally: python tools/refresh_profiles.py r1
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 400 python bench.py --steps 30 --warmup 5 2> $out/${tag}_bench.err | tail -1 > $out/${tag}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>> $out/${tag}_bench.err | tail -1 > $out/${tag}_bench_reference.json
# launch list of the timed region (per-launch durations, cold cache, serialised: compare shares)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-graph > $out/${tag}_launches_bench.log 2>&1
# one full capture of the heaviest kernels of the step
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k regex:'composite_rays_tma|reduce_members_batched|sel_classify|score_prologue_kernel' -c 8 -f -o $out/${tag}_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-graph > $out/${tag}_full.log 2>&1
timeout 400 python tools/perf_kernels.py > $out/${tag}_perf_kernels.jsonl 2> $out/${tag}_perf_kernels.err
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $out/${tag}_score16_launches.csv python tools/profile_score.py > $out/${tag}_score16.log 2>&1
# every kernel of one batched scoring call, full metric set (source page: per-instruction counts of the select / prologue kernels)
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $out/${tag}_select_full \
    python tools/profile_score.py > $out/${tag}_select_full.log 2>&1
timeout 120 python tools/diag_overlap.py 10 > $out/${tag}_diag_overlap.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/${tag}_nvidia_smi.csv
echo done
