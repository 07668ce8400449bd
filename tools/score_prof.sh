#!/bin/bash
# full ncu capture (source-level counters) of the kernels of one batched scoring call
#   gpurun --timeout 900 -- 'bash tools/score_prof.sh tag [kernel-regex]'
tag=${1:-prof}
rx=${2:-sel_classify|sel_fine_hist|score_prologue_kernel}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"$rx" -f -o $out/${tag}_full \
    python tools/profile_score.py > $out/${tag}_full.log 2>&1
tail -2 $out/${tag}_full.log
