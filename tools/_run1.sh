timeout 100 python -m pytest tests/test_gpu_select.py -x -q 2>&1 | tail -15 > gpurun_out/sel_tests.log
cat gpurun_out/sel_tests.log
grep -q passed gpurun_out/sel_tests.log || exit 1
rm -f gpurun_out/sel_perf.jsonl
timeout 60 python tools/perf_select.py 16 800 800 > gpurun_out/sel_perf.jsonl 2> gpurun_out/sel_perf.err
UB_PERF_STD_FLOOR=0 timeout 60 python tools/perf_select.py 16 800 800 >> gpurun_out/sel_perf.jsonl 2>> gpurun_out/sel_perf.err
timeout 60 python tools/perf_select.py 8 840 1297 >> gpurun_out/sel_perf.jsonl 2>> gpurun_out/sel_perf.err
timeout 60 python tools/perf_select.py 1 840 1297 >> gpurun_out/sel_perf.jsonl 2>> gpurun_out/sel_perf.err
cat gpurun_out/sel_perf.jsonl; tail -5 gpurun_out/sel_perf.err
cat > /tmp/score_once.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0,'.')
from uncertainty_nerf_gs_b200 import metrics as M
g=torch.Generator(device='cuda').manual_seed(0)
v,h,w=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
pred=torch.rand(v,h,w,3,device='cuda',generator=g); std=torch.clamp(0.1*torch.rand(v,h,w,1,device='cuda',generator=g),min=0.03)
gt=torch.clamp(pred+std*torch.randn(v,h,w,3,device='cuda',generator=g),0,1)
M.score_rgb_batch(pred,gt,std)
PY
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/sel_launches.csv python /tmp/score_once.py 16 800 800 > gpurun_out/sel_ncu.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/sel_launches_1v.csv python /tmp/score_once.py 1 840 1297 > gpurun_out/sel_ncu1.log 2>&1
