set -x
cat > /tmp/score_once.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0,'.')
from uncertainty_nerf_gs_b200 import metrics as M
g=torch.Generator(device='cuda').manual_seed(0)
v,h,w=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
pred=torch.rand(v,h,w,3,device='cuda',generator=g); std=torch.clamp(0.1*torch.rand(v,h,w,1,device='cuda',generator=g),min=0.03)
gt=torch.clamp(pred+std*torch.randn(v,h,w,3,device='cuda',generator=g),0,1)
M.score_rgb_batch(pred,gt,std)
M.score_rgb_batch(pred,gt,std)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sel_classify|sel_resolve" -s 2 -c 2 -o gpurun_out/sel_full -f python /tmp/score_once.py 16 800 800 > gpurun_out/sel_full.log 2>&1
tail -3 gpurun_out/sel_full.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/sel_launches_1v.csv python /tmp/score_once.py 1 840 1297 > gpurun_out/sel_ncu1.log 2>&1
