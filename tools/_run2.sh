set -x
cat > /tmp/score_once.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0,'.')
from uncertainty_nerf_gs_b200 import metrics as M
g=torch.Generator(device='cuda').manual_seed(0)
v,h,w=16,800,800
pred=torch.rand(v,h,w,3,device='cuda',generator=g); std=torch.clamp(0.1*torch.rand(v,h,w,1,device='cuda',generator=g),min=0.03)
gt=torch.clamp(pred+std*torch.randn(v,h,w,3,device='cuda',generator=g),0,1)
M.score_rgb_batch(pred,gt,std)
M.score_rgb_batch(pred,gt,std)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sel_classify|sel_cell_counts|sel_resolve|sel_alloc" -s 4 -c 4 -o gpurun_out/sel_full -f python /tmp/score_once.py > gpurun_out/sel_full.log 2>&1
tail -5 gpurun_out/sel_full.log
