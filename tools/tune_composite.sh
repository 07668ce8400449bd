#!/bin/bash
# (NCW, NST) sweep of composite_rays_tma<48, ., .> through the UB_COMPOSITE_NCW / UB_COMPOSITE_NST hooks:
# one full view (1 089 480 rays), CUDA events over 30 calls, rotating inputs.   gpurun -- 'bash tools/tune_composite.sh'
for cfg in ${UB_TUNE_CFGS:-"7 14" "8 16" "6 12" "6 18" "9 18"}; do
  set -- $cfg
  UB_COMPOSITE_NCW=$1 UB_COMPOSITE_NST=$2 timeout 90 python - <<PY
import os, sys, torch
sys.path.insert(0, '.')
from uncertainty_nerf_gs_b200 import ops, synthetic
dev = torch.device('cuda:0')
R = 1089480
ms = [synthetic.ray_samples(R, 48, seed=i, device=dev) for i in range(3)]
def run(m):
    return ops.composite_rays(m['density'], m['deltas'], m['starts'], m['ends'], m['rgb'], m['beta'], rays_per_chunk=1 << 15)
ref = None
for m in ms: run(m)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(30): run(ms[i % 3])
b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b) / 30
o = run(ms[0]); torch.cuda.synchronize()
chk = float(o['rgb'].double().sum()) + float(o['rgb_var'].double().nan_to_num().sum()) + float(o['depth'].double().sum())
print(f"NCW={os.environ['UB_COMPOSITE_NCW']:>2s} NST={os.environ['UB_COMPOSITE_NST']:>2s}: {t*1e3:7.1f} us  {R*1576/t/1e6:6.0f} GB/s  checksum {chk:.6f}")
PY
done
