"""Small composite_rays run whose stage ring wraps several times -- meant to be run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import ops, synthetic
R = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8 * 60 + 3
m = synthetic.ray_samples(R, 48, seed=0, device="cuda")
o = ops.composite_rays(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"], rays_per_chunk=1 << 15)
torch.cuda.synchronize()
print("ok", float(o["accumulation"].sum()))
