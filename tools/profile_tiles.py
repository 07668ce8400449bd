"""One active-splatfacto view (1 M Gaussians at 1297x840) between cudaProfilerStart/Stop -- for ncu:
    ncu --profile-from-start off --set full --import-source on -k regex:composite_tiles -o gpurun_out/tiles python tools/profile_tiles.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import binning, synthetic
from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs

dev = torch.device("cuda:0")
H, W, G = 840, 1297, 1_000_000
sc = synthetic.splat_scene(G, H, W, seed=0, device=dev)
ids, bins = binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], H, W)
bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
run = lambda: active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"], sc["betas"],
                                        ids, bins, H, W, bg)
for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
