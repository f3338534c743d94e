"""Runs only the PSFNoise renderer at the bench.py shape (for ncu / timing): python scripts/prof_psfnoise.py [B] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
import bench
from moleculardiffusion_mivit_b200 import experiments as X
from moleculardiffusion_mivit_b200.helpersGeneration import brownian_motion

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = bench.WORKLOADS["psfnoise_render"]
traj = brownian_motion(B, w["frames"], w["npos"], [1, 3, 5, 7, 9, 10.2], 1.0, seed=3, D_var=1.0, div=100.0, return_device=True)
for _ in range(2):
    X.trajs_to_vid_psf_noise(traj, w["npos"], True, w["props"], w["psf"], w["noise"], seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    v = X.trajs_to_vid_psf_noise(traj, w["npos"], True, w["props"], w["psf"], w["noise"], seed=1, seq_offset=i * B)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_ = B * (w["frames"] * w["npos"] * 16 + v[0].numel() * 4)
print("render_psfnoise B=%d: %.4f ms/launch (incl. the output allocation), %.1f GB/s algorithmic" % (B, ms, bytes_ / ms / 1e6))
