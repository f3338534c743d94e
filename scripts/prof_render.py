"""Runs only the V1 renderer at the bench.py shape (for ncu): python scripts/prof_render.py [B] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200.helpersGeneration import brownian_motion, derive_render_params, render_device

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
P = int(sys.argv[3]) if len(sys.argv) > 3 else 13
PROPS = {"particle_intensity": [4580, 500], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3, "resolution": 100e-9,
         "output_size": P, "upsampling_factor": 5, "background_intensity": [1420, 290], "poisson_noise": 100, "trajectory_unit": 1200}
traj = brownian_motion(B, 30, 10, [1, 3, 5, 7, 9, 10.2], 1.0, seed=3, D_var=1.0, div=100.0, return_device=True)
prm = derive_render_params(PROPS, 10, True)
prm.normalize, prm.norm_sub, prm.norm_div = 1, 1130.0, 4870.0
out = torch.empty((B, 30, P, P), dtype=torch.float32, device="cuda")
for _ in range(3):
    render_device(traj, prm, seed=1, out=out, out_seq_stride=30 * P * P)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    render_device(traj, prm, seed=1, seq_offset=i * B, out=out, out_seq_stride=30 * P * P)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_ = B * (300 * 16 + 30 * P * P * 4)
print("render_v1 B=%d P=%d: %.4f ms/launch, %.1f GB/s algorithmic" % (B, P, ms, bytes_ / ms / 1e6))
