"""Drop-in mirror of the reference's helpers/helpersFeatures.py public surface that the hot path uses: the
25-feature producer of the ViT's `features` input (ImagesFeatures experiment).  The arithmetic runs in
csrc/features.cu (one warp per trajectory, float64) through `mivit_diffusion_features`; nothing is computed on the host."""
import ctypes

import numpy as np

from . import _lib

__all__ = ["feature_names", "N_features", "compute_diffusion_features", "compute_features_for_multiple_trajectories"]

feature_names = [  # helpers/helpersFeatures.py:7-33 (order of the output array)
    "alpha", "diffusion_coefficient", "r_squared", "efficiency_log", "efficiency", "fractal_dimension", "gaussianity",
    "kurtosis", "msd_ratio", "trappedness", "trajectory_length", "mean_step_length", "mean_msd", "mean_dot_product",
    "fraction_same_direction", "fraction_positive_direction", "total_distance", "min_step", "max_step", "step_range",
    "avg_velocity", "step_cv", "fraction_small_steps", "fraction_large_steps", "convex_hull_area"]
N_features = len(feature_names)


def features_device(traj_dev, dt=1.0):
    """traj_dev: CUDA float64 (N, L, 2) -> CUDA float64 (N, 25), raw values (NaN / -inf as the reference returns them)."""
    import torch
    N, L, _ = traj_dev.shape
    out = torch.empty((N, N_features), dtype=torch.float64, device=traj_dev.device)
    _lib.check(_lib.lib().mivit_diffusion_features(_lib.ptr(traj_dev), N, L, float(dt), _lib.ptr(out), _lib.current_stream()))
    return out


def average_frames_device(traj_dev, n):
    """CUDA float64 (N, T, 2) -> (N, T // n, 2): helpers/helpersGeneration.py:48-74."""
    import torch
    N, T, _ = traj_dev.shape
    out = torch.empty((N, T // n, 2), dtype=torch.float64, device=traj_dev.device)
    _lib.check(_lib.lib().mivit_average_frames(_lib.ptr(traj_dev), N, T, int(n), _lib.ptr(out), _lib.current_stream()))
    return out


def _to_dev(a):
    import torch
    dev = _lib.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def compute_diffusion_features(trajectory, dt=1.0):
    """helpers/helpersFeatures.py:448-520: (L, 2) trajectory -> np.ndarray of the 25 features."""
    t = _to_dev(trajectory)
    return features_device(t[None, :, :2], dt)[0].cpu().numpy()


def compute_features_for_multiple_trajectories(trajectories, dt=1, nPosPerFrame=1):
    """helpers/helpersFeatures.py:524-568: (N, T, 2) -> (N, 25) with NaN replaced by 0; nPosPerFrame != 1 averages
    groups of nPosPerFrame positions first (:553-556; T must be a multiple, as in the reference's reshape)."""
    if nPosPerFrame != 1 and trajectories.shape[1] % nPosPerFrame != 0:
        raise ValueError("cannot reshape array of size %d into shape (%d,%d,newaxis)" % (
            trajectories.shape[1] * 2, trajectories.shape[1] // nPosPerFrame, nPosPerFrame))
    t = _to_dev(trajectories)
    if nPosPerFrame != 1:
        t = average_frames_device(t, nPosPerFrame)
    return np.nan_to_num(features_device(t, dt).cpu().numpy(), nan=0.0)
