"""Drop-in mirrors of the experiment-level renderer variants:

  trajs_to_vid_psf_noise   Experiments/PSFNoise/trainSettingsPSFNoise.py:196-263 (+ :266-309)
  trajs_to_vid_framerates  Experiments/Framerate/trainSettingsFramerate.py:170-202

The reference versions read module globals (N_PSF, N_Noise, part_mean, patch_size,
N_POSPERFRAME, originalNposPerFrame); here they are derived from the arguments, with the
globals' values available as keyword overrides."""
import ctypes

import numpy as np

from . import _lib
from .helpersGeneration import (_draw_seed, _to_device_f64, derive_render_params, normalize_images,  # noqa: F401
                                render_device, trajectories_to_video)

__all__ = ["trajs_to_vid_psf_noise", "trajs_to_vid_framerates"]


def trajs_to_vid_psf_noise(trajectories, nPosPerFrame, center=False, image_props={}, PSF_Settings=[], Noise_Settings=[], *,
                           part_mean=None, seed=None, seq_offset=0, _mean_noise=False):
    """(N,T,2) -> float32 (N, N_PSF, N_Noise, F, P, P); raw counts, no normalisation, no y flip,
    psf_division_factor ignored, background added twice for noise index >= 1 (reference quirks, kept).
    part_mean: the reference's module global used for the background sigma (:302); defaults to
    particle_intensity[0]."""
    import torch
    dev = _lib.require_cuda()
    N, T, _ = trajectories.shape
    if T % nPosPerFrame != 0:
        raise Exception("T is not divisble by posPerFrame")
    if list(PSF_Settings) == [] or list(Noise_Settings) == []:
        raise Exception("No settings given")
    prm = derive_render_params(image_props, nPosPerFrame, center, "psfnoise")
    prm.mean_noise = int(bool(_mean_noise))
    pm_glob = float(prm.part_mean if part_mean is None else part_mean)
    psf = np.asarray(PSF_Settings, dtype=np.float32)
    noi = np.asarray(Noise_Settings, dtype=np.float32)
    t_dev, was_host = _to_device_f64(trajectories, dev)
    F = T // nPosPerFrame
    out = torch.empty((N, len(psf), len(noi), F, prm.P, prm.P), dtype=torch.float32, device=dev)
    fp = ctypes.POINTER(ctypes.c_float)
    _lib.check(_lib.lib().mivit_render_psfnoise(_lib.ptr(t_dev), N, T, ctypes.byref(prm), psf.ctypes.data_as(fp), len(psf),
                                                noi.ctypes.data_as(fp), len(noi), pm_glob, _draw_seed(seed), int(seq_offset),
                                                _lib.ptr(out), _lib.current_stream()))
    return out.cpu().numpy() if was_host else out


def trajs_to_vid_framerates(trajectories, nPosPerFrame=[], center=False, image_props={}, *, originalNposPerFrame=10,
                            seed=None, seq_offset=0, _mean_noise=False):
    """(N,T,2) -> torch float32 (N, len(nPosPerFrame), T//nPosPerFrame[0], P, P), zero padded,
    each variant rendered with flux * n/originalNposPerFrame and normalised (fused).  Like the
    reference, every variant flips the caller's y in place, so variants alternate in y sign and an
    even number of variants leaves the caller's array unchanged.  Host input -> CPU tensor (as the
    reference returns); CUDA input -> CUDA tensor."""
    import torch
    dev = _lib.require_cuda()
    N, T, _ = trajectories.shape
    maxFrames = T // nPosPerFrame[0]
    part_flux, part_std = image_props["particle_intensity"]
    bg_mean, bg_sigma = image_props["background_intensity"][0], image_props["background_intensity"][1]
    P = int(image_props.get("output_size", 32))
    is_host = not isinstance(trajectories, torch.Tensor)
    out = torch.zeros((N, len(nPosPerFrame), maxFrames, P, P), dtype=torch.float32, device=dev)
    base_seed = _draw_seed(seed)
    t_dev = None
    for i, nSubPos in enumerate(nPosPerFrame):
        trajectories[:, :, 1] *= -1                      # the side effect of trajectories_to_video (:197)
        if T % nSubPos != 0:
            raise Exception("T is not divisible by nPosPerFrame")
        flux_i = part_flux * (nSubPos / originalNposPerFrame)
        props_i = image_props.copy()
        props_i["particle_intensity"] = [flux_i, part_std]
        prm = derive_render_params(props_i, nSubPos, center, "v1")
        prm.mean_noise = int(bool(_mean_noise))
        den = (bg_mean + flux_i) - (bg_mean - bg_sigma)
        if den == 0:
            raise ValueError("Denominator in normalization is zero. Check your inputs.")
        prm.normalize, prm.norm_sub, prm.norm_div = 1, float(bg_mean - bg_sigma), float(den)
        if t_dev is None:                                # upload once; the sign alternates via flip_y
            t_dev, _ = _to_device_f64(trajectories, dev)
            prm.flip_y = 0
        else:
            prm.flip_y = i % 2                           # uploaded copy carries the first flip
        view = out[:, i]
        _lib.check(_lib.lib().mivit_render_v1(_lib.ptr(t_dev), N, T, ctypes.byref(prm), (base_seed + i) & (2 ** 64 - 1),
                                              int(seq_offset), ctypes.c_void_p(view.data_ptr()),
                                              int(out.stride(0)), _lib.current_stream()))
    return out.cpu() if is_host else out
