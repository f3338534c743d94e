"""Drop-in mirror of the reference's helpers/helpersGeneration.py renderer API
 on top of the CUDA renderer (csrc/render.cu).

Same names, positional order, defaults, return types, error messages and side effects:
  trajectories_to_video   helpers/helpersGeneration.py:128-278 (flips the caller's y in place, :197)
  normalize_images        helpers/helpersGeneration.py:356-400
  brownian_motion         helpers/helpersGeneration.py:9-45   (device Philox source)
Extra keyword-only arguments (`seed`, `seq_offset`, ...) are additions; with seed=None a
seed is drawn from np.random so unseeded behaviour stays random and np.random.seed still
controls it.  Inputs may be numpy arrays (host: copied to the GPU and back inside the
call) or CUDA float64 torch tensors (device resident: a CUDA float32 tensor is returned).
"""
import ctypes

import numpy as np

from . import _lib

__all__ = ["trajectories_to_video", "trajectories_to_embeddings", "trajectories_to_video_multiple_settings", "trajs_to_vid_norm_rl", "generateTrajAndVideosBrownian", "create_gaussian_psf", "richardson_lucy_tv",
           "richardson_lucy_tv_iter_list", "apply_rl_tv_tensor", "apply_rl_tv_tensor_iter_list", "create_video_and_feature_pairs",
           "average_trajectories_frames", "average_trajs_add_error", "normalize_images", "brownian_motion", "derive_render_params",
           "DEFAULT_IMAGE_PROPS", "render_device"]

DEFAULT_IMAGE_PROPS = {  # helpers/helpersGeneration.py:205-222
    "particle_intensity": [500, 20],
    "NA": 1.46,
    "wavelength": 500e-9,
    "psf_division_factor": 1,
    "resolution": 100e-9,
    "output_size": 32,
    "upsampling_factor": 5,
    "background_intensity": [100, 10],
    "poisson_noise": 100,
    "trajectory_unit": 100,
}


def _draw_seed(seed):
    if seed is None:
        return int(np.random.randint(0, 2 ** 62, dtype=np.int64))
    return int(seed) & 0xFFFFFFFFFFFFFFFF


def derive_render_params(image_props, nPosPerFrame, center, variant="v1"):
    """Host-side scalar set-up of :225-247 (or trainSettingsPSFNoise.py:237-259) -> RenderParams."""
    d = dict(DEFAULT_IMAGE_PROPS)
    if variant in ("psfnoise", "multi"):
        d["poisson_noise"] = 1
    d.update(image_props)
    res, unit = d["resolution"], d["trajectory_unit"]
    if unit == -1:
        scale = 1.0
    elif variant in ("psfnoise", "multi"):
        scale = unit * 1e-9 / res
    else:
        scale = unit / (res * 1e9)
    U = int(d["upsampling_factor"])
    if variant == "psfnoise":
        fwhm = d["wavelength"] / 2 * d["NA"]
    else:
        fwhm = d["wavelength"] / 2 * d["NA"] / d["psf_division_factor"]
    sigma = U / res * fwhm / 2.355
    pm, pstd = d["particle_intensity"][0], d["particle_intensity"][1]
    bm, bs = d["background_intensity"][0], d["background_intensity"][1]
    p = _lib.RenderParams()
    p.scale, p.sigma_hr = float(scale), float(sigma)
    p.P, p.U, p.n = int(d["output_size"]), U, int(nPosPerFrame)
    p.center, p.flip_y = int(bool(center)), 0 if variant == "psfnoise" else 1
    p.draw = int(pm > 0.0001 and pstd > 0.0001)
    p.part_mean, p.part_std, p.bg_mean, p.bg_std = float(pm), float(pstd), float(bm), float(bs)
    p.poisson = float(d["poisson_noise"])
    p.normalize, p.norm_sub, p.norm_div, p.mean_noise = 0, 0.0, 1.0, 0
    return p


def _to_device_f64(trajectories, dev):
    """numpy (host) or torch tensor -> contiguous CUDA float64 tensor.  Returns (tensor, was_host)."""
    import torch
    if isinstance(trajectories, torch.Tensor):
        t = trajectories
        if t.device.type != "cuda":
            t = t.to(dev, non_blocking=True)
        return t.to(torch.float64).contiguous(), False
    host = torch.from_numpy(np.ascontiguousarray(trajectories, dtype=np.float64))
    return host.to(dev, non_blocking=False), True


def render_device(traj_dev, prm, seed, seq_offset=0, out=None, out_seq_stride=None, flip_y_applied=True):
    """Device-resident core: traj_dev CUDA float64 (N,T,2) -> CUDA float32 (N,F,P,P).
    prm.flip_y selects the sign the frames are rendered with; traj_dev is not modified."""
    import torch
    N, T, _ = traj_dev.shape
    F = T // prm.n
    if out is None:
        out = torch.empty((N, F, prm.P, prm.P), dtype=torch.float32, device=traj_dev.device)
        out_seq_stride = F * prm.P * prm.P
    _lib.check(_lib.lib().mivit_render_v1(_lib.ptr(traj_dev), N, T, ctypes.byref(prm), seed, int(seq_offset),
                                          _lib.ptr(out), int(out_seq_stride), _lib.current_stream()))
    return out


def trajectories_to_video(trajectories, nPosPerFrame, center=False, image_props={}, use_multiprocessing=False, *,
                          seed=None, seq_offset=0, normalize=None, _mean_noise=False):
    """helpers/helpersGeneration.py:128-278.  `use_multiprocessing` is accepted and ignored (the
    reference's branch is dead code: it raises TypeError, :344-348 vs :283).
    normalize=(background_mean, background_sigma, theoretical_max) fuses normalize_images."""
    import torch
    dev = _lib.require_cuda()
    N, T, _ = trajectories.shape
    # Invert the y axis IN PLACE, exactly like the reference (:197) -- callers rely on it
    # (trajs_to_vid_framerates alternates sign per variant; create_video_and_feature_pairs
    # computes features from the flipped trajectories).
    trajectories[:, :, 1] *= -1
    if T % nPosPerFrame != 0:
        raise Exception("T is not divisble by posPerFrame")
    prm = derive_render_params(image_props, nPosPerFrame, center, "v1")
    prm.flip_y = 0                      # the array we upload is already flipped
    prm.mean_noise = int(bool(_mean_noise))
    if normalize is not None:
        m, s, mx = normalize
        den = mx - (m - s)
        if den == 0:
            raise ValueError("Denominator in normalization is zero. Check your inputs.")
        prm.normalize, prm.norm_sub, prm.norm_div = 1, float(m - s), float(den)
    t_dev, was_host = _to_device_f64(trajectories, dev)
    out = render_device(t_dev, prm, _draw_seed(seed), seq_offset)
    if was_host:
        return out.cpu().numpy()
    return out


def trajectories_to_embeddings(trajectories, nPosPerFrame, embedding, center=False, image_props={}, *, seed=None, seq_offset=0,
                               normalize=None, return_frames=False, _mean_noise=False):
    """Renderer fused with the frame embedding (no reference counterpart as ONE call: it is
    `embedding(torch.Tensor(normalize_images(trajectories_to_video(...))))` of the training loops,
    helpers/helpersGeneration.py:128-278,356-400 + helpers/models.py:146-199) for `LinearProjectionEmbedding` /
    `CNNEmbedding`: returns CUDA float32 [N,F,E]; frames never touch HBM unless return_frames=True.
    Same side effect (in-place y flip), error text, image_props and RNG streams as trajectories_to_video, so
    `embedding(trajectories_to_video(..., seed=s))` and `trajectories_to_embeddings(..., seed=s)` see the same frames."""
    import torch
    from . import models as _m
    dev = _lib.require_cuda()
    if not isinstance(embedding, (_m.LinearProjectionEmbedding, _m.CNNEmbedding)):
        raise TypeError("the fused render+embed kernel covers LinearProjectionEmbedding and CNNEmbedding "
                        "(DeepResNetEmbedding needs batch statistics: render, then call the model)")
    N, T, _ = trajectories.shape
    trajectories[:, :, 1] *= -1
    if T % nPosPerFrame != 0:
        raise Exception("T is not divisble by posPerFrame")
    prm = derive_render_params(image_props, nPosPerFrame, center, "v1")
    prm.flip_y = 0
    prm.mean_noise = int(bool(_mean_noise))
    if normalize is not None:
        m, s, mx = normalize
        den = mx - (m - s)
        if den == 0:
            raise ValueError("Denominator in normalization is zero. Check your inputs.")
        prm.normalize, prm.norm_sub, prm.norm_div = 1, float(m - s), float(den)
    if isinstance(embedding, _m.LinearProjectionEmbedding):
        W, b = embedding.proj.weight, embedding.proj.bias
    else:
        W, b = embedding.conv.weight, embedding.conv.bias
    E = W.shape[0]
    if W.numel() != E * prm.P * prm.P:
        raise ValueError("embedding patch_size %d does not match output_size %d" % (int(round((W.numel() // E) ** 0.5)), prm.P))
    Wt = W.detach().to(device=dev, dtype=torch.float32).reshape(E, prm.P * prm.P).t().contiguous()
    bd = b.detach().to(device=dev, dtype=torch.float32).contiguous()
    t_dev, _ = _to_device_f64(trajectories, dev)
    F = T // nPosPerFrame
    emb = torch.empty((N, F, E), dtype=torch.float32, device=dev)
    frames = torch.empty((N, F, prm.P, prm.P), dtype=torch.float32, device=dev) if return_frames else None
    _lib.check(_lib.lib().mivit_render_embed_linear(_lib.ptr(t_dev), N, T, ctypes.byref(prm), _draw_seed(seed), int(seq_offset),
                                                    _lib.ptr(Wt), _lib.ptr(bd), E, _lib.ptr(emb), _lib.ptr(frames),
                                                    F * prm.P * prm.P, _lib.current_stream()))
    return (emb, frames) if return_frames else emb


def trajectories_to_video_multiple_settings(trajectories, nPosPerFrame, center=False, image_props={}, *, seed=None, seq_offset=0,
                                            _mean_noise=False):
    """helpers/helpersGeneration.py:422-492 (+ trajectory_to_mult_settings :494-540).  (N,T,2) -> four float32 (N,F,P,P) arrays:
    (no noise, + Gaussian background, + Poisson noise, Gaussian-filtered).  Like the reference it flips the caller's y IN PLACE
    (:432) and raises after the flip when T is not a multiple of nPosPerFrame; `poisson_noise` defaults to 1 here (:455)."""
    import torch
    dev = _lib.require_cuda()
    N, T, _ = trajectories.shape
    trajectories[:, :, 1] *= -1
    if T % nPosPerFrame != 0:
        raise Exception("T is not divisble by posPerFrame")
    prm = derive_render_params(image_props, nPosPerFrame, center, "multi")
    prm.flip_y = 0                      # already applied to the caller's array above
    prm.mean_noise = int(bool(_mean_noise))
    t_dev, was_host = _to_device_f64(trajectories, dev)
    F = T // nPosPerFrame
    outs = [torch.empty((N, F, prm.P, prm.P), dtype=torch.float32, device=dev) for _ in range(4)]
    _lib.check(_lib.lib().mivit_render_multi(_lib.ptr(t_dev), N, T, ctypes.byref(prm), _draw_seed(seed), int(seq_offset),
                                             _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]),
                                             _lib.current_stream()))
    return tuple(o.cpu().numpy() for o in outs) if was_host else tuple(outs)


def generateTrajAndVideosBrownian(Ds, nPart, nImages, nPosPerFrame, optics_props, *, seed=None, seq_offset=0):
    """helpers/helpersGeneration.py:402-417: nPart Brownian trajectories with D ~ N(Ds[0], Ds[1]) (the andi_datasets
    `single_state(L=0, alphas=1)` call, replaced by this package's device generator: statistical contract only, see
    oracle/trajectory_oracle.py) rendered with trajectories_to_video(center=True).  Returns (videos (N,F,P,P) float32 numpy,
    D (N,) float32 numpy) -- the reference's `labels[:, 0, 1]`."""
    T = int(nImages) * int(nPosPerFrame)
    s = _draw_seed(seed)
    trajs, D = brownian_motion(nPart, T, 1, [Ds[0]], 1.0, seed=s, seq_offset=seq_offset, D_var=Ds[1], return_D=True)
    videos = trajectories_to_video(trajs, nPosPerFrame, True, optics_props, seed=s, seq_offset=seq_offset)
    return videos, D


def create_gaussian_psf(size=9, sigma=1.3):
    """helpers/helpersGeneration.py:591-599 (a K x K float64 parameter array; host arithmetic as in the reference)."""
    if size % 2 == 0:
        size += 1  # ensure odd size for symmetry
    ax = np.arange(-size // 2 + 1, size // 2 + 1)
    x, y = np.meshgrid(ax, ax)
    psf = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    psf /= psf.sum()
    return psf


def _rl_tv_device(frames_dev, psf, iterations_list, tv_weight):
    """frames_dev: CUDA float32 [n, P, P] -> CUDA float32 [n, len(iterations_list), P, P] (include/mivit.h: mivit_rl_tv)."""
    import torch
    n, P = int(frames_dev.shape[0]), int(frames_dev.shape[-1])
    psf = np.ascontiguousarray(psf, dtype=np.float64)
    if psf.ndim != 2 or psf.shape[0] != psf.shape[1]:
        raise ValueError("psf must be a square 2-D array")
    its = np.asarray(list(iterations_list), dtype=np.int32)
    if its.size == 0 or np.any(np.diff(its) <= 0):
        raise ValueError("iterations_list must be a non-empty ascending list")
    psf_dev = torch.from_numpy(psf).to(frames_dev.device)
    out = torch.empty((n, its.size, P, P), dtype=torch.float32, device=frames_dev.device)
    _lib.check(_lib.lib().mivit_rl_tv(_lib.ptr(frames_dev), n, P, _lib.ptr(psf_dev), int(psf.shape[0]),
                                      its.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(its.size), float(tv_weight),
                                      _lib.ptr(out), _lib.current_stream()))
    return out


def richardson_lucy_tv(image, psf, iterations=20, tv_weight=0.01):
    """helpers/helpersGeneration.py:557-569: one P x P image -> its estimate after `iterations` iterations (float32)."""
    import torch
    dev = _lib.require_cuda()
    was_host = not torch.is_tensor(image)
    x = torch.as_tensor(np.asarray(image) if was_host else image).to(device=dev, dtype=torch.float32).contiguous()
    out = _rl_tv_device(x.reshape(1, x.shape[-2], x.shape[-1]), psf, [int(iterations) - 1], tv_weight)[0, 0]
    return out.cpu().numpy() if was_host else out


def richardson_lucy_tv_iter_list(image, psf, iterations_list, out_array, tv_weight=0.01):
    """helpers/helpersGeneration.py:571-587: fills out_array[k] with the estimate after iteration iterations_list[k] (0-based)
    and returns the last estimate."""
    import torch
    dev = _lib.require_cuda()
    x = torch.as_tensor(np.asarray(image)).to(device=dev, dtype=torch.float32).contiguous()
    out = _rl_tv_device(x.reshape(1, x.shape[-2], x.shape[-1]), psf, iterations_list, tv_weight)[0].cpu().numpy()
    for k in range(out.shape[0]):
        out_array[k] = out[k]
    return out[-1]


def apply_rl_tv_tensor(tensor, psf, n_iters=10, tv_weight=0.01):
    """helpers/helpersGeneration.py:603-614: [B, seq, H, W] torch tensor -> same shape / dtype / device."""
    import torch
    dev = _lib.require_cuda()
    B, seq, H, W = tensor.shape
    assert H == 9 and W == 9, "Only images of shape 9x9 are supported"
    x = tensor.detach().to(device=dev, dtype=torch.float32).contiguous().reshape(B * seq, H, W)
    out = _rl_tv_device(x, psf, [int(n_iters) - 1], tv_weight)[:, 0].reshape(B, seq, H, W)
    return out.to(dtype=tensor.dtype, device=tensor.device)


def apply_rl_tv_tensor_iter_list(tensor, psf, iterations_list=[2, 5, 10], tv_weight=0.01):
    """helpers/helpersGeneration.py:616-631: [B, seq, H, W] (torch tensor or numpy) -> numpy [B, len(iterations_list), seq, H, W]."""
    import torch
    dev = _lib.require_cuda()
    B, seq, H, W = tensor.shape
    assert H == 9 and W == 9, "Only images of shape 9x9 are supported"
    t = tensor.detach() if torch.is_tensor(tensor) else torch.from_numpy(np.ascontiguousarray(tensor))
    np_dtype = tensor.detach().cpu().numpy().dtype if torch.is_tensor(tensor) else tensor.dtype
    x = t.to(device=dev, dtype=torch.float32).contiguous().reshape(B * seq, H, W)
    out = _rl_tv_device(x, psf, iterations_list, tv_weight).reshape(B, seq, len(iterations_list), H, W).permute(0, 2, 1, 3, 4)
    return out.contiguous().cpu().numpy().astype(np_dtype, copy=False)


def trajs_to_vid_norm_rl(trajectories, nPosPerFrame, center, image_props, rl_iterations, poisson_index=2, *, seed=None,
                         seq_offset=0, _mean_noise=False):
    """helpers/helpersGeneration.py:635-658 (the Denoising experiments' generator): the four multiple-settings videos stacked
    to (N, 4, F, P, P), normalised with (bg_mean, bg_sigma, part_mean + bg_mean), plus the Richardson-Lucy/TV estimates
    (Gaussian PSF, sigma = 1) of the Poisson-noise videos after the iterations in `rl_iterations` ->
    numpy float32 (N, 4 + len(rl_iterations), F, P, P).  Everything stays on the GPU until the final copy."""
    import torch
    bg_mean, bg_sigma = image_props["background_intensity"]
    part_mean, part_sigma = image_props["particle_intensity"]
    psf = create_gaussian_psf(sigma=1)
    was_host = not torch.is_tensor(trajectories)
    if was_host:      # render from a device copy, keep the reference's in-place y flip of the caller's array (:432)
        dev = _lib.require_cuda()
        N, T, _ = trajectories.shape
        trajectories[:, :, 1] *= -1
        if T % nPosPerFrame != 0:
            raise Exception("T is not divisble by posPerFrame")
        t_dev = torch.from_numpy(np.ascontiguousarray(trajectories, dtype=np.float64)).to(dev)
        t_dev[:, :, 1] *= -1      # trajectories_to_video_multiple_settings flips once more below
    else:
        t_dev = trajectories
    vids = trajectories_to_video_multiple_settings(t_dev, nPosPerFrame, center=center, image_props=image_props, seed=seed,
                                                   seq_offset=seq_offset, _mean_noise=_mean_noise)
    videos = torch.stack(list(vids), dim=1)                                   # (N, 4, F, P, P)
    lo = bg_mean - bg_sigma
    den = (part_mean + bg_mean) - lo
    if den == 0:
        raise ValueError("Denominator in normalization is zero. Check the input values.")
    videos = (videos - lo) / den                                              # normalize_images (:389-395), float32
    to_rl = videos[:, poisson_index]
    N, F, P = to_rl.shape[0], to_rl.shape[1], to_rl.shape[-1]
    assert P == 9, "Only images of shape 9x9 are supported"
    rl = _rl_tv_device(to_rl.contiguous().reshape(N * F, P, P), psf, rl_iterations, 0.01)
    rl = rl.reshape(N, F, len(rl_iterations), P, P).permute(0, 2, 1, 3, 4)
    out = torch.cat([videos, rl], dim=1)
    return out.cpu().numpy() if was_host else out


def average_trajectories_frames(trajectories, nPosFrame):
    """helpers/helpersGeneration.py:48-74 (device kernel; numpy in, numpy out)."""
    from . import helpersFeatures as _hf
    return _hf.average_frames_device(_hf._to_dev(trajectories), nPosFrame).cpu().numpy()


def average_trajs_add_error(trajectories, nPosPerFrame, localization_uncertainty):
    """helpers/helpersGeneration.py:663-672: frame means and frame means + N(mu_loc, sigma_loc) localisation noise
    (drawn from the global np.random state like the reference)."""
    avg = average_trajectories_frames(trajectories, nPosPerFrame)
    mu, sigma = localization_uncertainty
    return avg, avg + np.random.normal(loc=mu, scale=sigma, size=avg.shape)


def create_video_and_feature_pairs(trajectories, nPosPerFrame, center, image_props, localization_uncertainty=(0, 0), dt=1.0, *,
                                   seed=None):
    """helpers/helpersGeneration.py:674-719: normalised videos (N,F,P,P) float32, the 25 diffusion features (N,25)
    float64 of the frame-averaged trajectories, and (trajectories, averaged, averaged + localisation error).
    As in the reference the features are computed AFTER trajectories_to_video flipped the caller's y axis in place."""
    from . import helpersFeatures as _hf
    bg_mean, bg_sigma = image_props["background_intensity"]
    part_mean, _ = image_props["particle_intensity"]
    videos = trajectories_to_video(trajectories, nPosPerFrame, center=center, image_props=image_props, seed=seed,
                                   normalize=(bg_mean, bg_sigma, part_mean + bg_mean))
    avg_dev = _hf.average_frames_device(_hf._to_dev(trajectories), nPosPerFrame)
    features = _hf.features_device(avg_dev, dt).cpu().numpy()
    avg = avg_dev.cpu().numpy()
    mu, sigma = localization_uncertainty
    avg_err = avg + np.random.normal(loc=mu, scale=sigma, size=avg.shape)
    return videos, features, (trajectories, avg, avg_err)


def normalize_images(images, background_mean=None, background_sigma=None, theoretical_max=None, clip_image=False):
    """helpers/helpersGeneration.py:356-400 -- elementwise; works on numpy arrays and torch tensors
    (host metadata only, the arithmetic of the array type is used as in the reference)."""
    import torch
    is_t = isinstance(images, torch.Tensor)
    if background_mean is None:
        background_mean = images.mean().item() if is_t else np.mean(images)
    if background_sigma is None:
        background_sigma = images.std(unbiased=False).item() if is_t else np.std(images)
    if theoretical_max is None:
        theoretical_max = images.max().item() if is_t else np.max(images)
    denominator = theoretical_max - (background_mean - background_sigma)
    if denominator == 0:
        raise ValueError("Denominator in normalization is zero. Check your inputs.")
    normalized = (images - (background_mean - background_sigma)) / denominator
    if clip_image:
        normalized = normalized.clamp(0, 1.5) if is_t else np.clip(normalized, 0, 1.5)
    return normalized, (background_mean, background_sigma, theoretical_max)


def brownian_motion(nparticles, nframes, nposframe, D, dt, startAtZero=False, *, seed=None, seq_offset=0,
                    D_var=0.0, div=1.0, return_device=False, return_D=False):
    """helpers/helpersGeneration.py:9-45 on the device: steps ~ N(0, 2 D dt / nposframe) per axis,
    cumulative sum.  Positions start at the origin (the reference starts at the first step unless
    startAtZero; with centring only differences matter).  D may be a scalar or a list of group
    means (global sequence id g uses group g % len(D)); D_var is the variance of the per-sequence D
    (the andi_datasets `Ds=[mean, var]` call sites, trainModelsPSFNoise.py:128-132)."""
    import torch
    dev = _lib.require_cuda()
    T = int(nframes) * int(nposframe)
    means = np.atleast_1d(np.asarray(D, dtype=np.float32)) * np.float32(dt / nposframe)
    vars_ = np.broadcast_to(np.atleast_1d(np.asarray(D_var, dtype=np.float32)) * np.float32((dt / nposframe) ** 2),
                            means.shape).astype(np.float32).copy()
    traj = torch.empty((nparticles, T, 2), dtype=torch.float64, device=dev)
    Dout = torch.empty((nparticles,), dtype=torch.float32, device=dev)
    fp = ctypes.POINTER(ctypes.c_float)
    _lib.check(_lib.lib().mivit_brownian(nparticles, T, means.ctypes.data_as(fp), vars_.ctypes.data_as(fp), len(means),
                                         float(div), _draw_seed(seed), int(seq_offset), _lib.ptr(traj), _lib.ptr(Dout),
                                         _lib.current_stream()))
    Dout = Dout / float(dt / nposframe)
    if not return_device:
        traj, Dout = traj.cpu().numpy(), Dout.cpu().numpy()
    return (traj, Dout) if return_D else traj
