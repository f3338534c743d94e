"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's
trajectory -> image-sequence renderer.  Never imported by the product package.

Parity status: PINNED.  tests/test_oracle_render.py checks this file against the
unmodified reference (imported through oracle/refshim.py when /root/reference is
present) and against the committed fixtures in tests/golden/ that
oracle/make_golden.py generated from the reference in the build container.

Reference code restated here (paths under /root/reference):
  helpers/helpersGeneration.py:77-97    gaussian_2d              -> _spot_literal
  helpers/helpersGeneration.py:128-278  trajectories_to_video    -> derive_params, render_v1
  helpers/helpersGeneration.py:283-319  trajectory_to_video      -> _render_sequence_v1
  helpers/helpersGeneration.py:356-400  normalize_images         -> normalize_images
  Experiments/PSFNoise/trainSettingsPSFNoise.py:196-309  trajs_to_vid_psf_noise -> render_psfnoise
  Experiments/Framerate/trainSettingsFramerate.py:170-202 trajs_to_vid_framerates -> render_framerates
  helpers/helpersGeneration.py:9-45     brownian_motion          -> brownian_oracle (in trajectory_oracle.py)

Two evaluation modes of one spot:
  'literal'   : the reference's algorithm verbatim -- a (G,G) float64 exp grid,
                renormalised by its on-grid maximum, accumulated into a float32 HR
                frame, block-averaged (the loops of :285-310).
  'separable' : the closed form the CUDA kernel uses:
                I * ay (x) ax with a[j] = exp(-((x_j-c)^2 - min_j (x_j-c)^2) / 2s^2),
                block-summed per axis, float32 like the kernel.
Both are checked against each other and against the reference.
"""
import numpy as np

from .noise import MeanNoise, NumpyNoise, PhiloxNoise  # noqa: F401

F32 = np.float32

DEFAULT_IMAGE_PROPS = {  # helpersGeneration.py:205-222
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


def derive_params(image_props, variant="v1"):
    """Scalar set-up of trajectories_to_video (:225-247) / trajs_to_vid_psf_noise (:237-259)."""
    d = dict(DEFAULT_IMAGE_PROPS)
    if variant in ("psfnoise", "multi"):
        d["poisson_noise"] = 1  # trainSettingsPSFNoise.py:232, helpersGeneration.py:455
    d.update(image_props)
    res = d["resolution"]
    unit = d["trajectory_unit"]
    if unit == -1:
        scale = 1.0
    elif variant in ("psfnoise", "multi"):
        scale = unit * 1e-9 / res          # trainSettingsPSFNoise.py:241, helpersGeneration.py:464
    else:
        scale = unit / (res * 1e9)         # helpersGeneration.py:231
    U = int(d["upsampling_factor"])
    if variant == "psfnoise":
        fwhm = d["wavelength"] / 2 * d["NA"]                              # :247 (psf_division_factor ignored)
    else:
        fwhm = d["wavelength"] / 2 * d["NA"] / d["psf_division_factor"]   # :239
    sigma = U / res * fwhm / 2.355                                        # :242
    return {
        "scale": float(scale), "sigma": float(sigma), "U": U, "P": int(d["output_size"]),
        "part_mean": float(d["particle_intensity"][0]), "part_std": float(d["particle_intensity"][1]),
        "bg_mean": float(d["background_intensity"][0]), "bg_std": float(d["background_intensity"][1]),
        "poisson": float(d["poisson_noise"]),
    }


def grid_axis(G):
    """np.linspace(-limit, limit, G) of gaussian_2d (:90-91)."""
    limit = (G - 1) // 2
    return np.linspace(-limit, limit, G), limit


def _spot_literal(xc, yc, sigma, G, amplitude):
    ax, _ = grid_axis(G)
    x, y = np.meshgrid(ax, ax)
    return amplitude * np.exp(-(((x - xc) ** 2) / (2 * sigma ** 2) + ((y - yc) ** 2) / (2 * sigma ** 2)))


def axis_profile_separable(c, sigma, G, U):
    """Block-mean axis profile of one sub-position, float32 like csrc/render.cu.
    c: float64 spot centre in HR pixels.  Returns (profile[P] float32, min_sq float64)."""
    limit = (G - 1) // 2
    step64 = (2.0 * limit / (G - 1)) if G > 1 else 1.0
    jc = int(np.clip(np.rint((c + limit) / step64), 0, G - 1))
    c0 = F32(c - (-limit + jc * step64))            # offset from nearest grid node, |c0| <= step/2 inside
    step = F32(step64)
    k = (np.arange(G, dtype=np.int32) - jc).astype(F32) * step
    inv2s2 = F32(1.0 / (2.0 * sigma * sigma))
    arg = (-(k * (k - F32(2.0) * c0)) * inv2s2).astype(F32)
    a = np.exp(arg).astype(F32)
    prof = a.reshape(G // U, U)
    acc = np.zeros(G // U, dtype=F32)
    for u in range(U):                               # sequential float32 sum, like the kernel
        acc = (acc + prof[:, u]).astype(F32)
    return (acc / F32(U)).astype(F32), float(c0) * float(c0)


def _frame_centres(traj_px, f, n, center, U):
    seg = traj_px[f * n:(f + 1) * n, :]
    if center:
        seg = seg - np.mean(seg, axis=0)             # :291
    return seg[:, 0] * U, seg[:, 1] * U              # :292-293


def _accumulate_frame(xs, ys, intens, sigmas, G, U, mode):
    """Noise-free LR frames for a list of PSF sigmas.  intens[p] float64 spot
    intensities.  Returns float32 array (len(sigmas), P, P)."""
    P = G // U
    out = np.zeros((len(sigmas), P, P), dtype=F32)
    for si, sigma in enumerate(sigmas):
        if mode == "literal":
            hr = np.zeros((G, G), dtype=F32)
            for p in range(len(xs)):
                spot = _spot_literal(xs[p], ys[p], sigma, G, intens[p])
                with np.errstate(divide="ignore", invalid="ignore"):
                    hr += intens[p] / np.max(spot) * spot          # :305-308 (f64 addend, f32 accumulate)
            out[si] = hr.reshape(P, U, P, U).mean(axis=(1, 3), dtype=F32)   # block_reduce(np.mean) :310
        else:
            lr = np.zeros((P, P), dtype=F32)
            for p in range(len(xs)):
                axp, mx = axis_profile_separable(xs[p], sigma, G, U)
                ayp, my = axis_profile_separable(ys[p], sigma, G, U)
                if (mx + my) / (2.0 * sigma * sigma) > 745.0:
                    # reference: spot underflows to 0 on the whole grid -> I/0*0 = NaN frame (:305-308)
                    lr[:] = np.nan
                    continue
                Ip = F32(intens[p])
                lr = (lr + (Ip * ayp)[:, None].astype(F32) * axp[None, :]).astype(F32)
            out[si] = lr
    return out


def render_v1(traj, n, center, image_props, noise=None, seq_offset=0, mode="separable", flip_y=True):
    """trajectories_to_video without the in-place side effect: `traj` (N,T,2) float64
    is NOT modified; flip_y=True applies the reference's y sign flip (:197).
    Returns float32 (N,F,P,P)."""
    traj = np.asarray(traj, dtype=np.float64)
    N, T, _ = traj.shape
    if T % n != 0:
        raise Exception("T is not divisble by posPerFrame")   # :200-201
    prm = derive_params(image_props, "v1")
    P, U = prm["P"], prm["U"]
    G = P * U
    F = T // n
    noise = noise if noise is not None else MeanNoise()
    tr = traj.copy()
    if flip_y:
        tr[:, :, 1] *= -1
    tr = tr * prm["scale"]
    out = np.zeros((N, F, P, P), dtype=F32)
    draw = prm["part_mean"] > 0.0001 and prm["part_std"] > 0.0001       # :299
    bmean, bstd = F32(prm["bg_mean"]), F32(prm["bg_std"])
    hi = F32(prm["bg_mean"] + 3 * prm["bg_std"])
    for s in range(N):
        seq = seq_offset + s
        zI = noise.intensity_z(seq, F * n, v1=True).reshape(F, n)
        zb, kpois = noise.pixel_v1(seq, F, P, prm["poisson"])
        for f in range(F):
            xs, ys = _frame_centres(tr[s], f, n, center, U)
            if draw:
                intens = (F32(prm["part_mean"] / n) + F32(prm["part_std"] / n) * zI[f]).astype(F32)   # :300
                lr = _accumulate_frame(xs, ys, intens.astype(np.float64), [prm["sigma"]], G, U, mode)[0]
            else:
                lr = np.zeros((P, P), dtype=F32)
            bg = np.clip((bmean + bstd * zb[f]).astype(F32), F32(0), hi)        # :312-313
            out[s, f] = (lr + bg).astype(F32)
        if prm["poisson"] != -1:                                                  # :316-317 (multiplicative)
            pn = F32(prm["poisson"])
            out[s] = ((out[s] * kpois).astype(F32) * (F32(1.0) / pn)).astype(F32)
    return out


def render_psfnoise(traj, n, center, image_props, psf_settings, noise_settings, noise=None,
                    seq_offset=0, mode="separable", part_mean_global=None):
    """trajs_to_vid_psf_noise (trainSettingsPSFNoise.py:196-309).  Returns float32
    (N, n_psf, n_noise, F, P, P).  No y flip, psf_division_factor ignored, one intensity
    per frame, background added twice for noise index >= 1 (the overwrite of
    out[psf,0,f] at :303-305), proper Poisson(x*pn)/pn."""
    traj = np.asarray(traj, dtype=np.float64)
    N, T, _ = traj.shape
    if T % n != 0:
        raise Exception("T is not divisble by posPerFrame")
    if list(psf_settings) == [] or list(noise_settings) == []:
        raise Exception("No settings given")
    prm = derive_params(image_props, "psfnoise")
    P, U = prm["P"], prm["U"]
    G = P * U
    F = T // n
    npsf, nnoise = len(psf_settings), len(noise_settings)
    noise = noise if noise is not None else MeanNoise()
    pm_glob = prm["part_mean"] if part_mean_global is None else float(part_mean_global)   # module global `part_mean` :302
    tr = traj * prm["scale"]
    out = np.zeros((N, npsf, nnoise, F, P, P), dtype=F32)
    draw = prm["part_mean"] > 0.0001 and prm["part_std"] > 0.0001
    sigmas = [prm["sigma"] / float(s) for s in psf_settings]                 # :290
    pn = F32(prm["poisson"])
    bmean = F32(prm["bg_mean"])
    for s in range(N):
        seq = seq_offset + s
        zI = noise.intensity_z(seq, F)
        for f in range(F):
            xs, ys = _frame_centres(tr[s], f, n, center, U)
            if draw:
                If = F32(prm["part_mean"]) + F32(prm["part_std"]) * zI[f]      # :279
                intens = np.full(n, F32(If / F32(n)), dtype=np.float64)       # :286
                base = _accumulate_frame(xs, ys, intens, sigmas, G, U, mode)
            else:
                base = np.zeros((npsf, P, P), dtype=F32)
            out[s, :, 0, f] = base
        for i in range(npsf):
            for j in range(nnoise):
                zb, poisson = noise.pixel(seq, F * P * P, variant=1 + i * nnoise + j)
                zb = zb.reshape(F, P, P)
                bstd = F32(pm_glob * float(noise_settings[j]))                 # :302
                hi = F32(prm["bg_mean"] + 3 * float(bstd))
                bg = np.clip((bmean + bstd * zb).astype(F32), F32(0), hi)
                v = (out[s, i, 0] + bg).astype(F32)                            # :303 (out[psf,0] already noisy for j>=1)
                lam = (v * pn).astype(F32)
                out[s, i, j] = (poisson(lam) / pn).astype(F32)                # :305
    return out


def gaussian_filter_nearest(frame, sigma=0.5, truncate=4.0):
    """scipy.ndimage.gaussian_filter(frame, sigma, mode='nearest', truncate=truncate) on a float32 image, which is what
    skimage.filters.gaussian(frame, sigma=0.5) runs (helpersGeneration.py:530): radius int(truncate*sigma + 0.5), normalised
    weights exp(-k^2 / 2 sigma^2) in float64, axis 0 then axis 1, float64 accumulation, float32 result after EACH axis."""
    r = int(truncate * float(sigma) + 0.5)
    k = np.arange(-r, r + 1, dtype=np.float64)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * k * k)
    w /= w.sum()
    a = np.asarray(frame, dtype=F32)
    for axis in (0, 1):
        n = a.shape[axis]
        acc = np.zeros(a.shape, dtype=np.float64)
        for i, wk in zip(range(-r, r + 1), w):
            idx = np.clip(np.arange(n) + i, 0, n - 1)
            acc += wk * np.take(a, idx, axis=axis).astype(np.float64)
        a = acc.astype(F32)
    return a


def render_multi(traj, n, center, image_props, noise=None, seq_offset=0, mode="separable", flip_y=True):
    """trajectories_to_video_multiple_settings / trajectory_to_mult_settings (helpersGeneration.py:422-540): one intensity per
    FRAME (:505) shared by its sub-positions (:512), four outputs per frame -- noise free (:523), + clipped Gaussian background
    (:524-525), + proper Poisson(x*pn)/pn (:527), + Gaussian filter sigma 0.5 of the Poisson frame (:530).  Returns four float32
    (N, F, P, P) arrays.  The reference flips the caller's y in place (:432); `flip_y` applies that sign here."""
    traj = np.asarray(traj, dtype=np.float64)
    N, T, _ = traj.shape
    if T % n != 0:
        raise Exception("T is not divisble by posPerFrame")
    prm = derive_params(image_props, "multi")
    P, U = prm["P"], prm["U"]
    G = P * U
    F = T // n
    noise = noise if noise is not None else MeanNoise()
    tr = traj.copy()
    if flip_y:
        tr[:, :, 1] *= -1
    tr = tr * prm["scale"]
    outs = [np.zeros((N, F, P, P), dtype=F32) for _ in range(4)]
    draw = prm["part_mean"] > 0.0001 and prm["part_std"] > 0.0001
    pn, bmean, bstd = F32(prm["poisson"]), F32(prm["bg_mean"]), F32(prm["bg_std"])
    hi = F32(prm["bg_mean"] + 3 * prm["bg_std"])
    for s in range(N):
        seq = seq_offset + s
        zI = noise.intensity_z(seq, F)
        for f in range(F):
            xs, ys = _frame_centres(tr[s], f, n, center, U)
            if draw:
                If = F32(prm["part_mean"]) + F32(prm["part_std"]) * zI[f]      # :505
                intens = np.full(n, float(If) / n, dtype=np.float64)          # :512 (a float64 quotient in the reference)
                outs[0][s, f] = _accumulate_frame(xs, ys, intens, [prm["sigma"]], G, U, mode)[0]
        zb, poisson = noise.pixel(seq, F * P * P, variant=0)
        bg = np.clip((bmean + bstd * zb.reshape(F, P, P)).astype(F32), F32(0), hi)
        outs[1][s] = (outs[0][s] + bg).astype(F32)                             # :524
        outs[2][s] = (poisson((outs[1][s] * pn).astype(F32)) / pn).astype(F32)  # :527
        for f in range(F):
            outs[3][s, f] = gaussian_filter_nearest(outs[2][s, f], 0.5)        # :530
    return tuple(outs)


def create_gaussian_psf(size=9, sigma=1.3):
    """helpersGeneration.py:591-599."""
    if size % 2 == 0:
        size += 1
    ax = np.arange(-size // 2 + 1, size // 2 + 1)
    x, y = np.meshgrid(ax, ax)
    psf = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    psf /= psf.sum()
    return psf


def tv_gradient(image):
    """helpersGeneration.py:542-555 (float32 in, float32 out)."""
    grad = np.zeros_like(image)
    dx = np.diff(image, axis=1, append=image[:, -1:])
    dy = np.diff(image, axis=0, append=image[-1:, :])
    eps = 1e-8
    mag = np.sqrt(dx ** 2 + dy ** 2 + eps)
    dx_norm = dx / mag
    dy_norm = dy / mag
    grad[:, :-1] -= dx_norm[:, :-1]
    grad[:, 1:] += dx_norm[:, :-1]
    grad[:-1, :] -= dy_norm[:-1, :]
    grad[1:, :] += dy_norm[:-1, :]
    return grad


def _conv_same(a, k):
    """scipy.signal.fftconvolve(a, k, mode='same') as the direct sum it equals (float64): out[i,j] = sum a[i+c-u, j+c-v] k[u,v]."""
    a = np.asarray(a, dtype=np.float64)
    K = k.shape[0]
    c = (K - 1) // 2
    H, W = a.shape
    pad = np.zeros((H + 2 * K, W + 2 * K), dtype=np.float64)
    pad[K:K + H, K:K + W] = a
    out = np.zeros((H, W), dtype=np.float64)
    for u in range(K):
        for v in range(K):
            out += k[u, v] * pad[K + c - u:K + c - u + H, K + c - v:K + c - v + W]
    return out


def richardson_lucy_tv_iter_list(image, psf, iterations_list, tv_weight=0.01, conv=_conv_same):
    """helpersGeneration.py:571-587: Richardson-Lucy with a total-variation step, estimates after the listed (0-based)
    iterations.  `conv` is the 'same'-mode convolution (the reference uses scipy.signal.fftconvolve)."""
    image = np.clip(image, 1e-6, None)
    psf_mirror = psf[::-1, ::-1]
    estimate = np.full(image.shape, 0.5, dtype=np.float32)
    out = [None] * len(iterations_list)
    for i in range(iterations_list[-1] + 1):
        relative_blur = image / (conv(estimate, psf) + 1e-6)
        correction = conv(relative_blur, psf_mirror)
        estimate *= correction
        tv_grad = tv_gradient(estimate)
        estimate -= tv_weight * tv_grad
        estimate = np.clip(estimate, 0, 1)
        if i in iterations_list:
            out[list(iterations_list).index(i)] = estimate.copy()
    return out


def _conv_fft(a, k):
    """The reference's own call.  scipy transforms the float32 estimate in SINGLE precision (rfftn of a float32 array), so the
    reference's result carries ~1e-7 relative noise per convolution, which the multiplicative RL update amplifies: the direct
    float64 sum (_conv_same, what the CUDA kernel computes) and this agree to 2e-5 / 1e-4 / 3e-4 after 3 / 6 / 11 iterations."""
    from scipy.signal import fftconvolve
    return fftconvolve(a, k, mode="same")


def render_norm_rl(traj, n, center, image_props, rl_iterations, poisson_index=2, noise=None, seq_offset=0, mode="separable",
                   flip_y=True, conv="direct"):
    """trajs_to_vid_norm_rl (helpersGeneration.py:635-658): the four multiple-settings outputs, normalised with
    (bg_mean, bg_sigma, part_mean + bg_mean), plus the Richardson-Lucy/TV estimates of the Poisson frames (PSF sigma = 1) after
    the listed iterations.  Returns float32 (N, 4 + len(rl_iterations), F, P, P)."""
    bg_mean, bg_sigma = image_props["background_intensity"]
    part_mean, _ = image_props["particle_intensity"]
    psf = create_gaussian_psf(sigma=1)
    videos = np.stack(render_multi(traj, n, center, image_props, noise, seq_offset, mode, flip_y), axis=1)
    videos, _ = normalize_images(videos, bg_mean, bg_sigma, part_mean + bg_mean)
    to_rl = videos[:, poisson_index]
    N, F, P, _ = to_rl.shape
    rl = np.empty((N, len(rl_iterations), F, P, P), dtype=to_rl.dtype)
    for b in range(N):
        for t in range(F):
            est = richardson_lucy_tv_iter_list(to_rl[b, t], psf, list(rl_iterations), conv=_conv_fft if conv == "fft" else _conv_same)
            for k, e in enumerate(est):
                rl[b, k, t] = e
    return np.concatenate([videos, rl], axis=1)


def normalize_images(images, background_mean=None, background_sigma=None, theoretical_max=None, clip_image=False):
    """helpersGeneration.py:356-400."""
    if background_mean is None:
        background_mean = np.mean(images)
    if background_sigma is None:
        background_sigma = np.std(images)
    if theoretical_max is None:
        theoretical_max = np.max(images)
    denominator = theoretical_max - (background_mean - background_sigma)
    if denominator == 0:
        raise ValueError("Denominator in normalization is zero. Check your inputs.")
    normalized = (images - (background_mean - background_sigma)) / denominator
    if clip_image:
        normalized = np.clip(normalized, 0, 1.5)
    return normalized, (background_mean, background_sigma, theoretical_max)


def render_framerates(traj, npos_list, center, image_props, noise_factory=None, seq_offset=0,
                      mode="separable", original_npos=10):
    """trajs_to_vid_framerates (trainSettingsFramerate.py:170-202).  Returns float32
    numpy (N, len(npos_list), T//npos_list[0], P, P), zero padded.  The reference's
    in-place y flip toggles the sign on every call, so variant i sees (-1)^(i+1) * y.
    noise_factory(i) -> noise source for variant i (default MeanNoise)."""
    traj = np.asarray(traj, dtype=np.float64)
    N, T, _ = traj.shape
    flux, pstd = image_props["particle_intensity"]
    bg_mean, bg_sigma = image_props["background_intensity"][0], image_props["background_intensity"][1]
    P = int(image_props.get("output_size", DEFAULT_IMAGE_PROPS["output_size"]))
    maxF = T // npos_list[0]
    out = np.zeros((N, len(npos_list), maxF, P, P), dtype=F32)
    for i, nsub in enumerate(npos_list):
        if T % nsub != 0:
            raise Exception("T is not divisible by nPosPerFrame")
        F = T // nsub
        flux_i = flux * (nsub / original_npos)
        props_i = dict(image_props)
        props_i["particle_intensity"] = [flux_i, pstd]
        t = traj.copy()
        if i % 2 == 1:          # flipped an even number of times before this call's own flip
            t[:, :, 1] *= -1    # pre-flip so that render_v1's flip gives +y
        vid = render_v1(t, nsub, center, props_i, noise=None if noise_factory is None else noise_factory(i),
                        seq_offset=seq_offset, mode=mode, flip_y=True)
        vid, _ = normalize_images(vid, bg_mean, bg_sigma, bg_mean + flux_i)
        out[:, i, :F] = vid
    return out
