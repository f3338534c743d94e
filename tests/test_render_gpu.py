"""GPU parity tests (-m gpu) of the CUDA renderer, called through the Python mirrors which go
through the C ABI (include/mivit.h).  Checker = oracle/ (numpy) + goldens from the reference."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import render_oracle as ro
from oracle.make_golden import C3_PROPS
from oracle.noise import PhiloxNoise
from oracle.trajectory_oracle import brownian_oracle

CLEAN = dict(C3_PROPS, background_intensity=[0, 0], poisson_noise=-1)
PSF = [2, 1.75, 1.5, 1.25, 1]
NOISE = [0, 1 / 50, 1 / 25, 1 / 20, 1 / 10, 1 / 5]
PSFNOISE_PROPS = dict(C3_PROPS, particle_intensity=[5000, 500], background_intensity=[5000, 0])
FRAMERATE_PROPS = dict(C3_PROPS, output_size=13)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return (np.load(os.path.join(golden_dir, "render_inputs.npz")),
            np.load(os.path.join(golden_dir, "render_golden.npz")),
            np.load(os.path.join(golden_dir, "render_noise_stats.npz")))


@pytest.fixture(scope="module")
def gen():
    from moleculardiffusion_mivit_b200 import helpersGeneration
    return helpersGeneration


@pytest.fixture(scope="module")
def exp():
    from moleculardiffusion_mivit_b200 import experiments
    return experiments


def relmax(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("key,sl,n,center,over", [
    ("v1_p9_center", slice(0, 8), 10, True, {}),
    ("v1_p9_nocenter", slice(0, 2), 10, False, {}),
    ("v1_p13_center", slice(0, 2), 10, True, {"output_size": 13}),
    ("v1_p8_center", slice(0, 1), 10, True, {"output_size": 8}),
    ("v1_p7u10_n15", slice(0, 1), 15, True, {"output_size": 7, "upsampling_factor": 10}),
])
def test_noise_free_matches_reference_golden(gold, gen, key, sl, n, center, over):
    inp, g, _ = gold
    t = inp["traj30"][sl].copy()
    props = dict(CLEAN, **over)
    out = gen.trajectories_to_video(t, n, center, props, _mean_noise=True)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == g[key].shape
    assert np.array_equal(t[:, :, 1], -inp["traj30"][sl][:, :, 1])      # in-place y flip (reference :197)
    flat, ref = out.reshape(*out.shape[:2], -1), g[key].reshape(*out.shape[:2], -1)
    assert np.array_equal(flat.argmax(-1), ref.argmax(-1))              # index / patch layout: exact
    assert relmax(out, g[key]) < 1e-5                                   # north_star tolerance (fp32)
    orc = ro.render_v1(inp["traj30"][sl], n, center, props)
    assert relmax(out, orc) < 2e-6


@pytest.mark.parametrize("wavelength,P", [(70e-9, 9), (120e-9, 13), (40e-9, 7)])
def test_sharp_psf_far_entries_underflow_without_nan(gold, gen, wavelength, P):
    """PSF narrower than a high-resolution pixel (sigma 0.3-0.9 sub-pixels): most axis-table entries underflow and the ratio of the
    geometric recurrence of the block samples (csrc/render.cu: axis_table_v1) exceeds 2^126 there -- the clamp must keep those
    entries at 0 (0 * inf would be NaN) while the spot itself stays within the fp32 tolerance of the oracle."""
    inp, _, _ = gold
    props = dict(CLEAN, wavelength=wavelength, output_size=P)
    t = inp["traj30"][:3].copy()
    out = gen.trajectories_to_video(t, 10, True, props, _mean_noise=True)
    orc = ro.render_v1(inp["traj30"][:3], 10, True, props)
    assert np.isfinite(orc).all(), "the oracle itself underflows to NaN frames for this PSF: not the case this test is about"
    assert np.isfinite(out).all()
    assert out.max() > 0
    flat, ref = out.reshape(*out.shape[:2], -1), orc.reshape(*orc.shape[:2], -1)
    assert np.array_equal(flat.argmax(-1), ref.argmax(-1))
    assert relmax(out, orc) < 1e-5


def test_fused_normalisation_and_background(gold, gen):
    inp, g, _ = gold
    out = gen.trajectories_to_video(inp["traj30"][:2].copy(), 10, True, dict(C3_PROPS, poisson_noise=-1),
                                    normalize=(1420, 290, 6000), _mean_noise=True)
    assert np.abs(out - g["v1_p9_bgmean_norm"]).max() < 5e-6
    raw = gen.trajectories_to_video(inp["traj30"][:2].copy(), 10, True, dict(C3_PROPS, poisson_noise=-1), _mean_noise=True)
    nrm, meta = gen.normalize_images(raw, 1420, 290, 6000)
    assert meta == (1420, 290, 6000) and np.abs(nrm - out).max() < 1e-6
    with pytest.raises(ValueError, match="Denominator"):
        gen.normalize_images(raw, 10, 0, 10)


def test_errors_and_edge_cases(gen, exp):
    with pytest.raises(Exception, match="T is not divisble by posPerFrame"):
        gen.trajectories_to_video(np.zeros((1, 25, 2)), 10, True, CLEAN)
    with pytest.raises(Exception, match="No settings given"):
        exp.trajs_to_vid_psf_noise(np.zeros((1, 20, 2)), 10, True, PSFNOISE_PROPS, [], [])
    out = gen.trajectories_to_video(np.zeros((0, 20, 2)), 10, True, CLEAN)
    assert out.shape == (0, 2, 9, 9)
    # particle_std <= 1e-4 -> no particle is drawn (reference :299): background only
    props = dict(C3_PROPS, particle_intensity=[4580, 0], poisson_noise=-1)
    out = gen.trajectories_to_video(np.zeros((1, 20, 2)), 10, True, props, _mean_noise=True)
    assert np.all(out == 1420)
    # a spot tens of pixels outside the image underflows in the reference and yields NaN frames
    far = np.zeros((1, 10, 2)); far[:, :, 0] = 1e4
    out = gen.trajectories_to_video(far, 10, False, CLEAN, _mean_noise=True)
    assert np.isnan(out).all()


def test_noisy_v1_matches_philox_oracle(gold, gen):
    inp, _, _ = gold
    t = inp["traj30"][:3]
    out = gen.trajectories_to_video(t.copy(), 10, True, C3_PROPS, seed=1234, seq_offset=5)
    orc = ro.render_v1(t, 10, True, C3_PROPS, noise=PhiloxNoise(1234), seq_offset=5)
    bad = np.abs(out - orc) > 1e-4 * np.abs(orc) + 1e-2
    assert bad.mean() < 2e-3, bad.mean()        # same counter stream -> same draws (rare fp ties differ)


def test_noisy_v1_statistics_vs_reference(gold, gen):
    inp, _, st = gold
    v = np.stack([gen.trajectories_to_video(inp["traj30"].copy(), 10, True, C3_PROPS, seed=50 + r)
                  for r in range(40)]).astype(np.float64)
    assert abs(v.mean() - st["v1_mean"]) / st["v1_mean"] < 2e-3
    assert abs(v.std() - st["v1_std"]) / st["v1_std"] < 5e-3
    assert np.abs(v.mean(axis=(0, 1, 2)) - st["v1_pix_mean"]).max() / st["v1_pix_mean"].max() < 5e-3
    assert np.abs(v.std(axis=(0, 1, 2)) / st["v1_pix_std"] - 1).max() < 6e-2   # 9600 samples/pixel: ~1% sampling error each, max over 81 pixels
    q = np.linspace(0, 1, len(st["v1_quantiles"]))
    s = np.sort(v.ravel())
    ks = np.abs(np.searchsorted(s, st["v1_quantiles"], side="right") / s.size - q).max()
    assert ks < 5e-3, ks


def test_noisy_p13_framerate_statistics_vs_reference(golden_dir, gold, gen):
    """The props bench.py renders (Framerate experiment, P = 13, n = 10): first / second moments, per-pixel maps and the KS
    statistic of the CUDA renderer against the reference's own np.random statistics (oracle/make_golden_r2.py)."""
    inp, _, _ = gold
    st = np.load(os.path.join(golden_dir, "render_noise_stats_p13.npz"))
    v = np.stack([gen.trajectories_to_video(inp["traj30"].copy(), 10, True, FRAMERATE_PROPS, seed=900 + r)
                  for r in range(24)]).astype(np.float64)
    assert abs(v.mean() - st["mean"]) / st["mean"] < 2e-3
    assert abs(v.std() - st["std"]) / st["std"] < 5e-3
    assert np.abs(v.mean(axis=(0, 1, 2)) - st["pix_mean"]).max() / st["pix_mean"].max() < 6e-3
    assert np.abs(v.std(axis=(0, 1, 2)) / st["pix_std"] - 1).max() < 8e-2
    q = np.linspace(0, 1, len(st["quantiles"]))
    srt = np.sort(v.ravel())
    ks = np.abs(np.searchsorted(srt, st["quantiles"], side="right") / srt.size - q).max()
    assert ks < 5e-3, ks
    orc = ro.render_v1(inp["traj30"][:2], 10, True, FRAMERATE_PROPS, noise=PhiloxNoise(900))
    out = gen.trajectories_to_video(inp["traj30"][:2].copy(), 10, True, FRAMERATE_PROPS, seed=900)
    bad = np.abs(out - orc) > 1e-4 * np.abs(orc) + 1e-2
    assert bad.mean() < 2e-3, bad.mean()                                  # odd P: the last pair of every row holds one pixel


def test_large_lambda_falls_back_to_ptrs(gold, gen):
    """poisson_noise > 380 is outside the alias table: PTRS on the pixel's own uniform stream, same draws as the oracle."""
    inp, _, _ = gold
    props = dict(C3_PROPS, poisson_noise=1000)
    out = gen.trajectories_to_video(inp["traj30"][:2].copy(), 10, True, props, seed=31)
    orc = ro.render_v1(inp["traj30"][:2], 10, True, props, noise=PhiloxNoise(31))
    bad = np.abs(out - orc) > 1e-4 * np.abs(orc) + 1e-2
    assert bad.mean() < 3e-3, bad.mean()
    clean = gen.trajectories_to_video(inp["traj30"][:2].copy(), 10, True, dict(props, poisson_noise=-1), seed=31)
    r = (out / clean).ravel()
    assert abs(r.mean() - 1) < 2e-3 and abs(r.std() - 1 / np.sqrt(1000)) < 2e-3


def test_psfnoise(gold, exp):
    inp, g, st = gold
    t = inp["traj20"][:1]
    out = exp.trajs_to_vid_psf_noise(t.copy(), 10, True, PSFNOISE_PROPS, PSF, NOISE, _mean_noise=True)
    assert out.shape == (1, 5, 6, 20, 9, 9) and out.dtype == np.float32
    assert relmax(out, g["psfnoise_mean"]) < 1e-5
    tt = inp["traj20"].copy()
    out = exp.trajs_to_vid_psf_noise(tt, 10, True, PSFNOISE_PROPS, PSF, NOISE, seed=99)
    assert np.array_equal(tt, inp["traj20"])                              # no y flip in this variant
    orc = ro.render_psfnoise(inp["traj20"], 10, True, PSFNOISE_PROPS, PSF, NOISE, noise=PhiloxNoise(99))
    bad = np.abs(out - orc) > 1e-4 * np.abs(orc) + 0.011                  # counts are multiples of 1/100
    assert bad.mean() < 5e-3, bad.mean()
    v = np.stack([exp.trajs_to_vid_psf_noise(inp["traj20"].copy(), 10, True, PSFNOISE_PROPS, PSF, NOISE, seed=7 + r)
                  for r in range(12)]).astype(np.float64)
    m, s = v.mean(axis=(0, 1, 4, 5, 6)), v.std(axis=(0, 1, 4, 5, 6))
    assert np.abs(m / st["psf_mean"] - 1).max() < 3e-3                    # incl. the 5236 / 10235 double-background quirk
    assert np.abs(s / st["psf_std"] - 1).max() < 2e-2
    qq = np.linspace(0, 1, st["psf_quantiles"].shape[-1])
    for i in range(5):
        for j in range(6):
            srt = np.sort(v[:, :, i, j].ravel())
            ks = np.abs(np.searchsorted(srt, st["psf_quantiles"][i, j], side="right") / srt.size - qq).max()
            assert ks < 0.02, (i, j, ks)


def test_framerates(gold, exp):
    import torch
    inp, g, _ = gold
    t = inp["traj30b"][:1].copy()
    out = exp.trajs_to_vid_framerates(t, [5, 10, 15, 20, 30, 50], True, FRAMERATE_PROPS, _mean_noise=True)
    assert isinstance(out, torch.Tensor) and out.device.type == "cpu" and tuple(out.shape) == (1, 6, 60, 13, 13)
    assert np.array_equal(t, inp["traj30b"][:1])                          # six flips restore the caller's array
    assert np.abs(out.numpy() - g["framerate_mean"]).max() < 5e-6
    assert torch.all(out[:, 5, 6:] == 0)


def test_brownian_source(gen):
    gm = [1, 3, 5, 7, 9, 10.2]
    import ctypes
    import torch
    from moleculardiffusion_mivit_b200 import _lib
    N, T = 64, 300
    traj = torch.empty((N, T, 2), dtype=torch.float64, device="cuda")
    D = torch.empty((N,), dtype=torch.float32, device="cuda")
    m = np.asarray(gm, dtype=np.float32); v = np.ones(6, dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    _lib.check(_lib.lib().mivit_brownian(N, T, m.ctypes.data_as(fp), v.ctypes.data_as(fp), 6, 100.0, 42, 10,
                                         _lib.ptr(traj), _lib.ptr(D), _lib.current_stream()))
    ot, oD = brownian_oracle(N, T, gm, [1.0] * 6, 100.0, 42, seq_offset=10)
    assert np.abs(D.cpu().numpy() - oD).max() < 1e-5
    assert np.abs(traj.cpu().numpy() - ot).max() < 1e-6
    tr = gen.brownian_motion(2000, 30, 10, 2.0, 1.0, seed=1)
    assert abs(np.diff(tr, axis=1).std() - np.sqrt(2 * 2.0 / 10)) < 0.01


def test_full_size_properties(gen):
    """BASELINE-size batch: sharding invariance (global sequence ids), determinism, linearity."""
    import torch
    from moleculardiffusion_mivit_b200.helpersGeneration import derive_render_params, render_device
    N, T = 4096, 300
    traj = gen.brownian_motion(N, 30, 10, [1, 3, 5, 7, 9, 10.2], 1.0, seed=3, D_var=1.0, div=100.0, return_device=True)
    prm = derive_render_params(FRAMERATE_PROPS, 10, True)
    a = render_device(traj, prm, seed=77)
    b = render_device(traj, prm, seed=77)
    assert torch.equal(a, b)
    h0 = render_device(traj[:1000].contiguous(), prm, seed=77, seq_offset=0)
    h1 = render_device(traj[1000:].contiguous(), prm, seed=77, seq_offset=1000)
    assert torch.equal(torch.cat([h0, h1]), a)
    assert torch.isfinite(a).all()
    clean = derive_render_params(dict(FRAMERATE_PROPS, background_intensity=[0, 0], poisson_noise=-1), 10, True)
    clean.mean_noise = 1
    c1 = render_device(traj, clean, seed=0)
    clean.part_mean *= 2
    c2 = render_device(traj, clean, seed=0)
    assert torch.allclose(c2, 2 * c1, rtol=1e-6, atol=0)
    # centred particle: flux conserved up to the part of the PSF that falls outside the 13x13 window
    tot = c1.sum(dim=(2, 3))
    assert (tot > 0).all()


# ---- renderer fused with the Linear / CNN frame embedding (north_star (1), BASELINE configs[2]) --------------------
@pytest.mark.parametrize("kind,P,E,n", [("linear", 9, 64, 10), ("cnn", 9, 32, 10), ("linear", 13, 128, 10), ("cnn", 7, 64, 15),
                                        ("linear", 15, 32, 30)])
@pytest.mark.parametrize("noisy", [False, True])
def test_fused_render_embed_matches_render_then_embed(gold, gen, kind, P, E, n, noisy):
    """emb = W . frame + b computed inside the render kernel == the reference order of operations
    (trajectories_to_video -> normalize_images -> embedding, helpers/models.py:160-164 / :188-197) on the SAME frames:
    noise-free frames are checked against the reference golden pipeline through the unfused renderer, noisy ones use
    identical Philox streams in both kernels."""
    import torch
    from moleculardiffusion_mivit_b200 import models as M
    inp, _, _ = gold
    props = dict(C3_PROPS if noisy else CLEAN, output_size=P, upsampling_factor=5)
    norm = (1420, 290, 6000)
    torch.manual_seed(P * 100 + E)
    emb_mod = (M.LinearProjectionEmbedding if kind == "linear" else M.CNNEmbedding)(P, E)
    t1, t2 = inp["traj30"][:6].copy(), inp["traj30"][:6].copy()
    frames = gen.trajectories_to_video(t1, n, True, props, seed=11, normalize=norm, _mean_noise=not noisy)
    emb, fr2 = gen.trajectories_to_embeddings(t2, n, emb_mod, True, props, seed=11, normalize=norm, return_frames=True,
                                              _mean_noise=not noisy)
    assert np.array_equal(t1, t2)                                            # same in-place y flip
    assert emb.shape == (6, 300 // n, E) and emb.dtype == torch.float32 and emb.is_cuda
    assert np.array_equal(fr2.cpu().numpy(), frames)                         # bit-identical frames (layout + values)
    W = (emb_mod.proj.weight if kind == "linear" else emb_mod.conv.weight).detach().double().reshape(E, P * P)
    b = (emb_mod.proj.bias if kind == "linear" else emb_mod.conv.bias).detach().double()
    ref = torch.from_numpy(frames).double().reshape(6, -1, P * P) @ W.t() + b
    err = (emb.cpu().double() - ref).abs().max().item()
    assert err < 1e-5 * max(1.0, ref.abs().max().item()), err                # fp32 dot product of P*P terms
    only = gen.trajectories_to_embeddings(inp["traj30"][:6].copy(), n, emb_mod, True, props, seed=11, normalize=norm,
                                          _mean_noise=not noisy)
    assert torch.equal(only, emb)                                            # frames_out = NULL path


@pytest.mark.parametrize("P,E,n,noisy", [(9, 64, 10, True), (13, 64, 10, True), (13, 128, 10, False), (7, 32, 15, True), (15, 39, 30, True)])
def test_fused_embed_weight_gradient_rerenders_identical_frames(gold, gen, P, E, n, noisy):
    """mivit_render_embed_linear_wgrad: dW = sum_frames demb (x) frame and db = sum_frames demb with the frames re-rendered inside
    the kernel, against the same products formed from the frames the forward kernel can write out."""
    import ctypes
    import torch
    from moleculardiffusion_mivit_b200 import _lib, models as M
    from moleculardiffusion_mivit_b200.helpersGeneration import derive_render_params
    inp, _, _ = gold
    props = dict(C3_PROPS if noisy else CLEAN, output_size=P)
    emb_mod = M.LinearProjectionEmbedding(P, E)
    t = inp["traj30"].copy()
    emb, frames = gen.trajectories_to_embeddings(t, n, emb_mod, True, props, seed=77, seq_offset=3, normalize=(1420, 290, 6000),
                                                 return_frames=True, _mean_noise=not noisy)
    N, Fr = frames.shape[0], frames.shape[1]
    demb = torch.randn(N, Fr, E, generator=torch.Generator().manual_seed(4)).cuda()
    prm = derive_render_params(props, n, True)
    prm.flip_y, prm.mean_noise = 0, int(not noisy)
    prm.normalize, prm.norm_sub, prm.norm_div = 1, 1130.0, 4870.0
    tdev = torch.from_numpy(t).cuda()                                     # already flipped by the call above
    dW = torch.zeros(E, P * P, device="cuda")
    db = torch.zeros(E, device="cuda")
    _lib.check(_lib.lib().mivit_render_embed_linear_wgrad(_lib.ptr(tdev), N, 300, ctypes.byref(prm), 77, 3, _lib.ptr(demb), E,
                                                          _lib.ptr(dW), _lib.ptr(db), _lib.current_stream()))
    ref_dW = demb.double().reshape(-1, E).t() @ frames.double().reshape(-1, P * P)
    ref_db = demb.double().sum(dim=(0, 1))
    assert (dW.double() - ref_dW).abs().max().item() < 2e-5 * ref_dW.abs().max().item()
    assert (db.double() - ref_db).abs().max().item() < 2e-5 * ref_db.abs().max().item()


def test_fused_render_embed_rejects_deepresnet(gen, gold):
    from moleculardiffusion_mivit_b200 import models as M
    inp, _, _ = gold
    with pytest.raises(TypeError, match="LinearProjectionEmbedding and CNNEmbedding"):
        gen.trajectories_to_embeddings(inp["traj30"][:1].copy(), 10, M.DeepResNetEmbedding(9, 64), True, C3_PROPS)
    with pytest.raises(Exception, match="T is not divisble by posPerFrame"):
        gen.trajectories_to_embeddings(inp["traj30"][:1].copy(), 7, M.LinearProjectionEmbedding(9, 64), True, C3_PROPS)


# ---- trajectories_to_video_multiple_settings (helpersGeneration.py:422-540) -------------------------------
@pytest.mark.parametrize("key,sl,n,center,over", [("p9_center", slice(0, 3), 10, True, {}),
                                                   ("p13_nocenter_n15", slice(3, 4), 15, False, {"output_size": 13})])
def test_multi_settings_noise_free_matches_reference_golden(golden_dir, gold, gen, key, sl, n, center, over):
    inp, _, _ = gold
    g = np.load(os.path.join(golden_dir, "render_multi_golden.npz"))
    t = inp["traj30"][sl].copy()
    outs = gen.trajectories_to_video_multiple_settings(t, n, center, dict(C3_PROPS, **over), _mean_noise=True)
    assert np.array_equal(t[:, :, 1], -inp["traj30"][sl][:, :, 1])          # the caller's y is flipped in place (:432)
    assert len(outs) == 4
    for name, a in zip(("none", "gauss", "poisson", "filter"), outs):
        ref = g["%s/%s" % (key, name)]
        assert a.dtype == np.float32 and a.shape == ref.shape
        assert relmax(a, ref) < 1e-5, (name, relmax(a, ref))
        # index / patch layout: the brightest pixel of every frame is where the reference puts it
        assert (a.reshape(a.shape[0], a.shape[1], -1).argmax(-1) == ref.reshape(ref.shape[0], ref.shape[1], -1).argmax(-1)).mean() > 0.99


def test_multi_settings_noisy_matches_philox_oracle_and_errors(gold, gen):
    inp, _, _ = gold
    t = inp["traj30"][:3]
    outs = gen.trajectories_to_video_multiple_settings(t.copy(), 10, True, C3_PROPS, seed=4321, seq_offset=2)
    orc = ro.render_multi(t, 10, True, C3_PROPS, noise=PhiloxNoise(4321), seq_offset=2)
    for name, a, b in zip(("none", "gauss", "poisson", "filter"), outs, orc):
        bad = np.abs(a - b) > 1e-4 * np.abs(b) + 1e-2
        # same counter streams -> same draws; a Poisson count that differs through an fp tie also moves its 5x5 filter footprint
        assert bad.mean() < (2e-2 if name == "filter" else 3e-3), (name, bad.mean())
    # defaults of this variant: poisson_noise = 1 (:455); the T % n check comes after the in-place flip
    t2 = np.zeros((1, 25, 2))
    t2[:, :, 1] = 1.0
    with pytest.raises(Exception, match="T is not divisble by posPerFrame"):
        gen.trajectories_to_video_multiple_settings(t2, 10, True, C3_PROPS)
    assert np.all(t2[:, :, 1] == -1.0)


def test_norm_rl_matches_reference_golden_and_oracle(golden_dir, gold, gen):
    """trajs_to_vid_norm_rl (helpersGeneration.py:635-658): normalised four outputs + Richardson-Lucy/TV estimates."""
    inp, _, _ = gold
    ref = np.load(os.path.join(golden_dir, "render_norm_rl_golden.npz"))["out"]
    t = inp["traj30"][:2].copy()
    out = gen.trajs_to_vid_norm_rl(t, 10, True, C3_PROPS, [2, 5, 10], _mean_noise=True)
    assert np.array_equal(t[:, :, 1], -inp["traj30"][:2][:, :, 1])              # the in-place y flip of the inner call
    assert out.shape == ref.shape == (2, 7, 30, 9, 9) and out.dtype == np.float32
    assert relmax(out[:, :4], ref[:, :4]) < 1e-5
    # The multiplicative RL update amplifies 1e-7-level differences of its INPUT (float32 renderer vs float64 reference frames,
    # and the reference's own single-precision FFT) by ~10x every few iterations: bounds per kept iteration (3, 6, 11 iterations).
    # The kernel's arithmetic itself is checked on identical inputs in test_rl_tv_helpers_match_oracle.
    orc = ro.render_norm_rl(inp["traj30"][:2], 10, True, C3_PROPS, [2, 5, 10])   # direct float64 convolution, like the kernel
    for k, tol in ((4, 2e-4), (5, 6e-4), (6, 2e-3)):
        assert np.abs(out[:, k] - orc[:, k]).max() < tol, (k, np.abs(out[:, k] - orc[:, k]).max())
        assert np.abs(out[:, k] - ref[:, k]).max() < tol, (k, np.abs(out[:, k] - ref[:, k]).max())
        assert np.abs(out[:, k] - ref[:, k]).mean() < 2e-6, (k, np.abs(out[:, k] - ref[:, k]).mean())


def test_rl_tv_helpers_match_oracle(gen):
    import torch
    rng = np.random.default_rng(3)
    imgs = rng.uniform(0.0, 1.2, size=(3, 4, 9, 9)).astype(np.float32)
    imgs[0, 0, 2, 3] = -0.5                                                      # clipped to 1e-6 (:558)
    psf = gen.create_gaussian_psf(sigma=1)
    assert psf.shape == (9, 9) and abs(psf.sum() - 1.0) < 1e-12 and np.allclose(psf, ro.create_gaussian_psf(sigma=1), rtol=0, atol=0)
    got = gen.apply_rl_tv_tensor_iter_list(imgs, psf, [0, 3, 7])
    assert got.shape == (3, 3, 4, 9, 9) and got.dtype == np.float32
    for b in range(3):
        for s in range(4):
            want = ro.richardson_lucy_tv_iter_list(imgs[b, s], psf, [0, 3, 7])
            for k in range(3):
                assert np.abs(got[b, k, s] - want[k]).max() < 5e-5, (b, s, k, np.abs(got[b, k, s] - want[k]).max())
    one = gen.richardson_lucy_tv(imgs[1, 2], psf, iterations=4)                  # 4 iterations = list index 3
    assert np.abs(one - got[1, 1, 2]).max() == 0.0
    outs = np.zeros((2, 9, 9), np.float32)
    last = gen.richardson_lucy_tv_iter_list(imgs[1, 2], psf, [3, 7], outs)
    assert np.array_equal(outs[0], got[1, 1, 2]) and np.array_equal(last, got[1, 2, 2])
    tens = gen.apply_rl_tv_tensor(torch.from_numpy(imgs), psf, n_iters=8)
    assert tuple(tens.shape) == (3, 4, 9, 9) and np.array_equal(tens.numpy(), got[:, 2])
    with pytest.raises(AssertionError, match="9x9"):
        gen.apply_rl_tv_tensor(torch.zeros(1, 1, 8, 8), psf)


def test_generate_traj_and_videos_brownian(gen):
    """helpersGeneration.py:402-417: trajectories from the device Brownian source (D ~ N(mean, var)) rendered centred."""
    props = dict(C3_PROPS, trajectory_unit=100)     # 1 trajectory unit = 1 pixel: D = 0.05 px^2 per sub-step stays in the 9 px frame
    vids, D = gen.generateTrajAndVideosBrownian([0.05, 1e-4], 64, 30, 10, props, seed=11)
    assert vids.shape == (64, 30, 9, 9) and vids.dtype == np.float32 and D.shape == (64,)
    assert np.all(D > 0) and abs(float(D.mean()) - 0.05) < 0.01 and 0.004 < float(D.std()) < 0.02
    assert np.isfinite(vids).all() and float(vids.mean()) > 1420.0     # background + particle
    vids2, D2 = gen.generateTrajAndVideosBrownian([0.05, 1e-4], 64, 30, 10, props, seed=11)
    assert np.array_equal(vids, vids2) and np.array_equal(D, D2)       # counter-based streams: reproducible
