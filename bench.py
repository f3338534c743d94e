#!/usr/bin/env python
"""bench.py -- benchmarks of the MiViT hot path on B200.

Metric (BASELINE.json): synthetic sequences/sec of  render + ViT train step;  render GB/s.

    python bench.py --gpus 1 --steps 20 --warmup 5                       # headline: BASELINE configs[1]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W   # configs[4]
    python bench.py --impl reference ...     # the reference algorithm on the host CPU cores (oracle port)
    python bench.py --config psfnoise_render | embeddings_linear_n | embeddings_cnn_n | embeddings_deepcnn_n | imagesfeatures

Workloads (`--config`; `config.workload` of the JSON line names the one that ran):
  framerate (default, BASELINE configs[1] "Framerate experiment: 30-frame sequences with full ViT training step on 1 B200"; with
      --gpus N it is configs[4], the data-parallel run): per step and per GPU, B Brownian trajectories (T=300 sub-steps, D groups of
      trainModelsFramerate.py:45) are generated on the device, rendered to 30 frames of 13x13 pixels with the Framerate
      experiment's image_props (n=10 sub-positions per frame, background + Poisson noise, fused normalisation), and pushed
      through one full training step (forward, MSE, backward, AdamW lr=1e-4) of the DeepResNet-embedding ViT (E64/H4/HD128/L6,
      regression token, no pos-enc, 506 081 parameters).
  psfnoise_render (configs[0]): trajs_to_vid_psf_noise of trainSettingsPSFNoise.py (T=200 -> 5 PSF x 6 noise x 20 frames of 9x9,
      197.6 KB per trajectory) + eval-mode ViT forward + MSE on each of the 30 variants (trainModelsPSFNoise.py:206-238).
  embeddings_{linear,cnn,deepcnn}_n (configs[2]): the Embeddings experiment's models (P=9, F=30, pos-enc on) trained FROM THE
      TRAJECTORIES: renderer fused with the frame embedding, frames re-rendered for the embedding's weight gradient
      (12 480 algorithmic bytes per sequence for linear / cnn); deepcnn renders to a frame buffer first.
  imagesfeatures (configs[3]): P=9, F=30 DeepResNet ViT + the 25 trajectory features (late fusion), features computed on the GPU
      from the frame-averaged trajectories (create_video_and_feature_pairs, helpersGeneration.py:674-719).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same metric through the public host-buffer
API (pinned H2D of the trajectories, D2H of the loss every step).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_GROUPS = [1, 3, 5, 7, 9, 10.2]          # TrainingDs_list means (variance 1), trainModelsFramerate.py:45
D_MAX = 10.0
BG_MEAN, BG_SIGMA = 1420, 290
PART_MEAN, PART_STD = 6000 - BG_MEAN, 500


def _props(P, **over):
    d = {"particle_intensity": [PART_MEAN, PART_STD], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3,
         "resolution": 100e-9, "output_size": P, "upsampling_factor": 5, "background_intensity": [BG_MEAN, BG_SIGMA],
         "poisson_noise": 100, "trajectory_unit": 1200}
    d.update(over)
    return d


def _model_cfg(embedding, size="n", **over):
    E, H, HD, L = {"s": (32, 2, 64, 3), "n": (64, 4, 128, 6), "b": (128, 8, 256, 12)}[size]
    d = dict(embedding=embedding, embed_dim=E, num_heads=H, hidden_dim=HD, num_layers=L, activation="relu",
             use_pos_encoding=False, use_regression_token=True)
    d.update(over)
    return d


# name -> workload description.  flops_per_seq: algorithmic training FLOPs, render_bytes_per_seq: 8d too.
WORKLOADS = {
    "framerate": dict(kind="train", P=13, npos=10, frames=30, props=_props(13), model=_model_cfg("deepresnet"),
                      flops_per_seq=8.773e9, from_traj=False, batch=1024,
                      text="framerate_P13_F30_deepcnn_n (BASELINE configs[1]): on-device Brownian T=300 -> render 30x13x13 (n=10, "
                           "bg+Poisson noise, normalised) -> DeepResNet-ViT E64/H4/HD128/L6 fwd+MSE+bwd+AdamW"),
    "psfnoise_render": dict(kind="psfnoise", P=9, npos=10, frames=20, batch=256,
                            props=_props(9, particle_intensity=[5000, 500], background_intensity=[5000, 0]),
                            psf=[2, 1.75, 1.5, 1.25, 1], noise=[0, 1 / 50, 1 / 25, 1 / 20, 1 / 10, 1 / 5],
                            model=_model_cfg("deepresnet"), flops_per_seq=30 * 939.2e6,
                            text="psfnoise_render (BASELINE configs[0]): on-device Brownian T=200 -> trajs_to_vid_psf_noise 5 PSF x 6 "
                                 "noise x 20 frames of 9x9 (197.6 KB per trajectory) -> eval-mode DeepResNet-ViT forward + MSE on each "
                                 "of the 30 variants"),
    "imagesfeatures": dict(kind="train", P=9, npos=10, frames=30, props=_props(9), batch=1024, from_traj=False, features=True,
                           model=_model_cfg("deepresnet", use_global_features=True, fusion_type="late"), flops_per_seq=4.227e9,
                           text="imagesfeatures_P9_F30_deepcnn_n_late (BASELINE configs[3]): on-device Brownian T=300 -> render 30x9x9 "
                                "+ 25 trajectory features on the GPU -> DeepResNet-ViT with late feature fusion fwd+MSE+bwd+AdamW"),
}
for _emb, _fl in (("linear", 0.042e9), ("cnn", 0.042e9), ("deepresnet", 4.227e9)):
    _nm = "embeddings_%s_n" % ("deepcnn" if _emb == "deepresnet" else _emb)
    WORKLOADS[_nm] = dict(kind="train", P=9, npos=10, frames=30, props=_props(9), batch=1024 if _emb == "deepresnet" else 4096,
                          model=_model_cfg(_emb, use_pos_encoding=True), flops_per_seq=_fl, from_traj=True,
                          text="%s_P9_F30 (BASELINE configs[2]): on-device Brownian T=300 -> training step FROM THE TRAJECTORIES (%s) "
                               "-> ViT E64/H4/HD128/L6 + pos-enc fwd+MSE+bwd+AdamW" % (
                                   _nm, "renderer fused with the frame embedding; frames re-rendered for its weight gradient, never "
                                   "in HBM" if _emb != "deepresnet" else "rendered into a frame buffer: BatchNorm needs all frames"))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_tflops_burst": d.get("bf16_tflops"),
                "source": "MEASURED_PEAKS.json (hbm_gbs, bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback of B200_PROFILING.md (6.65 TB/s, ~1.4 PFLOP/s sustained)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons DURING the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        load = s[len(s) // 4:] if s else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------ reference arm
def _render_slice(args):
    """Worker of the process pool: the literal numpy renderer on a slice of the step's trajectories."""
    traj, seed, name = args
    from oracle import render_oracle as ro
    from oracle.noise import NumpyNoise
    w = WORKLOADS[name]
    if w["kind"] == "psfnoise":
        return ro.render_psfnoise(traj, w["npos"], True, w["props"], w["psf"], w["noise"], noise=NumpyNoise(seed), mode="literal")
    vid = ro.render_v1(traj, w["npos"], True, w["props"], noise=NumpyNoise(seed), mode="literal")
    vid, _ = ro.normalize_images(vid, BG_MEAN, BG_SIGMA, BG_MEAN + PART_MEAN)
    return vid


def cpu_reference_step(name, n_seq, state, pool=None):
    """One bounded sample of the reference algorithm on the host: the literal numpy renderer (oracle/render_oracle.py) +
    normalisation (+ the 25 features, oracle/features_oracle.py) + the fp32 PyTorch model step (oracle/vit_oracle.py, torch
    intra-op threads = all cores).  pool = None: the renderer runs in ONE process, as the reference does (its
    use_multiprocessing branch raises TypeError); with a process pool the sequences of the step are rendered on all cores.
    Returns (t_render, t_model)."""
    import numpy as np
    import torch
    from oracle import vit_oracle as vo
    from oracle.trajectory_oracle import brownian_oracle
    w = WORKLOADS[name]
    T = w["npos"] * w["frames"]
    traj, D = brownian_oracle(n_seq, T, D_GROUPS, [1.0] * len(D_GROUPS), 100.0, seed=state["step"], seq_offset=0)
    t0 = time.perf_counter()
    if pool is None:
        vid = _render_slice((traj, state["step"], name))
    else:
        parts = np.array_split(np.arange(n_seq), min(n_seq, pool._processes))
        vid = np.concatenate(pool.map(_render_slice, [(traj[p], state["step"] * 1000 + i, name) for i, p in enumerate(parts)]))
    feats = None
    if w.get("features"):
        from oracle import features_oracle as fo
        flipped = traj.copy()
        flipped[:, :, 1] *= -1                              # create_video_and_feature_pairs sees the flipped trajectories
        feats = torch.from_numpy(np.nan_to_num(fo.features_batch(fo.average_frames(flipped, w["npos"]), 1.0), nan=0.0, posinf=0.0,
                                               neginf=0.0).astype(np.float32))
    t1 = time.perf_counter()
    y = torch.from_numpy((D / D_MAX).astype(np.float32)).unsqueeze(-1)
    if w["kind"] == "psfnoise":
        with torch.no_grad():
            for i in range(vid.shape[1]):
                for j in range(vid.shape[2]):
                    pred = vo.forward(state["sd"], state["cfg"], torch.from_numpy(np.ascontiguousarray(vid[:, i, j])), training=False)
                    torch.nn.functional.mse_loss(pred * D_MAX, y * D_MAX)
    else:
        x = torch.from_numpy(np.ascontiguousarray(vid))
        vo.train_step(state["sd"], state["opt"], state["cfg"], x, y, feats)
    t2 = time.perf_counter()
    state["step"] += 1
    return t1 - t0, t2 - t1


def _build_model(w, device=None):
    """Random-init model of the workload through this package's mirror classes (same constructors / init order as the reference)."""
    import torch
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import models as M
    m = w["model"]
    cls = {"deepresnet": M.DeepResNetEmbedding, "linear": M.LinearProjectionEmbedding, "cnn": M.CNNEmbedding}[m["embedding"]]
    torch.manual_seed(0)
    return M.GeneralTransformer(cls, {"patch_size": w["P"], "embed_dim": m["embed_dim"]}, m["embed_dim"], m["num_heads"],
                                m["hidden_dim"], m["num_layers"], M.MLPHead, F.relu, 0.0, m["use_pos_encoding"],
                                m["use_regression_token"], True, m.get("use_global_features", False), m.get("fusion_type", "early"),
                                25 if m.get("use_global_features") else None)


def cpu_reference_state(name):
    """Random-init weights of the workload's architecture (the mirror module's constructor only builds torch parameters on the
    host; nothing of the CUDA library runs) shared with the oracle through the state dict."""
    from oracle import vit_oracle as vo
    w = WORKLOADS[name]
    sd = {k: v.detach().clone() for k, v in _build_model(w).state_dict().items()}
    return {"sd": sd, "opt": vo.new_opt_state(sd), "cfg": dict(w["model"]), "step": 0}


def _cpu_baseline(name, n_seq, reps, warm):
    import multiprocessing as mp
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = cpu_reference_state(name)
    a, b = cpu_reference_step(name, n_seq, st)                     # as written: single-process renderer
    as_written = n_seq / (a + b)
    pool = mp.get_context("fork").Pool(cores)
    try:
        for _ in range(warm):
            cpu_reference_step(name, n_seq, st, pool)
        t0 = time.perf_counter()
        tr = tt = 0.0
        for _ in range(reps):
            a, b = cpu_reference_step(name, n_seq, st, pool)
            tr += a; tt += b
        el = time.perf_counter() - t0
    finally:
        pool.close()
    val = n_seq * reps / (tr + tt)
    w = WORKLOADS[name]
    what = "eval forward + MSE on 30 variants" if w["kind"] == "psfnoise" else "fp32 torch train step"
    sample = ("%d steps x %d sequences of the same workload: literal numpy renderer sharded over %d processes %.1f ms/seq + %s "
              "(%d intra-op threads) %.1f ms/seq; as written (renderer in one process, like the reference's loop): %.1f sequences/s" %
              (reps, n_seq, cores, 1e3 * tr / (n_seq * reps), what, cores, 1e3 * tt / (n_seq * reps), as_written))
    return {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "as_written_value": as_written, "sample": sample}, \
        1e3 * (tr + tt) / reps, el


def run_reference(args, rank):
    """Reference arm: the reference algorithm on ALL host cores -- the renderer sharded over a process pool (one process per core,
    each running the literal per-sequence loop), the model step with torch using every core.  The as-written figure (renderer in
    one process, like the reference's own loop) is reported beside it."""
    if rank != 0:
        return
    w = WORKLOADS[args.config]
    n_seq = args.cpu_seqs if w["kind"] != "psfnoise" else max(4, args.cpu_seqs // 4)
    cb, ms, el = _cpu_baseline(args.config, n_seq, args.steps, max(args.warmup, 1) if args.steps < 10 else 2)
    line = {"impl": "reference", "metric": "synthetic sequences/sec (render+ViT train step)", "value": cb["value"], "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["text"].split(":")[0], "batch_per_step": n_seq}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": el}
    print(json.dumps(line))


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from moleculardiffusion_mivit_b200 import _lib, experiments as X, helpersFeatures as HF
    from moleculardiffusion_mivit_b200.helpersGeneration import derive_render_params, render_device
    from moleculardiffusion_mivit_b200.models import TrajectorySource
    from moleculardiffusion_mivit_b200.training import MiViTTrainer

    w = WORKLOADS[args.config]
    P, NPOS, NFRAMES = w["P"], w["npos"], w["frames"]
    T = NPOS * NFRAMES
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL announces its version on STDOUT at the default debug level; rank 0's stdout must be the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B = args.batch if args.batch > 0 else w["batch"]
    model = _build_model(w).cuda()
    psf_mode = w["kind"] == "psfnoise"
    trainer = None
    if psf_mode:
        model.eval()
    else:
        model.train()
        trainer = MiViTTrainer(model, lr=1e-4, cuda_graph=not args.no_cuda_graph, sync_bn=args.sync_bn,
                               overlap_allreduce=not args.no_overlap, fused_allreduce=not args.nccl_allreduce)
    prm = derive_render_params(w["props"], NPOS, True)
    if not psf_mode:
        den = (BG_MEAN + PART_MEAN) - (BG_MEAN - BG_SIGMA)
        prm.normalize, prm.norm_sub, prm.norm_div = 1, float(BG_MEAN - BG_SIGMA), float(den)
    gm = np.asarray(D_GROUPS, dtype=np.float32)
    gv = np.ones(len(D_GROUPS), dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    traj = torch.empty((B, T, 2), dtype=torch.float64, device=dev)
    Dd = torch.empty((B,), dtype=torch.float32, device=dev)
    frames = torch.empty((B, NFRAMES, P, P), dtype=torch.float32, device=dev)
    labels = torch.empty((B, 1), dtype=torch.float32, device=dev)
    step_no = [0]

    def next_offset():
        off = (step_no[0] * world + rank) * B      # global sequence ids: the data set is sharding invariant
        step_no[0] += 1
        return off

    def gen_trajectories(off):
        _lib.check(L.mivit_brownian(B, T, gm.ctypes.data_as(fp), gv.ctypes.data_as(fp), len(gm), 100.0, args.seed, off,
                                    _lib.ptr(traj), _lib.ptr(Dd), _lib.current_stream()))
        torch.div(Dd.view(B, 1), D_MAX, out=labels)

    # PSFNoise: the eval-mode forward of one variant, captured once and replayed per variant (the 36 launches of a 256-sequence
    # forward are latency bound when issued one by one); --no-cuda-graph and the instrumented pass launch kernel by kernel
    psf_graph = {"g": None, "x": None, "pred": None, "n": 0, "on": not args.no_cuda_graph}

    def psf_forward(xv):
        if not psf_graph["on"]:
            return model(xv)
        if psf_graph["g"] is None:
            xs = torch.empty((B, NFRAMES, P, P), dtype=torch.float32, device=dev)
            xs.copy_(xv)
            model(xs)                               # eager once: workspace, tensor maps, function attributes
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.mivit_launch_count()
            with torch.cuda.graph(g):
                pred = model(xs)
            n_cap = int(L.mivit_launch_count() - n0)
            L.mivit_add_launch_count(-n_cap)        # captured, not executed
            psf_graph.update(g=g, x=xs, pred=pred, n=n_cap)
        psf_graph["x"].copy_(xv)
        psf_graph["g"].replay()
        L.mivit_add_launch_count(psf_graph["n"])    # the graph's kernel nodes are launches of this library too
        return psf_graph["pred"]

    def consume(off):
        """render + model step from the device-resident trajectories / labels of this step"""
        if psf_mode:
            vids = X.trajs_to_vid_psf_noise(traj, NPOS, True, w["props"], w["psf"], w["noise"], seed=args.seed, seq_offset=off)
            tot = torch.zeros((), device=dev)
            with torch.no_grad():
                for i in range(vids.shape[1]):
                    for j in range(vids.shape[2]):
                        pred = psf_forward(vids[:, i, j])
                        tot += torch.nn.functional.mse_loss(pred * D_MAX, labels * D_MAX)
            return tot
        if w.get("from_traj"):
            return trainer.train_step_from_trajectories(TrajectorySource(traj, prm, args.seed, off), labels)
        render_device(traj, prm, args.seed, seq_offset=off, out=frames, out_seq_stride=NFRAMES * P * P)
        feats = None
        if w.get("features"):      # create_video_and_feature_pairs: features of the frame-averaged (y-flipped) trajectories
            flipped = traj * torch.tensor([1.0, -1.0], dtype=torch.float64, device=dev)
            feats = torch.nan_to_num(HF.features_device(HF.average_frames_device(flipped, NPOS), 1.0), nan=0.0, posinf=0.0,
                                     neginf=0.0).float()
        return trainer.train_step(frames, labels, feats)

    def device_step():
        """generate -> render -> model step, all enqueued on the current stream."""
        off = next_offset()
        gen_trajectories(off)
        return consume(off)

    # host-buffer path (public API): pinned trajectories + labels in, loss out, every step
    h_traj = torch.empty((B, T, 2), dtype=torch.float64).pin_memory()
    h_lab = torch.empty((B, 1), dtype=torch.float32).pin_memory()
    device_step()
    torch.cuda.synchronize()
    h_traj.copy_(traj.cpu()); h_lab.copy_(labels.cpu())

    def e2e_step():
        off = next_offset()
        traj.copy_(h_traj, non_blocking=True)
        labels.copy_(h_lab, non_blocking=True)
        return float(consume(off).item())                              # D2H read of the loss = sync point

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            L.mivit_profile_enable(1)
        L.mivit_reset_launch_count()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        barrier()
        clocks = sampler.result()
        ms = e0.elapsed_time(e1)
        launches = int(L.mivit_launch_count())
        if profile:
            L.mivit_profile_enable(0)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, clocks, last

    for _ in range(max(args.warmup, 3)):
        device_step()
    # headline: the product path (CUDA-graph replay of the step unless --no-cuda-graph), no instrumentation
    ms, launches, clocks, last_loss = timed(device_step, args.steps)
    value = world * B * args.steps / (ms * 1e-3)
    # instrumented pass of the SAME step, launched kernel by kernel with CUDA events around every tagged launch on the
    # launching stream: per-kernel durations for the roofline (event records cannot live inside a replayed graph)
    if trainer is not None:
        graph_mode = trainer.cuda_graph
        trainer.cuda_graph = False
    psf_graph_on = psf_graph["on"]
    psf_graph["on"] = False
    device_step()
    ms_prof, _, _, _ = timed(device_step, args.steps, profile=True)
    psf_graph["on"] = psf_graph_on
    if trainer is not None:
        trainer.cuda_graph = graph_mode
    kt = (_lib.KernelTime * 64)()
    nk = L.mivit_profile_read(kt, 64)
    kernels = [{"name": kt[i].name.decode(), "launches": int(kt[i].launches), "ms": kt[i].total_ms, "work": kt[i].total_work}
               for i in range(nk)]

    for _ in range(2):
        e2e_step()
    ms_e2e, _, _, _ = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    step_ms = ms_prof / args.steps
    by_name = {k["name"]: k for k in kernels}

    def group(prefixes):
        ks = [k for k in kernels if any(k["name"].startswith(p) for p in prefixes)]
        return sum(k["ms"] for k in ks) / args.steps, sum(k["work"] for k in ks) / args.steps

    def hbm_roof(k, note):
        gbs = k["work"] / (k["ms"] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": k["name"], "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "traffic": None, "avg_launch_ms": k["ms"] / k["launches"],
                "share_of_step": k["ms"] / ms_prof, "peak_source": peaks["source"], "note": note}

    render_note = ("instruction-issue bound (Philox + Box-Muller + alias-table Poisson per pixel pair, 2 n P U exps per frame), "
                   "not HBM-bound: see DESIGN.md section 2")
    roof = None
    convs = [k for k in kernels if k["name"].startswith("conv_")]
    if convs and not psf_mode:
        top = max(convs, key=lambda k: k["ms"])
        ach = top["work"] / (top["ms"] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(top["name"])
        conv_ms, conv_work = group(["conv_"])
        bn_ms, bn_bytes = group(["bn_"])
        tr_ms, _ = group(["linear_tc", "attention", "layernorm", "encoder_"])
        roof = {"bound": "tensor", "kernel": top["name"], "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "peak_source": peaks["source"],
                "peak_burst": peaks.get("bf16_tflops_burst"),
                "frac_of_burst": (ach / peaks["bf16_tflops_burst"]) if peaks.get("bf16_tflops_burst") else None,
                "note": "frac > 1 means the kernel beats the cuBLAS bf16 GEMM figure measured under sustained load; "
                        "frac_of_burst is against the cuBLAS burst figure",
                "avg_launch_ms": top["ms"] / top["launches"], "share_of_step": top["ms"] / ms_prof,
                "all_convs_tflops": conv_work / (conv_ms * 1e-3) / 1e12, "all_convs_share_of_step": conv_ms / step_ms,
                # the other groups of the step, so the line cannot hide a regression outside the convolutions
                "groups_ms_per_step": {"convolutions": conv_ms, "batchnorm_streams": bn_ms, "transformer_tagged": tr_ms,
                                       "render": group(["render_"])[0], "step_instrumented": step_ms},
                "batchnorm_streams": {"ms_per_step": bn_ms, "share_of_step": bn_ms / step_ms,
                                      "gbs_algorithmic": bn_bytes / (bn_ms * 1e-3) / 1e9 if bn_ms else None,
                                      "frac_of_hbm_peak": (bn_bytes / (bn_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if bn_ms else None,
                                      "note": "HBM streams; algorithmic bytes = valid pixels x channels x 2 B x tensors touched "
                                              "(pad rows of the pitched layout excluded); DRAM bytes per kernel from ncu are in "
                                              "profiles/r02_ncu_bn.md"},
                "whole_step": {"model_tflops": value * w["flops_per_seq"] / 1e12 / world,
                               "frac_of_sustained": value * w["flops_per_seq"] / 1e12 / world / peaks["bf16_tflops"],
                               "frac_of_burst": (value * w["flops_per_seq"] / 1e12 / world / peaks["bf16_tflops_burst"])
                               if peaks.get("bf16_tflops_burst") else None},
                "timed_in": "instrumented pass (kernel-by-kernel launches with CUDA events), %.3f ms/step; the headline step "
                            "replays the same kernels as a CUDA graph" % step_ms}
    rk = by_name.get("render_v1") or by_name.get("render_embed_linear") or by_name.get("render_psfnoise")
    roof_render = hbm_roof(rk, render_note) if rk else None
    if roof is None:
        roof = roof_render          # render-dominated workloads: the renderer is the kernel the line is about
    cfg = {"workload": w["text"], "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
           "l2": ("activation working set %.1f GB per step >> 126 MB L2 (no flush needed)" %
                  (L.mivit_vit_workspace_bytes(ctypes.byref(model.vit_config(NFRAMES)), B) / 1e9)) if not psf_mode else
                 "rendered variants %.0f MB per step > 126 MB L2 (no flush needed)" % (B * 30 * NFRAMES * P * P * 4 / 1e6)}
    if psf_mode:
        cfg["launch"] = ("CUDA-graph replay of the eval forward, once per variant" if psf_graph["on"] else "kernel by kernel")
    if trainer is not None:
        cfg.update({
            "batchnorm": ("synchronised over the ranks (12 small in-graph all-reduces per step)" if trainer.sync_bn
                          else "per-rank batch statistics (stock DDP semantics)") if model._is_deep() else "none (no BatchNorm)",
            "launch": trainer.launch_description(),
            "allreduce": None if world == 1 else trainer.allreduce_description()})
    line = {"metric": "synthetic sequences/sec (render+ViT train step)" if not psf_mode else
                      "synthetic sequences/sec (PSFNoise render + ViT forward+loss on 30 variants)",
            "value": value, "unit": "sequences/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if model._is_deep() else "tf32", "data": "synthetic",
            "config": cfg, "model_tflops": value * w["flops_per_seq"] / 1e12 / world, "loss": float(last_loss.item()),
            "roofline": roof, "roofline_render": roof_render, "kernels": kernels, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sequences/s", "h2d_bytes_per_step": int(B * T * 2 * 8 + B * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches}
    if world == 1 and not args.no_cpu_baseline:
        n_seq = args.cpu_seqs if not psf_mode else max(4, args.cpu_seqs // 4)
        line["cpu_baseline"], _, _ = _cpu_baseline(args.config, n_seq, 3, 1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="framerate", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (0: the workload's default, 1024 for the headline)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seqs", type=int, default=32, help="sequences per CPU reference step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one gradient all-reduce after the whole backward")
    ap.add_argument("--sync-bn", action="store_true", help="N > 1: synchronised BatchNorm (parity mode) instead of per-rank statistics")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch the training step kernel by kernel instead of replaying it")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="N > 1: NCCL all-reduce + separate AdamW instead of the fused peer-memory all-reduce + AdamW kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return 0
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.stderr.write("bench.py: --gpus %d needs torchrun (WORLD_SIZE=%d); running replicas is not implemented\n" %
                         (args.gpus, world))
        return 2
    run_b200(args, rank, world, local_rank)
    return 0


if __name__ == "__main__":
    sys.exit(main())
