#!/usr/bin/env python
"""bench.py -- headline benchmark of the MiViT hot path on B200.

Metric (BASELINE.json): synthetic sequences/sec of  render + ViT train step.
Workload (BASELINE.json configs[1], "Framerate experiment: 30-frame sequences with full ViT
training step on 1 B200"): per step and per GPU, B Brownian trajectories (T=300 sub-steps, D groups
of trainModelsFramerate.py:45) are generated on the device, rendered to 30 frames of 13x13 pixels
with the Framerate experiment's image_props (n=10 sub-positions per frame, background + Poisson
noise, fused normalisation), and pushed through one full training step (forward, MSE, backward,
AdamW lr=1e-4) of the DeepResNet-embedding ViT (E64/H4/HD128/L6, regression token, no pos-enc,
506 081 parameters).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference algorithm on the host CPU cores (oracle port)

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same metric
through the public host-buffer API (pinned H2D of the trajectories, D2H of the loss every step).
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

# ---- workload: Experiments/Framerate/trainSettingsFramerate.py (image_props :62-81, model :40-46)
P, NPOS, NFRAMES = 13, 10, 30
T = NPOS * NFRAMES
BG_MEAN, BG_SIGMA = 1420, 290
PART_MEAN, PART_STD = 6000 - BG_MEAN, 500
IMAGE_PROPS = {"particle_intensity": [PART_MEAN, PART_STD], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3,
               "resolution": 100e-9, "output_size": P, "upsampling_factor": 5, "background_intensity": [BG_MEAN, BG_SIGMA],
               "poisson_noise": 100, "trajectory_unit": 1200}
D_GROUPS = [1, 3, 5, 7, 9, 10.2]          # TrainingDs_list means (variance 1), trainModelsFramerate.py:45
D_MAX = 10.0
EMBED, HEADS, HIDDEN, LAYERS = 64, 4, 128, 6
# algorithmic work per sequence
FLOPS_PER_SEQ_TRAIN = 8.773e9             # P13 F30 deepcnn_n: 2 924.4 MF forward x 3
RENDER_BYTES_PER_SEQ = T * 2 * 8 + NFRAMES * P * P * 4


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
    traj, seed = args
    from oracle import render_oracle as ro
    from oracle.noise import NumpyNoise
    vid = ro.render_v1(traj, NPOS, True, IMAGE_PROPS, noise=NumpyNoise(seed), mode="literal")
    vid, _ = ro.normalize_images(vid, BG_MEAN, BG_SIGMA, BG_MEAN + PART_MEAN)
    return vid


def cpu_reference_step(n_seq, state, pool=None):
    """One bounded sample of the reference algorithm on the host: the literal numpy renderer
    (oracle/render_oracle.py) + normalisation + one fp32 PyTorch training step (oracle/vit_oracle.py, torch
    intra-op threads = all cores).  pool = None: the renderer runs in ONE process, as the reference does (its
    use_multiprocessing branch raises TypeError); with a process pool the sequences of the step are rendered on all
    cores.  Returns (t_render, t_train)."""
    import numpy as np
    import torch
    from oracle import vit_oracle as vo
    from oracle.trajectory_oracle import brownian_oracle
    traj, D = brownian_oracle(n_seq, T, D_GROUPS, [1.0] * len(D_GROUPS), 100.0, seed=state["step"], seq_offset=0)
    t0 = time.perf_counter()
    if pool is None:
        vid = _render_slice((traj, state["step"]))
    else:
        parts = np.array_split(np.arange(n_seq), min(n_seq, pool._processes))
        vid = np.concatenate(pool.map(_render_slice, [(traj[p], state["step"] * 1000 + i) for i, p in enumerate(parts)]))
    t1 = time.perf_counter()
    x = torch.from_numpy(np.ascontiguousarray(vid))
    y = torch.from_numpy((D / D_MAX).astype(np.float32)).unsqueeze(-1)
    vo.train_step(state["sd"], state["opt"], state["cfg"], x, y)
    t2 = time.perf_counter()
    state["step"] += 1
    return t1 - t0, t2 - t1


def cpu_reference_state():
    import torch
    from oracle import vit_oracle as vo
    import torch.nn as nn
    torch.manual_seed(0)
    cfg = dict(embedding="deepresnet", embed_dim=EMBED, num_heads=HEADS, num_layers=LAYERS, activation="relu",
               use_pos_encoding=False, use_regression_token=True)
    sd = _random_state_dict()
    return {"sd": sd, "opt": vo.new_opt_state(sd), "cfg": cfg, "step": 0}


def _random_state_dict():
    """Random-init weights of the reference architecture via plain torch modules (host side, CPU)."""
    import torch
    import torch.nn as nn
    torch.manual_seed(0)
    sd = {}

    def lin(pre, o, i):
        m = nn.Linear(i, o)
        sd[pre + ".weight"], sd[pre + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()

    def bn(pre, c):
        sd[pre + ".weight"], sd[pre + ".bias"] = torch.ones(c), torch.zeros(c)
        sd[pre + ".running_mean"], sd[pre + ".running_var"] = torch.zeros(c), torch.ones(c)
        sd[pre + ".num_batches_tracked"] = torch.tensor(0)

    def conv(key, o, i, k):
        sd[key] = nn.Conv2d(i, o, k, bias=False).weight.detach().clone()

    conv("embedding.initial_conv.weight", 32, 1, 3); bn("embedding.bn1", 32)
    for b, (ci, co) in (("res_block1", (32, 64)), ("res_block2", (64, 128))):
        conv("embedding.%s.conv1.weight" % b, co, ci, 3); bn("embedding.%s.bn1" % b, co)
        conv("embedding.%s.conv2.weight" % b, co, co, 3); bn("embedding.%s.bn2" % b, co)
        conv("embedding.%s.skip.0.weight" % b, co, ci, 1); bn("embedding.%s.skip.1" % b, co)
    lin("embedding.fc", EMBED, 128)
    sd["norm.weight"], sd["norm.bias"] = torch.ones(EMBED), torch.zeros(EMBED)
    sd["reg_token"] = torch.randn(1, 1, EMBED)
    for l in range(LAYERS):
        p = "transformer.encoder_layers.%d." % l
        for s in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(p + "self_attn." + s, EMBED, EMBED)
        for nrm in ("norm1", "norm2"):
            sd[p + nrm + ".weight"], sd[p + nrm + ".bias"] = torch.ones(EMBED), torch.zeros(EMBED)
        lin(p + "feed_forward.fc1", HIDDEN, EMBED); lin(p + "feed_forward.fc2", EMBED, HIDDEN)
    sd["transformer.norm.weight"], sd["transformer.norm.bias"] = torch.ones(EMBED), torch.zeros(EMBED)
    lin("mlp_head.mlp.0", 128, EMBED); lin("mlp_head.mlp.3", 1, 128)
    return sd


def run_reference(args, rank):
    """Reference arm: the reference algorithm on ALL host cores -- the renderer sharded over a process pool (one
    process per core, each running the literal per-sequence loop), the training step with torch using every core.
    The as-written figure (renderer in one process, like the reference's own loop) is reported beside it."""
    if rank != 0:
        return
    import multiprocessing as mp
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_seq = args.cpu_seqs
    st = cpu_reference_state()
    a, b = cpu_reference_step(n_seq, st)                       # as written: single-process renderer
    as_written = n_seq / (a + b)
    pool = mp.get_context("fork").Pool(cores)
    try:
        for _ in range(max(args.warmup, 1) if args.steps < 10 else 2):
            cpu_reference_step(n_seq, st, pool)
        t0 = time.perf_counter()
        tr = tt = 0.0
        for _ in range(args.steps):
            a, b = cpu_reference_step(n_seq, st, pool)
            tr += a; tt += b
        el = time.perf_counter() - t0
    finally:
        pool.close()
    val = n_seq * args.steps / (tr + tt)
    line = {"impl": "reference", "metric": "synthetic sequences/sec (render+ViT train step)", "value": val, "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tr + tt) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "framerate_P13_F30_deepcnn_n (BASELINE configs[1])", "batch_per_step": n_seq},
            "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port",
                             "as_written_value": as_written,
                             "sample": "%d steps x %d sequences: literal numpy renderer sharded over %d processes + fp32 torch "
                                       "train step (%d intra-op threads); render %.1f ms/seq, train %.1f ms/seq; as written "
                                       "(renderer in one process, like the reference's loop): %.1f sequences/s" %
                                       (args.steps, n_seq, cores, cores, 1e3 * tr / (n_seq * args.steps),
                                        1e3 * tt / (n_seq * args.steps), as_written)},
            "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": el}
    print(json.dumps(line))


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import ctypes
    import numpy as np
    import torch
    import torch.nn.functional as F
    import torch.distributed as dist
    from moleculardiffusion_mivit_b200 import _lib, models as M
    from moleculardiffusion_mivit_b200.helpersGeneration import brownian_motion, derive_render_params, render_device
    from moleculardiffusion_mivit_b200.training import MiViTTrainer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL announces its version on STDOUT at the default debug level; rank 0's stdout must be the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B = args.batch
    torch.manual_seed(0)
    model = M.GeneralTransformer(M.DeepResNetEmbedding, {"patch_size": P, "embed_dim": EMBED}, EMBED, HEADS, HIDDEN, LAYERS,
                                 M.MLPHead, F.relu, 0.0, False, True, True).cuda().train()
    trainer = MiViTTrainer(model, lr=1e-4, cuda_graph=not args.no_cuda_graph, sync_bn=args.sync_bn,
                           overlap_allreduce=not args.no_overlap)
    prm = derive_render_params(IMAGE_PROPS, NPOS, True)
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

    def device_step():
        """generate -> render (+normalise) -> train step, all enqueued on the current stream."""
        off = (step_no[0] * world + rank) * B      # global sequence ids: the data set is sharding invariant
        step_no[0] += 1
        st = _lib.current_stream()
        _lib.check(L.mivit_brownian(B, T, gm.ctypes.data_as(fp), gv.ctypes.data_as(fp), len(gm), 100.0, args.seed, off,
                                    _lib.ptr(traj), _lib.ptr(Dd), st))
        render_device(traj, prm, args.seed, seq_offset=off, out=frames, out_seq_stride=NFRAMES * P * P)
        torch.div(Dd.view(B, 1), D_MAX, out=labels)
        return trainer.train_step(frames, labels)

    # host-buffer path (public API): pinned trajectories + labels in, loss out, every step
    h_traj = torch.empty((B, T, 2), dtype=torch.float64).pin_memory()
    h_lab = torch.empty((B, 1), dtype=torch.float32).pin_memory()
    device_step()
    torch.cuda.synchronize()
    h_traj.copy_(traj.cpu()); h_lab.copy_(labels.cpu())

    def e2e_step():
        off = (step_no[0] * world + rank) * B
        step_no[0] += 1
        traj.copy_(h_traj, non_blocking=True)
        labels.copy_(h_lab, non_blocking=True)
        render_device(traj, prm, args.seed, seq_offset=off, out=frames, out_seq_stride=NFRAMES * P * P)
        return float(trainer.train_step(frames, labels).item())      # D2H read of the loss = sync point

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
    graph_mode = trainer.cuda_graph
    trainer.cuda_graph = False
    device_step()
    ms_prof, _, _, _ = timed(device_step, args.steps, profile=True)
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
    convs = [k for k in kernels if k["name"].startswith("conv_")]
    roof = None
    if convs:
        top = max(convs, key=lambda k: k["ms"])
        ach = top["work"] / (top["ms"] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(top["name"])
        roof = {"bound": "tensor", "kernel": top["name"], "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "peak_source": peaks["source"],
                "peak_burst": peaks.get("bf16_tflops_burst"),
                "frac_of_burst": (ach / peaks["bf16_tflops_burst"]) if peaks.get("bf16_tflops_burst") else None,
                "note": "frac > 1 means the kernel beats the cuBLAS bf16 GEMM figure measured under sustained load; "
                        "frac_of_burst is against the cuBLAS burst figure",
                "avg_launch_ms": top["ms"] / top["launches"], "share_of_step": top["ms"] / ms_prof,
                "all_convs_tflops": sum(k["work"] for k in convs) / (sum(k["ms"] for k in convs) * 1e-3) / 1e12,
                "all_convs_share_of_step": sum(k["ms"] for k in convs) / ms_prof,
                "timed_in": "instrumented pass (kernel-by-kernel launches with CUDA events), %.3f ms/step; the headline step "
                            "replays the same kernels as a CUDA graph" % (ms_prof / args.steps)}
    rnd = [k for k in kernels if k["name"] == "render_v1"]
    roof_render = None
    if rnd:
        gbs = rnd[0]["work"] / (rnd[0]["ms"] * 1e-3) / 1e9
        roof_render = {"bound": "hbm", "kernel": "render_v1", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": gbs / peaks["hbm_gbs"], "traffic": None, "avg_launch_ms": rnd[0]["ms"] / rnd[0]["launches"],
                       "note": "ALU/RNG-bound (Philox + Poisson per pixel), not HBM-bound: see DESIGN.md"}
    line = {"metric": "synthetic sequences/sec (render+ViT train step)", "value": value, "unit": "sequences/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "framerate_P13_F30_deepcnn_n (BASELINE configs[1]): on-device Brownian T=300 -> render 30x13x13 "
                                   "(n=10, bg+Poisson noise, normalised) -> DeepResNet-ViT E64/H4/HD128/L6 fwd+MSE+bwd+AdamW",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "activation working set %.1f GB per step >> 126 MB L2 (no flush needed)" %
                             (L.mivit_vit_workspace_bytes(ctypes.byref(model.vit_config(NFRAMES)), B) / 1e9),
                       "batchnorm": "synchronised over the ranks (12 small all-reduces per step)" if trainer.sync_bn
                                    else "per-rank batch statistics (stock DDP semantics)",
                       "launch": "CUDA-graph replay of forward+loss+backward, eager AdamW"
                                 if (trainer.cuda_graph and not trainer.sync_bn) else "kernel by kernel",
                       "allreduce": None if world == 1 else
                                    ("two buckets, the non-embedding one overlapped with the image-embedding backward"
                                     if (trainer.overlap_allreduce and not trainer.sync_bn) else "one all-reduce after the backward")},
            "model_tflops": value * FLOPS_PER_SEQ_TRAIN / 1e12 / world, "loss": float(last_loss.item()),
            "roofline": roof, "roofline_render": roof_render, "kernels": kernels, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sequences/s", "h2d_bytes_per_step": int(B * T * 2 * 8 + B * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches}
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        st = cpu_reference_state()
        a, b = cpu_reference_step(args.cpu_seqs, st)
        as_written = args.cpu_seqs / (a + b)
        pool = mp.get_context("fork").Pool(cores)
        try:
            cpu_reference_step(args.cpu_seqs, st, pool)
            tr = tt = 0.0
            reps = 3
            for _ in range(reps):
                a, b = cpu_reference_step(args.cpu_seqs, st, pool)
                tr += a; tt += b
        finally:
            pool.close()
        cval = args.cpu_seqs * reps / (tr + tt)
        line["cpu_baseline"] = {"value": cval, "unit": "sequences/s", "cores": cores, "kind": "port", "as_written_value": as_written,
                                "sample": "%d steps x %d sequences of the same workload: literal numpy renderer sharded over %d "
                                          "processes %.1f ms/seq + fp32 torch train step (%d threads) %.1f ms/seq; as written "
                                          "(renderer in one process): %.1f sequences/s" %
                                          (reps, args.cpu_seqs, cores, 1e3 * tr / (args.cpu_seqs * reps), cores,
                                           1e3 * tt / (args.cpu_seqs * reps), as_written)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="sequences per GPU per step")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seqs", type=int, default=32, help="sequences per CPU reference step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one gradient all-reduce after the whole backward")
    ap.add_argument("--sync-bn", action="store_true", help="N > 1: synchronised BatchNorm (parity mode) instead of per-rank statistics")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch the training step kernel by kernel instead of replaying it")
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
