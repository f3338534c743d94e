"""ctypes binding of the C-ABI library (include/mivit.h).  There is no CPU fallback:
importing this module without a built libmivit_b200.so raises."""
import ctypes
import os
import shutil

from . import build as _build

_LIB = None


class RenderParams(ctypes.Structure):
    """struct mivit_render_params (include/mivit.h)."""
    _fields_ = [
        ("scale", ctypes.c_double), ("sigma_hr", ctypes.c_double),
        ("P", ctypes.c_int32), ("U", ctypes.c_int32), ("n", ctypes.c_int32), ("center", ctypes.c_int32),
        ("flip_y", ctypes.c_int32), ("draw", ctypes.c_int32),
        ("part_mean", ctypes.c_float), ("part_std", ctypes.c_float), ("bg_mean", ctypes.c_float),
        ("bg_std", ctypes.c_float), ("poisson", ctypes.c_float),
        ("normalize", ctypes.c_int32), ("norm_sub", ctypes.c_float), ("norm_div", ctypes.c_float),
        ("mean_noise", ctypes.c_int32),
    ]


class KernelTime(ctypes.Structure):
    """struct mivit_kernel_time (include/mivit.h)."""
    _fields_ = [("name", ctypes.c_char * 48), ("launches", ctypes.c_int64), ("total_ms", ctypes.c_double),
                ("total_work", ctypes.c_double)]


# typedef int (*mivit_allreduce_fn)(void* device_buf, int64_t n_floats, void* stream, void* user)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p)


class ResnetConfig(ctypes.Structure):
    """struct mivit_resnet_config (include/mivit.h)."""
    _fields_ = [("P", ctypes.c_int32), ("F", ctypes.c_int32), ("feature_size", ctypes.c_int32), ("ext_dim", ctypes.c_int32),
                ("hidden", ctypes.c_int32), ("single_prediction", ctypes.c_int32), ("activation", ctypes.c_int32),
                ("bn_eps", ctypes.c_float), ("bn_momentum", ctypes.c_float)]


class PeerComm(ctypes.Structure):
    """struct mivit_peer_comm (include/mivit.h)."""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("segment", ctypes.c_void_p * 8)]


class MivitError(RuntimeError):
    pass


def _declare(lib):
    c = ctypes
    vp, i32, i64, u64, f32, f64 = c.c_void_p, c.c_int32, c.c_int64, c.c_uint64, c.c_float, c.c_double
    fp = c.POINTER(c.c_float)
    sigs = {
        "mivit_abi_version": (i32, []),
        "mivit_last_error": (c.c_char_p, []),
        "mivit_launch_count": (i64, []),
        "mivit_reset_launch_count": (None, []),
        "mivit_add_launch_count": (None, [i64]),
        "mivit_set_allreduce_hook": (None, [ALLREDUCE_FN, vp, i32]),
        "mivit_profile_enable": (None, [i32]),
        "mivit_profile_read": (i32, [c.POINTER(KernelTime), i32]),
        "mivit_render_v1": (i32, [vp, i64, i32, c.POINTER(RenderParams), u64, u64, vp, i64, vp]),
        "mivit_render_multi": (i32, [vp, i64, i32, c.POINTER(RenderParams), u64, u64, vp, vp, vp, vp, vp]),
        "mivit_rl_tv": (i32, [vp, i64, i32, vp, i32, c.POINTER(c.c_int32), i32, f32, vp, vp]),
        "mivit_render_embed_linear": (i32, [vp, i64, i32, c.POINTER(RenderParams), u64, u64, vp, vp, i32, vp, vp, i64, vp]),
        "mivit_render_embed_linear_wgrad": (i32, [vp, i64, i32, c.POINTER(RenderParams), u64, u64, vp, i32, vp, vp, vp]),
        "mivit_poisson_alias_table": (i32, [f64, c.POINTER(c.c_uint32), c.POINTER(i32)]),
        "mivit_render_psfnoise": (i32, [vp, i64, i32, c.POINTER(RenderParams), fp, i32, fp, i32, f32, u64, u64, vp, vp]),
        "mivit_average_frames": (i32, [vp, i64, i32, i32, vp, vp]),
        "mivit_diffusion_features": (i32, [vp, i64, i32, f64, vp, vp]),
        "mivit_brownian": (i32, [i64, i32, fp, fp, i32, f64, u64, u64, vp, vp, vp]),
        "mivit_conv_pack_weights": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "mivit_conv_rows": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, vp]),
        "mivit_conv_rows_fused": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
        "mivit_conv_rows_wgrad": (i32, [vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
        "mivit_linear_tf32": (i32, [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "mivit_comm_segment_bytes": (i64, [i64]),
        "mivit_comm_grad_offset_bytes": (i64, []),
        "mivit_comm_alloc": (i32, [i64, c.POINTER(vp)]),
        "mivit_comm_free": (i32, [vp]),
        "mivit_comm_ipc_handle": (i32, [vp, c.POINTER(c.c_uint8)]),
        "mivit_comm_ipc_open": (i32, [c.POINTER(c.c_uint8), c.POINTER(vp)]),
        "mivit_comm_ipc_close": (i32, [vp]),
        "mivit_comm_set_lr": (i32, [c.POINTER(PeerComm), f32, vp]),
        "mivit_comm_set_step": (i32, [c.POINTER(PeerComm), i64, vp]),
        "mivit_allreduce_adamw": (i32, [c.POINTER(PeerComm), i32, i64, i64, vp, vp, vp, f32, f32, f32, f32, i32, vp, i32, vp]),
        "mivit_allreduce_small": (i32, [c.POINTER(PeerComm), i32, vp, i32, vp]),
        "mivit_set_bn_sync_comm": (i32, [c.POINTER(PeerComm)]),
        "mivit_vit_param_count": (i32, [vp]),
        "mivit_vit_param_sizes": (i32, [vp, c.POINTER(c.c_int64), i32]),
        "mivit_vit_workspace_bytes": (i64, [vp, i32]),
        "mivit_vit_forward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        "mivit_vit_backward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "mivit_vit_backward_part": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, i32, vp]),
        "mivit_vit_embedding_param_count": (i64, [vp]),
        "mivit_vit_pred_rows": (i32, [vp, i32]),
        "mivit_vit_forward_traj": (i32, [vp, i32, vp, i32, c.POINTER(RenderParams), u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        "mivit_vit_backward_traj": (i32, [vp, i32, vp, i32, c.POINTER(RenderParams), u64, u64, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        "mivit_vit_train_step_traj": (i32, [vp, i32, vp, i32, c.POINTER(RenderParams), u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                            vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, i64, i32, vp]),
        "mivit_resnet_param_count": (i32, [vp]),
        "mivit_resnet_param_sizes": (i32, [vp, c.POINTER(c.c_int64), i32]),
        "mivit_resnet_workspace_bytes": (i64, [vp, i32]),
        "mivit_resnet_pred_rows": (i32, [vp, i32]),
        "mivit_resnet_forward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        "mivit_resnet_backward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "mivit_resnet_train_step": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                          f32, f32, f32, f32, f32, i64, i32, vp]),
        "mivit_mse_loss": (i32, [vp, vp, i32, vp, vp, vp]),
        "mivit_adamw_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, f32, vp]),
        "mivit_vit_train_step": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                       f32, f32, f32, f32, f32, i64, i32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sigs


def lib():
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if shutil.which(_build.NVCC) or os.path.exists(_build.NVCC):
            if _build.needs_build():
                _build.build()
        if not os.path.exists(path):
            raise ImportError(
                "moleculardiffusion_mivit_b200: %s is missing and nvcc is not available to build it. "
                "Run `python -m moleculardiffusion_mivit_b200.build` (there is no CPU fallback)." % path)
        _LIB = ctypes.CDLL(path)
        _declare(_LIB)
    return _LIB


def check(rc):
    if rc != 0:
        raise MivitError(lib().mivit_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise MivitError("moleculardiffusion_mivit_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())
