"""ctypes binding of libb200vgan.so (the C ABI declared in include/b200vgan.h).

There is no CPU fallback: `load()` raises if the shared library is missing or a compute call is
made without an sm_100 GPU (the library itself reports that through bvg_last_error)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import build as _build

MODE_FP32, MODE_BF16, MODE_F16, MODE_FP32_TC = 0, 1, 2, 3
F32, BF16, F16 = 0, 1, 2
MAX_UPS, MAX_KERNELS, MAX_DILATIONS = 8, 4, 4

# every symbol include/b200vgan.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "bvg_last_error", "bvg_version", "bvg_device_check", "bvg_create", "bvg_destroy", "bvg_set_weight",
    "bvg_finalize", "bvg_plan_create", "bvg_plan_destroy", "bvg_plan_workspace_bytes", "bvg_plan_max_frames",
    "bvg_plan_num_launches", "bvg_forward", "bvg_forward_host", "bvg_activation1d", "bvg_conv1d",
    "bvg_conv_transpose1d", "bvg_forward_ragged", "bvg_plans_created", "bvg_plan_total_frames", "bvg_profile_enable", "bvg_profile_read",
    "bvg_activation1d_packed", "bvg_act_conv1d", "bvg_ecapa_workspace_bytes", "bvg_speaker_embedding",
    "bvg_forward_pcm16", "bvg_mel_frames", "bvg_log_mel", "bvg_timeline_merge",
]


class BvgConfig(C.Structure):
    _fields_ = [
        ("gpt_dim", C.c_int32),
        ("upsample_initial_channel", C.c_int32),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * MAX_UPS),
        ("upsample_kernel_sizes", C.c_int32 * MAX_UPS),
        ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * MAX_KERNELS),
        ("resblock_dilation_sizes", (C.c_int32 * MAX_DILATIONS) * MAX_KERNELS),
        ("num_dilations", C.c_int32),
        ("speaker_embedding_dim", C.c_int32),
        ("cond_in_each_up_layer", C.c_int32),
        ("num_mels", C.c_int32),
    ]


class BvgError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load(rebuild: bool = False) -> C.CDLL:
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    path = _build.LIB_PATH
    if rebuild or os.environ.get("B200VGAN_REBUILD") == "1" or not os.path.exists(path):
        path = _build.build(force=rebuild)
    lib = C.CDLL(path)
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    lib.bvg_last_error.restype = C.c_char_p
    lib.bvg_last_error.argtypes = []
    lib.bvg_version.restype = C.c_int
    lib.bvg_device_check.restype = C.c_int
    lib.bvg_create.argtypes = [C.POINTER(BvgConfig), C.POINTER(vp)]
    lib.bvg_destroy.argtypes = [vp]
    lib.bvg_destroy.restype = None
    lib.bvg_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64), i32, i32, vp]
    lib.bvg_finalize.argtypes = [vp, vp]
    lib.bvg_plan_create.argtypes = [vp, i32, C.POINTER(i32), i32, C.POINTER(vp)]
    lib.bvg_plan_destroy.argtypes = [vp]
    lib.bvg_plan_destroy.restype = None
    lib.bvg_plan_workspace_bytes.argtypes = [vp]
    lib.bvg_plan_workspace_bytes.restype = sz
    lib.bvg_plan_max_frames.argtypes = [vp]
    lib.bvg_plan_num_launches.argtypes = [vp]
    lib.bvg_forward_ragged.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, vp, sz, vp]
    lib.bvg_plans_created.argtypes = [vp]
    lib.bvg_plans_created.restype = C.c_int64
    lib.bvg_plan_total_frames.argtypes = [vp]
    lib.bvg_plan_total_frames.restype = C.c_int64
    lib.bvg_profile_enable.argtypes = [vp, i32]
    lib.bvg_profile_read.argtypes = [vp, vp, vp, vp, vp]
    lib.bvg_forward.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, sz, vp]
    lib.bvg_forward_pcm16.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, vp, sz, vp]
    lib.bvg_forward_host.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, sz, vp]
    lib.bvg_activation1d.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.bvg_activation1d_packed.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.bvg_conv1d.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.bvg_act_conv1d.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.bvg_ecapa_workspace_bytes.argtypes = [vp, i32, i32]
    lib.bvg_ecapa_workspace_bytes.restype = sz
    lib.bvg_speaker_embedding.argtypes = [vp, vp, i32, i32, vp, vp, vp, sz, vp]
    lib.bvg_conv_transpose1d.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.bvg_timeline_merge.argtypes = [vp, vp, vp, vp, vp, i32, C.c_int64, vp, C.c_int64, i32, C.c_float, vp, vp]
    lib.bvg_mel_frames.argtypes = [i32, i32]
    lib.bvg_log_mel.argtypes = [vp, i32, i32, i32, i32, i32, C.c_float, C.c_float, vp, i32, vp]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise BvgError(load().bvg_last_error().decode(errors="replace"))


def make_config(h) -> BvgConfig:
    """`h`: the `bigvgan` config node of the reference (attribute or item access, config.yaml:51-70)."""
    def get(k, default=None):
        if isinstance(h, dict):
            return h.get(k, default)
        return getattr(h, k, default) if hasattr(h, k) else h.get(k, default)

    cfg = BvgConfig()
    rates, ksz = list(get("upsample_rates")), list(get("upsample_kernel_sizes"))
    rks, rds = list(get("resblock_kernel_sizes")), [list(d) for d in get("resblock_dilation_sizes")]
    if str(get("resblock", "1")) != "1":
        raise BvgError("only resblock type '1' (AMPBlock1) is supported")
    if get("activation", "snakebeta") != "snakebeta" or not get("snake_logscale", True):
        raise BvgError("only activation=snakebeta with snake_logscale=true is supported")
    if get("feat_upsample", False):
        raise BvgError("feat_upsample=true is not supported")
    if len(rates) > MAX_UPS or len(rks) > MAX_KERNELS or any(len(d) != len(rds[0]) or len(d) > MAX_DILATIONS for d in rds):
        raise BvgError("architecture exceeds compiled limits")
    cfg.gpt_dim = int(get("gpt_dim"))
    cfg.upsample_initial_channel = int(get("upsample_initial_channel"))
    cfg.num_upsamples = len(rates)
    for i, (u, k) in enumerate(zip(rates, ksz)):
        cfg.upsample_rates[i] = int(u)
        cfg.upsample_kernel_sizes[i] = int(k)
    cfg.num_kernels = len(rks)
    cfg.num_dilations = len(rds[0])
    for j, k in enumerate(rks):
        cfg.resblock_kernel_sizes[j] = int(k)
        for m, d in enumerate(rds[j]):
            cfg.resblock_dilation_sizes[j][m] = int(d)
    cfg.speaker_embedding_dim = int(get("speaker_embedding_dim"))
    cfg.cond_in_each_up_layer = 1 if get("cond_d_vector_in_each_upsampling_layer", True) else 0
    cfg.num_mels = int(get("num_mels", 100))
    return cfg
