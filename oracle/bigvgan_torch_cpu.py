"""CPU baseline port of the reference's BigVGAN decode path, written with the SAME torch operators the
reference executes on a CPU host (F.conv1d / F.conv_transpose1d through oneDNN, multithreaded).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- like everything under ``oracle/`` this file is used by
``tests/`` and by ``bench.py``'s CPU arm (``--impl reference`` / ``cpu_baseline``); the product never
imports it.  It exists because the numpy oracle (``bigvgan_oracle.py``, the parity checker) is a poor
speed baseline: its FIR / snake steps are single-threaded numpy.  The reference is pure Python + torch,
so there is nothing to compile into ``oracle/_ref``; this restatement follows the reference module by
module (file:line cited per function) and is pinned to the numpy oracle and to the golden vectors the
unmodified reference produced (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import bigvgan_oracle as O


def _filter(dtype) -> torch.Tensor:
    """alias_free_torch/filter.py:29-58 (12 taps, cutoff 0.25, half width 0.3) as [1,1,12]."""
    return torch.from_numpy(O.kaiser_sinc_filter1d(0.25, 0.3, 12)).to(dtype).view(1, 1, 12)


def upsample1d(x: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """UpSample1d.forward -- alias_free_torch/resample.py:25-33 (ratio 2, kernel 12)."""
    C = x.shape[1]
    x = F.pad(x, (5, 5), mode="replicate")
    x = 2 * F.conv_transpose1d(x, f.expand(C, -1, -1), stride=2, groups=C)
    return x[..., 15:-15]


def downsample1d(x: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """DownSample1d / LowPassFilter1d.forward -- resample.py:46-49, filter.py:84-96 (stride 2)."""
    C = x.shape[1]
    x = F.pad(x, (5, 6), mode="replicate")
    return F.conv1d(x, f.expand(C, -1, -1), stride=2, groups=C)


def snakebeta(x, log_alpha, log_beta):
    """SnakeBeta.forward, alpha_logscale=True -- activations.py:109-122."""
    alpha = torch.exp(log_alpha)[None, :, None]
    beta = torch.exp(log_beta)[None, :, None]
    return x + (1.0 / (beta + 0.000000001)) * torch.pow(torch.sin(x * alpha), 2)


def activation1d(x, log_alpha, log_beta, f):
    """Activation1d.forward -- alias_free_torch/act.py:24-29."""
    return downsample1d(snakebeta(upsample1d(x, f), log_alpha, log_beta), f)


def amp_block1(x, sd, prefix, k, dilations, f):
    """AMPBlock1.forward -- models.py:65-74."""
    for i, d in enumerate(dilations):
        a1, a2 = f"{prefix}.activations.{2 * i}.act", f"{prefix}.activations.{2 * i + 1}.act"
        xt = activation1d(x, sd[a1 + ".alpha"], sd[a1 + ".beta"], f)
        xt = F.conv1d(xt, sd[f"{prefix}.convs1.{i}.weight"], sd[f"{prefix}.convs1.{i}.bias"], dilation=d,
                      padding=O.get_padding(k, d))
        xt = activation1d(xt, sd[a2 + ".alpha"], sd[a2 + ".beta"], f)
        xt = F.conv1d(xt, sd[f"{prefix}.convs2.{i}.weight"], sd[f"{prefix}.convs2.{i}.bias"], dilation=1,
                      padding=O.get_padding(k, 1))
        x = xt + x
    return x


def prepare_state_dict(sd: Dict[str, np.ndarray], dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in sd.items()
            if not k.startswith("speaker_encoder.") and not k.endswith("filter")}


@torch.inference_mode()
def bigvgan_forward_with_embedding(x, spk, sd, h=None, dtype=torch.float32) -> np.ndarray:
    """BigVGAN.forward after the speaker encoder -- models.py:210-250.  ``sd`` from prepare_state_dict
    (folded weights); x[B,T,gpt_dim], spk[B',1,512] numpy or torch.  Returns numpy [B,1,T*hop]."""
    h = dict(O.DEFAULT_H, **(h or {}))
    f = _filter(dtype)
    x = torch.as_tensor(np.asarray(x)).to(dtype).transpose(1, 2)                 # :220
    spk = torch.as_tensor(np.asarray(spk)).to(dtype).transpose(1, 2)             # :210
    x = F.conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)       # :224
    x = x + F.conv1d(spk, sd["cond_layer.weight"], sd["cond_layer.bias"])        # :226
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], stride=u, padding=(k - u) // 2)  # :230
        if h["cond_d_vector_in_each_upsampling_layer"]:
            x = x + F.conv1d(spk, sd[f"conds.{i}.weight"], sd[f"conds.{i}.bias"])                                 # :233-234
        xs = None
        for j in range(nk):                                                                                     # :237-242
            r = amp_block1(x, sd, f"resblocks.{i * nk + j}", h["resblock_kernel_sizes"][j],
                           h["resblock_dilation_sizes"][j], f)
            xs = r if xs is None else xs + r
        x = xs / nk                                                                                              # :243
    x = activation1d(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"], f)                     # :246
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)                                    # :247
    return torch.tanh(x).numpy()                                                                                # :248
