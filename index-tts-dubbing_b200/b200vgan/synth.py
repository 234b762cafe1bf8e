"""Deterministic synthetic weights, latents and speaker inputs for the BigVGAN2 decode path.

There is no network (no checkpoints), so every parity test, the golden-vector generator
(`oracle/gen_golden.py`, which loads the same tensors into the *unmodified reference*) and
`bench.py` use random-init weights of the `checkpoints/config.yaml` architecture produced here.
Each tensor has its own counter-based RNG stream keyed by (seed, crc32(name)), so the values do
not depend on generation order and are identical in the build container and on the GPU box.

Key layout = the reference's state dict: the folded layout
(`*.weight`, 913 keys) by default, or the checkpoint layout (`*.weight_g`/`*.weight_v`,
1029 keys) with ``weight_norm=True``.  `act.alpha/beta` are N(0, 0.5^2) in log scale
(the reference's zero init would hide per-channel indexing bugs).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Optional

import numpy as np

H_DEFAULT = dict(
    resblock="1",
    upsample_rates=[4, 4, 4, 4, 2, 2],
    upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    feat_upsample=False,
    speaker_embedding_dim=512,
    cond_d_vector_in_each_upsampling_layer=True,
    gpt_dim=1024,
    activation="snakebeta",
    snake_logscale=True,
    num_mels=100,
    sampling_rate=24000,
)

KAISER_TAPS = np.array([
    0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971,
    0.4432097971, 0.1285725832, -0.0576573834, -0.0255434588, 0.0093894657, 0.0020289647,
], dtype=np.float32)


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([int(seed), zlib.crc32(name.encode())])


def _normal(seed, name, shape, std=1.0, mean=0.0):
    return (mean + std * _rng(seed, name).standard_normal(shape, dtype=np.float32)).astype(np.float32)


def kaiser_filter() -> np.ndarray:
    """12-tap kaiser-sinc low-pass, identical for up and down sampling
    (reference: alias_free_torch/filter.py:29-58 with cutoff 0.25, half_width 0.3)."""
    half = 6
    A = 2.285 * (half - 1) * math.pi * (4 * 0.3) + 7.95
    beta = 0.1102 * (A - 8.7)
    t = np.arange(-half, half) + 0.5
    f = 2 * 0.25 * np.kaiser(12, beta) * np.sinc(2 * 0.25 * t)
    return (f / f.sum()).astype(np.float32)


def make_state_dict(seed: int = 1234, h: Optional[dict] = None, weight_norm: bool = False,
                    with_speaker_encoder: bool = True) -> Dict[str, np.ndarray]:
    h = dict(H_DEFAULT, **(h or {}))
    sd: Dict[str, np.ndarray] = {}
    filt = kaiser_filter().reshape(1, 1, 12)

    def conv(name, cout, cin, k, transposed=False, gain=1.0, taps_per_out=None):
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        fan = cin * (taps_per_out if taps_per_out else k)
        w = _normal(seed, name + ".weight", shape, std=gain / math.sqrt(fan))
        if weight_norm:
            # w = g * v / ||v||: pick v = w * r (r > 0 per dim-0 slice), g = ||w||
            r = np.exp(_normal(seed, name + ".wn_scale", (shape[0], 1, 1), std=0.3))
            nrm = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            sd[name + ".weight_g"] = nrm
            sd[name + ".weight_v"] = (w * r).astype(np.float32)
        else:
            sd[name + ".weight"] = w
        sd[name + ".bias"] = _normal(seed, name + ".bias", (cout,), std=0.05)

    def act(name, ch):
        sd[name + ".act.alpha"] = _normal(seed, name + ".act.alpha", (ch,), std=0.5)
        sd[name + ".act.beta"] = _normal(seed, name + ".act.beta", (ch,), std=0.5)
        sd[name + ".upsample.filter"] = filt.copy()
        sd[name + ".downsample.lowpass.filter"] = filt.copy()

    c0 = h["upsample_initial_channel"]
    conv("conv_pre", c0, h["gpt_dim"], 7)
    nk = len(h["resblock_kernel_sizes"])
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        conv(f"ups.{i}.0", ch, cin, k, transposed=True, taps_per_out=k // u)
        for j, ks in enumerate(h["resblock_kernel_sizes"]):
            p = f"resblocks.{i * nk + j}"
            for m in range(3):
                conv(f"{p}.convs1.{m}", ch, ch, ks, gain=0.7)
                conv(f"{p}.convs2.{m}", ch, ch, ks, gain=0.7)
            for m in range(6):
                act(f"{p}.activations.{m}", ch)
    act("activation_post", ch)
    conv("conv_post", 1, ch, 7, gain=0.1)

    # speaker conditioning (plain convs, never weight-normed: models.py:192-197)
    spk = h["speaker_embedding_dim"]
    wn_save, weight_norm = weight_norm, False
    conv("cond_layer", c0, spk, 1, gain=0.5)
    for i in range(len(h["upsample_rates"])):
        conv(f"conds.{i}", c0 // (2 ** (i + 1)), spk, 1, gain=0.5)
    weight_norm = wn_save

    if with_speaker_encoder:
        sd.update(make_ecapa_state_dict(seed, h["num_mels"], spk))
    return sd


def make_ecapa_state_dict(seed: int = 1234, num_mels: int = 100, lin_neurons: int = 512,
                          prefix: str = "speaker_encoder.") -> Dict[str, np.ndarray]:
    """ECAPA-TDNN parameters (key layout of ECAPA_TDNN.py:464-541 as seen in the reference state dict)."""
    sd: Dict[str, np.ndarray] = {}

    def conv(name, cout, cin, k):
        sd[prefix + name + ".conv.weight"] = _normal(seed, prefix + name + ".w", (cout, cin, k),
                                                     std=1.0 / math.sqrt(cin * k))
        sd[prefix + name + ".conv.bias"] = _normal(seed, prefix + name + ".b", (cout,), std=0.05)

    def bn(name, ch):
        sd[prefix + name + ".weight"] = _normal(seed, prefix + name + ".g", (ch,), std=0.1, mean=1.0)
        sd[prefix + name + ".bias"] = _normal(seed, prefix + name + ".beta", (ch,), std=0.1)
        sd[prefix + name + ".running_mean"] = _normal(seed, prefix + name + ".rm", (ch,), std=0.1)
        sd[prefix + name + ".running_var"] = (
            0.5 + _rng(seed, prefix + name + ".rv").random((ch,), dtype=np.float32)).astype(np.float32)
        sd[prefix + name + ".num_batches_tracked"] = np.zeros((), dtype=np.int64)

    def tdnn(name, cin, cout, k):
        conv(name + ".conv", cout, cin, k)
        bn(name + ".norm.norm", cout)

    tdnn("blocks.0", num_mels, 512, 5)
    for i in (1, 2, 3):
        tdnn(f"blocks.{i}.tdnn1", 512, 512, 1)
        for j in range(7):
            tdnn(f"blocks.{i}.res2net_block.blocks.{j}", 64, 64, 3)
        tdnn(f"blocks.{i}.tdnn2", 512, 512, 1)
        conv(f"blocks.{i}.se_block.conv1", 128, 512, 1)
        conv(f"blocks.{i}.se_block.conv2", 512, 128, 1)
    tdnn("mfa", 1536, 1536, 1)
    tdnn("asp.tdnn", 4608, 128, 1)
    conv("asp.conv", 1536, 128, 1)
    bn("asp_bn.norm", 3072)
    conv("fc", lin_neurons, 3072, 1)
    return sd


def make_latents(cfg_id: int, index: int, B: int, T: int, gpt_dim: int = 1024) -> np.ndarray:
    """x ~ N(0,1) fp32 [B,T,gpt_dim] (LayerNorm-like scale), seed = cfg_id*1000 + index."""
    return np.random.default_rng(cfg_id * 1000 + index).standard_normal((B, T, gpt_dim), dtype=np.float32)


def make_mel(seed: int = 7, Tm: int = 400, num_mels: int = 100, B: int = 1) -> np.ndarray:
    """Synthetic log-mel prompt N(-0.3, 2^2) [B,Tm,100]."""
    r = np.random.default_rng(seed)
    return (-0.3 + 2.0 * r.standard_normal((B, Tm, num_mels), dtype=np.float32)).astype(np.float32)


def make_speaker_embedding(seed: int = 7, B: int = 1, dim: int = 512) -> np.ndarray:
    """Stand-in ECAPA output [B,1,dim] for runs that do not exercise the speaker encoder."""
    r = np.random.default_rng(seed + 100003)
    return (0.5 * r.standard_normal((B, 1, dim), dtype=np.float32)).astype(np.float32)


def frames_for_seconds(sec: float, sr: int = 24000, hop: int = 1024) -> int:
    return int(math.ceil(sec * sr / hop))
