"""CPU oracle for the BigVGAN2 speech-code decoder path of scwf/index-tts-dubbing.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
path (``index-tts-dubbing_b200/``) never imports this file and has no CPU fallback.

It is a plain numpy *restatement* of the reference algorithm (layout ``[B, C, T]`` like
the reference), written step by step after the reference's own torch code -- it is
not a copy: every function cites the reference ``file:line`` it follows (paths are
relative to the reference checkout).  The dense / transposed / depthwise convolutions
that the reference delegates to ``torch.nn.functional`` (ATen, a third-party dependency;
reference pins ``torch>=2.1.2`` in ``setup.py:48``) are restated from their published
definition (cross-correlation with zero padding; scatter-add for the transposed form).

Parity pinning: the reference ships no golden vectors or known-answer tests for this
path.  The oracle is therefore pinned against outputs of the
*unmodified reference itself*, run on CPU in the build container by
``oracle/gen_golden.py`` and committed as fixtures under ``tests/golden/``
(``tests/test_oracle_golden.py`` checks every one of them).

Default arithmetic is float64 (the oracle's own rounding noise is then ~1e-15, far
below the 1e-3 parity gate); pass ``dtype=np.float32`` for the timed CPU baseline.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np

# ---------------------------------------------------------------------------------------
# Architecture constants of checkpoints/config.yaml:51-70 (bigvgan section)
# ---------------------------------------------------------------------------------------
DEFAULT_H = dict(
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


# ---------------------------------------------------------------------------------------
# Third-party arithmetic restated (torch.nn.functional.conv1d / conv_transpose1d / pad)
# ---------------------------------------------------------------------------------------
def pad1d(x: np.ndarray, left: int, right: int, mode: str) -> np.ndarray:
    """F.pad on the last axis; modes used by the path: constant(0), replicate, reflect."""
    if left == 0 and right == 0:
        return x
    np_mode = {"zeros": "constant", "replicate": "edge", "reflect": "reflect"}[mode]
    return np.pad(x, [(0, 0)] * (x.ndim - 1) + [(left, right)], mode=np_mode)


def conv1d(x, w, b=None, dilation=1, padding=0, stride=1):
    """Dense cross-correlation, ``x[B,Cin,L]``, ``w[Cout,Cin,k]`` (torch F.conv1d semantics).

    Call sites: models.py:26-41 (AMPBlock1 convs), :149 (conv_pre), :184 (conv_post),
    :192-197 (cond layers), ECAPA_TDNN.py via nnet/CNN.py:430-460.
    """
    B, Cin, L = x.shape
    Cout, Cin2, k = w.shape
    assert Cin == Cin2
    xp = pad1d(x, padding, padding, "zeros")
    Lp = xp.shape[-1]
    Lout = (Lp - dilation * (k - 1) - 1) // stride + 1
    out = np.zeros((B, Cout, Lout), dtype=x.dtype)
    for j in range(k):
        sl = xp[:, :, j * dilation: j * dilation + (Lout - 1) * stride + 1: stride]
        out += np.matmul(w[None, :, :, j], sl)
    if b is not None:
        out += b[None, :, None]
    return out


def conv_transpose1d(x, w, b, stride, padding):
    """Transposed conv, ``x[B,Cin,L]``, ``w[Cin,Cout,k]`` (torch F.conv_transpose1d).

    Call site: models.py:155-161, :230-231.  out_full[i*stride + kk] += w[:,:,kk]^T x[:,i];
    then ``padding`` samples are cropped from both ends.
    """
    B, Cin, L = x.shape
    Cin2, Cout, k = w.shape
    assert Cin == Cin2
    full = np.zeros((B, Cout, (L - 1) * stride + k), dtype=x.dtype)
    for kk in range(k):
        y = np.matmul(w[None, :, :, kk].transpose(0, 2, 1), x)  # [B,Cout,L]
        full[:, :, kk: kk + (L - 1) * stride + 1: stride] += y
    out = full[:, :, padding: full.shape[-1] - padding]
    if b is not None:
        out = out + b[None, :, None]
    return out


def depthwise_conv1d(x, f, stride):
    """``F.conv1d(x, f.expand(C,-1,-1), stride, groups=C)`` -- filter.py:93."""
    k = f.shape[-1]
    Lout = (x.shape[-1] - k) // stride + 1
    out = np.zeros(x.shape[:-1] + (Lout,), dtype=x.dtype)
    for j in range(k):
        out += f[j] * x[..., j: j + (Lout - 1) * stride + 1: stride]
    return out


def depthwise_conv_transpose1d(x, f, stride):
    """``F.conv_transpose1d(x, f.expand(C,-1,-1), stride, groups=C)`` -- resample.py:29-30."""
    k = f.shape[-1]
    L = x.shape[-1]
    out = np.zeros(x.shape[:-1] + ((L - 1) * stride + k,), dtype=x.dtype)
    for j in range(k):
        out[..., j: j + (L - 1) * stride + 1: stride] += f[j] * x
    return out


# ---------------------------------------------------------------------------------------
# alias_free_torch: filter.py / resample.py / act.py, activations.py
# ---------------------------------------------------------------------------------------
def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> np.ndarray:
    """alias_free_torch/filter.py:29-58.  Returns the taps as a 1-D float64 array."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = np.kaiser(kernel_size, beta)  # torch.kaiser_window(periodic=False)
    if even:
        time = np.arange(-half_size, half_size) + 0.5
    else:
        time = np.arange(kernel_size) - half_size
    if cutoff == 0:
        return np.zeros_like(time)
    filt = 2 * cutoff * window * np.sinc(2 * cutoff * time)
    filt /= filt.sum()
    return filt


def upsample1d(x, ratio=2, kernel_size=12):
    """alias_free_torch/resample.py:10-33 (UpSample1d)."""
    f = kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, kernel_size).astype(x.dtype)
    pad = kernel_size // ratio - 1
    pad_left = pad * ratio + (kernel_size - ratio) // 2
    pad_right = pad * ratio + (kernel_size - ratio + 1) // 2
    xp = pad1d(x, pad, pad, "replicate")
    y = ratio * depthwise_conv_transpose1d(xp, f, ratio)
    return y[..., pad_left: y.shape[-1] - pad_right]


def downsample1d(x, ratio=2, kernel_size=12):
    """alias_free_torch/resample.py:36-49 + filter.py:61-96 (LowPassFilter1d, stride=ratio)."""
    f = kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, kernel_size).astype(x.dtype)
    even = kernel_size % 2 == 0
    pad_left = kernel_size // 2 - int(even)
    pad_right = kernel_size // 2
    xp = pad1d(x, pad_left, pad_right, "replicate")
    return depthwise_conv1d(xp, f, ratio)


def snakebeta(x, log_alpha, log_beta):
    """activations.py:109-122 with ``alpha_logscale=True`` (config.yaml:69-70)."""
    alpha = np.exp(log_alpha)[None, :, None]
    beta = np.exp(log_beta)[None, :, None]
    return x + (1.0 / (beta + 0.000000001)) * np.sin(x * alpha) ** 2


def activation1d(x, log_alpha, log_beta):
    """alias_free_torch/act.py:24-29: upsample x2 -> SnakeBeta -> downsample x2."""
    y = upsample1d(x)
    y = snakebeta(y, log_alpha.astype(x.dtype), log_beta.astype(x.dtype))
    return downsample1d(y)


def activation1d_closed_form(x, log_alpha, log_beta):
    """Closed form of :func:`activation1d` -- what the CUDA kernels
    implement.  Kept here so the (non-GPU) tests can show it equals the step-by-step form.

    y[t] = sum_k f[k] s[clamp(2t+k-5, 0, 2L-1)],  s = u + sin^2(a u)/(b+1e-9),
    u[m] = 2 sum_{k == (m+5) mod 2} f[k] x[clamp((m+5-k)/2, 0, L-1)].
    """
    f = kaiser_sinc_filter1d(0.25, 0.3, 12).astype(x.dtype)
    L = x.shape[-1]
    m = np.arange(2 * L)
    u = np.zeros(x.shape[:-1] + (2 * L,), dtype=x.dtype)
    for k in range(12):
        sel = ((m + 5 - k) % 2) == 0
        idx = np.clip((m + 5 - k) // 2, 0, L - 1)
        u += np.where(sel, 2.0 * f[k], 0.0).astype(x.dtype) * x[..., idx]
    s = snakebeta(u, log_alpha.astype(x.dtype), log_beta.astype(x.dtype))
    t = np.arange(L)
    y = np.zeros_like(x)
    for k in range(12):
        y += f[k] * s[..., np.clip(2 * t + k - 5, 0, 2 * L - 1)]
    return y


# ---------------------------------------------------------------------------------------
# BigVGAN generator: models.py
# ---------------------------------------------------------------------------------------
def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """BigVGAN/utils.py:59-60."""
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """``remove_weight_norm`` (models.py:252-260): w = g * v / ||v||, norm over all dims but 0."""
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (g.astype(np.float64) * v.astype(np.float64) / norm).astype(v.dtype)


def fold_state_dict(sd: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Checkpoint layout (``*.weight_g`` / ``*.weight_v``) -> folded layout (``*.weight``)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            out[base + ".weight"] = fold_weight_norm(v, sd[base + ".weight_v"])
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def amp_block1(x, sd, prefix, kernel_size, dilations=(1, 3, 5)):
    """AMPBlock1.forward -- models.py:65-74."""
    for i, d in enumerate(dilations):
        a1 = f"{prefix}.activations.{2 * i}.act"
        a2 = f"{prefix}.activations.{2 * i + 1}.act"
        xt = activation1d(x, sd[a1 + ".alpha"], sd[a1 + ".beta"])
        xt = conv1d(xt, sd[f"{prefix}.convs1.{i}.weight"], sd[f"{prefix}.convs1.{i}.bias"],
                    dilation=d, padding=get_padding(kernel_size, d))
        xt = activation1d(xt, sd[a2 + ".alpha"], sd[a2 + ".beta"])
        xt = conv1d(xt, sd[f"{prefix}.convs2.{i}.weight"], sd[f"{prefix}.convs2.{i}.bias"],
                    dilation=1, padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def bigvgan_forward_with_embedding(x, spk, sd, h=None, dtype=np.float64):
    """BigVGAN.forward after the speaker encoder -- models.py:210-250.

    ``x[B,T,gpt_dim]`` latents, ``spk[B',1,512]`` (output of ECAPA, B' in {1,B}),
    ``sd`` folded state dict (numpy).  Returns ``wav[B,1,T*prod(upsample_rates)]``.
    """
    h = dict(DEFAULT_H, **(h or {}))
    sd = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()
          if not k.startswith("speaker_encoder.") and not k.endswith("filter")}
    x = np.asarray(x, dtype=dtype)
    spk = np.asarray(spk, dtype=dtype).transpose(0, 2, 1)  # [B',512,1]   models.py:210
    assert not h["feat_upsample"]
    x = x.transpose(0, 2, 1)  # models.py:220
    x = conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)  # :224
    x = x + conv1d(spk, sd["cond_layer.weight"], sd["cond_layer.bias"])  # :226
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], u, (k - u) // 2)  # :230
        if h["cond_d_vector_in_each_upsampling_layer"]:
            x = x + conv1d(spk, sd[f"conds.{i}.weight"], sd[f"conds.{i}.bias"])  # :233-234
        xs = None
        for j in range(nk):  # :237-242
            r = amp_block1(x, sd, f"resblocks.{i * nk + j}", h["resblock_kernel_sizes"][j],
                           h["resblock_dilation_sizes"][j])
            xs = r if xs is None else xs + r
        x = xs / nk  # :243
    x = activation1d(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"])  # :246
    x = conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)  # :247
    return np.tanh(x)  # :248


def bigvgan_forward(x, mel_ref, sd, h=None, dtype=np.float64, lens=None):
    """BigVGAN.forward -- models.py:201-250: ECAPA embedding of ``mel_ref[B',Tm,100]`` then decode."""
    spk = ecapa_forward(mel_ref, sd, prefix="speaker_encoder.", lengths=lens, dtype=dtype)
    return bigvgan_forward_with_embedding(x, spk, sd, h, dtype)


# ---------------------------------------------------------------------------------------
# ECAPA-TDNN speaker encoder: ECAPA_TDNN.py, nnet/CNN.py, nnet/normalization.py
# ---------------------------------------------------------------------------------------
def _sb_conv1d(x, sd, name, kernel_size, dilation=1):
    """SpeechBrain Conv1d, padding='same', padding_mode='reflect', stride 1 --
    nnet/CNN.py:430-433 (padding), :519-545 (get_padding_elem: floor(d(k-1)/2) per side)."""
    pad = (dilation * (kernel_size - 1)) // 2
    xp = pad1d(x, pad, pad, "reflect")
    return conv1d(xp, sd[name + ".conv.weight"], sd[name + ".conv.bias"], dilation=dilation)


def _bn_eval(x, sd, name, eps=1e-5):
    """BatchNorm1d in eval mode -- nnet/normalization.py:13-108 (torch.nn.BatchNorm1d, eps 1e-5)."""
    w, b = sd[name + ".norm.weight"], sd[name + ".norm.bias"]
    m, v = sd[name + ".norm.running_mean"], sd[name + ".norm.running_var"]
    scale = w / np.sqrt(v + eps)
    return x * scale[None, :, None] + (b - m * scale)[None, :, None]


def _tdnn_block(x, sd, name, kernel_size, dilation):
    """TDNNBlock.forward -- ECAPA_TDNN.py:126-128: norm(activation(conv(x)))."""
    y = _sb_conv1d(x, sd, name + ".conv", kernel_size, dilation)
    return _bn_eval(np.maximum(y, 0), sd, name + ".norm")


def _length_mask(lengths, L, dtype):
    """length_to_mask(lengths * L, max_len=L) -- ECAPA_TDNN.py:16-61."""
    # the reference multiplies the float32 relative lengths by L in float32 (0.6f*60 == 36.0f)
    prod = np.asarray(lengths, dtype=np.float32) * np.float32(L)
    return (np.arange(L, dtype=np.float32)[None, :] < prod[:, None]).astype(dtype)


def _res2net(x, sd, name, scale, kernel_size, dilation):
    """Res2NetBlock.forward -- ECAPA_TDNN.py:179-191."""
    ys = []
    y_i = None
    for i, x_i in enumerate(np.split(x, scale, axis=1)):
        if i == 0:
            y_i = x_i
        elif i == 1:
            y_i = _tdnn_block(x_i, sd, f"{name}.blocks.{i - 1}", kernel_size, dilation)
        else:
            y_i = _tdnn_block(x_i + y_i, sd, f"{name}.blocks.{i - 1}", kernel_size, dilation)
        ys.append(y_i)
    return np.concatenate(ys, axis=1)


def _se_block(x, sd, name, lengths):
    """SEBlock.forward -- ECAPA_TDNN.py:228-242."""
    L = x.shape[-1]
    if lengths is not None:
        mask = _length_mask(lengths, L, x.dtype)[:, None, :]
        s = (x * mask).sum(axis=2, keepdims=True) / mask.sum(axis=2, keepdims=True)
    else:
        s = x.mean(axis=2, keepdims=True)
    s = np.maximum(_sb_conv1d(s, sd, name + ".conv1", 1), 0)
    s = 1.0 / (1.0 + np.exp(-_sb_conv1d(s, sd, name + ".conv2", 1)))
    return s * x


def _se_res2net_block(x, sd, name, kernel_size, dilation, lengths):
    """SERes2NetBlock.forward -- ECAPA_TDNN.py:413-426 (in==out channels: no shortcut conv)."""
    residual = x
    y = _tdnn_block(x, sd, name + ".tdnn1", 1, 1)
    y = _res2net(y, sd, name + ".res2net_block", 8, kernel_size, dilation)
    y = _tdnn_block(y, sd, name + ".tdnn2", 1, 1)
    y = _se_block(y, sd, name + ".se_block", lengths)
    return y + residual


def _asp(x, sd, name, lengths, eps=1e-12):
    """AttentiveStatisticsPooling.forward -- ECAPA_TDNN.py:282-338 (global_context=True)."""
    B, C, L = x.shape

    def stats(x, m):
        mean = (m * x).sum(axis=2)
        std = np.sqrt(np.maximum((m * (x - mean[:, :, None]) ** 2).sum(axis=2), eps))
        return mean, std

    if lengths is None:
        lengths = np.ones(B, dtype=x.dtype)
    mask = _length_mask(lengths, L, x.dtype)[:, None, :]
    total = mask.sum(axis=2, keepdims=True)
    mean, std = stats(x, mask / total)
    attn = np.concatenate([x, np.repeat(mean[:, :, None], L, 2), np.repeat(std[:, :, None], L, 2)], axis=1)
    attn = _sb_conv1d(np.tanh(_tdnn_block(attn, sd, name + ".tdnn", 1, 1)), sd, name + ".conv", 1)
    attn = np.where(mask == 0, -np.inf, attn)
    attn = attn - attn.max(axis=2, keepdims=True)
    e = np.exp(attn)
    attn = e / e.sum(axis=2, keepdims=True)
    mean, std = stats(x, attn)
    return np.concatenate([mean, std], axis=1)[:, :, None]


def ecapa_forward(mel, sd, prefix="speaker_encoder.", lengths=None, dtype=np.float64):
    """ECAPA_TDNN.forward -- ECAPA_TDNN.py:543-581.  ``mel[B,Tm,100]`` -> ``[B,1,512]``.

    Architecture (ECAPA_TDNN.py:464-541 with the defaults of :470-481): TDNN(k5) ->
    3 x SE-Res2Net(k3, dilation 2/3/4) -> MFA 1x1 -> ASP -> BN -> fc 1x1.
    """
    sd = {k[len(prefix):]: np.asarray(v, dtype=dtype) for k, v in sd.items()
          if k.startswith(prefix) and not k.endswith("num_batches_tracked")}
    x = np.asarray(mel, dtype=dtype).transpose(0, 2, 1)
    if lengths is not None:
        lengths = np.asarray(lengths, dtype=dtype)
    xl = []
    x = _tdnn_block(x, sd, "blocks.0", 5, 1)
    xl.append(x)
    for i, d in zip((1, 2, 3), (2, 3, 4)):
        x = _se_res2net_block(x, sd, f"blocks.{i}", 3, d, lengths)
        xl.append(x)
    x = np.concatenate(xl[1:], axis=1)
    x = _tdnn_block(x, sd, "mfa", 1, 1)
    x = _asp(x, sd, "asp", lengths)
    # asp_bn is a BatchNorm1d whose inner module is ".norm" (state-dict key asp_bn.norm.*)
    w, b = sd["asp_bn.norm.weight"], sd["asp_bn.norm.bias"]
    m, v = sd["asp_bn.norm.running_mean"], sd["asp_bn.norm.running_var"]
    scale = w / np.sqrt(v + 1e-5)
    x = x * scale[None, :, None] + (b - m * scale)[None, :, None]
    x = conv1d(x, sd["fc.conv.weight"], sd["fc.conv.bias"])
    return x.transpose(0, 2, 1)


# ---------------------------------------------------------------------------------------
# Metric helper: log-mel of the reference front-end (utils/feature_extractors.py:24-50)
# ---------------------------------------------------------------------------------------
def _hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def _mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


def mel_filterbank(n_freqs=513, f_min=0.0, f_max=12000.0, n_mels=100, sample_rate=24000):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') restated."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = _mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))  # [n_freqs, n_mels]


def log_mel(wav, sample_rate=24000, n_fft=1024, hop=256, n_mels=100, clip_val=1e-7):
    """MelSpectrogramFeatures.forward (padding='center') -- feature_extractors.py:24-50,
    safe_log -- utils/common.py:110-121.  ``wav[..., N]`` -> ``[..., n_mels, frames]``."""
    wav = np.asarray(wav, dtype=np.float64)
    lead = wav.shape[:-1]
    w = wav.reshape(-1, wav.shape[-1])
    w = np.pad(w, [(0, 0), (n_fft // 2, n_fft // 2)], mode="reflect")
    n_frames = 1 + (w.shape[-1] - n_fft) // hop
    window = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)  # periodic hann
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = w[:, idx] * window
    spec = np.abs(np.fft.rfft(frames, axis=-1))  # power=1
    mel = spec @ mel_filterbank(n_fft // 2 + 1, 0.0, sample_rate / 2, n_mels, sample_rate)
    mel = np.log(np.maximum(mel, clip_val)).transpose(0, 2, 1)
    return mel.reshape(lead + mel.shape[1:])


def mel_l1(wav_a, wav_b) -> float:
    """Mean absolute log-mel difference (the mel-L1 of BASELINE.md section 5)."""
    return float(np.mean(np.abs(log_mel(wav_a) - log_mel(wav_b))))


def snr_db(ref, test) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    return float(10.0 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-300)))
