"""Drop-in `BigVGAN` generator module backed by libb200vgan.so.

Mirrors the reference's public surface for this path (reference: indextts/BigVGAN/models.py:130-260):

    g = BigVGAN(h, use_cuda_kernel=True)          # infer.py:110
    g.load_state_dict(ckpt["generator"])          # checkpoint layout (weight_g/weight_v, 1029 keys)
    g = g.to(device); g.remove_weight_norm(); g.eval()   # infer.py:114-117
    wav, _ = g(latent, mel_ref)                   # infer.py:458, :623   -> ([B,1,1024*T], None)

The torch side only holds parameters (so state dicts round-trip in either the checkpoint layout or
the folded layout) and device memory; every generator FLOP runs in hand-written sm_100a kernels
behind the C ABI.  There is NO torch / CPU fallback for the generator: construction fails without
the shared library, and forward fails without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import lib as _lib
from .ecapa import ECAPA_TDNN
from .synth import kaiser_filter

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def _hget(h, k, default=None):
    if isinstance(h, dict):
        return h.get(k, default)
    if hasattr(h, k):
        return getattr(h, k)
    try:
        return h.get(k, default)
    except Exception:
        return default


class _WNConv(nn.Module):
    """Parameter holder for a weight-normed Conv1d / ConvTranspose1d.  State-dict keys are
    `weight_g`, `weight_v`, `bias` until remove_weight_norm(), then `weight`, `bias`
    (torch.nn.utils.weight_norm semantics: g has shape [dim0,1,1], norm over all dims but 0)."""

    def __init__(self, shape: Tuple[int, int, int], n_bias: int):
        super().__init__()
        self.shape = tuple(shape)
        self.weight_g = nn.Parameter(torch.ones(shape[0], 1, 1))
        self.weight_v = nn.Parameter(torch.randn(*shape) * (1.0 / (shape[1] * shape[2]) ** 0.5))
        self.bias = nn.Parameter(torch.zeros(n_bias))
        with torch.no_grad():
            self.weight_g.copy_(self.weight_v.flatten(1).norm(dim=1).view(-1, 1, 1))

    @property
    def folded(self) -> bool:
        return "weight" in self._parameters and self._parameters["weight"] is not None

    def effective_weight(self) -> torch.Tensor:
        if self.folded:
            return self.weight.detach()
        v = self.weight_v.detach().double()
        g = self.weight_g.detach().double()
        nrm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        return (g * v / nrm).float()

    def remove_weight_norm(self):
        if self.folded:
            return
        w = self.effective_weight()
        del self._parameters["weight_g"]
        del self._parameters["weight_v"]
        self.weight = nn.Parameter(w)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        has_folded = prefix + "weight" in state_dict
        has_wn = prefix + "weight_g" in state_dict
        dev = self.bias.device
        if has_folded and not self.folded:
            del self._parameters["weight_g"]
            del self._parameters["weight_v"]
            self.weight = nn.Parameter(torch.zeros(*self.shape, device=dev))
        elif has_wn and self.folded:
            del self._parameters["weight"]
            self.weight_g = nn.Parameter(torch.ones(self.shape[0], 1, 1, device=dev))
            self.weight_v = nn.Parameter(torch.zeros(*self.shape, device=dev))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _Filter(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("filter", torch.from_numpy(kaiser_filter()).view(1, 1, 12).clone())


class _Down(nn.Module):
    def __init__(self):
        super().__init__()
        self.lowpass = _Filter()


class _SnakeBeta(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.alpha = nn.Parameter(torch.zeros(ch))  # log scale (activations.py:99-102)
        self.beta = nn.Parameter(torch.zeros(ch))


class _Activation1d(nn.Module):
    """Parameter holder with the reference key layout `act.{alpha,beta}`, `upsample.filter`,
    `downsample.lowpass.filter` (alias_free_torch/act.py:9-22)."""

    def __init__(self, ch):
        super().__init__()
        self.act = _SnakeBeta(ch)
        self.upsample = _Filter()
        self.downsample = _Down()


class _AMPBlock1(nn.Module):
    def __init__(self, ch, k, n_dil):
        super().__init__()
        self.convs1 = nn.ModuleList([_WNConv((ch, ch, k), ch) for _ in range(n_dil)])
        self.convs2 = nn.ModuleList([_WNConv((ch, ch, k), ch) for _ in range(n_dil)])
        self.activations = nn.ModuleList([_Activation1d(ch) for _ in range(2 * n_dil)])


class BigVGAN(nn.Module):
    PRECISIONS = ("auto", "fp32", "fp32tc", "fp16", "bf16")
    # what "auto" resolves to outside autocast: the fp32 tensor-core mode (fp32 tensors, waveform within 1e-4 of the
    # reference's fp32 output -- a tenth of the parity gate -- at 6x the speed of the CUDA-core mode); set to "fp32" for
    # the CUDA-core parity mode (1e-5)
    AUTO_FP32_PRECISION = "fp32tc"

    def __init__(self, h, use_cuda_kernel: bool = True, precision: str = "auto"):
        """`precision`: "fp32" (parity mode, the reference's default arithmetic on CUDA cores), "fp32tc" (fp32 storage and
        fp32-grade results with the convolutions on the tensor cores as three fp16 passes over split operands), "fp16" /
        "bf16" (tensor-core modes with 16-bit storage), or "auto" (default): what the reference module would compute in
        at this call site -- fp16 / bf16 under ``torch.amp.autocast`` with that dtype (infer.py:456, :613), fp32
        otherwise (``AUTO_FP32_PRECISION``: the fp32 tensor-core mode)."""
        super().__init__()
        self.h = h
        try:
            h["use_cuda_kernel"] = use_cuda_kernel  # models.py:141 does the same
        except Exception:
            pass
        self._cfg = _lib.make_config(h)
        self._libh = _lib.load()            # fails loudly if the extension is missing
        self.precision = precision
        cfg = self._cfg
        self.num_kernels = cfg.num_kernels
        self.num_upsamples = cfg.num_upsamples
        self.cond_in_each_up_layer = bool(cfg.cond_in_each_up_layer)
        c0, nd = cfg.upsample_initial_channel, cfg.num_dilations
        self.conv_pre = _WNConv((c0, cfg.gpt_dim, 7), c0)
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = c0
        self.hop = 1
        for i in range(cfg.num_upsamples):
            k, u = cfg.upsample_kernel_sizes[i], cfg.upsample_rates[i]
            self.hop *= u
            self.ups.append(nn.ModuleList([_WNConv((ch, ch // 2, k), ch // 2)]))
            ch //= 2
            for j in range(cfg.num_kernels):
                self.resblocks.append(_AMPBlock1(ch, cfg.resblock_kernel_sizes[j], nd))
        self.activation_post = _Activation1d(ch)
        self.conv_post = _WNConv((1, ch, 7), 1)
        spk = cfg.speaker_embedding_dim
        self.speaker_encoder = ECAPA_TDNN(int(_hget(h, "num_mels", 100)), lin_neurons=spk)
        self.cond_layer = nn.Conv1d(spk, c0, 1)
        if self.cond_in_each_up_layer:
            self.conds = nn.ModuleList([nn.Conv1d(spk, c0 >> (i + 1), 1) for i in range(cfg.num_upsamples)])
        # engine state
        self._handle = None
        self._engine_device = None
        self._engine_dirty = True
        self._plans: Dict[Tuple[Tuple[int, ...], int], C.c_void_p] = {}
        self._workspace: Optional[torch.Tensor] = None
        self._spk_cache: Dict[Tuple, torch.Tensor] = {}

    # ---- reference surface -------------------------------------------------------------------
    def remove_weight_norm(self):
        """Idempotent (models.py:252-260 would raise on a second call; a drop-in may be called twice)."""
        for m in self.modules():
            if isinstance(m, _WNConv):
                m.remove_weight_norm()
        self._engine_dirty = True

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine_dirty = True
        self._spk_cache.clear()
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._engine_dirty = True
        return r

    def forward(self, x, mel_ref, lens=None):
        """x [B,T,gpt_dim] latents, mel_ref [B',Tm,num_mels] -> (wav [B,1,T*hop] fp32, None)."""
        emb = self.speaker_embedding(mel_ref, lens)
        if emb.shape[0] == 2 * x.shape[0]:
            raise RuntimeError("contrastive branch (B' == 2B, models.py:205-209) is training-only")
        return self.forward_with_embedding(x, emb), None

    # ---- extensions --------------------------------------------------------------------------
    @torch.no_grad()
    def speaker_embedding(self, mel_ref, lens=None, cache_key=None):
        """ECAPA embedding [B',1,512]; pass `cache_key` (e.g. the prompt path) to reuse it."""
        if cache_key is not None and cache_key in self._spk_cache:
            return self._spk_cache[cache_key]
        dev = self.conv_pre.bias.device
        if dev.type != "cuda":
            raise _lib.BvgError("b200vgan has no CPU path: move the module to a CUDA (sm_100) device first")
        mel = mel_ref.to(device=dev, dtype=torch.float32).contiguous()
        Bp, Tm, M = mel.shape
        if M != self._cfg.num_mels:
            raise _lib.BvgError(f"mel width {M} != num_mels {self._cfg.num_mels}")
        rl = None if lens is None else torch.as_tensor(lens, dtype=torch.float32).to(dev).contiguous()
        with torch.cuda.device(dev):
            self._ensure_engine(dev)
            nbytes = int(self._libh.bvg_ecapa_workspace_bytes(self._handle, Bp, Tm))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            emb = torch.empty(Bp, 1, self._cfg.speaker_embedding_dim, device=dev, dtype=torch.float32)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self._libh.bvg_speaker_embedding(self._handle, mel.data_ptr(), Bp, Tm,
                                                        None if rl is None else rl.data_ptr(), emb.data_ptr(),
                                                        ws.data_ptr(), nbytes, stream))
        if cache_key is not None:
            self._spk_cache[cache_key] = emb
        return emb

    @torch.no_grad()
    def forward_with_embedding(self, x, emb, x_lens: Optional[Sequence[int]] = None, pcm16: bool = False):
        """Decode with a precomputed speaker embedding.  `x_lens` (host ints, latent frames per
        batch item) enables exact variable-length batches: item b is decoded as if alone.
        `pcm16=True` returns int16 PCM [B, T*hop] instead of the fp32 waveform: the reference callers'
        ``clamp(32767 * wav, -32767, 32767)`` + int16 cast (infer.py:462, :627, :650) fused into the last kernel."""
        if not x.is_cuda:
            raise _lib.BvgError("b200vgan has no CPU path: latents must live on a CUDA (sm_100) device")
        if x.dtype not in _DT:
            x = x.float()
        x = x.contiguous()
        B, T, D = x.shape
        if D != self._cfg.gpt_dim:
            raise _lib.BvgError(f"latent width {D} != gpt_dim {self._cfg.gpt_dim}")
        frames = tuple(int(v) for v in x_lens) if x_lens is not None else (T,) * B
        if len(frames) != B or max(frames) != T:
            raise _lib.BvgError("x_lens must have B entries and max(x_lens) == x.shape[1]")
        emb = emb.to(device=x.device, dtype=torch.float32).contiguous()
        if emb.shape[0] not in (1, B):
            raise _lib.BvgError("speaker embedding batch must be 1 or B")
        with torch.cuda.device(x.device):
            self._ensure_engine(x.device)
            plan = self._plan(frames, self._mode())
            ws = self._ensure_workspace(int(self._libh.bvg_plan_workspace_bytes(plan)), x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            if pcm16:
                pcm = torch.empty(B, T * self.hop, device=x.device, dtype=torch.int16)
                _lib.check(self._libh.bvg_forward_pcm16(self._handle, plan, x.data_ptr(), _DT[x.dtype], emb.data_ptr(),
                                                        emb.shape[0], pcm.data_ptr(), None, ws.data_ptr(), ws.numel(), stream))
                return pcm
            wav = torch.empty(B, 1, T * self.hop, device=x.device, dtype=torch.float32)
            _lib.check(self._libh.bvg_forward(self._handle, plan, x.data_ptr(), _DT[x.dtype], emb.data_ptr(),
                                              emb.shape[0], wav.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return wav

    @torch.no_grad()
    def forward_ragged(self, rows, frames: Sequence[int], emb, pcm16: bool = False, out: Optional[torch.Tensor] = None):
        """Variable-length batch without padding (bvg_forward_ragged): `rows` [sum(frames), gpt_dim] holds the
        segments' latent frames back to back, the result is ONE vector [sum(frames)*hop] (fp32 waveform, or int16
        PCM with `pcm16`) with segment b's samples starting at hop*sum(frames[:b]).  `out` may be a preallocated
        slice of a larger result arena.  Replaces the per-entry loop of srt_dubbing stretch_strategy.py:72-83 and the
        chunk concatenation of infer.py:439-463; each segment is decoded exactly as if alone."""
        if not rows.is_cuda:
            raise _lib.BvgError("b200vgan has no CPU path: latents must live on a CUDA (sm_100) device")
        if rows.dtype not in _DT:
            rows = rows.float()
        rows = rows.contiguous()
        frames = tuple(int(v) for v in frames)
        if rows.dim() != 2 or rows.shape[1] != self._cfg.gpt_dim or rows.shape[0] != sum(frames):
            raise _lib.BvgError("forward_ragged: rows must be [sum(frames), gpt_dim]")
        emb = emb.to(device=rows.device, dtype=torch.float32).contiguous()
        if emb.shape[0] not in (1, len(frames)):
            raise _lib.BvgError("speaker embedding batch must be 1 or the number of segments")
        n = rows.shape[0] * self.hop
        dt = torch.int16 if pcm16 else torch.float32
        if out is None:
            out = torch.empty(n, device=rows.device, dtype=dt)
        elif out.dtype != dt or out.numel() != n or not out.is_contiguous() or out.device != rows.device:
            raise _lib.BvgError("forward_ragged: `out` must be a contiguous device vector of sum(frames)*hop elements")
        with torch.cuda.device(rows.device):
            self._ensure_engine(rows.device)
            plan = self._plan(frames, self._mode())
            ws = self._ensure_workspace(int(self._libh.bvg_plan_workspace_bytes(plan)), rows.device)
            stream = torch.cuda.current_stream(rows.device).cuda_stream
            _lib.check(self._libh.bvg_forward_ragged(self._handle, plan, rows.data_ptr(), _DT[rows.dtype], emb.data_ptr(),
                                                     emb.shape[0], None if pcm16 else out.data_ptr(),
                                                     out.data_ptr() if pcm16 else None, ws.data_ptr(), ws.numel(), stream))
        return out

    def activation_kernel_name(self) -> str:
        """Which Activation1d kernel bvg_forward launches in the current precision mode (for bench reports)."""
        import os
        if self.resolved_precision() in ("fp32", "fp32tc"):
            return "act1d_c8_v3_kernel (Activation1d, register-streamed fp32)"
        if os.environ.get("BVG_ACT_MMA", "1") == "0":
            return "act1d_c8_v3_kernel (Activation1d, register-streamed, packed f32x2)"
        return "act1d_c8_mma_kernel (Activation1d: up-FIR and down-FIR as warp-level MMAs, SnakeBeta on CUDA cores)"

    def plans_created(self) -> int:
        return 0 if self._handle is None else int(self._libh.bvg_plans_created(self._handle))

    def resolved_precision(self) -> str:
        """The arithmetic the next forward will use ("auto" looks at the caller's autocast state)."""
        p = self.precision
        if p not in self.PRECISIONS:
            raise _lib.BvgError(f"precision must be one of {self.PRECISIONS}, got {p!r}")
        if p != "auto":
            return p
        if torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            return "fp16" if dt == torch.float16 else ("bf16" if dt == torch.bfloat16 else self.AUTO_FP32_PRECISION)
        return self.AUTO_FP32_PRECISION

    def _mode(self) -> int:
        return {"fp32": _lib.MODE_FP32, "bf16": _lib.MODE_BF16, "fp16": _lib.MODE_F16,
                "fp32tc": _lib.MODE_FP32_TC}[self.resolved_precision()]

    def num_launches(self, frames: Sequence[int]) -> int:
        return int(self._libh.bvg_plan_num_launches(self._plan(tuple(int(f) for f in frames), self._mode())))

    PROFILE_CLASSES = ("activation1d", "conv_tcgen05", "conv_cuda_core", "other")

    def profile_enable(self, on: bool = True):
        """Bracket every kernel launch of forward with CUDA events (see bvg_profile_read)."""
        if self._handle is None:
            raise _lib.BvgError("profile_enable: run one forward first (engine not built)")
        _lib.check(self._libh.bvg_profile_enable(self._handle, 1 if on else 0))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        ms, fl, by = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        _lib.check(self._libh.bvg_profile_read(self._handle, ms, fl, by, n))
        return {c: {"ms": ms[i], "flops": fl[i], "bytes": by[i], "launches": int(n[i])}
                for i, c in enumerate(self.PROFILE_CLASSES)}

    # ---- engine plumbing ---------------------------------------------------------------------
    def folded_state(self) -> Dict[str, torch.Tensor]:
        """Generator parameters in the folded layout, by reference state-dict name."""
        out = {}
        for name, m in self.named_modules():
            if isinstance(m, _WNConv):
                out[name + ".weight"] = m.effective_weight()
                out[name + ".bias"] = m.bias.detach()
            elif isinstance(m, _SnakeBeta):
                out[name + ".alpha"] = m.alpha.detach()
                out[name + ".beta"] = m.beta.detach()
        out["cond_layer.weight"] = self.cond_layer.weight.detach()
        out["cond_layer.bias"] = self.cond_layer.bias.detach()
        if self.cond_in_each_up_layer:
            for i, c in enumerate(self.conds):
                out[f"conds.{i}.weight"] = c.weight.detach()
                out[f"conds.{i}.bias"] = c.bias.detach()
        # speaker encoder: parameters and BatchNorm running statistics under their checkpoint names
        for k, v in self.speaker_encoder.state_dict().items():
            if not k.endswith("num_batches_tracked"):
                out["speaker_encoder." + k] = v.detach()
        return out

    def _release_engine(self):
        """Destroy the native handle, its plans and the workspace (they are bound to ONE device)."""
        for p in self._plans.values():
            self._libh.bvg_plan_destroy(p)
        self._plans.clear()
        if self._handle is not None:
            self._libh.bvg_destroy(self._handle)
        self._handle = None
        self._workspace = None
        self._engine_device = None
        self._spk_cache.clear()

    def _ensure_engine(self, device):
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._handle is not None and self._engine_device != device:
            # .to(another GPU): repacked weights, plan tables and the workspace live on the old device
            with torch.cuda.device(self._engine_device):
                torch.cuda.synchronize()
                self._release_engine()
        if self._handle is not None and not self._engine_dirty:
            return
        self._engine_device = device
        if self._handle is None:
            hd = C.c_void_p()
            _lib.check(self._libh.bvg_create(C.byref(self._cfg), C.byref(hd)))
            self._handle = hd
        stream = torch.cuda.current_stream(device).cuda_stream
        keep = []
        for name, t in self.folded_state().items():
            t = t.to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self._libh.bvg_set_weight(self._handle, name.encode(), t.data_ptr(), shape, t.dim(), 1, stream))
        _lib.check(self._libh.bvg_finalize(self._handle, stream))
        del keep
        self._engine_dirty = False

    def _plan(self, frames: Tuple[int, ...], mode: int):
        key = (frames, mode)
        p = self._plans.get(key)
        if p is None:
            if len(self._plans) >= 64:          # bound the cache for long dubbing jobs
                k0 = next(iter(self._plans))
                self._libh.bvg_plan_destroy(self._plans.pop(k0))
            p = C.c_void_p()
            arr = (C.c_int32 * len(frames))(*frames)
            _lib.check(self._libh.bvg_plan_create(self._handle, len(frames), arr, mode, C.byref(p)))
            self._plans[key] = p
        return p

    def _ensure_workspace(self, nbytes: int, device) -> torch.Tensor:
        ws = self._workspace
        if ws is None or ws.numel() < nbytes or ws.device != device:
            self._workspace = None
            ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
            self._workspace = ws     # contents do not matter: every forward re-clears the layout's guard rows
        return ws

    def __del__(self):
        try:
            self._release_engine()
        except Exception:
            pass
