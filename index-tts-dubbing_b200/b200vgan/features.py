"""GPU log-mel front-end with the reference's interface: `MelSpectrogramFeatures()(audio)` (indextts/utils/
feature_extractors.py:24-50), computed by the hand-written kernel behind `bvg_log_mel` (csrc/bvg_mel.cu) so the
prompt's conditioning mel never leaves the device on its way into `BigVGAN.forward(latent, mel.transpose(1, 2))`."""
from __future__ import annotations

import torch
from torch import nn

from . import lib as _lib


class MelSpectrogramFeatures(nn.Module):
    """Same constructor arguments and output as the reference class: audio [B, N] (or [N]) at `sample_rate` ->
    log-mel [B, n_mels, 1 + N // hop_length] fp32.  Only the configuration the reference deploys is implemented
    natively (n_fft = win_length = 1024, padding="center", normalize=False); anything else raises."""

    def __init__(self, sample_rate=24000, n_fft=1024, hop_length=256, win_length=None, n_mels=100, mel_fmin=0, mel_fmax=None,
                 normalize=False, padding="center"):
        super().__init__()
        if padding not in ["center", "same"]:
            raise ValueError("Padding must be 'center' or 'same'.")
        if padding != "center" or n_fft != 1024 or (win_length not in (None, n_fft)) or normalize:
            raise _lib.BvgError("b200vgan.MelSpectrogramFeatures: only n_fft=1024, win_length=n_fft, padding='center', "
                                "normalize=False (the reference's deployed configuration) has a native kernel")
        self.sample_rate, self.hop_length, self.n_mels = int(sample_rate), int(hop_length), int(n_mels)
        self.f_min = float(mel_fmin)
        self.f_max = -1.0 if mel_fmax is None else float(mel_fmax)
        self._libh = _lib.load()

    @torch.no_grad()
    def forward(self, audio, transposed: bool = False, **kwargs):
        if not audio.is_cuda:
            raise _lib.BvgError("b200vgan has no CPU path: move the audio to a CUDA (sm_100) device first")
        squeeze = audio.dim() == 1
        a = audio.reshape(-1, audio.shape[-1]).to(torch.float32).contiguous()
        B, N = a.shape
        frames = int(self._libh.bvg_mel_frames(N, self.hop_length))
        shape = (B, frames, self.n_mels) if transposed else (B, self.n_mels, frames)
        mel = torch.empty(shape, device=a.device, dtype=torch.float32)
        with torch.cuda.device(a.device):
            _lib.check(self._libh.bvg_log_mel(a.data_ptr(), B, N, self.sample_rate, self.hop_length, self.n_mels, self.f_min,
                                              self.f_max, mel.data_ptr(), 1 if transposed else 0,
                                              torch.cuda.current_stream(a.device).cuda_stream))
        return mel[0] if squeeze else mel
