"""ECAPA-TDNN speaker encoder: PARAMETER CONTAINER with the reference's state-dict key layout.

The product path does not run this module: `BigVGAN.speaker_embedding` uploads these tensors to
libb200vgan.so and calls `bvg_speaker_embedding` (csrc/bvg_ecapa.cu, fp32 CUDA kernels).  The module
tree reproduces the reference's key layout (reference: indextts/BigVGAN/ECAPA_TDNN.py:429-541, 231
keys, e.g. ``blocks.1.res2net_block.blocks.3.norm.norm.running_var``) so reference checkpoints load
unchanged with ``load_state_dict(strict=True)``.  `forward` restates ECAPA_TDNN.py:543-581 with torch
operators and is kept only as a debugging aid for tests; nothing in b200vgan calls it.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _SBConv(nn.Module):
    """SpeechBrain Conv1d wrapper: 'same' reflect padding (nnet/CNN.py:430-433, :519-545)."""

    def __init__(self, cin, cout, k, dilation=1):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, dilation=dilation)
        self.pad = (dilation * (k - 1)) // 2

    def forward(self, x):
        if self.pad:
            x = F.pad(x, (self.pad, self.pad), mode="reflect")
        return self.conv(x)


class _SBNorm(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.norm = nn.BatchNorm1d(ch)

    def forward(self, x):
        return self.norm(x)


class _TDNN(nn.Module):
    """conv -> ReLU -> BatchNorm (ECAPA_TDNN.py:126-128)."""

    def __init__(self, cin, cout, k, dilation):
        super().__init__()
        self.conv = _SBConv(cin, cout, k, dilation)
        self.norm = _SBNorm(cout)

    def forward(self, x):
        return self.norm(torch.relu(self.conv(x)))


class _Res2Net(nn.Module):
    def __init__(self, ch, scale, k, dilation):
        super().__init__()
        self.scale = scale
        self.blocks = nn.ModuleList([_TDNN(ch // scale, ch // scale, k, dilation) for _ in range(scale - 1)])

    def forward(self, x):  # ECAPA_TDNN.py:179-191
        ys, y = [], None
        for i, xi in enumerate(torch.chunk(x, self.scale, dim=1)):
            if i == 0:
                y = xi
            elif i == 1:
                y = self.blocks[i - 1](xi)
            else:
                y = self.blocks[i - 1](xi + y)
            ys.append(y)
        return torch.cat(ys, dim=1)


def _length_mask(lengths, L, dtype, device):
    # length_to_mask(lengths * L, max_len=L) -- ECAPA_TDNN.py:16-61
    lim = (lengths * L).to(device)
    return (torch.arange(L, device=device, dtype=lim.dtype)[None, :] < lim[:, None]).to(dtype)


class _SE(nn.Module):
    def __init__(self, ch, se):
        super().__init__()
        self.conv1 = _SBConv(ch, se, 1)
        self.conv2 = _SBConv(se, ch, 1)

    def forward(self, x, lengths=None):  # ECAPA_TDNN.py:228-242
        if lengths is not None:
            m = _length_mask(lengths, x.shape[-1], x.dtype, x.device)[:, None, :]
            s = (x * m).sum(dim=2, keepdim=True) / m.sum(dim=2, keepdim=True)
        else:
            s = x.mean(dim=2, keepdim=True)
        s = torch.sigmoid(self.conv2(torch.relu(self.conv1(s))))
        return s * x


class _SERes2Net(nn.Module):
    def __init__(self, ch, k, dilation):
        super().__init__()
        self.tdnn1 = _TDNN(ch, ch, 1, 1)
        self.res2net_block = _Res2Net(ch, 8, k, dilation)
        self.tdnn2 = _TDNN(ch, ch, 1, 1)
        self.se_block = _SE(ch, 128)

    def forward(self, x, lengths=None):  # ECAPA_TDNN.py:413-426
        y = self.tdnn2(self.res2net_block(self.tdnn1(x)))
        return self.se_block(y, lengths) + x


class _ASP(nn.Module):
    def __init__(self, ch, att=128):
        super().__init__()
        self.tdnn = _TDNN(ch * 3, att, 1, 1)
        self.conv = _SBConv(att, ch, 1)

    def forward(self, x, lengths=None):  # ECAPA_TDNN.py:282-338
        L = x.shape[-1]

        def stats(x, m, eps=1e-12):
            mean = (m * x).sum(2)
            std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
            return mean, std

        if lengths is None:
            lengths = torch.ones(x.shape[0], device=x.device)
        mask = _length_mask(lengths, L, x.dtype, x.device)[:, None, :]
        total = mask.sum(dim=2, keepdim=True)
        mean, std = stats(x, mask / total)
        attn = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, L), std.unsqueeze(2).expand(-1, -1, L)], dim=1)
        attn = self.conv(torch.tanh(self.tdnn(attn)))
        attn = attn.masked_fill(mask == 0, float("-inf"))
        attn = F.softmax(attn, dim=2)
        mean, std = stats(x, attn)
        return torch.cat((mean, std), dim=1).unsqueeze(2)


class ECAPA_TDNN(nn.Module):
    def __init__(self, input_size=100, lin_neurons=512):
        super().__init__()
        self.blocks = nn.ModuleList([_TDNN(input_size, 512, 5, 1)] +
                                    [_SERes2Net(512, 3, d) for d in (2, 3, 4)])
        self.mfa = _TDNN(1536, 1536, 1, 1)
        self.asp = _ASP(1536)
        self.asp_bn = _SBNorm(3072)
        self.fc = _SBConv(3072, lin_neurons, 1)

    @torch.no_grad()
    def forward(self, x, lengths=None):  # ECAPA_TDNN.py:543-581
        x = x.transpose(1, 2).float()
        xl = []
        for i, layer in enumerate(self.blocks):
            x = layer(x) if i == 0 else layer(x, lengths)
            xl.append(x)
        x = self.mfa(torch.cat(xl[1:], dim=1))
        x = self.asp_bn(self.asp(x, lengths))
        return self.fc(x).transpose(1, 2)
