"""ECAPA-TDNN speaker encoder: PARAMETER CONTAINER with the reference's state-dict key layout.

Nothing here computes: `BigVGAN.speaker_embedding` uploads these tensors to libb200vgan.so and calls
`bvg_speaker_embedding` (csrc/bvg_ecapa.cu, fp32 CUDA kernels).  The module tree only reproduces the
reference's key layout (reference: indextts/BigVGAN/ECAPA_TDNN.py:429-541, 231 keys, e.g.
``blocks.1.res2net_block.blocks.3.norm.norm.running_var``) so reference checkpoints load unchanged with
``load_state_dict(strict=True)``.  Calling a container raises: there is no torch / CPU path in the product
(the CPU restatement of ECAPA_TDNN.forward lives in oracle/bigvgan_oracle.py, test infrastructure).
"""
from __future__ import annotations

import torch.nn as nn


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("b200vgan.ecapa modules only hold parameters; use BigVGAN.speaker_embedding "
                           "(native kernels, csrc/bvg_ecapa.cu)")


class _SBConv(_Holder):
    """SpeechBrain Conv1d wrapper (nnet/CNN.py:305-545): key ``conv.{weight,bias}``."""

    def __init__(self, cin, cout, k, dilation=1):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, dilation=dilation)


class _SBNorm(_Holder):
    """SpeechBrain BatchNorm1d wrapper (nnet/normalization.py:13-108): key ``norm.*``."""

    def __init__(self, ch):
        super().__init__()
        self.norm = nn.BatchNorm1d(ch)


class _TDNN(_Holder):
    """TDNNBlock (ECAPA_TDNN.py:85-128): ``conv``, ``norm``."""

    def __init__(self, cin, cout, k, dilation):
        super().__init__()
        self.conv = _SBConv(cin, cout, k, dilation)
        self.norm = _SBNorm(cout)


class _Res2Net(_Holder):
    """Res2NetBlock (ECAPA_TDNN.py:131-191): ``blocks.{0..scale-2}``."""

    def __init__(self, ch, scale, k, dilation):
        super().__init__()
        self.blocks = nn.ModuleList([_TDNN(ch // scale, ch // scale, k, dilation) for _ in range(scale - 1)])


class _SE(_Holder):
    """SEBlock (ECAPA_TDNN.py:194-242): ``conv1``, ``conv2``."""

    def __init__(self, ch, se):
        super().__init__()
        self.conv1 = _SBConv(ch, se, 1)
        self.conv2 = _SBConv(se, ch, 1)


class _SERes2Net(_Holder):
    """SERes2NetBlock (ECAPA_TDNN.py:341-426)."""

    def __init__(self, ch, k, dilation):
        super().__init__()
        self.tdnn1 = _TDNN(ch, ch, 1, 1)
        self.res2net_block = _Res2Net(ch, 8, k, dilation)
        self.tdnn2 = _TDNN(ch, ch, 1, 1)
        self.se_block = _SE(ch, 128)


class _ASP(_Holder):
    """AttentiveStatisticsPooling (ECAPA_TDNN.py:245-338): ``tdnn``, ``conv``."""

    def __init__(self, ch, att=128):
        super().__init__()
        self.tdnn = _TDNN(ch * 3, att, 1, 1)
        self.conv = _SBConv(att, ch, 1)


class ECAPA_TDNN(_Holder):
    """ECAPA_TDNN (ECAPA_TDNN.py:429-541), channels [512,512,512,512,1536], kernels [5,3,3,3,1], dilations [1,2,3,4,1]."""

    def __init__(self, input_size=100, lin_neurons=512):
        super().__init__()
        self.blocks = nn.ModuleList([_TDNN(input_size, 512, 5, 1)] +
                                    [_SERes2Net(512, 3, d) for d in (2, 3, 4)])
        self.mfa = _TDNN(1536, 1536, 1, 1)
        self.asp = _ASP(1536)
        self.asp_bn = _SBNorm(3072)
        self.fc = _SBConv(3072, lin_neurons, 1)
