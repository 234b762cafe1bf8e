"""Activation-only micro run (hot-path packed kernel) for ncu / timing: B x C x T like a generator stage."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np, torch
from tests import gpu_util as G
from b200vgan import lib
B, C, T = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 96, 60160)))
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
x = torch.randn(B, C, T, device="cuda"); y = torch.empty_like(x)
la = (0.5 * torch.randn(C, device="cuda")); lb = (0.5 * torch.randn(C, device="cuda"))
L = G.L_()
for i in range(3):
    lib.check(L.bvg_activation1d_packed(x.data_ptr(), y.data_ptr(), la.data_ptr(), lb.data_ptr(), B, C, T, mode, None))
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
