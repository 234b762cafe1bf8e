"""Debug the fused act+conv kernel on small shapes: where does it differ from the unfused path?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np
from tests import gpu_util as G
from oracle import bigvgan_oracle as O
for (B, C, T, k, d) in ((1, 24, 700, 3, 1), (1, 24, 100, 3, 1), (1, 48, 700, 7, 3), (2, 96, 300, 11, 5), (1, 192, 400, 3, 1)):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    la = (0.3 * rng.standard_normal(C)).astype(np.float32); lb = (0.3 * rng.standard_normal(C)).astype(np.float32)
    w = (rng.standard_normal((C, C, k)) / np.sqrt(C * k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    act = O.activation1d(G.bf16_round(x), la.astype(np.float64), lb.astype(np.float64))
    ref = O.conv1d(G.bf16_round(act), G.bf16_round(w), b.astype(np.float64), dilation=d, padding=O.get_padding(k, d))
    os.environ["BVG_FUSE_ACT"] = "1"
    y, fused = G.act_conv1d(x, la, lb, w, b, None, k, d, 1)
    err = np.abs(y - ref)
    bad = np.argwhere(~np.isfinite(y) | (err > 3e-2 * np.abs(ref).max()))
    print(f"case {(B,C,T,k,d)} fused={fused} max-err {np.nanmax(err):.3e} (scale {np.abs(ref).max():.2f}) nan {np.isnan(y).sum()} nbad {len(bad)}")
    if len(bad):
        ts = np.unique(bad[:, 2]); cs = np.unique(bad[:, 1])
        print("   bad t range", ts[:10], "...", ts[-10:], " n_t", len(ts), " bad channels", cs[:16], "n_c", len(cs))
