"""GPU diagnostic: run the tcgen05 conv kernel on a few shapes and report the error against a
bf16-rounded fp64 reference, for both settings of the LBO/SBO debug knob.  Not a test."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))

import numpy as np
import torch

from oracle import bigvgan_oracle as O
from tests import gpu_util as G

CASES = [(1, 64, 64, 128, 1, 1), (1, 64, 64, 128, 3, 1), (1, 64, 64, 300, 3, 1), (1, 128, 64, 200, 3, 1),
         (1, 192, 192, 260, 7, 3), (1, 96, 96, 300, 11, 5), (1, 24, 24, 500, 7, 1), (2, 768, 768, 77, 3, 1)]
ok_any = False
for swap in ("0", "1"):
    os.environ["BVG_UMMA_SWAP"] = swap
    for (B, Cin, Cout, T, k, d) in CASES:
        rng = np.random.default_rng(0)
        x = rng.standard_normal((B, Cin, T)).astype(np.float32)
        w = (rng.standard_normal((Cout, Cin, k)) / np.sqrt(Cin * k)).astype(np.float32)
        ref = O.conv1d(G.bf16_round(x), G.bf16_round(w), None, dilation=d, padding=O.get_padding(k, d))
        try:
            t0 = time.time()
            y = G.conv1d(x, w, None, None, k, d, 1)
            err = np.abs(y - ref).max() / np.abs(ref).max()
            bad = np.argwhere(np.abs(y - ref) > 2e-2 * np.abs(ref).max())
            print(f"swap={swap} case={(B, Cin, Cout, T, k, d)} rel-err={err:.3e} nbad={len(bad)} "
                  f"first-bad={bad[:3].tolist()} dt={time.time() - t0:.2f}s", flush=True)
            ok_any |= err < 1e-2
        except Exception as e:
            print(f"swap={swap} case={(B, Cin, Cout, T, k, d)} EXC {e}", flush=True)
os.environ["BVG_UMMA_SWAP"] = "0"
sys.exit(0 if ok_any else 3)
