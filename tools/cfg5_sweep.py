"""Config 5 of BASELINE.json on one GPU: decode throughput for batch 1..64 x 2..30 s (device-resident latents,
CUDA-event time, 256 MiB L2 flush between timed steps), written as a markdown table.

    python tools/cfg5_sweep.py [--precision bf16|fp16|fp32] > profiles/r2_cfg5_sweep.md
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np
import torch

from b200vgan import synth
from b200vgan.model import BigVGAN

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--max-audio-s", type=float, default=1000.0, help="skip cells whose batch holds more audio than this")
a = ap.parse_args()
g = BigVGAN(dict(synth.H_DEFAULT), precision=a.precision)
sd = synth.make_state_dict(1234, with_speaker_encoder=False)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
g = g.to("cuda"); g.remove_weight_norm(); g.eval()
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
batches, seconds = (1, 2, 4, 8, 16, 32, 64), (2, 5, 10, 20, 30)
print(f"cfg5 sweep, one B200, precision mode {a.precision}: audio-seconds per second (ms per decode); "
      f"device-resident latents, CUDA events, {a.steps} timed steps after 2 warm-ups, 256 MiB L2 flush between steps\n")
print("| batch \\ seconds | " + " | ".join(f"{s} s" for s in seconds) + " |")
print("|---|" + "---|" * len(seconds))
for B in batches:
    cells = []
    for sec in seconds:
        T = synth.frames_for_seconds(sec)
        audio = B * T * 1024 / 24000.0
        if audio > a.max_audio_s:
            cells.append("-")
            continue
        x = torch.randn(B, T, 1024, device="cuda")
        for _ in range(2):
            g.forward_with_embedding(x, emb)
        torch.cuda.synchronize()
        ms = 0.0
        for i in range(a.steps):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.forward_with_embedding(x, emb); e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= a.steps
        cells.append(f"{audio / (ms * 1e-3):.0f} ({ms:.2f})")
        del x
    print(f"| {B} | " + " | ".join(cells) + " |")
    sys.stdout.flush()
