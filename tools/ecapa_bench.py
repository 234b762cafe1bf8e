"""Times the native ECAPA-TDNN speaker encoder (bvg_speaker_embedding) on a prompt-sized mel [B, Tm, 100]."""
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
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--tm", type=int, default=511)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--flush", type=int, default=1)
a = ap.parse_args()
g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
sd = synth.make_state_dict(1234, with_speaker_encoder=True)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
g = g.to("cuda")
g.remove_weight_norm()
g.eval()
mel = torch.from_numpy(synth.make_mel(seed=7, Tm=a.tm, B=a.batch)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
emb = g.speaker_embedding(mel)
torch.cuda.synchronize()
ts = []
for i in range(a.iters):
    if a.flush:
        flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    emb = g.speaker_embedding(mel)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"ecapa B={a.batch} Tm={a.tm}: median {sorted(ts)[len(ts) // 2]:.3f} ms  min {min(ts):.3f} ms  (|emb| max {float(emb.abs().max()):.3f})")
