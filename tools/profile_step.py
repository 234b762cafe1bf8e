"""Short single-GPU run for ncu: a few config-2 decodes (B=16 x 10 s, bf16 unless --precision fp32)."""
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
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=235)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
g = BigVGAN(dict(synth.H_DEFAULT), precision=a.precision)
sd = synth.make_state_dict(1234, with_speaker_encoder=False)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
g = g.to("cuda")
g.remove_weight_norm()
g.eval()
x = torch.from_numpy(synth.make_latents(2, 0, a.batch, a.frames)).cuda()
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
for _ in range(a.iters):
    wav = g.forward_with_embedding(x, emb)
torch.cuda.synchronize()
print("ok", float(wav.abs().max()))
