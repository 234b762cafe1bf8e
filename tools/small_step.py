"""A few tiny decodes (B=1, T=24 by default) for an ncu launch list: where does the fixed per-kernel time go?"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np, torch
from b200vgan import synth
from b200vgan.model import BigVGAN
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
sd = synth.make_state_dict(1234, with_speaker_encoder=False)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
g = g.to("cuda"); g.remove_weight_norm(); g.eval()
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
x = torch.from_numpy(synth.make_latents(2, 0, a.batch, a.frames)).cuda()
for _ in range(a.iters):
    y = g.forward_with_embedding(x, emb)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = g.forward_with_embedding(x, emb); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"decode B={a.batch} T={a.frames}: median {sorted(ts)[5]:.3f} ms")
