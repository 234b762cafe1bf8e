"""Does replaying the decode as a CUDA graph remove the inter-kernel gaps?  cfg2, eager vs graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np, torch
from b200vgan import synth
from b200vgan.model import BigVGAN
g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
sd = synth.make_state_dict(1234, with_speaker_encoder=False)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
g = g.to("cuda"); g.remove_weight_norm(); g.eval()
B, T = 16, 235
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
x = torch.from_numpy(synth.make_latents(2, 0, B, T)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g.forward_with_embedding(x, emb); torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    y = g.forward_with_embedding(x, emb)
def timeit(fn, n=8, warm=3):
    ts = []
    for i in range(n + warm):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= warm: ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
print("eager %.3f ms   graph %.3f ms" % (timeit(lambda: g.forward_with_embedding(x, emb)), timeit(graph.replay)))
