"""Is a small decode launch-bound?  CPU wall time of issuing one forward / one speaker-encoder call (no sync) next to
the device time of the same call, eager vs CUDA-graph replay, for cfg1 (B=1 x 5 s) and a few batch sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np, torch
from b200vgan import synth
from b200vgan.model import BigVGAN
g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
sd = synth.make_state_dict(1234, with_speaker_encoder=True)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
g = g.to("cuda"); g.remove_weight_norm(); g.eval()
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
mel = torch.from_numpy(synth.make_mel(seed=7, Tm=511, B=1)).cuda()

def timeit(fn, n=10, warm=3):
    dev, cpu = [], []
    for i in range(n + warm):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        if i >= warm: dev.append(e0.elapsed_time(e1)); cpu.append((t1 - t0) * 1e3)
    return sorted(dev)[len(dev) // 2], sorted(cpu)[len(cpu) // 2]

d, c = timeit(lambda: g.speaker_embedding(mel))
print(f"speaker_embedding [1,511,100]: device {d:.3f} ms, CPU issue {c:.3f} ms")
for B, T in [(1, 118), (1, 24), (4, 118), (16, 235)]:
    x = torch.from_numpy(synth.make_latents(2, 0, B, T)).cuda()
    g.forward_with_embedding(x, emb); torch.cuda.synchronize()
    d, c = timeit(lambda: g.forward_with_embedding(x, emb))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = g.forward_with_embedding(x, emb)
    dg, cg = timeit(graph.replay)
    print(f"decode B={B} T={T}: eager device {d:.3f} ms (CPU issue {c:.3f} ms, {g.num_launches([T] * B)} launches); graph replay device {dg:.3f} ms (CPU {cg:.3f} ms)")
