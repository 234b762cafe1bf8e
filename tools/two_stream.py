"""Experiment: does running two independent decodes on two CUDA streams overlap the FP32-bound activation
kernels of one with the tensor-bound convolutions of the other?  Compares 1 stream x B=16 against
2 streams x B=8 (same total work) and 2 streams x B=16."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np, torch
from b200vgan import synth
from b200vgan.model import BigVGAN

def make():
    g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
    sd = synth.make_state_dict(1234, with_speaker_encoder=False)
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    g = g.to("cuda"); g.remove_weight_norm(); g.eval()
    return g

T = 235
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
gs = [make(), make()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(nstream, B, steps=6, warm=3, delay=0):
    xs = [torch.from_numpy(synth.make_latents(2, i, B, T)).cuda() for i in range(nstream)]
    times = []
    for it in range(warm + steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nstream):
            streams[i].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(streams[i]):
                if i == 1 and delay:
                    torch.cuda._sleep(int(delay))   # phase offset: B's activations against A's convolutions
                gs[i].forward_with_embedding(xs[i], emb)
        for i in range(nstream):
            torch.cuda.current_stream().wait_stream(streams[i])
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    audio = nstream * B * T * 1024 / 24000
    print(f"{nstream} stream(s) x B={B} delay {delay}: {ms:.2f} ms  -> {audio / ms * 1e3:.0f} audio-s/s", flush=True)

run(1, 16); run(2, 8)
for d in (60000, 120000, 200000, 400000):
    run(2, 8, delay=d)
run(2, 16); run(2, 16, delay=150000)
