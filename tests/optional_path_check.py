"""Run by test_gpu_forward.py::test_optional_kernel_paths in a subprocess (the library reads its environment knobs once):
decodes the tiny golden case in bf16 mode and prints the SNR against the reference's fp32 waveform."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np
import torch
from b200vgan import synth
from b200vgan.model import BigVGAN
from oracle import bigvgan_oracle as O

g = np.load(os.path.join(ROOT, "tests", "golden", "forward_cfg1.npz"))
sd = synth.make_state_dict(1234, with_speaker_encoder=False)
m = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
m = m.to("cuda"); m.remove_weight_norm(); m.eval()
x = torch.from_numpy(synth.make_latents(1, 0, 1, 118)).cuda()
wav = m.forward_with_embedding(x, torch.from_numpy(g["emb"]).cuda()).cpu().numpy()
print("SNR_DB", O.snr_db(g["wav"], wav))
