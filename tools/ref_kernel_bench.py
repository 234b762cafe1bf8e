"""Time bvg_activation1d (our drop-in, [B,C,T] layout) against the reference's own fused kernel
(indextts/BigVGAN/alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-181), built for sm_100a by the
reference's setup.py when it was pip-installed into git-ignored baseline/_ref/ ("bar to beat on the
box = this kernel compiled -arch=sm_100a").  Also reports how far the two differ at the sequence edges (the
reference kernel does not follow the torch path there, SURVEY.md 8a last row) and in the interior.

    python tools/ref_kernel_bench.py            # [16,24,240640] fp32 and bf16  -> stdout (markdown table)
"""
import ctypes as C
import importlib.util
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
import numpy as np
import torch

from b200vgan import lib, synth

so = glob.glob(os.path.join(ROOT, "baseline", "_ref", "indextts", "BigVGAN", "alias_free_activation", "cuda",
                            "anti_alias_activation_cuda*.so"))
if not so:
    raise SystemExit("reference extension not found under baseline/_ref (pip install of /root/reference did not build it)")
spec = importlib.util.spec_from_file_location("anti_alias_activation_cuda", so[0])
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
L = lib.load()


def time_ms(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("| shape | dtype | reference kernel (sm_100a build) ms | bvg_activation1d ms | speed-up | GB/s ours | interior max-abs diff | edge (first/last 3) max-abs diff |")
print("|---|---|---|---|---|---|---|---|")
taps = torch.from_numpy(synth.kaiser_filter()).float().cuda().view(1, 1, 12)
for shape in ((16, 24, 240640), (16, 96, 60160), (16, 768, 940)):
    B, Cc, T = shape
    for dt, code in ((torch.float32, 0), (torch.bfloat16, 1)):
        g = torch.Generator(device="cuda").manual_seed(1)
        x = (0.8 * torch.randn(B, Cc, T, device="cuda", generator=g)).to(dt).contiguous()
        la = (0.5 * torch.randn(Cc, device="cuda", generator=g)).float()
        lb = (0.5 * torch.randn(Cc, device="cuda", generator=g)).float()
        y = torch.empty_like(x)
        s = torch.cuda.current_stream().cuda_stream
        ours = lambda: lib.check(L.bvg_activation1d(x.data_ptr(), y.data_ptr(), la.data_ptr(), lb.data_ptr(), B, Cc, T, code, s))  # noqa: E731
        theirs = lambda: ref.forward(x, taps, taps, la, lb)  # noqa: E731
        t_ref, t_our = time_ms(theirs), time_ms(ours)
        yr = ref.forward(x, taps, taps, la, lb).float()
        ours()
        torch.cuda.synchronize()
        d = (y.float() - yr).abs()
        edge = max(float(d[..., :3].max()), float(d[..., -3:].max()))
        inner = float(d[..., 3:-3].max())
        gbs = 2 * x.numel() * x.element_size() / (t_our * 1e-3) / 1e9
        print(f"| {list(shape)} | {str(dt).split('.')[-1]} | {t_ref:.3f} | {t_our:.3f} | {t_ref / t_our:.2f}x | {gbs:.0f} | {inner:.2e} | {edge:.2e} |")
