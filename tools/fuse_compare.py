"""Fused vs unfused activation must give bit-identical waveforms (same math, same rounding points)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
    import numpy as np, torch
    from b200vgan import synth
    from b200vgan.model import BigVGAN
    g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
    sd = synth.make_state_dict(1234, with_speaker_encoder=False)
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    g = g.to("cuda"); g.remove_weight_norm(); g.eval()
    emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
    outs = []
    for (B, T, lens) in ((1, 40, None), (3, 9, [9, 4, 1]), (2, 130, [130, 77])):
        x = torch.from_numpy(synth.make_latents(7, B, B, T)).cuda()
        outs.append(g.forward_with_embedding(x, emb, x_lens=lens).cpu().numpy())
    np.savez(sys.argv[1], *outs)
else:
    import numpy as np
    for flag in ("1", "0"):
        env = dict(os.environ, BVG_FUSE_ACT=flag)
        subprocess.check_call([sys.executable, __file__, f"/tmp/fuse_{flag}.npz"], env=env)
    a, b = np.load("/tmp/fuse_1.npz"), np.load("/tmp/fuse_0.npz")
    ok = True
    for k in a.files:
        d = np.abs(a[k] - b[k]).max()
        print(k, a[k].shape, "max |fused - unfused| =", d, "finite", np.isfinite(a[k]).all())
        ok &= d == 0.0
    print("IDENTICAL" if ok else "DIFFERENT")
    sys.exit(0 if ok else 1)
