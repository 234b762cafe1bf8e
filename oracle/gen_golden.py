"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference is imported from where it lies (nothing is copied); the only shims are a dummy
`matplotlib` module (BigVGAN/utils.py:7-13 imports it at module scope, it is absent here) and
an attribute-dict standing in for OmegaConf.  Weights come from
`b200vgan.synth` (deterministic, regenerated identically by the tests), loaded into the
reference module with `load_state_dict`; outputs are what the reference's own torch code
computes in fp32 (and fp64 for the full forward, to show the fp32 noise floor).
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("BVG_REF_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    pl = types.ModuleType("matplotlib.pylab")
    mpl.pylab = pl
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pylab"] = pl
    sys.path.insert(0, REF_ROOT)
    import indextts.BigVGAN.models as models
    return models


class H(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def ref_generator(models, sd_np, dtype=torch.float32):
    h = H(yaml.safe_load(open(f"{REF_ROOT}/checkpoints/config.yaml"))["bigvgan"])
    g = models.BigVGAN(h, use_cuda_kernel=False)
    g.remove_weight_norm()
    g.eval()
    missing = g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)
    print("load_state_dict:", missing)
    return g.to(dtype)


def prompt_mel():
    """Config 1's speaker input: the mel of the reference's own tests/sample_prompt.wav with the semantics of
    indextts/infer.py:509-514 (mean over channels, torchaudio Resample -> 24 kHz, MelSpectrogramFeatures).
    torchaudio.load needs torchcodec here, so the wav is read with scipy and scaled like torchaudio does
    for int16 PCM (/ 32768).  Returns (audio24k [1,N] fp32, mel [1,Tm,100] fp32, already transposed the way
    infer.py:458 passes it: auto_conditioning.transpose(1, 2))."""
    import scipy.io.wavfile as wavfile
    import torchaudio
    sys.path.insert(0, REF_ROOT)
    from indextts.utils.feature_extractors import MelSpectrogramFeatures
    sr, a = wavfile.read(os.path.join(REF_ROOT, "tests", "sample_prompt.wav"))
    audio = torch.from_numpy(a.astype(np.float32) / 32768.0).t()
    audio = torch.mean(audio, dim=0, keepdim=True)
    audio = torchaudio.transforms.Resample(sr, 24000)(audio)
    mel = MelSpectrogramFeatures()(audio)                     # [1, 100, Tm]
    return audio.numpy(), mel.transpose(1, 2).contiguous().numpy()


def merge_cases():
    """Random subtitle timelines for the timeline-merge fixture: out-of-order entries, overlaps (3-way too), an empty
    segment, segments running past the nominal end (array growth), loud sums (peak normalisation)."""
    rng = np.random.default_rng(77)
    cases = {}
    for name, n, sr, span, amp in (("a", 6, 2400, 3.0, 0.3), ("b", 14, 2400, 4.0, 0.9), ("c", 3, 16000, 0.5, 0.5), ("d", 9, 24000, 0.3, 1.2)):
        st = np.sort(rng.uniform(0.0, span, n))
        rng.shuffle(st)
        ln = rng.integers(50, int((0.6 if sr < 10000 else 0.15) * sr), n)
        if name == "b":
            ln[3] = 0                                 # empty segment
            st[5] = st[4] = st[6]                     # identical start times: 3-way overlap
        en = st + rng.uniform(0.1, 0.5, n)
        audio = [(amp * rng.standard_normal(int(l))).astype(np.float32) for l in ln]
        cases[name] = (sr, st, en, audio)
    return cases


def reference_audio_processor():
    """The unmodified srt_dubbing AudioProcessor; its logging / file-IO dependencies that this image lacks are stubbed."""
    import types
    for mod in ("colorama", "soundfile", "tqdm"):
        try:
            __import__(mod)
        except ImportError:
            m = types.ModuleType(mod)
            if mod == "colorama":
                class _C:
                    def __getattr__(self, k):
                        return ""
                m.Fore, m.Back, m.Style = _C(), _C(), _C()
                m.init = lambda *a, **k: None
            sys.modules[mod] = m
    # import the module file itself, not the package __init__ (which pulls the CLI and its text-processing dependencies)
    for name, path in (("srt_dubbing", os.path.join(REF_ROOT, "srt_dubbing")), ("srt_dubbing.src", os.path.join(REF_ROOT, "srt_dubbing", "src"))):
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    import importlib
    return importlib.import_module("srt_dubbing.src.audio_processor").AudioProcessor


def main():
    """`python oracle/gen_golden.py` regenerates everything; `python oracle/gen_golden.py cfg1 prompt` only the
    named groups (taps, act, amp, ecapa, tiny, cfg1, prompt, logmel, merge).  A full run reproduces the committed files
    (one torch seed, fixed order); a partial run only touches the named ones (their inputs do not depend on
    the torch random stream)."""
    from b200vgan import synth
    only = sys.argv[1:]
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    models = import_reference()
    from indextts.BigVGAN.alias_free_torch import Activation1d
    from indextts.BigVGAN.alias_free_torch.filter import kaiser_sinc_filter1d
    from indextts.BigVGAN import activations
    os.makedirs(GOLD, exist_ok=True)

    def want(name):
        return not only or name in only

    # 0. timeline merge of the dubbing tool (srt_dubbing/src/audio_processor.py:133-230) --------
    if only and "merge" in only or not only:
        AP = reference_audio_processor()
        out = {}
        for name, (sr, st, en, audio) in merge_cases().items():
            segs = [{"index": i + 1, "start_time": float(st[i]), "end_time": float(en[i]), "audio_data": audio[i]} for i in range(len(st))]
            ap = AP(sample_rate=sr)
            out[f"{name}_sr"] = np.int64(sr)
            out[f"{name}_start"], out[f"{name}_end"] = st, en
            out[f"{name}_len"] = np.array([len(a) for a in audio], dtype=np.int64)
            out[f"{name}_audio"] = np.concatenate(audio) if audio else np.zeros(0, np.float32)
            for flag in (False, True):
                out[f"{name}_merged_trunc{int(flag)}"] = ap._time_synchronized_merge([dict(s) for s in segs], flag, False)
            out[f"{name}_natural"] = ap._natural_concatenation([dict(s) for s in segs], False)
        np.savez_compressed(os.path.join(GOLD, "srt_merge.npz"), **out)
        if only == ["merge"]:
            return

    # 1. filter taps ------------------------------------------------------------------
    if want("taps"):
        taps = kaiser_sinc_filter1d(0.25, 0.3, 12).reshape(-1).numpy()
        np.savez(os.path.join(GOLD, "kaiser_taps.npz"), taps=taps)

    # 2. Activation1d(SnakeBeta, logscale) ---------------------------------------------
    if want("act"):
        rng = np.random.default_rng(11)
        cases = {}
        for name, (B, C, L) in {"a": (2, 16, 50), "b": (1, 8, 1), "c": (1, 24, 7), "d": (3, 8, 301)}.items():
            x = (1.5 * rng.standard_normal((B, C, L))).astype(np.float32)
            la = (0.5 * rng.standard_normal(C)).astype(np.float32)
            lb = (0.5 * rng.standard_normal(C)).astype(np.float32)
            act = Activation1d(activation=activations.SnakeBeta(C, alpha_logscale=True))
            act.act.alpha.data = torch.from_numpy(la)
            act.act.beta.data = torch.from_numpy(lb)
            y = act(torch.from_numpy(x)).numpy()
            cases.update({f"{name}_x": x, f"{name}_alpha": la, f"{name}_beta": lb, f"{name}_y": y})
        np.savez(os.path.join(GOLD, "activation1d.npz"), **cases)

    # 3. AMPBlock1 at a small width (reference class, own synthetic weights; uses the torch random stream,
    #    so it is only reproducible in a full run) ---------------------------------------------
    if want("amp"):
        h = H(yaml.safe_load(open(f"{REF_ROOT}/checkpoints/config.yaml"))["bigvgan"])
        h["use_cuda_kernel"] = False
        amp = {}
        for ks in (3, 7, 11):
            C, L = 16, 160
            blk = models.AMPBlock1(h, C, ks, (1, 3, 5), activation="snakebeta")
            blk.remove_weight_norm()
            blk.eval()
            sd = {k: ((0.8 / (C * ks) ** 0.5) * torch.randn_like(v) if v.ndim == 3 and v.shape[-1] != 12 else
                      (0.5 * torch.randn_like(v) if "act." in k else
                       (0.1 * torch.randn_like(v) if k.endswith("bias") else v)))
                  for k, v in blk.state_dict().items()}
            blk.load_state_dict(sd)
            x = torch.randn(2, C, L)
            y = blk(x)
            amp.update({f"k{ks}.{k}": v.numpy() for k, v in sd.items()})
            amp[f"k{ks}.x"] = x.numpy()
            amp[f"k{ks}.y"] = y.numpy()
        np.savez(os.path.join(GOLD, "ampblock1.npz"), **amp)

    # 4. full generator + ECAPA with the synthetic weights --------------------------------
    sd_np = synth.make_state_dict(seed=1234)
    g32 = g64 = None
    if want("ecapa") or want("tiny") or want("cfg1") or want("logmel"):
        g32 = ref_generator(models, sd_np)
    if want("tiny") or want("cfg1"):
        g64 = ref_generator(models, sd_np, torch.float64)
    # 4a. ECAPA on two prompts (one with relative lengths)
    if want("ecapa"):
        mel = synth.make_mel(seed=7, Tm=60, B=2)
        emb = g32.speaker_encoder(torch.from_numpy(mel)).numpy()
        lens = np.array([1.0, 0.6], dtype=np.float32)
        emb_l = g32.speaker_encoder(torch.from_numpy(mel), torch.from_numpy(lens)).numpy()
        np.savez(os.path.join(GOLD, "ecapa.npz"), mel=mel, emb=emb, lens=lens, emb_lens=emb_l)

    # 4b. tiny full forward: B=2 (shared prompt, B'=1), T=6
    if want("tiny"):
        x = synth.make_latents(0, 0, 2, 6)
        mel1 = synth.make_mel(seed=7, Tm=60, B=1)
        wav = g32(torch.from_numpy(x), torch.from_numpy(mel1))[0].numpy()
        wav64 = g64(torch.from_numpy(x).double(), torch.from_numpy(mel1).double())[0].numpy()
        emb1 = g32.speaker_encoder(torch.from_numpy(mel1)).numpy()
        print("tiny: amp", np.abs(wav).max(), "std", wav.std(), "fp32-vs-fp64 maxabs", np.abs(wav - wav64).max())
        np.savez(os.path.join(GOLD, "forward_tiny.npz"), x=x, mel=mel1, emb=emb1, wav=wav,
                 wav64=wav64.astype(np.float32))

    # 4c. the prompt fixture: tests/sample_prompt.wav -> 24 kHz audio and its log-mel [1,511,100] through the
    #     reference's own front-end (BASELINE.json config 1: "ECAPA embedding from tests/sample_prompt.wav");
    #     the audio is stored in fp32, exactly what the front-end saw.
    audio24 = melp = None
    if want("prompt") or want("cfg1"):
        audio24, melp = prompt_mel()
        print("prompt: audio", audio24.shape, "mel", melp.shape, "mean", melp.mean())
    if want("prompt"):
        np.savez_compressed(os.path.join(GOLD, "prompt.npz"), audio=audio24.astype(np.float32), mel=melp.astype(np.float32))

    # 4d. config 1 of BASELINE.json: B=1, T=118 (~5 s), speaker embedding from the prompt wav's mel
    if want("cfg1"):
        x = synth.make_latents(1, 0, 1, 118)
        mel4 = torch.from_numpy(melp)
        emb4 = g32.speaker_encoder(mel4).numpy()
        wav = g32(torch.from_numpy(x), mel4)[0].numpy()
        wav64 = g64(torch.from_numpy(x).double(), mel4.double())[0].numpy()
        print("cfg1: amp", np.abs(wav).max(), "std", wav.std(), "fp32-vs-fp64 maxabs", np.abs(wav - wav64).max())
        np.savez_compressed(os.path.join(GOLD, "forward_cfg1.npz"), emb=emb4, wav=wav.astype(np.float32),
                            fp32_noise=np.float32(np.abs(wav - wav64).max()))

    # 5. log-mel of the reference front-end (torchaudio) on one second of a generated waveform -----------
    if want("logmel"):
        sys.path.insert(0, REF_ROOT)
        from indextts.utils.feature_extractors import MelSpectrogramFeatures
        x = synth.make_latents(1, 0, 1, 118)
        wav = g32(torch.from_numpy(x), torch.from_numpy(synth.make_mel(seed=7, Tm=400, B=1)))[0].numpy()
        fe = MelSpectrogramFeatures()
        m = fe(torch.from_numpy(wav[:, 0, :24000])).numpy()
        np.savez_compressed(os.path.join(GOLD, "logmel.npz"), wav=wav[:, 0, :24000], mel=m)


if __name__ == "__main__":
    main()
