"""-m gpu: the GPU log-mel front-end (bvg_log_mel, csrc/bvg_mel.cu) against the reference's MelSpectrogramFeatures
outputs committed under tests/golden/ (oracle/gen_golden.py: logmel.npz from a decoded waveform, prompt.npz from the
reference's tests/sample_prompt.wav) and against the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu
MEL_ATOL = 2e-4    # log-mel, fp32 FFT vs the reference's (same bound as the oracle's own pin, tests/test_oracle_golden.py)


def _mel(audio, **kw):
    from b200vgan.features import MelSpectrogramFeatures
    out = MelSpectrogramFeatures()(torch.as_tensor(audio).cuda(), **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_log_mel_matches_reference_frontend(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    mel = _mel(g["wav"])
    assert mel.shape == g["mel"].shape == (1, 100, 94)
    err = np.abs(mel - g["mel"]).max()
    print("GPU log-mel vs reference (1 s decoded audio): max-abs", err)
    assert err <= MEL_ATOL


def test_prompt_mel_matches_reference_frontend(golden_dir):
    """cond_mel of config 1: the reference's tests/sample_prompt.wav resampled to 24 kHz -> [1,511,100] (infer.py:509-514)."""
    pr = np.load(os.path.join(golden_dir, "prompt.npz"))
    mel_t = _mel(pr["audio"], transposed=True)
    assert mel_t.shape == pr["mel"].shape == (1, 511, 100)
    # quiet bins carry the fp32 noise of BOTH STFTs (the fp64 oracle itself is 2.8e-4 from the reference here):
    # the same gate as the oracle's pin in tests/test_oracle_golden.py, plus the oracle as the tight check
    err = np.abs(mel_t - pr["mel"])
    ref64 = O.log_mel(pr["audio"]).transpose(0, 2, 1)
    err64 = np.abs(mel_t - ref64)
    print("GPU prompt mel vs reference: max-abs", err.max(), "mean", err.mean(), "| vs fp64 oracle: max-abs", err64.max())
    assert err.max() <= 1e-3 and err.mean() <= 2e-5
    assert err64.max() <= 1e-3 and err64.mean() <= 2e-5
    mel = _mel(pr["audio"])
    np.testing.assert_array_equal(mel.transpose(0, 2, 1), mel_t)


@pytest.mark.parametrize("B,N", [(1, 513), (3, 4000), (2, 24000 * 3 + 17)])
def test_log_mel_matches_oracle(B, N):
    rng = np.random.default_rng(N)
    wav = (0.3 * rng.standard_normal((B, N))).astype(np.float32)
    ref = O.log_mel(wav)
    mel = _mel(wav)
    assert mel.shape == ref.shape == (B, 100, 1 + N // 256)
    err = np.abs(mel - ref).max()
    print("GPU log-mel vs oracle", (B, N), err)
    assert err <= MEL_ATOL


def test_prompt_audio_to_waveform_stays_on_device(synth_sd, golden_dir):
    """audio -> GPU mel -> native ECAPA -> decode, against the reference's cfg1 waveform (which was made from the
    reference front-end's mel of the same audio)."""
    from b200vgan import synth
    from b200vgan.features import MelSpectrogramFeatures
    from b200vgan.model import BigVGAN
    g = BigVGAN(dict(synth.H_DEFAULT), precision="fp32")
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth_sd.items()})
    g = g.to("cuda"); g.remove_weight_norm(); g.eval()
    pr = np.load(os.path.join(golden_dir, "prompt.npz"))
    gold = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    cond_mel = MelSpectrogramFeatures()(torch.as_tensor(pr["audio"]).cuda())          # [1, 100, 511] like the reference
    wav, _ = g(torch.as_tensor(synth.make_latents(1, 0, 1, 118)).cuda(), cond_mel.transpose(1, 2))
    err = float(np.abs(wav.cpu().numpy() - gold["wav"]).max())
    print("audio -> mel -> ECAPA -> decode vs reference waveform: max-abs", err)
    assert err <= 1e-3


def test_log_mel_rejects_unsupported_configs():
    from b200vgan import lib
    from b200vgan.features import MelSpectrogramFeatures
    with pytest.raises(lib.BvgError):
        MelSpectrogramFeatures(n_fft=2048)
    with pytest.raises(lib.BvgError):
        MelSpectrogramFeatures()(torch.zeros(1, 100, device="cuda"))       # too short for reflect padding
    with pytest.raises(lib.BvgError):
        MelSpectrogramFeatures()(torch.zeros(1, 4000))                     # CPU tensor: no CPU path
