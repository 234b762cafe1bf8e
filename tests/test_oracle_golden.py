"""Pins the numpy oracle (oracle/bigvgan_oracle.py) to outputs of the UNMODIFIED reference
(fixtures written by oracle/gen_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest

from oracle import bigvgan_oracle as O


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_kaiser_taps(golden_dir):
    from b200vgan import synth
    taps = _g(golden_dir, "kaiser_taps.npz")["taps"]
    assert taps.shape == (12,)
    np.testing.assert_allclose(O.kaiser_sinc_filter1d(0.25, 0.3, 12), taps, atol=1e-7)
    np.testing.assert_allclose(synth.kaiser_filter(), taps, atol=1e-7)
    np.testing.assert_allclose(synth.KAISER_TAPS, taps, atol=1e-9)
    np.testing.assert_allclose(taps, taps[::-1], atol=0)        # symmetric
    assert abs(float(taps.astype(np.float64).sum()) - 1.0) < 1e-6


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_activation1d_matches_reference(golden_dir, case):
    g = _g(golden_dir, "activation1d.npz")
    x, la, lb = (g[f"{case}_{k}"].astype(np.float64) for k in ("x", "alpha", "beta"))
    y = O.activation1d(x, la, lb)
    np.testing.assert_allclose(y, g[f"{case}_y"], atol=5e-6)
    # the closed form the CUDA kernels implement is the same function
    np.testing.assert_allclose(O.activation1d_closed_form(x, la, lb), y, atol=1e-12)


@pytest.mark.parametrize("ks", [3, 7, 11])
def test_ampblock1_matches_reference(golden_dir, ks):
    g = _g(golden_dir, "ampblock1.npz")
    pre = f"k{ks}."
    sd = {"p." + k[len(pre):]: g[k].astype(np.float64) for k in g.files if k.startswith(pre)}
    y = O.amp_block1(sd["p.x"], sd, "p", ks)
    np.testing.assert_allclose(y, sd["p.y"], atol=2e-5)


def test_ecapa_matches_reference(golden_dir, synth_sd):
    g = _g(golden_dir, "ecapa.npz")
    np.testing.assert_allclose(O.ecapa_forward(g["mel"], synth_sd), g["emb"], atol=2e-5)
    np.testing.assert_allclose(O.ecapa_forward(g["mel"], synth_sd, lengths=g["lens"]), g["emb_lens"], atol=2e-5)


def test_forward_tiny_matches_reference(golden_dir, synth_sd):
    g = _g(golden_dir, "forward_tiny.npz")
    wav = O.bigvgan_forward(g["x"], g["mel"], synth_sd)
    assert wav.shape == g["wav"].shape == (2, 1, 6 * 1024)
    assert np.abs(wav - g["wav64"]).max() < 1e-6      # reference run in fp64
    assert np.abs(wav - g["wav"]).max() < 5e-5        # reference run in fp32 (its own rounding noise)
    assert O.mel_l1(wav[:, 0], g["wav"][:, 0]) < 1e-3


def test_weight_norm_fold(synth_sd):
    from b200vgan import synth
    wn = synth.make_state_dict(seed=1234, weight_norm=True, with_speaker_encoder=False)
    folded = O.fold_state_dict(wn)
    assert "conv_pre.weight_g" in wn and wn["ups.0.0.weight_g"].shape == (1536, 1, 1)
    for k in ("conv_pre.weight", "ups.0.0.weight", "resblocks.7.convs2.1.weight", "conv_post.weight"):
        np.testing.assert_allclose(folded[k], synth_sd[k], rtol=2e-6, atol=1e-8)


def test_log_mel_matches_reference_frontend(golden_dir):
    g = _g(golden_dir, "logmel.npz")
    np.testing.assert_allclose(O.log_mel(g["wav"]), g["mel"], atol=2e-4)


def test_conv_transpose_closed_form():
    """SURVEY 8a closed form used by the kernels: u interleaved phase convolutions."""
    rng = np.random.default_rng(3)
    for (k, u) in ((8, 4), (4, 4), (4, 2)):
        p = (k - u) // 2
        x = rng.standard_normal((1, 8, 9))
        w = rng.standard_normal((8, 16, k))
        y = O.conv_transpose1d(x, w, None, u, p)
        L = x.shape[-1]
        xp = np.concatenate([x, np.zeros((1, 8, 1))], axis=-1)
        y2 = np.zeros_like(y)
        for q in range(L + 1):
            for phi in range(u):
                n = q * u + phi - p
                if 0 <= n < L * u:
                    for m in range(k // u):
                        if q - m >= 0:
                            y2[0, :, n] += w[:, :, phi + m * u].T @ xp[0, :, q - m]
        np.testing.assert_allclose(y2, y, atol=1e-12)


def test_torch_cpu_port_matches_reference_golden(golden_dir):
    """The CPU baseline port (what bench.py's reference arm times) reproduces the unmodified reference."""
    import torch
    from oracle import bigvgan_torch_cpu as TC
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_tiny.npz"))
    sd = synth.make_state_dict(1234, with_speaker_encoder=False)
    y32 = TC.bigvgan_forward_with_embedding(g["x"], g["emb"], TC.prepare_state_dict(sd))
    np.testing.assert_allclose(y32, g["wav"], atol=2e-5)
    y64 = TC.bigvgan_forward_with_embedding(g["x"], g["emb"], TC.prepare_state_dict(sd, torch.float64), dtype=torch.float64)
    assert np.abs(y64 - g["wav64"]).max() < 2e-6      # the stored embedding is the reference's fp32 one
    yo = O.bigvgan_forward_with_embedding(g["x"], g["emb"], sd)
    assert np.abs(y64 - yo).max() < 1e-12            # torch port == numpy oracle on identical inputs


def test_cfg1_prompt_fixture_pins_oracle(golden_dir, synth_sd):
    """Config 1 of BASELINE.json: the speaker embedding comes from the reference's tests/sample_prompt.wav.
    prompt.npz holds the 24 kHz audio and the reference front-end's log-mel [1,511,100]; forward_cfg1.npz the
    reference's ECAPA embedding of that mel and its fp32 waveform.  Pins: oracle log-mel, oracle ECAPA and the
    torch-CPU port of the generator."""
    import torch
    from oracle import bigvgan_torch_cpu as TC
    from b200vgan import synth
    pr = np.load(os.path.join(golden_dir, "prompt.npz"))
    g = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    assert pr["mel"].shape == (1, 511, 100) and pr["audio"].shape == (1, 130560)
    mel = O.log_mel(pr["audio"])                       # [1, 100, 511]
    dm = np.abs(mel.transpose(0, 2, 1) - pr["mel"])     # the reference's STFT runs in fp32: quiet bins carry its noise
    assert dm.max() <= 1e-3 and dm.mean() <= 1e-5
    emb = O.ecapa_forward(pr["mel"], synth_sd)
    assert np.abs(emb - g["emb"]).max() <= 2e-5 * max(1.0, np.abs(g["emb"]).max())
    sd = synth.make_state_dict(1234, with_speaker_encoder=False)
    x = synth.make_latents(1, 0, 1, 118)
    y32 = TC.bigvgan_forward_with_embedding(x, g["emb"], TC.prepare_state_dict(sd))
    assert np.abs(y32 - g["wav"]).max() <= 5e-5        # fp32 summation-order noise (reference's own: fp32_noise)
    assert O.mel_l1(y32[:, 0], g["wav"][:, 0]) <= 1e-4


def _merge_cases(golden_dir):
    g = _g(golden_dir, "srt_merge.npz")
    for name in ("a", "b", "c", "d"):
        ln = g[f"{name}_len"]
        off = np.concatenate([[0], np.cumsum(ln)])
        segs = [{"index": i + 1, "start_time": float(g[f"{name}_start"][i]), "end_time": float(g[f"{name}_end"][i]),
                 "audio_data": g[f"{name}_audio"][off[i]:off[i + 1]]} for i in range(len(ln))]
        yield name, int(g[f"{name}_sr"]), segs, g


def test_timeline_merge_oracle_matches_reference(golden_dir):
    """oracle/srt_merge_oracle.py against the unmodified srt_dubbing AudioProcessor (audio_processor.py:70-230):
    placement, array growth, the previous-segment overlap rule, fp32 sums and the peak normalisation, bit for bit."""
    from oracle import srt_merge_oracle as M
    for name, sr, segs, g in _merge_cases(golden_dir):
        for flag in (False, True):
            np.testing.assert_array_equal(M.time_synchronized_merge(segs, sr, flag), g[f"{name}_merged_trunc{int(flag)}"])
        np.testing.assert_array_equal(M.natural_concatenation(segs), g[f"{name}_natural"])
