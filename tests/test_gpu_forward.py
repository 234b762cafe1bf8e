"""-m gpu: whole-generator parity against the reference's golden waveforms and the oracle.

Gates (BASELINE.md section 5 / north_star): fp32 mode waveform max-abs <= 1e-3 and mel-L1 <= 1e-3;
bf16 mode a stated SNR bound (BF16_SNR_DB below)."""
import os

import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu

FP32_MAXABS = 1e-3     # north_star gate
FP32_MEL_L1 = 1e-3     # north_star gate
BF16_SNR_DB = 28.5     # stated bound for the bf16 performance mode: within 3 dB of what is measured (31.4-31.9 dB)
FP16_SNR_DB = 40.0     # stated bound for the fp16 performance mode (3 more mantissa bits in every stored tensor)


def _oracle(x, emb, dtype=None):
    """The CPU oracle (torch-operator restatement of the reference path, pinned to the reference's golden
    waveforms in tests/test_oracle_golden.py), one call per batch item so ragged inputs need no padding."""
    from b200vgan import synth
    from oracle import bigvgan_torch_cpu as TC
    import torch as _t
    sd = TC.prepare_state_dict(synth.make_state_dict(1234, with_speaker_encoder=False))
    return TC.bigvgan_forward_with_embedding(np.asarray(x), np.asarray(emb), sd)


def _gate_fp32(tag, wav, ref):
    err = float(np.abs(wav - ref).max())
    mel = O.mel_l1(wav, ref)
    print(tag, "fp32 max-abs", err, "mel-L1", mel)
    assert err <= FP32_MAXABS and mel <= FP32_MEL_L1


@pytest.fixture(scope="module")
def gen(synth_sd):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from b200vgan import synth
    from b200vgan.model import BigVGAN
    g = BigVGAN(dict(synth.H_DEFAULT), use_cuda_kernel=True, precision="fp32")
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth_sd.items()})
    g = g.to("cuda")
    g.remove_weight_norm()
    g.eval()
    return g


def _run(g, x, emb, precision, lens=None):
    g.precision = precision
    wav = g.forward_with_embedding(torch.as_tensor(x).cuda(), torch.as_tensor(emb).cuda(), x_lens=lens)
    torch.cuda.synchronize()
    return wav.cpu().numpy()


def test_tiny_forward_fp32_matches_reference(gen, golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_tiny.npz"))
    wav = _run(gen, g["x"], g["emb"], "fp32")
    assert wav.shape == g["wav"].shape
    err = np.abs(wav - g["wav"]).max()
    print("tiny fp32 max-abs", err)
    assert err <= FP32_MAXABS
    assert O.mel_l1(wav[:, 0], g["wav"][:, 0]) <= FP32_MEL_L1


def test_dropin_call_with_speaker_encoder(gen, golden_dir):
    """The reference call site: wav, _ = bigvgan(latent, mel_ref)  (infer.py:458, :623)."""
    g = np.load(os.path.join(golden_dir, "forward_tiny.npz"))
    gen.precision = "fp32"
    wav, aux = gen(torch.as_tensor(g["x"]).cuda(), torch.as_tensor(g["mel"]).cuda())
    assert aux is None and tuple(wav.shape) == (2, 1, 6 * 1024) and wav.dtype == torch.float32
    assert np.abs(wav.cpu().numpy() - g["wav"]).max() <= FP32_MAXABS


def test_cfg1_fp32_matches_reference(gen, golden_dir):
    """Config 1 of BASELINE.json: B=1, T=118 (~5 s); the speaker embedding is the reference ECAPA's output for the
    mel of the reference's tests/sample_prompt.wav (tests/golden/prompt.npz, oracle/gen_golden.py)."""
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    x = synth.make_latents(1, 0, 1, 118)
    wav = _run(gen, x, g["emb"], "fp32")
    err = np.abs(wav - g["wav"]).max()
    mel = O.mel_l1(wav[:, 0], g["wav"][:, 0])
    print("cfg1 fp32 max-abs", err, "mel-L1", mel, "reference fp32-vs-fp64 noise", float(g["fp32_noise"]))
    assert err <= FP32_MAXABS and mel <= FP32_MEL_L1


def test_cfg1_dropin_call_from_prompt_mel(gen, golden_dir):
    """Config 1 through the reference call site: wav, _ = bigvgan(latent, mel_ref) with the prompt wav's mel
    [1,511,100] -- native ECAPA kernels + decode against the unmodified reference's waveform."""
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    pr = np.load(os.path.join(golden_dir, "prompt.npz"))
    gen.precision = "fp32"
    x = torch.as_tensor(synth.make_latents(1, 0, 1, 118)).cuda()
    emb = gen.speaker_embedding(torch.as_tensor(pr["mel"]).cuda()).cpu().numpy()
    assert np.abs(emb - g["emb"]).max() <= 2e-4
    wav, _ = gen(x, torch.as_tensor(pr["mel"]).cuda())
    _gate_fp32("cfg1 drop-in", wav.cpu().numpy()[0, 0], g["wav"][0, 0])


def test_cfg1_fp16_snr(gen, golden_dir):
    """fp16 storage mode (BVG_MODE_F16): what the reference deploys under torch.amp.autocast(float16), infer.py:456."""
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    x = synth.make_latents(1, 0, 1, 118)
    wav = _run(gen, x, g["emb"], "fp16")
    snr = O.snr_db(g["wav"], wav)
    print("cfg1 fp16 SNR dB", snr, "max-abs", np.abs(wav - g["wav"]).max(), "mel-L1", O.mel_l1(wav[:, 0], g["wav"][:, 0]))
    assert snr >= FP16_SNR_DB


def test_fp16_mode_saturates_instead_of_overflowing(gen):
    """Latents scaled x64 drive intermediate activations past fp16's range: stores saturate at +-65504 (no inf / NaN
    reaches the waveform), and at x8 the mode still tracks the fp32 mode."""
    from b200vgan import synth
    emb = synth.make_speaker_embedding(B=1)
    x = synth.make_latents(5, 0, 1, 40)
    big = _run(gen, 64.0 * x, emb, "fp16")
    assert np.isfinite(big).all() and np.abs(big).max() <= 1.0
    ref8 = _run(gen, 8.0 * x, emb, "fp32")
    got8 = _run(gen, 8.0 * x, emb, "fp16")
    snr = O.snr_db(ref8, got8)
    print("fp16 mode, latents x8: SNR dB vs fp32 mode", snr)
    assert np.isfinite(got8).all() and snr >= 30.0


def test_auto_precision_follows_autocast(gen, golden_dir):
    """precision="auto" (the constructor default): fp32 tensors like the reference module (the fp32 tensor-core mode)
    unless the caller is inside torch.amp.autocast, then the autocast dtype (infer.py:456, :613)."""
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    x = torch.as_tensor(synth.make_latents(1, 0, 1, 118)).cuda()
    emb = torch.as_tensor(g["emb"]).cuda()
    gen.precision = "auto"
    assert gen.resolved_precision() == "fp32tc"
    w32 = gen.forward_with_embedding(x, emb).cpu().numpy()
    with torch.amp.autocast("cuda", dtype=torch.float16):
        assert gen.resolved_precision() == "fp16"
        w16 = gen.forward_with_embedding(x, emb).cpu().numpy()
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        assert gen.resolved_precision() == "bf16"
    gen.precision = "fp32"
    assert np.abs(w32 - g["wav"]).max() <= FP32_MAXABS
    assert w16.dtype == np.float32 and O.snr_db(g["wav"], w16) >= FP16_SNR_DB


def test_cfg1_bf16_snr(gen, golden_dir):
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    x = synth.make_latents(1, 0, 1, 118)
    wav = _run(gen, x, g["emb"], "bf16")
    snr = O.snr_db(g["wav"], wav)
    print("cfg1 bf16 SNR dB", snr, "max-abs", np.abs(wav - g["wav"]).max(), "mel-L1", O.mel_l1(wav[:, 0], g["wav"][:, 0]))
    assert snr >= BF16_SNR_DB


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 0.0)])
def test_variable_length_batch_equals_single_decodes(gen, precision, tol):
    """Config 3 semantics: every segment of a ragged batch is decoded as if alone (exact per-segment
    lengths; the reference would need zero padding, which changes the last ~1 s -- SURVEY.md section 7)."""
    from b200vgan import synth
    lens = [7, 3, 5, 1]
    x = synth.make_latents(3, 0, 4, 7)
    emb = synth.make_speaker_embedding(B=1)
    batch = _run(gen, x, emb, precision, lens=lens)
    for b, n in enumerate(lens):
        single = _run(gen, x[b:b + 1, :n], emb, precision)
        got = batch[b, 0, : n * 1024]
        if tol == 0.0:
            np.testing.assert_array_equal(got, single[0, 0])       # same kernels, same summation order
        else:
            assert np.abs(got - single[0, 0]).max() <= tol
        assert not batch[b, 0, n * 1024:].any()                    # tail beyond the segment is zero


def test_per_item_speaker_embedding(gen):
    from b200vgan import synth
    x = synth.make_latents(3, 1, 2, 5)
    emb2 = synth.make_speaker_embedding(B=2)
    both = _run(gen, x, emb2, "fp32")
    for b in range(2):
        one = _run(gen, x[b:b + 1], emb2[b:b + 1], "fp32")
        assert np.abs(both[b] - one[0]).max() <= 2e-5


def test_tiny_forward_oracle_other_seed(gen):
    """Oracle (not golden) parity on fresh inputs: ragged batch, per-item embeddings."""
    from b200vgan import synth
    sd = synth.make_state_dict(1234, with_speaker_encoder=False)
    x = synth.make_latents(9, 3, 2, 4)
    emb = synth.make_speaker_embedding(seed=3, B=2)
    ref = O.bigvgan_forward_with_embedding(x, emb, sd)
    wav = _run(gen, x, emb, "fp32")
    assert np.abs(wav - ref).max() <= FP32_MAXABS
    assert O.mel_l1(wav[:, 0], ref[:, 0]) <= FP32_MEL_L1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_long_form_shift_equivariance(gen, precision):
    """Config 4 (60 s, T=1407): size-independent property.  The generator is fully convolutional, so
    far from the sequence ends (receptive field ~34 frames) decoding a window of the latents must
    reproduce the same samples; this exercises tile-halo vs sequence-edge handling at full size."""
    from b200vgan import synth
    T = 1407
    x = synth.make_latents(4, 0, 1, T)
    emb = synth.make_speaker_embedding(B=1)
    full = _run(gen, x, emb, precision)[0, 0]
    assert full.shape[0] == T * 1024 and np.isfinite(full).all()
    lo, hi, margin = 500, 900, 48
    win = _run(gen, x[:, lo:hi], emb, precision)[0, 0]
    a = full[(lo + margin) * 1024:(hi - margin) * 1024]
    b = win[margin * 1024:(hi - lo - margin) * 1024]
    err = np.abs(a - b).max()
    print(precision, "shift-equivariance max-abs", err)
    assert err <= (2e-5 if precision == "fp32" else 2e-2)


def test_cfg2_bf16_vs_fp32_mode(gen):
    """Config 2 shape (B=16 x 10 s): the bf16 tensor-core path against the fp32 CUDA-core path."""
    from b200vgan import synth
    x = synth.make_latents(2, 0, 16, 235)
    emb = synth.make_speaker_embedding(B=1)
    ref = _run(gen, x, emb, "fp32")
    wav = _run(gen, x, emb, "bf16")
    snr = O.snr_db(ref, wav)
    print("cfg2 bf16-vs-fp32-mode SNR dB", snr)
    assert snr >= BF16_SNR_DB


def test_cfg2_two_items_vs_oracle(gen):
    """Config 2 (B=16 x 10 s, T=235): items 3 and 12 of the batch against the CPU oracle decoding them alone --
    fp32 mode within the 1e-3 gates, bf16 mode within its SNR bound."""
    from b200vgan import synth
    x = synth.make_latents(2, 0, 16, 235)
    emb = synth.make_speaker_embedding(B=1)
    w32 = _run(gen, x, emb, "fp32")
    w16 = _run(gen, x, emb, "bf16")
    for i in (3, 12):
        ref = _oracle(x[i:i + 1], emb)[0, 0]
        _gate_fp32(f"cfg2 item {i}", w32[i, 0], ref)
        snr = O.snr_db(ref, w16[i, 0])
        print(f"cfg2 item {i} bf16 SNR dB", snr)
        assert snr >= BF16_SNR_DB


def test_cfg3_srt_segments_vs_oracle(gen):
    """Config 3: "parity checked per segment against the oracle run one segment at a time".
    Nine segments spanning the workload's length range (T = 24 ... 352) are taken out of the batched, ragged SRT
    decode (fp32 and bf16 modes) and compared with the CPU oracle's stand-alone decode of the same latents."""
    from b200vgan import sched, synth
    frames = sched.srt_workload()
    order = sorted(range(len(frames)), key=lambda i: (frames[i], i))
    pick = sorted({order[int(round(q * (len(order) - 1)))] for q in np.linspace(0.0, 1.0, 9)})
    assert frames[order[0]] == min(frames) and frames[order[-1]] == max(frames) and len(pick) >= 8
    # decode them together with their neighbours in the length-sorted order, so the batches are really ragged
    pos = {i: k for k, i in enumerate(order)}
    idx = sorted({order[min(max(pos[i] + d, 0), len(order) - 1)] for i in pick for d in (-1, 0, 1)})
    rng = np.random.default_rng(3)
    lat = {i: rng.standard_normal((frames[i], 1024), dtype=np.float32) for i in idx}
    lat_dev = [torch.from_numpy(lat[i]).cuda() if i in lat else None for i in range(len(frames))]
    emb_np = synth.make_speaker_embedding(B=1)
    emb = torch.from_numpy(emb_np).cuda()
    out = {}
    for precision in ("fp32", "bf16"):
        gen.precision = precision
        out[precision] = sched.decode_segments(gen, lat_dev, emb, indices=idx, max_batch_frames=2048, max_batch=8)
    for i in pick:
        ref = _oracle(lat[i][None], emb_np)[0, 0]
        assert out["fp32"][i].shape[0] == frames[i] * 1024 == ref.shape[0]
        _gate_fp32(f"cfg3 segment {i} (T={frames[i]})", out["fp32"][i].numpy(), ref)
        snr = O.snr_db(ref, out["bf16"][i].numpy())
        print(f"cfg3 segment {i} (T={frames[i]}) bf16 SNR dB", snr)
        assert snr >= BF16_SNR_DB


def test_cfg4_long_form_vs_oracle(gen):
    """Config 4 (60 s, B=1, T=1407): the full waveform against the CPU oracle, with the first and the last 1.5 s
    (where sequence-edge handling differs from tile-halo handling) asserted separately -- SURVEY 8d."""
    from b200vgan import synth
    T = 1407
    x = synth.make_latents(4, 0, 1, T)
    emb = synth.make_speaker_embedding(B=1)
    ref = _oracle(x, emb)[0, 0]
    w32 = _run(gen, x, emb, "fp32")[0, 0]
    assert w32.shape == ref.shape == (T * 1024,)
    edge = 36000                                    # 1.5 s at 24 kHz
    _gate_fp32("cfg4 whole", w32, ref)
    _gate_fp32("cfg4 first 1.5 s", w32[:edge], ref[:edge])
    _gate_fp32("cfg4 last 1.5 s", w32[-edge:], ref[-edge:])
    w16 = _run(gen, x, emb, "bf16")[0, 0]
    for tag, sl in (("whole", slice(None)), ("first 1.5 s", slice(0, edge)), ("last 1.5 s", slice(-edge, None))):
        snr = O.snr_db(ref[sl], w16[sl])
        print("cfg4", tag, "bf16 SNR dB", snr)
        assert snr >= BF16_SNR_DB


def test_cfg3_srt_workload_batched_equals_single(gen):
    """Config 3 of BASELINE.json: 512 variable-length SRT segments (1-15 s), decoded in length-bucketed
    ragged batches (the per-GPU part of the sharded dubbing job).  Spot-checked segments must equal
    their stand-alone decode bit for bit; every waveform must be finite and of the right length."""
    from b200vgan import sched, synth
    frames = sched.srt_workload()                       # 512 segments, ~4096 s of audio
    rng = np.random.default_rng(3)
    lat = [torch.from_numpy(rng.standard_normal((f, 1024), dtype=np.float32)).cuda() for f in frames]
    emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
    gen.precision = "bf16"
    mine = sched.lpt_shards(frames, 8)[3]               # the shard rank 3 of 8 would own
    out = sched.decode_segments(gen, lat, emb, indices=mine, max_batch_frames=4096, max_batch=32)
    assert sorted(out.keys()) == sorted(mine)
    for i in mine:
        assert out[i].shape[0] == frames[i] * 1024 and bool(torch.isfinite(out[i]).all())
    for i in (mine[0], mine[len(mine) // 2], mine[-1]):
        single = gen.forward_with_embedding(lat[i][None], emb)[0, 0].cpu()
        assert torch.equal(out[i], single)
    # the same job with 16-bit PCM straight from the last kernel (what the dubbing tool writes to disk)
    few = mine[:6]
    pcm = sched.decode_segments(gen, lat, emb, indices=few, max_batch_frames=4096, max_batch=32, int16=True)
    for i in few:
        ref = torch.clamp(32767 * out[i], -32767.0, 32767.0).type(torch.int16)
        assert pcm[i].dtype == torch.int16 and torch.equal(pcm[i], ref)


def test_decode_sentences_replaces_infer_fast_chunk_loop(gen, golden_dir):
    """sched.decode_sentences has the shape of the vocoder loop of IndexTTS.infer_fast (infer.py:439-463): a list of
    [1,T_i,1024] sentence latents + the prompt mel -> int16 waveforms in order.  Every sentence must equal its
    stand-alone decode through the reference call `bigvgan(latent, mel)` + the caller's clamp/cast, bit for bit
    (the reference's time-concatenation of chunk_size=2 sentences does not have that property)."""
    from b200vgan import sched
    pr = np.load(os.path.join(golden_dir, "prompt.npz"))
    mel = torch.as_tensor(pr["mel"]).cuda()
    rng = np.random.default_rng(17)
    lens = [31, 7, 118, 1, 64, 19, 64]
    lat = [torch.from_numpy(rng.standard_normal((1, n, 1024), dtype=np.float32)).cuda() for n in lens]
    for precision in ("fp32", "bf16"):
        gen.precision = precision
        before = gen.plans_created()
        wavs = sched.decode_sentences(gen, lat, mel, max_batch_frames=160, max_batch=4)
        assert len(wavs) == len(lens) and gen.plans_created() - before <= 4
        for l, n, w in zip(lat, lens, wavs):
            assert w.dtype == torch.int16 and tuple(w.shape) == (1, n * 1024) and not w.is_cuda
            wav, _ = gen(l, mel)
            ref = torch.clamp(32767 * wav.squeeze(1), -32767.0, 32767.0).type(torch.int16).cpu()
            assert torch.equal(w, ref)


def test_forward_ragged_equals_dense(gen):
    """bvg_forward_ragged (back-to-back rows in, back-to-back samples out) == bvg_forward on the padded batch."""
    from b200vgan import synth
    lens = [9, 2, 5]
    x = torch.as_tensor(synth.make_latents(7, 0, 3, 9)).cuda()
    emb = torch.as_tensor(synth.make_speaker_embedding(B=1)).cuda()
    rows = torch.cat([x[b, :n] for b, n in enumerate(lens)])
    for precision in ("fp32", "bf16"):
        gen.precision = precision
        dense = gen.forward_with_embedding(x, emb, x_lens=lens)
        flat = gen.forward_ragged(rows, lens, emb)
        pcm = gen.forward_ragged(rows, lens, emb, pcm16=True)
        off = 0
        for b, n in enumerate(lens):
            seg = dense[b, 0, : n * 1024]
            assert torch.equal(flat[off:off + n * 1024], seg)
            assert torch.equal(pcm[off:off + n * 1024], torch.clamp(32767 * seg, -32767.0, 32767.0).type(torch.int16))
            off += n * 1024


def test_forward_host_entry_point(gen):
    """bvg_forward_host (pinned host latents in, host waveform out, one call) == the device entry point."""
    import ctypes as C
    from b200vgan import lib, synth
    gen.precision = "bf16"
    x = torch.as_tensor(synth.make_latents(8, 0, 2, 6))
    emb = torch.as_tensor(synth.make_speaker_embedding(B=1)).cuda()
    ref = gen.forward_with_embedding(x.cuda(), emb).cpu()
    plan = gen._plan((6, 6), gen._mode())
    ws = gen._ensure_workspace(int(gen._libh.bvg_plan_workspace_bytes(plan)), emb.device)
    xh = x.pin_memory()
    wav_h = torch.empty(2, 1, 6 * 1024).pin_memory()
    x_dev, wav_dev = torch.empty_like(x, device="cuda"), torch.empty(2, 1, 6 * 1024, device="cuda")
    lib.check(gen._libh.bvg_forward_host(gen._handle, plan, xh.data_ptr(), lib.F32, x_dev.data_ptr(), emb.data_ptr(), 1,
                                         wav_h.data_ptr(), wav_dev.data_ptr(), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream))
    assert torch.equal(wav_h, ref)


def test_checkpoint_layout_gives_same_audio(gen, synth_sd):
    from b200vgan import synth
    from b200vgan.model import BigVGAN
    x = synth.make_latents(5, 0, 1, 4)
    emb = synth.make_speaker_embedding(B=1)
    ref = _run(gen, x, emb, "fp32")
    g2 = BigVGAN(dict(synth.H_DEFAULT), precision="fp32")
    wn = synth.make_state_dict(seed=1234, weight_norm=True)
    g2.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in wn.items()})
    g2 = g2.to("cuda")
    g2.remove_weight_norm()
    g2.eval()
    out = _run(g2, x, emb, "fp32")
    assert np.abs(out - ref).max() <= 1e-5


def test_native_speaker_encoder_matches_reference_golden(gen, golden_dir):
    """bvg_speaker_embedding (csrc/bvg_ecapa.cu) against the unmodified reference's ECAPA outputs,
    without and with relative lengths (ECAPA_TDNN.py:543-581)."""
    e = np.load(os.path.join(golden_dir, "ecapa.npz"))
    emb = gen.speaker_embedding(torch.as_tensor(e["mel"]).cuda()).cpu().numpy()
    assert emb.shape == e["emb"].shape
    err = np.abs(emb - e["emb"]).max()
    emb_l = gen.speaker_embedding(torch.as_tensor(e["mel"]).cuda(), torch.as_tensor(e["lens"])).cpu().numpy()
    err_l = np.abs(emb_l - e["emb_lens"]).max()
    print("native ECAPA max-abs", err, "with lens", err_l, "scale", np.abs(e["emb"]).max())
    assert err <= 2e-4 and err_l <= 2e-4


@pytest.mark.parametrize("B,Tm", [(1, 5), (3, 37), (2, 511)])
def test_native_speaker_encoder_matches_oracle(gen, synth_sd, B, Tm):
    """Short, ragged and prompt-sized mel inputs against the numpy oracle (fp64)."""
    rng = np.random.default_rng(B * 1000 + Tm)
    mel = (2.0 * rng.standard_normal((B, Tm, 100)) - 0.3).astype(np.float32)
    lens = None if B == 1 else np.linspace(0.45, 1.0, B).astype(np.float32)
    ref = O.ecapa_forward(mel, synth_sd, lengths=lens)
    emb = gen.speaker_embedding(torch.as_tensor(mel).cuda(), None if lens is None else torch.as_tensor(lens)).cpu().numpy()
    err = np.abs(emb - ref).max()
    print("native ECAPA vs oracle", (B, Tm), err, "scale", np.abs(ref).max())
    assert err <= 3e-4 * max(1.0, np.abs(ref).max())


def test_pcm16_output_matches_reference_postprocessing(gen):
    """bvg_forward_pcm16 == the callers' clamp(32767*wav, -32767, 32767).type(int16) (infer.py:462, :650)."""
    from b200vgan import synth
    x = torch.as_tensor(synth.make_latents(5, 0, 2, 9)).cuda()
    emb = torch.as_tensor(synth.make_speaker_embedding(B=1)).cuda()
    for precision in ("fp32", "bf16"):
        gen.precision = precision
        wav = gen.forward_with_embedding(x, emb, x_lens=[9, 4])
        pcm = gen.forward_with_embedding(x, emb, x_lens=[9, 4], pcm16=True)
        ref = torch.clamp(32767 * wav.squeeze(1), -32767.0, 32767.0).type(torch.int16)
        assert pcm.dtype == torch.int16 and tuple(pcm.shape) == (2, 9 * 1024)
        assert torch.equal(pcm, ref)
        assert int(pcm[1, 4 * 1024:].abs().max()) == 0


@pytest.mark.parametrize("B,sec", [(64, 2.0), (32, 5.0), (2, 30.0)])
def test_cfg5_sweep_corners_bf16_vs_fp32_mode(gen, B, sec):
    """Corners of BASELINE.json config 5 (B in 1..64 x 2..30 s): many short and few long utterances.
    Size-independent properties: bf16 path within its SNR bound of the fp32 path, and batch item i
    identical to the same latents decoded alone (batch invariance)."""
    from b200vgan import synth
    T = synth.frames_for_seconds(sec)
    x = synth.make_latents(5, B, B, T)
    emb = synth.make_speaker_embedding(B=1)
    ref = _run(gen, x, emb, "fp32")
    wav = _run(gen, x, emb, "bf16")
    assert wav.shape == (B, 1, T * 1024) and np.isfinite(wav).all()
    snr = O.snr_db(ref, wav)
    print("cfg5", (B, sec), "bf16-vs-fp32-mode SNR dB", snr)
    assert snr >= BF16_SNR_DB
    i = B - 1
    alone = _run(gen, x[i:i + 1], emb, "bf16")
    assert np.array_equal(alone[0], wav[i])


def test_speaker_encoder_rejects_wrong_mel_width(gen):
    from b200vgan import lib as L
    with pytest.raises(L.BvgError):
        gen.speaker_embedding(torch.zeros(1, 40, 80).cuda())


def test_forward_is_cuda_graph_capturable(gen):
    """bvg_forward allocates nothing and never synchronises (include/b200vgan.h): capture one decode in a
    CUDA graph, replay it on new latents, compare with the eager call bit for bit."""
    from b200vgan import synth
    gen.precision = "bf16"
    emb = torch.as_tensor(synth.make_speaker_embedding(B=1)).cuda()
    x_static = torch.as_tensor(synth.make_latents(6, 0, 2, 12)).cuda()
    eager0 = gen.forward_with_embedding(x_static, emb, x_lens=[12, 7]).clone()   # builds engine, plan, workspace
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_static = gen.forward_with_embedding(x_static, emb, x_lens=[12, 7])
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_static, eager0)
    x_new = torch.as_tensor(synth.make_latents(6, 1, 2, 12)).cuda()
    x_static.copy_(x_new)
    graph.replay()
    torch.cuda.synchronize()
    eager1 = gen.forward_with_embedding(x_new, emb, x_lens=[12, 7])
    torch.cuda.synchronize()
    assert torch.equal(y_static, eager1)


@pytest.mark.parametrize("env", [{"BVG_ACT_MMA": "0"}, {"BVG_ACT_MMA_UPLO": "1"}, {"BVG_FUSE_ACT": "1"},
                                 {"BVG_RES_MMA_MAXC": "0"}])
def test_optional_kernel_paths(env):
    """The opt-in / fallback kernels (register-streamed bf16 activation, hi+lo up-FIR taps, activation fused into
    the conv producer, residual in the epilogue) stay within the bf16 SNR bound on config 1."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "optional_path_check.py")],
                         env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    snr = float([ln for ln in out.stdout.splitlines() if ln.startswith("SNR_DB")][-1].split()[1])
    print(env, "cfg1 bf16 SNR dB", snr)
    assert snr >= BF16_SNR_DB


def test_latent_handoff_device_resident_and_graph_replay(gen):
    """bf16 channels-last latents appended on the device, decoded as one ragged batch; the CUDA-graph
    replay, the eager decode and the decode of the same latents given as fp32 all agree bit for bit (bf16 mode)."""
    from b200vgan import sched, synth
    emb = torch.as_tensor(synth.make_speaker_embedding(B=1)).cuda()
    frames = [9, 31, 4]
    lat32 = [torch.as_tensor(synth.make_latents(8, i, 1, f)).cuda() for i, f in enumerate(frames)]     # [1, T, 1024] like the GPT's
    gen.precision = "bf16"
    h = sched.LatentHandoff(gen, max_rows=64, dtype=torch.bfloat16)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                  # the producer runs on its own stream
        for l in lat32:
            h.append(l)
        ev = torch.cuda.Event()
        ev.record(side)
    h.wait_event(ev)
    eager = [w.clone() for w in h.decode(emb)]
    replay1 = [w.clone() for w in h.decode(emb, graph=True)]     # captures
    # new latents in the same buffer, same geometry: the replay must see them
    h.reset()
    for l in lat32:
        h.append(2.0 * l)
    replay2 = [w.clone() for w in h.decode(emb, graph=True)]     # replays
    eager2 = [w.clone() for w in h.decode(emb)]
    torch.cuda.synchronize()
    rows32 = torch.cat([l[0] for l in lat32])
    direct = gen.forward_ragged(rows32, frames, emb, pcm16=True)
    off = 0
    for i, n in enumerate(frames):
        assert eager[i].dtype == torch.int16 and eager[i].numel() == n * 1024
        assert torch.equal(eager[i], replay1[i]) and torch.equal(replay2[i], eager2[i])
        assert torch.equal(eager[i], direct[off * 1024:(off + n) * 1024])
        assert not torch.equal(eager[i], eager2[i])
        off += n
    gen.precision = "fp32"


# ---- fp32 tensor-core mode (BVG_MODE_FP32_TC, precision="fp32tc"): fp32 storage, convolutions as three bf16 tensor-core
# passes over split operands.  Same gates as the fp32 parity mode: max-abs <= 1e-3, mel-L1 <= 1e-3.

def test_fp32tc_tiny_and_cfg1_match_reference(gen, golden_dir):
    from b200vgan import synth
    g = np.load(os.path.join(golden_dir, "forward_tiny.npz"))
    wav = _run(gen, g["x"], g["emb"], "fp32tc")
    assert wav.shape == g["wav"].shape
    for b in range(wav.shape[0]):
        _gate_fp32(f"fp32tc tiny item {b}", wav[b, 0], g["wav"][b, 0])
    g1 = np.load(os.path.join(golden_dir, "forward_cfg1.npz"))
    x = synth.make_latents(1, 0, 1, 118)
    wav = _run(gen, x, g1["emb"], "fp32tc")
    _gate_fp32("fp32tc cfg1 (reference waveform)", wav[0, 0], g1["wav"][0, 0])


def test_fp32tc_dropin_call_under_no_autocast(gen, golden_dir):
    """wav, _ = bigvgan(latent, mel_ref) in the fp32 tensor-core mode, speaker encoder included."""
    g = np.load(os.path.join(golden_dir, "forward_tiny.npz"))
    gen.precision = "fp32tc"
    try:
        wav, aux = gen(torch.as_tensor(g["x"]).cuda(), torch.as_tensor(g["mel"]).cuda())
    finally:
        gen.precision = "fp32"
    assert aux is None and wav.dtype == torch.float32
    assert np.abs(wav.cpu().numpy() - g["wav"]).max() <= FP32_MAXABS


def test_fp32tc_cfg2_tracks_fp32_mode_and_oracle(gen):
    """Config 2 (B=16 x 10 s): every item against the FFMA parity mode, item 7 against the CPU oracle."""
    from b200vgan import synth
    x = synth.make_latents(2, 0, 16, 235)
    emb = synth.make_speaker_embedding(B=1)
    w32 = _run(gen, x, emb, "fp32")
    wtc = _run(gen, x, emb, "fp32tc")
    d = float(np.abs(wtc - w32).max())
    print("cfg2 fp32tc-vs-fp32-mode max-abs", d, "SNR dB", O.snr_db(w32.ravel(), wtc.ravel()))
    assert d <= 5e-4
    ref = _oracle(x[7:8], emb)[0, 0]
    _gate_fp32("fp32tc cfg2 item 7", wtc[7, 0], ref)


def test_fp32tc_ragged_and_long_form(gen):
    """Ragged batch (each segment = its stand-alone decode, as in the other modes) and the 60 s long form against the
    FFMA parity mode (which test_cfg4_long_form_vs_oracle gates against the oracle)."""
    from b200vgan import synth
    emb = synth.make_speaker_embedding(B=1)
    lens = [37, 5, 118, 64]
    x = synth.make_latents(3, 0, 4, max(lens))
    both = _run(gen, x, emb, "fp32tc", lens=lens)
    for b, n in enumerate(lens):
        one = _run(gen, x[b:b + 1, :n], emb, "fp32tc")
        assert np.array_equal(both[b, 0, :n * 1024], one[0, 0]), b
        assert not both[b, 0, n * 1024:].any()
    T = 1407
    x = synth.make_latents(4, 0, 1, T)
    w32 = _run(gen, x, emb, "fp32")[0, 0]
    wtc = _run(gen, x, emb, "fp32tc")[0, 0]
    d = float(np.abs(wtc - w32).max())
    print("cfg4 fp32tc-vs-fp32-mode max-abs", d)
    assert d <= 5e-4
    edge = 36000
    assert np.abs(wtc[:edge] - w32[:edge]).max() <= 5e-4 and np.abs(wtc[-edge:] - w32[-edge:]).max() <= 5e-4


def test_fp32tc_pcm16_and_ragged_rows(gen):
    """The fused int16 output and the back-to-back ragged entry point in the fp32 tensor-core mode."""
    from b200vgan import synth
    emb = torch.as_tensor(synth.make_speaker_embedding(B=1)).cuda()
    frames = [9, 33, 17]
    rows = torch.as_tensor(np.random.default_rng(5).standard_normal((sum(frames), 1024), dtype=np.float32)).cuda()
    gen.precision = "fp32tc"
    try:
        flat = gen.forward_ragged(rows, frames, emb).cpu().numpy()
        pcm = gen.forward_ragged(rows, frames, emb, pcm16=True).cpu().numpy()
    finally:
        gen.precision = "fp32"
    ref = gen.forward_ragged(rows, frames, emb).cpu().numpy()
    assert flat.shape == ref.shape == (sum(frames) * 1024,)
    assert np.abs(flat - ref).max() <= 5e-4
    want = np.clip(32767.0 * flat, -32767.0, 32767.0).astype(np.int16)
    assert np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_lockstep_blocks_equal_sequential_blocks(tmp_path):
    """BVG_ACT_GROUP (default 1): the three AMP blocks of a stage advance in lockstep and share their Activation1d launches.
    Same kernels on the same tensors in another order: the waveform must not change by a bit (bf16 and fp16 modes, ragged
    batch).  The switch is read once per process, so each setting runs in its own interpreter."""
    import subprocess
    import sys
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(os.getcwd(), "index-tts-dubbing_b200"))
import numpy as np, torch
from b200vgan import synth
from b200vgan.model import BigVGAN
g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
sd = synth.make_state_dict(1234, with_speaker_encoder=False)
g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
g = g.to("cuda"); g.remove_weight_norm(); g.eval()
x = torch.from_numpy(synth.make_latents(9, 0, 3, 40)).cuda()
emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
out = {}
for prec in ("bf16", "fp16"):
    g.precision = prec
    out[prec] = g.forward_with_embedding(x, emb, x_lens=[40, 7, 23]).cpu().numpy()
    out[prec + "_launches"] = np.int64(g.num_launches([40, 7, 23]))
np.savez(sys.argv[1], **out)
'''
    res = {}
    # "1": the default (lockstep, independent convolutions start under their predecessor's tail); "1c": lockstep with every
    # launch chained (BVG_PDL_INDEP=0); "0": the sequential block order
    for v, extra in (("0", {}), ("1", {}), ("1c", {"BVG_PDL_INDEP": "0"})):
        path = str(tmp_path / f"group{v}.npz")
        env = dict(os.environ, BVG_ACT_GROUP=v[0], **extra)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       timeout=600)
        res[v] = np.load(path)
    for prec in ("bf16", "fp16"):
        assert np.array_equal(res["0"][prec], res["1"][prec]), prec
        assert np.array_equal(res["1c"][prec], res["1"][prec]), prec
        assert int(res["1"][prec + "_launches"]) == int(res["0"][prec + "_launches"]) - 72   # 6 stages x 6 steps x 2 launches saved


def test_fp32tc_weight_reload_rebuilds_split_images():
    """load_state_dict on a model that already decoded in the fp32 tensor-core mode: the split weight images are rebuilt from
    the new weights (same handle), not reused."""
    from b200vgan import synth
    from b200vgan.model import BigVGAN
    g = BigVGAN(dict(synth.H_DEFAULT), precision="fp32tc")
    sd1 = synth.make_state_dict(1234, with_speaker_encoder=False)
    sd2 = synth.make_state_dict(4321, with_speaker_encoder=False)
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd1.items()}, strict=False)
    g = g.to("cuda")
    g.remove_weight_norm()
    g.eval()
    x = synth.make_latents(6, 0, 1, 5)
    emb = synth.make_speaker_embedding(B=1)
    w1 = _run(g, x, emb, "fp32tc")
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd2.items()}, strict=False)
    w2 = _run(g, x, emb, "fp32tc")
    ref2 = _run(g, x, emb, "fp32")
    assert np.abs(w2 - w1).max() > 1e-2            # other weights, other audio
    assert np.abs(w2 - ref2).max() <= 5e-4         # and it is the new weights' audio
