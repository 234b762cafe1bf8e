"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the torch module
mirrors the reference's state-dict layouts, scheduling / sharding logic, and the N>1 gather path
over gloo (world_size 2).  No compute call is made without a GPU."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from b200vgan import lib
    L = lib.load()
    header = open(os.path.join(ROOT, "include", "b200vgan.h")).read()
    declared = set(re.findall(r"\b(bvg_[a-z0-9_]+)\s*\(", header))
    declared -= {"bvg_config", "bvg_handle", "bvg_plan"}
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), f"libb200vgan.so does not export {s}"
    assert L.bvg_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from b200vgan import lib, synth
    L = lib.load()
    assert L.bvg_device_check() != 0
    assert b"no CPU fallback" in L.bvg_last_error() or b"CUDA" in L.bvg_last_error()
    cfg = lib.make_config(synth.H_DEFAULT)
    hd = ctypes.c_void_p()
    assert L.bvg_create(ctypes.byref(cfg), ctypes.byref(hd)) != 0
    x = np.zeros((1, 8, 4), np.float32)
    a = np.zeros(8, np.float32)
    rc = L.bvg_activation1d(x.ctypes.data, x.ctypes.data, a.ctypes.data, a.ctypes.data, 1, 8, 4, 0, None)
    assert rc != 0


def test_config_struct_matches_reference_yaml():
    from b200vgan import lib, synth
    cfg = lib.make_config(dict(synth.H_DEFAULT))
    assert cfg.gpt_dim == 1024 and cfg.upsample_initial_channel == 1536
    assert list(cfg.upsample_rates)[:6] == [4, 4, 4, 4, 2, 2]
    assert list(cfg.upsample_kernel_sizes)[:6] == [8, 8, 4, 4, 4, 4]
    assert [list(r)[:3] for r in cfg.resblock_dilation_sizes][:3] == [[1, 3, 5]] * 3
    with pytest.raises(lib.BvgError):
        lib.make_config(dict(synth.H_DEFAULT, activation="snake"))


@pytest.fixture(scope="module")
def module_cpu():
    from b200vgan import synth
    from b200vgan.model import BigVGAN
    torch.manual_seed(0)
    return BigVGAN(dict(synth.H_DEFAULT), use_cuda_kernel=True)


def test_state_dict_layouts(module_cpu, synth_sd):
    from b200vgan import synth
    g = module_cpu
    keys = set(g.state_dict().keys())
    assert len(keys) == 1029                       # checkpoint layout
    assert {"conv_pre.weight_g", "conv_pre.weight_v", "ups.0.0.weight_g",
            "resblocks.17.activations.5.downsample.lowpass.filter", "activation_post.upsample.filter",
            "speaker_encoder.blocks.1.res2net_block.blocks.6.norm.norm.num_batches_tracked",
            "speaker_encoder.asp_bn.norm.running_var", "speaker_encoder.fc.conv.weight",
            "cond_layer.weight", "conds.5.bias"} <= keys
    assert tuple(g.state_dict()["ups.0.0.weight_g"].shape) == (1536, 1, 1)
    wn = synth.make_state_dict(seed=1234, weight_norm=True)
    assert set(wn.keys()) == keys
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in wn.items()})
    g.remove_weight_norm()
    g.remove_weight_norm()                         # idempotent
    keys2 = set(g.state_dict().keys())
    assert len(keys2) == 913 and set(synth_sd.keys()) == keys2
    folded = g.folded_state()
    for k in ("conv_pre.weight", "ups.3.0.weight", "resblocks.9.convs1.2.weight", "conv_post.weight"):
        np.testing.assert_allclose(folded[k].numpy(), synth_sd[k], rtol=2e-6, atol=1e-8)
    # folded layout loads too, and the module can go back to the checkpoint layout
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth_sd.items()})
    np.testing.assert_array_equal(g.folded_state()["resblocks.0.convs2.0.weight"].numpy(),
                                  synth_sd["resblocks.0.convs2.0.weight"])
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in wn.items()})
    assert len(g.state_dict()) == 1029


def test_speaker_encoder_is_a_parameter_container_only(module_cpu):
    """ECAPA_TDNN.py:429-541 key layout (231 keys) and no torch compute path in the product."""
    enc = module_cpu.speaker_encoder
    keys = set(enc.state_dict().keys())
    assert len(keys) == 231
    assert {"blocks.0.conv.conv.weight", "blocks.1.res2net_block.blocks.3.norm.norm.running_var",
            "blocks.3.se_block.conv2.conv.bias", "mfa.conv.conv.weight", "asp.tdnn.norm.norm.weight",
            "asp.conv.conv.weight", "asp_bn.norm.running_mean", "fc.conv.weight"} <= keys
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 40, 100))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_forward_refuses_cpu(module_cpu):
    from b200vgan import lib
    with pytest.raises(lib.BvgError):
        module_cpu.forward_with_embedding(torch.zeros(1, 4, 1024), torch.zeros(1, 1, 512))


def test_lpt_and_batches():
    from b200vgan import sched
    frames = sched.srt_workload()
    assert len(frames) == 512 and min(frames) >= 24 and max(frames) <= 352
    for world in (1, 2, 4, 8):
        shards = sched.lpt_shards(frames, world)
        assert sorted(i for s in shards for i in s) == list(range(512))
        loads = [sum(frames[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(frames)          # LPT bound
    batches = sched.make_batches(range(512), frames, max_batch_frames=2048, max_batch=16)
    assert sorted(i for b in batches for i in b) == list(range(512))
    for b in batches:
        assert len(b) <= 16 and (len(b) == 1 or sum(frames[i] for i in b) <= 2048)
        assert all(frames[b[k]] >= frames[b[k + 1]] for k in range(len(b) - 1))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))
    from b200vgan import sched
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = sched.srt_workload(n=37, seed=5)
    mine = sched.lpt_shards(frames, world)[rank]
    # stand-in for decode_shard's result: this rank's segments back to back in ONE int16 vector + an index
    index, off = {}, 0
    for i in mine:
        index[i] = (off, frames[i] * 4)
        off += frames[i] * 4
    flat = torch.empty(off, dtype=torch.int16)
    for i, (o, n) in index.items():
        flat[o:o + n] = i
    out = sched.gather_results(sched.ShardResult(flat=flat, host=flat, index=index), len(frames))
    if rank == 0:
        ok = all(out[i].shape[0] == frames[i] * 4 and out[i].dtype == torch.int16 and
                 int(out[i][0]) == i and int(out[i][-1]) == i for i in range(len(frames)))
        q.put(ok)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_precision_resolution_follows_autocast(module_cpu):
    """precision="auto" (constructor default): fp32 tensors like the reference module (the fp32 tensor-core mode, or the
    CUDA-core parity mode when AUTO_FP32_PRECISION says so), the autocast dtype inside an autocast region (infer.py:456,
    :613); explicit modes are taken as given; anything else is rejected."""
    from b200vgan import lib
    m = module_cpu
    assert m.precision == "auto" and m.resolved_precision() == "fp32tc" and m._mode() == lib.MODE_FP32_TC
    m.AUTO_FP32_PRECISION = "fp32"          # (instance override of the class default)
    assert m.resolved_precision() == "fp32" and m._mode() == lib.MODE_FP32
    del m.AUTO_FP32_PRECISION
    for p, mode in (("fp32", lib.MODE_FP32), ("fp32tc", lib.MODE_FP32_TC), ("bf16", lib.MODE_BF16), ("fp16", lib.MODE_F16)):
        m.precision = p
        assert m.resolved_precision() == p and m._mode() == mode
    m.precision = "int8"
    with pytest.raises(lib.BvgError):
        m.resolved_precision()
    m.precision = "auto"


def test_mel_frontend_constructor_mirrors_reference():
    """b200vgan.MelSpectrogramFeatures takes the reference's constructor arguments (feature_extractors.py:25-27); only the
    deployed configuration has a native kernel, and there is no CPU path."""
    import b200vgan
    from b200vgan import lib
    f = b200vgan.MelSpectrogramFeatures()
    assert (f.sample_rate, f.hop_length, f.n_mels) == (24000, 256, 100)
    with pytest.raises(ValueError):
        b200vgan.MelSpectrogramFeatures(padding="valid")
    for kw in ({"n_fft": 2048}, {"padding": "same"}, {"normalize": True}, {"win_length": 512}):
        with pytest.raises(lib.BvgError):
            b200vgan.MelSpectrogramFeatures(**kw)
    with pytest.raises(lib.BvgError):
        f(torch.zeros(1, 4000))
    L = lib.load()
    assert L.bvg_mel_frames(130560, 256) == 511 and L.bvg_mel_frames(24000, 256) == 94
