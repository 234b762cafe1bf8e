"""Timeline merge of the dubbing tool: host placement logic on CPU, the GPU kernel under -m gpu.
Reference: srt_dubbing/src/audio_processor.py:70-230; fixture tests/golden/srt_merge.npz holds the unmodified
reference's outputs (oracle/gen_golden.py, group "merge")."""
import os

import numpy as np
import pytest

from oracle import srt_merge_oracle as M


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "srt_merge.npz"))
    for name in ("a", "b", "c", "d"):
        ln = g[f"{name}_len"]
        off = np.concatenate([[0], np.cumsum(ln)])
        segs = [{"index": i + 1, "start_time": float(g[f"{name}_start"][i]), "end_time": float(g[f"{name}_end"][i]),
                 "audio_data": g[f"{name}_audio"][off[i]:off[i + 1]]} for i in range(len(ln))]
        yield name, int(g[f"{name}_sr"]), segs, g


def test_host_placement_matches_oracle(golden_dir):
    """b200vgan.timeline.plan_time_synchronized (product host logic) == the oracle's placement, on the fixture cases and
    on random timelines (empty segments, ties, overflow past the nominal end)."""
    from b200vgan import timeline as T
    rng = np.random.default_rng(5)
    trials = [(sr, [s["start_time"] for s in segs], [s["end_time"] for s in segs], [len(s["audio_data"]) for s in segs])
              for _, sr, segs, _ in _cases(golden_dir)]
    for _ in range(200):
        n = int(rng.integers(1, 12))
        st = rng.uniform(0, 2.0, n).round(int(rng.integers(1, 4))).tolist()
        trials.append((int(rng.choice([8000, 16000, 24000])), st, (np.array(st) + rng.uniform(0, 1, n)).tolist(),
                       [int(v) for v in rng.integers(0, 9000, n) * (rng.uniform(size=n) > 0.1)]))
    for sr, st, en, ln in trials:
        for flag in (False, True):
            assert T.plan_time_synchronized(st, en, ln, sr, flag) == M.plan_time_synchronized(st, en, ln, sr, flag)


@pytest.mark.gpu
@pytest.mark.parametrize("flag", [False, True])
def test_gpu_timeline_merge_is_bit_exact(golden_dir, flag):
    import torch
    from b200vgan import timeline as T
    for name, sr, segs, g in _cases(golden_dir):
        dsegs = [dict(s, audio_data=torch.as_tensor(s["audio_data"]).cuda()) for s in segs]
        out = T.merge_time_synchronized(dsegs, sr, truncate_on_overflow=flag)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"{name}_merged_trunc{int(flag)}"])
        np.testing.assert_array_equal(T.natural_concatenation(dsegs).cpu().numpy(), g[f"{name}_natural"])


@pytest.mark.gpu
def test_gpu_timeline_merge_of_decoded_segments(synth_sd):
    """Vocoder -> timeline without leaving the device: decode three ragged segments (fp32), place them on a subtitle
    timeline, compare with the oracle merge of the same waveforms."""
    import torch
    from b200vgan import sched, synth, timeline as T
    from b200vgan.model import BigVGAN
    g = BigVGAN(dict(synth.H_DEFAULT), precision="bf16")
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth_sd.items()})
    g = g.to("cuda"); g.remove_weight_norm(); g.eval()
    emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).cuda()
    frames = [12, 30, 7]
    lat = [torch.from_numpy(synth.make_latents(3, i, 1, f)[0]).cuda() for i, f in enumerate(frames)]
    res = sched.decode_shard(g, lat, emb, to_host=False, int16=False)
    starts = [0.2, 0.9, 0.6]     # the third overlaps the second
    segs = [{"index": i + 1, "start_time": starts[i], "end_time": starts[i] + 1.0, "audio_data": res.segment(i, on_host=False)} for i in range(3)]
    out = T.merge_time_synchronized(segs, 24000, truncate_on_overflow=True)
    torch.cuda.synchronize()
    ref = M.time_synchronized_merge([dict(s, audio_data=s["audio_data"].cpu().numpy()) for s in segs], 24000, True)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_timeline_merge_rejects_cpu_tensors():
    import torch
    from b200vgan import lib, timeline as T
    with pytest.raises(lib.BvgError):
        T.merge_time_synchronized([{"start_time": 0.0, "audio_data": torch.zeros(10)}], 24000)
