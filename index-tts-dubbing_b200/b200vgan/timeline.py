"""Timeline merge of decoded subtitle segments on the GPU, with the reference's interface and arithmetic:
`AudioProcessor._time_synchronized_merge` / `_natural_concatenation` (srt_dubbing/src/audio_processor.py:70-230).

The decoded segments stay on the device between the vocoder (`sched.decode_shard`) and the finished timeline; only the
placement arithmetic (a few integers per segment) runs on the host.  Bit-exact against the reference for fp32 segments
(tests/golden/srt_merge.npz): segments are added in start-time order, the peak normalisation divides in fp32."""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import torch

from . import lib as _lib

DYNAMIC_BUFFER_SIZE = 1024    # srt_dubbing/src/config.py:20
MAX_AMPLITUDE = 1.0           # srt_dubbing/src/config.py:21
MAX_COVER = 16                # most segments the kernel adds into one sample


def plan_time_synchronized(start_times: Sequence[float], end_times: Sequence[float], lengths: Sequence[int], sample_rate: int,
                           truncate_on_overflow: bool):
    """Placement of audio_processor.py:157-218: (order, start sample per sorted segment or -1 for an empty one, total samples).
    Note the reference's overlap rule compares with the PREVIOUS segment's nominal start, not its shifted one."""
    order = sorted(range(len(start_times)), key=lambda i: start_times[i])
    max_end = 0.0
    for i in order:
        max_end = max(max_end, start_times[i] + lengths[i] / sample_rate if lengths[i] > 0 else end_times[i])
    total = int(max_end * sample_rate) + DYNAMIC_BUFFER_SIZE
    starts: List[int] = []
    for k, i in enumerate(order):
        if lengths[i] == 0:
            starts.append(-1)
            continue
        s = int(start_times[i] * sample_rate)
        if not truncate_on_overflow and k > 0:
            j = order[k - 1]
            prev_end = int(start_times[j] * sample_rate) + lengths[j]
            if s < prev_end:
                s = prev_end
        if s + lengths[i] > total:
            total = s + lengths[i] + DYNAMIC_BUFFER_SIZE
        starts.append(s)
    return order, starts, total


@torch.no_grad()
def merge_time_synchronized(segments: List[Dict[str, Any]], sample_rate: int = 24000, truncate_on_overflow: bool = False) -> torch.Tensor:
    """`segments`: dicts with 'start_time', optional 'end_time', and 'audio_data' = a 1-D fp32 CUDA tensor (e.g. a view of
    `ShardResult.flat`).  Returns the merged fp32 timeline on the same device."""
    if not segments:
        raise _lib.BvgError("merge_time_synchronized: no segments")
    audio = [s["audio_data"] for s in segments]
    dev = next((a.device for a in audio if isinstance(a, torch.Tensor)), None)
    if dev is None or dev.type != "cuda" or any((not isinstance(a, torch.Tensor)) or a.device != dev or a.dtype != torch.float32 or a.dim() != 1
                                                 for a in audio):
        raise _lib.BvgError("merge_time_synchronized: every audio_data must be a 1-D fp32 tensor on one CUDA device (no CPU path)")
    st = [float(s["start_time"]) for s in segments]
    en = [float(s.get("end_time", s["start_time"])) for s in segments]
    ln = [int(a.numel()) for a in audio]
    order, starts, total = plan_time_synchronized(st, en, ln, sample_rate, truncate_on_overflow)
    ref_order = [(i, s) for i, s in zip(order, starts) if s >= 0]  # the order the reference adds in
    keep = sorted(ref_order, key=lambda p: p[1])                   # the kernel wants ascending starts (stable sort)
    rank_of = {i: r for r, (i, _) in enumerate(ref_order)}          # the kernel adds covering segments by this rank
    flat = torch.cat([audio[i] for i, _ in keep]) if keep else torch.zeros(0, device=dev)
    src, off = [], 0
    for i, _ in keep:
        src.append(off)
        off += ln[i]
    # coverage bound of the kernel
    ends = sorted((s + ln[i], s) for i, s in keep)
    active, cover, ei = 0, 0, 0
    for i, s in keep:
        while ei < len(ends) and ends[ei][0] <= s:
            active -= 1
            ei += 1
        active += 1
        cover = max(cover, active)
    if cover > MAX_COVER:
        raise _lib.BvgError(f"merge_time_synchronized: {cover} segments overlap at one sample (limit {MAX_COVER})")
    L = _lib.load()
    out = torch.empty(total, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        t_src = torch.tensor(src, dtype=torch.int64).to(dev, non_blocking=True)
        t_dst = torch.tensor([s for _, s in keep], dtype=torch.int64).to(dev, non_blocking=True)
        t_n = torch.tensor([ln[i] for i, _ in keep], dtype=torch.int32).to(dev, non_blocking=True)
        t_rank = torch.tensor([rank_of[i] for i, _ in keep], dtype=torch.int32).to(dev, non_blocking=True)
        peak = torch.empty(1, device=dev, dtype=torch.int32)
        _lib.check(L.bvg_timeline_merge(flat.data_ptr(), t_src.data_ptr(), t_dst.data_ptr(), t_n.data_ptr(), t_rank.data_ptr(), len(keep),
                                        max([ln[i] for i, _ in keep], default=0), out.data_ptr(), total,
                                        0 if truncate_on_overflow else 1, MAX_AMPLITUDE, peak.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream))
    return out


@torch.no_grad()
def natural_concatenation(segments: List[Dict[str, Any]]) -> torch.Tensor:
    """audio_processor.py:86-127: subtitle-index order, empty segments skipped, back to back (one device `cat`)."""
    parts = [s["audio_data"] for s in sorted(segments, key=lambda x: x.get("index", 0))]
    parts = [p for p in parts if p.numel()]
    return torch.cat(parts) if parts else torch.zeros(0)
