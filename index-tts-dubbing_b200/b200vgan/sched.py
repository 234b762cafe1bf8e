"""Segment scheduling for multi-utterance / SRT-dubbing decode.

Utterances are independent, so the path shards with NO data-path collective: one process per GPU,
each decodes its own shard in length-bucketed, exact-variable-length batches and returns the
waveforms to the host.  The only communication is the final host-side gather of results
(`gather_waveforms`, torch.distributed: NCCL on GPUs, gloo in the CPU tests).

The reference decodes one subtitle entry at a time on one device
(srt_dubbing/src/strategies/stretch_strategy.py:72-83, indextts/infer.py:622-631); this module
defines the batched, sharded replacement of that loop.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def lpt_shards(frames: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of segments to `world` ranks.
    Returns, per rank, the segment indices it owns (cost model: decode time ~ latent frames)."""
    order = sorted(range(len(frames)), key=lambda i: (-int(frames[i]), i))
    loads = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(frames[i])
    return shards


def make_batches(indices: Sequence[int], frames: Sequence[int], max_batch_frames: int = 4096,
                 max_batch: int = 64) -> List[List[int]]:
    """Length-sorted bucketing: consecutive segments (longest first) share a batch while the sum of
    frames stays under `max_batch_frames`.  Variable lengths inside a batch are exact (per-segment
    lengths are passed to the kernels), so bucketing only serves load balance, not correctness."""
    order = sorted(indices, key=lambda i: (-int(frames[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    tot = 0
    for i in order:
        f = int(frames[i])
        if cur and (tot + f > max_batch_frames or len(cur) >= max_batch):
            batches.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += f
    if cur:
        batches.append(cur)
    return batches


def srt_workload(n: int = 512, seed: int = 2026, lo: float = 1.0, hi: float = 15.0, sr: int = 24000,
                 hop: int = 1024) -> List[int]:
    """Config 3 of BASELINE.json: n segment durations U(lo,hi) seconds -> latent frame counts."""
    dur = np.random.default_rng(seed).uniform(lo, hi, size=n)
    return [int(np.ceil(d * sr / hop)) for d in dur]


@torch.no_grad()
def decode_segments(model, latents: Sequence[torch.Tensor], emb: torch.Tensor, indices: Optional[Sequence[int]] = None,
                    max_batch_frames: int = 4096, max_batch: int = 64, to_host: bool = True,
                    int16: bool = False) -> Dict[int, torch.Tensor]:
    """Decode the given segments (`latents[i]`: [T_i, gpt_dim] on the model's device) in batches.
    Returns {segment index: waveform [T_i*hop]} (on the host, pinned, if `to_host`)."""
    if indices is None:
        indices = list(range(len(latents)))
    frames = [int(l.shape[0]) for l in latents]
    out: Dict[int, torch.Tensor] = {}
    for batch in make_batches(indices, frames, max_batch_frames, max_batch):
        T = max(frames[i] for i in batch)
        x = torch.zeros(len(batch), T, latents[batch[0]].shape[1], device=latents[batch[0]].device,
                        dtype=latents[batch[0]].dtype)
        for k, i in enumerate(batch):
            x[k, : frames[i]] = latents[i]
        # int16: the callers' clamp(32767*wav) + int16 cast (infer.py:627-628, :650), fused into the decode's last kernel
        wav = model.forward_with_embedding(x, emb, x_lens=[frames[i] for i in batch], pcm16=int16)
        for k, i in enumerate(batch):
            w = (wav[k] if int16 else wav[k, 0])[: frames[i] * model.hop]
            out[i] = w.to("cpu", non_blocking=False) if to_host else w
    return out


def gather_waveforms(local: Dict[int, torch.Tensor], n_total: int, group=None) -> Optional[List[torch.Tensor]]:
    """Host-side gather of per-rank results onto rank 0 (returns None elsewhere).
    One variable-length `gather_object` -- never on the hot path."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return [local[i] for i in range(n_total)]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    payload = {i: w.cpu() for i, w in local.items()}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged: Dict[int, torch.Tensor] = {}
    for part in gathered:
        merged.update(part)
    missing = [i for i in range(n_total) if i not in merged]
    if missing:
        raise RuntimeError(f"gather_waveforms: segments {missing[:8]} were decoded by no rank")
    return [merged[i] for i in range(n_total)]
