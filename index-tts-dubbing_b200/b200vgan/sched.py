"""Segment scheduling for multi-utterance / SRT-dubbing decode.

Utterances are independent, so the path shards with NO data-path collective: one process per GPU,
each decodes its own shard in length-bucketed, exact-variable-length ("ragged") batches.  The only
communication is the final gather of the 16-bit results onto rank 0 (`gather_results`,
torch.distributed: NCCL on GPUs, gloo in the CPU tests) -- never on the hot path.

What this replaces in the reference:
  * srt_dubbing/src/strategies/stretch_strategy.py:72-83 -- one `tts.infer` (one B=1 vocoder call) per
    subtitle entry, each followed by its own blocking `.cpu()`;
  * indextts/infer.py:439-463 (`infer_fast`) -- `chunk_size = 2` sentence latents concatenated along TIME and
    decoded as one B=1 sequence (which smears the sentences' boundaries into each other), then
    `clamp(32767 * wav)` + `.cpu()` per chunk (infer.py:462-463).
Here every segment is decoded exactly as if alone (per-segment lengths reach the kernels, the packed
layout separates segments by zero guard rows), B > 1 per launch sequence, int16 PCM comes straight out of
the last kernel, and one asynchronous device-to-host copy per batch lands in a pinned result arena.

Data movement per batch (`decode_shard`): the segments' latents are gathered into one contiguous
[sum T_i, gpt_dim] device matrix (one `torch.cat` of device tensors, or one cudaMemcpyAsync per pinned
host tensor -- no zero fill, no padding), `BigVGAN.forward_ragged` writes the batch's slice of the
shard's result vector, and a copy stream moves that slice to the host while the next batch decodes.
"""
from __future__ import annotations

from dataclasses import dataclass
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


@dataclass
class ShardResult:
    """One rank's decoded segments: ONE flat vector (int16 PCM or fp32) plus an index.
    `flat` lives on the device (`host` is its pinned mirror when `to_host` was requested and is complete after
    `wait()`); `index[i] = (offset, samples)` for every segment this rank decoded."""
    flat: torch.Tensor
    host: Optional[torch.Tensor]
    index: Dict[int, Tuple[int, int]]
    done: Optional[torch.cuda.Event] = None
    batches: int = 0

    def wait(self) -> "ShardResult":
        if self.done is not None:
            self.done.synchronize()
            self.done = None
        return self

    def segment(self, i: int, on_host: bool = True) -> torch.Tensor:
        off, n = self.index[i]
        if on_host and self.host is not None:
            self.wait()
            return self.host[off:off + n]
        return self.flat[off:off + n]


@torch.no_grad()
def decode_shard(model, latents: Sequence[Optional[torch.Tensor]], emb: torch.Tensor,
                 indices: Optional[Sequence[int]] = None, max_batch_frames: int = 4096, max_batch: int = 64,
                 to_host: bool = True, int16: bool = True) -> ShardResult:
    """Decode the segments `indices` of `latents` (latents[i]: [T_i, gpt_dim], on the model's device or in
    (ideally pinned) host memory; entries this rank does not own may be None) in ragged batches.
    Stream-ordered: returns without synchronising; `ShardResult.wait()` / `.segment()` do."""
    if indices is None:
        indices = [i for i, l in enumerate(latents) if l is not None]
    indices = list(indices)
    dev = model.conv_pre.bias.device
    hop = model.hop
    frames = {i: int(latents[i].shape[0]) for i in indices}
    batches = make_batches(indices, frames, max_batch_frames, max_batch)
    total = sum(frames.values())
    dt = torch.int16 if int16 else torch.float32
    flat = torch.empty(total * hop, device=dev, dtype=dt)
    host = torch.empty(total * hop, dtype=dt, pin_memory=True) if to_host else None
    index: Dict[int, Tuple[int, int]] = {}
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev) if to_host else None
    gpt_dim = latents[indices[0]].shape[1] if indices else 0
    off = 0
    for batch in batches:
        nfr = [frames[i] for i in batch]
        rows_n = sum(nfr)
        parts = [latents[i] for i in batch]
        if all(p.is_cuda for p in parts):
            rows = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)           # one gather kernel
        else:
            rows = torch.empty(rows_n, gpt_dim, device=dev, dtype=parts[0].dtype)
            r = 0
            for p_, n in zip(parts, nfr):                                             # one cudaMemcpyAsync per segment
                rows[r:r + n].copy_(p_, non_blocking=True)
                r += n
        out = flat[off * hop:(off + rows_n) * hop]
        model.forward_ragged(rows, nfr, emb, pcm16=int16, out=out)
        r = off
        for i, n in zip(batch, nfr):
            index[i] = (r * hop, n * hop)
            r += n
        if to_host:                                                                   # one async D2H per batch
            ev = torch.cuda.Event()
            ev.record(main)
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                host[off * hop:(off + rows_n) * hop].copy_(out, non_blocking=True)
        off += rows_n
    done = None
    if to_host:
        done = torch.cuda.Event()
        done.record(copy_stream)
        flat.record_stream(copy_stream)
    return ShardResult(flat=flat, host=host, index=index, done=done, batches=len(batches))


@torch.no_grad()
def decode_segments(model, latents: Sequence[Optional[torch.Tensor]], emb: torch.Tensor,
                    indices: Optional[Sequence[int]] = None, max_batch_frames: int = 4096, max_batch: int = 64,
                    to_host: bool = True, int16: bool = False) -> Dict[int, torch.Tensor]:
    """`decode_shard` with the results split per segment: {segment index: waveform [T_i*hop]} (views of one
    pinned host vector if `to_host`, of one device vector otherwise)."""
    res = decode_shard(model, latents, emb, indices, max_batch_frames, max_batch, to_host, int16)
    res.wait()
    return {i: res.segment(i, on_host=to_host) for i in res.index}


@torch.no_grad()
def decode_sentences(model, latents: Sequence[torch.Tensor], cond_mel: torch.Tensor, max_batch_frames: int = 4096,
                     max_batch: int = 64, cache_key=None) -> List[torch.Tensor]:
    """Drop-in for the vocoder loop of `IndexTTS.infer_fast` (indextts/infer.py:439-463):

        wavs = sched.decode_sentences(self.bigvgan, all_latents, auto_conditioning.transpose(1, 2))
        wav = torch.cat(wavs, dim=1)                         # infer.py:473

    `latents`: the per-sentence GPT latents [1, T_i, gpt_dim] in sentence order (infer.py:441),
    `cond_mel`: the prompt mel [1, Tm, num_mels] (`auto_conditioning.transpose(1, 2)`).
    Returns, in the same order, int16 waveforms [1, T_i*hop] on the host -- already
    `clamp(32767 * wav, -32767, 32767)` (infer.py:462) and cast (infer.py:488, :492).  Sentences are decoded as a true
    B > 1 ragged batch instead of being concatenated along time, so each one equals its stand-alone decode."""
    emb = model.speaker_embedding(cond_mel, cache_key=cache_key)
    segs = [l[0] if l.dim() == 3 else l for l in latents]
    out = decode_segments(model, segs, emb, None, max_batch_frames, max_batch, to_host=True, int16=True)
    return [out[i][None] for i in range(len(segs))]


def gather_results(res: ShardResult, n_total: int, group=None, device: Optional[torch.device] = None):
    """Gather every rank's flat result vector onto rank 0: returns the list of per-segment host tensors (views of
    one pinned buffer) on rank 0, None elsewhere.  Traffic: one all_gather of a [n_total, 2] index table (who owns
    what, where) and ONE gather of the flat vectors padded to the longest shard -- device to device over NCCL
    (NVLink) followed by one device-to-host copy per shard on rank 0, or host tensors over gloo.  Never on the hot
    path."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        res.wait()
        return [res.segment(i) for i in range(n_total)]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = res.flat.device if on_gpu else torch.device("cpu")
    table = torch.full((n_total, 2), -1, dtype=torch.int64)
    for i, (off, n) in res.index.items():
        table[i, 0], table[i, 1] = off, n
    table = table.to(dev)
    tables = [torch.empty_like(table) for _ in range(world)]
    dist.all_gather(tables, table, group=group)
    tables = [t.cpu() for t in tables]
    sizes = [int(t[:, 1].clamp(min=0).sum()) for t in tables]
    pad = max(sizes)
    if on_gpu:
        src = res.flat
    else:
        res.wait()
        src = res.host if res.host is not None else res.flat.cpu()
    mine = torch.zeros(pad, dtype=src.dtype, device=dev)
    mine[: src.numel()].copy_(src)
    parts = [torch.empty(pad, dtype=src.dtype, device=dev) for _ in range(world)] if rank == 0 else None
    # byte views: gloo has no int16 collectives
    dist.gather(mine.view(torch.uint8), [p_.view(torch.uint8) for p_ in parts] if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    host = torch.empty(world, pad, dtype=src.dtype, pin_memory=torch.cuda.is_available())
    for r in range(world):
        host[r].copy_(parts[r], non_blocking=on_gpu)
    if on_gpu:
        torch.cuda.current_stream(dev).synchronize()
    out: List[Optional[torch.Tensor]] = [None] * n_total
    for r, t in enumerate(tables):
        for i in torch.nonzero(t[:, 1] >= 0).flatten().tolist():
            off, n = int(t[i, 0]), int(t[i, 1])
            out[i] = host[r, off:off + n]
    missing = [i for i, w in enumerate(out) if w is None]
    if missing:
        raise RuntimeError(f"gather_results: segments {missing[:8]} were decoded by no rank")
    return out


class LatentHandoff:
    """Device-resident hand-off between the GPT latent producer and the vocoder.

    Reference: `latent = self.gpt(..., return_latent=True)` (indextts/gpt/model.py:462-488 `get_logits` returns
    `final_norm(last_hidden_state)[:, :T]`, infer.py:614-619) is a [1, T, 1024] tensor per sentence that the reference
    hands to `self.bigvgan(latent, ...)` one sentence (or one time-concatenated chunk, infer.py:439-458) at a time,
    fp32 or autocast-fp16, followed by `.cpu()` per call.

    Contract offered to the producer instead:
      * layout   rows: ONE persistent [max_rows, gpt_dim] matrix, sentence after sentence ("channels-last" is what
                 `final_norm` already emits, so `append` is a single strided copy with the dtype cast fused; a producer
                 that can write in place uses `rows(n)` as its `out=` target and skips even that);
      * dtype    bf16 by default (what the bf16 mode stores anyway: bit-identical results to fp32 latents), fp16 / fp32 accepted;
      * stream   everything is ordered on torch's current stream -- no host synchronisation; a producer on another
                 stream records an event after its last write and the consumer stream waits on it (`wait_event`);
      * graph    `decode(graph=True)` captures the whole ragged decode (pack -> 162 launches -> int16 PCM) for this
                 sentence geometry into a CUDA graph on first use and replays it afterwards: the buffers the graph
                 reads and writes are owned here, so their addresses never change (one launch per decode from the host).
    """

    def __init__(self, model, max_rows: int, dtype: torch.dtype = torch.bfloat16, int16: bool = True):
        dev = model.conv_pre.bias.device
        if dev.type != "cuda":
            raise RuntimeError("LatentHandoff: the model must live on a CUDA (sm_100) device")
        self.model, self.int16 = model, int16
        self.buf = torch.empty(int(max_rows), model._cfg.gpt_dim, device=dev, dtype=dtype)
        self.out = torch.empty(int(max_rows) * model.hop, device=dev, dtype=torch.int16 if int16 else torch.float32)
        self.emb = torch.empty(1, 1, model._cfg.speaker_embedding_dim, device=dev, dtype=torch.float32)
        self.frames: List[int] = []
        self.used = 0
        self._graphs: Dict[Tuple[Tuple[int, ...], str], "torch.cuda.CUDAGraph"] = {}

    def reset(self) -> None:
        self.frames, self.used = [], 0

    def rows(self, n: int) -> torch.Tensor:
        """The next `n` rows of the hand-off matrix, for a producer that writes its latent in place; counted as one sentence."""
        if self.used + n > self.buf.shape[0]:
            raise RuntimeError(f"LatentHandoff: {self.used + n} rows exceed the buffer ({self.buf.shape[0]})")
        v = self.buf[self.used:self.used + n]
        self.frames.append(int(n))
        self.used += int(n)
        return v

    def append(self, latent: torch.Tensor) -> None:
        """`latent`: [1, T, gpt_dim] or [T, gpt_dim] on the device (any float dtype): one asynchronous cast-and-copy."""
        lat = latent[0] if latent.dim() == 3 else latent
        self.rows(lat.shape[0]).copy_(lat, non_blocking=True)

    def wait_event(self, event: "torch.cuda.Event") -> None:
        """Order the decode after a producer that ran on another stream."""
        torch.cuda.current_stream(self.buf.device).wait_event(event)

    @torch.no_grad()
    def decode(self, emb: torch.Tensor, graph: bool = False) -> List[torch.Tensor]:
        """Decode every appended sentence as one exact ragged batch.  Returns per-sentence views of the device result
        vector (int16 PCM, or fp32 with int16=False), valid until the next `decode`."""
        if not self.frames:
            return []
        m, frames = self.model, tuple(self.frames)
        self.emb.copy_(emb.reshape(1, 1, -1), non_blocking=True)
        rows, out = self.buf[:self.used], self.out[:self.used * m.hop]
        if not graph:
            m.forward_ragged(rows, frames, self.emb, pcm16=self.int16, out=out)
        else:
            key = (frames, m.resolved_precision())
            g = self._graphs.get(key)
            if g is None:
                m.forward_ragged(rows, frames, self.emb, pcm16=self.int16, out=out)    # builds the plan, uploads its tables
                torch.cuda.current_stream(self.buf.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    m.forward_ragged(rows, frames, self.emb, pcm16=self.int16, out=out)
                if len(self._graphs) >= 32:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = g
            g.replay()
        res, off = [], 0
        for n in frames:
            res.append(out[off * m.hop:(off + n) * m.hop])
            off += n
        return res
