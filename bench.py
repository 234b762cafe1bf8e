#!/usr/bin/env python
"""bench.py -- BigVGAN2 speech-code decode throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (torch-CPU port)

A "step" is one decode of the configuration BASELINE.json quotes the metric on: config 2,
batch 16 x 10 s of synthetic latents (T = 235 frames -> 240 640 samples each), bf16 tensor-core
mode, on every GPU (weak scaling: per-GPU work is fixed).  `value` = audio-seconds produced by all
ranks / max-over-ranks device time, inputs resident in HBM.  `e2e` = the same metric through the
public module call with HOST (pinned) buffers: H2D of the latents and D2H of the waveform inside
the timed region.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "index-tts-dubbing_b200"))

import numpy as np  # noqa: E402

SR, HOP = 24000, 1024
CFG2_B, CFG2_SEC = 16, 10.0
METRIC = "vocoded audio-sec/sec (RTF^-1)"
UNIT = "audio-s/s"
CONV_FLOP_PER_FRAME = 2 * 1.4586e9      # dense + transposed conv MACs per latent frame


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def _src_hash(*names):
    import hashlib
    h = hashlib.sha1()
    for n in names:
        h.update(open(os.path.join(ROOT, "index-tts-dubbing_b200", "csrc", n), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the `ncu --set full`
    capture of the current release (tools/ncu_traffic.py writes profiles/roofline_traffic.json next to the capture
    summary, keyed by a hash of the kernel's source).  A capture taken from other kernel source is reported as null
    with the reason, never silently reused."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(p))
    except (OSError, ValueError):
        return {"bytes": None, "source": "no capture summary (profiles/roofline_traffic.json)"}
    if d.get("kernel_source_sha1_16") != _src_hash("bvg_conv_umma.cu"):
        return {"bytes": None, "source": f"stale: {d.get('capture')} was taken from other kernel source"}
    return {"bytes": d.get("dram_bytes"), "source": f"{d.get('what')}, {d.get('capture')}"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower() == "active" for r in self.rows)]
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
#   kind "reference": the UNMODIFIED reference module (indextts.BigVGAN.models.BigVGAN, its own torch code path,
#                     use_cuda_kernel=False) imported from git-ignored baseline/_ref/ (pip-installed there from
#                     /root/reference in the build container; it travels to the GPU box with the snapshot);
#   kind "port":      oracle/bigvgan_torch_cpu.py (the same path restated with the same torch operators) when that
#                     install is absent.
# Both run fp32 on every host thread, on B=1 x 235 latent frames (one 10-s item of config 2) per step.
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
CPU_FRAMES = 235


def _host_threads():
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, ncpu))
    return torch.get_num_threads()


def _load_reference_module():
    """The unmodified reference generator on CPU with the synthetic weights, or None if baseline/_ref is absent."""
    if not os.path.isdir(os.path.join(REF_DIR, "indextts", "BigVGAN")):
        return None
    import types
    import torch
    from b200vgan import synth
    try:
        import matplotlib  # noqa: F401  (BigVGAN/utils.py imports it at module scope; absent in this image)
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pylab = types.ModuleType("matplotlib.pylab")
        sys.modules["matplotlib"], sys.modules["matplotlib.pylab"] = mpl, mpl.pylab
    sys.path.insert(0, REF_DIR)
    try:
        from indextts.BigVGAN.models import BigVGAN as RefBigVGAN
    except Exception as e:   # noqa: BLE001
        print(f"bench.py: reference import from baseline/_ref failed ({e!r}); timing the port", file=sys.stderr)
        return None

    class H(dict):           # stands in for the OmegaConf node (attribute + item access)
        __getattr__ = dict.__getitem__

        def __setattr__(self, k, v):
            self[k] = v
    g = RefBigVGAN(H(synth.H_DEFAULT), use_cuda_kernel=False)
    g.remove_weight_norm()
    g.eval()
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(1234).items()}, strict=True)
    return g


def cpu_decode_rate(sample_frames: int, repeats: int, warmup: int):
    """Returns (audio-seconds per second, per-step seconds, threads, kind, description)."""
    import torch
    from b200vgan import synth
    threads = _host_threads()
    ref = _load_reference_module()
    if ref is not None:
        kind, what = "reference", "unmodified reference indextts.BigVGAN.models.BigVGAN.forward(latent, mel_ref) from baseline/_ref (torch CPU path, fp32)"
        mel = torch.from_numpy(synth.make_mel(seed=7, Tm=400, B=1))

        def step(x):
            with torch.no_grad():
                ref(torch.from_numpy(x), mel)
    else:
        from oracle import bigvgan_torch_cpu as TC      # the ONLY place bench.py executes oracle/
        kind, what = "port", "torch-CPU port of the reference path (same F.conv1d/conv_transpose1d operators, oneDNN, fp32)"
        sd = TC.prepare_state_dict(synth.make_state_dict(1234, with_speaker_encoder=False))
        emb = synth.make_speaker_embedding(B=1)

        def step(x):
            TC.bigvgan_forward_with_embedding(x, emb, sd)
    times = []
    for i in range(warmup + repeats):
        x = synth.make_latents(2, i, 1, sample_frames)
        t0 = time.perf_counter()
        step(x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    audio_s = sample_frames * HOP / SR
    return audio_s * len(times) / sum(times), sum(times) / len(times), threads, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = CPU_FRAMES
    rate, step_s, cores, kind, what = cpu_decode_rate(frames, max(1, args.steps), max(0, args.warmup))
    sample = f"B=1 x {frames} latent frames ({frames * HOP / SR:.2f} s audio: one item of cfg2) per step, {what}, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: BigVGAN2 decode batch 16 x 10 s synthetic latents (CPU arm: bounded sample)",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def srt_leg(g, emb, dev, world, rank, dist):
    """Config 3 of BASELINE.json: the 512-segment SRT dubbing job (1-15 s each, ~4096 s of audio), LPT-sharded over
    the ranks, each rank decoding its shard in ragged batches from PINNED HOST latents to int16 PCM, results gathered
    on rank 0's host.  Wall-clocked end to end (max over ranks).  Strong scaling: the job is fixed as N grows; at
    N > 1 rank 0 also runs the whole job alone afterwards so the line carries its own single-GPU reference."""
    import torch
    from b200vgan import sched
    frames = sched.srt_workload()
    n = len(frames)
    shards = sched.lpt_shards(frames, world)
    loads = [sum(frames[i] for i in sh) for sh in shards]
    audio_s = sum(frames) * HOP / SR

    def host_latents(indices, seed):
        tot = sum(frames[i] for i in indices)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        big = torch.empty(tot, 1024, dtype=torch.float32).pin_memory()
        big.copy_(torch.randn(tot, 1024, device=dev, generator=gen))
        lat, off = [None] * n, 0
        for i in indices:
            lat[i] = big[off:off + frames[i]]
            off += frames[i]
        return lat

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_pass(lat, indices, gather):
        """Returns scalars only: the result tensors (pinned host arenas, device vectors) die here, so the next pass
        re-uses their memory from torch's caching allocators instead of paying cudaHostAlloc / cudaMalloc again --
        the steady state of a dubbing service that handles one job after another."""
        sync_all()
        p0 = g.plans_created()
        t0 = time.perf_counter()
        res = sched.decode_shard(g, lat, emb, indices, max_batch_frames=4096, max_batch=64, to_host=not gather, int16=True)
        if gather:
            torch.cuda.synchronize(dev)
            t1 = time.perf_counter()
            out = sched.gather_results(res, n)
        else:
            res.wait()
            t1 = time.perf_counter()
            out = [res.segment(i) for i in indices]
        t2 = time.perf_counter()
        ok = True
        if rank == 0:
            ok = all(out[k].numel() == frames[i] * HOP for k, i in enumerate(range(n) if gather else indices))
        return t2 - t0, t2 - t1, g.plans_created() - p0, res.batches, ok

    lat = host_latents(shards[rank], 777 + rank)
    first = one_pass(lat, shards[rank], world > 1)          # builds plans, warms the allocators and the table arena
    one_pass(lat, shards[rank], world > 1)
    timed = one_pass(lat, shards[rank], world > 1)
    t = torch.tensor([timed[0], timed[1]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, gather_s = float(t[0]), float(t[1])
    ok = timed[4]
    out = {"audio_s_per_s": audio_s / wall, "wall_ms": wall * 1e3, "segments": n, "audio_s": audio_s,
           "gather_ms": gather_s * 1e3 if world > 1 else 0.0, "host_copy_tail_ms": gather_s * 1e3 if world == 1 else None,
           "plans_created_first_pass": first[2], "plans_created": timed[2], "batches_rank0": timed[3],
           "max_rank_imbalance": max(loads) / (sum(loads) / world), "results_complete": bool(ok),
           "h2d_bytes": sum(frames) * 1024 * 4, "d2h_bytes": sum(frames) * HOP * 2, "scaling": "strong",
           "what": "512 SRT segments U(1,15) s (sched.srt_workload), LPT-sharded, ragged batches <= 4096 frames, pinned host "
                   "fp32 latents -> int16 PCM gathered on rank 0's host; wall clock, max over ranks"}
    if world > 1:
        del lat, timed, first
        single = None
        if rank == 0:
            lat_all = host_latents(list(range(n)), 555)
            # the whole job on rank 0 alone (no barrier inside: the other ranks wait below)
            def alone():
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                res = sched.decode_shard(g, lat_all, emb, list(range(n)), max_batch_frames=4096, max_batch=64, to_host=True, int16=True)
                res.wait()
                return time.perf_counter() - t0
            alone()
            alone()
            single = alone()
        dist.barrier()
        if rank == 0:
            out["single_gpu_wall_ms"] = single * 1e3
            out["strong_eff"] = single / (world * wall)
    else:
        out["strong_eff"] = 1.0
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from b200vgan import synth
    from b200vgan.model import BigVGAN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- b200vgan has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, T = args.batch, synth.frames_for_seconds(args.seconds)
    audio_s_per_step = B * T * HOP / SR
    g = BigVGAN(dict(synth.H_DEFAULT), use_cuda_kernel=True, precision=args.precision)
    sd = synth.make_state_dict(1234)
    g.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    g = g.to(dev)
    g.remove_weight_norm()
    g.eval()
    emb = torch.from_numpy(synth.make_speaker_embedding(B=1)).to(dev)

    nrot = 4   # rotate distinct input batches; the ~GBs of per-step workspace traffic also exceed the 126 MB L2
    lat_host = [torch.from_numpy(synth.make_latents(2, rank * 100 + i, B, T)).pin_memory() for i in range(nrot)]
    lat_dev = [t.to(dev) for t in lat_host]
    mel_host = torch.from_numpy(synth.make_mel(seed=7, Tm=511, B=1)).pin_memory()   # prompt-sized mel [1,511,100]
    wav_host = torch.empty(B, 1, T * HOP, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2, written between timed steps
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(step, warm):
        """W warm-ups, then args.steps steps, each bracketed by its own CUDA-event pair on the launch stream with an
        L2 flush in between (outside the pair); barrier + synchronize on both sides.  Returns device ms (sum), wall s."""
        for i in range(warm):
            step(i)
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        w0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record(stream)
            step(i)
            ev[i][1].record(stream)
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev), time.perf_counter() - w0

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    # ---- device-resident throughput ("value"): no per-launch instrumentation ---------------------
    dev_ms, wall = timed_loop(lambda i: g.forward_with_embedding(lat_dev[i % nrot], emb), max(3, args.warmup))

    # ---- end-to-end through the reference call site, host buffers ("e2e") -------------------------
    #      wav, _ = bigvgan(latent, mel_ref): H2D of latents + prompt mel, ECAPA + decode, D2H of the waveform.
    #      Copies run on their own streams, ordered by events, so the H2D of step i+1 and the D2H of step i-1 overlap
    #      the decode of step i (what a serving loop does); every step's copies are inside the timed region.
    h2d_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    wav_hosts = [wav_host, torch.empty_like(wav_host).pin_memory()]

    class Pipe:
        """Two-deep software pipeline around a decode call: stage(i) uploads step i's inputs on the copy stream,
        run(i) decodes them on the compute stream and hands the result to the download stream."""
        def __init__(self, call):
            self.call, self.inputs, self.drained = call, {}, [None, None]
            self.consumed = [None, None]
            # two device input slots, allocated once: no allocator traffic (and no cross-stream block hand-over) per step
            self.x = [torch.empty_like(lat_dev[0]) for _ in range(2)]
            self.mel = [torch.empty(mel_host.shape, device=dev) for _ in range(2)]

        def stage(self, i):
            k = i & 1
            with torch.cuda.stream(h2d_s):
                if self.consumed[k] is not None:
                    h2d_s.wait_event(self.consumed[k])          # the decode that read this slot two steps ago is done
                self.x[k].copy_(lat_host[i % nrot], non_blocking=True)
                self.mel[k].copy_(mel_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d_s)
            self.inputs[i] = (self.x[k], self.mel[k], ev)

        def run(self, i):
            if i not in self.inputs:
                self.stage(i)
            x, mel, ev = self.inputs.pop(i)
            stream.wait_event(ev)
            self.stage(i + 1)                                   # next step's upload overlaps this decode
            flush.fill_(i & 0xFF)                               # L2 flush between steps (inside the timed region: ~40 us)
            wav = self.call(x, mel)
            done = torch.cuda.Event()
            done.record(stream)
            self.consumed[i & 1] = done
            if self.drained[i & 1] is not None:                 # the host buffer's previous download has finished
                self.drained[i & 1].synchronize()
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(done)
                wav_hosts[i & 1].copy_(wav, non_blocking=True)
                wav.record_stream(d2h_s)
                fin = torch.cuda.Event()
                fin.record(d2h_s)
            self.drained[i & 1] = fin

        def finish(self):
            for f in self.drained:
                if f is not None:
                    stream.wait_event(f)                        # the timed region's end event waits for the last downloads
            self.inputs.clear()

    def piped_loop(call, warm):
        pipe = Pipe(call)
        for i in range(warm):
            pipe.run(i)
        pipe.finish()
        torch.cuda.synchronize(dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pipe.inputs.clear()
        for i in range(args.steps):
            pipe.run(i)
        pipe.finish()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    e2e_ms = piped_loop(lambda x, mel: g(x, mel)[0], 4)

    # ---- same with the speaker embedding cached per prompt (forward_with_embedding) ----------------
    e2e_emb_ms = piped_loop(lambda x, mel: g.forward_with_embedding(x, emb), 4)

    # ---- and strictly serial on one stream (upload, decode, download back to back), for comparison ----
    def e2e_step(i):
        x = lat_host[i % nrot].to(dev, non_blocking=True)
        mel = mel_host.to(dev, non_blocking=True)
        wav, _ = g(x, mel)
        wav_host.copy_(wav, non_blocking=True)
    e2e_serial_ms, _ = timed_loop(e2e_step, 2)
    clocks = sampler.stop()

    # ---- per-kernel-class split: a SEPARATE loop with every launch bracketed by events -------------
    g.profile_enable(True)
    g.profile_read()
    prof_ms, _ = timed_loop(lambda i: g.forward_with_embedding(lat_dev[i % nrot], emb), 1)
    prof = g.profile_read()
    g.profile_enable(False)
    # the warm-up step of that loop was profiled too: normalise by the launches actually recorded
    prof_steps = args.steps + 1

    srt = None
    if not args.no_srt:
        srt = srt_leg(g, emb, dev, world, rank, dist)

    t = torch.tensor([dev_ms, e2e_ms, e2e_emb_ms, e2e_serial_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_emb_ms, e2e_serial_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    if rank == 0:
        pk = peaks()
        value = world * audio_s_per_step * args.steps / (dev_ms / 1e3)
        e2e_value = world * audio_s_per_step * args.steps / (e2e_ms / 1e3)
        e2e_emb_value = world * audio_s_per_step * args.steps / (e2e_emb_ms / 1e3)
        conv = prof["conv_tcgen05"] if prof["conv_tcgen05"]["launches"] else prof["conv_cuda_core"]
        tc = args.precision != "fp32" and prof["conv_tcgen05"]["launches"] > 0
        f32 = args.precision in ("fp32", "fp32tc")
        conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
        peak_tf = pk["bf16_tflops_sustained"]
        act = prof["activation1d"]
        act_gbs = act["bytes"] / (act["ms"] * 1e-3) / 1e9 if act["ms"] > 0 else 0.0
        step_ms_prof = sum(v["ms"] for v in prof.values()) / prof_steps
        traffic = ncu_traffic()
        roofline = {
            "bound": "tensor", "kernel": "conv_umma_kernel (tcgen05 implicit-GEMM conv1d)" if tc else "conv_simt_kernel",
            "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tflops / peak_tf,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
            "traffic": traffic.get("bytes"), "traffic_source": traffic.get("source"),
            "launches_per_step": conv["launches"] / prof_steps, "ms_per_step": conv["ms"] / prof_steps,
            "share_of_step": conv["ms"] / prof_steps / step_ms_prof if step_ms_prof else None,
            "measured_in": "separate instrumented loop (per-launch CUDA events), not the `value` loop",
        }
        roofline_act = {
            "bound": "hbm", "kernel": g.activation_kernel_name(), "achieved": act_gbs, "peak": pk["hbm_gbs"],
            "unit": "GB/s", "frac": act_gbs / pk["hbm_gbs"], "traffic": None,
            "launches_per_step": act["launches"] / prof_steps, "ms_per_step": act["ms"] / prof_steps,
            "share_of_step": act["ms"] / prof_steps / step_ms_prof if step_ms_prof else None,
        }
        # The activation's compute ceiling next to its HBM roofline: MUFU (2 cos per element, 16 lanes per SM per clock) in
        # the tensor-core-FIR modes; the FP32 pipe (31 lane-operations per element, 128 lanes per SM per clock) in fp32 mode.
        sm_mhz = (clocks or {}).get("sm_mhz") or 0
        if act["ms"] > 0 and sm_mhz:
            esize = 4 if f32 else 2
            elems = act["bytes"] / (2.0 * esize)
            ops_per_elem, lanes, pipe = (31.0, 128, "fp32") if f32 else (2.0, 16, "mufu")
            ach = ops_per_elem * elems / (act["ms"] * 1e-3) / 1e12
            peak = 148 * lanes * sm_mhz * 1e6 / 1e12
            roofline_act["compute_pipe"] = {"pipe": pipe, "achieved_Tlaneops": ach, "peak_Tlaneops": peak, "frac": ach / peak,
                                            "note": f"{ops_per_elem:g} {pipe} lane-ops per element x elements / event time vs 148 SMs x {lanes} lanes x median SM clock"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, step_s, cores, kind, what = cpu_decode_rate(CPU_FRAMES, 2, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"B=1 x {CPU_FRAMES} latent frames ({CPU_FRAMES * HOP / SR:.2f} s audio: one item of cfg2), {what}, "
                             f"{cores} threads, mean of 2 after 1 warm-up"}
        launches = g.num_launches([T] * B) * args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32", "fp32tc": "f32 (3 x bf16 tensor-core passes)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": f"cfg2: BigVGAN2 decode batch {B} x {args.seconds:g} s synthetic latents per GPU "
                                   f"(T={T} frames, {B * T * HOP} samples), random-init weights of checkpoints/config.yaml",
                       "precision_mode": args.precision, "l2": "256 MiB flush write between timed steps + 4 rotating input batches",
                       "sharding": "independent utterances per GPU, no data-path collective"},
            "realtime_factor_per_gpu": value / world,
            "conv_tflops_whole_step": world * CONV_FLOP_PER_FRAME * B * T * args.steps / (dev_ms / 1e3) / 1e12,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * T * 1024 * 4 + mel_host.numel() * 4,
                    "d2h_bytes_per_step": B * T * HOP * 4, "ms_per_step": e2e_ms / args.steps,
                    "call": "wav, _ = bigvgan(latent, mel_ref)  (infer.py:458): pinned-host latents + prompt mel -> device, native "
                            "ECAPA speaker encoder + decode, fp32 waveform -> pinned host; uploads / downloads on copy streams "
                            "(two-deep pipeline, every step's copies and a 256 MiB L2 flush inside the timed region)",
                    "serial_one_stream": {"value": world * audio_s_per_step * args.steps / (e2e_serial_ms / 1e3),
                                          "ms_per_step": e2e_serial_ms / args.steps,
                                          "what": "the same call with upload, decode and download back to back on one stream, L2 flush between steps"}},
            "e2e_cached_embedding": {"value": e2e_emb_value, "unit": UNIT, "ms_per_step": e2e_emb_ms / args.steps,
                                     "call": "forward_with_embedding(latent, emb): speaker embedding cached per prompt"},
            "gpu_launches": launches,
            "roofline": roofline, "roofline_activation": roofline_act,
            "kernel_classes_ms_per_step": {k: v["ms"] / prof_steps for k, v in prof.items()},
            "instrumented_ms_per_step": prof_ms / args.steps,
            "srt": srt,
            "cpu_baseline": cpu, "clocks": clocks, "wall_s_timed_region": wall,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else this process or its libraries print
    (e.g. NCCL's version banner, which is written to fd 1 from C) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32", "fp32tc"])
    ap.add_argument("--batch", type=int, default=CFG2_B)
    ap.add_argument("--seconds", type=float, default=CFG2_SEC)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-srt", action="store_true", help="skip the config-3 SRT leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
