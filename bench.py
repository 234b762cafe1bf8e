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


def ncu_dram_bytes(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE captured launch of the dominant kernel (ncu --set full summary
    committed under profiles/; a capture, not a live measurement -- the live numbers are the CUDA-event times)."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, found = 0.0, 0
    try:
        for ln in open(path):
            ln = ln.strip()
            for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
                if ln.startswith(key):
                    u = ln[len(key):ln.index("]")]
                    total += float(ln.split("=")[1].replace(",", "")) * unit.get(u, 1.0)
                    found += 1
    except OSError:
        return {}
    if found != 2:
        return {}
    return {"bytes": total, "source": "ncu --set full, one launch (stage 0, k=11 conv, 768 ch, 15040 rows: 59 MB algorithmic incl. 13 MB weights), "
                                      + os.path.relpath(path, ROOT)}


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
# CPU reference arm / cpu_baseline: the CPU port of the reference path (oracle/bigvgan_torch_cpu.py)
# ------------------------------------------------------------------------------------------------
def cpu_decode_rate(sample_frames: int, repeats: int, warmup: int):
    """Times the CPU port of the reference path (oracle/bigvgan_torch_cpu.py: the reference's own torch
    operators on the host cores, fp32, all threads) on B=1 x sample_frames latent frames.
    Returns (audio-seconds per second, per-step seconds, threads used)."""
    import torch
    from oracle import bigvgan_torch_cpu as TC      # the ONLY place bench.py executes oracle/
    from b200vgan import synth
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, ncpu))
    sd = TC.prepare_state_dict(synth.make_state_dict(1234, with_speaker_encoder=False))
    emb = synth.make_speaker_embedding(B=1)
    times = []
    for i in range(warmup + repeats):
        x = synth.make_latents(2, i, 1, sample_frames)
        t0 = time.perf_counter()
        TC.bigvgan_forward_with_embedding(x, emb, sd)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    audio_s = sample_frames * HOP / SR
    return audio_s * len(times) / sum(times), sum(times) / len(times), torch.get_num_threads()


CPU_PORT = "torch-CPU port of the reference path (same F.conv1d/conv_transpose1d operators, oneDNN, fp32)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 47
    rate, step_s, cores = cpu_decode_rate(frames, max(1, args.steps), max(0, args.warmup))
    sample = f"B=1 x {frames} latent frames ({frames * HOP / SR:.2f} s audio) per step, {CPU_PORT}, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: BigVGAN2 decode batch 16 x 10 s synthetic latents (CPU arm: bounded sample)",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
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
    wav_host = torch.empty(B, 1, T * HOP, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2, written between timed steps
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") --------------------------------------------------
    for i in range(max(3, args.warmup)):
        g.forward_with_embedding(lat_dev[i % nrot], emb)
    torch.cuda.synchronize(dev)
    g.profile_enable(True)
    g.profile_read()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                 # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record(stream)
        g.forward_with_embedding(lat_dev[i % nrot], emb)
        ev[i][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = g.profile_read()
    g.profile_enable(False)

    # ---- end-to-end through the public call with host buffers ("e2e") ---------------------------
    def e2e_step(i):
        x = lat_host[i % nrot].to(dev, non_blocking=True)          # H2D from pinned memory
        wav = g.forward_with_embedding(x, emb)
        wav_host.copy_(wav, non_blocking=True)                     # D2H of the step's result
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev2[i][0].record(stream)
        e2e_step(i)
        ev2[i][1].record(stream)
    barrier()
    e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        value = world * audio_s_per_step * args.steps / (dev_ms / 1e3)
        e2e_value = world * audio_s_per_step * args.steps / (e2e_ms / 1e3)
        conv = prof["conv_tcgen05"] if prof["conv_tcgen05"]["launches"] else prof["conv_cuda_core"]
        tc = args.precision == "bf16" and prof["conv_tcgen05"]["launches"] > 0
        conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
        peak_tf = pk["bf16_tflops_sustained"]
        act = prof["activation1d"]
        act_gbs = act["bytes"] / (act["ms"] * 1e-3) / 1e9 if act["ms"] > 0 else 0.0
        step_ms_prof = sum(v["ms"] for v in prof.values()) / args.steps
        ncu_traffic = ncu_dram_bytes(os.path.join(ROOT, "profiles", "r1_rc5_ncu_full_conv_s0k11.txt"))
        roofline = {
            "bound": "tensor", "kernel": "conv_umma_kernel (tcgen05 implicit-GEMM conv1d)" if tc else "conv_simt_kernel",
            "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tflops / peak_tf,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
            "traffic": ncu_traffic.get("bytes"), "traffic_source": ncu_traffic.get("source"),
            "launches_per_step": conv["launches"] / args.steps, "ms_per_step": conv["ms"] / args.steps,
            "share_of_step": conv["ms"] / args.steps / step_ms_prof if step_ms_prof else None,
        }
        roofline_act = {
            "bound": "hbm", "kernel": ("act1d_c8_mma_kernel (Activation1d: up-FIR and down-FIR as warp-level MMAs, SnakeBeta on CUDA cores)" if args.precision == "bf16" else "act1d_c8_v3_kernel (Activation1d, register-streamed fp32)"), "achieved": act_gbs, "peak": pk["hbm_gbs"],
            "unit": "GB/s", "frac": act_gbs / pk["hbm_gbs"], "traffic": None,
            "launches_per_step": act["launches"] / args.steps, "ms_per_step": act["ms"] / args.steps,
            "share_of_step": act["ms"] / args.steps / step_ms_prof if step_ms_prof else None,
        }
        # The activation is not HBM-bound.  bf16 mode (act1d_c8_mma_kernel: FIRs on the tensor cores): the nearest hardware
        # ceiling is the MUFU rate (2 cos per element, 16 lanes per SM per clock); fp32 mode (register-streamed kernel): the
        # FP32 pipe (31 lane-operations per element, 128 lanes per SM per clock, tools/fma_bench.cu).  Report that ceiling too.
        sm_mhz = (clocks or {}).get("sm_mhz") or 0
        if act["ms"] > 0 and sm_mhz:
            elems = act["bytes"] / (2.0 * (2 if args.precision == "bf16" else 4))
            ops_per_elem, lanes, pipe = (2.0, 16, "mufu") if args.precision == "bf16" else (31.0, 128, "fp32")
            ach = ops_per_elem * elems / (act["ms"] * 1e-3) / 1e12
            peak = 148 * lanes * sm_mhz * 1e6 / 1e12
            roofline_act["compute_pipe"] = {"pipe": pipe, "achieved_Tlaneops": ach, "peak_Tlaneops": peak, "frac": ach / peak,
                                            "note": f"{ops_per_elem:g} {pipe} lane-ops per element x elements / event time vs 148 SMs x {lanes} lanes x median SM clock"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            frames = 47
            rate, step_s, cores = cpu_decode_rate(frames, 4, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"B=1 x {frames} latent frames ({frames * HOP / SR:.2f} s audio), {CPU_PORT}, "
                             f"{cores} threads, mean of 4 after 1 warm-up"}
        launches = g.num_launches([T] * B) * args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": f"cfg2: BigVGAN2 decode batch {B} x {args.seconds:g} s synthetic latents per GPU "
                                   f"(T={T} frames, {B * T * HOP} samples), random-init weights of checkpoints/config.yaml",
                       "precision_mode": args.precision, "l2": "256 MiB flush write between timed steps + 4 rotating input batches",
                       "sharding": "independent utterances per GPU, no data-path collective"},
            "realtime_factor_per_gpu": value / world,
            "conv_tflops_whole_step": world * CONV_FLOP_PER_FRAME * B * T * args.steps / (dev_ms / 1e3) / 1e12,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * T * 1024 * 4,
                    "d2h_bytes_per_step": B * T * HOP * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "roofline": roofline, "roofline_activation": roofline_act,
            "kernel_classes_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
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
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=CFG2_B)
    ap.add_argument("--seconds", type=float, default=CFG2_SEC)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
