#!/usr/bin/env python
"""Benchmark of the NOMAD embedding hot path on B200 (BASELINE.json configs[1]:
"batch embedding: 256 x 4 s synthetic 16 kHz clips, wav2vec2-base NOMAD head").

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port)

One "step" = one pass of the hot path (waveform -> 256-d NOMAD embedding) over one batch of 256 synthetic
4-second clips per GPU.  N > 1 (torchrun): utterances shard across ranks with no data-path collective
(weak scaling: every rank embeds its own 256-clip batch).  Rank 0 prints ONE JSON line.

* ``value``  : utterance-seconds embedded per second, waveforms already resident in HBM, CUDA-event timed.
* ``e2e``    : same metric through ``nomad_b200_embed_host`` with pinned HOST buffers (H2D of the waveforms
               and D2H of the embeddings inside the timed region).
* ``roofline``: the dominant kernel (tcgen05 GEMM) timed in situ with one CUDA-event pair per launch on the
               launching stream during the timed region; achieved = sum(2MNK) / sum(duration), against the
               measured sustained fp16/bf16 tensor peak in MEASURED_PEAKS.json.
* ``cpu_baseline``: the oracle (torch CPU fp32 restatement of the reference arithmetic) on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "utterance-seconds embedded per second"
UNIT = "utt-s/s"
CLIPS = 256
CLIP_SECONDS = 4
SR = 16000
WORKLOAD = "configs[1]: batch embedding, 256 x 4 s synthetic 16 kHz clips, wav2vec2-base + NOMAD head"
CPU_SAMPLE_CLIPS = 16
PAIR_N, PAIR_M = 100_000, 1_000
FALLBACK_PEAK_TFLOPS = 1400.0  # B200_PROFILING.md: sustained ~1.4 PFLOP/s (burst fallback 1590)
# dram__bytes_read.sum + dram__bytes_write.sum per tensor-core GEMM launch, averaged over the 55 GEMM launches of one
# step (ncu, profiles/r01_v11_gemm_dram.csv: 31.4 GB per step; the algorithmic operand + result bytes of those
# launches are 33.7 GB)
GEMM_DRAM_TRAFFIC_BYTES = 571.5e6


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_batch(clips: int, seed: int):
    import torch
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(clips, CLIP_SECONDS * SR, generator=g)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(steps: int, warmup: int, min_seconds: float = 0.0):
    """Oracle port on the host cores: `steps` passes over CPU_SAMPLE_CLIPS x 4 s clips.  -> (utt-s/s, ms/step, cores)"""
    import torch

    from nomad_b200.weights import random_state_dict
    from oracle import w2v_oracle as O  # CPU baseline leg: the one place bench.py may execute oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_state_dict(1234)
    wav = synth_batch(CPU_SAMPLE_CLIPS, 0)
    with torch.no_grad():
        for _ in range(warmup):
            O.embed(sd, wav)
        times = []
        t_all = time.perf_counter()
        i = 0
        while i < steps or (time.perf_counter() - t_all) < min_seconds:
            t0 = time.perf_counter()
            O.embed(sd, wav)
            times.append(time.perf_counter() - t0)
            i += 1
    sec = sum(times) / len(times)
    return CPU_SAMPLE_CLIPS * CLIP_SECONDS / sec, sec * 1e3, cores, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    rate, ms, cores, n = cpu_reference_rate(steps, min(args.warmup, 1))
    sample = (f"{CPU_SAMPLE_CLIPS} x {CLIP_SECONDS} s clips per step (1/16 of the 256-clip batch), oracle port of the "
              f"reference's torch-CPU fp32 path, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": CPU_SAMPLE_CLIPS, "clip_seconds": CLIP_SECONDS,
                   "weights": "random-init(seed=1234)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from nomad_b200 import _lib
    from nomad_b200.dist import init_from_env
    from nomad_b200.engine import Engine
    from nomad_b200.weights import load_state_dict

    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner on the
    # first collective) goes to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = init_from_env("nccl")
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sd, source = load_state_dict(None, 1234)
    eng = Engine(sd, local)
    lib = _lib.load()

    N = CLIP_SECONDS * SR
    wav_cpu = synth_batch(CLIPS, seed=rank)
    off = np.arange(CLIPS + 1, dtype=np.int64) * N
    wav_dev = wav_cpu.reshape(-1).to(dev)
    out = torch.empty((CLIPS, 256), dtype=torch.float32, device=dev)
    wav_pin = wav_cpu.reshape(-1).pin_memory()
    wav_pin_np = wav_pin.numpy()
    emb_pin = torch.empty((CLIPS, 256), dtype=torch.float32).pin_memory()
    emb_pin_np = emb_pin.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + sync; device time via CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step_dev = lambda: eng.embed_packed(wav_dev, off, out)
    step_host = lambda: eng.embed_host(wav_pin_np, off, emb_pin_np)

    for _ in range(max(3, args.warmup)):
        step_dev()
    step_host()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("NOMAD_B200_BENCH_NO_SAMPLER") != "1":
        sampler.start()
        time.sleep(0.3)
    n0 = eng.launch_count()
    total_ms = timed(step_dev, args.steps)
    launches = eng.launch_count() - n0
    e2e_ms = timed(step_host, args.steps)
    # roofline leg: the same K steps again with one CUDA-event pair around every tensor-core GEMM launch (kept out
    # of the `value` region: ~1100 extra event records per 10 steps cost ~5 % of the step)
    lib.nomad_b200_profile_gemm(1)
    prof_ms = timed(step_dev, args.steps)
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_int64()
    lib.nomad_b200_profile_gemm_read(C.byref(g_ms), C.byref(g_fl), C.byref(g_n))
    lib.nomad_b200_profile_gemm(0)
    # second named quantity of BASELINE.json's metric: pairwise distances per second (config 4 shape per GPU:
    # 100 k degraded x 1 k NMR embeddings resident in HBM, matrix materialised + fp64 row means)
    gq = torch.Generator().manual_seed(100 + rank)
    deg_e = torch.nn.functional.normalize(torch.randn(PAIR_N, 256, generator=gq), dim=1).to(dev)
    nmr_e = torch.nn.functional.normalize(torch.randn(PAIR_M, 256, generator=gq), dim=1).to(dev)
    for _ in range(3):
        eng.cdist_mean(deg_e, nmr_e)
    pair_ms = timed(lambda: eng.cdist_mean(deg_e, nmr_e), 20) / 20
    pair_mean_ms = timed(lambda: eng.cdist_mean(deg_e, nmr_e, want_matrix=False), 20) / 20
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)

    utt_s_per_step = CLIPS * CLIP_SECONDS * world
    ms_per_step = total_ms / args.steps
    value = utt_s_per_step / (ms_per_step / 1e3)
    e2e_value = utt_s_per_step / (e2e_ms / args.steps / 1e3)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak = float(json.load(open(peaks_path)).get("bf16_tflops_sustained", FALLBACK_PEAK_TFLOPS))
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
    else:
        peak, peak_src = FALLBACK_PEAK_TFLOPS, "fallback (B200_PROFILING.md sustained figure)"
    achieved = (g_fl.value / 1e12) / (g_ms.value / 1e3) if g_ms.value > 0 else 0.0

    from nomad_b200.weights import flops_embed
    step_flops = CLIPS * flops_embed(N)
    # CPU baseline: rank 0, single-GPU runs only (a reported baseline, not the target)
    cpu = None
    if world == 1:
        cpu_rate, cpu_ms, cores, cpu_n = cpu_reference_rate(1, 1, min_seconds=10.0)
        cpu = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cpu_n} pass(es) over {CPU_SAMPLE_CLIPS} x {CLIP_SECONDS} s clips (1/16 of the batch), "
                         f"oracle torch-CPU fp32, {cores} threads, {cpu_ms:.0f} ms/pass"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu_per_step": CLIPS, "clip_seconds": CLIP_SECONDS,
                   "frames_per_clip": 199, "weights": source, "accumulate": "fp32",
                   "l2": "per-step working set ~6 GB of activations >> 126 MB L2 (no explicit flush needed)",
                   "parallelism": f"dp{world} (utterance-sharded, no data-path collective)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(CLIPS * N * 4),
                "d2h_bytes_per_step": int(CLIPS * 256 * 4), "api": "nomad_b200_embed_host (pinned host buffers)"},
        "gpu_launches": int(launches),
        "step_tflops": step_flops * world / (ms_per_step / 1e3) / 1e12,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_pair_kernel (tcgen05 cta_group::2, all 55 GEMM launches of the step)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": GEMM_DRAM_TRAFFIC_BYTES, "peak_source": peak_src,
                     "launches_timed": int(g_n.value), "kernel_ms_per_step": g_ms.value / args.steps,
                     "kernel_share_of_step": g_ms.value / prof_ms,
                     "profiled_ms_per_step": prof_ms / args.steps},
        "pairwise": {"metric": "pairwise distances per second", "n": PAIR_N, "m": PAIR_M, "n_gpus": world,
                     "value": PAIR_N * PAIR_M * world / (pair_ms / 1e3), "unit": "pairs/s",
                     "write_GBps_per_gpu": PAIR_N * PAIR_M * 4 / (pair_ms / 1e3) / 1e9,
                     "means_only_value": PAIR_N * PAIR_M * world / (pair_mean_ms / 1e3),
                     "roof": "3-pass split-fp16 Gram: 1536 FLOP/pair on the tensor roof, 4 B/pair written on the HBM roof"},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
