"""In-situ kernel timeline of the bench step (CUPTI via torch.profiler): per-kernel totals, gaps, host enqueue time."""
import os, sys, time, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
B, N = 256, 64000
eng = Engine(random_state_dict(1234), 0)
gen = torch.Generator().manual_seed(0)
wav = (0.1 * torch.randn(B * N, generator=gen)).cuda()
off = np.arange(B + 1, dtype=np.int64) * N
out = torch.empty(B, 256, device="cuda")
for _ in range(5):
    eng.embed_packed(wav, off, out)
torch.cuda.synchronize()
t0 = time.perf_counter(); eng.embed_packed(wav, off, out); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0):.2f} ms, until done {1e3*(t2-t0):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        eng.embed_packed(wav, off, out)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
print("cuda events:", len(evs))
agg = collections.OrderedDict()
busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    k = e.name[:48]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += d; busy += d
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"span {span/3/1e3:.2f} ms/step, busy {busy/3/1e3:.2f} ms/step, idle {100*(1-busy/span):.1f}%")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t/3/1e3:8.3f} ms/step  n/step={n/3:5.1f}  {k}")
gaps = [evs[i+1].time_range.start - evs[i].time_range.end for i in range(len(evs)-1)]
gaps = [g for g in gaps if g < 1000]
print(f"median gap {np.median(gaps):.2f} us, mean {np.mean(gaps):.2f} us, sum/step {sum(gaps)/3/1e3:.2f} ms")
# per-launch GEMM durations of the last profiled step, labelled by their position in the launch sequence
gem = [e for e in evs if "gemm_tc" in e.name]
per = len(gem) // 3
last = gem[-per:]
labels = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "proj"] + [f"{n}.{l}" for l in range(12) for n in ("qkv", "out", "fc1", "fc2")]
FL = {"conv1": 2 * 256 * 6400 * 512 * 1536, "conv2": 2 * 256 * 3200 * 512 * 1536, "conv3": 2 * 256 * 1600 * 512 * 1536,
      "conv4": 2 * 256 * 800 * 512 * 1536, "conv5": 2 * 256 * 400 * 512 * 1024, "conv6": 2 * 256 * 200 * 512 * 1024,
      "proj": 2 * 51200 * 768 * 512, "qkv": 2 * 51200 * 2304 * 768, "out": 2 * 51200 * 768 * 768,
      "fc1": 2 * 51200 * 3072 * 768, "fc2": 2 * 51200 * 768 * 3072}
if len(last) == len(labels):
    agg2 = collections.OrderedDict()
    for lab, e in zip(labels, last):
        k = lab.split(".")[0]
        d = (e.time_range.end - e.time_range.start)
        agg2.setdefault(k, []).append(d)
    for k, v in agg2.items():
        us = float(np.mean(v))
        print(f"  gemm {k:6s} n={len(v):2d}  {us:8.1f} us/launch  {FL[k] / us / 1e6:7.1f} TFLOP/s")
