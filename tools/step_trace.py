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
