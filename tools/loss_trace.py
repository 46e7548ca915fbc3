"""Kernel timeline of the NOMAD loss step (BASELINE config 3: 32 x 2 s estimate/clean pairs, forward + backward)."""
import collections, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
B, N = 32, 32000
eng = Engine(random_state_dict(1234), 0)
g = torch.Generator().manual_seed(0)
eng.set_loss_head(torch.randn(256, 768, generator=g) * 0.03, torch.zeros(256))
est = (0.1 * torch.randn(B, N, generator=g)).cuda()
clean = (0.1 * torch.randn(B, N, generator=g)).cuda()
for _ in range(3):
    eng.loss_fwd_bwd(est, clean, 0.1, True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        eng.loss_fwd_bwd(est, clean, 0.1, True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict(); busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    k = e.name[:60]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += d; busy += d
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"span {span/3/1e3:.2f} ms/step, busy {busy/3/1e3:.2f} ms/step, launches/step {len(evs)/3:.0f}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{t/3/1e3:8.3f} ms/step  n/step={n/3:5.1f}  {k}")
