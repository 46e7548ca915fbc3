"""Kernel timeline of the triplet fine-tuning step at the reference's training shape (8 triplets x 10 s)."""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.triplet import triplet_loss_and_grads
from nomad_b200.weights import random_state_dict
B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 160000
sd = random_state_dict(1234)
eng = Engine(sd, 0)
g = torch.Generator().manual_seed(0)
A, P, Nn = ((0.1 * torch.randn(B, N, generator=g)).cuda() for _ in range(3))
for _ in range(2):
    triplet_loss_and_grads(eng, sd, A, P, Nn)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        triplet_loss_and_grads(eng, sd, A, P, Nn)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict(); busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    k = e.name[:70]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += d; busy += d
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"span {span/2/1e3:.2f} ms/step, busy {busy/2/1e3:.2f} ms/step, launches/step {len(evs)/2:.0f}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{t/2/1e3:8.3f} ms/step  n/step={n/2:5.1f}  {k}")
