"""Timeline of one host-buffer embed call (nomad_b200_embed_host): where the H2D copies sit relative to the kernels."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
B, N = 256, 64000
eng = Engine(random_state_dict(1234), 0)
wav = (0.1 * torch.randn(B * N, generator=torch.Generator().manual_seed(0))).pin_memory()
wav_np = wav.numpy()
off = np.arange(B + 1, dtype=np.int64) * N
out = np.empty((B, 256), dtype=np.float32)
for _ in range(4):
    eng.embed_host(wav_np, off, out)
t0 = time.perf_counter()
for _ in range(5):
    eng.embed_host(wav_np, off, out)
print(f"wall per call {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.embed_host(wav_np, off, out)
    eng.embed_host(wav_np, off, out)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t00 = evs[0].time_range.start
half = len(evs) // 2
for e in evs[half:half + 12] + evs[-4:]:
    print(f"{(e.time_range.start - t00) / 1e3:9.3f} ms  +{(e.time_range.end - e.time_range.start):8.1f} us  {e.name[:60]}")
first, last = evs[half], evs[-1]
print(f"second call: first GPU activity -> last: {(last.time_range.end - first.time_range.start) / 1e3:.2f} ms; "
      f"gap between calls {(evs[half].time_range.start - evs[half - 1].time_range.end):.0f} us")
