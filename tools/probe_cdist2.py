"""cdist timing at the BASELINE configs[4] sizes (matrix materialised / means only)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
eng = Engine(random_state_dict(1234), 0)
g = torch.Generator(device="cuda").manual_seed(0)
for n, m in ((100_000, 1000), (100_000, 8192), (1_000_000, 128), (100_000, 2048)):
    a = torch.nn.functional.normalize(torch.randn(n, 256, device="cuda", generator=g), dim=1)
    b = torch.nn.functional.normalize(torch.randn(m, 256, device="cuda", generator=g), dim=1)
    for wm in (True, False):
        for _ in range(3):
            eng.cdist_mean(a, b, want_matrix=wm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.cdist_mean(a, b, want_matrix=wm)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"cdist {n} x {m} matrix={wm}: {ms:.3f} ms  {n*m/ms/1e9:.1f} Gpairs/s  write {n*m*4/ms/1e9 if wm else 0:.2f} TB/s  tensor {1536.0*n*m/ms/1e9:.0f} TFLOP/s")
