"""GPU probe: distance-matrix sweep (BASELINE config 5): tensor-core Gram kernel vs fp32 direct kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
eng = Engine(random_state_dict(1234), 0)
gen = torch.Generator(device="cuda").manual_seed(0)
def unit(n):
    return torch.nn.functional.normalize(torch.randn(n, 256, device="cuda", generator=gen), dim=1)
for n, m in ((1000, 128), (10000, 1024), (100000, 1000), (100000, 8192), (1000000, 2048)):
    a, b = unit(n), unit(m)
    a[7] = b[3]                      # exact duplicate
    a[11] = torch.nn.functional.normalize(b[5] + 1e-3 * a[11], dim=0)  # near duplicate
    idx = torch.tensor([0, 7, 11, n // 2, n - 1], device="cuda")
    ref = torch.cdist(a[idx].double(), b.double())
    for impl, name in ((0, "tcgen05"), (1, "fp32-direct")):
        if impl == 1 and n * m > 3e9:
            continue
        for want in (True, False):
            dm, mean = eng.cdist_mean(a, b, want_matrix=want, gemm_impl=impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):  # caller-owned outputs: a fresh 3 GB torch.empty per call put cudaMalloc inside the timed loop
                dm, mean = eng.cdist_mean(a, b, want_matrix=want, gemm_impl=impl, out_dm=dm, out_mean=mean)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            err = float((dm[idx].double() - ref).abs().max()) if want else float("nan")
            merr = float((mean[idx] - ref.mean(1)).abs().max())
            print(f"n={n} m={m} {name:11s} matrix={int(want)}: {ms:8.3f} ms  {n*m/ms/1e6:9.1f} Gpairs/s  "
                  f"write {n*m*4/ms/1e6 if want else 0:7.1f} GB/s  max err {err:.2e} mean err {merr:.2e} d(dup)={float(dm[7,3]) if want else -1:.1e}", flush=True)
        del dm
