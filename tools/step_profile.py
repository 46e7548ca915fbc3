"""One warm-up + N timed embed steps of the bench workload (256 x 4 s); meant to run under ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine  # noqa: E402
from nomad_b200.weights import random_state_dict  # noqa: E402

B = int(os.environ.get("B", 256))
N = int(os.environ.get("N", 64000))
steps = int(os.environ.get("STEPS", 1))
eng = Engine(random_state_dict(1234), 0)
gen = torch.Generator().manual_seed(0)
wav = (0.1 * torch.randn(B * N, generator=gen)).cuda()
off = np.arange(B + 1, dtype=np.int64) * N
out = torch.empty(B, 256, device="cuda")
eng.embed_packed(wav, off, out)
torch.cuda.synchronize()
n0 = eng.launch_count()
for _ in range(steps):
    eng.embed_packed(wav, off, out)
torch.cuda.synchronize()
print("launches per step:", (eng.launch_count() - n0) // steps, "first emb:", out[0, :4].tolist())
