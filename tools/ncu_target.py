"""Short ncu target: a few passes of the bench workload (256 x 4 s) through the C ABI, nothing else.
usage: ncu ... python tools/ncu_target.py [steps] [clips]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = 64000
eng = Engine(random_state_dict(1234), 0)
gen = torch.Generator().manual_seed(0)
wav = (0.1 * torch.randn(B * N, generator=gen)).cuda()
off = np.arange(B + 1, dtype=np.int64) * N
out = torch.empty(B, 256, device="cuda")
for _ in range(steps):
    eng.embed_packed(wav, off, out)
torch.cuda.synchronize()
print("done", float(out.abs().sum()))
