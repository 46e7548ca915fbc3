import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
sd = random_state_dict(1234)
eng = Engine(sd, 0)
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_loss.npz"))
eng.set_loss_head(torch.from_numpy(g["head_w"]), torch.from_numpy(g["head_b"]))
gen = torch.Generator().manual_seed(5)
est = (0.1 * torch.randn(4, 1, 16384, generator=gen)).cuda()
loss, grad = eng.loss_fwd_bwd(est, est.clone(), 0.1)
print("loss(x,x) =", loss.item(), "grad max", grad.abs().max().item(), "nan", torch.isnan(grad).any().item(), "nnz", (grad != 0).sum().item())
layers, emb = eng.layers(torch.cat([est, est]).squeeze(1))
print("layer diffs between identical halves:", [(layers[l][:4] - layers[l][4:]).abs().max().item() for l in range(12)], (emb[:4]-emb[4:]).abs().max().item())
