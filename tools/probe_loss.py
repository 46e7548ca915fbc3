"""GPU probe: NOMAD loss value + d loss / d estimate vs the golden reference run and the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine  # noqa: E402
from nomad_b200.weights import random_state_dict  # noqa: E402
from oracle import w2v_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
sd = random_state_dict(1234)
eng = Engine(sd, 0)
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_loss.npz"))
hw, hb = torch.from_numpy(g["head_w"]), torch.from_numpy(g["head_b"])
eng.set_loss_head(hw, hb)
est, clean = torch.from_numpy(g["est"]).cuda(), torch.from_numpy(g["clean"]).cuda()
for impl in (1, 0):
    eng.set_gemm_impl(impl)
    for fgm in (1.0, 0.1):
        loss, grad = eng.loss_fwd_bwd(est, clean, fgm, with_grad=True)
        torch.cuda.synchronize()
        ref_l, ref_g = float(g[f"loss_fgm{fgm}"]), g[f"grad_fgm{fgm}"].reshape(2, -1)
        gg = grad.cpu().numpy()
        cos = float((gg * ref_g).sum() / (np.linalg.norm(gg) * np.linalg.norm(ref_g) + 1e-30))
        print(f"impl={impl} fgm={fgm}: loss {loss.item():.6f} ref {ref_l:.6f} rel {abs(loss.item()-ref_l)/ref_l:.2e} | "
              f"grad max|ref| {np.abs(ref_g).max():.3e} max err {np.abs(gg-ref_g).max():.3e} cos {cos:.6f} "
              f"norm ratio {np.linalg.norm(gg)/np.linalg.norm(ref_g):.4f} nan={np.isnan(gg).any()}", flush=True)
    l2, _ = eng.loss_fwd_bwd(est, clean, 1.0, with_grad=False)
    print(f"impl={impl} forward-only loss {l2.item():.6f}", flush=True)
# BASELINE config 4: 32 x 2 s pairs
eng.set_gemm_impl(0)
gen = torch.Generator().manual_seed(0)
e4 = (0.1 * torch.randn(32, 1, 32000, generator=gen)).cuda()
c4 = (0.1 * torch.randn(32, 1, 32000, generator=gen)).cuda()
for _ in range(2):
    loss, grad = eng.loss_fwd_bwd(e4, c4, 0.1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    loss, grad = eng.loss_fwd_bwd(e4, c4, 0.1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"C4 loss fwd+bwd 32 x 2 s pairs: {ms:.2f} ms/step -> {32/ms*1e3:.0f} pairs/s; loss {loss.item():.5f} grad finite {torch.isfinite(grad).all().item()}", flush=True)
# oracle check on a subset (4 pairs) at full length
with torch.enable_grad():
    es = e4[:4].cpu().clone().requires_grad_(True)
    lo = O.nomad_forward(sd, hw, hb, es, c4[:4].cpu(), feature_grad_mult=0.1)
    lo.backward()
l4, g4 = eng.loss_fwd_bwd(e4[:4], c4[:4], 0.1)
rg = es.grad.reshape(4, -1).numpy()
gg = g4.cpu().numpy()
cos = float((gg * rg).sum() / (np.linalg.norm(gg) * np.linalg.norm(rg)))
print(f"4 x 2 s pairs vs oracle: loss {l4.item():.6f} ref {lo.item():.6f}; grad cos {cos:.6f} max err {np.abs(gg-rg).max():.3e} max|ref| {np.abs(rg).max():.3e}", flush=True)
