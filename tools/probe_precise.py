"""fp32-class mode diagnostics: (1) error of the split-operand tensor-core GEMM against fp64 as a function of K (is the
TMEM accumulation the limit?), (2) per-layer error growth of the fp32-mode forward against the fp64 oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200 import _lib
from nomad_b200.engine import Engine
from nomad_b200.weights import random_state_dict
from oracle import w2v_oracle as O
lib = _lib.load()
dev = torch.device("cuda:0")


def split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi, lo


g = torch.Generator().manual_seed(0)
for (M, N, K) in ((2048, 768, 512), (2048, 768, 768), (2048, 768, 1536), (2048, 768, 3072), (2048, 3072, 768), (2048, 48, 6144)):
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.04).to(dev)
    sc = 2.0 ** np.floor(np.log2(16000.0 / float(w.abs().max())))
    ah, al = split(a)
    bh, bl = split(w * sc)
    c = torch.empty(M, N, device=dev)
    _lib.check(lib.nomad_b200_gemm_split(ah.data_ptr(), al.data_ptr(), K, bh.data_ptr(), bl.data_ptr(), M, N, K, float(1.0 / sc),
                                         None, c.data_ptr(), None, None, N, 8, torch.cuda.current_stream().cuda_stream), "gemm_split")
    torch.cuda.synchronize()
    a_eff = ah.double() + al.double()
    b_eff = (bh.double() + bl.double()) / sc
    ref_eff = a_eff @ b_eff.T                       # what a perfect accumulator would give for the operands as stored
    ref_full = a.double() @ w.double().T            # the true product
    drop = (al.double() @ (bl.double() / sc).T)     # the lo x lo term the kernel leaves out
    f32 = (a @ w.T)                                 # cuBLAS fp32 (TF32 off by default)
    s = float(ref_full.abs().max())
    print(f"M{M} N{N} K{K}: |out|max {s:.2f}  kernel-vs-stored-operands {float((c.double() - ref_eff).abs().max()):.3e}  "
          f"mean-signed {float((c.double() - ref_eff).mean()):+.2e}  rel-bias {float(((c.double() - ref_eff) * ref_eff.sign()).mean() / ref_eff.abs().mean()):+.2e}  "
          f"operand-split {float((ref_eff - ref_full).abs().max()):.3e}  lo*lo {float(drop.abs().max()):.3e}  "
          f"kernel-vs-true {float((c.double() - ref_full).abs().max()):.3e}  cublas-fp32-vs-true {float((f32.double() - ref_full).abs().max()):.3e}")

sd = random_state_dict(1234)
eng = Engine(sd, 0, precision="fp32")
gen = torch.Generator().manual_seed(3)
wav = 0.1 * torch.randn(3, 32000, generator=gen)
layers, emb = eng.layers(wav.cuda())
with torch.no_grad():
    ref = O.ssl_layers(sd, wav, dtype=torch.float64)
    r32 = O.ssl_layers(sd, wav, dtype=torch.float32)
    e64 = O.embed(sd, wav, dtype=torch.float64)
for l in range(12):
    print(f"layer {l:2d}: ours-vs-fp64 {float((layers[l].cpu().double() - ref[l]).abs().max()):.3e}   torch-fp32-vs-fp64 "
          f"{float((r32[l].double() - ref[l]).abs().max()):.3e}   |x|max {float(ref[l].abs().max()):.2f}")
print(f"embedding: ours-vs-fp64 {float((emb.cpu().double() - e64).abs().max()):.3e}")
eng.set_precision("fp16")
l16, e16 = eng.layers(wav.cuda())
print(f"fp16 mode: layer 11 {float((l16[11].cpu().double() - ref[11]).abs().max()):.3e}  embedding {float((e16.cpu().double() - e64).abs().max()):.3e}")
