"""GPU probe: CUDA path vs the CPU oracle, per layer, for both GEMM back ends."""
import os
import sys
import time

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
gold = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(gold, "ref_small.npz"))

for impl in (1, 0):
    eng.set_gemm_impl(impl)
    wav = torch.from_numpy(g["wav_b"]).cuda()
    layers, emb = eng.layers(wav)
    torch.cuda.synchronize()
    layers = layers.cpu().numpy()
    ref = g["layers_b"]
    errs = [float(np.abs(layers[l] - ref[l]).max()) for l in range(12)]
    print(f"impl={impl} layer max-abs errs:", " ".join(f"{e:.2e}" for e in errs), flush=True)
    print(f"impl={impl} |ref| max {np.abs(ref[-1]).max():.3f}  nan={np.isnan(layers).any()}", flush=True)
    e = eng.embed([torch.from_numpy(g["wav_b"][i]) for i in range(3)]).cpu().numpy()
    print(f"impl={impl} emb (batch) max-abs err {np.abs(e - g['emb_b']).max():.3e}", flush=True)
    lens = g["lens"].tolist()
    flat = torch.from_numpy(g["wav_v"])
    waves, o = [], 0
    for n in lens:
        waves.append(flat[o:o + n]); o += n
    ev = eng.embed(waves).cpu().numpy()
    print(f"impl={impl} emb (varlen {lens}) per-utt max-abs err:",
          " ".join(f"{x:.2e}" for x in np.abs(ev - g['emb_v']).max(axis=1)), flush=True)

# bundled wavs
import wave
def load(p):
    with wave.open(p, "rb") as w:
        return torch.from_numpy(np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768.0)
gp = np.load(os.path.join(gold, "ref_predict.npz"))
nmr = [load(os.path.join(gold, "wavs", "nmr-data", f)) for f in gp["nmr_files"]]
deg = [load(os.path.join(gold, "wavs", "test-data", f)) for f in gp["deg_files"]]
e_all = eng.embed(nmr + deg)
e_n, e_d = e_all[:4], e_all[4:]
print("bundled emb err nmr %.3e deg %.3e" % (np.abs(e_n.cpu().numpy() - gp["nmr_emb"]).max(), np.abs(e_d.cpu().numpy() - gp["deg_emb"]).max()), flush=True)
dm, mean = eng.cdist_mean(e_d, e_n)
print("bundled dm err %.3e avg err %.3e" % (np.abs(dm.cpu().numpy() - gp["dm"]).max(), np.abs(mean.cpu().numpy() - gp["avg"]).max()))
print("dm\n", dm.cpu().numpy(), "\nref\n", gp["dm"], flush=True)
gc = np.load(os.path.join(gold, "ref_cdist.npz"))
dm, mean = eng.cdist_mean(torch.from_numpy(gc["a"]).cuda(), torch.from_numpy(gc["b"]).cuda())
print("cdist fixture: dm err %.3e mean err %.3e" % (np.abs(dm.cpu().numpy() - gc["dm"]).max(), np.abs(mean.cpu().numpy() - gc["avg"]).max()), flush=True)

# timing: 256 x 4 s
eng.set_gemm_impl(0)
B, N = 256, 64000
gen = torch.Generator().manual_seed(0)
wav = (0.1 * torch.randn(B * N, generator=gen)).cuda()
off = np.arange(B + 1, dtype=np.int64) * N
out = torch.empty(B, 256, device="cuda")
for _ in range(3):
    eng.embed_packed(wav, off, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.embed_packed(wav, off, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"embed 256x4s: {ms:.2f} ms/step -> {B * 4 / ms * 1e3:.0f} utt-s/s, {B * O.flops_embed(N) / ms / 1e9:.1f} TFLOP/s", flush=True)
# oracle check of a few utterances of the big batch
with torch.no_grad():
    ref = O.embed(sd, wav[: 2 * N].cpu().reshape(2, N)).numpy()
print("big-batch emb err (first 2 utts): %.3e" % np.abs(out[:2].cpu().numpy() - ref).max(), flush=True)
