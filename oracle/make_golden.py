"""Generate the committed golden fixtures under ``tests/golden/`` by running the
reference's own ``nomad.py`` (verbatim, under ``oracle/ref_shims.py``).

Run in the build container only:  ``python oracle/make_golden.py``
(needs ``/root/reference``).  TEST INFRASTRUCTURE ONLY.

Weights: ``nomad_b200.weights.random_state_dict(1234)`` (real checkpoint is not
shipped and there is no network).  Fixtures:

* ``wavs/{nmr-data,test-data}/*.wav``  the reference's bundled example inputs
  (``data/nmr-data``, ``data/test-data``), copied byte for byte;
* ``ref_predict.npz``   ``Nomad.predict('dir', ...)`` on them: embeddings,
  unrounded cdist matrix, both returned DataFrames and both CSV texts;
* ``ref_small.npz``     ``TripletModel`` embeddings + all 12 layer outputs on
  short synthetic clips, batch and per-file;
* ``ref_loss.npz``      ``Nomad.forward`` value and d loss/d estimate on a small
  pair batch, with the freshly initialised loss head that produced them;
* ``ref_cdist.npz``     scipy ``cdist`` + row mean on seeded embeddings.
"""
from __future__ import annotations

import io
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nomad_b200.weights import random_state_dict  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLD, exist_ok=True)
    # 1. bundled wavs
    for sub in ("nmr-data", "test-data"):
        dst = os.path.join(GOLD, "wavs", sub)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(ref_shims.REFERENCE_ROOT, "data", sub))):
            shutil.copyfile(os.path.join(ref_shims.REFERENCE_ROOT, "data", sub, f), os.path.join(dst, f))

    sd = random_state_dict(1234)
    work = tempfile.mkdtemp(prefix="nomad_ref_")
    mod, ref = ref_shims.load_reference_nomad(sd, work, feature_grad_mult=1.0)

    # 2. predict('dir') on the bundled wavs (listing order = os.listdir of the fixture dirs)
    nmr_dir = os.path.join(GOLD, "wavs", "nmr-data")
    deg_dir = os.path.join(GOLD, "wavs", "test-data")
    res_dir = os.path.join(work, "res")
    os.makedirs(res_dir)
    nmr_emb = ref.get_embeddings(nmr_dir).set_index("filename")
    deg_emb = ref.get_embeddings(deg_dir).set_index("filename")
    df_avg, df_dm = ref.predict("dir", nmr_dir, deg_dir, res_dir)
    from scipy.spatial.distance import cdist
    dm = cdist(deg_emb, nmr_emb)
    np.savez_compressed(
        os.path.join(GOLD, "ref_predict.npz"),
        nmr_files=np.array([os.path.basename(x) for x in nmr_emb.index]),
        deg_files=np.array([os.path.basename(x) for x in deg_emb.index]),
        nmr_emb=nmr_emb.to_numpy(dtype=np.float32), deg_emb=deg_emb.to_numpy(dtype=np.float32),
        dm=dm, avg=dm.mean(axis=1),
        df_avg_index=np.array([str(x) for x in df_avg.index]), df_avg_values=df_avg["NOMAD"].to_numpy(),
        df_dm_index=np.array([str(x) for x in df_dm.index]), df_dm_columns=np.array([str(x) for x in df_dm.columns]), df_dm_values=df_dm.to_numpy(),
        csv_avg=np.array(open(os.path.join(res_dir, "nomad_avg.csv")).read()),
        csv_scores=np.array(open(os.path.join(res_dir, "nomad_scores.csv")).read()),
    )

    # 3. small synthetic clips: batch (equal length) + per-file variable length
    g = torch.Generator().manual_seed(0)
    wav_b = 0.1 * torch.randn(3, 4000, generator=g)
    with torch.no_grad():
        res = ref.model.ssl_model(wav_b, mask=False, features_only=True)
        layers_b = torch.stack([x[0].permute(1, 0, 2) for x in res["layer_results"]])  # (12,B,T,768)
        emb_b = ref.model(wav_b.unsqueeze(1))
        lens = [400, 401, 719, 720, 1000, 2345, 5000, 6789]
        wav_v = [0.1 * torch.randn(n, generator=g) for n in lens]
        emb_v = torch.cat([ref.model(w.reshape(1, 1, -1)) for w in wav_v], 0)
    np.savez_compressed(
        os.path.join(GOLD, "ref_small.npz"),
        wav_b=wav_b.numpy(), layers_b=layers_b.numpy(), emb_b=emb_b.numpy(),
        lens=np.array(lens), wav_v=np.concatenate([w.numpy() for w in wav_v]), emb_v=emb_v.numpy(),
    )

    # 4. loss forward/backward (HEAD semantics: both streams recorded; grad wrt estimate)
    est = (0.1 * torch.randn(2, 1, 4000, generator=g)).requires_grad_(True)
    clean = 0.1 * torch.randn(2, 1, 4000, generator=g)
    out = {}
    for fgm in (1.0, 0.1):
        ref.model.ssl_model.feature_grad_mult = fgm
        if est.grad is not None:
            est.grad = None
        loss = ref.forward(est, clean)
        loss.backward()
        out[f"loss_fgm{fgm}"] = loss.detach().numpy()
        out[f"grad_fgm{fgm}"] = est.grad.detach().numpy().copy()
    ref.model.ssl_model.feature_grad_mult = 1.0
    np.savez_compressed(
        os.path.join(GOLD, "ref_loss.npz"),
        est=est.detach().numpy(), clean=clean.numpy(),
        head_w=ref.lossnet_layers.embedding_layer[1].weight.detach().numpy(),
        head_b=ref.lossnet_layers.embedding_layer[1].bias.detach().numpy(), **out)

    # 5. cdist + row mean
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 256)).astype(np.float32)
    b = rng.standard_normal((19, 256)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    a[3] = b[5] + 1e-3 * a[3]  # a near-duplicate pair: small distance, cancellation-prone
    a[3] /= np.linalg.norm(a[3])
    dmc = cdist(a, b)
    np.savez_compressed(os.path.join(GOLD, "ref_cdist.npz"), a=a, b=b, dm=dmc, avg=np.mean(dmc, axis=1))
    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        p = os.path.join(GOLD, f)
        if os.path.isfile(p):
            print(f"  {f}: {os.path.getsize(p)} bytes")


if __name__ == "__main__":
    main()
