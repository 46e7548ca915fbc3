"""Timings of the BASELINE.json configurations other than the bench.py headline (configs[2..4]) on ONE B200.
They are parity-test cases, not bench lines; this script records what they cost so DESIGN.md / profiles can quote
them.  CUDA-event timing, warm-up first, synthetic inputs (seeded), random-init weights (seed 1234).

    python tools/bench_configs.py [c3] [c4] [c5]      (default: all)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine  # noqa: E402
from nomad_b200.nomad import plan_batches  # noqa: E402
from nomad_b200.weights import flops_embed, random_state_dict  # noqa: E402


def ev_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def c1(eng):
    """configs[0]: nomad.predict('dir') on the bundled wavs (2 degraded vs 4 NMR), through the drop-in API; the oracle
    port of the reference arithmetic on the host cores next to it."""
    import tempfile
    import wave

    from nomad_b200.nomad import Nomad
    from nomad_b200.weights import random_state_dict as rsd
    from oracle import w2v_oracle as O
    nmr = os.path.join(ROOT, "tests", "golden", "wavs", "nmr-data")
    deg = os.path.join(ROOT, "tests", "golden", "wavs", "test-data")
    sd = rsd(1234)
    nomad = Nomad(state_dict=sd)
    with tempfile.TemporaryDirectory() as td:
        nomad.predict("dir", nmr, deg, td)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            nomad.predict("dir", nmr, deg, td)
        torch.cuda.synchronize()
        ours = (time.perf_counter() - t0) / 5
    waves, secs = [], 0.0
    for d in (nmr, deg):
        for f in os.listdir(d):
            with wave.open(os.path.join(d, f), "rb") as w:
                pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768.0
            waves.append(torch.from_numpy(pcm))
            secs += len(pcm) / 16000.0
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        e = O.embed_each(sd, waves).numpy()
    O.cdist_mean(e[4:], e[:4])
    cpu = time.perf_counter() - t0
    return {"config": "c1: nomad.predict('dir') on the bundled wavs (6 files, %.1f utt-s), wall clock incl. file reads and CSV "
                      "writes" % secs, "ours_s": ours, "oracle_cpu_s": cpu, "cores": os.cpu_count(), "speedup": cpu / ours}


def lib(eng, B=256, N=64000):
    """the oracle's plain torch ops on the SAME B200 (cuDNN convs, cuBLAS GEMMs, eager attention) -- the
    "library-kernel Blackwell path" -- on the bench workload, fp32 with TF32 tensor cores and fp16 autocast."""
    from nomad_b200.weights import random_state_dict as rsd
    from oracle import w2v_oracle as O
    sd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in rsd(1234).items()}
    g = torch.Generator().manual_seed(0)
    wav = (0.1 * torch.randn(B, N, generator=g)).cuda()
    out = {"config": "lib: oracle torch ops on cuda:0 (cuDNN / cuBLAS eager), %d x %.0f s" % (B, N / 16000)}
    ref = None
    for name, tf32, amp in (("fp32_tf32", True, False), ("fp16_autocast", True, True), ("fp32_ieee", False, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                return O.embed(sd, wav)
        try:
            ms = ev_time(run, 3, warm=2)
            e = run().float()
            if name == "fp32_ieee":
                ref = e
            out[name] = {"ms": ms, "utt_s_per_s": B * N / 16000 / (ms / 1e3), "tflops": B * flops_embed(N) / (ms / 1e3) / 1e12}
            out[name + "_emb"] = e.cpu()
        except Exception as ex:  # e.g. out of memory at this batch size
            out[name] = {"error": str(ex)[:200]}
    if ref is not None:
        for name in ("fp32_tf32", "fp16_autocast"):
            if name + "_emb" in out:
                out[name]["max_abs_err_vs_fp32"] = float((out[name + "_emb"].cuda() - ref).abs().max())
    ours = eng.embed_packed(wav.reshape(-1), np.arange(B + 1, dtype=np.int64) * N)
    if ref is not None:
        out["nomad_b200_max_abs_err_vs_fp32"] = float((ours - ref).abs().max())
    for k in [k for k in out if k.endswith("_emb")]:
        del out[k]
    return out


def c3(eng, n_utts=1250):
    """corpus scoring slice: n_utts variable-length (1-20 s) utterances, length-bucketed batches, one GPU.
    (100 k utterances on 8 GPUs = 12.5 k per GPU; this is a 1/10 sample of one GPU's share.)"""
    rng = np.random.default_rng(0)
    dur = rng.uniform(1.0, 20.0, size=n_utts)
    lens = (16000 * dur).astype(np.int64)
    g = torch.Generator().manual_seed(0)
    wav = (0.1 * torch.randn(int(lens.sum()), generator=g)).cuda()
    off = np.zeros(n_utts + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    batches = plan_batches(lens.tolist(), 256 * 64000)
    out = torch.empty(n_utts, 256, device="cuda")

    def run():
        for idx in batches:
            ls = [int(lens[i]) for i in idx]
            w = torch.cat([wav[off[i]:off[i + 1]] for i in idx])
            o = np.zeros(len(idx) + 1, dtype=np.int64)
            np.cumsum(np.asarray(ls, dtype=np.int64), out=o[1:])
            out[torch.as_tensor(idx, device="cuda")] = eng.embed_packed(w, o)
    ms = ev_time(run, 2, warm=1)
    utt_s = float(lens.sum()) / 16000.0
    fl = sum(flops_embed(int(n)) for n in lens)
    return {"config": "c3 slice: %d utts, 1-20 s, masked varlen, 1 GPU" % n_utts, "batches": len(batches),
            "ms": ms, "utt_s": utt_s, "utt_s_per_s": utt_s / (ms / 1e3), "tflops": fl / (ms / 1e3) / 1e12}


def c4(eng, B=32, N=32000):
    g = torch.Generator().manual_seed(0)
    eng.set_loss_head(torch.randn(256, 768, generator=g) * 0.03, torch.zeros(256))  # LossNetLayers' fresh nn.Linear
    est = (0.1 * torch.randn(B, N, generator=g)).cuda()
    clean = (0.1 * torch.randn(B, N, generator=g)).cuda()
    ms = ev_time(lambda: eng.loss_fwd_bwd(est, clean, 0.1, True), 10)
    ms_f = ev_time(lambda: eng.loss_fwd_bwd(est, clean, 0.1, False), 10)
    fl = 3.0 * B * flops_embed(N)
    return {"config": "c4: NOMAD loss fwd+bwd, %d x %.1f s pairs" % (B, N / 16000), "ms_fwd_bwd": ms, "ms_fwd_only": ms_f,
            "pairs_per_s": B / (ms / 1e3), "algorithmic_tflops": fl / (ms / 1e3) / 1e12}


def c5(eng):
    rows = []
    g = torch.Generator().manual_seed(0)
    for n in (1000, 10000, 100000, 1000000):
        for m in (128, 512, 2048, 8192):
            if n * m * 4 > 40e9:
                want = [False]
            else:
                want = [True, False]
            a = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).cuda()
            b = torch.nn.functional.normalize(torch.randn(m, 256, generator=g), dim=1).cuda()
            for wm in want:
                ms = ev_time(lambda: eng.cdist_mean(a, b, want_matrix=wm), 5 if n * m > 1e9 else 20)
                pairs = float(n) * m
                rows.append({"n": n, "m": m, "matrix": wm, "ms": ms, "pairs_per_s": pairs / (ms / 1e3),
                             "write_GBps": (pairs * 4 / (ms / 1e3) / 1e9) if wm else None,
                             "gram_tflops": pairs * 512 / (ms / 1e3) / 1e12})
            del a, b
    return {"config": "c5: distance sweep, D = 256, 1 GPU", "rows": rows}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
    eng = Engine(random_state_dict(1234), 0)
    for w in which:
        t0 = time.time()
        r = {"c1": c1, "lib": lib, "c3": c3, "c4": c4, "c5": c5}[w](eng)
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)
