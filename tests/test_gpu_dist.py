"""Multi-GPU scoring: ``predict_sharded`` equals ``Nomad.predict``.  With one visible GPU the
sharded entry point runs with world size 1; with >= 2 GPUs it is launched under torchrun (NCCL) on 2 ranks and on
every visible GPU."""
import os
import subprocess
import sys
import wave

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch
from nomad_b200.dist import init_from_env, predict_sharded
from nomad_b200.nomad import Nomad
from nomad_b200.weights import random_state_dict
rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
nomad = Nomad(device=f"cuda:{{local}}", state_dict=random_state_dict(1234))
nomad.window_files = 7          # several read windows per rank
res = predict_sharded(nomad, "dir", {nmr!r}, {deg!r}, {out!r}, matrix={matrix!r})
assert (res is not None) == (rank == 0)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
'''


def _single(state_dict, nmr, deg, out):
    from nomad_b200.nomad import Nomad
    return Nomad(state_dict=state_dict).predict("dir", nmr, deg, out)


def _write_wav(path, pcm, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(pcm.shape[1]); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.tobytes())


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    """23 degraded + 9 NMR files, 0.3-6 s, mono/stereo, 16 / 44.1 / 48 kHz (the device ingest resamples)."""
    root = tmp_path_factory.mktemp("corpus")
    rng = np.random.default_rng(12)
    for sub, n in (("nmr", 9), ("deg", 23)):
        (root / sub).mkdir()
        for i in range(n):
            sr = (16000, 16000, 44100, 48000)[i % 4]
            ch = 2 if i % 5 == 0 else 1
            frames = int(sr * rng.uniform(0.3, 6.0))
            pcm = (rng.standard_normal((frames, ch)) * 3000).astype(np.int16)
            _write_wav(root / sub / f"{sub}_{i:03d}.wav", pcm, sr)
    return str(root / "nmr"), str(root / "deg")


def test_predict_sharded_world1_equals_predict(state_dict, golden_dir, tmp_path):
    from nomad_b200.dist import predict_sharded
    from nomad_b200.nomad import Nomad
    nmr, deg = os.path.join(golden_dir, "wavs", "nmr-data"), os.path.join(golden_dir, "wavs", "test-data")
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    avg1, dm1 = _single(state_dict, nmr, deg, str(a))
    avg2, dm2 = predict_sharded(Nomad(state_dict=state_dict), "dir", nmr, deg, str(b))
    pd.testing.assert_frame_equal(avg1, avg2)
    pd.testing.assert_frame_equal(dm1, dm2)
    assert open(a / "nomad_scores.csv").read() == open(b / "nomad_scores.csv").read()


def _torchrun(script, nproc, port):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                          capture_output=True, text=True, timeout=900)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("nproc", sorted({2, max(2, min(8, torch.cuda.device_count()))}))
def test_predict_sharded_ranks_nccl(state_dict, corpus, tmp_path, nproc):
    """NCCL ranks produce byte-identical CSVs to the single-GPU ``predict``: every utterance is embedded exactly as if
    alone (so its embedding does not depend on the rank / batch it lands in) and row means are summed in a fixed order."""
    nmr, deg = corpus
    a, b, c = tmp_path / "a", tmp_path / "b", tmp_path / "c"
    a.mkdir(); b.mkdir(); c.mkdir()
    avg1, dm1 = _single(state_dict, nmr, deg, str(a))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, nmr=nmr, deg=deg, out=str(b), matrix=None))
    r = _torchrun(script, nproc, 29533 + nproc)
    assert r.returncode == 0, r.stderr[-3000:]
    assert open(a / "nomad_avg.csv").read() == open(b / "nomad_avg.csv").read()
    assert open(a / "nomad_scores.csv").read() == open(b / "nomad_scores.csv").read()
    # row-sharded scores: every rank writes its rows, rank 0 the means; together they are the same table
    script.write_text(WORKER.format(root=ROOT, nmr=nmr, deg=deg, out=str(c), matrix="local"))
    r = _torchrun(script, nproc, 29633 + nproc)
    assert r.returncode == 0, r.stderr[-3000:]
    assert open(a / "nomad_avg.csv").read() == open(c / "nomad_avg.csv").read()
    assert not (c / "nomad_scores.csv").exists()
    parts = pd.concat([pd.read_csv(c / f"nomad_scores.rank{k}.csv") for k in range(nproc)]).set_index("Test File")
    full = pd.read_csv(a / "nomad_scores.csv").set_index("Test File")
    assert sorted(parts.index) == sorted(full.index) and list(parts.columns) == list(full.columns)
    np.testing.assert_array_equal(parts.loc[full.index].to_numpy(), full.to_numpy())


def test_batched_file_ingest_equals_per_file_loop(state_dict, corpus):
    """``Nomad.embed_files`` (threaded reader, one pinned H2D + one conversion launch per batch of 16 kHz mono files,
    device resampling for the rest) against the reference's per-file host loop (``nomad.py:172-186``)."""
    from nomad_b200.nomad import Nomad
    nmr, deg = corpus
    nomad = Nomad(state_dict=state_dict)
    paths = [os.path.join(deg, f) for f in sorted(os.listdir(deg))]
    nomad.window_files = 5
    fast = nomad.embed_files(np.array(paths)).cpu().numpy()
    ref = np.stack([nomad.engine.embed([nomad.load_processing(p).reshape(-1)]).cpu().numpy()[0] for p in paths])
    for i, p in enumerate(paths):
        with wave.open(p, "rb") as w:
            same_rate_mono = w.getframerate() == 16000 and w.getnchannels() == 1
        if same_rate_mono:
            np.testing.assert_array_equal(fast[i], ref[i])            # same samples, same arithmetic: same bits
        else:
            assert np.abs(fast[i] - ref[i]).max() <= 1e-4, p          # GPU resampler vs torchaudio (1e-6 on the waveform)
