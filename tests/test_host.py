"""Host-side logic of the reference-facing API that needs no GPU."""
import os
import wave

import numpy as np
import pandas as pd
import pytest
import torch

from nomad_b200 import audio
from nomad_b200.dist import shard_by_cost
from nomad_b200.nomad import Nomad, NomadLoss, plan_batches


def test_plan_batches_covers_everything_in_budget():
    rng = np.random.default_rng(0)
    lens = rng.integers(16000, 320000, size=500).tolist()
    batches = plan_batches(lens, 64 * 16000)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(500))
    for b in batches:
        assert len(b) == 1 or sum(lens[i] for i in b) <= 64 * 16000
        ls = [lens[i] for i in b]
        assert ls == sorted(ls)  # length-bucketed
    assert plan_batches([], 10) == []
    assert plan_batches([5, 50, 5], 10) == [[0, 2], [1]]


def test_shard_by_cost_balanced_and_deterministic():
    costs = [10, 1, 1, 1, 7, 3, 3, 2]
    s = shard_by_cost(costs, 2)
    assert sorted(i for p in s for i in p) == list(range(8))
    loads = [sum(costs[i] for i in p) for p in s]
    assert abs(loads[0] - loads[1]) <= 2
    assert s == shard_by_cost(costs, 2)
    assert shard_by_cost([1.0], 4) == [[0], [], [], []]


def test_load_processing_matches_torchaudio_semantics(golden_dir, tmp_path):
    p = os.path.join(golden_dir, "wavs", "nmr-data", "MJ60_10.wav")
    w = audio.load_processing(p)
    assert w.dtype == torch.float32 and w.shape == (1, 27225)
    with wave.open(p, "rb") as f:
        pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")
    np.testing.assert_array_equal(w.numpy()[0], pcm.astype(np.float32) / 32768.0)
    # a DataFrame row arrives as ndarray (nomad.py:194-195)
    assert audio.load_processing(np.array([p])).shape == (1, 27225)
    # stereo -> mean of the two channels (nomad.py:199-200); trim to 10 s (nomad.py:208-210)
    st = tmp_path / "st.wav"
    a = (np.arange(170000) % 2000 - 1000).astype("<i2")
    b = (-a // 2).astype("<i2")
    with wave.open(str(st), "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(16000)
        f.writeframes(np.stack([a, b], 1).tobytes())
    m = audio.load_processing(str(st))
    np.testing.assert_allclose(m.numpy()[0], (a.astype(np.float32) + b.astype(np.float32)) / 2 / 32768.0, atol=1e-7)
    assert audio.load_processing(str(st), trim=True).shape == (1, 160000)
    # resample 8 kHz -> 16 kHz doubles the length (torchaudio.transforms.Resample, nomad.py:203-205)
    lo = tmp_path / "lo.wav"
    with wave.open(str(lo), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(8000)
        f.writeframes(a[:8000].tobytes())
    assert audio.load_processing(str(lo)).shape == (1, 16000)


def test_predict_argument_errors_match_reference_messages(tmp_path):
    n = Nomad.__new__(Nomad)  # argument validation happens before any device work (nomad.py:83-99)
    with pytest.raises(Exception, match="nmr_path not specified"):
        n.predict("dir", None, "x")
    with pytest.raises(Exception, match="test_path not specified"):
        n.predict("dir", "x", None)
    with pytest.raises(Exception, match="Path to the non-matching reference files .* does not exist"):
        n.predict("dir", str(tmp_path / "nope"), str(tmp_path))
    with pytest.raises(Exception, match="Path to the test files .* does not exist"):
        n.predict("dir", str(tmp_path), str(tmp_path / "nope"))
    with pytest.raises(Exception, match="File .* does not exist"):
        n.predict("csv", str(tmp_path / "a.csv"), str(tmp_path / "b.csv"))
    with pytest.raises(Exception, match="Mode value wav is not valid. Valid values are dir and csv"):
        n.predict("wav", str(tmp_path), str(tmp_path))
    bad = tmp_path / "bad.csv"
    pd.DataFrame({"path": ["a.wav"]}).to_csv(bad, index=False)
    with pytest.raises(Exception, match="not including a column called filename"):
        n.get_embeddings(str(bad))


def test_write_results_reproduces_reference_csv(golden_dir, tmp_path):
    """Feeding the reference's own unrounded matrix through our frame/CSV writer gives its CSV bytes."""
    g = np.load(os.path.join(golden_dir, "ref_predict.npz"))
    n = Nomad.__new__(Nomad)
    deg = ["/some/dir/" + f for f in g["deg_files"]]
    nmr = ["/other/" + f for f in g["nmr_files"]]
    df_avg, df_dm = n.write_results(deg, nmr, g["dm"], g["avg"], str(tmp_path))
    assert open(tmp_path / "nomad_avg.csv").read() == str(g["csv_avg"])
    assert open(tmp_path / "nomad_scores.csv").read() == str(g["csv_scores"])
    assert list(df_avg.index) == list(g["df_avg_index"]) and df_avg.index.name == "Test File"
    assert list(df_dm.columns) == list(g["df_dm_columns"])
    np.testing.assert_array_equal(df_dm.to_numpy(), g["df_dm_values"])
    np.testing.assert_array_equal(df_avg["NOMAD"].to_numpy(), g["df_avg_values"])
    # default location: results-csv/<DD-MM-YYYY_HH-MM-SS>/<ts>_nomad_{avg,scores}.csv (nomad.py:123-133)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        n.write_results(deg, nmr, g["dm"], g["avg"], None)
        (d,) = os.listdir("results-csv")
        assert sorted(os.listdir(os.path.join("results-csv", d))) == [f"{d}_nomad_avg.csv", f"{d}_nomad_scores.csv"]
    finally:
        os.chdir(cwd)


def test_nomad_loss_module_matches_reference_formula():
    torch.manual_seed(0)
    ref = [torch.randn(2, 5, 8) for _ in range(12)] + [torch.randn(2, 4)]
    tst = [torch.randn(2, 5, 8) for _ in range(12)] + [torch.randn(2, 4)]
    exp = sum(torch.nn.functional.l1_loss(t, r) for r, t in zip(ref, tst))
    assert torch.allclose(NomadLoss()(ref, tst), exp)


def test_csv_writer_is_byte_identical_to_pandas(tmp_path):
    """nomad_b200_write_scores_csv vs ``DataFrame.round(3)`` + ``to_csv`` (reference nomad.py:113-120, 138-139): same
    bytes, including rounding ties, exact integers, tiny and huge magnitudes, signed zero, NaN / inf and labels that
    need quoting."""
    import pandas as pd
    from nomad_b200 import _lib
    rng = np.random.default_rng(11)
    base = rng.uniform(0.0, 2.0, size=(257, 9))
    special = np.array([[0.0005, 0.0015, 0.0025, 1.0, 2.0, 1e-5, -1e-5, 123456.7891, 1e17],
                        [np.nan, np.inf, -np.inf, -0.0, 0.0, 0.9995, 0.28, 1e-4, 5e-4],
                        [0.1235, 0.1245, 0.3265, 1e16, 12345678.0005, 3.0, 0.001, 0.01, 0.1]])
    values = np.vstack([special, base])
    rows = [f"utt_{i}" for i in range(values.shape[0])]
    rows[1], rows[2], rows[3] = "with,comma", 'with"quote', " lead space"
    cols = [f"nmr{j}" for j in range(9)]
    cols[4] = "a,b"
    for decimals in (3, 0):
        ref = pd.DataFrame(values, columns=cols).round(decimals)
        ref["Test File"] = rows
        ref.set_index("Test File", inplace=True)
        p_ref, p_got = tmp_path / f"ref{decimals}.csv", tmp_path / f"got{decimals}.csv"
        ref.reset_index().to_csv(p_ref, index=False)
        _lib.write_scores_csv(p_got, "Test File", rows, cols, values, decimals=decimals, threads=3)
        assert p_got.read_bytes() == p_ref.read_bytes()
    # one-column frame (nomad_avg.csv) and an empty one
    avg = values[:, 0]
    ref = pd.DataFrame({"Test File": rows, "NOMAD": avg}).set_index("Test File").round(3)
    ref.reset_index().to_csv(tmp_path / "a_ref.csv", index=False)
    _lib.write_scores_csv(tmp_path / "a_got.csv", "Test File", rows, ["NOMAD"], avg)
    assert (tmp_path / "a_got.csv").read_bytes() == (tmp_path / "a_ref.csv").read_bytes()
    _lib.write_scores_csv(tmp_path / "e.csv", "Test File", [], ["NOMAD"], np.zeros((0, 1)))
    assert (tmp_path / "e.csv").read_bytes() == b"Test File,NOMAD\n"


def test_csv_writer_fuzz_against_pandas(tmp_path):
    """Random magnitudes (1e-9 .. 1e19, both signs), every rounding setting the reference could plausibly use, and no
    rounding at all (full shortest-repr digits): the bytes still equal pandas'."""
    import pandas as pd
    from nomad_b200 import _lib
    rng = np.random.default_rng(123)
    mant = rng.uniform(-10.0, 10.0, size=(400, 7))
    expo = rng.integers(-9, 19, size=(400, 7))
    values = mant * (10.0 ** expo)
    values[::13, 3] = np.round(values[::13, 3])          # exact integers
    values[::17, 5] = rng.integers(-5, 5, size=values[::17, 5].shape) / 8.0  # exact binary fractions (rounding ties)
    rows = [f"r{i}" for i in range(values.shape[0])]
    cols = [f"c{j}" for j in range(values.shape[1])]
    for decimals in (0, 1, 3, 6, -1):
        ref = pd.DataFrame(values, columns=cols)
        if decimals >= 0:
            ref = ref.round(decimals)
        ref.insert(0, "Test File", rows)
        p_ref, p_got = tmp_path / f"ref{decimals}.csv", tmp_path / f"got{decimals}.csv"
        ref.to_csv(p_ref, index=False)
        _lib.write_scores_csv(p_got, "Test File", rows, cols, values, decimals=decimals, threads=2)
        assert p_got.read_bytes() == p_ref.read_bytes(), decimals


def test_native_wav_reader_against_the_wave_module(tmp_path):
    """``nomad_b200_wav_probe`` / ``nomad_b200_wav_read_pcm16`` (host threads, no GPU): headers and samples equal what
    the ``wave`` module reads; anything that is not plain 16-bit PCM is reported as not handled (frames = -1)."""
    import wave

    from nomad_b200 import _lib
    rng = np.random.default_rng(0)
    paths, data = [], []
    for i, (sr, ch, n) in enumerate([(16000, 1, 5000), (44100, 2, 3001), (16000, 1, 1), (8000, 1, 0), (48000, 1, 70001)]):
        p = str(tmp_path / f"f{i}.wav")
        pcm = (rng.standard_normal((n, ch)) * 3000).astype(np.int16)
        with wave.open(p, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.tobytes())
        paths.append(p); data.append(pcm)
    (tmp_path / "bad.wav").write_bytes(b"not a wav file at all")
    with wave.open(str(tmp_path / "f8.wav"), "wb") as w:       # 8-bit PCM: not handled natively
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(16000); w.writeframes(b"\x80" * 100)
    paths += [str(tmp_path / "bad.wav"), str(tmp_path / "missing.wav"), str(tmp_path / "f8.wav")]
    sr, ch, fr, off = _lib.wav_probe(paths, 3)
    assert list(sr[:5]) == [16000, 44100, 16000, 8000, 48000] and list(ch[:5]) == [1, 2, 1, 1, 1]
    assert list(fr) == [5000, 3001, 1, 0, 70001, -1, -1, -1]
    ns = np.array([max(f, 0) * c for f, c in zip(fr, ch)], np.int64)
    dst = np.zeros(len(paths), np.int64)
    np.cumsum(ns[:-1], out=dst[1:])
    buf = np.full(int(ns.sum()) + 4, 12345, np.int16)
    _lib.wav_read_pcm16(paths, off, ns, dst, buf.ctypes.data, 4)
    for i, pcm in enumerate(data):
        np.testing.assert_array_equal(buf[dst[i]: dst[i] + ns[i]], pcm.reshape(-1))
    assert (buf[int(ns.sum()):] == 12345).all()


def test_every_environment_switch_of_the_library_is_documented():
    """Doc-drift guard: each NOMAD_B200_* variable the native sources read appears in INTEGRATION.md's table."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for path in glob.glob(os.path.join(root, "nomad_b200", "csrc", "*.cu")) + glob.glob(os.path.join(root, "nomad_b200", "csrc", "*.cuh")):
        names |= set(re.findall(r'getenv\("(NOMAD_B200_[A-Z0-9_]+)"\)', open(path).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert names, "no switches found: the scan is broken"
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing
