"""Golden vectors for the ingest row (``Nomad.load_processing``, reference nomad.py:192-212).

Runs the reference's own operations -- int16 PCM / 32768 as ``torchaudio.load`` returns it, mean of the first two
channels (nomad.py:199-200), ``torchaudio.transforms.Resample(sr, 16000)`` (nomad.py:203-205), 10 s trim
(nomad.py:208-210) -- on small seeded PCM buffers and stores inputs + outputs in tests/golden/ref_ingest.npz.

    python oracle/make_golden_ingest.py
"""
import os

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_load_processing(pcm: np.ndarray, sr: int, target_sr: int = 16000, trim: bool = False) -> np.ndarray:
    wave = torch.from_numpy(pcm.astype(np.float32) / 32768.0).T.contiguous()  # torchaudio.load: (channels, N)
    if wave.shape[0] > 1:
        wave = ((wave[0, :] + wave[1, :]) / 2).unsqueeze(0)
    if sr != target_sr:
        wave = torchaudio.transforms.Resample(sr, target_sr)(wave)
        sr = target_sr
    if trim and wave.shape[1] > sr * 10:
        wave = wave[:, : sr * 10]
    return wave[0].numpy()


def main():
    rng = np.random.default_rng(20240611)
    out = {}
    cases = [(8000, 1, 0.30, False), (22050, 2, 0.25, False), (44100, 2, 0.20, False), (48000, 1, 0.20, False),
             (16000, 2, 0.10, False), (11025, 3, 0.15, False), (8000, 1, 10.5, True)]
    for i, (sr, ch, sec, trim) in enumerate(cases):
        n = int(sr * sec) + 37
        t = np.arange(n)[:, None] / sr
        pcm = (6000 * np.sin(2 * np.pi * (220.0 * (1 + np.arange(ch))[None, :]) * t) + 2500 * rng.standard_normal((n, ch)))
        pcm = np.clip(pcm, -32768, 32767).astype(np.int16)
        y = reference_load_processing(pcm, sr, 16000, trim)
        out[f"pcm{i}"] = pcm if not trim else pcm[:: 1]
        out[f"sr{i}"] = np.int64(sr)
        out[f"trim{i}"] = np.int64(trim)
        out[f"out{i}"] = y if not trim else np.concatenate([y[:2000], y[-2000:]])  # keep the fixture small
        out[f"len{i}"] = np.int64(y.shape[0])
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_ingest.npz"), **out)
    print("wrote", os.path.join(ROOT, "tests", "golden", "ref_ingest.npz"))


if __name__ == "__main__":
    main()
