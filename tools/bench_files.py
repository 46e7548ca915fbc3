"""End-to-end scoring FROM FILES: `nomad.predict('dir')` over a synthetic corpus of 16 kHz mono PCM16
wavs written to /dev/shm -- file reads, PCM ingest, H2D, embedding, distance matrix, CSV writing all inside the clock.

    python tools/bench_files.py [n_deg] [n_nmr]
"""
import json, os, shutil, sys, tempfile, time, wave
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nomad_b200.nomad import Nomad
from nomad_b200.weights import random_state_dict

n_deg = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n_nmr = int(sys.argv[2]) if len(sys.argv) > 2 else 300
base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
root = tempfile.mkdtemp(prefix="nomad_files_", dir=base)
rng = np.random.default_rng(0)
secs = 0.0
for sub, n in (("nmr", n_nmr), ("deg", n_deg)):
    os.makedirs(os.path.join(root, sub))
    for i in range(n):
        d = rng.uniform(1.0, 20.0)
        pcm = (rng.standard_normal(int(16000 * d)) * 3000).astype(np.int16)
        with wave.open(os.path.join(root, sub, f"{sub}_{i:05d}.wav"), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
        secs += len(pcm) / 16000.0
out = os.path.join(root, "out"); os.makedirs(out)
nomad = Nomad(state_dict=random_state_dict(1234))
res = {}
for threads in (1, nomad.reader_threads):
    nomad.reader_threads = threads
    nomad.predict("dir", os.path.join(root, "nmr"), os.path.join(root, "deg"), out)   # warm-up (page cache, workspaces)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nomad.predict("dir", os.path.join(root, "nmr"), os.path.join(root, "deg"), out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[f"reader_threads_{threads}"] = {"seconds": dt, "utt_s_per_s": secs / dt}
print(json.dumps({"config": f"predict('dir') from files: {n_deg} degraded + {n_nmr} NMR wavs, 1-20 s, 16 kHz mono PCM16 on {base}, "
                            f"wall clock incl. reads, ingest, embedding, {n_deg}x{n_nmr} distances and both CSVs",
                  "utt_s": secs, "cores": os.cpu_count(), **res}))
shutil.rmtree(root, ignore_errors=True)
