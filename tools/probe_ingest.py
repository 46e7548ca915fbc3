"""GPU probe: device ingest (PCM16 -> mono -> torchaudio-identical resample) timing per rate pair, 64 x 20 s utterances."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for sr, ch in ((16000, 1), (44100, 2), (48000, 1), (8000, 1), (22050, 1)):
    U = 64
    pcm = [(rng.standard_normal((sr * 20, ch)) * 5000).astype(np.int16) for _ in range(2)]
    dev = [torch.from_numpy(p).cuda() for p in pcm]
    n_out = int(lib.nomad_b200_ingest_out_samples(sr * 20, sr, 16000, 0))
    out = torch.empty(U, n_out, device="cuda")

    def run():
        for u in range(U):
            d = dev[u & 1]
            _lib.check(lib.nomad_b200_ingest_pcm16(C.c_void_p(d.data_ptr()), sr * 20, ch, sr, 16000, 0,
                                                   C.c_void_p(out[u].data_ptr()), st))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"ingest {U} x 20 s @ {sr} Hz x{ch}: {ms * 1e3 / U:7.1f} us/utt -> {U * 20 / ms * 1e3:10.0f} utt-s/s", flush=True)
