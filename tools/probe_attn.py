"""GPU probe of the attention core through the C ABI: parity against torch fp32 and timing.
usage: python tools/probe_attn.py [bench|mixed]   (knobs: NOMAD_B200_FA_SLEEP_TMA / _MMA)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200 import _lib  # noqa: E402


def run(Ts, reps=20, check=True):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    Ts = np.asarray(Ts, dtype=np.int32)
    rows = ((Ts + 0) + 0).copy()
    frame0 = np.zeros(len(Ts), dtype=np.int32)
    f = 0
    for u, t in enumerate(Ts):
        frame0[u] = f
        f += int(t) + 1  # one padding row between utterances
    frames = f
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(frames, 2304, generator=g) * 1.5)
    qkv[:, :768] *= 0.125 * 3.0  # q scaled; wide score range to exercise the running-max rescale
    qkv = qkv.to(torch.float16).to(dev)
    out = torch.zeros(frames, 768, dtype=torch.float16, device=dev)
    lse = torch.zeros(frames, 12, dtype=torch.float32, device=dev)
    wsb = lib.nomad_b200_attention_workspace_bytes(Ts.ctypes.data_as(C.POINTER(C.c_int32)), len(Ts))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        _lib.check(lib.nomad_b200_attention_f16(qkv.data_ptr(), frames, frame0.ctypes.data_as(C.POINTER(C.c_int32)),
                                                Ts.ctypes.data_as(C.POINTER(C.c_int32)), len(Ts), out.data_ptr(),
                                                lse.data_ptr(), ws.data_ptr(), wsb, st), "attention")
    call()
    torch.cuda.synchronize()
    err = lerr = 0.0
    if check:
        for u in list(range(min(4, len(Ts)))) + [len(Ts) - 1]:
            t, f0 = int(Ts[u]), int(frame0[u])
            x = qkv[f0:f0 + t].float().view(t, 3, 12, 64)
            q, k, v = x[:, 0].transpose(0, 1), x[:, 1].transpose(0, 1), x[:, 2].transpose(0, 1)
            s = q @ k.transpose(1, 2)
            ref = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(t, 768)
            err = max(err, float((out[f0:f0 + t].float() - ref).abs().max()))
            lerr = max(lerr, float((lse[f0:f0 + t] - torch.logsumexp(s, -1).transpose(0, 1)).abs().max()))
    from torch.profiler import ProfilerActivity, profile
    for _ in range(3):
        call()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            call()
        torch.cuda.synchronize()
    ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "attention" in e.name]
    us = sum(e.time_range.end - e.time_range.start for e in ks) / max(1, len(ks))
    fl = sum(4.0 * 12 * 64 * float(t) * float(t) for t in Ts)
    print(f"utts={len(Ts)} T=[{Ts.min()}..{Ts.max()}] frames={frames}: {us:8.1f} us/kernel (CUPTI)  "
          f"{fl / us / 1e6:7.1f} TFLOP/s  out err {err:.2e} lse err {lerr:.2e}", flush=True)
    return err, lerr


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "bench"
    if what == "bench":
        run([199] * 256)
    elif what == "mixed":
        rng = np.random.default_rng(0)
        run([1, 2, 63, 64, 65, 127, 128, 129, 199, 256, 257, 511, 999, 1500])
        run(list(rng.integers(49, 1000, size=64)))
        run([99] * 64)
        run([999] * 48)
