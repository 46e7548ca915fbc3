"""GPU probe for the tcgen05 GEMM: one case per process (a bad kernel poisons the context).
usage: python tools/probe_gemm.py <case>|all"""
import ctypes as C
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200 import _lib  # noqa: E402

CASES = {
    # name: (M, N, K, lda(0=K), k_wrap, batch, flags, note)
    "tiny": (128, 256, 64, 0, 0, 1, 8),
    "k128": (256, 256, 128, 0, 0, 1, 8),
    "simt_tiny": (128, 256, 64, 0, 0, 1, 8),
    "mtail": (1000, 768, 512, 0, 0, 1, 1 | 2 | 16),
    "ffn": (5000, 3072, 768, 0, 0, 1, 1 | 2 | 16),
    "resid": (4100, 768, 3072, 0, 0, 1, 1 | 4 | 8 | 16),
    "n128": (700, 128, 256, 0, 0, 1, 8),
    "n64": (700, 64, 256, 0, 0, 1, 8),
    "overlap": (999, 512, 1536, 1024, 0, 1, 2 | 16),
    "wrap": (999, 512, 1536, 1024, 1024, 1, 2 | 16),
    "overlap2": (3999, 512, 1024, 1024, 0, 1, 2 | 16),
    "posconv": (300, 48, 6144, 48, 0, 16, 16),
    "big": (51200, 3072, 768, 0, 0, 1, 1 | 2 | 16),
    "big768": (51200, 768, 3072, 0, 0, 1, 1 | 4 | 8),
    "bigqkv": (51200, 2304, 768, 0, 0, 1, 1 | 16),
    "bigconv": (409600, 512, 1536, 1024, 0, 1, 2 | 16),
    "bigout": (51200, 768, 768, 0, 0, 1, 1 | 4 | 8),
    "bigconv2": (204800, 512, 1536, 1024, 0, 1, 2 | 16),
    # epilogue knock-outs on the out-projection shape (which part of the residual epilogue costs what)
    "out_h": (51200, 768, 768, 0, 0, 1, 16),
    "out_bh": (51200, 768, 768, 0, 0, 1, 1 | 16),
    "out_f": (51200, 768, 768, 0, 0, 1, 8),
    "out_fh": (51200, 768, 768, 0, 0, 1, 8 | 16),
    "out_rf": (51200, 768, 768, 0, 0, 1, 4 | 8),
    "out_rfh": (51200, 768, 768, 0, 0, 1, 1 | 4 | 8 | 16),
    "out_rh": (51200, 768, 768, 0, 0, 1, 1 | 4 | 16),
}


def run_case(name):
    M, N, K, lda, wrap, batch, flags = CASES[name]
    impl = 1 if name.startswith("simt") else 0
    lib = C.CDLL(_lib.LIB_PATH)
    for fn in ("nomad_b200_gemm_f16", "nomad_b200_last_error"):
        getattr(lib, fn).restype, getattr(lib, fn).argtypes = _lib.PROTOTYPES[fn]
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    lda_ = lda or K
    if lda and batch == 1:
        rows = M + (K + lda - 1) // lda + 1
        a_buf = (torch.randn(rows * lda, generator=g) * 0.5).to(torch.float16).to(dev)
        a_rows = rows if wrap else M
        a_bs = 0
        idx = (torch.arange(M, device=dev)[:, None] * lda + torch.arange(K, device=dev)[None, :])
        a_mat = a_buf[idx].float()[None]
    elif lda:  # grouped, overlapping rows (pos-conv like): per batch buffer of (M + K/lda) rows of lda
        rows = M + K // lda
        a_buf = (torch.randn(batch, rows * lda, generator=g) * 0.5).to(torch.float16).to(dev)
        a_rows = M
        a_bs = rows * lda
        idx = (torch.arange(M, device=dev)[:, None] * lda + torch.arange(K, device=dev)[None, :])
        a_mat = a_buf[:, idx].float()
    else:
        a_buf = (torch.randn(batch, M, K, generator=g) * 0.5).to(torch.float16).to(dev)
        a_rows = M
        a_bs = M * K
        a_mat = a_buf.float()
    b = (torch.randn(batch, N, K, generator=g) * 0.05).to(torch.float16).to(dev)
    bias = torch.randn(batch, N, generator=g).to(dev)
    ldc = N * batch if batch > 1 else N
    c_bs = N if batch > 1 else 0
    resid = torch.randn(M, ldc, generator=g).to(dev)
    cf = torch.full((M, ldc), float("nan"), device=dev)
    ch = torch.full((M, ldc), float("nan"), device=dev, dtype=torch.float16)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        r = lib.nomad_b200_gemm_f16(a_buf.data_ptr(), a_rows, lda_, wrap, b.data_ptr(), M, N, K, batch, a_bs, N * K,
                                     c_bs, bias.data_ptr(), resid.data_ptr(), cf.data_ptr(), ch.data_ptr(), ldc,
                                     flags, impl, st)
        if r != 0:
            raise RuntimeError(lib.nomad_b200_last_error().decode())

    call()
    torch.cuda.synchronize()
    # reference
    ref = torch.einsum("bmk,bnk->bmn", a_mat, b.float())
    if flags & 1:
        ref = ref + bias[:, None, :]
    if flags & 2:
        ref = torch.nn.functional.gelu(ref)
    ref = ref.permute(1, 0, 2).reshape(M, batch * N)
    if flags & 4:
        ref = ref + resid
    out = {}
    if flags & 8:
        out["f32"] = (cf - ref).abs().max().item()
    if flags & 16:
        out["op_t"] = (ch.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    msg = f"[{name}] M={M} N={N} K={K} lda={lda_} wrap={wrap} batch={batch} flags={flags} impl={impl} max|ref|={scale:.3f} err={out}"
    if M * N * K * batch > 1e10:
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        msg += f" time={ms:.3f} ms {2.0 * M * N * K * batch / ms / 1e9:.1f} TFLOP/s"
        if not lda and batch == 1:  # cuBLAS fp16 on the same shape (no epilogue), for reference
            a2, b2 = a_buf[0], b[0].t().contiguous()
            for _ in range(3):
                torch.matmul(a2, b2)
            e0.record()
            for _ in range(10):
                torch.matmul(a2, b2)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / 10
            msg += f" | cuBLAS {ms2:.3f} ms {2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    tol = 0.02 * max(scale, 1.0)
    bad = [k for k, v in out.items() if not (v <= tol)]
    return 1 if bad else 0


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        rc = 0
        for name in CASES:
            try:
                p = subprocess.run([sys.executable, __file__, name], timeout=180, capture_output=True, text=True)
                print(p.stdout.strip() or f"[{name}] no output", flush=True)
                if p.returncode != 0:
                    rc = 1
                    print(f"[{name}] FAILED rc={p.returncode}\n" + "\n".join(p.stderr.strip().splitlines()[-6:]), flush=True)
            except subprocess.TimeoutExpired:
                rc = 1
                print(f"[{name}] TIMEOUT (hang)", flush=True)
        sys.exit(rc)
    sys.exit(run_case(which))
