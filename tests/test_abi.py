"""The C-ABI library loads, exports every symbol include/nomad_b200.h declares, and its compute entry
points fail loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from nomad_b200 import _lib
from nomad_b200.weights import conv_out_lengths, expected_shapes, random_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "nomad_b200.h")).read()
    declared = set(re.findall(r"NOMAD_B200_API[^;]*?\b(nomad_b200_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nomad_b200_version().decode().startswith("nomad_b200")


def test_no_runtime_link_dependency_on_cuda_driver():
    # the library must load on a box without libcuda (cudart static, driver entry points at run time)
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out and "not found" not in out


def test_num_frames_matches_conv_arithmetic(lib):
    for n in list(range(380, 460)) + [719, 720, 721, 4000, 16000, 16384, 32000, 64000, 163360, 223840, 320000]:
        exp = conv_out_lengths(n)[-1] if n >= 400 else 0
        assert lib.nomad_b200_num_frames(n) == exp, n
    assert lib.nomad_b200_num_frames(399) == 0 and lib.nomad_b200_num_frames(400) == 1
    assert lib.nomad_b200_num_frames(64000) == 199 and lib.nomad_b200_num_frames(32000) == 99


def test_workspace_planning_is_host_only(lib):
    off = np.array([0, 64000, 64000 + 400, 64000 + 400 + 320000], dtype=np.int64)
    n = lib.nomad_b200_embed_workspace_bytes(off.ctypes.data_as(C.POINTER(C.c_int64)), 3)
    assert n > 0
    # conv0 activation dominates: ~1.5 KB per conv0 row
    rows = sum((conv_out_lengths(int(x))[0] + 63) // 64 * 64 for x in np.diff(off))
    assert n > rows * 512 * 2
    # an utterance shorter than 400 samples is rejected with the reference's complaint
    bad = np.array([0, 399], dtype=np.int64)
    assert lib.nomad_b200_embed_workspace_bytes(bad.ctypes.data_as(C.POINTER(C.c_int64)), 1) == 0
    assert b"at least 400" in lib.nomad_b200_last_error()
    assert lib.nomad_b200_layers_workspace_bytes(32, 32000) > 0
    assert lib.nomad_b200_loss_workspace_bytes(32, 32000, 1) > lib.nomad_b200_layers_workspace_bytes(32, 32000)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(lib):
    sd = random_state_dict(1)
    name = b"embedding_layer.1.bias"
    arr = np.ascontiguousarray(sd[name.decode()].numpy())
    t = (_lib.Tensor * 1)()
    t[0].name, t[0].data, t[0].numel = name, arr.ctypes.data_as(C.c_void_p), arr.size
    h = C.c_void_p()
    assert lib.nomad_b200_create(C.byref(h), t, 1, 0, 0) != 0
    assert b"no CPU fallback" in lib.nomad_b200_last_error()
    from nomad_b200.engine import Engine
    with pytest.raises(_lib.NomadB200Error):
        Engine(sd, 0)
    from nomad_b200.nomad import Nomad
    with pytest.raises(RuntimeError):
        Nomad(device="cpu")


def test_state_dict_layout():
    sd = random_state_dict(1234)
    shapes = expected_shapes()
    assert set(sd) == set(shapes)
    assert sum(v.numel() for v in sd.values()) > 94_000_000
    sd2 = random_state_dict(1234)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
