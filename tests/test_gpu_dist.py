"""Multi-GPU scoring: ``predict_sharded`` equals ``Nomad.predict``.  With one visible GPU the
sharded entry point runs with world size 1; with >= 2 GPUs it is launched under torchrun (NCCL)."""
import os
import subprocess
import sys

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
res = predict_sharded(nomad, "dir", {nmr!r}, {deg!r}, {out!r})
if rank == 0:
    assert res is not None
else:
    assert res is None
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
'''


def _single(state_dict, nmr, deg, out):
    from nomad_b200.nomad import Nomad
    return Nomad(state_dict=state_dict).predict("dir", nmr, deg, out)


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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_predict_sharded_two_ranks_nccl(state_dict, golden_dir, tmp_path):
    nmr, deg = os.path.join(golden_dir, "wavs", "nmr-data"), os.path.join(golden_dir, "wavs", "test-data")
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    avg1, dm1 = _single(state_dict, nmr, deg, str(a))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, nmr=nmr, deg=deg, out=str(b)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    avg2 = pd.read_csv(b / "nomad_avg.csv").set_index("Test File")
    dm2 = pd.read_csv(b / "nomad_scores.csv").set_index("Test File")
    # same listing order and the same scores (each utterance is embedded independently of its batch/rank)
    assert list(avg2.index) == list(avg1.index) and list(dm2.columns) == list(dm1.columns)
    np.testing.assert_allclose(avg2["NOMAD"].to_numpy(), avg1["NOMAD"].to_numpy(), atol=1e-9)
    np.testing.assert_allclose(dm2.to_numpy(), dm1.to_numpy(), atol=1e-9)
