"""world_size-2 gloo test of the multi-GPU host logic (sharding, the all-gather of NMR embeddings,
row-sharded scoring, restoring listing order) with the oracle standing in for the device kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_embed(seed_rows):
    # deterministic "embedding" of utterance i: unit vector from its index
    def fn(idx):
        out = []
        for i in idx:
            g = torch.Generator().manual_seed(1000 + seed_rows[i])
            v = torch.randn(256, generator=g)
            out.append(v / v.norm())
        return torch.stack(out) if out else torch.zeros(0, 256)
    return fn


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nomad_b200.dist import sharded_embeddings, sharded_scores
    from oracle import w2v_oracle as O

    def cdist_fn(a, b, want):
        dm, mean = O.cdist_mean(a.numpy(), b.numpy())
        return (torch.from_numpy(dm).float() if want else None), torch.from_numpy(mean)

    nmr_cost = [5, 1, 9, 3, 3, 7, 2]
    deg_cost = [4, 4, 1, 8, 2, 6, 3, 3, 5, 1, 7]
    nmr = sharded_embeddings(nmr_cost, _fake_embed(list(range(7))), 256, torch.device("cpu"))
    res = sharded_scores(deg_cost, _fake_embed(list(range(100, 111))), nmr, cdist_fn, torch.device("cpu"))
    if rank == 0:
        q.put((nmr.numpy(), res[0], res[1], res[2]))
    else:
        assert res is None
        q.put(nmr.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_scoring_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full = next(o for o in outs if isinstance(o, tuple))
    other = next(o for o in outs if not isinstance(o, tuple))
    nmr, dm, mean, deg = full
    # every rank holds the same, complete NMR set in listing order
    np.testing.assert_array_equal(nmr, other)
    sys.path.insert(0, ROOT)
    from oracle import w2v_oracle as O
    exp_nmr = _fake_embed(list(range(7)))(list(range(7))).numpy()
    exp_deg = _fake_embed(list(range(100, 111)))(list(range(11))).numpy()
    np.testing.assert_array_equal(nmr, exp_nmr)
    np.testing.assert_array_equal(deg, exp_deg)          # listing order restored bit-exactly
    rdm, rmean = O.cdist_mean(exp_deg, exp_nmr)
    np.testing.assert_allclose(dm, rdm, atol=1e-6)
    np.testing.assert_allclose(mean, rmean, atol=1e-12)
