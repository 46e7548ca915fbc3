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


def _worker(rank, world, port, q, matrix):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nomad_b200.dist import score_sharded, shard_by_cost
    from oracle import w2v_oracle as O

    def cdist_fn(a, b, want):
        dm, mean = O.cdist_mean(a.numpy(), b.numpy())
        return (torch.from_numpy(dm).float() if want else None), torch.from_numpy(mean)

    nmr_cost = [5, 1, 9, 3, 3, 7, 2]
    deg_cost = [4, 4, 1, 8, 2, 6, 3, 3, 5, 1, 7]
    res = score_sharded(nmr_cost, _fake_embed(list(range(7))), deg_cost, _fake_embed(list(range(100, 111))), cdist_fn,
                        torch.device("cpu"), matrix=matrix, want_emb=True)
    assert res["local_rows"] == shard_by_cost(deg_cost, world)[rank]
    np_ = lambda t: None if t is None else t.numpy()
    q.put((rank, np_(res["nmr"]), np_(res["dm"]), np_(res["mean"]), np_(res["emb"]), res["local_rows"],
           np_(res["local_mean"]), np_(res["local_dm"])))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, matrix):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, matrix)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=100) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return outs


def _expected():
    sys.path.insert(0, ROOT)
    from oracle import w2v_oracle as O
    exp_nmr = _fake_embed(list(range(7)))(list(range(7))).numpy()
    exp_deg = _fake_embed(list(range(100, 111)))(list(range(11))).numpy()
    rdm, rmean = O.cdist_mean(exp_deg, exp_nmr)
    return exp_nmr, exp_deg, rdm, rmean


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_ranks_scoring_equals_single_process(world):
    outs = _run(world, "root")
    exp_nmr, exp_deg, rdm, rmean = _expected()
    for o in outs:                                        # every rank holds the complete NMR set in listing order
        np.testing.assert_array_equal(o[1], exp_nmr)
    _, _, dm, mean, deg, _, _, _ = outs[0]
    np.testing.assert_array_equal(deg, exp_deg)           # listing order restored bit-exactly
    np.testing.assert_allclose(dm, rdm, atol=1e-6)
    np.testing.assert_allclose(mean, rmean, atol=1e-12)
    for o in outs[1:]:                                    # results go to rank 0 ONLY (nothing is all-gathered)
        assert o[2] is None and o[3] is None and o[4] is None


@pytest.mark.timeout(120)
def test_row_sharded_matrix_stays_with_the_ranks():
    outs = _run(2, "local")
    exp_nmr, exp_deg, rdm, rmean = _expected()
    assert outs[0][2] is None                             # no matrix on rank 0 ...
    np.testing.assert_allclose(outs[0][3], rmean, atol=1e-12)   # ... but all the means
    seen = []
    for o in outs:
        rows, lmean, ldm = o[5], o[6], o[7]
        np.testing.assert_allclose(ldm, rdm[rows], atol=1e-6)
        np.testing.assert_allclose(lmean, rmean[rows], atol=1e-12)
        seen += rows
    assert sorted(seen) == list(range(11))


def test_single_process_is_the_identity():
    sys.path.insert(0, ROOT)
    from nomad_b200.dist import score_sharded
    from oracle import w2v_oracle as O

    def cdist_fn(a, b, want):
        dm, mean = O.cdist_mean(a.numpy(), b.numpy())
        return (torch.from_numpy(dm).float() if want else None), torch.from_numpy(mean)

    exp_nmr, exp_deg, rdm, rmean = _expected()
    res = score_sharded([5, 1, 9, 3, 3, 7, 2], _fake_embed(list(range(7))), [4, 4, 1, 8, 2, 6, 3, 3, 5, 1, 7],
                        _fake_embed(list(range(100, 111))), cdist_fn, torch.device("cpu"), want_emb=True)
    np.testing.assert_array_equal(res["nmr"].numpy(), exp_nmr)
    np.testing.assert_array_equal(res["emb"].numpy(), exp_deg)
    np.testing.assert_allclose(res["dm"].numpy(), rdm, atol=1e-6)
    np.testing.assert_allclose(res["mean"].numpy(), rmean, atol=1e-12)
