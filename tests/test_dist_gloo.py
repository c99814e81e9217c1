"""world_size-2 `gloo` test of the multi-GPU plumbing on CPU: view sharding, the packed
all-reduce of per-mask sums / counts, and the finalised means against a single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xmask3d_b200 import dist as xd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(seed, n_views, n, k, c):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((n, c)).astype(np.float32)
    member = rng.random((n_views, k, n)) < 0.1
    member[:, 3] = False                      # a mask nobody hits
    return feat, member


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feat, member = _make(5, 7, 400, 6, 16)
        mine = xd.shard_views(7, world, rank)
        # per-view float32 sums, exactly what xm3d_pool_batch hands back per (scene, view) segment
        sums = np.stack([member[v].astype(np.float32) @ feat for v in mine]) if len(mine) else np.zeros((0, 6, 16), np.float32)
        cnts = np.stack([member[v].sum(1) for v in mine]) if len(mine) else np.zeros((0, 6), np.int64)
        work, finish = xd.allreduce_mask_sums(torch.from_numpy(sums), torch.from_numpy(cnts), async_op=True)
        tot, n = finish()
        mean = xd.finalize_mean(tot, n)
        q.put((rank, list(mine), tot.numpy(), n.numpy(), mean.numpy()))
    finally:
        dist.destroy_process_group()


def test_allreduce_mask_sums_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    feat, member = _make(5, 7, 400, 6, 16)
    ref_sum = sum(member[v].astype(np.float64) @ feat.astype(np.float64) for v in range(7))
    ref_cnt = member.sum((0, 2))
    assert sorted(res[0][1] + res[1][1]) == list(range(7))           # views covered exactly once
    for _, _, tot, n, mean in res:
        assert np.array_equal(n, ref_cnt)
        assert np.abs(tot - ref_sum).max() < 1e-4
        ref_mean = ref_sum / np.maximum(ref_cnt, 1)[:, None]
        assert np.abs(mean - ref_mean).max() < 1e-5 and np.all(mean[3] == 0)
    assert np.array_equal(res[0][2], res[1][2])                       # identical on both ranks


def _vote_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        n, t, n_views = 300, 7, 5
        seen = rng.random((n_views, n)) < 0.3
        cls = rng.integers(0, t, (n_views, n))
        votes = torch.zeros((n, t), dtype=torch.int32)
        counter = torch.zeros(n, dtype=torch.int32)
        for v in xd.shard_views(n_views, world, rank):            # this rank's views (numpy restatement of :642-647)
            idx = np.nonzero(seen[v])[0]
            votes[idx, cls[v][idx]] += 1
            counter[idx] += 1
        xd.allreduce_votes(votes, counter)
        q.put((rank, votes.numpy(), counter.numpy()))
    finally:
        dist.destroy_process_group()


def test_allreduce_votes_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vote_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    n, t, n_views = 300, 7, 5
    seen = rng.random((n_views, n)) < 0.3
    cls = rng.integers(0, t, (n_views, n))
    ref_v, ref_c = np.zeros((n, t), np.int32), np.zeros(n, np.int32)
    for v in range(n_views):
        idx = np.nonzero(seen[v])[0]
        ref_v[idx, cls[v][idx]] += 1
        ref_c[idx] += 1
    for _, votes, counter in res:
        assert np.array_equal(votes, ref_v) and np.array_equal(counter, ref_c)


def test_sharding_properties():
    for n, w in ((100, 8), (7, 2), (3, 4), (0, 2), (20, 1)):
        parts = [list(xd.shard_views(n, w, r)) for r in range(w)]
        assert sum(parts, []) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    rng = np.random.default_rng(1)
    weights = rng.lognormal(np.log(150e3), 0.5, 312) * 20
    for w in (1, 2, 4, 8):
        sh = xd.shard_scenes(weights, w)
        assert sorted(sum(sh, [])) == list(range(312))
        load = np.array([weights[s].sum() for s in sh])
        assert load.max() / load.mean() < 1.02                      # near-perfect balance
    assert xd.shard_scenes([1.0, 1.0, 1.0], 2) == [[0, 2], [1]]       # deterministic tie-breaking


def test_single_process_passthrough():
    s = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    c = torch.tensor([[1, 0, 2], [3, 0, 1]])
    tot, n = xd.allreduce_mask_sums(s, c)
    assert torch.equal(tot, s.double().sum(0)) and n.tolist() == [4, 0, 3]
    m = xd.finalize_mean(tot, n)
    assert torch.all(m[1] == 0) and torch.allclose(m[0], (tot[0] / 4).float())
