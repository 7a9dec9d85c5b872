"""world_size-2 gloo tests (CPU) of the batch-sharding plumbing: the compute function is the CPU
oracle, the thing under test is partition / gather / world-size independence."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rlvae_b200 import distributed as D


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 128, 1000, (1 << 20) + 3):
        for w in (1, 2, 3, 4, 8):
            b = [D.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import metric_oracle as O
        from rlvae_b200.synthetic import make_hmc_streams, make_points, make_synthetic_metric
        sm = make_synthetic_metric(64, 8, seed=0)
        t = (sm.centroids.clone(), sm.metric_matrices.clone(), sm.temperature, sm.regularization)

        class FakeMetric:          # table replication: non-src ranks start with garbage
            centroids = t[0] if rank == 0 else torch.zeros_like(t[0])
            metric_matrices = t[1] if rank == 0 else torch.zeros_like(t[1])
            temperature = torch.tensor(t[2] if rank == 0 else 0.0)
            regularization = torch.tensor(t[3] if rank == 0 else 0.0)
            _tab = 'stale'
        fm = FakeMetric()
        D.broadcast_tables(fm, src=0)
        assert fm._tab is None and torch.equal(fm.centroids, t[0]) and torch.equal(fm.metric_matrices, t[1])
        tt = (fm.centroids, fm.metric_matrices, float(fm.temperature), float(fm.regularization))

        n = 37                                           # ragged split: 18 + 19
        z = make_points(n, 8, seed=1)

        def fn(zz):
            return {'ginv': O.inverse_metric(zz, *tt), 'logdet': O.log_det_metric(zz, *tt), 'tag': 'x'}
        out = D.sharded_apply(fn, z, gather=True)
        full = fn(z)
        ok = (out['ginv'].shape == (n, 8, 8) and torch.allclose(out['ginv'], full['ginv'], atol=1e-6)
              and torch.allclose(out['logdet'], full['logdet'], atol=1e-5) and out['tag'] == 'x')
        # scalar reduction
        lo, hi = D.shard_bounds(n, world, rank)
        s = D.all_reduce_scalar(full['logdet'][lo:hi].sum().clone())
        ok = ok and torch.allclose(s, full['logdet'].sum(), atol=1e-4)
        # HMC streams: a rank's chains equal the same rows of the single-process run
        z0, gam, acc = make_hmc_streams(10, 8, 2, seed=2)
        z0r, gr, ar = D.shard_hmc_streams(z0, gam, acc)
        mine = O.hmc_sample(tt, z0r, gr, ar, 3, 0.03)
        whole = O.hmc_sample(tt, z0, gam, acc, 3, 0.03)
        lo, hi = D.shard_bounds(10, world, rank)
        ok = ok and torch.allclose(mine, whole[lo:hi], atol=1e-6)
        gathered = D.all_gather_rows(mine, 10)
        ok = ok and torch.allclose(gathered, whole, atol=1e-6)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_matches_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
