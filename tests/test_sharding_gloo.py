"""CPU tests of the multi-GPU host logic (N > 1 path): the round-robin ray sharding tiles the index space, and
world_size-2 processes that each reduce their own shard and all-reduce the partial histograms over `gloo`
reproduce the single-process histogram.  The per-shard work is done by the CPU oracle here (a checker standing in
for the GPU); the exchange step is exactly bench.py's: one sum all-reduce of double[5*Nr + 8]."""
import os
import socket
import sys

import numpy as np
import pytest

from raytrace_cpu_b200 import _abi as A
from raytrace_cpu_b200.sharding import make_shard, shard_indices, shard_work_count
import oracle_bindings as ob

SPIN = 0.998


@pytest.mark.parametrize("count,nranks,chunk", [(125863, 2, 65536), (125863, 3, 4096), (125863, 8, 0), (1002000, 4, 0), (1000, 8, 64), (65536, 2, 65536), (5, 4, 2), (0, 2, 16)])
def test_shards_tile_the_index_space(count, nranks, chunk):
    seen = np.zeros(count, dtype=np.int32)
    for r in range(nranks):
        s = make_shard(count, r, nranks, chunk)
        idx = shard_indices(s)
        assert len(idx) <= shard_work_count(s) <= len(idx) + max(chunk, 65536)
        assert (np.diff(idx) > 0).all()
        seen[idx] += 1
    assert (seen == 1).all()


def test_single_rank_shard_is_the_identity():
    s = make_shard(1234)
    assert (shard_indices(s) == np.arange(1234)).all() and shard_work_count(s) == 1234
    s = make_shard(1234, begin=100, end=200)
    assert (shard_indices(s) == np.arange(100, 200)).all()


def test_round_robin_balances_expensive_rows():
    """Cost varies strongly with cos(alpha) (= row of the grid): every rank must see every part of the range."""
    n_beta, n_rows = 1000, 1001
    cost = np.repeat(np.linspace(0, 1, n_rows) ** 4, n_beta)       # a steep cost profile over rows
    loads = [cost[shard_indices(make_shard(n_rows * n_beta, r, 8))].sum() for r in range(8)]
    assert max(loads) / min(loads) < 1.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import raytrace_cpu_b200 as K

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = ob.Oracle()
    src = ob.make_pointsource([0, 10, 1e-3, 1.5707], 0.04, 0.04, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    bins = K.emissivity_bins(SPIN, Nr=20, src=src)
    tr = orc.tracer(SPIN, A.TOL)
    rays = orc.pointsource_rays(tr, src)
    mine = shard_indices(make_shard(len(rays), rank, world, chunk=512))
    part = rays[mine].copy()                      # this rank's shard of the ray grid
    orc.redshift_start(part, tr, 0.0)
    orc.trace(part, tr, ob.make_stop(), ob.make_run(A.RK4), 2)
    orc.range_phi(part)
    orc.redshift(part, tr, -1.0)
    h = orc.bin_emissivity(part, bins)
    packed = np.concatenate([h.N.astype(np.float64), h.flux, h.emis, h.redshift_sum, h.time_sum, h.counters.astype(np.float64)])
    t = torch.from_numpy(packed)
    dist.all_reduce(t)                            # the one exchange step of the multi-GPU path
    if rank == 0:
        np.save(os.path.join(tmpdir, "reduced.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_reproduces_the_whole(tmp_path):
    import torch.multiprocessing as mp
    import raytrace_cpu_b200 as K

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced = A.EmisResult.from_packed(20, np.load(tmp_path / "reduced.npy"))

    orc = ob.Oracle()
    src = ob.make_pointsource([0, 10, 1e-3, 1.5707], 0.04, 0.04, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    bins = K.emissivity_bins(SPIN, Nr=20, src=src)
    tr = orc.tracer(SPIN, A.TOL)
    rays = orc.pointsource_rays(tr, src)
    orc.redshift_start(rays, tr, 0.0)
    orc.trace(rays, tr, ob.make_stop(), ob.make_run(A.RK4))
    orc.range_phi(rays)
    orc.redshift(rays, tr, -1.0)
    whole = orc.bin_emissivity(rays, bins)
    assert (reduced.N == whole.N).all()                                  # integer work: exact
    assert (reduced.counters == whole.counters).all()
    for f in ("flux", "emis", "redshift_sum", "time_sum"):              # sums: order of addition differs
        np.testing.assert_allclose(getattr(reduced, f), getattr(whole, f), rtol=1e-12)
