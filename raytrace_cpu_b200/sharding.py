"""Host-side description of how rays shard across GPUs (kt_shard in include/kerrtrace.h).

Rays are independent (raytracer.cpp:102), so the flat ray index space [begin, end) is cut into chunks of
`chunk` rays dealt round-robin to ranks: chunk c belongs to rank c % nranks.  Round-robin (rather than one
contiguous block per rank) matters because the cost of a ray varies 2-70x with its direction: every rank
sees the whole range of cos(alpha).  No data-path collective is needed; the only exchange is one sum-reduce of
the per-rank partial histogram at the end.

`shard_indices` is the numpy statement of the ordinal -> index map the kernel applies (kt_trace.cuh:
ordinal_to_index); tests/test_sharding_gloo.py checks that the shards of all ranks tile the index space.
"""
import numpy as np

from . import _abi as A

def auto_chunk(total, nranks):
    """Default chunk size (kt_api.cu: set_work): ~256 chunks per rank, a multiple of 32, within [32, 65536]."""
    c = (total // (max(1, nranks) * 256)) // 32 * 32
    return min(65536, max(32, c))


def _chunk_of(shard):
    return shard.chunk if shard.chunk > 0 else auto_chunk(max(0, shard.end - shard.begin), shard.nranks)


def make_shard(count, rank=0, nranks=1, chunk=0, begin=0, end=None):
    s = A.Shard()
    s.begin = begin
    s.end = count if end is None else min(end, count)
    s.rank, s.nranks, s.chunk = rank, nranks, chunk
    return s


def shard_work_count(shard):
    """Number of work ordinals the kernel hands out for this shard (kt_api.cu: set_work); an upper bound of the
    rays in the shard -- ordinals that map beyond `end` are dropped by the kernel."""
    total = max(0, shard.end - shard.begin)
    nranks = max(1, shard.nranks)
    chunk = _chunk_of(shard)
    if nranks == 1:
        return total
    nchunks = (total + chunk - 1) // chunk
    mine = (nchunks - shard.rank + nranks - 1) // nranks if nchunks > shard.rank else 0
    return mine * chunk


def shard_indices(shard):
    """All ray indices owned by this shard, in the order the work queue hands them out."""
    nranks = max(1, shard.nranks)
    chunk = _chunk_of(shard)
    q = np.arange(shard_work_count(shard), dtype=np.int64)
    if nranks == 1:
        idx = shard.begin + q
    else:
        c, w = q // chunk, q % chunk
        idx = shard.begin + (c * nranks + shard.rank) * chunk + w
    return idx[idx < shard.end]
