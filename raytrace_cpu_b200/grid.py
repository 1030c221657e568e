"""BASELINE config 5: lamppost emissivity over a spin x height grid, rays sharded across the GPUs of one box.

Every (spin, height) point is one fused kernel launch per rank (its shard of that point's ray grid) accumulating into its
own slot of ONE packed device buffer `double[n_points, 5*Nr + 8]`; after the last launch a single NCCL sum all-reduce
combines all points (n_points * 4.9 KB) -- "one batched NCCL reduce" (SURVEY.md section 8e).  torch supplies the device
buffer, the stream and the collective; the arithmetic is libkerrtrace's.
"""
import math

import numpy as np

from . import _abi as A
from . import api
from .sharding import make_shard


def lamppost_grid(spins, heights, rays_per_point, Nr=100, n_beta=None, method=api.Integrator.RK4, r_disc=500.0, gamma=2.0,
                  math_mode=A.MATH_EXACT, ctx=None, rank=0, world=1, group=None):
    """Returns {(spin, height): EmisResult} (whole-job results on every rank).  The source sits on the axis (theta = 1e-3)
    at radius `height`; the ray grid is BASELINE config 1's (cos alpha in [-0.995, 0.995), beta in [-pi, pi))."""
    import torch
    import torch.distributed as dist

    ctx = ctx or api.get_context()
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    points = [(a, h) for a in spins for h in heights]
    width = 5 * Nr + A.NCOUNTERS
    buf = torch.zeros((len(points), width), dtype=torch.float64, device=dev)
    n_beta = n_beta or max(100, int(round(math.sqrt(rays_per_point))))
    n_rows = max(1, int(round(rays_per_point / n_beta)))
    dcos, dbeta = (0.995 + 0.995) / n_rows, 2 * math.pi / n_beta
    bins_of = {}
    for i, (a, h) in enumerate(points):
        src = api.pointsource_desc([0.0, h, 1e-3, 1.5707], 0.0, dcos, dbeta, -0.995, 0.995, -math.pi, math.pi)
        bins = api.emissivity_bins(a, Nr=Nr, r_disc=r_disc, gamma=gamma, src=src)
        bins_of[(a, h)] = bins
        count = ctx.L.kt_pointsource_count(api.C.byref(src), None, None)
        if count < 0:
            raise api.KerrtraceError(ctx.L.kt_last_error().decode())
        shard = make_shard(count, rank, world) if world > 1 else None
        api.emissivity_dev(src, a, bins, buf[i].data_ptr(), method=method, shard=shard, ctx=ctx, math_mode=math_mode)
    if world > 1:
        dist.all_reduce(buf, group=group)      # the one exchange step: all grid points in one message
    host = buf.cpu().numpy()
    return {p: A.EmisResult.from_packed(Nr, host[i]) for i, p in enumerate(points)}, bins_of
