"""Developer experiment: the host-buffer trace (kt_trace == Raytracer::run_raytrace on rays[]) staged vs zero-copy,
against the device-resident ray-set kernel and the fused kernel on the same grid.  Usage: gpu_e2e.py [rays] [fast]"""
import ctypes as C
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402

n_rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1.6e7)
mode = A.MATH_FAST if "fast" in sys.argv else A.MATH_EXACT
tag = os.environ.get("KT_LIB", "default") + ("/fast" if mode else "/exact") + f"/mb{os.environ.get('KT_MIN_BLOCKS', 'def')}"
SPIN = 0.998


def wall(f, reps=2):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


ctx = K.get_context(0)
ctx.set_timeout(600)
dcos, dbeta, _ = bench.grid_for(n_rays, n_beta=2000)
src = K.pointsource_desc(bench.SOURCE, 0.0, dcos, dbeta, bench.COS0, bench.COSMAX, bench.BETA0, bench.BETAMAX)
bins = K.emissivity_bins(SPIN, Nr=100, src=src)
t = wall(lambda: K.emissivity(src, SPIN, bins, ctx=ctx, math_mode=mode))
print(f"[{tag}] fused emissivity      {n_rays:.2e} rays: {t*1e3:8.1f} ms  {n_rays/t:.3e} rays/s", flush=True)
if "fusedonly" in sys.argv:
    sys.exit(0)

ps = K.PointSource(bench.SOURCE, 0.0, SPIN, A.TOL, dcos, dbeta, bench.COS0, bench.COSMAX, bench.BETA0, bench.BETAMAX, ctx=ctx, math_mode=mode)
ps.redshift_start()
n = ps.get_count()
pristine = torch.empty(n * 144, dtype=torch.uint8).pin_memory()
work = torch.empty(n * 144, dtype=torch.uint8).pin_memory()
K.lib.check(ctx.L.kt_rayset_download(ps._rs, C.c_void_p(pristine.data_ptr()), n))
tr = K.api.make_tracer(SPIN, A.TOL)
stop = K.api.thetalim_stop(math.pi / 2)
run = K.api.make_run(K.Integrator.RK4, 1000.0, -1, mode)


def dev_resident():
    K.lib.check(ctx.L.kt_rayset_trace(ps._rs, C.byref(tr), C.byref(stop), C.byref(run)))


best = 1e30
for _ in range(2):
    K.lib.check(ctx.L.kt_rayset_upload(ps._rs, C.c_void_p(pristine.data_ptr()), n))
    best = min(best, wall(dev_resident, 1))
print(f"[{tag}] rayset kernel (HBM)   {n:.2e} rays: {best*1e3:8.1f} ms  {n/best:.3e} rays/s", flush=True)
want = np.zeros(n, dtype=A.RAY_DTYPE)
K.lib.check(ctx.L.kt_rayset_download(ps._rs, want.ctypes.data_as(C.c_void_p), n))

for zc in (0, 1):
    os.environ["KT_ZEROCOPY"] = str(zc)
    c2 = K.api.Context(0)
    c2.set_timeout(600)
    best = 1e30
    for _ in range(2):
        work.copy_(pristine)
        best = min(best, wall(lambda: K.lib.check(c2.L.kt_trace(c2.h, C.byref(tr), C.byref(stop), C.byref(run), C.c_void_p(work.data_ptr()), n)), 1))
    got = work.numpy().view(A.RAY_DTYPE)
    same = all(np.array_equal(got[f], want[f], equal_nan=True) for f in ("r", "theta", "phi", "t", "pt", "pr", "ptheta", "pphi", "steps", "status"))
    print(f"[{tag}] kt_trace zerocopy={zc}   {n:.2e} rays: {best*1e3:8.1f} ms  {n/best:.3e} rays/s  identical to the HBM run: {same}", flush=True)
    c2.close()
