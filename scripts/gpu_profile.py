"""Minimal driver for ncu: runs the fused lamppost-emissivity kernel a few times (same workload as bench.py, any size).

    ncu --set full -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof python scripts/gpu_profile.py --rays 8e6 --math fast
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=float, default=8e6)
ap.add_argument("--math", default="exact", choices=["exact", "fast"])
ap.add_argument("--integrator", default="rk4", choices=["rk4", "rk45"])
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--min-blocks", type=int, default=0)
ap.add_argument("--h", type=float, default=10.0)
ap.add_argument("--rayset", action="store_true", help="profile the ray-set kernel (144-B records in HBM, results written back) instead")
args = ap.parse_args()

ctx = K.get_context(0)
ctx.set_timeout(600.0)
print(f"DFMA dependent latency: {ctx.fp64_latency_cycles():.2f} cycles/op", flush=True)
if args.min_blocks:
    ctx.set_min_blocks(args.min_blocks)
dcos, dbeta, _ = bench.grid_for(int(args.rays), n_beta=max(100, int(round(args.rays ** 0.5))))
src = K.pointsource_desc([0.0, args.h, 1e-3, 1.5707], 0.0, dcos, dbeta, bench.COS0, bench.COSMAX, bench.BETA0, bench.BETAMAX)
bins = K.emissivity_bins(bench.SPIN, Nr=100, src=src)
mode = A.MATH_FAST if args.math == "fast" else A.MATH_EXACT
meth = K.Integrator.RK45 if args.integrator == "rk45" else K.Integrator.RK4
if args.rayset:
    import ctypes as C
    import math
    ps = K.PointSource([0.0, args.h, 1e-3, 1.5707], 0.0, bench.SPIN, A.TOL, dcos, dbeta, bench.COS0, bench.COSMAX, bench.BETA0, bench.BETAMAX,
                       ctx=ctx, math_mode=mode)
    ps.redshift_start()
    tr = K.api.make_tracer(bench.SPIN, A.TOL)
    stop = K.api.thetalim_stop(math.pi / 2)
    run = K.api.make_run(meth, 1000.0, -1, mode)
    t0 = time.perf_counter()
    K.lib.check(ctx.L.kt_rayset_trace(ps._rs, C.byref(tr), C.byref(stop), C.byref(run)))
    ctx.synchronize() if hasattr(ctx, "synchronize") else K.lib.check(ctx.L.kt_synchronize(ctx.h))
    dt = time.perf_counter() - t0
    n = ps.get_count()
    print(f"ray-set kernel: {n} slots, {dt * 1e3:.1f} ms, {n / dt:.3e} rays/s; algorithmic HBM traffic {n * (128 + 88) / 1e9:.2f} GB "
          f"-> {n * (128 + 88) / dt / 1e9:.1f} GB/s", flush=True)
    sys.exit(0)
for i in range(args.reps):
    t0 = time.perf_counter()
    out = K.emissivity(src, bench.SPIN, bins, method=meth, ctx=ctx, math_mode=mode)
    dt = time.perf_counter() - t0
    c = out.counters
    print(f"rep {i}: {dt * 1e3:.1f} ms  {c[A.CNT_TRACED] / dt:.3e} rays/s  steps/ray {c[A.CNT_STEPS] / c[A.CNT_TRACED]:.1f}  "
          f"alg {259.0 * c[A.CNT_STEPS] / dt / 1e12:.2f} TFLOP/s  counters {c.tolist()}", flush=True)
