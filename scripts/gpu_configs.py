"""Times the BASELINE.json configs 1-4 through the fused entry points on one GPU (developer report for DESIGN.md)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402
import oracle_bindings as ob  # noqa: E402

SPIN = 0.998
ctx = K.get_context(0)
ctx.set_timeout(900.0)
modes = [("exact", A.MATH_EXACT), ("fast", A.MATH_FAST)]
only = os.environ.get("CFG_ONLY", "C1,C2,C3,C4").split(",")
n3 = int(float(os.environ.get("CFG_C3_RAYS", "1e7")))
mb = int(os.environ.get("CFG_MINB", "0"))
if mb:
    ctx.set_min_blocks(mb)


def timed(f, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


# C1: emissivity, lamppost h = 10, 1e6 rays, RK4
src = ob.make_pointsource([0, 10, 1e-3, 1.5707], 0.00199, 0.00628318530717959, cosalpha0=-0.995, cosalphamax=0.995,
                          beta0=-3.14159265358979, betamax=3.14159265358979)
bins = K.emissivity_bins(SPIN, Nr=100, src=src)
for name, m in modes if "C1" in only else []:
    dt, out = timed(lambda: K.emissivity(src, SPIN, bins, ctx=ctx, math_mode=m))
    c = out.counters
    print(f"C1 emissivity 1e6 rays RK4 [{name}]: {dt*1e3:.1f} ms, {c[A.CNT_TRACED]/dt:.3e} rays/s, counters {c.tolist()}", flush=True)

# C2: 1024x1024 image plane, incl 30, dist 1e4, RK4
plane = ob.make_imageplane(1e4, 30.0, -30.0, 30.0, 1024)
ib = K.image_bins(plane, SPIN, 30.0, Nx=1024)
for name, m in modes if "C2" in only else []:
    dt, out = timed(lambda: K.disc_image(plane, SPIN, ib, method=K.Integrator.RK4, ctx=ctx, math_mode=m))
    c = out.counters
    print(f"C2 image 1025^2 rays RK4 [{name}]: {dt*1e3:.1f} ms, {c[A.CNT_TRACED]/dt:.3e} rays/s, disc hits {c[A.CNT_DISC]}, "
          f"steps/ray {c[A.CNT_STEPS]/c[A.CNT_TRACED]:.0f}, counters {c.tolist()}", flush=True)

# C3: emissivity RK45, r = 5, tolerance sweep (1e7 rays per tolerance here)
g3 = int(round(n3 ** 0.5))
src3 = ob.make_pointsource([0, 5, 1e-3, 0], 1.99 / g3, 2 * np.pi / g3, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
bins3 = K.emissivity_bins(SPIN, Nr=30, src=src3)
tols = [float(t) for t in os.environ.get("CFG_TOLS", "1e-6,1e-8,1e-10").split(",")]
for tol in tols if "C3" in only else []:
    for name, m in modes:
        dt, out = timed(lambda: K.emissivity(src3, SPIN, bins3, method=K.Integrator.RK45, rk45_tol=tol, ctx=ctx, math_mode=m), reps=1)
        c = out.counters
        print(f"C3 emissivity RK45 tol={tol:g} {n3:.0e} rays [{name}]: {dt*1e3:.1f} ms, {c[A.CNT_TRACED]/dt:.3e} rays/s, steps/ray "
              f"{c[A.CNT_STEPS]/c[A.CNT_TRACED]:.0f}, steplim rays {c[A.CNT_STEPLIM]}", flush=True)

# C4: caustic_discplane, bundles 501^2 x 5, dist 500, RK45 tol 1e-8 and RK4
plane4 = ob.make_imageplane(500.0, 30.0, -20.0, 20.0, 500, bundles=True)
stop4 = ob.make_stop(A.STOP_ISCODISC, r_isco=K.kerr_isco(SPIN), r_out=50.0)
for integ in (K.Integrator.RK4, K.Integrator.RK45) if "C4" in only else []:
    for name, m in modes:
        dt, out = timed(lambda: K.plane_rays(plane4, SPIN, stop4, method=integ, r_max=550.0, ctx=ctx, math_mode=m), reps=1)
        live = out["steps"] != -1
        print(f"C4 bundles 501^2x5 {integ.name} [{name}]: {dt*1e3:.1f} ms, {live.sum()/dt:.3e} rays/s, "
              f"hits {(out['status'][live] & 1).sum()}, steplim {(out['status'][live] & 8 > 0).sum()}", flush=True)
    for name, m in modes:
        dt, out = timed(lambda: K.caustic_discplane(500.0, 30.0, SPIN, 50.0, 500, x0=-20.0, xmax=20.0, method=integ, ctx=ctx, math_mode=m), reps=1)
        print(f"C4 caustic_discplane (maps + Jacobian on device) 501^2x5 {integ.name} [{name}]: {dt*1e3:.1f} ms, disc pixels {out.disc_count}, "
              f"suppressed {out.suppressed}", flush=True)
