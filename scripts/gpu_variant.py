"""Developer experiment driver: parity discrepancy vs the oracle + bulk timing for the library selected by KT_LIB."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402
import oracle_bindings as ob  # noqa: E402

mode = A.MATH_FAST if "fast" in sys.argv else A.MATH_EXACT
ctx = K.get_context(0)
ctx.set_timeout(300)
O = ob.Oracle()
SPIN = 0.998
tag = os.environ.get("KT_LIB", "default") + ("/fast" if mode else "/exact")

for h, d in ((5.0, 0.05), (10.0, 0.02)):
    src = ob.make_pointsource([0, h, 1e-3, 0], d, d, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    tr = O.tracer(SPIN, A.TOL)
    want = O.pointsource_rays(tr, src)
    O.redshift_start(want, tr, 0.0)
    O.trace(want, tr, ob.make_stop(), ob.make_run(A.RK4))
    O.redshift(want, tr, -1.0)
    ps = K.PointSource([0, h, 1e-3, 0], 0.0, SPIN, A.TOL, d, d, -0.995, 0.995, -np.pi, np.pi, ctx=ctx, math_mode=mode)
    ps.redshift_start()
    ps.run_raytrace(K.Integrator.RK4, np.pi / 2, 1000.0, 0)
    ps.redshift(-1.0)
    got = ps.rays
    import hashlib
    # bit-identity between builds: hash of every result field of every used ray (NaN payloads canonicalised)
    names = ("t", "r", "theta", "phi", "pt", "pr", "ptheta", "pphi", "redshift", "steps", "status", "rdot_sign", "thetadot_sign",
             "rdot_flips", "equatorial_crossings")
    fields = [np.nan_to_num(np.ascontiguousarray(got[f][want["steps"] != -1]), nan=-777.0) for f in names]
    digest = hashlib.sha1(b"".join(f.tobytes() for f in fields)).hexdigest()[:16]
    used = want["steps"] != -1
    same = (got["status"] == want["status"]) & used
    rel = np.abs(got["r"][same] - want["r"][same]) / np.abs(want["r"][same])
    relg = np.abs(got["redshift"][same] - want["redshift"][same]) / np.abs(want["redshift"][same])
    relg = relg[np.isfinite(relg)]
    print(f"[{tag}] h={h} n={used.sum()}: status differ {(~same & used).sum()}, steps differ {((got['steps'] != want['steps']) & used).sum()}, "
          f"r: bit-equal {np.mean(rel == 0):.4f} >1e-12 {np.mean(rel > 1e-12):.4f} >1e-9 {np.mean(rel > 1e-9):.5f} max {rel.max():.2e}; "
          f"g >1e-9 {np.mean(relg > 1e-9):.5f} sha {digest}", flush=True)

dcos, dbeta, _ = bench.grid_for(int(1.5e7), n_beta=3873)
src = K.pointsource_desc([0.0, 10.0, 1e-3, 1.5707], 0.0, dcos, dbeta, bench.COS0, bench.COSMAX, bench.BETA0, bench.BETAMAX)
bins = K.emissivity_bins(SPIN, Nr=100, src=src)
for i in range(2):
    t0 = time.perf_counter()
    out = K.emissivity(src, SPIN, bins, ctx=ctx, math_mode=mode)
    dt = time.perf_counter() - t0
c = out.counters
print(f"[{tag}] 1.5e7 rays: {dt*1e3:.1f} ms {c[A.CNT_TRACED]/dt:.3e} rays/s counters {c.tolist()}", flush=True)

if "rk45" in sys.argv:
    dcos, dbeta, _ = bench.grid_for(int(2e6), n_beta=1414)
    src = K.pointsource_desc([0.0, 10.0, 1e-3, 1.5707], 0.0, dcos, dbeta, bench.COS0, bench.COSMAX, bench.BETA0, bench.BETAMAX)
    bins = K.emissivity_bins(SPIN, Nr=100, src=src)
    for i in range(2):
        t0 = time.perf_counter()
        out = K.emissivity(src, SPIN, bins, method=K.Integrator.RK45, ctx=ctx, math_mode=mode)
        dt = time.perf_counter() - t0
    c = out.counters
    print(f"[{tag}] RK45 2e6 rays: {dt*1e3:.1f} ms {c[A.CNT_TRACED]/dt:.3e} rays/s counters {c.tolist()} hist-sum {float(np.nansum(out.emis)):.17g}", flush=True)
