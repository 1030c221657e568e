"""Developer probe for a GPU box: quick parity diagnostics against the oracle + timing sweeps.
Not part of the product or the test-suite; prints a report and writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402
from oracle_bindings import Oracle, make_imageplane, make_pointsource, make_run, make_stop  # noqa: E402

OUT = {}
O = Oracle()
ctx = K.get_context(0)
ctx.set_timeout(float(os.environ.get("PROBE_TIMEOUT", "60")))
print("SMs", ctx.sm_count)
t0 = time.time()
peak = ctx.fp64_peak_tflops()
print(f"fp64 DFMA peak {peak:.2f} TFLOP/s ({time.time()-t0:.2f}s)")
OUT["fp64_peak_tflops"] = peak
spin = 0.998


def compare(rg, ro, label):
    used = ro["steps"] != -1
    same_status = (rg["status"] == ro["status"])[used].mean()
    same_steps = (rg["steps"] == ro["steps"])[used].mean()
    both = used & (rg["status"] == ro["status"])
    res = {"n": int(used.sum()), "status_agree": float(same_status), "steps_agree": float(same_steps)}
    for f in ("r", "theta", "phi", "t", "redshift", "k", "h", "Q", "emit"):
        a, b = rg[f][both], ro[f][both]
        fin = np.isfinite(a) & np.isfinite(b)
        rel = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)
        if len(rel):
            res[f] = {"median": float(np.median(rel)), "p99": float(np.quantile(rel, 0.99)), "p999": float(np.quantile(rel, 0.999)),
                      "max": float(rel.max()), "frac_gt_1e-9": float((rel > 1e-9).mean()),
                      "nan_mismatch": int((np.isnan(a) != np.isnan(b)).sum())}
    print(label, json.dumps(res))
    OUT[label] = res


# ---- 1. rayset path vs oracle ---------------------------------------------------------------------------
src = make_pointsource([0, 5, 1e-3, 0], 0.05, 0.05, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
for mode in (A.MATH_EXACT, A.MATH_FAST):
    for integ in (A.RK4, A.RK45):
        for stopk in (A.STOP_THETALIM, A.STOP_FLATDISC, A.STOP_ISCODISC):
            tr = O.tracer(spin, A.TOL)
            ro = O.pointsource_rays(tr, src)
            O.redshift_start(ro, tr, 0.0)
            stop = make_stop(stopk, r_isco=O.kerr_isco(spin), r_out=50.0)
            run = make_run(integ)
            O.trace(ro, tr, stop, run)
            O.range_phi(ro)
            ps = K.PointSource([0, 5, 1e-3, 0], 0.0, spin, A.TOL, 0.05, 0.05, -0.995, 0.995, -np.pi, np.pi, ctx=ctx, math_mode=mode)
            ps.redshift_start()
            t0 = time.time()
            if stopk == A.STOP_THETALIM:
                ps.run_raytrace(K.Integrator(integ), np.pi / 2, 1000.0, 0)
                ctx.synchronize()
                dt = time.time() - t0
                ps.range_phi()
                ps.redshift(-1.0)
                O.redshift(ro, tr, -1.0)
            else:
                dest = K.FlatDiscDestination() if stopk == A.STOP_FLATDISC else K.DiscWithISCODestination(O.kerr_isco(spin), 50.0)
                ps.run_raytrace(dest, K.Integrator(integ), 1000.0, 0)
                ctx.synchronize()
                dt = time.time() - t0
                ps.range_phi()
                ps.redshift(dest)
                O.redshift_dest(ro, tr, stop)
            ctx.check_aborted()
            compare(ps.rays, ro, f"ps5167 mode={mode} integ={integ} stop={stopk} ({dt*1e3:.1f} ms)")

# image plane + bundles
for bundles in (False, True):
    for integ in (A.RK4, A.RK45):
        nx = 33 if integ == A.RK45 else 32
        ip = make_imageplane(1000.0, 30.0, -20, 20, nx, bundles=bundles)
        tr = O.tracer(-spin, A.PRECISION)
        ro = O.imageplane_rays(tr, ip)
        O.redshift_start(ro, tr, 0.0, True)
        stop = make_stop(A.STOP_ISCODISC, r_isco=O.kerr_isco(spin), r_out=50.0)
        O.trace(ro, tr, stop, make_run(integ, r_max=1100.0))
        O.redshift_dest(ro, tr, stop, True)
        cls = K.ImagePlaneBundles if bundles else K.ImagePlane
        pl = cls(1000.0, 30.0, -20, 20, 40.0 / nx, -20, 20, 40.0 / nx, spin, 0.0, ctx=ctx)
        pl.redshift_start()
        dest = K.DiscWithISCODestination(O.kerr_isco(spin), 50.0)
        pl.run_raytrace(dest, K.Integrator(integ), 1100.0, 0)
        pl.redshift(dest, True)
        ctx.check_aborted()
        compare(pl.rays, ro, f"plane bundles={bundles} integ={integ}")

# ---- 2. fused emissivity vs oracle ----------------------------------------------------------------------
src = make_pointsource([0, 10, 1e-3, 1.5707], 0.01, 0.01, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
tr = O.tracer(spin, A.TOL)
ro = O.pointsource_rays(tr, src)
O.redshift_start(ro, tr, 0.0)
t0 = time.time()
O.trace(ro, tr, make_stop(), make_run(A.RK4))
t_cpu = time.time() - t0
O.range_phi(ro)
O.redshift(ro, tr, -1.0)
bins = K.emissivity_bins(spin, Nr=100, src=src)
bo = O.bin_emissivity(ro, bins)
print(f"oracle 125863 rays RK4 {t_cpu:.2f}s = {int((ro['steps']!=-1).sum())/t_cpu:.3e} rays/s; counters", bo.counters.tolist())
for mode in (A.MATH_EXACT, A.MATH_FAST):
    t0 = time.time()
    bg = K.emissivity(src, spin, bins, ctx=ctx, math_mode=mode)
    dt = time.time() - t0
    dN = bg.N - bo.N
    ok = bo.N >= 100
    rel_e = np.abs(bg.emis - bo.emis)[ok] / bo.emis[ok]
    rel_g = np.abs(bg.redshift_sum - bo.redshift_sum)[ok] / bo.redshift_sum[ok]
    print(f"fused emissivity mode={mode} {dt*1e3:.1f} ms counters {bg.counters.tolist()} sum|dN|={int(np.abs(dN).sum())} "
          f"max rel emis {rel_e.max():.3e} max rel gsum {rel_g.max():.3e}")
    OUT[f"fused_emis_mode{mode}"] = {"counters": bg.counters.tolist(), "oracle_counters": bo.counters.tolist(),
                                     "sum_abs_dN": int(np.abs(dN).sum()), "max_rel_emis": float(rel_e.max())}

# ---- 3. timing sweeps (C1 and larger), events via wall clock around a synchronous call -------------------
def c1_src(scale=1):
    return make_pointsource([0, 10, 1e-3, 1.5707], 0.00199 / scale, 2 * np.pi / (1000 * scale), cosalpha0=-0.995, cosalphamax=0.995,
                            beta0=-3.14159265358979, betamax=3.14159265358979)


for scale, label in ((1, "1e6"), (3, "9e6")):
    src = c1_src(scale)
    bins = K.emissivity_bins(spin, Nr=100, src=src)
    for mode in (A.MATH_EXACT, A.MATH_FAST):
        for mb in (3, 4, 5):
            ctx.set_min_blocks(mb)
            best = None
            for rep in range(2):
                t0 = time.time()
                bg = K.emissivity(src, spin, bins, ctx=ctx, math_mode=mode)
                dt = time.time() - t0
                best = dt if best is None else min(best, dt)
            c = bg.counters
            flops = 259.0 * c[A.CNT_STEPS]
            print(f"C1x{label} mode={mode} minblocks={mb}: {best*1e3:.1f} ms, {c[A.CNT_TRACED]/best:.3e} rays/s, "
                  f"steps/ray {c[A.CNT_STEPS]/max(c[A.CNT_TRACED],1):.1f}, alg {flops/best/1e12:.2f} TFLOP/s, counters {c.tolist()}")
            OUT[f"time_{label}_mode{mode}_mb{mb}"] = {"ms": best * 1e3, "rays_per_s": float(c[A.CNT_TRACED] / best),
                                                      "alg_tflops": flops / best / 1e12, "counters": c.tolist()}
ctx.set_min_blocks(0)

# RK45 fused timing (C3-like, r=5, tol 1e-8)
src = make_pointsource([0, 5, 1e-3, 0], 0.004, 0.004, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
bins = K.emissivity_bins(spin, Nr=30, src=src)
for mode in (A.MATH_EXACT, A.MATH_FAST):
    t0 = time.time()
    bg = K.emissivity(src, spin, bins, method=K.Integrator.RK45, ctx=ctx, math_mode=mode)
    dt = time.time() - t0
    c = bg.counters
    print(f"RK45 fused mode={mode}: {dt*1e3:.1f} ms, {c[A.CNT_TRACED]/dt:.3e} rays/s, counters {c.tolist()}")
    OUT[f"rk45_mode{mode}"] = {"ms": dt * 1e3, "rays_per_s": float(c[A.CNT_TRACED] / dt), "counters": c.tolist()}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(OUT, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
print("probe done")
