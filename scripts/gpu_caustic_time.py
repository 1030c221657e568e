"""Developer timing: kt_caustic_discplane (trace + maps + Jacobian on device) vs kt_imageplane_rays (trace only), repeated."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402
import oracle_bindings as ob  # noqa: E402

SPIN = 0.998
ctx = K.get_context(0)
ctx.set_timeout(600.0)
plane4 = ob.make_imageplane(500.0, 30.0, -20.0, 20.0, 500, bundles=True)
stop4 = ob.make_stop(A.STOP_ISCODISC, r_isco=K.kerr_isco(SPIN), r_out=50.0)
for name, m in (("exact", A.MATH_EXACT), ("fast", A.MATH_FAST)):
    for i in range(3):
        t0 = time.perf_counter()
        K.plane_rays(plane4, SPIN, stop4, method=K.Integrator.RK4, r_max=550.0, ctx=ctx, math_mode=m)
        t1 = time.perf_counter()
        out = K.caustic_discplane(500.0, 30.0, SPIN, 50.0, 500, x0=-20.0, xmax=20.0, method=K.Integrator.RK4, ctx=ctx, math_mode=m)
        t2 = time.perf_counter()
        print(f"[{name}] rep {i}: plane_rays {1e3*(t1-t0):.1f} ms, caustic_discplane {1e3*(t2-t1):.1f} ms, disc pixels {out.disc_count}", flush=True)
