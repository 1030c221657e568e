import os, sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import raytrace_cpu_b200 as K
from raytrace_cpu_b200 import _abi as A
import oracle_bindings as ob
import test_parity_gpu as T
from golden_cases import CASES
ctx = K.get_context(0); ctx.set_timeout(300)
for name in ("ps_r5_rk4_thetalim", "plane24_rk4_flatplane"):
    case = CASES[name]
    g = np.load(os.path.join(T.GOLDEN, name + ".npz"))
    for mode in (A.MATH_EXACT, A.MATH_FAST):
        init, got = T.run_gpu_case(case, ctx, mode)
        want = g["final"]
        plane = case["source"] != "point"
        used = (want["steps"] != -1) & ~T.degenerate_mask(want, plane)
        same = (got["status"] == want["status"]) & used
        print(name, "mode", mode, "used", used.sum(), "status differ", (used & ~same).sum())
        for f in ("r", "theta", "t", "redshift", "phi"):
            a, b = got[f][same], want[f][same]
            with np.errstate(invalid="ignore"):
                r = T.rel(a, b)
            print("   ", f, "frac>1e-9", np.nanmean(r > 1e-9), "frac>1e-12", np.nanmean(r > 1e-12), "median", np.nanmedian(r), "nan mismatch", (np.isnan(a) != np.isnan(b)).sum())
        for f in ("steps", "rdot_flips", "equatorial_crossings"):
            print("   ", f, "equal frac", (got[f][same] == want[f][same]).mean())
        fs, fv = T.fma_discrepancy(name, case, want, 1e-9)
        print("    fma oracle discrepancy", fs, fv)
