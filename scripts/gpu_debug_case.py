"""Developer aid: run one golden case on the GPU and print the rays whose terminal status differs from the fixture."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytrace_cpu_b200 as K  # noqa: E402
from raytrace_cpu_b200 import _abi as A  # noqa: E402
from golden_cases import CASES  # noqa: E402
import test_parity_gpu as T  # noqa: E402

ctx = K.get_context(0)
ctx.set_timeout(120)
for name in sys.argv[1:]:
    case = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    for mode in (A.MATH_EXACT, A.MATH_FAST):
        init, got = T.run_gpu_case(case, ctx, mode)
        want = g["final"]
        used = want["steps"] != -1
        bad = np.flatnonzero((got["status"] != want["status"]) & used)
        print(f"== {name} mode={mode}: {len(bad)} status mismatches of {used.sum()}")
        for i in bad[:12]:
            a, b = got[i], want[i]
            print(f"  ray {i} alpha={b['alpha']:.4f} beta={b['beta']:.4f}  GPU status={a['status']} steps={a['steps']} r={a['r']:.6g} th={a['theta']:.6g} ph={a['phi']:.6g} flips={a['rdot_flips']}"
                  f" | REF status={b['status']} steps={b['steps']} r={b['r']:.6g} th={b['theta']:.6g} ph={b['phi']:.6g} flips={b['rdot_flips']}")
        same = (got["status"] == want["status"]) & used
        for f in ("r", "theta", "phi", "t", "redshift"):
            d = np.abs(got[f][same] - want[f][same]) / np.maximum(np.abs(want[f][same]), 1e-300)
            d = d[np.isfinite(d)]
            print(f"  {f}: median {np.median(d):.2e} p99 {np.quantile(d, .99):.2e} frac>1e-9 {np.mean(d > 1e-9):.4f}")
        for f in ("k", "h", "Q", "emit", "r", "theta", "phi"):
            d = np.abs(init[f][used] - g["init"][f][used]) / np.maximum(np.abs(g["init"][f][used]), 1e-300)
            d = d[np.isfinite(d)]
            print(f"  init {f}: max rel {d.max():.2e}")
