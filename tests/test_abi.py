"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/kerrtrace.h declares,
the Python mirrors of its structs match the C layouts, the host-side scalars agree with the oracle, and the
product fails loudly (no CPU fallback) when no GPU is present.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import raytrace_cpu_b200 as K
from raytrace_cpu_b200 import _abi as A
from raytrace_cpu_b200 import lib
import oracle_bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kerrtrace.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kt_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(lib.LIB_PATH), "run __graft_entry__.build() first"
    L = lib.load()
    assert L.kt_version() == 100


def test_every_declared_symbol_is_exported():
    syms = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (kt_[a-z0-9_]+)", syms))
    declared = declared_functions()
    assert len(declared) >= 45
    missing = [f for f in declared if f not in exported]
    assert not missing, f"declared in kerrtrace.h but not exported: {missing}"
    # and the Python binding knows all of them
    assert not [f for f in declared if f not in lib.EXPORTS], "lib.EXPORTS is out of date"


def test_cold_path_test_build_exports_the_same_abi():
    """libkerrtrace_cold.so (csrc/Makefile: cold -- every guard-free quotient / root forced onto its out-of-line recomputation) is
    what test_parity_gpu.py::test_cold_paths_are_bit_identical runs against the default build: it must exist and be a drop-in."""
    cold = os.path.join(os.path.dirname(lib.LIB_PATH), "libkerrtrace_cold.so")
    assert os.path.exists(cold), "run __graft_entry__.build() first (make -C raytrace_cpu_b200/csrc all cold)"
    syms = subprocess.run(["nm", "-D", "--defined-only", cold], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (kt_[a-z0-9_]+)", syms))
    assert not [f for f in declared_functions() if f not in exported]


def test_no_torch_or_cxx_types_in_the_abi():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)   # comments may mention torch streams
    assert "torch" not in src and "std::" not in src and "template" not in src
    assert 'extern "C"' in src


@pytest.mark.parametrize("name,mirror", [
    ("kt_stop", A.Stop), ("kt_tracer", A.Tracer), ("kt_run", A.Run), ("kt_pointsource", A.PointSourceDesc),
    ("kt_imageplane", A.ImagePlaneDesc), ("kt_shard", A.Shard), ("kt_emis_bins", A.EmisBins), ("kt_emis_out", A.EmisOut),
    ("kt_image_bins", A.ImageBins), ("kt_image_out", A.ImageOut),
])
def test_struct_mirrors_match_c_layout(name, mirror):
    assert lib.load().kt_abi_sizeof(name.encode()) == C.sizeof(mirror)


def test_ray_layout_is_the_reference_ray_struct():
    L = lib.load()
    assert L.kt_abi_sizeof(b"kt_ray") == A.RAY_DTYPE.itemsize == 144      # sizeof(Ray<double>), raytracer.h:65-78
    assert L.kt_abi_sizeof(b"kt_ray_min") == A.RAY_MIN_DTYPE.itemsize == 56
    names = ["t", "r", "theta", "phi", "pt", "pr", "ptheta", "pphi", "k", "h", "Q", "emit", "redshift", "steps", "status",
             "rdot_sign", "thetadot_sign", "rdot_flips", "equatorial_crossings", "alpha", "beta"]
    assert list(A.RAY_DTYPE.names) == names
    assert A.RAY_DTYPE.fields["steps"][1] == 104 and A.RAY_DTYPE.fields["alpha"][1] == 128
    if ob.ref_available():
        assert ob.Ref().L.ref_sizeof_ray() == 144


def test_host_scalars_match_oracle_and_reference_printouts():
    orc = ob.Oracle()
    for a in (0.0, 0.3, 0.9, 0.998, -0.998):
        assert K.kerr_isco(a) == orc.kerr_isco(a)
        assert K.kerr_isco(a, -1) == orc.kerr_isco(a, -1)
        assert K.kerr_horizon(a) == orc.kerr_horizon(a)
    # values the reference's tests print: horizon(0.998)=1.063214, r_isco(0.998)=1.236971
    assert f"{K.kerr_horizon(0.998):.6f}" == "1.063214"
    assert f"{K.kerr_isco(0.998):.6f}" == "1.236971"


@pytest.mark.parametrize("d,count", [(0.2, 354), (0.05, 5167), (0.01, 125863)])
def test_pointsource_counts_match_reference_printouts(d, count):
    # ray counts for dcosalpha = dbeta = 0.2 / 0.05 / 0.01
    L = lib.load()
    if d == 0.2:
        src = ob.make_pointsource([0, 5, 1e-3, 0], d, d)                       # raytrace_rk4_test defaults
    else:
        src = ob.make_pointsource([0, 5, 1e-3, 0], d, d, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    na, nb = C.c_int32(), C.c_int32()
    n = L.kt_pointsource_count(C.byref(src), C.byref(na), C.byref(nb))
    assert n == count
    assert (n, na.value, nb.value) == ob.Oracle().pointsource_count(src)


def test_imageplane_counts():
    L = lib.load()
    orc = ob.Oracle()
    for nx, bundles in ((32, False), (1024, False), (500, True), (33, True)):
        ip = ob.make_imageplane(1e4, 30.0, -30, 30, nx, bundles=bundles)
        a, b = C.c_int32(), C.c_int32()
        n = L.kt_imageplane_count(C.byref(ip), C.byref(a), C.byref(b))
        assert (n, a.value, b.value) == orc.imageplane_count(ip)
        assert a.value == nx + 1 and n == (nx + 1) ** 2 * (5 if bundles else 1)   # fencepost: Nx+1 rays per axis


def test_tracer_defaults():
    tr = K.api.make_tracer(0.998, 100.0)
    o = ob.Oracle().tracer(0.998, 100.0)
    for f, _ in A.Tracer._fields_:
        assert getattr(tr, f) == getattr(o, f), f
    assert (tr.theta_precision, tr.max_tstep, tr.max_phistep, tr.maxtstep_rlim, tr.rk45_tol) == (50.0, 1.0, 0.1, 100.0, 1e-8)


def test_fails_loudly_without_a_gpu():
    """The product path must not fall back to anything when the device is missing."""
    L = lib.load()
    if L.kt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(K.KerrtraceError, match="no CPU fallback"):
        K.Context(0)
    assert L.kt_create(0) is None
    assert b"no CUDA device" in L.kt_last_error()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    pkg = os.path.join(ROOT, "raytrace_cpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "kerr_oracle" not in text and "oracle_bindings" not in text, os.path.join(dirpath, f)
    deps = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_degenerate_grids_are_refused():
    """A zero / non-finite grid step (inf-to-int cast on the host, divide by a zero count on the device) or an axis without a
    single slot makes kt_pointsource_count / kt_imageplane_count return -1 with a message; the entry points then fail."""
    L = lib.load()
    na, nb = C.c_int32(7), C.c_int32(7)
    for kw in (dict(dcos=0.0), dict(dbeta=0.0), dict(dcos=float("nan")), dict(dbeta=float("inf")), dict(dcos=-0.1)):
        src = ob.make_pointsource([0, 5, 1e-3, 0], kw.get("dcos", 0.05), kw.get("dbeta", 0.05))
        assert L.kt_pointsource_count(C.byref(src), C.byref(na), C.byref(nb)) == -1, kw
        assert (na.value, nb.value) == (0, 0)
        assert b"point-source grid" in L.kt_last_error()
    ip = ob.make_imageplane(1e4, 30.0, -30, 30, 32)
    assert L.kt_imageplane_count(C.byref(ip), None, None) == 33 * 33
    ip.dx = 0.0
    assert L.kt_imageplane_count(C.byref(ip), C.byref(na), C.byref(nb)) == -1
    ip.dx = ip.dy
    ip.dy = float("nan")
    assert L.kt_imageplane_count(C.byref(ip), None, None) == -1
    assert b"image-plane grid" in L.kt_last_error()
