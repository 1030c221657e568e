"""Loader and raw ctypes prototypes of libkerrtrace.so (C ABI: include/kerrtrace.h).

The engine has no CPU fallback: if the shared library is missing, or no CUDA device is present when a
context is created, this module raises -- it never routes to the oracle or to any other implementation.
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
# KT_LIB=<variant> loads an experimental build libkerrtrace_<variant>.so (csrc/Makefile: VARIANT=...)
_VARIANT = os.environ.get("KT_LIB", "")
LIB_PATH = os.path.join(_HERE, f"libkerrtrace_{_VARIANT}.so" if _VARIANT else "libkerrtrace.so")


class KerrtraceError(RuntimeError):
    pass


# every symbol include/kerrtrace.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "kt_device_count", "kt_create", "kt_destroy", "kt_last_error", "kt_version", "kt_set_stream", "kt_synchronize",
    "kt_sm_count", "kt_kernel_launches", "kt_fp64_peak", "kt_fp64_latency", "kt_selftest_divsqrt", "kt_set_min_blocks", "kt_host_alloc", "kt_host_free",
    "kt_abi_sizeof", "kt_set_timeout", "kt_aborted", "kt_tracer_init", "kt_kerr_horizon", "kt_kerr_isco", "kt_pointsource_count", "kt_imageplane_count",
    "kt_rayset_create", "kt_rayset_destroy", "kt_rayset_count", "kt_rayset_device_ptr", "kt_rayset_upload",
    "kt_rayset_download", "kt_rayset_init_pointsource", "kt_rayset_init_imageplane", "kt_rayset_redshift_start",
    "kt_rayset_trace", "kt_rayset_redshift", "kt_rayset_redshift_dest", "kt_rayset_range_phi",
    "kt_rayset_calculate_momentum", "kt_rayset_bin_emissivity", "kt_rayset_bin_image",
    "kt_trace", "kt_trace_paths_count", "kt_trace_paths", "kt_redshift_start", "kt_redshift", "kt_redshift_dest", "kt_range_phi", "kt_pointsource_init",
    "kt_imageplane_init", "kt_pointsource_emissivity", "kt_pointsource_emissivity_dev", "kt_imageplane_image",
    "kt_imageplane_image_dev", "kt_imageplane_rays", "kt_caustic_discplane",
]

_lib = None


def load():
    """Load libkerrtrace.so (once).  Raises KerrtraceError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KerrtraceError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C raytrace_cpu_b200/csrc`).  There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    d, i32, i64, vp = C.c_double, C.c_int, C.c_int64, C.c_void_p
    TR, ST, RN = C.POINTER(A.Tracer), C.POINTER(A.Stop), C.POINTER(A.Run)
    PS, IP, SH = C.POINTER(A.PointSourceDesc), C.POINTER(A.ImagePlaneDesc), C.POINTER(A.Shard)
    EB, EO, IB, IO = C.POINTER(A.EmisBins), C.POINTER(A.EmisOut), C.POINTER(A.ImageBins), C.POINTER(A.ImageOut)

    def proto(name, restype, *argtypes):
        f = getattr(L, name)
        f.restype = restype
        f.argtypes = list(argtypes)

    proto("kt_device_count", i32)
    proto("kt_create", vp, i32)
    proto("kt_destroy", None, vp)
    proto("kt_last_error", C.c_char_p)
    proto("kt_version", i32)
    proto("kt_set_stream", i32, vp, vp)
    proto("kt_synchronize", i32, vp)
    proto("kt_sm_count", i32, vp)
    proto("kt_kernel_launches", i64, vp)
    proto("kt_fp64_peak", i32, vp, A.c_double_p)
    proto("kt_fp64_latency", i32, vp, A.c_double_p)
    proto("kt_selftest_divsqrt", i32, vp, i32, i32, A.c_int64_p)
    proto("kt_set_min_blocks", i32, vp, i32)
    proto("kt_host_alloc", vp, C.c_size_t)
    proto("kt_host_free", None, vp)
    proto("kt_abi_sizeof", i32, C.c_char_p)
    proto("kt_set_timeout", i32, vp, d)
    proto("kt_aborted", i32, vp)
    proto("kt_tracer_init", None, TR, d, d)
    proto("kt_kerr_horizon", d, d)
    proto("kt_kerr_isco", d, d, i32)
    proto("kt_pointsource_count", i64, PS, A.c_int32_p, A.c_int32_p)
    proto("kt_imageplane_count", i64, IP, A.c_int32_p, A.c_int32_p)
    proto("kt_rayset_create", vp, vp, i64)
    proto("kt_rayset_destroy", None, vp)
    proto("kt_rayset_count", i64, vp)
    proto("kt_rayset_device_ptr", vp, vp)
    proto("kt_rayset_upload", i32, vp, vp, i64)
    proto("kt_rayset_download", i32, vp, vp, i64)
    proto("kt_rayset_init_pointsource", i32, vp, TR, PS)
    proto("kt_rayset_init_imageplane", i32, vp, TR, IP)
    proto("kt_rayset_redshift_start", i32, vp, TR, d, i32, i32)
    proto("kt_rayset_trace", i32, vp, TR, ST, RN)
    proto("kt_rayset_redshift", i32, vp, TR, d, i32, i32, i32)
    proto("kt_rayset_redshift_dest", i32, vp, TR, ST, i32)
    proto("kt_rayset_range_phi", i32, vp, d, d)
    proto("kt_rayset_calculate_momentum", i32, vp, TR)
    proto("kt_rayset_bin_emissivity", i32, vp, EB, EO)
    proto("kt_rayset_bin_image", i32, vp, IB, IO)
    proto("kt_trace", i32, vp, TR, ST, RN, vp, i64)
    proto("kt_trace_paths_count", i32, vp, TR, ST, RN, vp, i64, C.POINTER(A.PathOpts), A.c_int64_p)
    proto("kt_trace_paths", i32, vp, TR, ST, RN, vp, i64, C.POINTER(A.PathOpts), A.c_int64_p, A.c_double_p)
    proto("kt_redshift_start", i32, vp, TR, vp, i64, d, i32, i32)
    proto("kt_redshift", i32, vp, TR, vp, i64, d, i32, i32, i32)
    proto("kt_redshift_dest", i32, vp, TR, ST, vp, i64, i32)
    proto("kt_range_phi", i32, vp, vp, i64, d, d)
    proto("kt_pointsource_init", i32, vp, TR, PS, vp, i64)
    proto("kt_imageplane_init", i32, vp, TR, IP, vp, i64)
    proto("kt_pointsource_emissivity", i32, vp, PS, TR, ST, RN, EB, SH, EO)
    proto("kt_pointsource_emissivity_dev", i32, vp, PS, TR, ST, RN, EB, SH, vp)
    proto("kt_imageplane_image", i32, vp, IP, TR, ST, RN, IB, SH, IO)
    proto("kt_imageplane_image_dev", i32, vp, IP, TR, ST, RN, IB, SH, vp)
    proto("kt_imageplane_rays", i32, vp, IP, TR, ST, RN, i32, d, i32, vp, i64)
    proto("kt_caustic_discplane", i32, vp, IP, TR, ST, RN, C.POINTER(A.CausticCfg), C.POINTER(A.CausticOut))
    _lib = L
    return L


def last_error():
    return load().kt_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise KerrtraceError(f"kerrtrace error {rc}: {last_error()}")
