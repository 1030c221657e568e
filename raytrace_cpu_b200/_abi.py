"""ctypes / numpy mirrors of the POD types in include/kerrtrace.h.

Pure declarations: importing this module loads no shared library.  The layouts are checked against the C
side by tests/test_abi.py (sizeof via kt_abi_sizeof) and against the reference's `Ray<double>` (144 bytes,
raytracer.h:65-78) by the oracle tests.
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)

# status bits, raytracer.h:58-63
STATUS_DEST = 1 << 0
STATUS_HORIZON = 1 << 1
STATUS_RLIM = 1 << 2
STATUS_STEPLIM = 1 << 3
STATUS_ERGO = 1 << 4
STATUS_NEG_ENERGY = 1 << 5

# raytracer.h:19-46
PRECISION = 100.0
TOL = 100.0
THETA_PRECISION = 50.0
MAXDT = 1.0
MAXDT_RLIM = 100.0
MAXDPHI = 0.1
RLIM = 1000.0
STEPLIM = 10_000_000
RK45_STEPLIM = 100_000
MIN_STEP = 1e-3

EULER, RK4, RK45 = 0, 1, 2
STOP_THETALIM, STOP_FLATDISC, STOP_ISCODISC, STOP_FLATPLANE = 0, 1, 2, 3
MATH_EXACT, MATH_FAST = 0, 1
RUN_RETIRE_STUCK = 1
HIT_Z, HIT_THETA, HIT_DEST = 0, 1, 2
REDSHIFT_V, REDSHIFT_DEST = 0, 1
CNT_SLOTS, CNT_TRACED, CNT_STEPS, CNT_DEST, CNT_HORIZON, CNT_RLIM, CNT_STEPLIM, CNT_DISC = range(8)
NCOUNTERS = 8

RAY_DTYPE = np.dtype(
    [
        ("t", "f8"), ("r", "f8"), ("theta", "f8"), ("phi", "f8"),
        ("pt", "f8"), ("pr", "f8"), ("ptheta", "f8"), ("pphi", "f8"),
        ("k", "f8"), ("h", "f8"), ("Q", "f8"),
        ("emit", "f8"), ("redshift", "f8"),
        ("steps", "i4"), ("status", "i4"),
        ("rdot_sign", "i4"), ("thetadot_sign", "i4"),
        ("rdot_flips", "i4"), ("equatorial_crossings", "i4"),
        ("alpha", "f8"), ("beta", "f8"),
    ],
    align=True,
)
assert RAY_DTYPE.itemsize == 144

RAY_MIN_DTYPE = np.dtype(
    [
        ("r", "f8"), ("theta", "f8"), ("phi", "f8"), ("t", "f8"), ("redshift", "f8"),
        ("steps", "i4"), ("status", "i4"), ("rdot_flips", "i4"), ("equatorial_crossings", "i4"),
    ],
    align=True,
)
assert RAY_MIN_DTYPE.itemsize == 56


class Stop(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("_pad", C.c_int32),
        ("theta_lim", C.c_double), ("r_isco", C.c_double), ("r_out", C.c_double),
        ("incl", C.c_double), ("phi0", C.c_double), ("z_s", C.c_double),
    ]


class Tracer(C.Structure):
    _fields_ = [
        ("spin", C.c_double), ("horizon", C.c_double), ("precision", C.c_double),
        ("theta_precision", C.c_double), ("max_tstep", C.c_double), ("max_phistep", C.c_double),
        ("maxtstep_rlim", C.c_double), ("rk45_tol", C.c_double),
    ]


class Run(C.Structure):
    _fields_ = [
        ("integrator", C.c_int32), ("steplim", C.c_int32), ("r_max", C.c_double),
        ("math_mode", C.c_int32), ("flags", C.c_int32),
    ]


class PathOpts(C.Structure):  # kt_path_opts
    _fields_ = [("write_step", C.c_int32), ("write_cartesian", C.c_int32), ("write_rmax", C.c_double), ("write_rmin", C.c_double)]


class PointSourceDesc(C.Structure):
    _fields_ = [
        ("pos", C.c_double * 4), ("V", C.c_double), ("E", C.c_double),
        ("dcosalpha", C.c_double), ("dbeta", C.c_double),
        ("cosalpha0", C.c_double), ("cosalphamax", C.c_double),
        ("beta0", C.c_double), ("betamax", C.c_double),
    ]


class ImagePlaneDesc(C.Structure):
    _fields_ = [
        ("dist", C.c_double), ("incl_deg", C.c_double),
        ("x0", C.c_double), ("xmax", C.c_double), ("dx", C.c_double),
        ("y0", C.c_double), ("ymax", C.c_double), ("dy", C.c_double),
        ("phi0", C.c_double),
        ("bundles", C.c_int32), ("_pad", C.c_int32),
        ("eps_frac", C.c_double),
    ]


class Shard(C.Structure):
    _fields_ = [
        ("begin", C.c_int64), ("end", C.c_int64),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("chunk", C.c_int64),
    ]


class EmisBins(C.Structure):
    _fields_ = [
        ("Nr", C.c_int32), ("logbin", C.c_int32),
        ("r_min", C.c_double), ("dr", C.c_double), ("r_isco", C.c_double),
        ("gamma", C.c_double), ("n_primary", C.c_double),
        ("hit_filter", C.c_int32), ("redshift_mode", C.c_int32),
        ("theta_lim", C.c_double), ("V", C.c_double),
    ]


class EmisOut(C.Structure):
    _fields_ = [
        ("N", c_int64_p), ("flux", c_double_p), ("emis", c_double_p),
        ("redshift_sum", c_double_p), ("time_sum", c_double_p),
        ("counters", C.c_int64 * NCOUNTERS),
    ]


class ImageBins(C.Structure):
    _fields_ = [
        ("img_Nx", C.c_int32), ("img_Ny", C.c_int32),
        ("x0", C.c_double), ("y0", C.c_double), ("img_dx", C.c_double), ("img_dy", C.c_double),
        ("flip_image", C.c_int32), ("_pad", C.c_int32),
        ("r_isco", C.c_double), ("r_disc", C.c_double),
        ("q1", C.c_double), ("rb1", C.c_double), ("q2", C.c_double), ("rb2", C.c_double), ("q3", C.c_double),
        ("hit_filter", C.c_int32), ("redshift_mode", C.c_int32), ("reverse", C.c_int32), ("_pad2", C.c_int32),
        ("theta_lim", C.c_double), ("V", C.c_double),
    ]


class ImageOut(C.Structure):
    _fields_ = [
        ("flux", c_double_p), ("r", c_double_p), ("phi", c_double_p), ("enshift", c_double_p),
        ("time", c_double_p), ("emis", c_double_p), ("N", c_int32_p),
        ("counters", C.c_int64 * NCOUNTERS),
    ]


class CausticCfg(C.Structure):  # kt_caustic_cfg
    _fields_ = [
        ("img_Nx", C.c_int32), ("img_Ny", C.c_int32),
        ("r_isco", C.c_double), ("r_disc", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("redshift_mode", C.c_int32), ("suppress", C.c_int32), ("V", C.c_double),
    ]


class CausticOut(C.Structure):  # kt_caustic_out
    _fields_ = [
        ("det_J", c_double_p), ("sign_J", c_double_p), ("order", c_int32_p), ("hit", c_int32_p),
        ("radius", c_double_p), ("phi", c_double_p), ("x_disc", c_double_p), ("y_disc", c_double_p), ("redshift", c_double_p),
        ("disc_count", C.c_int64), ("cnt_horizon", C.c_int64), ("cnt_rlim", C.c_int64), ("cnt_steplim", C.c_int64),
        ("cnt_outofrange", C.c_int64), ("cnt_other", C.c_int64), ("suppressed", C.c_int64),
        ("counters", C.c_int64 * NCOUNTERS),
    ]


CAUSTIC_SENTINEL = 1e30


def dptr(a):
    """double* view of a C-contiguous float64 numpy array."""
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_double_p)


class EmisResult:
    """Host-side owner of the arrays behind a kt_emis_out."""

    def __init__(self, Nr):
        self.N = np.zeros(Nr, dtype=np.int64)
        self.flux = np.zeros(Nr)
        self.emis = np.zeros(Nr)
        self.redshift_sum = np.zeros(Nr)
        self.time_sum = np.zeros(Nr)
        self.c = EmisOut(
            self.N.ctypes.data_as(c_int64_p), dptr(self.flux), dptr(self.emis), dptr(self.redshift_sum),
            dptr(self.time_sum),
        )

    @property
    def counters(self):
        return np.array(list(self.c.counters), dtype=np.int64)

    @classmethod
    def from_packed(cls, Nr, packed):
        """Unpack the double[5*Nr + NCOUNTERS] layout of kt_pointsource_emissivity_dev."""
        out = cls(Nr)
        packed = np.asarray(packed, dtype=np.float64)
        out.N[:] = np.rint(packed[0:Nr]).astype(np.int64)
        out.flux[:] = packed[Nr:2 * Nr]
        out.emis[:] = packed[2 * Nr:3 * Nr]
        out.redshift_sum[:] = packed[3 * Nr:4 * Nr]
        out.time_sum[:] = packed[4 * Nr:5 * Nr]
        for i in range(NCOUNTERS):
            out.c.counters[i] = int(round(packed[5 * Nr + i]))
        return out


class CausticResult:
    """Host-side owner of the nine maps behind a kt_caustic_out; maps are [ix, iy] like Array2D (the FITS extensions of
    caustic_discplane.cpp:527-596 are their transposes)."""

    DOUBLE_MAPS = ("det_J", "sign_J", "radius", "phi", "x_disc", "y_disc", "redshift")
    INT_MAPS = ("order", "hit")
    SCALARS = ("disc_count", "cnt_horizon", "cnt_rlim", "cnt_steplim", "cnt_outofrange", "cnt_other", "suppressed")

    def __init__(self, nx, ny):
        for f in self.DOUBLE_MAPS:
            setattr(self, f, np.zeros((nx, ny)))
        for f in self.INT_MAPS:
            setattr(self, f, np.zeros((nx, ny), dtype=np.int32))
        self.c = CausticOut(dptr(self.det_J), dptr(self.sign_J), self.order.ctypes.data_as(c_int32_p),
                            self.hit.ctypes.data_as(c_int32_p), dptr(self.radius), dptr(self.phi), dptr(self.x_disc),
                            dptr(self.y_disc), dptr(self.redshift))

    def __getattr__(self, name):
        if name in CausticResult.SCALARS:
            return getattr(self.c, name)
        raise AttributeError(name)

    @property
    def counters(self):
        return np.array(list(self.c.counters), dtype=np.int64)


class ImageResult:
    """Host-side owner of the arrays behind a kt_image_out; maps are [ix, iy] like Array2D."""

    FIELDS = ("flux", "r", "phi", "enshift", "time", "emis")

    def __init__(self, nx, ny):
        for f in self.FIELDS:
            setattr(self, f, np.zeros((nx, ny)))
        self.N = np.zeros((nx, ny), dtype=np.int32)
        self.c = ImageOut(*[dptr(getattr(self, f)) for f in self.FIELDS], self.N.ctypes.data_as(c_int32_p))

    @property
    def counters(self):
        return np.array(list(self.c.counters), dtype=np.int64)

    @classmethod
    def from_packed(cls, nx, ny, packed):
        out = cls(nx, ny)
        packed = np.asarray(packed, dtype=np.float64)
        npix = nx * ny
        for i, f in enumerate(cls.FIELDS):
            getattr(out, f)[...] = packed[i * npix:(i + 1) * npix].reshape(nx, ny)
        out.N[...] = np.rint(packed[6 * npix:7 * npix]).astype(np.int32).reshape(nx, ny)
        for i in range(NCOUNTERS):
            out.c.counters[i] = int(round(packed[7 * npix + i]))
        return out
