"""Python mirror of the reference's host interface for the hot path, bound to libkerrtrace.so.

Same class names, method names, argument order, defaults and semantics as the reference's C++ API
(SURVEY.md section 8b): `Raytracer`, `PointSource`, `ImagePlane`, `ImagePlaneBundles`, `Integrator`,
`FlatDiscDestination`, `DiscWithISCODestination`, `FlatPlaneDestination`  (raytracer.h, pointsource.h,
imageplane.h, imageplane_bundles.h, ray_destination.h).  Rays live in HBM as an AoS `kt_ray` set; `.rays`
downloads them as a numpy structured array with the reference's field names.

`emissivity()`, `disc_image()` and `plane_rays()` are the fused paths (whole apps in one kernel).
Everything computes on the GPU through the C ABI; there is no CPU fallback.
"""
import ctypes as C
import enum
import math

import numpy as np

from . import _abi as A
from . import lib as _lib
from .lib import KerrtraceError, check

M_PI = math.pi
M_PI_2 = math.pi / 2


class Integrator(enum.IntEnum):  # raytracer.h:83
    Euler = A.EULER
    RK4 = A.RK4
    RK45 = A.RK45


# ---------------------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------------------
class Context:
    """One engine context = one GPU (one process per GPU)."""

    def __init__(self, device=0):
        self.L = _lib.load()
        self.h = self.L.kt_create(int(device))
        if not self.h:
            raise KerrtraceError(_lib.last_error())
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.kt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Issue all work on an existing CUDA stream (pass torch.cuda.current_stream().cuda_stream).
        0 means the legacy default stream (what torch uses unless told otherwise), None the context's own stream."""
        if cuda_stream is None:
            handle = None
        else:
            handle = C.c_void_p(cuda_stream if cuda_stream else 1)   # cudaStreamLegacy == (cudaStream_t)0x1
        check(self.L.kt_set_stream(self.h, handle))

    def synchronize(self):
        check(self.L.kt_synchronize(self.h))

    @property
    def sm_count(self):
        return self.L.kt_sm_count(self.h)

    @property
    def kernel_launches(self):
        return self.L.kt_kernel_launches(self.h)

    def set_min_blocks(self, n):
        check(self.L.kt_set_min_blocks(self.h, int(n)))

    def set_timeout(self, seconds):
        """Arm (seconds > 0) or disarm the watchdog of the persistent trace kernels."""
        check(self.L.kt_set_timeout(self.h, float(seconds)))

    def check_aborted(self):
        check(self.L.kt_aborted(self.h))

    def fp64_latency_cycles(self):
        v = C.c_double()
        check(self.L.kt_fp64_latency(self.h, C.byref(v)))
        return v.value

    def fp64_peak_tflops(self):
        v = C.c_double()
        check(self.L.kt_fp64_peak(self.h, C.byref(v)))
        return v.value


_default_ctx = {}


def get_context(device=None):
    import os

    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---------------------------------------------------------------------------------------------------------
# host-side scalars
# ---------------------------------------------------------------------------------------------------------
def kerr_horizon(a):
    return _lib.load().kt_kerr_horizon(a)


def kerr_isco(a, sign=+1):
    return _lib.load().kt_kerr_isco(a, sign)


def make_tracer(spin, precision=A.PRECISION):
    tr = A.Tracer()
    _lib.load().kt_tracer_init(C.byref(tr), spin, precision)
    return tr


# ---------------------------------------------------------------------------------------------------------
# stop surfaces: ray_destination.h:85-204
# ---------------------------------------------------------------------------------------------------------
class RayDestination:
    """Base of the device-side stop surfaces.  Arbitrary user subclasses cannot run on the GPU."""

    def _stop(self):
        raise KerrtraceError("only FlatDiscDestination, DiscWithISCODestination and FlatPlaneDestination are offloaded")


class FlatDiscDestination(RayDestination):
    def __init__(self, theta_lim=M_PI_2):
        self.theta_lim = theta_lim

    def _stop(self):
        s = A.Stop()
        s.kind, s.theta_lim = A.STOP_FLATDISC, self.theta_lim
        return s


class DiscWithISCODestination(RayDestination):
    def __init__(self, r_isco, r_out=-1, theta_lim=M_PI_2):
        self.r_isco, self.r_out, self.theta_lim = r_isco, r_out, theta_lim

    def _stop(self):
        s = A.Stop()
        s.kind, s.theta_lim, s.r_isco, s.r_out = A.STOP_ISCODISC, self.theta_lim, self.r_isco, self.r_out
        return s


class FlatPlaneDestination(RayDestination):
    def __init__(self, incl, phi0, z_s):
        self.incl, self.phi0, self.z_s = incl, phi0, z_s

    def projection(self, r, theta, phi):
        return r * (np.sin(theta) * np.sin(self.incl) * np.cos(phi - self.phi0) + np.cos(theta) * np.cos(self.incl))

    def source_coords(self, r, theta, phi):
        X = r * np.sin(theta) * np.cos(phi)
        Y = r * np.sin(theta) * np.sin(phi)
        Z = r * np.cos(theta)
        x_s = -X * np.sin(self.phi0) + Y * np.cos(self.phi0)
        y_s = -X * np.cos(self.incl) * np.cos(self.phi0) - Y * np.cos(self.incl) * np.sin(self.phi0) + Z * np.sin(self.incl)
        return x_s, y_s

    def _stop(self):
        s = A.Stop()
        s.kind, s.incl, s.phi0, s.z_s = A.STOP_FLATPLANE, self.incl, self.phi0, self.z_s
        return s


def thetalim_stop(theta_max):
    s = A.Stop()
    s.kind, s.theta_lim = A.STOP_THETALIM, theta_max
    return s


def make_run(method, r_max, steplim=-1, math_mode=A.MATH_EXACT, flags=0):
    r = A.Run()
    r.integrator, r.r_max, r.steplim, r.math_mode, r.flags = int(method), r_max, steplim, math_mode, flags
    return r


class TextOutput:
    """include/text_output.h: columns of width 20, scientific notation, 8 digits -- `setw(20) << scientific << setprecision(8)`."""

    def __init__(self, filename, append=False, precision=8, colwidth=20):
        self._own = isinstance(filename, (str, bytes)) or hasattr(filename, "__fspath__")
        self._f = open(filename, "a" if append else "w") if self._own else filename
        self.precision, self.width = precision, colwidth

    def row(self, values):
        self._f.write("".join(f"{v:{self.width}.{self.precision}e}" for v in values) + "\n")

    def newline(self, n=1):
        self._f.write("\n" * n)

    def close(self):
        if self._own and self._f:
            self._f.close()
            self._f = None


# ---------------------------------------------------------------------------------------------------------
# Raytracer<double>: raytracer.h:86-198
# ---------------------------------------------------------------------------------------------------------
class Raytracer:
    def __init__(self, num_rays, spin, precision=A.PRECISION, init_max_phistep=A.MAXDPHI, init_max_tstep=A.MAXDT,
                 ctx=None, math_mode=A.MATH_EXACT):
        self.ctx = ctx or get_context()
        self.L = self.ctx.L
        self.nRays = int(num_rays)
        self.spin = spin
        self._tr = make_tracer(spin, precision)
        self._tr.max_phistep = init_max_phistep
        self._tr.max_tstep = init_max_tstep
        self.horizon = self._tr.horizon
        self.math_mode = math_mode
        self._rs = self.L.kt_rayset_create(self.ctx.h, self.nRays)
        if not self._rs:
            raise KerrtraceError(_lib.last_error())
        self._host = None

    def __del__(self):
        try:
            if getattr(self, "_rs", None):
                self.L.kt_rayset_destroy(self._rs)
                self._rs = None
        except Exception:
            pass

    # ---- data access -----------------------------------------------------------------------------------
    @property
    def rays(self):
        """The ray set as a numpy structured array (fields named as struct Ray, raytracer.h:65-78)."""
        if self._host is None:
            buf = np.zeros(self.nRays, dtype=A.RAY_DTYPE)
            if self.nRays:
                check(self.L.kt_rayset_download(self._rs, buf.ctypes.data_as(C.c_void_p), self.nRays))
            self._host = buf
        return self._host

    def upload(self, rays):
        assert rays.dtype == A.RAY_DTYPE and len(rays) == self.nRays
        rays = np.ascontiguousarray(rays)
        check(self.L.kt_rayset_upload(self._rs, rays.ctypes.data_as(C.c_void_p), self.nRays))
        self._host = None

    def _dirty(self):
        self._host = None

    def get_count(self):
        return self.nRays

    # ---- configuration -----------------------------------------------------------------------------------
    def set_boundary(self, r=-1):
        self._tr.horizon = r if r > 0 else kerr_horizon(self.spin)
        self.horizon = self._tr.horizon

    def calculate_horizon(self):
        return kerr_horizon(self.spin)

    def set_precision(self, precision=A.PRECISION, theta_precision=A.THETA_PRECISION):
        """No-op, as in the reference (self-assignment bug, raytracer.h:169-178)."""

    def set_rk45_tol(self, tol):
        self._tr.rk45_tol = tol

    def get_rk45_tol(self):
        return self._tr.rk45_tol

    def set_max_tstep(self, max_, rlim=A.MAXDT_RLIM):
        self._tr.max_tstep = max_
        self._tr.maxtstep_rlim = rlim

    def set_max_phistep(self, max_):
        self._tr.max_phistep = max_

    # ---- the hot path ------------------------------------------------------------------------------------
    def run_raytrace(self, *args, **kw):
        """Both reference overloads:
        run_raytrace(method=Integrator.Euler, theta_max=pi/2, r_max=1000, show_progress=1, outfile=None, ..., steplim=-1)
        run_raytrace(dest, method=Integrator.Euler, r_max=1000, show_progress=1, outfile=None, ..., steplim=-1)
        """
        names_theta = ["method", "theta_max", "r_max", "show_progress", "outfile", "write_step", "write_rmax",
                       "write_rmin", "write_cartesian", "steplim"]
        names_dest = ["dest", "method", "r_max", "show_progress", "outfile", "write_step", "write_rmax", "write_rmin",
                      "write_cartesian", "steplim"]
        use_dest = (args and isinstance(args[0], RayDestination)) or "dest" in kw
        names = names_dest if use_dest else names_theta
        p = dict(method=Integrator.Euler, theta_max=M_PI_2, r_max=1000, show_progress=1, outfile=None, write_step=1,
                 write_rmax=-1, write_rmin=-1, write_cartesian=True, steplim=-1, dest=None)
        for n, v in zip(names, args):
            p[n] = v
        p.update(kw)
        if use_dest:
            assert p["method"] != Integrator.Euler, "Integrator::Euler does not support RayDestination stopping conditions"
            stop = p["dest"]._stop()
        else:
            stop = thetalim_stop(p["theta_max"])
        run = make_run(p["method"], p["r_max"], p["steplim"], self.math_mode)
        if p["outfile"] is not None:
            return self._trace_paths(stop, run, p)
        check(self.L.kt_rayset_trace(self._rs, C.byref(self._tr), C.byref(stop), C.byref(run)))
        self._dirty()

    def _trace_paths(self, stop, run, p):
        """run_raytrace with an outfile (raytracer.cpp:86-100): `outfile` is a TextOutput, a path or a writable text
        stream.  Returns (offsets, points): ray i wrote points[offsets[i]:offsets[i + 1]], rows (t, x, y, z) or (t, r, theta, phi)."""
        opts = A.PathOpts()
        opts.write_step, opts.write_cartesian = int(p["write_step"]), int(bool(p["write_cartesian"]))
        opts.write_rmax, opts.write_rmin = float(p["write_rmax"]), float(p["write_rmin"])
        host = np.ascontiguousarray(self.rays)
        n = self.nRays
        before = host["steps"].copy()
        offsets = np.zeros(n + 1, dtype=np.int64)
        ptr = host.ctypes.data_as(C.c_void_p)
        args = (self.ctx.h, C.byref(self._tr), C.byref(stop), C.byref(run), ptr, n, C.byref(opts))
        check(self.L.kt_trace_paths_count(*args, offsets.ctypes.data_as(A.c_int64_p)))
        points = np.zeros((int(offsets[n]), 4), dtype=np.float64)
        check(self.L.kt_trace_paths(*args, offsets.ctypes.data_as(A.c_int64_p), points.ctypes.data_as(A.c_double_p)))
        self.upload(host)
        steplim = run.steplim if run.steplim > 0 else (A.RK45_STEPLIM if run.integrator == A.RK45 else A.STEPLIM)
        out = p["outfile"] if isinstance(p["outfile"], TextOutput) else TextOutput(p["outfile"])
        for i in range(n):
            if before[i] < 0 or before[i] >= steplim:      # skipped rays write nothing, not even the blank lines
                continue
            for row in points[offsets[i]:offsets[i + 1]]:
                out.row(row)
            out.newline(2)
        if out is not p["outfile"]:
            out.close()
        return offsets, points

    def redshift_start(self, V, reverse=False, projradius=False):
        check(self.L.kt_rayset_redshift_start(self._rs, C.byref(self._tr), V, int(reverse), int(projradius)))
        self._dirty()

    def redshift(self, V_or_dest, reverse=False, projradius=False, motion=0):
        if isinstance(V_or_dest, RayDestination):
            stop = V_or_dest._stop()
            check(self.L.kt_rayset_redshift_dest(self._rs, C.byref(self._tr), C.byref(stop), int(reverse)))
        else:
            check(self.L.kt_rayset_redshift(self._rs, C.byref(self._tr), float(V_or_dest), int(reverse), int(projradius), motion))
        self._dirty()

    def range_phi(self, min=-M_PI, max=M_PI):  # noqa: A002 (reference argument names)
        check(self.L.kt_rayset_range_phi(self._rs, min, max))
        self._dirty()

    def calculate_momentum(self):
        check(self.L.kt_rayset_calculate_momentum(self._rs, C.byref(self._tr)))
        self._dirty()

    # ---- device-side versions of the apps' binning loops ----------------------------------------------------
    def bin_emissivity(self, bins):
        out = A.EmisResult(bins.Nr)
        check(self.L.kt_rayset_bin_emissivity(self._rs, C.byref(bins), C.byref(out.c)))
        return out

    def bin_image(self, bins):
        out = A.ImageResult(bins.img_Nx, bins.img_Ny)
        check(self.L.kt_rayset_bin_image(self._rs, C.byref(bins), C.byref(out.c)))
        return out


def pointsource_desc(pos, V, dcosalpha, dbeta, cosalpha0=-0.999999, cosalphamax=0.995, beta0=-0.995 * M_PI,
                     betamax=M_PI, E=1.0):
    s = A.PointSourceDesc()
    s.pos[:] = list(pos)
    s.V, s.E = V, E
    s.dcosalpha, s.dbeta = dcosalpha, dbeta
    s.cosalpha0, s.cosalphamax, s.beta0, s.betamax = cosalpha0, cosalphamax, beta0, betamax
    return s


def imageplane_desc(dist, inc, x0, xmax, dx, y0, ymax, dy, phi, bundles=False, eps_frac=0.01):
    p = A.ImagePlaneDesc()
    p.dist, p.incl_deg = dist, inc
    p.x0, p.xmax, p.dx = x0, xmax, dx
    p.y0, p.ymax, p.dy = y0, ymax, dy
    p.phi0 = phi
    p.bundles = int(bundles)
    p.eps_frac = eps_frac
    return p


class PointSource(Raytracer):
    """pointsource.h:24: PointSource(pos, V, spin, tol, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax, E)"""

    def __init__(self, pos, V, spin, tol, dcosalpha, dbeta, cosalpha0=-0.999999, cosalphamax=0.995,
                 beta0=-0.995 * M_PI, betamax=M_PI, E=1.0, ctx=None, math_mode=A.MATH_EXACT):
        self.desc = pointsource_desc(pos, V, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax, E)
        na, nb = C.c_int32(), C.c_int32()
        n = _lib.load().kt_pointsource_count(C.byref(self.desc), C.byref(na), C.byref(nb))
        if n < 0:
            raise KerrtraceError(_lib.load().kt_last_error().decode())
        self.n_cosalpha, self.n_beta = na.value, nb.value
        self.velocity, self.energy = V, E
        super().__init__(n, spin, tol, ctx=ctx, math_mode=math_mode)
        check(self.L.kt_rayset_init_pointsource(self._rs, C.byref(self._tr), C.byref(self.desc)))

    def redshift_start(self, *args, **kw):
        if args or kw:
            return super().redshift_start(*args, **kw)
        return super().redshift_start(self.velocity)  # pointsource.cpp:67-73

    def redshift(self, V_or_dest, *args, **kw):
        return super().redshift(V_or_dest, *args, **kw)


class ImagePlane(Raytracer):
    """imageplane.h:27: ImagePlane(dist, inc, x0, xmax, dx, y0, ymax, dy, spin, phi, precision)"""

    _BUNDLES = False

    def __init__(self, dist, inc, x0, xmax, dx, y0, ymax, dy, spin, phi, precision=A.PRECISION, eps_frac=0.01,
                 ctx=None, math_mode=A.MATH_EXACT):
        self.desc = imageplane_desc(dist, inc, x0, xmax, dx, y0, ymax, dy, phi, self._BUNDLES, eps_frac)
        nx, ny = C.c_int32(), C.c_int32()
        n = _lib.load().kt_imageplane_count(C.byref(self.desc), C.byref(nx), C.byref(ny))
        if n < 0:
            raise KerrtraceError(_lib.load().kt_last_error().decode())
        self.Nx, self.Ny = nx.value, ny.value
        self.m_x0, self.m_dx, self.m_y0, self.m_dy = x0, dx, y0, dy
        super().__init__(n, -1 * spin, precision, ctx=ctx, math_mode=math_mode)  # imageplane.cpp:12 negates the spin
        check(self.L.kt_rayset_init_imageplane(self._rs, C.byref(self._tr), C.byref(self.desc)))

    def redshift_start(self, *args, **kw):
        if args or kw:
            return super().redshift_start(*args, **kw)
        return super().redshift_start(0, True)  # imageplane.cpp:124-130

    def redshift(self, *args, **kw):
        # ImagePlane::redshift(bool projradius) -> Raytracer::redshift(-1, true, projradius) (imageplane.cpp:133-140);
        # apps call redshift(-1.0, true), which binds to the base overload through the using-declaration
        if len(args) == 1 and isinstance(args[0], bool) and not kw:
            return super().redshift(-1, True, args[0])
        return super().redshift(*args, **kw)

    def get_x_index(self, ix):
        return int(ix / self.Ny)

    def get_y_index(self, ix):
        return ix % self.Ny

    def ray_x(self, ix):
        return self.m_x0 + self.get_x_index(ix) * self.m_dx

    def ray_y(self, ix):
        return self.m_y0 + self.get_y_index(ix) * self.m_dy


class ImagePlaneBundles(ImagePlane):
    """imageplane_bundles.h:140: five rays (centre, E, W, N, S) per pixel."""

    _BUNDLES = True
    RAYS_PER_BUNDLE = 5

    def __init__(self, dist, inc_deg, x0, xmax, dx, y0, ymax, dy, spin, phi, precision=A.PRECISION, eps_frac=0.01,
                 ctx=None, math_mode=A.MATH_EXACT):
        super().__init__(dist, inc_deg, x0, xmax, dx, y0, ymax, dy, spin, phi, precision, eps_frac, ctx, math_mode)
        self.eps_x, self.eps_y = eps_frac * dx, eps_frac * dy

    def centre_ray(self, ix, iy):
        return (ix * self.Ny + iy) * 5 + 0

    def east_ray(self, ix, iy):
        return (ix * self.Ny + iy) * 5 + 1

    def west_ray(self, ix, iy):
        return (ix * self.Ny + iy) * 5 + 2

    def north_ray(self, ix, iy):
        return (ix * self.Ny + iy) * 5 + 3

    def south_ray(self, ix, iy):
        return (ix * self.Ny + iy) * 5 + 4


# ---------------------------------------------------------------------------------------------------------
# fused paths
# ---------------------------------------------------------------------------------------------------------
def emissivity_bins(spin, Nr=100, r_min=-1.0, r_disc=500.0, logbin=True, gamma=2.0, src=None, hit_filter=A.HIT_Z,
                    redshift_mode=A.REDSHIFT_V, theta_lim=M_PI_2, V=-1.0):
    """Bin set-up of emissivity.cpp:57-61 (r_isco through the float-rounded kerr_isco)."""
    b = A.EmisBins()
    r_isco = kerr_isco(spin, +1)
    if r_min < 0:
        r_min = r_isco
    b.Nr, b.logbin = Nr, int(logbin)
    b.r_min = r_min
    b.dr = math.exp(math.log(r_disc / r_min) / Nr) if logbin else (r_disc - r_min) / Nr
    b.r_isco, b.gamma = r_isco, gamma
    if src is not None:  # long num_primary_rays = ((cmax-c0)/dc) * ((bmax-b0)/db): emissivity.cpp:61
        b.n_primary = float(int(((src.cosalphamax - src.cosalpha0) / src.dcosalpha) * ((src.betamax - src.beta0) / src.dbeta)))
    else:
        b.n_primary = 1.0
    b.hit_filter, b.redshift_mode, b.theta_lim, b.V = hit_filter, redshift_mode, theta_lim, V
    return b


def bin_radii(bins):
    i = np.arange(bins.Nr)
    return bins.r_min * np.power(bins.dr, i) if bins.logbin else bins.r_min + i * bins.dr


def emissivity(src, spin, bins, method=Integrator.RK4, stop=None, r_max=1000.0, tol=A.TOL, steplim=-1, rk45_tol=None,
               shard=None, ctx=None, math_mode=A.MATH_EXACT, tracer=None, flags=0):
    """Whole emissivity app in one kernel: PointSource -> trace -> redshift -> radial histogram (raw sums)."""
    ctx = ctx or get_context()
    tr = tracer or make_tracer(spin, tol)
    if rk45_tol is not None:
        tr.rk45_tol = rk45_tol
    stop = stop if stop is not None else thetalim_stop(M_PI_2)
    if isinstance(stop, RayDestination):
        stop = stop._stop()
    run = make_run(method, r_max, steplim, math_mode, flags)
    out = A.EmisResult(bins.Nr)
    check(ctx.L.kt_pointsource_emissivity(ctx.h, C.byref(src), C.byref(tr), C.byref(stop), C.byref(run), C.byref(bins),
                                          C.byref(shard) if shard is not None else None, C.byref(out.c)))
    return out


def emissivity_dev(src, spin, bins, dev_ptr, method=Integrator.RK4, stop=None, r_max=1000.0, tol=A.TOL, steplim=-1,
                   rk45_tol=None, shard=None, ctx=None, math_mode=A.MATH_EXACT, tracer=None):
    """Asynchronous variant accumulating into a caller-owned device buffer double[5*Nr + 8] (e.g. a torch tensor)."""
    ctx = ctx or get_context()
    tr = tracer or make_tracer(spin, tol)
    if rk45_tol is not None:
        tr.rk45_tol = rk45_tol
    stop = stop if stop is not None else thetalim_stop(M_PI_2)
    if isinstance(stop, RayDestination):
        stop = stop._stop()
    run = make_run(method, r_max, steplim, math_mode)
    check(ctx.L.kt_pointsource_emissivity_dev(ctx.h, C.byref(src), C.byref(tr), C.byref(stop), C.byref(run), C.byref(bins),
                                              C.byref(shard) if shard is not None else None, C.c_void_p(dev_ptr)))


def image_bins(plane, spin, r_disc, img_Nx=None, img_Ny=None, flip_image=True, q=(3, 4, 3, 10, 3), hit_filter=A.HIT_Z,
               redshift_mode=A.REDSHIFT_V, reverse=True, theta_lim=M_PI_2, V=-1.0, Nx=None, Ny=None):
    """Bin set-up of imageplane_disc_image.cpp:78-108."""
    b = A.ImageBins()
    nx = Nx if Nx is not None else int(round((plane.xmax - plane.x0) / plane.dx))
    b.img_Nx = img_Nx or nx
    b.img_Ny = img_Ny or b.img_Nx  # the app defaults img_Ny to img_Nx (imageplane_disc_image.cpp:57)
    b.x0, b.y0 = plane.x0, plane.y0
    b.img_dx = (plane.xmax - plane.x0) / b.img_Nx
    b.img_dy = (plane.ymax - plane.y0) / b.img_Ny
    b.flip_image = int(flip_image)
    b.r_isco, b.r_disc = kerr_isco(spin, +1), r_disc
    b.q1, b.rb1, b.q2, b.rb2, b.q3 = q
    b.hit_filter, b.redshift_mode, b.reverse, b.theta_lim, b.V = hit_filter, redshift_mode, int(reverse), theta_lim, V
    return b


def disc_image(plane, spin, bins, method=Integrator.RK4, stop=None, r_max=None, precision=A.PRECISION, steplim=-1,
               rk45_tol=None, shard=None, ctx=None, math_mode=A.MATH_EXACT):
    """Whole imageplane_disc_image app in one kernel: ImagePlane -> trace -> redshift -> 7 sum maps."""
    ctx = ctx or get_context()
    tr = make_tracer(-1 * spin, precision)
    if rk45_tol is not None:
        tr.rk45_tol = rk45_tol
    stop = stop if stop is not None else thetalim_stop(M_PI_2)
    if isinstance(stop, RayDestination):
        stop = stop._stop()
    run = make_run(method, 1.1 * plane.dist if r_max is None else r_max, steplim, math_mode)
    out = A.ImageResult(bins.img_Nx, bins.img_Ny)
    check(ctx.L.kt_imageplane_image(ctx.h, C.byref(plane), C.byref(tr), C.byref(stop), C.byref(run), C.byref(bins),
                                    C.byref(shard) if shard is not None else None, C.byref(out.c)))
    return out


def plane_rays(plane, spin, stop, method=Integrator.RK4, r_max=None, precision=A.PRECISION, steplim=-1, rk45_tol=None,
               redshift_mode=A.REDSHIFT_DEST, V=-1.0, range_phi=False, ctx=None, math_mode=A.MATH_EXACT):
    """Trace an image plane (or bundles) and return compact per-ray results (caustic_discplane.cpp:209-217)."""
    ctx = ctx or get_context()
    tr = make_tracer(-1 * spin, precision)
    if rk45_tol is not None:
        tr.rk45_tol = rk45_tol
    if isinstance(stop, RayDestination):
        stop = stop._stop()
    run = make_run(method, 1.1 * plane.dist if r_max is None else r_max, steplim, math_mode)
    n = ctx.L.kt_imageplane_count(C.byref(plane), None, None)
    if n < 0:
        raise KerrtraceError(ctx.L.kt_last_error().decode())
    out = np.zeros(n, dtype=A.RAY_MIN_DTYPE)
    check(ctx.L.kt_imageplane_rays(ctx.h, C.byref(plane), C.byref(tr), C.byref(stop), C.byref(run), redshift_mode, V,
                                   int(range_phi), out.ctypes.data_as(C.c_void_p), n))
    return out


def caustic_discplane(dist, incl, spin, r_disc, Nx, Ny=None, x0=None, xmax=None, y0=None, ymax=None, plane_phi0=0.0,
                      method=Integrator.RK45, rk45_tol=1e-8, precision=A.PRECISION, bundle_eps_frac=0.01, suppress=True,
                      steplim=-1, ctx=None, math_mode=A.MATH_EXACT):
    """The whole of caustic_discplane.cpp:56-493 with its par-file parameter names: an (Nx + 1) x (Ny + 1) image plane (ray
    bundles when bundle_eps_frac > 0, else single rays and grid-neighbour differences) traced to the disc annulus
    [r_isco, r_disc] with DiscWithISCODestination, redshift(dest, reverse), the nine per-pixel maps, the Jacobian determinant
    of the lens mapping and the suppression pass -- one C ABI call, everything on the device.  -> A.CausticResult"""
    ctx = ctx or get_context()
    x0 = -1 * r_disc if x0 is None else x0
    xmax = r_disc if xmax is None else xmax
    y0 = x0 if y0 is None else y0
    ymax = xmax if ymax is None else ymax
    Ny = Nx if Ny is None else Ny
    dx, dy = (xmax - x0) / Nx, (ymax - y0) / Ny
    use_bundles = bundle_eps_frac > 0.0
    plane = imageplane_desc(dist, incl, x0, xmax, dx, y0, ymax, dy, plane_phi0, use_bundles, bundle_eps_frac if use_bundles else 0.01)
    r_isco = kerr_isco(spin, +1)
    tr = make_tracer(-1 * spin, precision)
    tr.rk45_tol = rk45_tol
    stop = DiscWithISCODestination(r_isco, r_disc)._stop()
    run = make_run(method, 1.1 * dist, steplim, math_mode)
    cfg = A.CausticCfg()
    cfg.img_Nx, cfg.img_Ny = Nx + 1, Ny + 1
    cfg.r_isco, cfg.r_disc, cfg.dx, cfg.dy = r_isco, r_disc, dx, dy
    cfg.redshift_mode, cfg.suppress, cfg.V = A.REDSHIFT_DEST, int(suppress), -1.0
    out = A.CausticResult(Nx + 1, Ny + 1)
    check(ctx.L.kt_caustic_discplane(ctx.h, C.byref(plane), C.byref(tr), C.byref(stop), C.byref(run), C.byref(cfg), C.byref(out.c)))
    return out
