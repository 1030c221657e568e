"""ctypes bindings of the two CPU checkers (test infrastructure; never imported by the product):

* ``Oracle``  -- oracle/liboracle.so, the plain-C restatement (oracle/kerr_oracle.c)
* ``Ref``     -- oracle/_ref/libref_oracle.so, the UNMODIFIED reference compiled by oracle/Makefile from
                 /root/reference (present in the build container; the prebuilt .so travels to the GPU box)

Both are driven with the same kt_* POD structs as the product so the parity tests read the same for all
three implementations.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from raytrace_cpu_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_oracle.so")


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


def _rays_ptr(rays):
    assert rays.dtype == A.RAY_DTYPE and rays.flags.c_contiguous
    return rays.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.orc_kerr_horizon.restype = C.c_double
        L.orc_kerr_horizon.argtypes = [C.c_double]
        L.orc_kerr_isco.restype = C.c_double
        L.orc_kerr_isco.argtypes = [C.c_double, C.c_int]
        L.orc_tracer_init.argtypes = [C.POINTER(A.Tracer), C.c_double, C.c_double]
        L.orc_pointsource_count.restype = C.c_int64
        L.orc_pointsource_count.argtypes = [C.POINTER(A.PointSourceDesc), A.c_int32_p, A.c_int32_p]
        L.orc_imageplane_count.restype = C.c_int64
        L.orc_imageplane_count.argtypes = [C.POINTER(A.ImagePlaneDesc), A.c_int32_p, A.c_int32_p]
        L.orc_pointsource_init.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.POINTER(A.PointSourceDesc)]
        L.orc_imageplane_init.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.POINTER(A.ImagePlaneDesc)]
        L.orc_redshift_start.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.c_double, C.c_int, C.c_int]
        L.orc_trace.restype = C.c_int64
        L.orc_trace.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.POINTER(A.Stop), C.POINTER(A.Run), C.c_int]
        L.orc_trace_paths.restype = C.c_int64
        L.orc_trace_paths.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.POINTER(A.Stop), C.POINTER(A.Run),
                                      C.POINTER(A.PathOpts), A.c_int64_p, A.c_double_p]
        L.orc_redshift.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_redshift_dest.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer), C.POINTER(A.Stop), C.c_int]
        L.orc_range_phi.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double]
        L.orc_calculate_momentum.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.Tracer)]
        L.orc_bin_emissivity.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.EmisBins), C.POINTER(A.EmisOut)]
        L.orc_bin_image.argtypes = [C.c_void_p, C.c_int64, C.POINTER(A.ImageBins), C.POINTER(A.ImageOut)]
        L.orc_caustic_maps.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_double, C.c_double, C.POINTER(A.CausticCfg),
                                       C.POINTER(A.CausticOut)]
        L.orc_momentum_from_consts.argtypes = [A.c_double_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                               C.c_double, C.c_double, C.c_double]

    def tracer(self, spin, precision=A.PRECISION):
        tr = A.Tracer()
        self.L.orc_tracer_init(C.byref(tr), spin, precision)
        return tr

    def kerr_isco(self, a, sign=+1):
        return self.L.orc_kerr_isco(a, sign)

    def kerr_horizon(self, a):
        return self.L.orc_kerr_horizon(a)

    def pointsource_count(self, src):
        na, nb = C.c_int32(), C.c_int32()
        n = self.L.orc_pointsource_count(C.byref(src), C.byref(na), C.byref(nb))
        return n, na.value, nb.value

    def imageplane_count(self, ip):
        nx, ny = C.c_int32(), C.c_int32()
        n = self.L.orc_imageplane_count(C.byref(ip), C.byref(nx), C.byref(ny))
        return n, nx.value, ny.value

    def pointsource_rays(self, tr, src):
        n, _, _ = self.pointsource_count(src)
        rays = np.zeros(n, dtype=A.RAY_DTYPE)
        self.L.orc_pointsource_init(_rays_ptr(rays), n, C.byref(tr), C.byref(src))
        return rays

    def imageplane_rays(self, tr, ip):
        n, _, _ = self.imageplane_count(ip)
        rays = np.zeros(n, dtype=A.RAY_DTYPE)
        self.L.orc_imageplane_init(_rays_ptr(rays), n, C.byref(tr), C.byref(ip))
        return rays

    def redshift_start(self, rays, tr, V, reverse=False, projradius=False):
        self.L.orc_redshift_start(_rays_ptr(rays), len(rays), C.byref(tr), V, int(reverse), int(projradius))

    def trace(self, rays, tr, stop, run, nthreads=0):
        return self.L.orc_trace(_rays_ptr(rays), len(rays), C.byref(tr), C.byref(stop), C.byref(run), nthreads)

    def trace_paths(self, rays, tr, stop, run, opts):
        """run_raytrace with an outfile: rays updated in place; returns (offsets[n + 1], points[total, 4])."""
        n = len(rays)
        offsets = np.zeros(n + 1, dtype=np.int64)
        scratch = rays.copy()
        total = self.L.orc_trace_paths(_rays_ptr(scratch), n, C.byref(tr), C.byref(stop), C.byref(run), C.byref(opts),
                                       offsets.ctypes.data_as(A.c_int64_p), None)
        points = np.zeros((total, 4))
        self.L.orc_trace_paths(_rays_ptr(rays), n, C.byref(tr), C.byref(stop), C.byref(run), C.byref(opts),
                               offsets.ctypes.data_as(A.c_int64_p), points.ctypes.data_as(A.c_double_p))
        return offsets, points

    def redshift(self, rays, tr, V, reverse=False, projradius=False, motion=0):
        self.L.orc_redshift(_rays_ptr(rays), len(rays), C.byref(tr), V, int(reverse), int(projradius), motion)

    def redshift_dest(self, rays, tr, stop, reverse=False):
        self.L.orc_redshift_dest(_rays_ptr(rays), len(rays), C.byref(tr), C.byref(stop), int(reverse))

    def range_phi(self, rays, lo=-np.pi, hi=np.pi):
        self.L.orc_range_phi(_rays_ptr(rays), len(rays), lo, hi)

    def calculate_momentum(self, rays, tr):
        self.L.orc_calculate_momentum(_rays_ptr(rays), len(rays), C.byref(tr))

    def bin_emissivity(self, rays, bins):
        out = A.EmisResult(bins.Nr)
        self.L.orc_bin_emissivity(_rays_ptr(rays), len(rays), C.byref(bins), C.byref(out.c))
        return out

    def bin_image(self, rays, bins):
        out = A.ImageResult(bins.img_Nx, bins.img_Ny)
        self.L.orc_bin_image(_rays_ptr(rays), len(rays), C.byref(bins), C.byref(out.c))
        return out

    def momentum_from_consts(self, k, h, Q, rs, ts, r, theta, a):
        p = np.zeros(4)
        self.L.orc_momentum_from_consts(A.dptr(p), k, h, Q, rs, ts, r, theta, a)
        return p

    def caustic_maps(self, rays, plane_Ny, bundles, eps_x, eps_y, cfg):
        out = A.CausticResult(cfg.img_Nx, cfg.img_Ny)
        self.L.orc_caustic_maps(_rays_ptr(rays), len(rays), plane_Ny, int(bundles), eps_x, eps_y, C.byref(cfg), C.byref(out.c))
        return out


def parse_paths(text, steps_before, steplim):
    """Text written by run_raytrace(..., outfile) -> (offsets[n + 1], points[total, 4]).  Every propagated ray (steps >= 0 and
    < steplim before the call) contributes its point lines followed by two blank lines; skipped rays contribute nothing."""
    blocks, cur, blanks = [], [], 0
    for line in text.split("\n")[:-1]:
        if line.strip():
            cur.append([float(x) for x in line.split()])
        else:
            blanks += 1
            if blanks == 2:
                blocks.append(cur)
                cur, blanks = [], 0
    assert not cur and blanks == 0, "trailing partial block"
    live = [i for i, s in enumerate(steps_before) if 0 <= s < steplim]
    assert len(blocks) == len(live), (len(blocks), len(live))
    n = len(steps_before)
    counts = np.zeros(n, dtype=np.int64)
    for i, b in zip(live, blocks):
        counts[i] = len(b)
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts)
    pts = np.array([row for b in blocks for row in b], dtype=np.float64).reshape(-1, 4)
    return offsets, pts


def ref_available():
    return os.path.exists(REF_SO)


def _stop_par(stop):
    arr = (C.c_double * 6)(stop.theta_lim, stop.r_isco, stop.r_out, stop.incl, stop.phi0, stop.z_s)
    return arr


class RefSource:
    """One live reference object (PointSource / ImagePlane / ImagePlaneBundles <double>)."""

    def __init__(self, L, handle):
        self.L, self.h = L, handle

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    def count(self):
        return self.L.ref_count(self.h)

    def rays(self, copy=True):
        n = self.count()
        buf = (C.c_char * (n * 144)).from_address(self.L.ref_rays(self.h))
        arr = np.frombuffer(buf, dtype=A.RAY_DTYPE, count=n)
        return arr.copy() if copy else arr

    def set_rk45_tol(self, tol):
        self.L.ref_set_rk45_tol(self.h, tol)

    def set_max_tstep(self, max_, rlim):
        self.L.ref_set_max_tstep(self.h, max_, rlim)

    def set_max_phistep(self, max_):
        self.L.ref_set_max_phistep(self.h, max_)

    def redshift_start(self):
        self.L.ref_source_redshift_start(self.h)

    def redshift_start_V(self, V, reverse=False, projradius=False):
        self.L.ref_redshift_start(self.h, V, int(reverse), int(projradius))

    def run(self, stop, run):
        if stop.kind == A.STOP_THETALIM:
            self.L.ref_run_thetalim(self.h, run.integrator, stop.theta_lim, run.r_max, run.steplim if run.steplim > 0 else -1)
        else:
            self.L.ref_run_dest(self.h, run.integrator, stop.kind, _stop_par(stop), run.r_max,
                                run.steplim if run.steplim > 0 else -1)
        return self.L.ref_last_trace_seconds(self.h)

    def run_paths(self, stop, run, opts, precision=17):
        """run_raytrace with an outfile through the reference's own TextOutput; the text is parsed back into
        (offsets[n + 1], points[total, 4]) -- with 17 digits the columns are the exact doubles.  `before` steps decide
        which rays were propagated (each of those ends with two blank lines, raytracer.cpp:99)."""
        import tempfile
        before = self.rays()["steps"].copy()
        steplim = run.steplim if run.steplim > 0 else (A.RK45_STEPLIM if run.integrator == A.RK45 else A.STEPLIM)
        with tempfile.TemporaryDirectory() as tmp:
            fn = os.path.join(tmp, "paths.dat")
            if stop.kind == A.STOP_THETALIM:
                self.L.ref_run_thetalim_paths(self.h, run.integrator, stop.theta_lim, run.r_max, run.steplim if run.steplim > 0 else -1,
                                              fn.encode(), precision, opts.write_step, opts.write_rmax, opts.write_rmin, opts.write_cartesian)
            else:
                self.L.ref_run_dest_paths(self.h, run.integrator, stop.kind, _stop_par(stop), run.r_max,
                                          run.steplim if run.steplim > 0 else -1, fn.encode(), precision, opts.write_step,
                                          opts.write_rmax, opts.write_rmin, opts.write_cartesian)
            text = open(fn).read()
        return parse_paths(text, before, steplim)

    def redshift(self, V, reverse=False, projradius=False, motion=0):
        self.L.ref_redshift_V(self.h, V, int(reverse), int(projradius), motion)

    def redshift_dest(self, stop, reverse=False):
        self.L.ref_redshift_dest(self.h, stop.kind, _stop_par(stop), int(reverse))

    def range_phi(self, lo=-np.pi, hi=np.pi):
        self.L.ref_range_phi(self.h, lo, hi)

    def calculate_momentum(self):
        self.L.ref_calculate_momentum(self.h)


class Ref:
    def __init__(self, path=REF_SO):
        L = self.L = C.CDLL(path)
        d = C.c_double
        L.ref_sizeof_ray.restype = C.c_int
        L.ref_set_num_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        L.ref_pointsource_create.restype = C.c_void_p
        L.ref_pointsource_create.argtypes = [A.c_double_p] + [d] * 10
        L.ref_imageplane_create.restype = C.c_void_p
        L.ref_imageplane_create.argtypes = [d] * 11
        L.ref_bundles_create.restype = C.c_void_p
        L.ref_bundles_create.argtypes = [d] * 12
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_count.restype = C.c_int64
        L.ref_count.argtypes = [C.c_void_p]
        L.ref_rays.restype = C.c_void_p
        L.ref_rays.argtypes = [C.c_void_p]
        L.ref_horizon.restype = d
        L.ref_horizon.argtypes = [C.c_void_p]
        L.ref_last_trace_seconds.restype = d
        L.ref_last_trace_seconds.argtypes = [C.c_void_p]
        L.ref_set_rk45_tol.argtypes = [C.c_void_p, d]
        L.ref_set_max_tstep.argtypes = [C.c_void_p, d, d]
        L.ref_set_max_phistep.argtypes = [C.c_void_p, d]
        L.ref_source_redshift_start.argtypes = [C.c_void_p]
        L.ref_redshift_start.argtypes = [C.c_void_p, d, C.c_int, C.c_int]
        L.ref_run_thetalim.argtypes = [C.c_void_p, C.c_int, d, d, C.c_int]
        L.ref_run_dest.argtypes = [C.c_void_p, C.c_int, C.c_int, A.c_double_p, d, C.c_int]
        L.ref_run_thetalim_paths.argtypes = [C.c_void_p, C.c_int, d, d, C.c_int, C.c_char_p, C.c_int, C.c_int, d, d, C.c_int]
        L.ref_run_dest_paths.argtypes = [C.c_void_p, C.c_int, C.c_int, A.c_double_p, d, C.c_int, C.c_char_p, C.c_int, C.c_int, d, d, C.c_int]
        L.ref_redshift_V.argtypes = [C.c_void_p, d, C.c_int, C.c_int, C.c_int]
        L.ref_redshift_dest.argtypes = [C.c_void_p, C.c_int, A.c_double_p, C.c_int]
        L.ref_range_phi.argtypes = [C.c_void_p, d, d]
        L.ref_calculate_momentum.argtypes = [C.c_void_p]
        L.ref_kerr_horizon.restype = d
        L.ref_kerr_horizon.argtypes = [d]
        L.ref_kerr_isco.restype = d
        L.ref_kerr_isco.argtypes = [d, C.c_int]
        L.ref_integrate_disc_area.restype = d
        L.ref_integrate_disc_area.argtypes = [d, d, d]
        L.ref_momentum_from_consts.argtypes = [A.c_double_p, d, d, d, C.c_int, C.c_int, d, d, d, d]
        assert L.ref_sizeof_ray() == 144

    def pointsource(self, src, spin, tol=A.TOL):
        pos = (C.c_double * 4)(*src.pos)
        h = self.L.ref_pointsource_create(pos, src.V, spin, tol, src.dcosalpha, src.dbeta, src.cosalpha0,
                                          src.cosalphamax, src.beta0, src.betamax, src.E)
        return RefSource(self.L, h)

    def imageplane(self, ip, spin, precision=A.PRECISION):
        """spin is the PHYSICAL spin (the reference negates it itself, imageplane.cpp:12)."""
        if ip.bundles:
            h = self.L.ref_bundles_create(ip.dist, ip.incl_deg, ip.x0, ip.xmax, ip.dx, ip.y0, ip.ymax, ip.dy, spin,
                                          ip.phi0, precision, ip.eps_frac)
        else:
            h = self.L.ref_imageplane_create(ip.dist, ip.incl_deg, ip.x0, ip.xmax, ip.dx, ip.y0, ip.ymax, ip.dy, spin,
                                             ip.phi0, precision)
        return RefSource(self.L, h)

    def set_num_threads(self, n):
        self.L.ref_set_num_threads(int(n))
        return self.L.ref_max_threads()

    def kerr_isco(self, a, sign=+1):
        return self.L.ref_kerr_isco(a, sign)

    def kerr_horizon(self, a):
        return self.L.ref_kerr_horizon(a)

    def integrate_disc_area(self, rmin, rmax, a):
        return self.L.ref_integrate_disc_area(rmin, rmax, a)

    def momentum_from_consts(self, k, h, Q, rs, ts, r, theta, a):
        p = np.zeros(4)
        self.L.ref_momentum_from_consts(A.dptr(p), k, h, Q, rs, ts, r, theta, 0.0, a)
        return p


# ---- small constructors shared by the tests -------------------------------------------------------------

def make_pointsource(pos, dcosalpha, dbeta, V=0.0, E=1.0, cosalpha0=-0.999999, cosalphamax=0.995,
                     beta0=-0.995 * np.pi, betamax=np.pi):
    s = A.PointSourceDesc()
    s.pos[:] = pos
    s.V, s.E = V, E
    s.dcosalpha, s.dbeta = dcosalpha, dbeta
    s.cosalpha0, s.cosalphamax, s.beta0, s.betamax = cosalpha0, cosalphamax, beta0, betamax
    return s


def make_imageplane(dist, incl_deg, x0, xmax, nx, y0=None, ymax=None, ny=None, phi0=0.0, bundles=False, eps_frac=0.01):
    y0 = x0 if y0 is None else y0
    ymax = xmax if ymax is None else ymax
    ny = nx if ny is None else ny
    p = A.ImagePlaneDesc()
    p.dist, p.incl_deg = dist, incl_deg
    p.x0, p.xmax, p.dx = x0, xmax, (xmax - x0) / nx
    p.y0, p.ymax, p.dy = y0, ymax, (ymax - y0) / ny
    p.phi0 = phi0
    p.bundles = int(bundles)
    p.eps_frac = eps_frac
    return p


def make_stop(kind=A.STOP_THETALIM, theta_lim=np.pi / 2, r_isco=0.0, r_out=-1.0, incl=0.0, phi0=0.0, z_s=0.0):
    s = A.Stop()
    s.kind, s.theta_lim, s.r_isco, s.r_out, s.incl, s.phi0, s.z_s = kind, theta_lim, r_isco, r_out, incl, phi0, z_s
    return s


def make_path_opts(write_step=10, write_cartesian=True, write_rmax=-1.0, write_rmin=-1.0):
    o = A.PathOpts()
    o.write_step, o.write_cartesian, o.write_rmax, o.write_rmin = write_step, int(write_cartesian), write_rmax, write_rmin
    return o


def make_run(integrator=A.RK4, r_max=1000.0, steplim=-1, math_mode=A.MATH_EXACT):
    r = A.Run()
    r.integrator, r.steplim, r.r_max, r.math_mode = integrator, steplim, r_max, math_mode
    return r


def same_bits(a, b):
    """Bitwise equality of two float arrays, treating any-NaN == any-NaN."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return ((a == b) | (np.isnan(a) & np.isnan(b)))
