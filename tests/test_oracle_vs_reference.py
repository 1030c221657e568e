"""Pins the CPU oracle (oracle/kerr_oracle.c) to the reference.

* against the committed golden fixtures (outputs of the unmodified reference, scripts/make_golden.py):
  BIT-EXACT for every ray field, on every machine;
* against the live reference build oracle/_ref/libref_oracle.so where it exists (build container; the prebuilt
  .so also travels to the GPU box): bit-exact on larger grids and on the API corner cases;
* against the text output of the reference's own emissivity_rd application (binning loop + area division).
"""
import os

import numpy as np
import pytest

from raytrace_cpu_b200 import _abi as A
import oracle_bindings as ob
from golden_cases import CASES, PATH_CASES, R_ISCO, SPIN, build_source, tracer_for

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLOAT_FIELDS = [n for n in A.RAY_DTYPE.names if A.RAY_DTYPE[n].kind == "f"]
INT_FIELDS = [n for n in A.RAY_DTYPE.names if A.RAY_DTYPE[n].kind == "i"]


@pytest.fixture(scope="module")
def orc():
    return ob.Oracle()


def assert_bit_identical(got, want, skip=()):
    assert (got["steps"] == want["steps"]).all()
    assert (got["status"] == want["status"]).all()
    used = want["steps"] != -1
    for f in INT_FIELDS:
        if f not in skip:
            assert (got[f][used] == want[f][used]).all(), f
    for f in FLOAT_FIELDS:
        if f not in skip:
            bad = ~ob.same_bits(got[f][used], want[f][used])
            assert not bad.any(), f"{f}: {int(bad.sum())} rays differ"


def run_oracle_case(orc, case):
    src = build_source(case)
    tr = tracer_for(case, orc.tracer)
    point = case["source"] == "point"
    rays = orc.pointsource_rays(tr, src) if point else orc.imageplane_rays(tr, src)
    orc.redshift_start(rays, tr, src.V if point else 0.0, not point)
    init = rays.copy()
    stop = ob.make_stop(**case["stop"])
    run = ob.make_run(case["integrator"], r_max=case["r_max"], steplim=case.get("steplim", -1))
    if "paths" in case:
        offsets, points = orc.trace_paths(rays, tr, stop, run, ob.make_path_opts(**case["paths"]))
        return init, rays, offsets, points
    orc.trace(rays, tr, stop, run)
    if case["redshift"] == "V":
        orc.redshift(rays, tr, -1.0, not point)
    else:
        orc.redshift_dest(rays, tr, stop, not point)
    if case.get("range_phi", True):
        orc.range_phi(rays)
    return init, rays


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_fixture_bit_for_bit(orc, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    init, final = run_oracle_case(orc, CASES[name])
    assert len(final) == len(g["final"])
    assert_bit_identical(init, g["init"], skip=("redshift",))     # redshift is not yet defined before the trace
    assert_bit_identical(final, g["final"])


@pytest.mark.parametrize("name", sorted(PATH_CASES))
def test_oracle_reproduces_golden_ray_paths_bit_for_bit(orc, name):
    """run_raytrace with an outfile (raytracer.cpp:86-100 and the write blocks of the five propagators): the points the
    reference wrote through its TextOutput at 17 significant digits, the per-ray point counts and the final rays."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    init, final, offsets, points = run_oracle_case(orc, PATH_CASES[name])
    assert_bit_identical(init, g["init"], skip=("redshift",))
    assert_bit_identical(final, g["final"], skip=("redshift",))
    assert (offsets == g["offsets"]).all() and offsets[-1] > 1000
    assert points.shape == g["points"].shape and ob.same_bits(points, g["points"]).all()


CAUSTIC_FITS = {"DET_J": "det_J", "SIGN_J": "sign_J", "ORDER": "order", "HIT": "hit", "RADIUS": "radius", "PHI": "phi",
                "X_DISC": "x_disc", "Y_DISC": "y_disc", "REDSHIFT": "redshift"}


def oracle_caustic(orc, par):
    """caustic_discplane.cpp:104-493 through the oracle: plane -> redshift_start -> trace -> redshift(dest, reverse) -> maps."""
    import raytrace_cpu_b200 as K
    dx, dy = (par["xmax"] - par["x0"]) / par["Nx"], (par["ymax"] - par["y0"]) / par["Ny"]
    bundles = par["bundle_eps_frac"] > 0
    ip = K.imageplane_desc(par["dist"], par["incl"], par["x0"], par["xmax"], dx, par["y0"], par["ymax"], dy, par["plane_phi0"],
                           bundles, par["bundle_eps_frac"] if bundles else 0.01)
    tr = orc.tracer(-par["spin"], par["precision"])
    tr.rk45_tol = par["rk45_tol"]
    rays = orc.imageplane_rays(tr, ip)
    orc.redshift_start(rays, tr, 0.0, True)
    r_isco = orc.kerr_isco(par["spin"], +1)
    stop = ob.make_stop(A.STOP_ISCODISC, r_isco=r_isco, r_out=par["r_disc"])
    orc.trace(rays, tr, stop, ob.make_run(A.RK45 if par["rk45"] else A.RK4, r_max=1.1 * par["dist"]))
    orc.redshift_dest(rays, tr, stop, True)
    cfg = A.CausticCfg()
    cfg.img_Nx, cfg.img_Ny = par["Nx"] + 1, par["Ny"] + 1
    cfg.r_isco, cfg.r_disc, cfg.dx, cfg.dy = r_isco, par["r_disc"], dx, dy
    cfg.redshift_mode, cfg.suppress, cfg.V = A.REDSHIFT_DEST, 1, -1.0
    _, _, plane_ny = orc.imageplane_count(ip)
    return orc.caustic_maps(rays, plane_ny, bundles, par["bundle_eps_frac"] * dx, par["bundle_eps_frac"] * dy, cfg)


@pytest.mark.parametrize("case", ["app_caustic_rk4", "app_caustic_rk45", "app_caustic_grid_rk4"])
def test_oracle_caustic_maps_match_the_reference_app_bit_for_bit(orc, case):
    """tests/golden/app_caustic_*.fits were written by the reference's own caustic_discplane binary (CPU build); every pixel of
    its nine extensions -- hit mask, disc coordinates, image order, redshift, Jacobian determinant (bundle and
    grid-neighbour variants), sentinel and suppressed pixels -- must come out of the oracle identically."""
    from fits_reader import by_extname, read_fits
    from golden_cases import APP_PARS, caustic_par
    hdus = read_fits(os.path.join(GOLDEN, case + ".fits"))
    want = by_extname(hdus)
    got = oracle_caustic(orc, caustic_par(APP_PARS["caustic_discplane"][case]))
    assert got.disc_count == hdus[0][0]["DISC_N"] > 100
    for ext, field in CAUSTIC_FITS.items():
        a, b = getattr(got, field).astype(np.float64), want[ext].T       # FITS frame[k * Nx + j] = data[j][k] (fits_output.h:170-181)
        assert a.shape == b.shape
        assert ob.same_bits(a, b).all(), ext
    assert (got.det_J == A.CAUSTIC_SENTINEL).sum() > 0 and np.isnan(got.det_J).sum() > 0


def test_golden_scalars(orc):
    s = np.load(os.path.join(GOLDEN, "scalars.npz"))
    for a, ip, ir, hz in zip(s["spins"], s["isco_pro"], s["isco_retro"], s["horizon"]):
        assert orc.kerr_isco(a, +1) == ip and orc.kerr_isco(a, -1) == ir
        assert orc.kerr_horizon(a) == hz or (np.isnan(hz) and np.isnan(orc.kerr_horizon(a)))
    assert orc.kerr_isco(SPIN) == R_ISCO


def test_oracle_binning_matches_reference_emissivity_rd_app(orc):
    """tests/golden/emissivity_rd_app.dat is the 7-column output of the reference's own emissivity_rd binary
    (text_output.h: scientific, precision 8).  Columns: r, area, N, flux/area, emis/area, <g>, <t>."""
    dat = np.genfromtxt(os.path.join(GOLDEN, "emissivity_rd_app.dat"))
    assert dat.shape == (30, 7)
    src = ob.make_pointsource([0, 10, 1e-3, 1.5707], 0.02, 0.02, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    tr = orc.tracer(SPIN, A.TOL)
    rays = orc.pointsource_rays(tr, src)
    orc.redshift_start(rays, tr, 0.0)
    stop = ob.make_stop(A.STOP_FLATDISC, theta_lim=1.5707963)
    orc.trace(rays, tr, stop, ob.make_run(A.RK4))
    orc.range_phi(rays)
    orc.redshift_dest(rays, tr, stop)
    import raytrace_cpu_b200 as K
    bins = K.emissivity_bins(SPIN, Nr=30, src=src, hit_filter=A.HIT_THETA, redshift_mode=A.REDSHIFT_DEST, theta_lim=1.5707963)
    out = orc.bin_emissivity(rays, bins)
    r_bins = K.api.bin_radii(bins)
    np.testing.assert_allclose(r_bins, dat[:, 0], rtol=2e-8)
    assert (out.N == dat[:, 2].astype(np.int64)).all()            # ray counts per annulus: exact
    area = dat[:, 1]
    full = out.N > 0
    np.testing.assert_allclose(out.flux[full] / area[full], dat[full, 3], rtol=2e-8)
    np.testing.assert_allclose(out.emis[full] / area[full], dat[full, 4], rtol=2e-8)
    np.testing.assert_allclose(out.redshift_sum[full] / out.N[full], dat[full, 5], rtol=2e-8)
    np.testing.assert_allclose(out.time_sum[full] / out.N[full], dat[full, 6], rtol=2e-8)
    assert np.isnan(dat[~full, 5]).all()                           # 0/0 in empty bins
    if ob.ref_available():                                         # the area column is the reference's integrate_disc_area
        ref = ob.Ref()
        dr = bins.dr
        want = np.array([ref.integrate_disc_area(r, r * dr, SPIN) for r in r_bins])
        np.testing.assert_allclose(want, area, rtol=2e-8)


# ---------------------------------------------------------------------------------------------------------
# live reference (skipped where oracle/_ref was not built)
# ---------------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref/libref_oracle.so not built")


@needs_ref
@pytest.mark.parametrize("integ", [A.RK4, A.RK45])
@pytest.mark.parametrize("stopk", [A.STOP_THETALIM, A.STOP_FLATDISC, A.STOP_ISCODISC])
def test_live_reference_lamppost_5167(orc, integ, stopk):
    """The reference's own emissivity_rk45_test grid (5 167 slots), every integrator x stop-surface pair."""
    ref = ob.Ref()
    src = ob.make_pointsource([0, 5, 1e-3, 0], 0.05, 0.05, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    tr = orc.tracer(SPIN, A.TOL)
    h = ref.pointsource(src, SPIN)
    rays = orc.pointsource_rays(tr, src)
    assert len(rays) == h.count() == 5167
    h.redshift_start()
    orc.redshift_start(rays, tr, 0.0)
    stop = ob.make_stop(stopk, r_isco=R_ISCO, r_out=50.0)
    run = ob.make_run(integ)
    h.run(stop, run)
    orc.trace(rays, tr, stop, run)
    h.range_phi()
    orc.range_phi(rays)
    if stopk == A.STOP_THETALIM:
        h.redshift(-1.0)
        orc.redshift(rays, tr, -1.0)
    else:
        h.redshift_dest(stop)
        orc.redshift_dest(rays, tr, stop)
    assert_bit_identical(rays, h.rays())


@needs_ref
def test_live_reference_redshift_variants_and_reentrancy(orc):
    ref = ob.Ref()
    src = ob.make_pointsource([0, 5, 1e-3, 0], 0.2, 0.2, V=1 / (SPIN + 5 ** 1.5))
    tr = orc.tracer(SPIN, A.TOL)
    h = ref.pointsource(src, SPIN)
    rays = orc.pointsource_rays(tr, src)
    # V == -1 in redshift_start: the by-value V is overwritten by ray 0 (raytracer.cpp:391-394)
    h.redshift_start_V(-1.0, False, True)
    orc.redshift_start(rays, tr, -1.0, False, True)
    assert_bit_identical(rays, h.rays(), skip=("redshift",))
    # two passes with a small step limit: steps accumulate, stuck rays are flagged negative (raytracer.cpp:965-967)
    stop = ob.make_stop()
    for steplim in (200, 1000):
        run = ob.make_run(A.RK4, steplim=steplim)
        h.run(stop, run)
        orc.trace(rays, tr, stop, run)
        assert_bit_identical(rays, h.rays(), skip=("redshift",))
    assert (rays["steps"] < -1).any()
    for V, rev, pj, mo in [(-1.0, False, True, 0), (0.05, False, False, 0), (-0.3, False, False, 1), (0.2, True, False, 1)]:
        h.redshift(V, rev, pj, mo)
        orc.redshift(rays, tr, V, rev, pj, mo)
        assert_bit_identical(rays, h.rays())
    h.calculate_momentum()
    orc.calculate_momentum(rays, tr)
    assert_bit_identical(rays, h.rays())


@needs_ref
@pytest.mark.parametrize("tl", [-0.3, 0.0])
@pytest.mark.parametrize("integ", [A.RK4, A.RK45])
def test_live_reference_negative_and_zero_thetalim(orc, tl, integ):
    ref = ob.Ref()
    src = ob.make_pointsource([0, 8, 1.2, 0], 0.2, 0.2)
    tr = orc.tracer(SPIN, A.TOL)
    h = ref.pointsource(src, SPIN)
    rays = orc.pointsource_rays(tr, src)
    h.redshift_start()
    orc.redshift_start(rays, tr, 0.0)
    stop = ob.make_stop(A.STOP_THETALIM, theta_lim=tl)
    run = ob.make_run(integ, r_max=200.0, steplim=20000)
    h.run(stop, run)
    orc.trace(rays, tr, stop, run)
    assert_bit_identical(rays, h.rays(), skip=("redshift",))


SPIN_SWEEP = (0.0, 0.5, -0.9, 1.0)     # Schwarzschild, moderate, retrograde, extremal (horizon = isco = 1, Delta has a double root)


def spin_sweep_setup(orc, spin, stopk):
    """An off-axis source (h != 0 for every ray) at r = 6, theta = 0.7 above a disc with the spin's own ISCO."""
    src = ob.make_pointsource([0, 6, 0.7, 0.3], 0.1, 0.1, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    stop = ob.make_stop(stopk, r_isco=orc.kerr_isco(spin, +1), r_out=50.0)
    return src, stop


def oracle_spin_sweep(orc, spin, integ, stopk):
    src, stop = spin_sweep_setup(orc, spin, stopk)
    tr = orc.tracer(spin, A.TOL)
    rays = orc.pointsource_rays(tr, src)
    orc.redshift_start(rays, tr, 0.0)
    orc.trace(rays, tr, stop, ob.make_run(integ, r_max=200.0, steplim=20000))
    orc.range_phi(rays)
    if stopk == A.STOP_THETALIM:
        orc.redshift(rays, tr, -1.0)
    else:
        orc.redshift_dest(rays, tr, stop)
    return rays


@needs_ref
@pytest.mark.parametrize("integ,stopk", [(A.EULER, A.STOP_THETALIM), (A.RK4, A.STOP_THETALIM), (A.RK4, A.STOP_ISCODISC),
                                         (A.RK45, A.STOP_THETALIM), (A.RK45, A.STOP_ISCODISC)])
@pytest.mark.parametrize("spin", SPIN_SWEEP)
def test_live_reference_spin_sweep(orc, spin, stopk, integ):
    """Every other fixture is a = +-0.998; the restatement must follow the reference at any spin, a = 0 and |a| = 1 included."""
    ref = ob.Ref()
    src, stop = spin_sweep_setup(orc, spin, stopk)
    h = ref.pointsource(src, spin)
    assert ref.kerr_isco(spin, +1) == orc.kerr_isco(spin, +1) and ref.kerr_horizon(spin) == orc.kerr_horizon(spin)
    h.redshift_start()
    h.run(stop, ob.make_run(integ, r_max=200.0, steplim=20000))
    h.range_phi()
    if stopk == A.STOP_THETALIM:
        h.redshift(-1.0)
    else:
        h.redshift_dest(stop)
    rays = oracle_spin_sweep(orc, spin, integ, stopk)
    assert len(rays) == h.count() == 1334
    assert_bit_identical(rays, h.rays())
    st = rays["status"][rays["steps"] != -1]
    # the grid exercises every ending: the disc / theta limit, escape, and capture (RK45 stalls there until the step limit)
    assert (st & A.STATUS_DEST).any() and (st & A.STATUS_RLIM).any() and (st & (A.STATUS_HORIZON | A.STATUS_STEPLIM)).any()


# set_max_tstep(max, rlim) / set_max_phistep(max) (raytracer.h:187-196): tighter caps, a cap radius inside the grid, and both
# "off" spellings (0 and negative: the propagators test `> 0`, raytracer.cpp:861-866)
STEP_CAPS = {"tight": (0.25, 50.0, 0.02), "t_off": (0.0, 100.0, 0.1), "phi_off": (1.0, 100.0, -1.0), "both_off": (-1.0, 0.0, 0.0),
             "wide_rlim": (0.5, 1e4, 0.1)}


def oracle_step_caps(orc, caps, integ):
    src = ob.make_pointsource([0, 5, 1e-3, 0], 0.1, 0.1, cosalpha0=-0.995, cosalphamax=0.995, beta0=-np.pi, betamax=np.pi)
    tr = orc.tracer(SPIN, A.TOL)
    tr.max_tstep, tr.maxtstep_rlim, tr.max_phistep = STEP_CAPS[caps] if isinstance(caps, str) else caps
    rays = orc.pointsource_rays(tr, src)
    orc.redshift_start(rays, tr, 0.0)
    orc.trace(rays, tr, ob.make_stop(), ob.make_run(integ, r_max=300.0, steplim=50000))
    orc.range_phi(rays)
    orc.redshift(rays, tr, -1.0)
    return src, rays


@needs_ref
@pytest.mark.parametrize("integ", [A.RK4, A.RK45])
@pytest.mark.parametrize("caps", sorted(STEP_CAPS))
def test_live_reference_step_caps(orc, caps, integ):
    ref = ob.Ref()
    src, rays = oracle_step_caps(orc, caps, integ)
    h = ref.pointsource(src, SPIN)
    mt, rl, mp = STEP_CAPS[caps]
    h.set_max_tstep(mt, rl)
    h.set_max_phistep(mp)
    h.redshift_start()
    h.run(ob.make_stop(), ob.make_run(integ, r_max=300.0, steplim=50000))
    h.range_phi()
    h.redshift(-1.0)
    assert_bit_identical(rays, h.rays())
    if integ == A.RK4 and caps != "t_off":   # the caps change the step schedule (the default t cap alone never binds on this grid)
        _, base = oracle_step_caps(orc, (A.MAXDT, A.MAXDT_RLIM, A.MAXDPHI), integ)
        assert (base["steps"] != rays["steps"]).any()


@needs_ref
@pytest.mark.parametrize("bundles", [False, True])
def test_live_reference_image_plane(orc, bundles):
    ref = ob.Ref()
    ip = ob.make_imageplane(1000.0, 30.0, -20, 20, 32, bundles=bundles)
    tr = orc.tracer(-SPIN, A.PRECISION)
    h = ref.imageplane(ip, SPIN)
    rays = orc.imageplane_rays(tr, ip)
    rr = h.rays(copy=False)
    rr["rdot_flips"][:] = 0            # uninitialised in the reference
    rr["equatorial_crossings"][:] = 0
    h.redshift_start()
    orc.redshift_start(rays, tr, 0.0, True)
    assert_bit_identical(rays, h.rays(), skip=("redshift",))
    stop = ob.make_stop()
    run = ob.make_run(A.RK4, r_max=1100.0)
    h.run(stop, run)
    orc.trace(rays, tr, stop, run)
    h.redshift(-1.0, True)
    orc.redshift(rays, tr, -1.0, True)
    h.range_phi()
    orc.range_phi(rays)
    assert_bit_identical(rays, h.rays())
    if not bundles:
        assert np.isnan(rays["r"]).sum() == 1     # the exact-centre pixel (b = 0 -> asin(0/0)), SURVEY 8c


# image-plane geometries beyond the fixtures' (30 / 60 degrees, square grids, phi0 = 0): near face-on, near edge-on, a rotated
# observer, rectangular grids with dx != dy (the x coordinate steps by dy, imageplane.cpp:45), other spins
PLANE_SWEEP = {
    "faceon":   dict(spin=0.998, dist=800.0, incl=3.0, x0=-12.0, xmax=12.0, nx=21, ny=21, phi0=0.0),
    "edgeon":   dict(spin=0.998, dist=800.0, incl=87.0, x0=-12.0, xmax=12.0, nx=21, ny=21, phi0=0.0),
    "rotated":  dict(spin=0.5, dist=600.0, incl=45.0, x0=-10.0, xmax=10.0, nx=19, ny=19, phi0=0.7),
    "rect":     dict(spin=0.9, dist=600.0, incl=70.0, x0=-15.0, xmax=15.0, nx=30, ny=12, y0=-6.0, ymax=9.0, phi0=-2.0),
    "schwarz":  dict(spin=0.0, dist=600.0, incl=50.0, x0=-10.0, xmax=10.0, nx=19, ny=19, phi0=0.0),
    "retro":    dict(spin=-0.7, dist=600.0, incl=50.0, x0=-10.0, xmax=10.0, nx=19, ny=19, phi0=0.0),
}


def plane_sweep_setup(orc, name):
    c = PLANE_SWEEP[name]
    ip = ob.make_imageplane(c["dist"], c["incl"], c["x0"], c["xmax"], c["nx"], y0=c.get("y0"), ymax=c.get("ymax"), ny=c["ny"],
                            phi0=c["phi0"])
    # the tracer of an image plane stores the negated spin (imageplane.cpp:12); the disc's ISCO is that of the hole's own spin
    # a disc smaller than the field of view: the rays outside r_out cross the plane and go on to the horizon or to r_max
    stop = ob.make_stop(A.STOP_ISCODISC, r_isco=orc.kerr_isco(c["spin"], +1), r_out=8.0)
    return c, ip, stop


def oracle_plane_sweep(orc, name, integ):
    c, ip, stop = plane_sweep_setup(orc, name)
    tr = orc.tracer(-c["spin"], A.PRECISION)
    rays = orc.imageplane_rays(tr, ip)
    orc.redshift_start(rays, tr, 0.0, True)
    orc.trace(rays, tr, stop, ob.make_run(integ, r_max=1.1 * c["dist"], steplim=20000))
    orc.redshift_dest(rays, tr, stop, True)
    orc.range_phi(rays)
    return rays


@needs_ref
@pytest.mark.parametrize("integ", [A.RK4, A.RK45])
@pytest.mark.parametrize("name", sorted(PLANE_SWEEP))
def test_live_reference_plane_sweep(orc, name, integ):
    ref = ob.Ref()
    c, ip, stop = plane_sweep_setup(orc, name)
    h = ref.imageplane(ip, c["spin"])
    rr = h.rays(copy=False)
    rr["rdot_flips"][:] = 0            # uninitialised in the reference
    rr["equatorial_crossings"][:] = 0
    h.redshift_start()
    h.run(stop, ob.make_run(integ, r_max=1.1 * c["dist"], steplim=20000))
    h.redshift_dest(stop, True)
    h.range_phi()
    rays = oracle_plane_sweep(orc, name, integ)
    assert len(rays) == h.count() == (c["nx"] + 1) * (c["ny"] + 1)
    assert_bit_identical(rays, h.rays())
    st = rays["status"][rays["steps"] != -1]
    assert (st & A.STATUS_DEST).any() and (st & A.STATUS_RLIM).any()     # every geometry sees the disc and the sky behind it


# Two more corners of the parameter space: a moving (V != 0) source with E != 1 at a = 0.7 whose rays stop on a tilted
# FlatPlaneDestination (the overshoot stop of caustic_plane.cpp:151), and ImagePlaneBundles from a rotated, highly inclined
# observer at a = 0.5 onto an ISCO disc smaller than the field of view; RK45 runs at a non-default tolerance.
def corner_setup(orc, which):
    if which == "moving_flatplane":
        src = ob.make_pointsource([0, 4, 1.0, 0.2], 0.1, 0.1, V=0.05, E=2.5)
        stop = ob.make_stop(A.STOP_FLATPLANE, incl=1.0, phi0=0.4, z_s=30.0)
        return dict(point=True, spin=0.7, src=src, stop=stop, r_max=300.0)
    ip = ob.make_imageplane(700.0, 75.0, -9.0, 9.0, 13, bundles=True, phi0=1.1)
    stop = ob.make_stop(A.STOP_ISCODISC, r_isco=orc.kerr_isco(0.5, +1), r_out=7.0)
    return dict(point=False, spin=0.5, src=ip, stop=stop, r_max=770.0)


def oracle_corner(orc, which, integ):
    c = corner_setup(orc, which)
    tr = orc.tracer(c["spin"], A.TOL) if c["point"] else orc.tracer(-c["spin"], A.PRECISION)
    tr.rk45_tol = 1e-7
    rays = orc.pointsource_rays(tr, c["src"]) if c["point"] else orc.imageplane_rays(tr, c["src"])
    orc.redshift_start(rays, tr, c["src"].V if c["point"] else 0.0, not c["point"])
    orc.trace(rays, tr, c["stop"], ob.make_run(integ, r_max=c["r_max"], steplim=30000))
    orc.redshift_dest(rays, tr, c["stop"], not c["point"])
    return rays


@needs_ref
@pytest.mark.parametrize("integ", [A.RK4, A.RK45])
@pytest.mark.parametrize("which", ["moving_flatplane", "rotated_bundles"])
def test_live_reference_corner_cases(orc, which, integ):
    ref = ob.Ref()
    c = corner_setup(orc, which)
    h = ref.pointsource(c["src"], c["spin"]) if c["point"] else ref.imageplane(c["src"], c["spin"])
    h.set_rk45_tol(1e-7)
    if not c["point"]:
        rr = h.rays(copy=False)
        rr["rdot_flips"][:] = 0            # uninitialised in the reference
        rr["equatorial_crossings"][:] = 0
    h.redshift_start()
    h.run(c["stop"], ob.make_run(integ, r_max=c["r_max"], steplim=30000))
    h.redshift_dest(c["stop"], not c["point"])
    rays = oracle_corner(orc, which, integ)
    assert len(rays) == h.count() == (1333 if c["point"] else 980)
    assert_bit_identical(rays, h.rays())
    assert (rays["status"] & A.STATUS_DEST).sum() > 100 and (rays["status"] & A.STATUS_RLIM).sum() > 300


@needs_ref
@pytest.mark.parametrize("integ", [A.EULER, A.RK4, A.RK45])
def test_live_reference_ray_paths(orc, integ):
    """The serial outfile branch on a second pass over already-traced rays too: rays that hit the step limit (steps < 0) are
    skipped without even their blank lines, the others continue from where they stopped."""
    ref = ob.Ref()
    src = ob.make_pointsource([0, 6, 0.7, 0], 0.3, 0.3)
    tr = orc.tracer(SPIN, A.TOL)
    h = ref.pointsource(src, SPIN)
    rays = orc.pointsource_rays(tr, src)
    stop = ob.make_stop(A.STOP_THETALIM, theta_lim=0.0)
    for steplim, opts in ((150, ob.make_path_opts(4, True)), (400, ob.make_path_opts(5, False, 30.0, 3.0))):
        run = ob.make_run(integ, r_max=80.0, steplim=steplim)
        woff, wpts = h.run_paths(stop, run, opts)
        off, pts = orc.trace_paths(rays, tr, stop, run, opts)
        assert (off == woff).all() and pts.shape == wpts.shape and ob.same_bits(pts, wpts).all()
        assert_bit_identical(rays, h.rays(), skip=("redshift", "emit"))
    assert (rays["steps"] < -1).any()


@needs_ref
def test_live_reference_euler_image_plane(orc):
    ref = ob.Ref()
    ip = ob.make_imageplane(200.0, 75.0, -12, 12, 21)
    tr = orc.tracer(-SPIN, A.PRECISION)
    h = ref.imageplane(ip, SPIN)
    rays = orc.imageplane_rays(tr, ip)
    rr = h.rays(copy=False)
    rr["rdot_flips"][:] = 0
    rr["equatorial_crossings"][:] = 0
    stop = ob.make_stop()
    run = ob.make_run(A.EULER, r_max=300.0)
    h.run(stop, run)
    orc.trace(rays, tr, stop, run)
    assert_bit_identical(rays, h.rays(), skip=("redshift", "emit"))


def test_rk45_nan_ray_is_retired_not_hung(orc):
    """The reference spins forever on the NaN centre pixel under RK45; the oracle and the CUDA path
    retire the ray with STEPLIM set and a negated step count instead."""
    ip = ob.make_imageplane(1000.0, 30.0, -20, 20, 8)     # 9x9 grid containing x = y = 0 exactly
    tr = orc.tracer(-SPIN, A.PRECISION)
    rays = orc.imageplane_rays(tr, ip)
    orc.redshift_start(rays, tr, 0.0, True)
    centre = np.flatnonzero((rays["alpha"] == 0) & (rays["beta"] == 0))
    assert len(centre) == 1 and np.isnan(rays["h"][centre[0]])
    orc.trace(rays, tr, ob.make_stop(), ob.make_run(A.RK45, r_max=1100.0))
    c = rays[centre[0]]
    assert c["status"] & A.STATUS_STEPLIM and c["steps"] < 0
    others = np.delete(np.arange(len(rays)), centre)
    assert (rays["steps"][others] > 0).all()
