"""Definitions of the golden fixtures in tests/golden/ (generated from the unmodified reference by
scripts/make_golden.py).  Shared by the generator, the oracle tests and the GPU parity tests."""
import numpy as np

from raytrace_cpu_b200 import _abi as A
import oracle_bindings as ob

SPIN = 0.998
R_ISCO = 1.2369706630706787   # kerr_isco<double>(0.998, +1) through its float temporaries (kerr.h:29-31)

CASES = {
    # the reference's raytrace_rk4_test grid (src/tests/raytrace_rk4_test.cpp:26-32): 354 slots, Keplerian source
    "ps354_rk4_thetalim": dict(source="point", spin=SPIN, pos=[0, 5, 1e-3, 0], V=1 / (SPIN + 5 ** 1.5), dcos=0.2, dbeta=0.2,
                               defaults=True, integrator=A.RK4, stop=dict(kind=A.STOP_THETALIM), r_max=1000.0, redshift="V"),
    # the Euler leg of the same test (src/tests/raytrace_rk4_test.cpp:40-42) and an image plane under Euler, the integrator
    # the apps get when they leave run_raytrace's default (raytracer.h:112)
    "ps354_euler_thetalim": dict(source="point", spin=SPIN, pos=[0, 5, 1e-3, 0], V=1 / (SPIN + 5 ** 1.5), dcos=0.2, dbeta=0.2,
                                 defaults=True, integrator=A.EULER, stop=dict(kind=A.STOP_THETALIM), r_max=1000.0, redshift="V"),
    "plane16_euler_thetalim": dict(source="plane", spin=SPIN, dist=1e3, incl=60.0, x0=-15.0, xmax=15.0, n=15, integrator=A.EULER,
                                   stop=dict(kind=A.STOP_THETALIM), r_max=1.1e3, redshift="V"),
    # coarse versions of the emissivity_rk45_test grid (src/tests/emissivity_rk45_test.cpp:37-50)
    "ps_r5_rk4_thetalim": dict(source="point", spin=SPIN, pos=[0, 5, 1e-3, 0], V=0.0, dcos=0.1, dbeta=0.1, integrator=A.RK4,
                               stop=dict(kind=A.STOP_THETALIM), r_max=1000.0, redshift="V"),
    "ps_r5_rk45_thetalim": dict(source="point", spin=SPIN, pos=[0, 5, 1e-3, 0], V=0.0, dcos=0.1, dbeta=0.1, integrator=A.RK45,
                                stop=dict(kind=A.STOP_THETALIM), r_max=1000.0, redshift="V"),
    # BASELINE config 1 geometry (h = 10), FlatDiscDestination as emissivity_rd.cpp:101-106
    "ps_h10_rk4_flatdisc": dict(source="point", spin=SPIN, pos=[0, 10, 1e-3, 1.5707], V=0.0, dcos=0.1, dbeta=0.1, integrator=A.RK4,
                                stop=dict(kind=A.STOP_FLATDISC), r_max=1000.0, redshift="dest"),
    "ps_h10_rk45_iscodisc": dict(source="point", spin=SPIN, pos=[0, 10, 1e-3, 1.5707], V=0.0, dcos=0.1, dbeta=0.1, integrator=A.RK45,
                                 stop=dict(kind=A.STOP_ISCODISC, r_isco=R_ISCO, r_out=400.0), r_max=1000.0, redshift="dest",
                                 rk45_tol=1e-9),
    # BASELINE config 2 geometry at low resolution (even grid: contains the NaN centre pixel, RK4 survives it)
    "plane32_rk4_thetalim": dict(source="plane", spin=SPIN, dist=1e4, incl=30.0, x0=-30.0, xmax=30.0, n=32, integrator=A.RK4,
                                 stop=dict(kind=A.STOP_THETALIM), r_max=1.1e4, redshift="V"),
    # BASELINE config 4 geometry at low resolution: bundles, ISCO disc, RK45 (caustic_discplane.par_example)
    "bundles21_rk45_iscodisc": dict(source="bundles", spin=SPIN, dist=500.0, incl=30.0, x0=-20.0, xmax=20.0, n=20, integrator=A.RK45,
                                    stop=dict(kind=A.STOP_ISCODISC, r_isco=R_ISCO, r_out=50.0), r_max=550.0, redshift="dest",
                                    range_phi=False),
    # flat source plane behind the hole (caustic_plane.cpp:151), RK4
    "plane24_rk4_flatplane": dict(source="plane", spin=SPIN, dist=500.0, incl=60.0, x0=-12.0, xmax=12.0, n=24, integrator=A.RK4,
                                  stop=dict(kind=A.STOP_FLATPLANE, incl=np.radians(60.0), phi0=0.0, z_s=50.0), r_max=550.0,
                                  redshift="dest", range_phi=False),
}


# Ray-path dumping (run_raytrace with an outfile, raytracer.cpp:86-100): the trace_rays / trace_rays_imageplane call shapes
# (src/ray_paths/trace_rays.cpp:71, trace_rays_imageplane.cpp:61) plus RK4 / RK45 with a radial window and a destination.
# Fixtures hold the points the reference wrote (17 significant digits: the exact doubles), their per-ray offsets and the rays.
PATH_CASES = {
    "paths_ps_euler_cart": dict(source="point", spin=SPIN, pos=[0, 5, 1e-3, 0], V=1 / (SPIN + 5 ** 1.5), dcos=0.4, dbeta=0.4,
                                defaults=True, integrator=A.EULER, stop=dict(kind=A.STOP_THETALIM), r_max=100.0,
                                paths=dict(write_step=10, write_cartesian=True, write_rmax=-1.0, write_rmin=-1.0)),
    "paths_plane_euler_cart": dict(source="plane", spin=SPIN, dist=100.0, incl=60.0, x0=-8.0, xmax=8.0, n=7, integrator=A.EULER,
                                   stop=dict(kind=A.STOP_THETALIM, theta_lim=0.0), r_max=150.0,
                                   paths=dict(write_step=10, write_cartesian=True, write_rmax=-1.0, write_rmin=-1.0)),
    "paths_ps_rk4_window_bl": dict(source="point", spin=SPIN, pos=[0, 5, 1e-3, 0], V=0.0, dcos=0.4, dbeta=0.4, defaults=True,
                                   integrator=A.RK4, stop=dict(kind=A.STOP_THETALIM), r_max=100.0,
                                   paths=dict(write_step=3, write_cartesian=False, write_rmax=20.0, write_rmin=2.0)),
    "paths_ps_rk45_iscodisc": dict(source="point", spin=SPIN, pos=[0, 10, 1e-3, 1.5707], V=0.0, dcos=0.4, dbeta=0.4, defaults=True,
                                   integrator=A.RK45, stop=dict(kind=A.STOP_ISCODISC, r_isco=R_ISCO, r_out=400.0), r_max=1000.0,
                                   steplim=2000, paths=dict(write_step=9, write_cartesian=True, write_rmax=-1.0, write_rmin=-1.0)),
}


def caustic_par(text):
    """caustic_discplane par text -> dict of the parameters the app reads (caustic_discplane.cpp:56-100)."""
    kv = dict(line.split(" = ") for line in text.strip().split("\n"))
    return dict(dist=float(kv["dist"]), incl=float(kv["incl"]), spin=float(kv["spin"]), r_disc=float(kv["r_disc"]), Nx=int(kv["Nx"]),
                Ny=int(kv["Ny"]), x0=float(kv["x0"]), xmax=float(kv["xmax"]), y0=float(kv["y0"]), ymax=float(kv["ymax"]),
                plane_phi0=float(kv["plane_phi0"]), rk45=kv["integrator"] == "rk45", rk45_tol=float(kv.get("rk45_tol", 1e-8)),
                precision=float(kv["precision"]), bundle_eps_frac=float(kv["bundle_eps_frac"]))


def build_source(case):
    if case["source"] == "point":
        if case.get("defaults"):
            return ob.make_pointsource(case["pos"], case["dcos"], case["dbeta"], V=case["V"])
        return ob.make_pointsource(case["pos"], case["dcos"], case["dbeta"], V=case["V"], cosalpha0=-0.995, cosalphamax=0.995,
                                   beta0=-np.pi, betamax=np.pi)
    return ob.make_imageplane(case["dist"], case["incl"], case["x0"], case["xmax"], case["n"], bundles=case["source"] == "bundles")


def tracer_for(case, make_tracer):
    """make_tracer(spin, precision) -> kt_tracer; image planes store the negated spin (imageplane.cpp:12)."""
    if case["source"] == "point":
        tr = make_tracer(case["spin"], A.TOL)
    else:
        tr = make_tracer(-case["spin"], A.PRECISION)
    if "rk45_tol" in case:
        tr.rk45_tol = case["rk45_tol"]
    return tr


# par files (minus the outfile line) of the application-level golden runs: the reference's own binaries, CPU build.
# BASELINE config 2 / config 4 geometries at low resolution.  imageplane_disc_image must use rk4: its 49 x 49 grid holds
# the exact-centre pixel on which the reference's RK45 never terminates.
APP_PARS = {
    "imageplane_disc_image": {
        "app_image_rk4": "dist = 10000\nincl = 30\nspin = 0.998\nr_disc = 30\nx0 = -30\nxmax = 30\ny0 = -30\nymax = 30\n"
                         "Nx = 48\nimg_Nx = 48\nq1 = 3\nrb1 = 4\nq2 = 3\nrb2 = 10\nq3 = 3\nintegrator = rk4\nshow_progress = 0\n",
    },
    # emissivity.cpp, the RK45 flagship (BASELINE config 3 geometry, r = 5): built from upstream's source with its own
    # one-token fix of line 79 -- CPU build for the golden file, shim build for the GPU run
    "emissivity": {
        "app_emissivity_rk45": "source = 0 5 1E-3 0\nV = 0\nspin = 0.998\ndcosalpha = 0.02\ndbeta = 0.02\nNr = 30\nlogbin_r = 1\n"
                               "show_progress = 0\ncosalpha0 = -0.995\ncosalphamax = 0.995\n",
    },
    # the sibling caustic applications: a flat source plane behind the hole (FlatPlaneDestination, bundles) and
    # the celestial sphere (theta_max = 0 disables the theta stop, custom steplim; caustic_sourceplane.cpp:155)
    "caustic_plane": {
        "app_caustic_plane_rk4": "dist = 500\nincl = 60\nplane_phi0 = 0\nspin = 0.998\nz_s = 50\nr_max = 600\nx0 = -12\nxmax = 12\nNx = 24\n"
                                 "integrator = rk4\nprecision = 100\nshow_progress = 0\nbundle_eps_frac = 0.01\n",
    },
    "caustic_sourceplane": {
        "app_caustic_sourceplane_rk4": "dist = 300\nincl = 60\nplane_phi0 = 0\nspin = 0.998\nx0 = -12\nxmax = 12\nNx = 25\n"
                                       "integrator = rk4\nprecision = 100\nshow_progress = 0\nsteplim = 20000\n",
    },
    # src/ray_paths/trace_rays.cpp, trace_rays_imageplane.cpp: Euler + outfile (text: t x y z per written point)
    "trace_rays": {
        "app_trace_rays": "source = 0 5 1E-3 0\nV = 0\nspin = 0.998\ndcosalpha = 0.4\ndbeta = 0.9\nr_max = 60\nshow_progress = 0\n"
                          "write_step = 10\n",
    },
    "trace_rays_imageplane": {
        "app_trace_rays_imageplane": "dist = 100\nincl = 60\nspin = 0.998\nx0 = -8\nxmax = 8\nNx = 6\ny0 = -8\nymax = 8\nNy = 6\n"
                                     "write_step = 10\nplane_phi0 = 100\ntol = 0\n",
    },
    "caustic_discplane": {
        "app_caustic_rk4": "dist = 500\nincl = 30\nplane_phi0 = 0\nspin = 0.998\nr_disc = 50\nx0 = -20\nxmax = 20\ny0 = -20\nymax = 20\n"
                           "Nx = 40\nNy = 40\nintegrator = rk4\nprecision = 100\nshow_progress = 0\nbundle_eps_frac = 0.01\n",
        "app_caustic_rk45": "dist = 500\nincl = 30\nplane_phi0 = 0\nspin = 0.998\nr_disc = 50\nx0 = -20\nxmax = 20\ny0 = -20\nymax = 20\n"
                            "Nx = 24\nNy = 24\nintegrator = rk45\nrk45_tol = 1e-8\nprecision = 100\nshow_progress = 0\nbundle_eps_frac = 0.01\n",
        # bundle_eps_frac = 0: single rays, Jacobian from the grid neighbours (caustic_discplane.cpp:337-448); odd ray count
        # per axis would hold the exact-centre pixel (NaN ray), 36 steps -> 37 rays does too, so RK4
        "app_caustic_grid_rk4": "dist = 500\nincl = 30\nplane_phi0 = 0\nspin = 0.998\nr_disc = 50\nx0 = -20\nxmax = 20\ny0 = -20\nymax = 20\n"
                                "Nx = 36\nNy = 36\nintegrator = rk4\nprecision = 100\nshow_progress = 0\nbundle_eps_frac = 0\n",
    },
}
