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
    "caustic_discplane": {
        "app_caustic_rk4": "dist = 500\nincl = 30\nplane_phi0 = 0\nspin = 0.998\nr_disc = 50\nx0 = -20\nxmax = 20\ny0 = -20\nymax = 20\n"
                           "Nx = 40\nNy = 40\nintegrator = rk4\nprecision = 100\nshow_progress = 0\nbundle_eps_frac = 0.01\n",
        "app_caustic_rk45": "dist = 500\nincl = 30\nplane_phi0 = 0\nspin = 0.998\nr_disc = 50\nx0 = -20\nxmax = 20\ny0 = -20\nymax = 20\n"
                            "Nx = 24\nNy = 24\nintegrator = rk45\nrk45_tol = 1e-8\nprecision = 100\nshow_progress = 0\nbundle_eps_frac = 0.01\n",
    },
}
