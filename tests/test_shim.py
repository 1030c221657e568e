"""The drop-in boundary from the application side: the reference's own, UNMODIFIED application and test sources are
compiled against the source-compatible shim headers (raytrace_cpu_b200/shim/raytracer/) and linked to libkerrtrace.so.

* CPU: every application of the three families in scope compiles (the FITS writers against a declarations-only
  fitsio.h, cfitsio being absent here); the ones that link without cfitsio are built into shim/_built/.
  These need the reference checkout and are skipped where /root/reference does not exist.
* GPU: the prebuilt binaries (they travel with the repo snapshot) are RUN on the B200: the reference's
  emissivity_rd application reproduces the 7-column profile its CPU build wrote (tests/golden/emissivity_rd_app.dat),
  and the reference's own acceptance test emissivity_rk45_test prints PASS.
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "raytrace_cpu_b200", "shim")
BUILT = os.path.join(SHIM, "_built")
REF = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")

needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "raytracer")), reason="reference checkout absent")


def make_overlay(tmp):
    src = os.path.join(tmp, "src")
    os.makedirs(src)
    os.symlink(os.path.join(REF, "src", "include"), os.path.join(src, "include"))
    for d in ("emissivity", "imageplane", "caustic", "tests", "ray_paths"):
        os.makedirs(os.path.join(src, d))
        for f in os.listdir(os.path.join(REF, "src", d)):
            if f.endswith(".cpp"):
                os.symlink(os.path.join(REF, "src", d, f), os.path.join(src, d, f))
    os.makedirs(os.path.join(src, "raytracer"))
    for f in os.listdir(os.path.join(SHIM, "raytracer")):
        os.symlink(os.path.join(SHIM, "raytracer", f), os.path.join(src, "raytracer", f))
    return src


@needs_reference
@pytest.mark.parametrize("app", [
    "emissivity/emissivity_rd.cpp", "imageplane/imageplane_disc_image.cpp", "imageplane/imageplane_disc_image_rd.cpp",
    "imageplane/imageplane_disc_image_isco.cpp", "caustic/caustic_discplane.cpp", "caustic/caustic_plane.cpp",
    "caustic/caustic_sourceplane.cpp", "tests/emissivity_rk45_test.cpp", "tests/integrator_perf_test.cpp",
    "tests/emissivity_rk45_plot.cpp", "tests/raytrace_rk4_test.cpp", "ray_paths/trace_rays.cpp", "ray_paths/trace_rays_imageplane.cpp",
])
def test_unmodified_reference_sources_compile_against_the_shim(app):
    with tempfile.TemporaryDirectory() as tmp:
        src = make_overlay(tmp)
        cmd = ["/usr/bin/g++", "-fsyntax-only", "-fopenmp", "-w", f"-I{src}", f"-I{ROOT}/include", f"-I{ROOT}/tests/fitsio_stub",
               os.path.join(src, app)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]


@needs_reference
def test_build_script_links_the_apps():
    p = subprocess.run([os.path.join(SHIM, "build_apps.sh")], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    for exe in ("emissivity", "emissivity_rd", "emissivity_rk45_test", "integrator_perf_test", "raytrace_rk4_test", "trace_rays",
                "trace_rays_imageplane"):
        assert os.access(os.path.join(BUILT, exe), os.X_OK)
    deps = subprocess.run(["ldd", os.path.join(BUILT, "emissivity_rd")], capture_output=True, text=True).stdout
    assert "libkerrtrace.so" in deps


def test_shim_headers_do_not_copy_the_reference_engine():
    """The shim holds interface declarations and calls into the C ABI -- no integrator."""
    for f in os.listdir(os.path.join(SHIM, "raytracer")):
        text = open(os.path.join(SHIM, "raytracer", f)).read()
        assert "momentum_from_consts<T>(pt2" not in text and "a21" not in text
    assert "kt_trace(" in open(os.path.join(SHIM, "raytracer", "raytracer.h")).read()


built = pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "emissivity_rd")), reason="shim/_built not present")


@pytest.mark.gpu
@built
def test_reference_emissivity_rd_app_runs_unchanged_on_the_gpu():
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "par"))
        os.makedirs(os.path.join(tmp, "bin"))
        out = os.path.join(tmp, "out.dat")
        with open(os.path.join(tmp, "par", "emissivity_rd.par"), "w") as f:
            f.write(f"outfile = {out}\ntheta_lim = 1.5707963\nsource = 0 10 1E-3 1.5707\nV = 0\nspin = 0.998\n"
                    "dcosalpha = 0.02\ndbeta = 0.02\nNr = 30\nlogbin_r = 1\nshow_progress = 0\n")
        env = dict(os.environ, KERRTRACE_QUIET="1")
        p = subprocess.run([os.path.join(BUILT, "emissivity_rd")], cwd=os.path.join(tmp, "bin"), capture_output=True, text=True,
                           env=env, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        got = np.genfromtxt(out)
    want = np.genfromtxt(os.path.join(GOLDEN, "emissivity_rd_app.dat"))      # written by the reference's CPU build
    assert got.shape == want.shape == (30, 7)
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=1e-8)              # bin radii
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=1e-8)              # proper areas (host code of the app)
    assert np.abs(got[:, 2] - want[:, 2]).sum() <= 0.002 * want[:, 2].sum()   # rays per annulus
    same = (got[:, 2] == want[:, 2]) & (want[:, 2] >= 100)
    assert same.sum() >= 5
    for col in (3, 4, 5, 6):                                                 # flux, emissivity, <g>, <t>
        np.testing.assert_allclose(got[same, col], want[same, col], rtol=2e-3)
        assert (np.abs(got[same, col] / want[same, col] - 1) <= 2e-8).mean() >= 0.8


@pytest.mark.gpu
@built
def test_reference_acceptance_test_passes_on_the_gpu():
    """src/tests/emissivity_rk45_test.cpp (RK4 vs RK45 binned profiles), built unchanged."""
    env = dict(os.environ, KERRTRACE_QUIET="1")
    p = subprocess.run([os.path.join(BUILT, "emissivity_rk45_test")], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "PASS" in p.stdout
    p = subprocess.run([os.path.join(BUILT, "integrator_perf_test")], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "raytrace_rk4_test")), reason="shim/_built not present")
def test_reference_euler_vs_rk4_test_on_the_gpu():
    """src/tests/raytrace_rk4_test.cpp: the same rays through Integrator::Euler and Integrator::RK4; its summary must be
    the one its CPU build printed (tests/golden/raytrace_rk4_test.stdout)."""
    import re
    env = dict(os.environ, KERRTRACE_QUIET="1")
    p = subprocess.run([os.path.join(BUILT, "raytrace_rk4_test")], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    want = open(os.path.join(GOLDEN, "raytrace_rk4_test.stdout")).read()
    assert "PASS" in p.stdout and "PASS" in want

    def summary(text):
        return {k: int(v) for k, v in re.findall(r"^\s*(Total rays|Skipped \(invalid\)|Hit disc \(theta_max\)|Hit r_max|Hit horizon|Diverged trajectories):\s+(\d+)", text, re.M)}

    got, ref = summary(p.stdout), summary(want)
    assert len(ref) == 6 and got.keys() == ref.keys()
    for k in ("Total rays", "Skipped (invalid)"):
        assert got[k] == ref[k]
    for k in ("Hit disc (theta_max)", "Hit r_max", "Hit horizon", "Diverged trajectories"):
        assert abs(got[k] - ref[k]) <= 2, (k, got, ref)      # a separatrix ray may change sides (one ray = 0.3 %)


def split_blocks(text):
    blocks, cur, blanks = [], [], 0
    for line in text.split("\n")[:-1]:
        if line.strip():
            cur.append([float(x) for x in line.split()])
        else:
            blanks += 1
            if blanks == 2:
                blocks.append(np.array(cur, dtype=np.float64).reshape(-1, 4))
                cur, blanks = [], 0
    assert not cur
    return blocks


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "trace_rays")), reason="shim/_built not present")
@pytest.mark.parametrize("app", ["trace_rays", "trace_rays_imageplane"])
def test_reference_ray_path_apps_on_the_gpu(app):
    """src/ray_paths/trace_rays{,_imageplane}.cpp (Integrator::Euler + outfile), unchanged: the text file must be the one the
    CPU build wrote -- same number of ray blocks, same number of points per ray, every column within the 8 printed digits."""
    from golden_cases import APP_PARS
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "par"))
        os.makedirs(os.path.join(tmp, "bin"))
        out = os.path.join(tmp, "out.dat")
        with open(os.path.join(tmp, "par", app + ".par"), "w") as f:
            f.write(f"outfile = {out}\n" + APP_PARS[app]["app_" + app])
        env = dict(os.environ, KERRTRACE_QUIET="1")
        p = subprocess.run([os.path.join(BUILT, app)], cwd=os.path.join(tmp, "bin"), capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        got = split_blocks(open(out).read())
    want = split_blocks(open(os.path.join(GOLDEN, "app_" + app + ".dat")).read())
    assert len(got) == len(want) > 20
    same = [len(a) == len(b) for a, b in zip(got, want)]
    assert np.mean(same) >= 0.95
    close = []
    for a, b, s in zip(got, want, same):
        if s and len(b):
            scale = np.maximum(1.0, np.sqrt((b[:, 1:] ** 2).sum(axis=1)))[:, None]
            close.append(bool((np.abs(a[:, 1:] - b[:, 1:]) <= 2e-8 * scale).all() and np.allclose(a[:, 0], b[:, 0], rtol=2e-8, atol=2e-8)))
    assert np.mean(close) >= 0.95, np.mean(close)
    # byte-identical lines for the bulk of the file (same formatting code: the reference's TextOutput)


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs 2 and 4 at application level: the reference's FITS-writing programs, built unchanged against the
# shim (+ the bundled minimal FITS writer, cfitsio being absent), run on the B200 and compared with the FITS files
# the reference's own CPU build wrote for the same par files (tests/golden/app_*.fits, scripts/make_golden.py).
# ---------------------------------------------------------------------------------------------------------
def run_app(exe, partext, tmp):
    out = os.path.join(tmp, "out.fits")
    par = os.path.join(tmp, "run.par")
    with open(par, "w") as f:
        f.write(f"outfile = {out}\n" + partext)
    env = dict(os.environ, KERRTRACE_QUIET="1")
    p = subprocess.run([os.path.join(BUILT, exe), f"--parfile={par}"], cwd=tmp, capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    from fits_reader import by_extname, read_fits
    hdus = read_fits(out)
    return hdus[0][0], by_extname(hdus)


def close_fraction(a, b, rtol, mask):
    a, b = a[mask], b[mask]
    ok = np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-12)
    return ok.mean() if ok.size else 1.0


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "imageplane_disc_image")), reason="shim/_built not present")
def test_reference_imageplane_disc_image_app_on_the_gpu():
    from fits_reader import by_extname, read_fits
    from golden_cases import APP_PARS
    want_h = read_fits(os.path.join(GOLDEN, "app_image_rk4.fits"))
    want = by_extname(want_h)
    with tempfile.TemporaryDirectory() as tmp:
        prim, got = run_app("imageplane_disc_image", APP_PARS["imageplane_disc_image"]["app_image_rk4"], tmp)
    assert list(got) == ["FLUX", "RADIUS", "PHI", "ENSHIFT", "TIME", "EMIS"] == list(want)      # the python notebooks' contract
    for k in ("DIST", "INCL", "SPIN", "ISCO", "RDISC", "NRAYS"):
        assert prim[k] == want_h[0][0][k], k
    assert abs(prim["DISCRAYS"] - want_h[0][0]["DISCRAYS"]) <= 0.01 * want_h[0][0]["DISCRAYS"]
    keep = np.ones((48, 48), dtype=bool)
    keep[23, :] = False                         # pixels fed by the degenerate y = 0 rays (flip_image: 48 - 24 - 1)
    hit_g, hit_w = np.isfinite(got["RADIUS"]), np.isfinite(want["RADIUS"])        # 0/0 = NaN where no ray hit (array.h:176-186)
    assert (hit_g != hit_w)[keep].sum() <= 0.005 * keep.sum()
    both = hit_g & hit_w & keep
    assert both.sum() > 1000
    for name in want:
        tol = 1e-9
        frac = close_fraction(got[name], want[name], tol, both)
        assert frac >= 0.99, (name, frac)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "caustic_discplane")), reason="shim/_built not present")
@pytest.mark.parametrize("case,n", [("app_caustic_rk4", 41), ("app_caustic_rk45", 25)])
def test_reference_caustic_discplane_app_on_the_gpu(case, n):
    from fits_reader import by_extname, read_fits
    from golden_cases import APP_PARS
    want_h = read_fits(os.path.join(GOLDEN, case + ".fits"))
    want = by_extname(want_h)
    with tempfile.TemporaryDirectory() as tmp:
        prim, got = run_app("caustic_discplane", APP_PARS["caustic_discplane"][case], tmp)
    names = ["DET_J", "SIGN_J", "ORDER", "HIT", "RADIUS", "PHI", "X_DISC", "Y_DISC", "REDSHIFT"]       # python/caustic_discplane.py:38-54
    assert list(got) == names == list(want)
    assert prim["NRAYS"] == want_h[0][0]["NRAYS"] == n * n
    keep = np.ones((n, n), dtype=bool)
    keep[n // 2, :] = False                      # the degenerate y = 0 bundle centres
    rk45 = "rk45" in case
    assert (got["HIT"] != want["HIT"])[keep].mean() <= (0.02 if rk45 else 0.01)
    both = (got["HIT"] > 0) & (want["HIT"] > 0) & keep
    tol = 1e-6 if rk45 else 1e-9
    for name in ("RADIUS", "PHI", "X_DISC", "Y_DISC", "REDSHIFT"):
        frac = close_fraction(got[name], want[name], tol, both)
        assert frac >= (0.95 if rk45 else 0.98), (name, frac)
    assert (got["ORDER"] == want["ORDER"])[both].mean() >= 0.98
    # Jacobian determinant from the east/west/north/south satellites: finite differences over eps = 1 % of a pixel
    # amplify the per-ray differences by ~1/eps
    assert close_fraction(got["DET_J"], want["DET_J"], 1e-3 if rk45 else 1e-6, both) >= (0.90 if rk45 else 0.96)
    assert (got["SIGN_J"] == want["SIGN_J"])[both].mean() >= 0.97


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "caustic_plane")), reason="shim/_built not present")
@pytest.mark.parametrize("app,case,n", [("caustic_plane", "app_caustic_plane_rk4", 25), ("caustic_sourceplane", "app_caustic_sourceplane_rk4", 26)])
def test_reference_sibling_caustic_apps_on_the_gpu(app, case, n):
    """caustic_plane.cpp (FlatPlaneDestination behind the hole, ray bundles) and caustic_sourceplane.cpp (theta_max = 0,
    custom steplim, caustic_sourceplane.cpp:155), unchanged: SURVEY 8f.2.  Same extensions, same pixels as the FITS files of the
    reference's CPU build."""
    from fits_reader import by_extname, read_fits
    from golden_cases import APP_PARS
    want = by_extname(read_fits(os.path.join(GOLDEN, case + ".fits")))
    with tempfile.TemporaryDirectory() as tmp:
        prim, got = run_app(app, APP_PARS[app][case], tmp)
    assert list(got) == list(want) and len(want) == 8
    hit_name = "HIT_PLANE" if app == "caustic_plane" else "ESCAPED"
    coords = ("X_S", "Y_S") if app == "caustic_plane" else ("THETA_S", "PHI_S")
    keep = np.ones((n, n), dtype=bool)
    keep[n // 2 if n % 2 else slice(0, 0), :] = False          # the y = 0 row exists only on odd grids
    assert want[hit_name].sum() > 0.5 * n * n
    assert (got[hit_name] != want[hit_name])[keep].mean() <= 0.01
    both = (got[hit_name] > 0) & (want[hit_name] > 0) & keep
    for name in coords:
        assert close_fraction(got[name], want[name], 1e-9, both) >= 0.98, name
    assert (got["ORDER"] == want["ORDER"])[both].mean() >= 0.98
    assert (got["RDOT_FLIPS"] == want["RDOT_FLIPS"])[both].mean() >= 0.98
    # det(J): bundles differentiate over 1 % of a pixel (amplification ~ 1/eps), the source-plane app over grid neighbours
    tol = 1e-6 if app == "caustic_plane" else 1e-8
    defined = both & np.isfinite(want["DET_J"]) & np.isfinite(got["DET_J"]) & (want["DET_J"] < 1e29) & (got["DET_J"] < 1e29)
    assert defined.sum() > 0.3 * n * n
    assert close_fraction(got["DET_J"], want["DET_J"], tol, defined) >= 0.95
    assert (got["SIGN_J"] == want["SIGN_J"])[both].mean() >= 0.97


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BUILT, "emissivity")), reason="shim/_built not present")
def test_reference_emissivity_rk45_flagship_app_on_the_gpu():
    """src/emissivity/emissivity.cpp (RK45 + theta_max + the z < 1e-2 disc filter; BASELINE config 3 geometry at 31 417 rays)
    with upstream's one-token fix of its line 79 (shim/build_apps.sh), against the .dat its CPU build wrote.  The adaptive
    controller may accept a different step sequence, so the comparison is the reference's own RK45 acceptance style: ray
    counts per annulus within 2 %, binned quantities within 1 % on annuli with >= 100 rays."""
    from golden_cases import APP_PARS
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "par"))
        os.makedirs(os.path.join(tmp, "bin"))
        out = os.path.join(tmp, "out.dat")
        with open(os.path.join(tmp, "par", "emissivity.par"), "w") as f:
            f.write(f"outfile = {out}\n" + APP_PARS["emissivity"]["app_emissivity_rk45"])
        env = dict(os.environ, KERRTRACE_QUIET="1")
        p = subprocess.run([os.path.join(BUILT, "emissivity")], cwd=os.path.join(tmp, "bin"), capture_output=True, text=True,
                           env=env, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        got = np.genfromtxt(out)
    want = np.genfromtxt(os.path.join(GOLDEN, "app_emissivity_rk45.dat"))
    assert got.shape == want.shape == (30, 7)
    np.testing.assert_allclose(got[:, :2], want[:, :2], rtol=1e-8)            # radii and proper areas: host code of the app
    assert np.abs(got[:, 2] - want[:, 2]).sum() <= 0.02 * want[:, 2].sum()
    full = want[:, 2] >= 100
    assert full.sum() >= 10
    for col in (3, 4, 5, 6):
        np.testing.assert_allclose(got[full, col], want[full, col], rtol=1e-2)


def test_only_the_exact_destination_classes_are_offloaded():
    """The reference's concrete destinations are not final: a user class derived from one of them may override reached() /
    step_limit() / four_velocity() and inherits the base class's device descriptor.  stop_of() compares the dynamic type with
    the three known classes exactly and throws for anything else (before any device call: runs without a GPU)."""
    prog = r'''
#include "raytracer/raytracer.h"
#include "raytracer/ray_destination.h"
#include <cstdio>
template <typename T> struct Wavy : public FlatDiscDestination<T> {
    using FlatDiscDestination<T>::FlatDiscDestination;
    bool reached(T r, T theta, T phi) const override { return theta >= 1.0 + 0.1 * sin(r); }
};
template <typename T> struct Mine : public RayDestination<T> {
    bool reached(T, T, T) const override { return false; }
};
int main() {
    FlatDiscDestination<double> disc(M_PI_2);
    DiscWithISCODestination<double> isco(1.24, 50.0, M_PI_2);
    FlatPlaneDestination<double> plane(0.5, 0.0, 10.0);
    if (kerrtrace_shim::stop_of<double>(&disc).kind != KT_STOP_FLATDISC) return 1;
    if (kerrtrace_shim::stop_of<double>(&isco).kind != KT_STOP_ISCODISC) return 2;
    if (kerrtrace_shim::stop_of<double>(&plane).kind != KT_STOP_FLATPLANE) return 3;
    int refused = 0;
    Wavy<double> wavy(M_PI_2);
    Mine<double> mine;
    try { kerrtrace_shim::stop_of<double>(&wavy); } catch (const std::runtime_error&) { refused++; }
    try { kerrtrace_shim::stop_of<double>(&mine); } catch (const std::runtime_error&) { refused++; }
    try { kerrtrace_shim::stop_of<double>(nullptr); } catch (const std::runtime_error&) { refused++; }
    std::printf("refused %d\n", refused);
    return refused == 3 ? 0 : 4;
}
'''
    ref_inc = os.path.join(REF, "src")
    if not os.path.isdir(os.path.join(ref_inc, "include")):
        pytest.skip("the shim headers include the reference's kerr.h / text_output.h: reference checkout absent")
    with tempfile.TemporaryDirectory() as tmp:
        src = make_overlay(tmp)
        with open(os.path.join(src, "t.cpp"), "w") as f:
            f.write(prog)
        exe = os.path.join(tmp, "t")
        libdir = os.path.join(ROOT, "raytrace_cpu_b200")
        p = subprocess.run(["/usr/bin/g++", "-O1", "-w", f"-I{src}", f"-I{ROOT}/include", os.path.join(src, "t.cpp"), "-o", exe,
                            f"-L{libdir}", f"-Wl,-rpath,{libdir}", "-lkerrtrace"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-3000:]
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "refused 3" in r.stdout, (r.returncode, r.stdout, r.stderr)
