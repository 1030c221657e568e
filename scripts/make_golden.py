"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_oracle.so, built by
oracle/Makefile from /root/reference) and tests/golden/emissivity_rd_app.dat from the reference's own
emissivity_rd application compiled out of tree.  Run in the build container (needs /root/reference):

    python scripts/make_golden.py

The reference ships no golden vectors, so these fixtures -- outputs of the reference
itself on deterministic grids -- are what pins the oracle and the CUDA path on machines without /root/reference.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from raytrace_cpu_b200 import _abi as A  # noqa: E402
import oracle_bindings as ob  # noqa: E402
from golden_cases import CASES, PATH_CASES, build_source  # noqa: E402

from golden_cases import APP_PARS  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def run_case(ref, case):
    src = build_source(case)
    if case["source"] == "point":
        h = ref.pointsource(src, case["spin"], A.TOL)
    else:
        h = ref.imageplane(src, case["spin"], A.PRECISION)
        # the reference never initialises these two counters for image planes
        rr = h.rays(copy=False)
        rr["rdot_flips"][:] = 0
        rr["equatorial_crossings"][:] = 0
    if "rk45_tol" in case:
        h.set_rk45_tol(case["rk45_tol"])
    h.redshift_start()
    init = h.rays()
    stop = ob.make_stop(**case["stop"])
    run = ob.make_run(case["integrator"], r_max=case["r_max"], steplim=case.get("steplim", -1))
    if "paths" in case:
        offsets, points = h.run_paths(stop, run, ob.make_path_opts(**case["paths"]))
        return init, h.rays(), offsets, points
    h.run(stop, run)
    reverse = case["source"] != "point"
    if case["redshift"] == "V":
        h.redshift(-1.0, reverse)
    else:
        h.redshift_dest(stop, reverse)
    if case.get("range_phi", True):
        h.range_phi()
    return init, h.rays()


FORCE = "--force" in sys.argv   # default: only fixtures that do not exist yet are generated


def missing(filename):
    return FORCE or not os.path.exists(os.path.join(GOLDEN, filename))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ob.Ref()
    for name, case in CASES.items():
        if not missing(name + ".npz"):
            continue
        init, final = run_case(ref, case)
        used = final["steps"] != -1
        # zero the fields the reference leaves uninitialised in unused slots so the fixture is deterministic
        for arr in (init, final):
            for f in arr.dtype.names:
                if f not in ("steps", "status"):
                    arr[f][~used] = 0
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), init=init, final=final)
        print(f"{name}: {len(final)} slots, {int(used.sum())} rays, status histogram "
              f"{ {int(s): int((final['status'][used] == s).sum()) for s in np.unique(final['status'][used])} }")

    for name, case in PATH_CASES.items():
        if not missing(name + ".npz"):
            continue
        init, final, offsets, points = run_case(ref, case)
        used = final["steps"] != -1
        for arr in (init, final):
            for f in arr.dtype.names:
                if f not in ("steps", "status"):
                    arr[f][~used] = 0
        for arr in (init, final):          # the path fixtures pin positions and bookkeeping, not the energies
            arr["redshift"][:] = 0
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), init=init, final=final, offsets=offsets, points=points)
        print(f"{name}: {int(used.sum())} rays, {len(points)} points")

    # scalars the reference prints
    scal = {"spins": np.array([0.0, 0.1, 0.5, 0.9, 0.998, -0.5, -0.998])}
    scal["isco_pro"] = np.array([ref.kerr_isco(a, +1) for a in scal["spins"]])
    scal["isco_retro"] = np.array([ref.kerr_isco(a, -1) for a in scal["spins"]])
    scal["horizon"] = np.array([ref.kerr_horizon(a) for a in scal["spins"]])
    if missing("scalars.npz"):
        np.savez(os.path.join(GOLDEN, "scalars.npz"), **scal)

    # the reference's own emissivity_rd application, end to end (text output, 7 columns)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "emissivity_rd")
        subprocess.run(["/usr/bin/g++", "-O2", "-fopenmp", "-ffp-contract=off", "-w", f"-I{REF}/src", f"-I{REF}/src/raytracer",
                        f"{REF}/src/emissivity/emissivity_rd.cpp", f"{REF}/src/raytracer/raytracer.cpp",
                        f"{REF}/src/raytracer/pointsource.cpp", "-o", exe], check=True)
        # --parfile keeps a dangling c_str() in the app: use its default relative path instead
        os.makedirs(os.path.join(tmp, "par"))
        os.makedirs(os.path.join(tmp, "bin"))
        par = os.path.join(tmp, "par", "emissivity_rd.par")
        out = os.path.join(GOLDEN, "emissivity_rd_app.dat")
        with open(par, "w") as f:
            f.write(f"outfile = {out}\ntheta_lim = 1.5707963\nsource = 0 10 1E-3 1.5707\nV = 0\nspin = 0.998\n"
                    "dcosalpha = 0.02\ndbeta = 0.02\nNr = 30\nlogbin_r = 1\nshow_progress = 0\n")
        if missing("emissivity_rd_app.dat"):
            subprocess.run([exe], check=True, stdout=subprocess.DEVNULL, cwd=os.path.join(tmp, "bin"))
            print("emissivity_rd_app.dat:", sum(1 for _ in open(out)), "lines")

        # emissivity.cpp with upstream's one-token fix of line 79 (build-time sed, as shim/build_apps.sh does)
        if missing("app_emissivity_rk45.dat"):
            fixed = os.path.join(tmp, "emissivity_fixed.cpp")
            text = open(f"{REF}/src/emissivity/emissivity.cpp").read()
            assert text.count("disc_r[ir]*dr : disc_r + dr, spin)") == 1
            open(fixed, "w").write(text.replace("disc_r[ir]*dr : disc_r + dr, spin)", "disc_r[ir]*dr : dr, spin)"))
            exe = os.path.join(tmp, "emissivity")
            subprocess.run(["/usr/bin/g++", "-O2", "-fopenmp", "-ffp-contract=off", "-w", f"-I{REF}/src", f"-I{REF}/src/raytracer",
                            f"-I{REF}/src/emissivity", fixed, f"{REF}/src/raytracer/raytracer.cpp", f"{REF}/src/raytracer/pointsource.cpp",
                            "-o", exe], check=True)
            out = os.path.join(GOLDEN, "app_emissivity_rk45.dat")
            with open(os.path.join(tmp, "par", "emissivity.par"), "w") as f:
                f.write(f"outfile = {out}\n" + APP_PARS["emissivity"]["app_emissivity_rk45"])
            subprocess.run([exe], check=True, stdout=subprocess.DEVNULL, cwd=os.path.join(tmp, "bin"))
            print("app_emissivity_rk45.dat:", sum(1 for _ in open(out)), "lines")

        # the reference's own Euler-vs-RK4 consistency test (stdout) and its ray-path dumping applications (text files)
        if missing("raytrace_rk4_test.stdout"):
            exe = os.path.join(tmp, "raytrace_rk4_test")
            subprocess.run(["/usr/bin/g++", "-O2", "-fopenmp", "-ffp-contract=off", "-w", f"-I{REF}/src", f"-I{REF}/src/raytracer",
                            f"{REF}/src/tests/raytrace_rk4_test.cpp", f"{REF}/src/raytracer/raytracer.cpp",
                            f"{REF}/src/raytracer/pointsource.cpp", "-o", exe], check=True)
            res = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
            with open(os.path.join(GOLDEN, "raytrace_rk4_test.stdout"), "w") as f:
                f.write(f"# exit code {res.returncode}\n" + res.stdout)
            print("raytrace_rk4_test.stdout: exit", res.returncode)
        for app, imp in (("trace_rays", "pointsource.cpp"), ("trace_rays_imageplane", "imageplane.cpp")):
            exe = os.path.join(tmp, app)
            built = False
            for parname, text in APP_PARS[app].items():
                if not missing(parname + ".dat"):
                    continue
                if not built:
                    subprocess.run(["/usr/bin/g++", "-O2", "-fopenmp", "-ffp-contract=off", "-w", f"-I{REF}/src", f"-I{REF}/src/raytracer",
                                    f"{REF}/src/ray_paths/{app}.cpp", f"{REF}/src/raytracer/raytracer.cpp",
                                    f"{REF}/src/raytracer/{imp}", "-o", exe], check=True)
                    built = True
                out = os.path.join(GOLDEN, parname + ".dat")
                # trace_rays' --parfile keeps a dangling c_str() (trace_rays.cpp:29-33, the same UB as SURVEY appendix A.16):
                # both apps are run on their default relative path ../par/<app>.par instead
                par = os.path.join(tmp, "par", app + ".par")
                with open(par, "w") as f:
                    f.write(f"outfile = {out}\n" + text)
                subprocess.run([exe], check=True, stdout=subprocess.DEVNULL, cwd=os.path.join(tmp, "bin"))
                print(parname + ".dat:", sum(1 for _ in open(out)), "lines")

        # the reference's FITS-writing applications (CPU build of the reference, linked with the bundled minimal
        # FITS writer because cfitsio is not installed here): golden multi-extension FITS files
        minifits = os.path.join(ROOT, "raytrace_cpu_b200", "shim", "minifits")
        obj = os.path.join(tmp, "minifits.o")
        subprocess.run(["/usr/bin/gcc", "-O2", "-c", os.path.join(minifits, "minifits.c"), "-o", obj], check=True)
        for app, src_file, imp in (("imageplane_disc_image", "imageplane/imageplane_disc_image.cpp", "imageplane.cpp"),
                                   ("caustic_discplane", "caustic/caustic_discplane.cpp", "imageplane.cpp"),
                                   ("caustic_plane", "caustic/caustic_plane.cpp", "imageplane.cpp"),
                                   ("caustic_sourceplane", "caustic/caustic_sourceplane.cpp", "imageplane.cpp")):
            if not any(missing(parname + ".fits") for parname in APP_PARS[app]):
                continue
            exe = os.path.join(tmp, app)
            subprocess.run(["/usr/bin/g++", "-O2", "-fopenmp", "-ffp-contract=off", "-w", f"-I{REF}/src", f"-I{REF}/src/raytracer",
                            f"-I{minifits}", f"{REF}/src/{src_file}", f"{REF}/src/raytracer/raytracer.cpp",
                            f"{REF}/src/raytracer/{imp}", obj, "-o", exe], check=True)
            for parname, text in APP_PARS[app].items():
                if not missing(parname + ".fits"):
                    continue
                out = os.path.join(GOLDEN, parname + ".fits")
                par = os.path.join(tmp, parname + ".par")
                with open(par, "w") as f:
                    f.write(f"outfile = {out}\n" + text)
                subprocess.run([exe, f"--parfile={par}"], check=True, stdout=subprocess.DEVNULL, cwd=os.path.join(tmp, "bin"))
                print(parname + ".fits:", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
