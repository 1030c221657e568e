#!/bin/sh
# Builds the reference's own, UNMODIFIED application sources against the kerrtrace shim:
# an overlay tree is made of symlinks into the reference checkout with src/raytracer/ replaced by the shim
# headers; nothing is copied.  Usage: build_apps.sh [reference-root] [output-dir]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
REF=${1:-/root/reference}
OUT=${2:-$HERE/_built}
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
OVL=$(mktemp -d)
trap 'rm -rf "$OVL"' EXIT
mkdir -p "$OVL/src" "$OUT"
# real directories holding per-file symlinks, so that the sources' relative includes ("../raytracer/...",
# "../include/...") resolve inside the overlay and not through the physical parent of a symlinked directory
ln -s "$REF/src/include" "$OVL/src/include"
for d in emissivity imageplane caustic tests ray_paths; do
    mkdir "$OVL/src/$d"
    ln -s "$REF/src/$d"/*.cpp "$OVL/src/$d/"
done
mkdir "$OVL/src/raytracer"
ln -s "$HERE/raytracer"/*.h "$OVL/src/raytracer/"
FLAGS="-O2 -fopenmp -w -I$OVL/src -I$REPO/include -L$REPO/raytrace_cpu_b200 -Wl,-rpath,$REPO/raytrace_cpu_b200"
$CXX $FLAGS "$OVL/src/emissivity/emissivity_rd.cpp" -o "$OUT/emissivity_rd" -lkerrtrace
# emissivity.cpp (the RK45 flagship app) does not compile upstream either: line 79 adds a double to a pointer
# (`disc_r + dr`).  The overlay holds a build-time copy with upstream's own one-token fix from emissivity_rd.cpp:88 (`dr`);
# nothing else differs and nothing is committed.
rm "$OVL/src/emissivity/emissivity.cpp"
sed 's/disc_r\[ir\]\*dr : disc_r + dr, spin)/disc_r[ir]*dr : dr, spin)/' "$REF/src/emissivity/emissivity.cpp" > "$OVL/src/emissivity/emissivity.cpp"
$CXX $FLAGS "$OVL/src/emissivity/emissivity.cpp" -o "$OUT/emissivity" -lkerrtrace
$CXX $FLAGS "$OVL/src/tests/emissivity_rk45_test.cpp" -o "$OUT/emissivity_rk45_test" -lkerrtrace
$CXX $FLAGS "$OVL/src/tests/integrator_perf_test.cpp" -o "$OUT/integrator_perf_test" -lkerrtrace
# Euler (the reference's own Euler-vs-RK4 consistency test) and the ray-path dumping applications (Euler + outfile)
$CXX $FLAGS "$OVL/src/tests/raytrace_rk4_test.cpp" -o "$OUT/raytrace_rk4_test" -lkerrtrace
$CXX $FLAGS "$OVL/src/ray_paths/trace_rays.cpp" -o "$OUT/trace_rays" -lkerrtrace
$CXX $FLAGS "$OVL/src/ray_paths/trace_rays_imageplane.cpp" -o "$OUT/trace_rays_imageplane" -lkerrtrace
# the FITS-writing applications, with the bundled minimal FITS writer standing in for cfitsio (absent from this image)
/usr/bin/gcc -O2 -c "$HERE/minifits/minifits.c" -o "$OVL/minifits.o"
for app in imageplane/imageplane_disc_image imageplane/imageplane_disc_image_rd imageplane/imageplane_disc_image_isco \
           caustic/caustic_discplane caustic/caustic_plane caustic/caustic_sourceplane; do
    $CXX $FLAGS -I"$HERE/minifits" "$OVL/src/$app.cpp" "$OVL/minifits.o" -o "$OUT/$(basename $app)" -lkerrtrace
done
echo "built: $(ls "$OUT")"
