/*
 * kerr_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm for the hot path (wilkinsdr/raytrace_cpu,
 * src/raytracer/raytracer.cpp, pointsource.cpp, imageplane.cpp, imageplane_bundles.h, ray_destination.h,
 * include/kerr.h and the apps' binning loops).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (raytrace_cpu_b200/) never does.
 *
 * PARITY PIN: tests/test_oracle_vs_reference.py checks this restatement bit-for-bit (NaN-aware) against the
 * unmodified reference compiled from /root/reference (oracle/_ref/libref_oracle.so, built by oracle/Makefile)
 * and against the committed fixtures in tests/golden/ that scripts/make_golden.py generated from that build.
 *
 * POD types are shared with the product's public header so buffers are interchangeable in tests.
 */
#ifndef KERR_ORACLE_H_
#define KERR_ORACLE_H_

#include "../include/kerrtrace.h"

#ifdef __cplusplus
extern "C" {
#endif

double orc_kerr_horizon(double a);
double orc_kerr_isco(double a, int sign);
void orc_tracer_init(kt_tracer* tr, double spin, double precision);
int64_t orc_pointsource_count(const kt_pointsource* s, int32_t* n_cosalpha, int32_t* n_beta);
int64_t orc_imageplane_count(const kt_imageplane* p, int32_t* Nx, int32_t* Ny);

void orc_rays_reset(kt_ray* rays, int64_t n);
void orc_pointsource_init(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_pointsource* s);
void orc_imageplane_init(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_imageplane* p);
void orc_redshift_start(kt_ray* rays, int64_t n, const kt_tracer* tr, double V, int reverse, int projradius);
/* returns the number of rays actually propagated; nthreads <= 0 -> OpenMP default */
int64_t orc_trace(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_stop* stop, const kt_run* run, int nthreads);
/* the serial `outfile` path of run_raytrace: offsets[n+1], points[4 * offsets[n]] (NULL = count only) */
int64_t orc_trace_paths(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_stop* stop, const kt_run* run,
                        const kt_path_opts* opts, int64_t* offsets, double* points);
void orc_redshift(kt_ray* rays, int64_t n, const kt_tracer* tr, double V, int reverse, int projradius, int motion);
void orc_redshift_dest(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_stop* stop, int reverse);
void orc_range_phi(kt_ray* rays, int64_t n, double lo, double hi);
void orc_calculate_momentum(kt_ray* rays, int64_t n, const kt_tracer* tr);
void orc_bin_emissivity(const kt_ray* rays, int64_t n, const kt_emis_bins* b, kt_emis_out* out);
void orc_bin_image(const kt_ray* rays, int64_t n, const kt_image_bins* b, kt_image_out* out);
/* caustic_discplane.cpp:146-493 on an already traced + redshifted ray set (plane_Ny = ImagePlane::Ny) */
void orc_caustic_maps(const kt_ray* rays, int64_t n, int32_t plane_Ny, int bundles, double eps_x, double eps_y,
                      const kt_caustic_cfg* cfg, kt_caustic_out* out);
void orc_momentum_from_consts(double* p4, double k, double h, double Q, int rdot_sign, int thetadot_sign, double r,
                              double theta, double a);

#ifdef __cplusplus
}
#endif
#endif
