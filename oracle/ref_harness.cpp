// oracle/ref_harness.cpp  --  TEST INFRASTRUCTURE ONLY (never on the product path).
//
// A thin extern "C" wrapper around the UNMODIFIED reference classes, compiled against the reference
// sources where they lie (/root/reference/src, never copied into this repo) by oracle/Makefile into
// oracle/_ref/libref_oracle.so.  It lets tests/ and bench.py's cpu_baseline / --impl reference leg drive
// the reference's own CPU implementation of the hot path:
//
//   PointSource<double>        pointsource.h:24,   pointsource.cpp:11-64
//   ImagePlane<double>         imageplane.h:27,    imageplane.cpp:11-121
//   ImagePlaneBundles<double>  imageplane_bundles.h:140-181
//   Raytracer<double>::run_raytrace (both overloads, with and without an outfile)   raytracer.cpp:63-127, 972-1034
//   redshift_start / redshift / range_phi              raytracer.cpp:342-622
//   FlatDisc / DiscWithISCO / FlatPlane destinations   ray_destination.h:85-204
//   kerr_horizon / kerr_isco / integrate_disc_area     kerr.h:13-32, disc.h:125-141
//
// Nothing here re-implements any arithmetic: every call forwards to the reference object.
#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>

#include "raytracer/pointsource.h"
#include "raytracer/imageplane.h"
#include "raytracer/imageplane_bundles.h"
#include "raytracer/ray_destination.h"
#include "include/disc.h"

namespace {

enum SourceKind { SRC_POINT = 0, SRC_PLANE = 1, SRC_BUNDLES = 2 };

struct Handle {
    int kind;
    PointSource<double>* ps = nullptr;
    ImagePlane<double>* ip = nullptr;
    ImagePlaneBundles<double>* ib = nullptr;
    Raytracer<double>* base = nullptr;
    double last_trace_seconds = 0;
};

// The reference prints progress chatter on std::cout from ctors and every phase; silence it.
struct Quiet {
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};

// stop kinds shared with include/kerrtrace.h (KT_STOP_*)
RayDestination<double>* make_dest(int kind, const double* p) {
    switch (kind) {
        case 1: return new FlatDiscDestination<double>(p[0]);
        case 2: return new DiscWithISCODestination<double>(p[1], p[2], p[0]);
        case 3: return new FlatPlaneDestination<double>(p[3], p[4], p[5]);
        default: return nullptr;
    }
}

Integrator to_integrator(int m) {
    return m == 2 ? Integrator::RK45 : (m == 1 ? Integrator::RK4 : Integrator::Euler);
}

}  // namespace

extern "C" {

int ref_sizeof_ray() { return (int)sizeof(Ray<double>); }

// OpenMP team size of the reference's parallel loop (torchrun exports OMP_NUM_THREADS=1; the CPU baseline must
// use every host thread it can)
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }

void* ref_pointsource_create(const double* pos, double V, double spin, double tol, double dcosalpha, double dbeta,
                             double cosalpha0, double cosalphamax, double beta0, double betamax, double E) {
    Quiet q;
    double p[4] = {pos[0], pos[1], pos[2], pos[3]};
    Handle* h = new Handle;
    h->kind = SRC_POINT;
    h->ps = new PointSource<double>(p, V, spin, tol, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax, E);
    h->base = h->ps;
    return h;
}

void* ref_imageplane_create(double dist, double incl_deg, double x0, double xmax, double dx, double y0, double ymax,
                            double dy, double spin, double phi, double precision) {
    Quiet q;
    Handle* h = new Handle;
    h->kind = SRC_PLANE;
    h->ip = new ImagePlane<double>(dist, incl_deg, x0, xmax, dx, y0, ymax, dy, spin, phi, precision);
    h->base = h->ip;
    return h;
}

void* ref_bundles_create(double dist, double incl_deg, double x0, double xmax, double dx, double y0, double ymax,
                         double dy, double spin, double phi, double precision, double eps_frac) {
    Quiet q;
    Handle* h = new Handle;
    h->kind = SRC_BUNDLES;
    h->ib = new ImagePlaneBundles<double>(dist, incl_deg, x0, xmax, dx, y0, ymax, dy, spin, phi, precision, eps_frac);
    h->base = h->ib;
    return h;
}

void ref_destroy(void* hv) {
    Quiet q;
    Handle* h = (Handle*)hv;
    delete h->ps;
    delete h->ip;
    delete h->ib;
    delete h;
}

int64_t ref_count(void* hv) { return ((Handle*)hv)->base->get_count(); }
void* ref_rays(void* hv) { return ((Handle*)hv)->base->rays; }
double ref_horizon(void* hv) { return ((Handle*)hv)->base->calculate_horizon(); }
double ref_last_trace_seconds(void* hv) { return ((Handle*)hv)->last_trace_seconds; }

void ref_set_rk45_tol(void* hv, double tol) { ((Handle*)hv)->base->set_rk45_tol(tol); }
void ref_set_max_tstep(void* hv, double m, double rlim) { ((Handle*)hv)->base->set_max_tstep(m, rlim); }
void ref_set_max_phistep(void* hv, double m) { ((Handle*)hv)->base->set_max_phistep(m); }

// the source's own no-argument redshift_start() (pointsource.cpp:67, imageplane.cpp:124, imageplane_bundles.h:192)
void ref_source_redshift_start(void* hv) {
    Quiet q;
    Handle* h = (Handle*)hv;
    if (h->kind == SRC_POINT) h->ps->redshift_start();
    else if (h->kind == SRC_PLANE) h->ip->redshift_start();
    else h->ib->redshift_start();
}

void ref_redshift_start(void* hv, double V, int reverse, int projradius) {
    Quiet q;
    ((Handle*)hv)->base->redshift_start(V, reverse != 0, projradius != 0);
}

// run_raytrace(method, theta_max, r_max, show_progress=0, ..., steplim)   raytracer.cpp:63
void ref_run_thetalim(void* hv, int method, double theta_max, double r_max, int steplim) {
    Quiet q;
    Handle* h = (Handle*)hv;
    auto t0 = std::chrono::high_resolution_clock::now();
    h->base->run_raytrace(to_integrator(method), theta_max, r_max, 0, nullptr, 1, -1, -1, true, steplim);
    auto t1 = std::chrono::high_resolution_clock::now();
    h->last_trace_seconds = std::chrono::duration<double>(t1 - t0).count();
}

// run_raytrace(method, theta_max, r_max, 0, &outfile, write_step, write_rmax, write_rmin, write_cartesian, steplim): the
// serial ray-path dumping branch (raytracer.cpp:86-100).  The text goes to `filename` through the reference's own TextOutput;
// precision 17 makes the columns round-trip to the exact doubles (the apps use the default 8).
void ref_run_thetalim_paths(void* hv, int method, double theta_max, double r_max, int steplim, const char* filename,
                            int precision, int write_step, double write_rmax, double write_rmin, int write_cartesian) {
    Quiet q;
    Handle* h = (Handle*)hv;
    TextOutput out(filename, false, precision, precision + 12);
    h->base->run_raytrace(to_integrator(method), theta_max, r_max, 0, &out, write_step, write_rmax, write_rmin,
                          write_cartesian != 0, steplim);
    out.close();
}

void ref_run_dest_paths(void* hv, int method, int stop_kind, const double* stop_par, double r_max, int steplim,
                        const char* filename, int precision, int write_step, double write_rmax, double write_rmin,
                        int write_cartesian) {
    Quiet q;
    Handle* h = (Handle*)hv;
    RayDestination<double>* d = make_dest(stop_kind, stop_par);
    TextOutput out(filename, false, precision, precision + 12);
    h->base->run_raytrace(d, to_integrator(method), r_max, 0, &out, write_step, write_rmax, write_rmin, write_cartesian != 0,
                          steplim);
    out.close();
    delete d;
}

// run_raytrace(dest, method, r_max, show_progress=0, ..., steplim)        raytracer.cpp:972
// stop_par = {theta_lim, r_isco, r_out, incl, phi0, z_s}
void ref_run_dest(void* hv, int method, int stop_kind, const double* stop_par, double r_max, int steplim) {
    Quiet q;
    Handle* h = (Handle*)hv;
    RayDestination<double>* d = make_dest(stop_kind, stop_par);
    auto t0 = std::chrono::high_resolution_clock::now();
    h->base->run_raytrace(d, to_integrator(method), r_max, 0, nullptr, 1, -1, -1, true, steplim);
    auto t1 = std::chrono::high_resolution_clock::now();
    h->last_trace_seconds = std::chrono::duration<double>(t1 - t0).count();
    delete d;
}

void ref_redshift_V(void* hv, double V, int reverse, int projradius, int motion) {
    Quiet q;
    ((Handle*)hv)->base->redshift(V, reverse != 0, projradius != 0, motion);
}

void ref_redshift_dest(void* hv, int stop_kind, const double* stop_par, int reverse) {
    Quiet q;
    RayDestination<double>* d = make_dest(stop_kind, stop_par);
    ((Handle*)hv)->base->redshift(d, reverse != 0);
    delete d;
}

void ref_range_phi(void* hv, double lo, double hi) {
    Quiet q;
    ((Handle*)hv)->base->range_phi(lo, hi);
}

void ref_calculate_momentum(void* hv) { ((Handle*)hv)->base->calculate_momentum(); }

double ref_kerr_horizon(double a) { return kerr_horizon<double>(a); }
double ref_kerr_isco(double a, int sign) { return kerr_isco<double>(a, sign); }
double ref_integrate_disc_area(double rmin, double rmax, double a) { return integrate_disc_area<double>(rmin, rmax, a); }
void ref_cartesian(double* xyz, double r, double theta, double phi, double a) {
    cartesian<double>(xyz[0], xyz[1], xyz[2], r, theta, phi, a);
}
void ref_momentum_from_consts(double* p, double k, double h, double Q, int rs, int ts, double r, double theta,
                              double phi, double a) {
    momentum_from_consts<double>(p[0], p[1], p[2], p[3], k, h, Q, rs, ts, r, theta, phi, a);
}

}  // extern "C"
