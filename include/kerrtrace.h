/*
 * kerrtrace.h -- C ABI of the B200-native Kerr null-geodesic ray-tracing engine.
 *
 * This is the drop-in boundary for ONE hot path of wilkinsdr/raytrace_cpu: per-ray integration of null
 * geodesics (fixed-step RK4 / adaptive RK45 in Boyer-Lindquist coordinates) plus the reduction of disc hits
 * into emissivity profiles and images.  The reference has no FFI layer; the path sits behind the C++ class
 * API `Raytracer<T>` / `PointSource<T>` / `ImagePlane<T>` / `ImagePlaneBundles<T>` / `RayDestination<T>`
 * (SURVEY.md section 8b).  Every entry point below cites the reference interface it replaces; all
 * `file:line` citations are relative to the reference tree (src/...).  The source-compatible C++ headers in
 * raytrace_cpu_b200/shim/ bind these entry points under the reference's own class names.
 *
 * Conventions: plain C, POD structs, caller-owned buffers, no exceptions.  Every function returns 0 on
 * success or a negative KT_ERR_* code; kt_last_error() returns a thread-local message.  One kt_ctx per
 * device (one process per GPU); a ctx is thread-compatible, not thread-safe.
 * Arithmetic is FP64 throughout (dtype "f64"): the reference instantiates float too but no app uses it.
 */
#ifndef KERRTRACE_H_
#define KERRTRACE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KT_VERSION 100

/* ---- error codes ---- */
#define KT_OK 0
#define KT_ERR_INVALID (-1)     /* bad argument */
#define KT_ERR_CUDA (-2)        /* CUDA runtime error (message has the cudaError string) */
#define KT_ERR_NOMEM (-3)       /* host or device allocation failed */
#define KT_ERR_UNSUPPORTED (-4) /* valid in the reference but not offloaded (user-defined RayDestination subclasses) */
#define KT_ERR_TIMEOUT (-5)     /* the optional watchdog fired: a trace kernel abandoned its launch */

/* ---- ray status bit flags: raytracer.h:58-63 ---- */
#define KT_STATUS_DEST (1 << 0)
#define KT_STATUS_HORIZON (1 << 1)
#define KT_STATUS_RLIM (1 << 2)
#define KT_STATUS_STEPLIM (1 << 3)
#define KT_STATUS_ERGO (1 << 4)
#define KT_STATUS_NEG_ENERGY (1 << 5)

/* ---- compile-time defaults of the reference: raytracer.h:19-46 ---- */
#define KT_PRECISION 100.0
#define KT_THETA_PRECISION 50.0
#define KT_MAXDT 1.0
#define KT_MAXDT_RLIM 100.0
#define KT_MAXDPHI 0.1
#define KT_RLIM 1000.0
#define KT_STEPLIM 10000000
#define KT_RK45_STEPLIM 100000
#define KT_MIN_STEP 1e-3

/* struct Ray<double>: raytracer.h:65-78.  Identical layout (144 bytes) so `Ray<double>* rays` can be passed
 * straight through.  steps: -1 unused slot, 0 initialised, +n traced, -n hit the step limit. */
typedef struct kt_ray {
    double t, r, theta, phi;
    double pt, pr, ptheta, pphi;
    double k, h, Q;
    double emit, redshift;
    int32_t steps;
    int32_t status;
    int32_t rdot_sign, thetadot_sign;
    int32_t rdot_flips;
    int32_t equatorial_crossings;
    double alpha, beta;
} kt_ray;

/* enum class Integrator: raytracer.h:83 */
enum { KT_EULER = 0, KT_RK4 = 1, KT_RK45 = 2 };

/* Stop surface.  THETALIM = run_raytrace(method, theta_max, r_max, ...) (raytracer.cpp:63); the others are
 * the three concrete RayDestination classes (ray_destination.h:85-204) for run_raytrace(dest, ...)
 * (raytracer.cpp:972).  User-defined RayDestination subclasses cannot run on the device. */
enum { KT_STOP_THETALIM = 0, KT_STOP_FLATDISC = 1, KT_STOP_ISCODISC = 2, KT_STOP_FLATPLANE = 3 };
typedef struct kt_stop {
    int32_t kind;
    int32_t _pad;
    double theta_lim;    /* THETALIM: theta_max; FLATDISC / ISCODISC: theta_lim */
    double r_isco, r_out;/* ISCODISC (r_out <= 0: no outer edge) */
    double incl, phi0, z_s; /* FLATPLANE (radians) */
} kt_stop;

/* Arithmetic mode.  EXACT evaluates every expression in the reference's operation order with IEEE
 * + - * / sqrt and no FMA contraction (only libm-class functions differ from glibc by <= 2 ulp); it is
 * the default and the benchmarked mode.  FAST evaluates the same right-hand side as a function through
 * cheaper, algebraically equivalent expressions with FMA contraction; same step
 * schedule, results differ at the 1e-15-per-step level -- the class of the reference built with
 * -march=native -- and it is ~1.5x faster. */
enum { KT_MATH_EXACT = 0, KT_MATH_FAST = 1 };

/* Members of Raytracer<T>: raytracer.h:87-99, ctor raytracer.cpp:13-37. */
typedef struct kt_tracer {
    double spin;            /* as stored by the reference: ImagePlane* pass -spin (imageplane.cpp:12) */
    double horizon;         /* kerr_horizon(spin) or set_boundary(r) */
    double precision;       /* ctor argument (TOL = 100 for PointSource) */
    double theta_precision; /* THETA_PRECISION */
    double max_tstep;       /* MAXDT */
    double max_phistep;     /* MAXDPHI */
    double maxtstep_rlim;   /* MAXDT_RLIM */
    double rk45_tol;        /* 1e-8; set_rk45_tol() */
} kt_tracer;

/* Arguments of one run_raytrace() call. */
typedef struct kt_run {
    int32_t integrator; /* KT_EULER (theta_max stop only, as raytracer.cpp:983 asserts), KT_RK4 or KT_RK45 */
    int32_t steplim;    /* <= 0: STEPLIM for RK4, RK45_STEPLIM for RK45 (raytracer.cpp:80) */
    double r_max;
    int32_t math_mode;  /* KT_MATH_EXACT (default 0) or KT_MATH_FAST */
    int32_t flags;      /* 0 (reference behaviour) or KT_RUN_RETIRE_STUCK */
} kt_run;

/* Opt-in, NOT reference behaviour: under RK45 a ray falling into the hole never crosses the horizon --
 * the per-step cap |r - r_h| / |pr| / precision has no floor (raytracer.cpp:1421-1434), so it creeps for RK45_STEPLIM =
 * 100 000 steps and ends as a step-limit failure (steps < 0), which every consumer discards.  At h = 10 these are 1-4 % of
 * the rays and 77-93 % of all steps.  With this flag an ingoing ray closer to the horizon than 1e-4 of the gap to the
 * innermost photon orbit (below which no radial turning point exists) is retired at once with STEPLIM set and its step
 * count negated: the same rays fail, after ~1 500 steps instead of 100 000; histograms and images are unchanged. */
#define KT_RUN_RETIRE_STUCK 1

/* PointSource<T> constructor arguments: pointsource.h:24, pointsource.cpp:11-64. */
typedef struct kt_pointsource {
    double pos[4]; /* t, r, theta, phi of the source */
    double V;      /* angular velocity of the source */
    double E;      /* photon energy in the source frame (default 1) */
    double dcosalpha, dbeta;
    double cosalpha0, cosalphamax; /* defaults -0.999999, 0.995 */
    double beta0, betamax;         /* defaults -0.995*pi, pi */
} kt_pointsource;

/* ImagePlane<T> / ImagePlaneBundles<T> constructor arguments: imageplane.h:27, imageplane.cpp:11-121,
 * imageplane_bundles.h:140-181.  spin is taken from kt_tracer (already negated). */
typedef struct kt_imageplane {
    double dist, incl_deg;
    double x0, xmax, dx;
    double y0, ymax, dy;
    double phi0;
    int32_t bundles;  /* 0: ImagePlane (1 ray/pixel); 1: ImagePlaneBundles (5 rays/pixel) */
    int32_t _pad;
    double eps_frac;  /* bundles only (default 0.01) */
} kt_imageplane;

/* Which rays of a source's flat index space this call processes.  Rays [begin, end) are cut into chunks
 * of `chunk` rays dealt round-robin: chunk c belongs to rank (c % nranks).  {0, count, 0, 1, 0} = all. */
typedef struct kt_shard {
    int64_t begin, end;
    int32_t rank, nranks;
    int64_t chunk; /* <= 0: automatic (about 256 chunks per rank, a multiple of 32 in [32, 65536]) */
} kt_shard;

/* Emissivity binning: emissivity.cpp:57-61,96-126; emissivity_rd.cpp:112-137. */
enum { KT_HIT_Z = 0 /* r cos(theta) < 1e-2, emissivity.cpp:103 */, KT_HIT_THETA = 1 /* theta >= theta_lim - 1e-3, emissivity_rd.cpp:116 */,
       KT_HIT_DEST = 2 /* status & DEST (no r_isco test), imageplane_disc_image_isco.cpp:129; images only */ };
enum { KT_REDSHIFT_V = 0 /* redshift(V, reverse, projradius) */, KT_REDSHIFT_DEST = 1 /* redshift(dest, reverse) */ };
typedef struct kt_emis_bins {
    int32_t Nr;
    int32_t logbin;
    double r_min;
    double dr;       /* log bins: ratio exp(log(r_disc/r_min)/Nr); linear: width */
    double r_isco;   /* kerr_isco<double>(spin,+1) as computed by the host (float-rounded, kerr.h:29-31) */
    double gamma;
    double n_primary;/* (double)(long)(((cmax-c0)/dc)*((bmax-b0)/db)) */
    int32_t hit_filter;
    int32_t redshift_mode;
    double theta_lim;  /* KT_HIT_THETA */
    double V;          /* KT_REDSHIFT_V: receiver angular velocity, -1 = Keplerian (raytracer.cpp:516-520) */
} kt_emis_bins;

enum { KT_CNT_SLOTS = 0, KT_CNT_TRACED = 1, KT_CNT_STEPS = 2, KT_CNT_DEST = 3, KT_CNT_HORIZON = 4, KT_CNT_RLIM = 5,
       KT_CNT_STEPLIM = 6, KT_CNT_DISC = 7, KT_NCOUNTERS = 8 };
/* Raw sums (division by N and by the proper area is the caller's job, emissivity.cpp:128-134).
 * counters: slots visited, rays traced (steps >= 0 before the trace), sum of |steps|, rays per terminal
 * status, disc_count (emissivity.cpp:124). */
typedef struct kt_emis_out {
    int64_t* N;            /* [Nr] */
    double* flux;          /* [Nr]  sum 1/(n_primary*g) */
    double* emis;          /* [Nr]  sum g^-gamma */
    double* redshift_sum;  /* [Nr] */
    double* time_sum;      /* [Nr] */
    int64_t counters[KT_NCOUNTERS];
} kt_emis_out;

/* Image binning: imageplane_disc_image.cpp:20-28,122-161.  Maps are [ix*img_Ny + iy] like Array2D. */
typedef struct kt_image_bins {
    int32_t img_Nx, img_Ny;
    double x0, y0, img_dx, img_dy;
    int32_t flip_image;
    int32_t _pad;
    double r_isco, r_disc;
    double q1, rb1, q2, rb2, q3; /* powerlaw3 emissivity */
    int32_t hit_filter;    /* KT_HIT_Z (imageplane_disc_image.cpp:130), KT_HIT_THETA (_rd.cpp:119), KT_HIT_DEST (_isco.cpp:129) */
    int32_t redshift_mode; /* KT_REDSHIFT_V: redshift(V, reverse) ; KT_REDSHIFT_DEST: redshift(dest, reverse) */
    int32_t reverse;       /* apps pass true except imageplane_disc_image_rd.cpp:113 (default-argument false) */
    int32_t _pad2;
    double theta_lim;      /* KT_HIT_THETA */
    double V;              /* KT_REDSHIFT_V receiver velocity (-1 = Keplerian) */
} kt_image_bins;
typedef struct kt_image_out {
    double *flux, *r, *phi, *enshift, *time, *emis; /* raw sums, [img_Nx*img_Ny] each */
    int32_t* N;
    int64_t counters[KT_NCOUNTERS];
} kt_image_out;

/* Ray-path dumping: the `outfile, write_step, write_rmax, write_rmin, write_cartesian` arguments of run_raytrace
 * (raytracer.cpp:63-68) and the write blocks of the five propagators (raytracer.cpp:293-312, 923-942, 1209-1228,
 * 1552-1571, 1850-1869).  After every write_step-th step that neither crossed the horizon nor reached the destination
 * a ray inside the radial window writes one point: (t, x, y, z) through cartesian() (kerr.h:40-56) when write_cartesian,
 * else (t, r, theta, phi).  A ray that leaves the window after its first point stops there. */
typedef struct kt_path_opts {
    int32_t write_step;
    int32_t write_cartesian;
    double write_rmax, write_rmin; /* < 0: unbounded */
} kt_path_opts;

/* Compact per-ray result for source-generated rays (caustic maps, parity tests): 56 bytes instead of 144. */
typedef struct kt_ray_min {
    double r, theta, phi, t, redshift;
    int32_t steps, status, rdot_flips, equatorial_crossings;
} kt_ray_min;

/* Caustic / critical-curve maps of the image plane: caustic_discplane.cpp:146-202 (per-pixel maps and the two
 * helper predicates), :209-334 (bundle Jacobian from the east/west/north/south satellite rays, kt_imageplane::bundles = 1),
 * :337-448 (grid-neighbour Jacobian, bundles = 0) and :451-493 (suppression of alternating-sign pixels).
 * Maps are [ix * img_Ny + iy] like Array2D. */
#define KT_CAUSTIC_SENTINEL 1e30    /* det_J where the image order changes inside the stencil (caustic_discplane.cpp:152) */
typedef struct kt_caustic_cfg {
    int32_t img_Nx, img_Ny;    /* rays per axis: the app's Nx + 1, Ny + 1 (caustic_discplane.cpp:129-130) */
    double r_isco, r_disc;     /* valid_hit: steps > 0, r_isco <= r < r_disc, redshift > 0 (caustic_discplane.cpp:177-182) */
    double dx, dy;             /* grid spacing, the denominators of the grid-neighbour Jacobian (caustic_discplane.cpp:435-438) */
    int32_t redshift_mode;     /* KT_REDSHIFT_DEST as the app (caustic_discplane.cpp:217) or KT_REDSHIFT_V with V */
    int32_t suppress;          /* 1: run the alternating-sign suppression pass (the app always does) */
    double V;
} kt_caustic_cfg;
typedef struct kt_caustic_out {
    double *det_J, *sign_J;    /* NaN / 0 where undefined, KT_CAUSTIC_SENTINEL / 0 at order boundaries */
    int32_t *order, *hit;      /* order = max(floor(|phi| / 2 pi), rdot_flips / 2), -1 where no hit */
    double *radius, *phi, *x_disc, *y_disc, *redshift;
    int64_t disc_count;        /* pixels whose centre ray hit the disc */
    int64_t cnt_horizon, cnt_rlim, cnt_steplim, cnt_outofrange, cnt_other; /* the app's diagnostic line (caustic_discplane.cpp:256-274) */
    int64_t suppressed;
    int64_t counters[KT_NCOUNTERS];
} kt_caustic_out;

typedef struct kt_ctx kt_ctx;
typedef struct kt_rayset kt_rayset;

/* ---- context ---- */
int kt_device_count(void);
kt_ctx* kt_create(int device);          /* NULL on failure, see kt_last_error() */
void kt_destroy(kt_ctx*);
const char* kt_last_error(void);
int kt_version(void);
int kt_set_stream(kt_ctx*, void* cuda_stream); /* run on a caller stream (e.g. torch's current stream); NULL = own stream */
int kt_synchronize(kt_ctx*);
int kt_sm_count(kt_ctx*);
int64_t kt_kernel_launches(kt_ctx*);    /* number of kernels this ctx has launched so far */
/* DFMA micro-benchmark: measured FP64 FMA peak in TFLOP/s (2 flop per FMA) -- the roofline denominator. */
int kt_fp64_peak(kt_ctx*, double* tflops);
/* bit-exactness self-test of the engine's IEEE division / square root (see csrc/kt_device.cuh: xdiv, xsqrt):
 * counts[5] = {wrong quotients, wrong roots, quotients sent to the slow path, roots sent to the slow path, samples};
 * mode 3 tests the exact mode's sincos against the CUDA library's: counts = {wrong sines, wrong cosines, 0, 0, samples} */
int kt_selftest_divsqrt(kt_ctx*, int mode, int samples_per_thread, int64_t* counts);
/* dependent-issue latency of one DFMA in SM cycles (one warp, serial chain): explains the latency-bound regime */
int kt_fp64_latency(kt_ctx*, double* cycles_per_op);
/* tuning: resident 128-thread CTAs per SM the trace kernel is built for (3, 4, 5 or 6; 0 = the measured best for the integrator and arithmetic mode).  Also
 * read from the environment variable KT_MIN_BLOCKS at kt_create(). */
int kt_set_min_blocks(kt_ctx*, int min_blocks);
/* Optional watchdog for the persistent trace kernels (seconds; 0 = off, the default; also read from the
 * environment variable KT_TIMEOUT_S).  When armed, a kernel that runs longer abandons its launch and the
 * next synchronising call (or kt_aborted) returns KT_ERR_TIMEOUT. */
int kt_set_timeout(kt_ctx*, double seconds);
int kt_aborted(kt_ctx*);
/* pinned host memory for the host-buffer entry points (plain malloc'ed buffers work too, only slower) */
void* kt_host_alloc(size_t bytes);
void kt_host_free(void*);
/* sizeof() of a POD struct of this header by name ("kt_ray", "kt_stop", ...), -1 if unknown: lets a
 * foreign-language binding verify its mirror of the layout */
int kt_abi_sizeof(const char* struct_name);

/* ---- helpers that restate host-side scalars of the reference ---- */
void kt_tracer_init(kt_tracer*, double spin, double precision); /* Raytracer ctor, raytracer.cpp:13-37 */
double kt_kerr_horizon(double a);                 /* kerr.h:13-20 */
double kt_kerr_isco(double a, int sign);          /* kerr.h:22-32 (float temporaries) */
int64_t kt_pointsource_count(const kt_pointsource*, int32_t* n_cosalpha, int32_t* n_beta); /* pointsource.cpp:12-17 */
int64_t kt_imageplane_count(const kt_imageplane*, int32_t* Nx, int32_t* Ny); /* imageplane.cpp:12-14, imageplane_bundles.h:146-150 */

/* ---- device-resident ray sets (AoS kt_ray in HBM; what Raytracer<T>::rays mirrors) ---- */
kt_rayset* kt_rayset_create(kt_ctx*, int64_t n);  /* all slots steps=-1, status=0 (raytracer.cpp:45-49), rest zero */
void kt_rayset_destroy(kt_rayset*);
int64_t kt_rayset_count(const kt_rayset*);
void* kt_rayset_device_ptr(kt_rayset*);
int kt_rayset_upload(kt_rayset*, const kt_ray* host, int64_t n);
int kt_rayset_download(kt_rayset*, kt_ray* host, int64_t n);
int kt_rayset_init_pointsource(kt_rayset*, const kt_tracer*, const kt_pointsource*); /* PointSource ctor */
int kt_rayset_init_imageplane(kt_rayset*, const kt_tracer*, const kt_imageplane*);   /* ImagePlane / Bundles ctor */
int kt_rayset_redshift_start(kt_rayset*, const kt_tracer*, double V, int reverse, int projradius); /* raytracer.cpp:342 */
int kt_rayset_trace(kt_rayset*, const kt_tracer*, const kt_stop*, const kt_run*);    /* run_raytrace, both overloads */
int kt_rayset_redshift(kt_rayset*, const kt_tracer*, double V, int reverse, int projradius, int motion); /* raytracer.cpp:420 */
int kt_rayset_redshift_dest(kt_rayset*, const kt_tracer*, const kt_stop*, int reverse); /* raytracer.cpp:450 */
int kt_rayset_range_phi(kt_rayset*, double lo, double hi);                            /* raytracer.cpp:603 */
int kt_rayset_calculate_momentum(kt_rayset*, const kt_tracer*);                       /* raytracer.cpp:704 */
/* bin a traced ray set whose redshift has already been computed (the apps' serial loops) */
int kt_rayset_bin_emissivity(kt_rayset*, const kt_emis_bins*, kt_emis_out* host_out);
int kt_rayset_bin_image(kt_rayset*, const kt_image_bins*, kt_image_out* host_out);

/* ---- host-buffer entry points (what the shim's Raytracer<double> methods call; copies included) ---- */
int kt_trace(kt_ctx*, const kt_tracer*, const kt_stop*, const kt_run*, kt_ray* rays, int64_t n);
/* run_raytrace(..., outfile != 0, write_step, write_rmax, write_rmin, write_cartesian), the serial ordered-output path
 * (raytracer.cpp:86-100, 995-1008), in two calls so that the caller owns every buffer:
 *   kt_trace_paths_count: offsets[n + 1] = exclusive prefix sums of the number of points each ray writes (offsets[n] =
 *                         total); `rays` is not modified.
 *   kt_trace_paths      : the trace itself -- `rays` updated in place exactly as kt_trace does, and ray i's points
 *                         written to points[4 * offsets[i] .. 4 * offsets[i + 1]) in step order.
 * The caller then emits the text (ray order, points, two blank lines after every ray that was propagated: rays with
 * steps < 0 or steps >= steplim are skipped entirely, raytracer.cpp:91-92,99) -- the shim's Raytracer does exactly that
 * through the reference's own TextOutput. */
int kt_trace_paths_count(kt_ctx*, const kt_tracer*, const kt_stop*, const kt_run*, const kt_ray* rays, int64_t n,
                         const kt_path_opts*, int64_t* offsets);
int kt_trace_paths(kt_ctx*, const kt_tracer*, const kt_stop*, const kt_run*, kt_ray* rays, int64_t n, const kt_path_opts*,
                   const int64_t* offsets, double* points);
int kt_redshift_start(kt_ctx*, const kt_tracer*, kt_ray* rays, int64_t n, double V, int reverse, int projradius);
int kt_redshift(kt_ctx*, const kt_tracer*, kt_ray* rays, int64_t n, double V, int reverse, int projradius, int motion);
int kt_redshift_dest(kt_ctx*, const kt_tracer*, const kt_stop*, kt_ray* rays, int64_t n, int reverse);
int kt_range_phi(kt_ctx*, kt_ray* rays, int64_t n, double lo, double hi);
int kt_pointsource_init(kt_ctx*, const kt_tracer*, const kt_pointsource*, kt_ray* rays, int64_t n);
int kt_imageplane_init(kt_ctx*, const kt_tracer*, const kt_imageplane*, kt_ray* rays, int64_t n);

/* ---- fused paths: init -> redshift_start -> trace -> range_phi -> redshift -> bin, rays never leave the SM ----
 * kt_pointsource_emissivity == the whole of emissivity{,_rd}.cpp:88-126 (and emissivity_rk45_test.cpp:69-113).
 * kt_imageplane_image       == imageplane_disc_image{,_rd,_isco}.cpp:109-161.
 * kt_imageplane_rays        == caustic_discplane.cpp:209-217 (trace + redshift; per-ray results for the maps).
 * The *_dev variants accumulate into caller-provided DEVICE buffers on the ctx stream without synchronising
 * (layout below), so partial results can be combined across GPUs with one NCCL reduce. */
int kt_pointsource_emissivity(kt_ctx*, const kt_pointsource*, const kt_tracer*, const kt_stop*, const kt_run*,
                              const kt_emis_bins*, const kt_shard* /* NULL = all rays */, kt_emis_out* host_out);
/* dev_hist: double[5*Nr + KT_NCOUNTERS] = {N, flux, emis, redshift_sum, time_sum}[Nr] then counters, all as
 * doubles (integers < 2^53 are exact) so a single sum-reduce combines them.  Must be zeroed by the caller. */
int kt_pointsource_emissivity_dev(kt_ctx*, const kt_pointsource*, const kt_tracer*, const kt_stop*, const kt_run*,
                                  const kt_emis_bins*, const kt_shard*, double* dev_hist);
int kt_imageplane_image(kt_ctx*, const kt_imageplane*, const kt_tracer*, const kt_stop*, const kt_run*,
                        const kt_image_bins*, const kt_shard*, kt_image_out* host_out);
/* dev_maps: double[7*img_Nx*img_Ny + KT_NCOUNTERS] = {flux, r, phi, enshift, time, emis, N} then counters. */
int kt_imageplane_image_dev(kt_ctx*, const kt_imageplane*, const kt_tracer*, const kt_stop*, const kt_run*,
                            const kt_image_bins*, const kt_shard*, double* dev_maps);
/* redshift_mode: KT_REDSHIFT_V with V (apps pass -1) or KT_REDSHIFT_DEST; reverse is always true for planes. */
int kt_imageplane_rays(kt_ctx*, const kt_imageplane*, const kt_tracer*, const kt_stop*, const kt_run*,
                       int redshift_mode, double V, int range_phi, kt_ray_min* host_out, int64_t n);

/* kt_caustic_discplane == caustic_discplane.cpp:204-493 in one call: image plane (or bundles) -> trace -> redshift ->
 * per-pixel maps -> Jacobian determinant -> suppression pass, all on the device; only the nine maps come back. */
int kt_caustic_discplane(kt_ctx*, const kt_imageplane*, const kt_tracer*, const kt_stop*, const kt_run*, const kt_caustic_cfg*,
                         kt_caustic_out* host_out);

#ifdef __cplusplus
}
#endif
#endif /* KERRTRACE_H_ */
