// raytracer.h (kerrtrace shim) -- source-compatible replacement of the reference's src/raytracer/raytracer.h.
//
// Drop this directory in place of the reference's src/raytracer/ (headers only: raytracer.cpp, pointsource.cpp and
// imageplane.cpp are no longer compiled) and link libkerrtrace.so: the emissivity / imageplane / caustic applications
// then build UNCHANGED and run their geodesic integration on a B200 through the C ABI of include/kerrtrace.h.
// See INTEGRATION.md.  Everything an application touches keeps its name, signature and default arguments
// (reference raytracer.h:19-198): struct Ray<T>, RAY_STATUS_*, the tuning macros, enum class Integrator and
// class Raytracer<T> with its public `rays` array.
//
// What runs where: run_raytrace (Euler, RK4, RK45; theta_max or one of the three concrete RayDestination classes; with
// or without an outfile for ray-path dumping), redshift_start, redshift, range_phi and calculate_momentum run on the GPU
// and leave their results in the host array `rays`, exactly as the CPU code does.  With an outfile the engine returns
// every ray's points and this header writes them through the reference's own TextOutput, in ray order, so the text is
// formatted by the same code as in the CPU build.  User-defined RayDestination subclasses and T = float are not
// offloaded: they throw std::runtime_error (keep the reference's CPU build for those).
#ifndef RAYTRACER_H_
#define RAYTRACER_H_

// tuning constants of the reference (raytracer.h:19-46); applications use TOL and PRECISION
#define PRECISION 100
#define TOL 100
#define THETA_PRECISION 50
#define MAXDT 1
#define MAXDT_RLIM 100
#define MAXDPHI 0.1
#define RLIM 1000
#define STEPLIM 10000000
#define RK45_STEPLIM 100000
#define THREAD_STEPLIM 10000000
#define MIN_STEP 1E-3
#define COUNT_MIN 100

#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
using namespace std;   // the reference header does this and the applications rely on it

#include "../include/kerr.h"
#include "../include/text_output.h"
#include "../include/progress_bar.h"

#include "kerrtrace.h"

constexpr int RAY_STATUS_DEST       = KT_STATUS_DEST;
constexpr int RAY_STATUS_HORIZON    = KT_STATUS_HORIZON;
constexpr int RAY_STATUS_RLIM       = KT_STATUS_RLIM;
constexpr int RAY_STATUS_STEPLIM    = KT_STATUS_STEPLIM;
constexpr int RAY_STATUS_ERGO       = KT_STATUS_ERGO;
constexpr int RAY_STATUS_NEG_ENERGY = KT_STATUS_NEG_ENERGY;

// Same members, same order, same size as the reference's Ray<T> (and as kt_ray for T = double).
template <typename T>
struct Ray
{
    T t, r, theta, phi;
    T pt, pr, ptheta, pphi;
    T k, h, Q;
    T emit, redshift;
    int steps;
    int status;
    int rdot_sign, thetadot_sign;
    int rdot_flips;
    int equatorial_crossings;
    T alpha, beta;
};
static_assert(sizeof(Ray<double>) == sizeof(kt_ray), "Ray<double> must stay layout-compatible with kt_ray");

template <typename T> class RayDestination;

enum class Integrator { Euler, RK4, RK45 };

namespace kerrtrace_shim {

inline bool quiet() { return std::getenv("KERRTRACE_QUIET") != nullptr; }

inline void check(int rc, const char* what)
{
    if (rc != KT_OK) throw std::runtime_error(std::string("kerrtrace: ") + what + ": " + kt_last_error());
}

// one engine context per process (device from KERRTRACE_DEVICE, default 0)
inline kt_ctx* context()
{
    static kt_ctx* ctx = nullptr;
    if (!ctx) {
        const char* dev = std::getenv("KERRTRACE_DEVICE");
        ctx = kt_create(dev ? std::atoi(dev) : 0);
        if (!ctx) throw std::runtime_error(std::string("kerrtrace: ") + kt_last_error());
    }
    return ctx;
}

template <typename T> kt_stop stop_of(const RayDestination<T>* dest);   // ray_destination.h

}  // namespace kerrtrace_shim

template <typename T>
class Raytracer
{
    static_assert(std::is_same<T, double>::value, "the kerrtrace shim offloads Raytracer<double> only");

private:
    kt_tracer cfg;      // precision, theta_precision, max_tstep, max_phistep, maxtstep_rlim, rk45_tol (+ spin, horizon)
    int math_mode;

    void sync_cfg() { cfg.spin = spin; cfg.horizon = horizon; }

protected:
    int nRays;
    T spin;
    T horizon;

    kt_ctx* engine() const { return kerrtrace_shim::context(); }
    const kt_tracer* tracer() { sync_cfg(); return &cfg; }

public:
    Ray<T>* rays;

    Raytracer(int num_rays, T spin_par, T init_precision = PRECISION, T init_max_phistep = MAXDPHI, T init_max_tstep = MAXDT)
        : nRays(num_rays), spin(spin_par)
    {
        kt_tracer_init(&cfg, spin_par, init_precision);
        cfg.max_phistep = init_max_phistep;
        cfg.max_tstep = init_max_tstep;
        horizon = cfg.horizon;
        const char* mm = std::getenv("KERRTRACE_MATH");
        math_mode = (mm && std::string(mm) == "fast") ? KT_MATH_FAST : KT_MATH_EXACT;
        if (!kerrtrace_shim::quiet()) {
            cout << "Event horizon at " << horizon << endl;
            cout << "Allocating memory (" << nRays * sizeof(Ray<T>) / (1 << 20) << "MB)" << endl;
        }
        // Create (or look up) the engine context FIRST: kt_create selects KERRTRACE_DEVICE, so the pinned allocation below opens
        // its CUDA context on that device and not on device 0 (one process per GPU must not touch GPU 0 from every rank).
        (void)engine();
        // pinned host memory: the copies of every offloaded call then run at full PCIe rate
        rays = static_cast<Ray<T>*>(kt_host_alloc(sizeof(Ray<T>) * (nRays > 0 ? nRays : 1)));
        if (!rays) throw std::bad_alloc();
        for (int i = 0; i < nRays; i++) {
            rays[i] = Ray<T>();      // the reference leaves everything but steps/status uninitialised
            rays[i].steps = -1;
            rays[i].status = 0;
        }
    }

    ~Raytracer()
    {
        if (!kerrtrace_shim::quiet()) cout << "Cleaning up raytracer" << endl;
        kt_host_free(rays);
    }

    Raytracer(const Raytracer&) = delete;
    Raytracer& operator=(const Raytracer&) = delete;

    void run_raytrace(Integrator method = Integrator::Euler, T theta_max = M_PI / 2, T r_max = 1000, int show_progress = 1,
                      TextOutput* outfile = 0, int write_step = 1, T write_rmax = -1, T write_rmin = -1,
                      bool write_cartesian = true, int steplim = -1)
    {
        kt_stop stop = {};
        stop.kind = KT_STOP_THETALIM;
        stop.theta_lim = theta_max;
        trace(stop, method, r_max, outfile, steplim, write_step, write_rmax, write_rmin, write_cartesian);
    }

    void run_raytrace(RayDestination<T>* dest, Integrator method = Integrator::Euler, T r_max = 1000, int show_progress = 1,
                      TextOutput* outfile = 0, int write_step = 1, T write_rmax = -1, T write_rmin = -1,
                      bool write_cartesian = true, int steplim = -1)
    {
        assert(method != Integrator::Euler && "Integrator::Euler does not support RayDestination stopping conditions");
        trace(kerrtrace_shim::stop_of(dest), method, r_max, outfile, steplim, write_step, write_rmax, write_rmin, write_cartesian);
    }

    void redshift_start(T V, bool reverse = false, bool projradius = false)
    {
        if (!kerrtrace_shim::quiet()) cout << "Calculating initial energies" << endl;
        kerrtrace_shim::check(kt_redshift_start(engine(), tracer(), as_kt(), nRays, V, reverse, projradius), "redshift_start");
    }

    void redshift(T V, bool reverse = false, bool projradius = false, int motion = 0)
    {
        if (!kerrtrace_shim::quiet()) cout << "Calculating ray redshifts..." << endl;
        kerrtrace_shim::check(kt_redshift(engine(), tracer(), as_kt(), nRays, V, reverse, projradius, motion), "redshift");
    }

    void redshift(RayDestination<T>* dest, bool reverse = false, bool projradius = false, int motion = 0)
    {
        if (!kerrtrace_shim::quiet()) cout << "Calculating ray redshifts..." << endl;
        const kt_stop stop = kerrtrace_shim::stop_of(dest);
        kerrtrace_shim::check(kt_redshift_dest(engine(), tracer(), &stop, as_kt(), nRays, reverse), "redshift(dest)");
    }

    void range_phi(T min = -1 * M_PI, T max = M_PI)
    {
        if (!kerrtrace_shim::quiet()) cout << "Putting phi co-ordinate into range [" << min << "," << max << "]" << endl;
        kerrtrace_shim::check(kt_range_phi(engine(), as_kt(), nRays, min, max), "range_phi");
    }

    void calculate_momentum()
    {
        // a cheap O(n) loop: evaluated on the host with the reference's own inline function (kerr.h:299-335)
        for (int i = 0; i < nRays; i++)
            momentum_from_consts<T>(rays[i].pt, rays[i].pr, rays[i].ptheta, rays[i].pphi, rays[i].k, rays[i].h, rays[i].Q,
                                    rays[i].rdot_sign, rays[i].thetadot_sign, rays[i].r, rays[i].theta, rays[i].phi, spin);
    }

    int get_count() { return nRays; }

    void set_boundary(T r = -1) { horizon = (r > 0) ? r : kerr_horizon<T>(spin); }
    T calculate_horizon() { return kerr_horizon<T>(spin); }

    // both overloads are no-ops in the reference (self-assignment, raytracer.h:169-178) and stay no-ops here
    void set_precision(T) {}
    void set_precision(T, T) {}

    void set_rk45_tol(T tol) { cfg.rk45_tol = tol; }
    T get_rk45_tol() const { return cfg.rk45_tol; }
    void set_max_tstep(T max, T rlim = MAXDT_RLIM) { cfg.max_tstep = max; cfg.maxtstep_rlim = rlim; }
    void set_max_phistep(T max) { cfg.max_phistep = max; }

    // shim extension: KT_MATH_EXACT (default) or KT_MATH_FAST; also selectable with KERRTRACE_MATH=fast
    void set_math_mode(int mode) { math_mode = mode; }

protected:
    kt_ray* as_kt() { return reinterpret_cast<kt_ray*>(rays); }

private:
    void trace(const kt_stop& stop, Integrator method, T r_max, TextOutput* outfile, int steplim, int write_step, T write_rmax,
               T write_rmin, bool write_cartesian)
    {
        static const char* names[] = {"Running raytracer...", "Running raytracer (RK4)...", "Running raytracer (RK45/DOPRI5)..."};
        if (!kerrtrace_shim::quiet()) cout << names[static_cast<int>(method)] << endl;     // raytracer.cpp:75-81
        kt_run run = {};
        run.integrator = (method == Integrator::RK45) ? KT_RK45 : (method == Integrator::RK4) ? KT_RK4 : KT_EULER;
        run.steplim = steplim;
        run.r_max = r_max;
        run.math_mode = math_mode;
        if (outfile == nullptr) {
            kerrtrace_shim::check(kt_trace(engine(), tracer(), &stop, &run, as_kt(), nRays), "run_raytrace");
            return;
        }
        // serial path of the reference (raytracer.cpp:86-100): rays in index order, one line per written point, two blank
        // lines after every ray that was propagated; rays with steps < 0 or steps >= steplim are skipped entirely
        kt_path_opts opts = {};
        opts.write_step = write_step;
        opts.write_cartesian = write_cartesian ? 1 : 0;
        opts.write_rmax = write_rmax;
        opts.write_rmin = write_rmin;
        const int effective_steplim = (steplim > 0) ? steplim : (method == Integrator::RK45) ? RK45_STEPLIM : STEPLIM;
        std::vector<int> before(nRays > 0 ? nRays : 1);
        for (int i = 0; i < nRays; i++) before[i] = rays[i].steps;
        std::vector<int64_t> offsets(nRays + 1, 0);
        kerrtrace_shim::check(kt_trace_paths_count(engine(), tracer(), &stop, &run, as_kt(), nRays, &opts, offsets.data()),
                              "run_raytrace (path count)");
        std::vector<double> pts(4 * (offsets[nRays] > 0 ? offsets[nRays] : 1));
        kerrtrace_shim::check(kt_trace_paths(engine(), tracer(), &stop, &run, as_kt(), nRays, &opts, offsets.data(), pts.data()),
                              "run_raytrace (paths)");
        for (int ray = 0; ray < nRays; ray++) {
            if (before[ray] < 0) continue;
            else if (before[ray] >= effective_steplim) continue;
            for (int64_t q = offsets[ray]; q < offsets[ray + 1]; q++)
                (*outfile) << pts[4 * q] << pts[4 * q + 1] << pts[4 * q + 2] << pts[4 * q + 3] << endl;
            outfile->newline(2);
        }
    }
};

#endif /* RAYTRACER_H_ */
