// kt_api.cu -- the C ABI of include/kerrtrace.h: contexts, device-resident ray sets, launch orchestration.
// Host code only (kernels live in kt_kernels.cu, compiled once per arithmetic mode).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "kt_types.h"
#include "kt_kernels.h"

using kt::CausticMaps;
using kt::KernelTable;
using kt::LaunchCfg;
using kt::PlaneFrame;
using kt::PointSourceFrame;
using kt::TraceParams;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define KT_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) return fail(KT_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* kt_last_error(void) { return g_err; }
extern "C" int kt_version(void) { return KT_VERSION; }

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
struct kt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;     // the stream work is issued on (own_stream or the caller's)
    unsigned long long* d_queue = nullptr;
    double* d_scratch = nullptr;       // small reduction buffer (histograms, counters)
    size_t scratch_bytes = 0;
    kt_ray* d_rays = nullptr;          // staging for the host-buffer entry points
    size_t rays_bytes = 0;
    double* h_pinned = nullptr;        // pinned bounce buffer for small results
    size_t pinned_bytes = 0;
    long long launches = 0;
    int min_blocks = 0;
    double timeout_s = 0;              // watchdog for the persistent kernels (0 = off)
    // chunk pipeline of the host-buffer trace: 4 streams, each with its own work queue and statistics block
    static const int NPIPE = 4;
    cudaStream_t pipe[NPIPE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_done[NPIPE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_start = nullptr;
    unsigned long long* d_pipe_queue = nullptr;   // NPIPE counters
    double* d_pipe_cnt = nullptr;                 // NPIPE x KT_NCOUNTERS
    long long pipe_chunk = 1ll << 40;             // rays per chunk: default = no chunking (see kt_trace); KT_PIPE_CHUNK overrides
    int carveout = -1;                            // KT_CARVEOUT: preferred shared-memory carve-out of the trace kernels (percent)
    int zerocopy = 1;                             // kt_trace on pinned host arrays: kernel addresses them directly (KT_ZEROCOPY=0: staged)
};

struct kt_rayset {
    kt_ctx* ctx;
    long long n;
    kt_ray* d;
};

extern "C" int kt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int ctx_bind(kt_ctx* c)
{
    if (!c) return fail(KT_ERR_INVALID, "null context");
    KT_CUDA(cudaSetDevice(c->device));
    return KT_OK;
}

static int ensure_scratch(kt_ctx* c, size_t bytes)
{
    if (c->scratch_bytes >= bytes) return KT_OK;
    if (c->d_scratch) cudaFree(c->d_scratch);
    c->d_scratch = nullptr;
    c->scratch_bytes = 0;
    if (cudaMalloc(&c->d_scratch, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(KT_ERR_NOMEM, "device allocation of %zu bytes failed", bytes);
    }
    c->scratch_bytes = bytes;
    return KT_OK;
}

static int ensure_pinned(kt_ctx* c, size_t bytes)
{
    if (c->pinned_bytes >= bytes) return KT_OK;
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr;
    c->pinned_bytes = 0;
    if (cudaHostAlloc(&c->h_pinned, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return fail(KT_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes);
    }
    c->pinned_bytes = bytes;
    return KT_OK;
}

static int ensure_rays(kt_ctx* c, size_t bytes)
{
    if (c->rays_bytes >= bytes) return KT_OK;
    if (c->d_rays) cudaFree(c->d_rays);
    c->d_rays = nullptr;
    c->rays_bytes = 0;
    if (cudaMalloc(&c->d_rays, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(KT_ERR_NOMEM, "device allocation of %zu bytes failed", bytes);
    }
    c->rays_bytes = bytes;
    return KT_OK;
}

extern "C" kt_ctx* kt_create(int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(KT_ERR_CUDA, "no CUDA device available (%s): the kerrtrace engine has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        fail(KT_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
        return nullptr;
    }
    kt_ctx* c = new (std::nothrow) kt_ctx;
    if (!c) {
        fail(KT_ERR_NOMEM, "host allocation of the context failed");
        return nullptr;
    }
    c->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c->d_queue, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(c->d_queue, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        fail(KT_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        kt_destroy(c);      // (frees whichever of the stream / queue already exists)
        return nullptr;
    }
    c->sm_count = prop.multiProcessorCount;
    c->stream = c->own_stream;
    const char* mb = getenv("KT_MIN_BLOCKS");
    c->min_blocks = mb ? atoi(mb) : 0;
    const char* to = getenv("KT_TIMEOUT_S");
    c->timeout_s = to ? atof(to) : 0;
    const char* pc = getenv("KT_PIPE_CHUNK");
    if (pc && atoll(pc) > 0) c->pipe_chunk = atoll(pc);
    const char* co = getenv("KT_CARVEOUT");
    if (co) c->carveout = atoi(co);
    const char* zc = getenv("KT_ZEROCOPY");
    if (zc) c->zerocopy = atoi(zc) != 0;
    bool pipe_ok = cudaEventCreateWithFlags(&c->pipe_start, cudaEventDisableTiming) == cudaSuccess &&
                   cudaMalloc(&c->d_pipe_queue, kt_ctx::NPIPE * sizeof(unsigned long long)) == cudaSuccess &&
                   cudaMalloc(&c->d_pipe_cnt, kt_ctx::NPIPE * KT_NCOUNTERS * sizeof(double)) == cudaSuccess;
    for (int i = 0; i < kt_ctx::NPIPE && pipe_ok; i++)
        pipe_ok = cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&c->pipe_done[i], cudaEventDisableTiming) == cudaSuccess;
    if (!pipe_ok) {
        fail(KT_ERR_CUDA, "context creation failed (pipeline streams): %s", cudaGetErrorString(cudaGetLastError()));
        kt_destroy(c);
        return nullptr;
    }
    if (ensure_scratch(c, 1 << 20) != KT_OK || ensure_pinned(c, 1 << 20) != KT_OK) {
        kt_destroy(c);
        return nullptr;
    }
    return c;
}

extern "C" void kt_destroy(kt_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    // c->stream may be a caller-owned stream (kt_set_stream) that the caller has already destroyed -- Python finalisers run in
    // any order at interpreter exit -- so it is never touched here: the whole device is drained instead.
    cudaDeviceSynchronize();
    if (c->d_queue) cudaFree(c->d_queue);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->d_rays) cudaFree(c->d_rays);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (int i = 0; i < kt_ctx::NPIPE; i++) {
        if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
        if (c->pipe_done[i]) cudaEventDestroy(c->pipe_done[i]);
    }
    if (c->pipe_start) cudaEventDestroy(c->pipe_start);
    if (c->d_pipe_queue) cudaFree(c->d_pipe_queue);
    if (c->d_pipe_cnt) cudaFree(c->d_pipe_cnt);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int kt_set_stream(kt_ctx* c, void* s)
{
    if (!c) return fail(KT_ERR_INVALID, "null context");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return KT_OK;
}

extern "C" int kt_synchronize(kt_ctx* c)
{
    if (int rc = ctx_bind(c)) return rc;
    KT_CUDA(cudaStreamSynchronize(c->stream));
    return KT_OK;
}

extern "C" int kt_sm_count(kt_ctx* c) { return c ? c->sm_count : 0; }
extern "C" int64_t kt_kernel_launches(kt_ctx* c) { return c ? c->launches : 0; }
extern "C" int kt_set_min_blocks(kt_ctx* c, int mb)
{
    if (!c) return fail(KT_ERR_INVALID, "null context");
    c->min_blocks = mb;
    return KT_OK;
}

extern "C" int kt_set_timeout(kt_ctx* c, double seconds)
{
    if (!c) return fail(KT_ERR_INVALID, "null context");
    c->timeout_s = seconds;
    return KT_OK;
}

// synchronise and report (then clear) whether a watchdog-armed kernel abandoned its launch
extern "C" int kt_aborted(kt_ctx* c)
{
    if (int rc = ctx_bind(c)) return rc;
    int flag = 0;
    KT_CUDA(cudaMemcpyAsync(&flag, c->d_queue + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    if (flag) {
        KT_CUDA(cudaMemsetAsync(c->d_queue + 1, 0, sizeof(unsigned long long), c->stream));
        return fail(KT_ERR_TIMEOUT, "trace kernel exceeded the %.1f s watchdog and abandoned its launch: results are incomplete", c->timeout_s);
    }
    return KT_OK;
}

extern "C" void* kt_host_alloc(size_t bytes)
{
    void* p = nullptr;
    // mapped + portable: every device of the process can address it, which is what lets kt_trace run zero-copy on it
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        fail(KT_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes);
        return nullptr;
    }
    return p;
}
extern "C" void kt_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

static LaunchCfg cfg_of(kt_ctx* c)
{
    LaunchCfg cfg;
    cfg.stream = c->stream;
    cfg.sm_count = c->sm_count;
    cfg.min_blocks = c->min_blocks;
    cfg.carveout = c->carveout;
    return cfg;
}

static const KernelTable* table_for(int math_mode)
{
    return math_mode == KT_MATH_FAST ? kt_kernel_table_fast() : kt_kernel_table_exact();
}

// ---------------------------------------------------------------------------------------------------------
// host-side scalars of the reference
// ---------------------------------------------------------------------------------------------------------
extern "C" double kt_kerr_horizon(double a) { return 1 + 1 * std::sqrt((1 - a) * (1 + a)); } // kerr.h:13-20

extern "C" double kt_kerr_isco(double a, int sign)
{
    // kerr.h:22-32: A and B are held in float and the closing expression is evaluated in float
    const float A = (float)(1. + std::pow(1. - a * a, 1. / 3.) * (std::pow(1. + a, 1. / 3.) + std::pow(1. - a, 1. / 3.)));
    const float B = (float)std::sqrt(3. * a * a + (double)(A * A));
    const float lhs = 3 + B;
    const float inner = (3 - A) * (3 + A + 2 * B);
    const float res = lhs - (float)sign * sqrtf(inner);
    return (double)res;
}

extern "C" void kt_tracer_init(kt_tracer* tr, double spin, double precision)
{
    tr->spin = spin;
    tr->horizon = kt_kerr_horizon(spin);
    tr->precision = precision;
    tr->theta_precision = KT_THETA_PRECISION;
    tr->max_tstep = KT_MAXDT;
    tr->max_phistep = KT_MAXDPHI;
    tr->maxtstep_rlim = KT_MAXDT_RLIM;
    tr->rk45_tol = 1e-8;
}

// A grid axis is usable when its step is finite and non-zero and the slot count (lo..hi)/step + 1 is a finite number in
// [1, 2^31): anything else would cast inf / NaN to an integer on the host and divide by a zero count on the device.
static bool axis_ok(double lo, double hi, double step)
{
    if (!std::isfinite(lo) || !std::isfinite(hi) || !std::isfinite(step) || step == 0) return false;
    const double f = ((hi - lo) / step) + 1;
    return std::isfinite(f) && f >= 1 && f < 2147483648.0;
}

extern "C" int64_t kt_pointsource_count(const kt_pointsource* s, int32_t* n_cosalpha, int32_t* n_beta)
{
    if (!s || !axis_ok(s->cosalpha0, s->cosalphamax, s->dcosalpha) || !axis_ok(s->beta0, s->betamax, s->dbeta)) {
        if (n_cosalpha) *n_cosalpha = 0;
        if (n_beta) *n_beta = 0;
        fail(KT_ERR_INVALID, "point-source grid: steps must be finite and non-zero and each axis must hold 1 .. 2^31-1 slots");
        return -1;
    }
    const double fa = ((s->cosalphamax - s->cosalpha0) / s->dcosalpha) + 1;
    const double fb = ((s->betamax - s->beta0) / s->dbeta) + 1;
    if (n_cosalpha) *n_cosalpha = (int32_t)fa;
    if (n_beta) *n_beta = (int32_t)fb;
    // the reference truncates the double product into `int num_rays`; we keep 64 bits so that grids beyond
    // 2^31 rays (which the reference cannot represent) remain addressable by the fused paths
    return (int64_t)(fa * fb);
}

extern "C" int64_t kt_imageplane_count(const kt_imageplane* p, int32_t* Nx, int32_t* Ny)
{
    if (!p || !axis_ok(p->x0, p->xmax, p->dx) || !axis_ok(p->y0, p->ymax, p->dy)) {
        if (Nx) *Nx = 0;
        if (Ny) *Ny = 0;
        fail(KT_ERR_INVALID, "image-plane grid: steps must be finite and non-zero and each axis must hold 1 .. 2^31-1 slots");
        return -1;
    }
    const double fx = ((p->xmax - p->x0) / p->dx) + 1;
    const double fy = ((p->ymax - p->y0) / p->dy) + 1;
    if (Nx) *Nx = (int32_t)fx;
    if (Ny) *Ny = (int32_t)fy;
    const int64_t slots = (int64_t)(fx * fy);
    return p->bundles ? slots * 5 : slots;
}

static int make_params(TraceParams& P, const kt_tracer* tr, const kt_stop* stop, const kt_run* run)
{
    if (!tr) return fail(KT_ERR_INVALID, "null tracer");
    memset(&P, 0, sizeof(P));
    P.tr = *tr;
    if (stop) P.stop = *stop;
    if (run) {
        if (run->integrator != KT_EULER && run->integrator != KT_RK4 && run->integrator != KT_RK45)
            return fail(KT_ERR_INVALID, "unknown integrator %d", run->integrator);
        // the reference asserts this (raytracer.cpp:983): Euler only exists for the theta_max overload
        if (run->integrator == KT_EULER && stop && stop->kind != KT_STOP_THETALIM)
            return fail(KT_ERR_INVALID, "Integrator::Euler does not support RayDestination stopping conditions");
        P.integrator = run->integrator;
        P.rlim = run->r_max;
        P.steplim = run->steplim > 0 ? run->steplim : (run->integrator == KT_RK45 ? KT_RK45_STEPLIM : KT_STEPLIM); // raytracer.cpp:80
    }
    if (stop && (stop->kind < KT_STOP_THETALIM || stop->kind > KT_STOP_FLATPLANE)) return fail(KT_ERR_INVALID, "unknown stop kind %d", stop->kind);
    P.stop._pad = 1 << P.stop.kind;    // one-hot copy of the kind for the device's bit tests (kt_device.cuh: dest_reached)
    const double a = tr->spin;
    if (run && (run->flags & KT_RUN_RETIRE_STUCK) && run->integrator == KT_RK45) {
        // innermost (prograde) circular photon orbit of |a|: no radial turning point exists below it
        const double r_ph = 2 * (1 + std::cos(2.0 / 3.0 * std::acos(-std::fabs(a) > -1 ? -std::fabs(a) : -1.0)));
        const double gap = r_ph - tr->horizon;
        P.stuck_eps = gap > 0 ? 1e-4 * gap : 0.0;
    }
    // reciprocals of the per-run constant divisors (kt_types.h): usable when the divisor is an ordinary number whose
    // significand is not all ones
    {
        auto usable = [](double b) {
            uint64_t u;
            memcpy(&u, &b, sizeof(u));
            const int e = (int)((u >> 52) & 0x7ff) - 1023;
            return std::isfinite(b) && b != 0 && e > -200 && e < 200 && (u & 0xfffffffffffffull) != 0xfffffffffffffull;
        };
        P.inv_precision = 1.0 / tr->precision;
        P.inv_theta_precision = 1.0 / tr->theta_precision;
        P.inv_six = 1.0 / 6.0;
        P.divc_ok = usable(tr->precision) && usable(tr->theta_precision);
        const double tl = P.stop.theta_lim;
        P.th_mode = tl > 0 ? 1 : (tl < 0 ? 2 : 0);
        P.tcap_on = tr->max_tstep > 0 && tr->maxtstep_rlim > 0;
        P.phcap_on = tr->max_phistep > 0;
        P.cap_num_t = tr->max_tstep > 0 ? tr->max_tstep : 1.0;
        P.cap_num_ph = tr->max_phistep > 0 ? tr->max_phistep : 1.0;
        // (+inf limits are fine: their bit patterns order above every finite number; NaNs are not)
        // ... and moderate magnitudes, the premises of the range-propagated acceptance (kt_trace.cuh: rhs_head, rk4_step_v3)
        auto within = [](double x, int lo, int hi) { return std::isfinite(x) && std::fabs(x) >= std::ldexp(1.0, lo) && std::fabs(x) <= std::ldexp(1.0, hi); };
        P.hot_ok = P.divc_ok && P.rlim > 0 && tr->horizon > 0 && within(tr->horizon, -20, 20) && !std::isnan(tl) &&
                   std::isfinite(tr->max_tstep) && std::isfinite(tr->max_phistep) && !std::isnan(tr->maxtstep_rlim) &&
                   (a == 0 || within(a, -20, 10)) && within(tr->precision, -30, 30) && within(tr->theta_precision, -30, 30) &&
                   within(P.cap_num_t, -100, 100) && within(P.cap_num_ph, -100, 100);
    }
    P.a = a;
    P.two_a = 2 * a;
    P.two_a_a = 2 * a * a;
    P.aa = a * a;
    P.nranks = 1;
    return KT_OK;
}

static int set_work(TraceParams& P, kt_ctx* c, long long count, const kt_shard* sh)
{
    P.queue = c->d_queue;
    P.abort_flag = reinterpret_cast<int*>(c->d_queue + 1);
    P.timeout_ns = c->timeout_s > 0 ? (unsigned long long)(c->timeout_s * 1e9) : 0ull;
    if (count < 0) return fail(KT_ERR_INVALID, "invalid source grid");   // kt_*_count has recorded why
    if (!sh) {
        P.begin = 0; P.end = count; P.rank = 0; P.nranks = 1; P.chunk = 65536; P.n_work = count;
        return KT_OK;
    }
    // a shard outside these bounds would make ordinal_to_index produce negative or foreign indices (duplicated / missing rays)
    const int nr = sh->nranks > 1 ? sh->nranks : 1;
    if (sh->begin < 0 || sh->end < sh->begin)
        return fail(KT_ERR_INVALID, "kt_shard: need 0 <= begin <= end (got begin %lld, end %lld)", (long long)sh->begin, (long long)sh->end);
    if (sh->nranks < 0 || (nr > 1 && (sh->rank < 0 || sh->rank >= nr)))
        return fail(KT_ERR_INVALID, "kt_shard: need 0 <= rank < nranks (got rank %d of %d)", (int)sh->rank, (int)sh->nranks);
    if (sh->chunk < 0) return fail(KT_ERR_INVALID, "kt_shard: negative chunk");
    long long b = sh->begin > count ? count : sh->begin;
    long long e = sh->end > count ? count : sh->end;
    P.begin = b; P.end = e;
    P.nranks = sh->nranks > 1 ? sh->nranks : 1;
    P.rank = P.nranks > 1 ? sh->rank : 0;
    const long long total = e - b;
    if (sh->chunk > 0) {
        P.chunk = sh->chunk;
    } else {
        // auto: ~256 chunks per rank so that every rank samples the whole cost range, 32 <= chunk <= 65536
        long long c = total / ((long long)P.nranks * 256);
        c = (c / 32) * 32;
        P.chunk = c < 32 ? 32 : (c > 65536 ? 65536 : c);
    }
    if (P.nranks == 1) { P.n_work = e - b; return KT_OK; }
    const long long nchunks = (total + P.chunk - 1) / P.chunk;
    const long long mine = nchunks > P.rank ? (nchunks - P.rank + P.nranks - 1) / P.nranks : 0;
    P.n_work = mine * P.chunk;   // upper bound; the kernel drops ordinals that map beyond `end`
    return KT_OK;
}

// source-frame tetrad of calculate_constants (raytracer.cpp:632-650), evaluated once on the host with the
// reference's expressions and the host libm (so these numbers are bit-identical to the CPU path)
static void make_ps_frame(PointSourceFrame& F, const kt_tracer* tr, const kt_pointsource* s)
{
    const double a = tr->spin, r = s->pos[1], th = s->pos[2], V = s->V;
    const double rho2 = r * r + (a * std::cos(th)) * (a * std::cos(th));
    const double del = r * r - 2 * r + a * a;
    const double sig2 = (r * r + a * a) * (r * r + a * a) - a * a * del * std::sin(th) * std::sin(th);
    const double e2nu = rho2 * del / sig2;
    const double e2psi = sig2 * std::sin(th) * std::sin(th) / rho2;
    const double omega = 2 * a * r / sig2;
    memset(&F, 0, sizeof(F));
    for (int i = 0; i < 4; i++) F.pos[i] = s->pos[i];
    F.E = s->E;
    F.V = V;
    F.et0 = (1 / std::sqrt(e2nu)) / std::sqrt(1 - (V - omega) * (V - omega) * e2psi / e2nu);
    F.et3 = (1 / std::sqrt(e2nu)) * V / std::sqrt(1 - (V - omega) * (V - omega) * e2psi / e2nu);
    F.e10 = (V - omega) * std::sqrt(e2psi / e2nu) / std::sqrt(e2nu - (V - omega) * (V - omega) * e2psi);
    F.e13 = (1 / std::sqrt(e2nu * e2psi)) * (e2nu + V * omega * e2psi - omega * omega * e2psi) /
            std::sqrt(e2nu - (V - omega) * (V - omega) * e2psi);
    F.e22 = -1 / std::sqrt(rho2);
    F.e31 = std::sqrt(del / rho2);
    F.rho2 = rho2;
    F.sn = std::sin(th);
    F.cs = std::cos(th);
    F.tn = std::tan(th);
    F.cosalpha0 = s->cosalpha0; F.dcosalpha = s->dcosalpha; F.cosalphamax = s->cosalphamax;
    F.beta0 = s->beta0; F.dbeta = s->dbeta; F.betamax = s->betamax;
    int32_t na, nb;
    kt_pointsource_count(s, &na, &nb);
    F.n_cosalpha = na;
    F.n_beta = nb;
}

static void make_plane_frame(PlaneFrame& F, const kt_imageplane* p)
{
    memset(&F, 0, sizeof(F));
    F.D = p->dist;
    F.incl = p->incl_deg * M_PI / 180;   // imageplane.cpp:23
    F.phi0 = p->phi0;
    F.x0 = p->x0; F.y0 = p->y0; F.dy = p->dy;
    F.eps_x = p->eps_frac * p->dx;       // imageplane_bundles.h:151-152
    F.eps_y = p->eps_frac * p->dy;
    F.sin_i = std::sin(F.incl);
    F.cos_i = std::cos(F.incl);
    int32_t nx, ny;
    kt_imageplane_count(p, &nx, &ny);
    F.Nx = nx; F.Ny = ny;
    F.bundles = p->bundles;
}

// ---------------------------------------------------------------------------------------------------------
// device-resident ray sets
// ---------------------------------------------------------------------------------------------------------
extern "C" kt_rayset* kt_rayset_create(kt_ctx* c, int64_t n)
{
    if (ctx_bind(c) != KT_OK) return nullptr;
    if (n < 0) { fail(KT_ERR_INVALID, "negative ray count"); return nullptr; }
    kt_rayset* rs = new (std::nothrow) kt_rayset;
    if (!rs) return nullptr;
    rs->ctx = c; rs->n = n; rs->d = nullptr;
    if (n > 0) {
        if (cudaMalloc(&rs->d, (size_t)n * sizeof(kt_ray)) != cudaSuccess) {
            cudaGetLastError();
            fail(KT_ERR_NOMEM, "device allocation of %lld rays failed", (long long)n);
            delete rs;
            return nullptr;
        }
        cudaError_t e = kt_kernel_table_exact()->rayset_reset(cfg_of(c), rs->d, n);
        c->launches++;
        if (e != cudaSuccess) {
            fail(KT_ERR_CUDA, "rayset reset: %s", cudaGetErrorString(e));
            cudaFree(rs->d);
            delete rs;
            return nullptr;
        }
    }
    return rs;
}

extern "C" void kt_rayset_destroy(kt_rayset* rs)
{
    if (!rs) return;
    cudaSetDevice(rs->ctx->device);
    cudaDeviceSynchronize();   // not the context's stream: it may be a caller-owned stream that no longer exists (see kt_destroy)
    if (rs->d) cudaFree(rs->d);
    delete rs;
}

extern "C" int64_t kt_rayset_count(const kt_rayset* rs) { return rs ? rs->n : 0; }
extern "C" void* kt_rayset_device_ptr(kt_rayset* rs) { return rs ? rs->d : nullptr; }

extern "C" int kt_rayset_upload(kt_rayset* rs, const kt_ray* host, int64_t n)
{
    if (!rs || !host || n < 0 || n > rs->n) return fail(KT_ERR_INVALID, "bad upload arguments");
    if (int rc = ctx_bind(rs->ctx)) return rc;
    KT_CUDA(cudaMemcpyAsync(rs->d, host, (size_t)n * sizeof(kt_ray), cudaMemcpyHostToDevice, rs->ctx->stream));
    KT_CUDA(cudaStreamSynchronize(rs->ctx->stream));
    return KT_OK;
}

extern "C" int kt_rayset_download(kt_rayset* rs, kt_ray* host, int64_t n)
{
    if (!rs || !host || n < 0 || n > rs->n) return fail(KT_ERR_INVALID, "bad download arguments");
    if (int rc = ctx_bind(rs->ctx)) return rc;
    KT_CUDA(cudaMemcpyAsync(host, rs->d, (size_t)n * sizeof(kt_ray), cudaMemcpyDeviceToHost, rs->ctx->stream));
    KT_CUDA(cudaStreamSynchronize(rs->ctx->stream));
    return KT_OK;
}

// ---- operations on a raw device pointer (shared by the rayset and the host-buffer entry points) ----------
static int dev_init_pointsource(kt_ctx* c, kt_ray* d, long long n, const kt_tracer* tr, const kt_pointsource* s)
{
    if (!s) return fail(KT_ERR_INVALID, "null source");
    TraceParams P;
    if (int rc = make_params(P, tr, nullptr, nullptr)) return rc;
    if (kt_pointsource_count(s, nullptr, nullptr) < 0) return KT_ERR_INVALID;   // degenerate grid (message recorded)
    PointSourceFrame F;
    make_ps_frame(F, tr, s);
    KT_CUDA(kt_kernel_table_exact()->init_pointsource(cfg_of(c), d, n, P, F));
    c->launches++;
    return KT_OK;
}

static int dev_init_imageplane(kt_ctx* c, kt_ray* d, long long n, const kt_tracer* tr, const kt_imageplane* p)
{
    if (!p) return fail(KT_ERR_INVALID, "null image plane");
    TraceParams P;
    if (int rc = make_params(P, tr, nullptr, nullptr)) return rc;
    if (kt_imageplane_count(p, nullptr, nullptr) < 0) return KT_ERR_INVALID;    // degenerate grid (message recorded)
    PlaneFrame F;
    make_plane_frame(F, p);
    KT_CUDA(kt_kernel_table_exact()->init_imageplane(cfg_of(c), d, n, P, F));
    c->launches++;
    return KT_OK;
}

static int dev_redshift_start(kt_ctx* c, kt_ray* d, long long n, const kt_tracer* tr, double V, int reverse, int projradius)
{
    TraceParams P;
    if (int rc = make_params(P, tr, nullptr, nullptr)) return rc;
    if (n <= 0) return KT_OK;
    if (V == -1) {
        // raytracer.cpp:391-394 overwrites the by-value V inside the loop: ray 0 defines it for every ray
        kt_ray first;
        KT_CUDA(cudaMemcpyAsync(&first, d, sizeof(kt_ray), cudaMemcpyDeviceToHost, c->stream));
        KT_CUDA(cudaStreamSynchronize(c->stream));
        const double a = reverse ? -1 * tr->spin : tr->spin;
        if (projradius) V = 1 / (a + first.r * std::sin(first.theta) * std::sqrt(first.r * std::sin(first.theta)));
        else V = 1 / (a + first.r * std::sqrt(first.r));
    }
    KT_CUDA(kt_kernel_table_exact()->redshift_start(cfg_of(c), d, n, P, V, reverse, projradius));
    c->launches++;
    return KT_OK;
}

static int dev_trace(kt_ctx* c, kt_ray* d, long long n, const kt_tracer* tr, const kt_stop* stop, const kt_run* run)
{
    if (!stop || !run) return fail(KT_ERR_INVALID, "null stop / run");
    TraceParams P;
    if (int rc = make_params(P, tr, stop, run)) return rc;
    if (n <= 0) return KT_OK;
    if (int rc = set_work(P, c, n, nullptr)) return rc;
    if (int rc = ensure_scratch(c, KT_NCOUNTERS * sizeof(double))) return rc;
    KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, KT_NCOUNTERS * sizeof(double), c->stream));
    KT_CUDA(table_for(run->math_mode)->trace_rayset(cfg_of(c), P, d, c->d_scratch));
    c->launches++;
    return KT_OK;
}

static int dev_redshift(kt_ctx* c, kt_ray* d, long long n, const kt_tracer* tr, double V, int reverse, int projradius, int motion)
{
    TraceParams P;
    if (int rc = make_params(P, tr, nullptr, nullptr)) return rc;
    if (n <= 0) return KT_OK;
    KT_CUDA(kt_kernel_table_exact()->redshift_V(cfg_of(c), d, n, P, V, reverse, projradius, motion));
    c->launches++;
    return KT_OK;
}

static int dev_redshift_dest(kt_ctx* c, kt_ray* d, long long n, const kt_tracer* tr, const kt_stop* stop, int reverse)
{
    TraceParams P;
    if (int rc = make_params(P, tr, stop, nullptr)) return rc;
    if (n <= 0) return KT_OK;
    KT_CUDA(kt_kernel_table_exact()->redshift_dest(cfg_of(c), d, n, P, reverse));
    c->launches++;
    return KT_OK;
}

static int dev_range_phi(kt_ctx* c, kt_ray* d, long long n, double lo, double hi)
{
    if (n <= 0) return KT_OK;
    KT_CUDA(kt_kernel_table_exact()->range_phi(cfg_of(c), d, n, lo, hi));
    c->launches++;
    return KT_OK;
}

#define RS_CHECK(rs)                                             \
    if (!(rs)) return fail(KT_ERR_INVALID, "null ray set");       \
    if (int rc__ = ctx_bind((rs)->ctx)) return rc__;

extern "C" int kt_rayset_init_pointsource(kt_rayset* rs, const kt_tracer* tr, const kt_pointsource* s)
{
    RS_CHECK(rs);
    return dev_init_pointsource(rs->ctx, rs->d, rs->n, tr, s);
}
extern "C" int kt_rayset_init_imageplane(kt_rayset* rs, const kt_tracer* tr, const kt_imageplane* p)
{
    RS_CHECK(rs);
    return dev_init_imageplane(rs->ctx, rs->d, rs->n, tr, p);
}
extern "C" int kt_rayset_redshift_start(kt_rayset* rs, const kt_tracer* tr, double V, int reverse, int projradius)
{
    RS_CHECK(rs);
    return dev_redshift_start(rs->ctx, rs->d, rs->n, tr, V, reverse, projradius);
}
extern "C" int kt_rayset_trace(kt_rayset* rs, const kt_tracer* tr, const kt_stop* stop, const kt_run* run)
{
    RS_CHECK(rs);
    return dev_trace(rs->ctx, rs->d, rs->n, tr, stop, run);
}
extern "C" int kt_rayset_redshift(kt_rayset* rs, const kt_tracer* tr, double V, int reverse, int projradius, int motion)
{
    RS_CHECK(rs);
    return dev_redshift(rs->ctx, rs->d, rs->n, tr, V, reverse, projradius, motion);
}
extern "C" int kt_rayset_redshift_dest(kt_rayset* rs, const kt_tracer* tr, const kt_stop* stop, int reverse)
{
    RS_CHECK(rs);
    return dev_redshift_dest(rs->ctx, rs->d, rs->n, tr, stop, reverse);
}
extern "C" int kt_rayset_range_phi(kt_rayset* rs, double lo, double hi)
{
    RS_CHECK(rs);
    return dev_range_phi(rs->ctx, rs->d, rs->n, lo, hi);
}
extern "C" int kt_rayset_calculate_momentum(kt_rayset* rs, const kt_tracer* tr)
{
    RS_CHECK(rs);
    TraceParams P;
    if (int rc = make_params(P, tr, nullptr, nullptr)) return rc;
    if (rs->n <= 0) return KT_OK;
    KT_CUDA(kt_kernel_table_exact()->calc_momentum(cfg_of(rs->ctx), rs->d, rs->n, P));
    rs->ctx->launches++;
    return KT_OK;
}

// ---- reductions: packed device buffer -> host structs ------------------------------------------------------
static void unpack_counters(const double* packed, int64_t* counters)
{
    for (int i = 0; i < KT_NCOUNTERS; i++) counters[i] = (int64_t)llround(packed[i]);
}

static int fetch_emis(kt_ctx* c, const double* d_hist, int Nr, kt_emis_out* out)
{
    const size_t nd = (size_t)5 * Nr + KT_NCOUNTERS;
    if (int rc = ensure_pinned(c, nd * sizeof(double))) return rc;
    KT_CUDA(cudaMemcpyAsync(c->h_pinned, d_hist, nd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    const double* h = c->h_pinned;
    for (int i = 0; i < Nr; i++) {
        if (out->N) out->N[i] = (int64_t)llround(h[i]);
        if (out->flux) out->flux[i] = h[Nr + i];
        if (out->emis) out->emis[i] = h[2 * Nr + i];
        if (out->redshift_sum) out->redshift_sum[i] = h[3 * Nr + i];
        if (out->time_sum) out->time_sum[i] = h[4 * Nr + i];
    }
    unpack_counters(h + 5 * (size_t)Nr, out->counters);
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

static int fetch_image(kt_ctx* c, const double* d_maps, long long npix, kt_image_out* out)
{
    const size_t nd = (size_t)7 * npix + KT_NCOUNTERS;
    if (int rc = ensure_pinned(c, nd * sizeof(double))) return rc;
    KT_CUDA(cudaMemcpyAsync(c->h_pinned, d_maps, nd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    const double* h = c->h_pinned;
    double* dst[6] = {out->flux, out->r, out->phi, out->enshift, out->time, out->emis};
    for (int m = 0; m < 6; m++)
        if (dst[m]) memcpy(dst[m], h + (size_t)m * npix, (size_t)npix * sizeof(double));
    if (out->N)
        for (long long i = 0; i < npix; i++) out->N[i] = (int32_t)llround(h[6 * (size_t)npix + i]);
    unpack_counters(h + 7 * (size_t)npix, out->counters);
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

extern "C" int kt_rayset_bin_emissivity(kt_rayset* rs, const kt_emis_bins* b, kt_emis_out* out)
{
    RS_CHECK(rs);
    if (!b || !out || b->Nr <= 0) return fail(KT_ERR_INVALID, "bad emissivity bins");
    kt_ctx* c = rs->ctx;
    const size_t nd = (size_t)5 * b->Nr + KT_NCOUNTERS;
    if (int rc = ensure_scratch(c, nd * sizeof(double))) return rc;
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, nd * sizeof(double), c->stream));
    if (rs->n > 0) {
        KT_CUDA(kt_kernel_table_exact()->bin_emissivity(cfg_of(c), rs->d, rs->n, *b, std::log(b->dr), c->d_scratch));
        c->launches++;
    }
    return fetch_emis(c, c->d_scratch, b->Nr, out);
}

extern "C" int kt_rayset_bin_image(kt_rayset* rs, const kt_image_bins* b, kt_image_out* out)
{
    RS_CHECK(rs);
    if (!b || !out || b->img_Nx <= 0 || b->img_Ny <= 0) return fail(KT_ERR_INVALID, "bad image bins");
    kt_ctx* c = rs->ctx;
    const long long npix = (long long)b->img_Nx * b->img_Ny;
    const size_t nd = (size_t)7 * npix + KT_NCOUNTERS;
    if (int rc = ensure_scratch(c, nd * sizeof(double))) return rc;
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, nd * sizeof(double), c->stream));
    if (rs->n > 0) {
        KT_CUDA(kt_kernel_table_exact()->bin_image(cfg_of(c), rs->d, rs->n, *b, c->d_scratch));
        c->launches++;
    }
    return fetch_image(c, c->d_scratch, npix, out);
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer entry points: H2D -> op -> D2H on the ctx stream (what the shim's Raytracer methods call)
// ---------------------------------------------------------------------------------------------------------
static int stage_in(kt_ctx* c, const kt_ray* rays, int64_t n, bool upload)
{
    if (!rays || n < 0) return fail(KT_ERR_INVALID, "bad ray buffer");
    if (int rc = ctx_bind(c)) return rc;
    if (n == 0) return KT_OK;
    if (int rc = ensure_rays(c, (size_t)n * sizeof(kt_ray))) return rc;
    if (upload) KT_CUDA(cudaMemcpyAsync(c->d_rays, rays, (size_t)n * sizeof(kt_ray), cudaMemcpyHostToDevice, c->stream));
    return KT_OK;
}

static int stage_out(kt_ctx* c, kt_ray* rays, int64_t n)
{
    if (n == 0) return KT_OK;
    KT_CUDA(cudaMemcpyAsync(rays, c->d_rays, (size_t)n * sizeof(kt_ray), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

// run_raytrace on a HOST ray array: H2D copy -> persistent trace kernel -> D2H copy.
// The array CAN be cut into up to NPIPE chunks, each on its own stream, so that copies overlap kernels
// (KT_PIPE_CHUNK=<rays>), but that is off by default: measured on B200 (1.6e7 rays, profiles/README.md) the copies are
// cheap (2.3 GB each way in 41 ms = 55 GB/s, 12 % of the call) while every chunk brings its own straggling rays
// (the kernel of a 4e6-ray chunk lives 360 ms for 130 ms of bulk work), and chunk kernels queued behind them
// lose more than the overlap wins: 1 chunk 711 ms, 2 chunks 765 ms, 4 chunks 886 ms.
extern "C" int kt_trace(kt_ctx* c, const kt_tracer* tr, const kt_stop* stop, const kt_run* run, kt_ray* rays, int64_t n)
{
    if (!stop || !run) return fail(KT_ERR_INVALID, "null stop / run");
    if (!rays || n < 0) return fail(KT_ERR_INVALID, "bad ray buffer");
    if (int rc = ctx_bind(c)) return rc;
    if (n == 0) return KT_OK;
    TraceParams P0;
    if (int rc = make_params(P0, tr, stop, run)) return rc;
    const KernelTable* T = table_for(run->math_mode);
    // Zero-copy: when the application's array is pinned host memory the device can address (kt_host_alloc, which is what
    // the shim's Raytracer::rays uses, or any cudaHostAlloc / cudaHostRegister'ed buffer), the persistent kernel reads each
    // ray's record over PCIe when a lane picks the ray up and writes the result record back when the ray retires: the
    // transfers hide completely under the integration (144 B in + 88 B out against ~1.4e5 flop per ray) and no device
    // staging copy exists.  KT_ZEROCOPY=0 forces the staged H2D -> kernel -> D2H path.
    if (c->zerocopy) {
        cudaPointerAttributes pa;
        memset(&pa, 0, sizeof(pa));
        // (the kernel moves records with 128-bit accesses: a host array that is not 16-byte aligned goes through the staged copy)
        if (cudaPointerGetAttributes(&pa, rays) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
            (reinterpret_cast<uintptr_t>(pa.devicePointer) & 15u) == 0) {
            TraceParams P = P0;
            if (int rc = set_work(P, c, n, nullptr)) return rc;
            if (int rc = ensure_scratch(c, KT_NCOUNTERS * sizeof(double))) return rc;
            KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
            KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, KT_NCOUNTERS * sizeof(double), c->stream));
            KT_CUDA(T->trace_rayset(cfg_of(c), P, static_cast<kt_ray*>(pa.devicePointer), c->d_scratch));
            c->launches++;
            KT_CUDA(cudaStreamSynchronize(c->stream));
            return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
        }
        cudaGetLastError();   // not a CUDA-visible pointer: staged path
    }
    if (int rc = stage_in(c, rays, n, false)) return rc;
    long long nchunks = n / c->pipe_chunk;
    nchunks = nchunks < 1 ? 1 : (nchunks > kt_ctx::NPIPE ? kt_ctx::NPIPE : nchunks);
    const bool trace_pipe = getenv("KT_PIPE_TRACE") != nullptr;   // developer aid: print the per-chunk timeline
    cudaEvent_t ev[kt_ctx::NPIPE][4] = {};
    const long long chunk = ((n + nchunks - 1) / nchunks + 31) / 32 * 32;
    int k = 0;
    // Everything that can fail after the first asynchronous operation runs inside `enqueue`: whatever it returns, the pipe
    // streams are drained before this call returns (no D2H copy may still be writing into the caller's array once the caller
    // has been told the call failed) and the timeline events are destroyed.
    auto enqueue = [&]() -> int {
        if (trace_pipe)
            for (int i = 0; i < kt_ctx::NPIPE; i++)
                for (int j = 0; j < 4; j++) KT_CUDA(cudaEventCreate(&ev[i][j]));
        KT_CUDA(cudaEventRecord(c->pipe_start, c->stream));
        for (int i = 0; i < kt_ctx::NPIPE; i++) KT_CUDA(cudaStreamWaitEvent(c->pipe[i], c->pipe_start, 0));
        for (long long off = 0; off < n; off += chunk, k++) {
            const long long len = (n - off < chunk) ? (n - off) : chunk;
            const int si = k % kt_ctx::NPIPE;
            cudaStream_t st = c->pipe[si];
            kt_ray* d = c->d_rays + off;
            if (trace_pipe) cudaEventRecord(ev[si][0], st);
            KT_CUDA(cudaMemcpyAsync(d, rays + off, (size_t)len * sizeof(kt_ray), cudaMemcpyHostToDevice, st));
            if (trace_pipe) cudaEventRecord(ev[si][1], st);
            TraceParams P = P0;
            if (int rc = set_work(P, c, len, nullptr)) return rc;
            P.queue = c->d_pipe_queue + si;
            // the stream's queue counter and its slice of the statistics are zeroed on the stream that uses them
            KT_CUDA(cudaMemsetAsync(P.queue, 0, sizeof(unsigned long long), st));
            KT_CUDA(cudaMemsetAsync(c->d_pipe_cnt + si * KT_NCOUNTERS, 0, KT_NCOUNTERS * sizeof(double), st));
            LaunchCfg cfg = cfg_of(c);
            cfg.stream = st;
            KT_CUDA(T->trace_rayset(cfg, P, d, c->d_pipe_cnt + si * KT_NCOUNTERS));
            c->launches++;
            if (trace_pipe) cudaEventRecord(ev[si][2], st);
            KT_CUDA(cudaMemcpyAsync(rays + off, d, (size_t)len * sizeof(kt_ray), cudaMemcpyDeviceToHost, st));
            if (trace_pipe) cudaEventRecord(ev[si][3], st);
        }
        for (int i = 0; i < kt_ctx::NPIPE; i++) {
            KT_CUDA(cudaEventRecord(c->pipe_done[i], c->pipe[i]));
            KT_CUDA(cudaStreamWaitEvent(c->stream, c->pipe_done[i], 0));
        }
        KT_CUDA(cudaStreamSynchronize(c->stream));
        return KT_OK;
    };
    const int rc = enqueue();
    if (rc != KT_OK) {
        for (int i = 0; i < kt_ctx::NPIPE; i++) cudaStreamSynchronize(c->pipe[i]);
        cudaGetLastError();
    } else if (trace_pipe) {
        for (int i = 0; i < k && i < kt_ctx::NPIPE; i++) {
            float t[4];
            for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], ev[0][0], ev[i][j]);
            fprintf(stderr, "kt_trace chunk %d: H2D %.1f-%.1f ms, kernel done %.1f ms, D2H done %.1f ms\n", i, t[0], t[1], t[2], t[3]);
        }
    }
    for (int i = 0; i < kt_ctx::NPIPE; i++)
        for (int j = 0; j < 4; j++)
            if (ev[i][j]) cudaEventDestroy(ev[i][j]);
    if (rc != KT_OK) return rc;
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}
// ---------------------------------------------------------------------------------------------------------
// ray-path dumping (run_raytrace with an outfile, raytracer.cpp:86-100 / 995-1008): counting pass, then writing pass
// ---------------------------------------------------------------------------------------------------------
static int check_path_args(const kt_stop* stop, const kt_run* run, const kt_ray* rays, int64_t n, const kt_path_opts* o)
{
    if (!stop || !run || !o) return fail(KT_ERR_INVALID, "null stop / run / path options");
    if (!rays || n < 0) return fail(KT_ERR_INVALID, "bad ray buffer");
    if (o->write_step <= 0) return fail(KT_ERR_INVALID, "write_step must be positive");
    return KT_OK;
}

struct DevBuf {   // scoped device allocation
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes)
    {
        if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(KT_ERR_NOMEM, "device allocation of %zu bytes failed", bytes);
        }
        return KT_OK;
    }
};

static int run_paths(kt_ctx* c, const kt_tracer* tr, const kt_stop* stop, const kt_run* run, const kt_ray* rays, int64_t n,
                     const kt_path_opts* o, long long* d_counts, const long long* d_offsets, double* d_points)
{
    TraceParams P;
    if (int rc = make_params(P, tr, stop, run)) return rc;
    if (int rc = ensure_rays(c, (size_t)n * sizeof(kt_ray))) return rc;
    if (int rc = ensure_scratch(c, KT_NCOUNTERS * sizeof(double))) return rc;
    KT_CUDA(cudaMemcpyAsync(c->d_rays, rays, (size_t)n * sizeof(kt_ray), cudaMemcpyHostToDevice, c->stream));
    if (int rc = set_work(P, c, n, nullptr)) return rc;
    KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, KT_NCOUNTERS * sizeof(double), c->stream));
    KT_CUDA(table_for(run->math_mode)->trace_paths(cfg_of(c), P, c->d_rays, c->d_scratch, *o, d_counts, d_offsets, d_points));
    c->launches++;
    return KT_OK;
}

extern "C" int kt_trace_paths_count(kt_ctx* c, const kt_tracer* tr, const kt_stop* stop, const kt_run* run, const kt_ray* rays,
                                    int64_t n, const kt_path_opts* o, int64_t* offsets)
{
    if (int rc = check_path_args(stop, run, rays, n, o)) return rc;
    if (!offsets) return fail(KT_ERR_INVALID, "null offsets");
    if (int rc = ctx_bind(c)) return rc;
    offsets[0] = 0;
    if (n == 0) return KT_OK;
    DevBuf counts;
    if (int rc = counts.alloc((size_t)n * sizeof(long long))) return rc;
    KT_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)n * sizeof(long long), c->stream));   // skipped rays write no points
    if (int rc = run_paths(c, tr, stop, run, rays, n, o, static_cast<long long*>(counts.p), nullptr, nullptr)) return rc;
    // counts land in offsets[1..n], then an in-place inclusive scan turns them into exclusive prefix sums
    KT_CUDA(cudaMemcpyAsync(offsets + 1, counts.p, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 1; i <= n; i++) offsets[i] += offsets[i - 1];
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

extern "C" int kt_trace_paths(kt_ctx* c, const kt_tracer* tr, const kt_stop* stop, const kt_run* run, kt_ray* rays, int64_t n,
                              const kt_path_opts* o, const int64_t* offsets, double* points)
{
    if (int rc = check_path_args(stop, run, rays, n, o)) return rc;
    if (!offsets) return fail(KT_ERR_INVALID, "null offsets");
    if (int rc = ctx_bind(c)) return rc;
    if (n == 0) return KT_OK;
    const long long total = offsets[n];
    if (total < 0 || (total > 0 && !points)) return fail(KT_ERR_INVALID, "bad offsets / null points");
    DevBuf d_off, d_pts;
    if (int rc = d_off.alloc((size_t)(n + 1) * sizeof(long long))) return rc;
    if (int rc = d_pts.alloc((size_t)total * 4 * sizeof(double))) return rc;
    KT_CUDA(cudaMemcpyAsync(d_off.p, offsets, (size_t)(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    if (int rc = run_paths(c, tr, stop, run, rays, n, o, nullptr, static_cast<const long long*>(d_off.p), static_cast<double*>(d_pts.p)))
        return rc;
    KT_CUDA(cudaMemcpyAsync(rays, c->d_rays, (size_t)n * sizeof(kt_ray), cudaMemcpyDeviceToHost, c->stream));
    if (total > 0) KT_CUDA(cudaMemcpyAsync(points, d_pts.p, (size_t)total * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

extern "C" int kt_redshift_start(kt_ctx* c, const kt_tracer* tr, kt_ray* rays, int64_t n, double V, int reverse, int projradius)
{
    if (int rc = stage_in(c, rays, n, true)) return rc;
    if (int rc = dev_redshift_start(c, c->d_rays, n, tr, V, reverse, projradius)) return rc;
    return stage_out(c, rays, n);
}
extern "C" int kt_redshift(kt_ctx* c, const kt_tracer* tr, kt_ray* rays, int64_t n, double V, int reverse, int projradius, int motion)
{
    if (int rc = stage_in(c, rays, n, true)) return rc;
    if (int rc = dev_redshift(c, c->d_rays, n, tr, V, reverse, projradius, motion)) return rc;
    return stage_out(c, rays, n);
}
extern "C" int kt_redshift_dest(kt_ctx* c, const kt_tracer* tr, const kt_stop* stop, kt_ray* rays, int64_t n, int reverse)
{
    if (int rc = stage_in(c, rays, n, true)) return rc;
    if (int rc = dev_redshift_dest(c, c->d_rays, n, tr, stop, reverse)) return rc;
    return stage_out(c, rays, n);
}
extern "C" int kt_range_phi(kt_ctx* c, kt_ray* rays, int64_t n, double lo, double hi)
{
    if (int rc = stage_in(c, rays, n, true)) return rc;
    if (int rc = dev_range_phi(c, c->d_rays, n, lo, hi)) return rc;
    return stage_out(c, rays, n);
}
extern "C" int kt_pointsource_init(kt_ctx* c, const kt_tracer* tr, const kt_pointsource* s, kt_ray* rays, int64_t n)
{
    if (int rc = stage_in(c, rays, n, false)) return rc;
    if (int rc = dev_init_pointsource(c, c->d_rays, n, tr, s)) return rc;
    return stage_out(c, rays, n);
}
extern "C" int kt_imageplane_init(kt_ctx* c, const kt_tracer* tr, const kt_imageplane* p, kt_ray* rays, int64_t n)
{
    if (int rc = stage_in(c, rays, n, false)) return rc;
    if (int rc = dev_init_imageplane(c, c->d_rays, n, tr, p)) return rc;
    return stage_out(c, rays, n);
}

// ---------------------------------------------------------------------------------------------------------
// fused paths
// ---------------------------------------------------------------------------------------------------------
extern "C" int kt_pointsource_emissivity_dev(kt_ctx* c, const kt_pointsource* s, const kt_tracer* tr, const kt_stop* stop,
                                             const kt_run* run, const kt_emis_bins* b, const kt_shard* sh, double* dev_hist)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!s || !stop || !run || !b || !dev_hist || b->Nr <= 0) return fail(KT_ERR_INVALID, "bad emissivity arguments");
    TraceParams P;
    if (int rc = make_params(P, tr, stop, run)) return rc;
    PointSourceFrame F;
    make_ps_frame(F, tr, s);
    const long long count = kt_pointsource_count(s, nullptr, nullptr);
    if (int rc = set_work(P, c, count, sh)) return rc;
    if (P.n_work <= 0) return KT_OK;
    KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
    KT_CUDA(table_for(run->math_mode)->trace_emis(cfg_of(c), P, F, *b, std::log(b->dr), dev_hist));
    c->launches++;
    return KT_OK;
}

extern "C" int kt_pointsource_emissivity(kt_ctx* c, const kt_pointsource* s, const kt_tracer* tr, const kt_stop* stop,
                                         const kt_run* run, const kt_emis_bins* b, const kt_shard* sh, kt_emis_out* out)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!b || !out || b->Nr <= 0) return fail(KT_ERR_INVALID, "bad emissivity arguments");
    const size_t nd = (size_t)5 * b->Nr + KT_NCOUNTERS;
    if (int rc = ensure_scratch(c, nd * sizeof(double))) return rc;
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, nd * sizeof(double), c->stream));
    if (int rc = kt_pointsource_emissivity_dev(c, s, tr, stop, run, b, sh, c->d_scratch)) return rc;
    return fetch_emis(c, c->d_scratch, b->Nr, out);
}

extern "C" int kt_imageplane_image_dev(kt_ctx* c, const kt_imageplane* p, const kt_tracer* tr, const kt_stop* stop,
                                       const kt_run* run, const kt_image_bins* b, const kt_shard* sh, double* dev_maps)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!p || !stop || !run || !b || !dev_maps || b->img_Nx <= 0 || b->img_Ny <= 0) return fail(KT_ERR_INVALID, "bad image arguments");
    TraceParams P;
    if (int rc = make_params(P, tr, stop, run)) return rc;
    PlaneFrame F;
    make_plane_frame(F, p);
    const long long count = kt_imageplane_count(p, nullptr, nullptr);
    if (int rc = set_work(P, c, count, sh)) return rc;
    if (P.n_work <= 0) return KT_OK;
    KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
    KT_CUDA(table_for(run->math_mode)->trace_image(cfg_of(c), P, F, *b, 1, dev_maps));
    c->launches++;
    return KT_OK;
}

extern "C" int kt_imageplane_image(kt_ctx* c, const kt_imageplane* p, const kt_tracer* tr, const kt_stop* stop, const kt_run* run,
                                   const kt_image_bins* b, const kt_shard* sh, kt_image_out* out)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!b || !out || b->img_Nx <= 0 || b->img_Ny <= 0) return fail(KT_ERR_INVALID, "bad image arguments");
    const long long npix = (long long)b->img_Nx * b->img_Ny;
    const size_t nd = (size_t)7 * npix + KT_NCOUNTERS;
    if (int rc = ensure_scratch(c, nd * sizeof(double))) return rc;
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, nd * sizeof(double), c->stream));
    if (int rc = kt_imageplane_image_dev(c, p, tr, stop, run, b, sh, c->d_scratch)) return rc;
    return fetch_image(c, c->d_scratch, npix, out);
}

// image plane (or bundles) -> trace -> redshift: compact per-ray results left in c->d_rays (count slots, the first
// Nx * Ny * (bundles ? 5 : 1) of them live)
static int dev_plane_rays(kt_ctx* c, const kt_imageplane* p, const kt_tracer* tr, const kt_stop* stop, const kt_run* run,
                          int redshift_mode, double V, int range_phi, PlaneFrame& F, long long count)
{
    TraceParams P;
    if (int rc = make_params(P, tr, stop, run)) return rc;
    make_plane_frame(F, p);
    if (int rc = set_work(P, c, count, nullptr)) return rc;
    kt_image_bins B;
    memset(&B, 0, sizeof(B));
    B.redshift_mode = redshift_mode;
    B.reverse = 1;
    B.V = V;
    const size_t out_bytes = (size_t)count * sizeof(kt_ray_min);
    if (int rc = ensure_rays(c, out_bytes + KT_NCOUNTERS * sizeof(double))) return rc;
    kt_ray_min* d_out = reinterpret_cast<kt_ray_min*>(c->d_rays);
    double* d_cnt = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(c->d_rays) + out_bytes);
    KT_CUDA(cudaMemsetAsync(d_out, 0, out_bytes + KT_NCOUNTERS * sizeof(double), c->stream));
    KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
    KT_CUDA(table_for(run->math_mode)->trace_planerays(cfg_of(c), P, F, B, range_phi, d_out, d_cnt));
    c->launches++;
    return KT_OK;
}

extern "C" int kt_caustic_discplane(kt_ctx* c, const kt_imageplane* p, const kt_tracer* tr, const kt_stop* stop, const kt_run* run,
                                    const kt_caustic_cfg* cfg, kt_caustic_out* out)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!p || !stop || !run || !cfg || !out) return fail(KT_ERR_INVALID, "bad arguments");
    if (cfg->img_Nx <= 0 || cfg->img_Ny <= 0) return fail(KT_ERR_INVALID, "bad map size %d x %d", cfg->img_Nx, cfg->img_Ny);
    const long long count = kt_imageplane_count(p, nullptr, nullptr);
    if (count < 0) return KT_ERR_INVALID;   // degenerate grid (message recorded by kt_imageplane_count)
    const long long npix = (long long)cfg->img_Nx * cfg->img_Ny;
    // one block of the context's scratch buffer (it only ever grows: no cudaMalloc / cudaFree per call): 8 double maps (the 7
    // outputs + the sign snapshot), 2 int maps, 7 counters
    const size_t bytes = (size_t)npix * (8 * sizeof(double) + 2 * sizeof(int)) + 8 * sizeof(unsigned long long);
    if (int rc = ensure_scratch(c, bytes)) return rc;
    PlaneFrame F;
    if (count > 0) {
        if (int rc = dev_plane_rays(c, p, tr, stop, run, cfg->redshift_mode, cfg->V, 0, F, count)) return rc;
    } else {
        make_plane_frame(F, p);
    }
    const long long live = (long long)F.Nx * F.Ny * (F.bundles ? 5 : 1);
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, bytes, c->stream));
    CausticMaps m;
    double* dp = c->d_scratch;
    m.det_J = dp; m.sign_J = dp + npix; m.radius = dp + 2 * npix; m.phi = dp + 3 * npix; m.x_disc = dp + 4 * npix;
    m.y_disc = dp + 5 * npix; m.redshift = dp + 6 * npix; m.sign_snap = dp + 7 * npix;
    m.cnt = reinterpret_cast<unsigned long long*>(dp + 8 * npix);
    m.order = reinterpret_cast<int*>(m.cnt + 8);
    m.hit = m.order + npix;
    const KernelTable* T = kt_kernel_table_exact();
    const kt_ray_min* d_rays = reinterpret_cast<const kt_ray_min*>(c->d_rays);
    for (int pass = 0; pass < (cfg->suppress ? 3 : 2); pass++) {
        KT_CUDA(T->caustic_pass(cfg_of(c), pass, d_rays, count > 0 ? live : 0, (int)F.Ny, F.bundles, F.eps_x, F.eps_y, *cfg, m));
        c->launches++;
    }
    struct { void* dst; const void* src; size_t n; } copies[] = {
        {out->det_J, m.det_J, sizeof(double)}, {out->sign_J, m.sign_J, sizeof(double)}, {out->radius, m.radius, sizeof(double)},
        {out->phi, m.phi, sizeof(double)}, {out->x_disc, m.x_disc, sizeof(double)}, {out->y_disc, m.y_disc, sizeof(double)},
        {out->redshift, m.redshift, sizeof(double)}, {out->order, m.order, sizeof(int)}, {out->hit, m.hit, sizeof(int)}};
    for (auto& cp : copies)
        if (cp.dst) KT_CUDA(cudaMemcpyAsync(cp.dst, cp.src, (size_t)npix * cp.n, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long cnt[8];
    KT_CUDA(cudaMemcpyAsync(cnt, m.cnt, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    double tc[KT_NCOUNTERS] = {0};
    if (count > 0)
        KT_CUDA(cudaMemcpyAsync(tc, reinterpret_cast<unsigned char*>(c->d_rays) + (size_t)count * sizeof(kt_ray_min), sizeof(tc),
                                cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    out->disc_count = (int64_t)cnt[0]; out->cnt_horizon = (int64_t)cnt[1]; out->cnt_rlim = (int64_t)cnt[2];
    out->cnt_steplim = (int64_t)cnt[3]; out->cnt_outofrange = (int64_t)cnt[4]; out->cnt_other = (int64_t)cnt[5];
    out->suppressed = (int64_t)cnt[6];
    unpack_counters(tc, out->counters);
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

extern "C" int kt_imageplane_rays(kt_ctx* c, const kt_imageplane* p, const kt_tracer* tr, const kt_stop* stop, const kt_run* run,
                                  int redshift_mode, double V, int range_phi, kt_ray_min* host_out, int64_t n)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!p || !stop || !run || !host_out) return fail(KT_ERR_INVALID, "bad arguments");
    TraceParams P;
    if (int rc = make_params(P, tr, stop, run)) return rc;
    PlaneFrame F;
    make_plane_frame(F, p);
    const long long count = kt_imageplane_count(p, nullptr, nullptr);
    if (count < 0) return KT_ERR_INVALID;   // degenerate grid (message recorded by kt_imageplane_count)
    if (n < count) return fail(KT_ERR_INVALID, "output buffer holds %lld rays, the plane has %lld", (long long)n, count);
    if (count == 0) return KT_OK;
    if (int rc = set_work(P, c, count, nullptr)) return rc;
    kt_image_bins B;
    memset(&B, 0, sizeof(B));
    B.redshift_mode = redshift_mode;
    B.reverse = 1;
    B.V = V;
    const size_t out_bytes = (size_t)count * sizeof(kt_ray_min);
    if (int rc = ensure_rays(c, out_bytes + KT_NCOUNTERS * sizeof(double))) return rc;
    kt_ray_min* d_out = reinterpret_cast<kt_ray_min*>(c->d_rays);
    double* d_cnt = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(c->d_rays) + out_bytes);
    // surplus tail slots (steps = -1 in the reference) are never written by the kernel: pre-fill them
    KT_CUDA(cudaMemsetAsync(d_out, 0, out_bytes + KT_NCOUNTERS * sizeof(double), c->stream));
    KT_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned long long), c->stream));
    KT_CUDA(table_for(run->math_mode)->trace_planerays(cfg_of(c), P, F, B, range_phi, d_out, d_cnt));
    c->launches++;
    KT_CUDA(cudaMemcpyAsync(host_out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    const long long live = (long long)F.Nx * F.Ny * (F.bundles ? 5 : 1);
    for (long long i = live; i < count; i++) host_out[i].steps = -1;
    return c->timeout_s > 0 ? kt_aborted(c) : KT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// FP64 roofline denominator
// ---------------------------------------------------------------------------------------------------------
extern "C" int kt_fp64_peak(kt_ctx* c, double* tflops)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!tflops) return fail(KT_ERR_INVALID, "null output");
    const int blocks = c->sm_count * 4, threads = 256, iters = 1024;
    struct Events {      // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
    KT_CUDA(cudaEventCreate(&ev.a));
    KT_CUDA(cudaEventCreate(&ev.b));
    const cudaEvent_t e0 = ev.a, e1 = ev.b;
    const KernelTable* T = kt_kernel_table_fast();
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        KT_CUDA(cudaEventRecord(e0, c->stream));
        KT_CUDA(T->dfma_peak(cfg_of(c), c->d_scratch, blocks, threads, iters));
        c->launches++;
        KT_CUDA(cudaEventRecord(e1, c->stream));
        KT_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        KT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fmas = (double)blocks * threads * (double)iters * 16.0 * 8.0;
        const double tf = 2.0 * fmas / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    return KT_OK;
}

// dependent-issue latency of one DFMA in SM cycles (a single warp running a serial chain): developer diagnostic
extern "C" int kt_fp64_latency(kt_ctx* c, double* cycles_per_op)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!cycles_per_op) return fail(KT_ERR_INVALID, "null output");
    KT_CUDA(kt_kernel_table_fast()->dfma_peak(cfg_of(c), c->d_scratch, -1, 32, 4096));
    c->launches++;
    double h[3];
    KT_CUDA(cudaMemcpyAsync(h, c->d_scratch, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    *cycles_per_op = h[0] / h[1];
    return KT_OK;
}

// Bit-exactness self-test of the deferred-guard division / square root used by the exact arithmetic mode.
// mode 0: operands uniform over all bit patterns; 1: moderate exponents; 2: with special values.
// counts[5] = {wrong quotients, wrong roots, quotients sent to the slow path, roots sent to the slow path, samples}
extern "C" int kt_selftest_divsqrt(kt_ctx* c, int mode, int samples_per_thread, int64_t* counts)
{
    if (int rc = ctx_bind(c)) return rc;
    if (!counts) return fail(KT_ERR_INVALID, "null output");
    KT_CUDA(cudaMemsetAsync(c->d_scratch, 0, 5 * sizeof(unsigned long long), c->stream));
    if (mode >= 100)   // developer self-test of the fast mode's lean right-hand side (kt_kernels.cu: k_selftest_lean); 101: negated spin
        KT_CUDA(kt_kernel_table_fast()->dfma_peak(cfg_of(c), c->d_scratch, -7, mode == 100 ? 1 : 0, samples_per_thread));
    else
        KT_CUDA(kt_kernel_table_exact()->dfma_peak(cfg_of(c), c->d_scratch, -2, mode, samples_per_thread));
    c->launches++;
    unsigned long long h[5];
    KT_CUDA(cudaMemcpyAsync(h, c->d_scratch, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    KT_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 5; i++) counts[i] = (int64_t)h[i];
    return KT_OK;
}

// sizeof() of every POD struct of the ABI, so bindings can check their mirrors (tests/test_abi.py)
extern "C" int kt_abi_sizeof(const char* name)
{
    if (!strcmp(name, "kt_ray")) return (int)sizeof(kt_ray);
    if (!strcmp(name, "kt_ray_min")) return (int)sizeof(kt_ray_min);
    if (!strcmp(name, "kt_stop")) return (int)sizeof(kt_stop);
    if (!strcmp(name, "kt_tracer")) return (int)sizeof(kt_tracer);
    if (!strcmp(name, "kt_run")) return (int)sizeof(kt_run);
    if (!strcmp(name, "kt_pointsource")) return (int)sizeof(kt_pointsource);
    if (!strcmp(name, "kt_imageplane")) return (int)sizeof(kt_imageplane);
    if (!strcmp(name, "kt_shard")) return (int)sizeof(kt_shard);
    if (!strcmp(name, "kt_emis_bins")) return (int)sizeof(kt_emis_bins);
    if (!strcmp(name, "kt_emis_out")) return (int)sizeof(kt_emis_out);
    if (!strcmp(name, "kt_image_bins")) return (int)sizeof(kt_image_bins);
    if (!strcmp(name, "kt_image_out")) return (int)sizeof(kt_image_out);
    if (!strcmp(name, "kt_path_opts")) return (int)sizeof(kt_path_opts);
    if (!strcmp(name, "kt_caustic_cfg")) return (int)sizeof(kt_caustic_cfg);
    if (!strcmp(name, "kt_caustic_out")) return (int)sizeof(kt_caustic_out);
    return -1;
}
