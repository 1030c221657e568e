// kt_kernels.cu -- IO policies, kernel instantiations and launchers.  Compiled once per arithmetic mode
// (Makefile: -DKT_FAST=0 -fmad=false  ->  kt_kernels_exact.o ;  -DKT_FAST=1 -fmad=true  ->  kt_kernels_fast.o).
#include <cstddef>
#include <type_traits>

#include "kt_kernels.h"
#include "kt_trace.cuh"

namespace kt {
namespace KT_MODE_NS {

// =========================================================================================================
// IO policy 1: device-resident AoS ray set in, same ray set out  (Raytracer<T>::run_raytrace on `rays`)
// =========================================================================================================
struct RaysetIO {
    // Retirement is a handful of stores, so it runs from the lane (no outbox); the inbox carries pr / ptheta only
    // because a theta turning point on a ray's very first step leaves them as they were (raytracer.cpp:824-829).  Small
    // staging matters in the exact mode: its spills live in L1, which is what shared memory leaves over.
    static constexpr int ND_IN = 10, NI_IN = 8, ND_OUT = 0, NI_OUT = 0;
    static constexpr bool DIRECT_RETIRE = true;
    static constexpr bool KEEP_MOM = true;     // Ray::pt..pphi are written back (raytracer.cpp:956-959)
    static constexpr bool PATHS = false;
    static constexpr bool NEED_PHI = true;
    static constexpr bool NEED_INFO = true;    // status bits, rdot_flips and equatorial_crossings are written back
    static constexpr bool NEED_FAILED = true;  // ... and so are t and phi of a ray that failed on the step limit
    static constexpr bool COOP_RECORDS = true; // a refill's 32 records are loaded by the warp as one coalesced span (stage())
    static constexpr int RECORD_STAGE_BYTES = 32 * 144;
    template <class H> __device__ void setup_paths(H&) {}
    __device__ long long path_first(long long) { return 0; }
    __device__ void path_done(long long, int) {}
    kt_ray* rays;
    double* counters;
    unsigned char* smem;
    const double2* staged;    // this lane's record in the warp's staging area (set by stage(); nullptr: read it from `rays`)

    __device__ void block_begin() {}
    __device__ void block_end() {}
    __device__ double* counters_out() { return counters; }

    // Coalesced load of initial conditions.  The 32 ordinals a warp draws for a refill are 32 CONSECUTIVE records, 4 608
    // contiguous bytes: the warp copies them into shared memory with nine fully coalesced 512-byte requests (lane l moves the
    // 16-byte words l, l + 32, ...), and each lane then picks its own record's words from there.  Compared with every lane
    // reading its own record (32 requests of 32 bytes per load instruction, 144 bytes apart) this is 8x fewer, 16x larger
    // memory requests -- which is what counts when `rays` is the application's pinned host array and every request is a PCIe
    // read.  (Sharded launches -- ordinals dealt in chunks -- and the tail of the array fall back to per-lane reads.)
    __device__ void stage(const TraceParams& P, long long base_ordinal)
    {
        staged = nullptr;
        if (P.nranks > 1) return;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const long long first = P.begin + base_ordinal;
        long long n = P.end - first;
        n = n < 0 ? 0 : (n > 32 ? 32 : n);
        double2* st = reinterpret_cast<double2*>(smem + warp * RECORD_STAGE_BYTES);
        const double2* g = reinterpret_cast<const double2*>(rays + first);
        const int words = (int)n * 9;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int w = lane + 32 * i;
            if (w < words) st[w] = g[w];
        }
        __syncwarp();
        staged = st + 9 * lane;
    }

    // Coalesced stores of results.  A finished lane deposits the words the trace owns -- w[0..3] = t .. pphi and the 24 bytes
    // steps | status, signs, flips | crossings -- with its record index in the warp's staging area (free between refills: 96 of a
    // slot's 144 bytes); flush_one() then writes one ray's record with seven lanes: four 16-byte words (one 64-byte request)
    // and three 8-byte words (one 24-byte request).  k, h, Q, emit, redshift, alpha, beta are not the trace's to write.
    __device__ void deposit(const RayOut& o, LaneCounters& c)
    {
        int st = o.steps0;
        if (o.steps > 0) st += o.steps;
        if (o.status & KT_STATUS_STEPLIM) st = -st;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double2* slot = reinterpret_cast<double2*>(smem + warp * RECORD_STAGE_BYTES + lane * 96);
        slot[0] = make_double2(o.t, o.r);
        slot[1] = make_double2(o.th, o.ph);
        slot[2] = make_double2(o.pt, o.pr);
        slot[3] = make_double2(o.pth, o.pph);
        slot[4] = make_double2(__hiloint2double(o.status, st), __hiloint2double(o.sth, o.sr));
        slot[5] = make_double2(__hiloint2double(o.cross, o.flips), __longlong_as_double(o.idx));
        count_terminal(c, o.steps, o.status);
    }
    __device__ void flush_one(int src)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const unsigned char* slot = smem + warp * RECORD_STAGE_BYTES + src * 96;
        const long long idx = *reinterpret_cast<const long long*>(slot + 88);
        unsigned char* rec = reinterpret_cast<unsigned char*>(rays + idx);
        if (lane < 4) reinterpret_cast<double2*>(rec)[lane] = reinterpret_cast<const double2*>(slot)[lane];
        else if (lane < 7) reinterpret_cast<double*>(rec + 104)[lane - 4] = reinterpret_cast<const double*>(slot + 64)[lane - 4];
    }

    // run_raytrace skips rays with steps < 0 or steps >= steplim: raytracer.cpp:116-117
    // The 144-byte record as nine 16-byte words: w[0] = {t, r}, w[1] = {theta, phi}, w[2] = {pt, pr}, w[3] = {ptheta, pphi},
    // w[4] = {k, h}, w[5] = {Q, emit}, w[6] = {redshift, steps | status}, w[7] = {rdot_sign | thetadot_sign, rdot_flips |
    // equatorial_crossings}, w[8] = {alpha, beta}.  Records are read and written with 128-bit accesses (LDG.128 / STG.128):
    // consecutive lanes of a refill take consecutive records, so a warp's loads cover one contiguous 4.6-KB span -- in HBM, or
    // over PCIe when `rays` is the application's pinned host array.
    __device__ bool make(const TraceParams& P, long long idx, RayIn& in, LaneCounters&)
    {
        static_assert(sizeof(kt_ray) == 144 && offsetof(kt_ray, steps) == 104 && offsetof(kt_ray, rdot_sign) == 112, "kt_ray layout");
        const double2* w = staged ? staged : reinterpret_cast<const double2*>(rays + idx);
        const double2 w6 = w[6];
        const int st = __double2loint(w6.y), status = __double2hiint(w6.y);
        if (st < 0 || st >= P.steplim) return false;
        const double2 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4], w5 = w[5], w7 = w[7];
        in.t = w0.x; in.r = w0.y; in.th = w1.x; in.ph = w1.y;
        in.pt = w2.x; in.pr = w2.y; in.pth = w3.x; in.pph = w3.y;
        in.k = w4.x; in.h = w4.y; in.Q = w5.x;
        in.emit = 0; in.step0 = 0;
        in.steps0 = st; in.status0 = status;
        in.sr = __double2loint(w7.x); in.sth = __double2hiint(w7.x); in.flips = __double2loint(w7.y); in.cross = __double2hiint(w7.y);
        return true;
    }
    __device__ void pack_in(Stage<ND_IN, NI_IN>& b, int slot, const RayIn& in)
    {
        b.d[0][slot] = in.t; b.d[1][slot] = in.r; b.d[2][slot] = in.th; b.d[3][slot] = in.ph;
        b.d[4][slot] = in.k; b.d[5][slot] = in.h; b.d[6][slot] = in.Q;
        b.d[7][slot] = in.pr; b.d[8][slot] = in.pth; b.d[9][slot] = in.step0;
        b.i[0][slot] = (int)(in.idx & 0xffffffffll); b.i[1][slot] = (int)(in.idx >> 32);
        b.i[2][slot] = in.steps0; b.i[3][slot] = in.status0; b.i[4][slot] = in.sr; b.i[5][slot] = in.sth;
        b.i[6][slot] = in.flips; b.i[7][slot] = in.cross;
    }
    __device__ void unpack_in(const Stage<ND_IN, NI_IN>& b, int slot, RayIn& in)
    {
        in.t = b.d[0][slot]; in.r = b.d[1][slot]; in.th = b.d[2][slot]; in.ph = b.d[3][slot];
        in.k = b.d[4][slot]; in.h = b.d[5][slot]; in.Q = b.d[6][slot];
        in.pt = 0; in.pr = b.d[7][slot]; in.pth = b.d[8][slot]; in.pph = 0;   // pt / pphi: rewritten by the first step
        in.step0 = b.d[9][slot];
        in.idx = ((long long)b.i[1][slot] << 32) | (unsigned int)b.i[0][slot];
        in.steps0 = b.i[2][slot]; in.status0 = b.i[3][slot]; in.sr = b.i[4][slot]; in.sth = b.i[5][slot];
        in.flips = b.i[6][slot]; in.cross = b.i[7][slot];
        in.emit = 0; in.alpha = 0; in.beta = 0;
    }
    __device__ void pack_out(Stage<ND_OUT, NI_OUT>&, int, const RayOut&) {}      // unused: DIRECT_RETIRE
    __device__ void unpack_out(const Stage<ND_OUT, NI_OUT>&, int, RayOut&) {}
    // write-back: raytracer.cpp:952-967
    __device__ void retire(const TraceParams&, const RayOut& o, LaneCounters& c)
    {
        int st = o.steps0;
        if (o.steps > 0) st += o.steps;
        if (o.status & KT_STATUS_STEPLIM) st = -st;
        double2* w = reinterpret_cast<double2*>(rays + o.idx);
        w[0] = make_double2(o.t, o.r);
        w[1] = make_double2(o.th, o.ph);
        w[2] = make_double2(o.pt, o.pr);
        w[3] = make_double2(o.pth, o.pph);
        // (k, h, Q, emit, redshift, alpha, beta are not the trace's to write: w[4], w[5], w[8] and the first half of w[6] stay)
        reinterpret_cast<double*>(w + 6)[1] = __hiloint2double(o.status, st);                       // steps | status
        w[7] = make_double2(__hiloint2double(o.sth, o.sr), __hiloint2double(o.cross, o.flips));     // signs, flips | crossings
        count_terminal(c, o.steps, o.status);
    }
};

// =========================================================================================================
// IO policy 1b: the same ray set, dumping ray paths  (run_raytrace with an outfile: raytracer.cpp:86-100)
//   points == nullptr: counting pass, counts[idx] = number of points ray idx writes
//   points != nullptr: writing pass, ray idx writes its points from offsets[idx] on
// =========================================================================================================
struct PathIO : RaysetIO {
    static constexpr bool PATHS = true;
    static constexpr bool NEED_PHI = true;
    kt_path_opts opts;
    long long* counts;
    const long long* offsets;
    double* points;
    template <class H> __device__ void setup_paths(H& h) { h.opts = opts; h.points = points; }
    __device__ long long path_first(long long idx) { return points ? offsets[idx] : 0; }
    __device__ void path_done(long long idx, int npts) { if (!points) counts[idx] = npts; }
};

// =========================================================================================================
// Shared pieces of the fused (source -> reduction) policies
// =========================================================================================================

// final steps value as the reference leaves it in rays[].steps (steps0 == 0 for freshly generated rays)
__device__ __forceinline__ int final_steps(int steps, int status)
{
    int st = 0;
    if (steps > 0) st = steps;
    if (status & KT_STATUS_STEPLIM) st = -st;
    return st;
}

// =========================================================================================================
// IO policy 2: PointSource -> emissivity histogram  (emissivity{,_rd}.cpp:88-126 in one kernel)
// =========================================================================================================
// PHI = false: the azimuth is not integrated (RK4).  Nothing in the emissivity reduction reads phi -- the histogram is over r -- so
// unless the stop surface does (FlatPlaneDestination: ray_destination.h:170-204), its accumulator and update are dropped: 5 FP64
// instructions and 4 registers per step.
template <bool PHI>
struct EmisIOT {
    static constexpr int ND_IN = 5, NI_IN = 3, ND_OUT = 7, NI_OUT = 3;
    static constexpr bool DIRECT_RETIRE = false;
    static constexpr bool KEEP_MOM = false;
    static constexpr bool PATHS = false;
    static constexpr bool NEED_PHI = PHI;
    // the histogram reads r, theta, t, the redshift and the terminal class of a ray: its ERGO / NEG_ENERGY bits, radial turning
    // points and equatorial crossings are unobservable (emissivity.cpp:96-126), so the RK4 step does not evaluate them
    static constexpr bool NEED_INFO = false;
    static constexpr bool NEED_FAILED = false;   // a ray with steps <= 0 is discarded before anything of it is read
    static constexpr bool COOP_RECORDS = false;
    __device__ void stage(const TraceParams&, long long) {}
    __device__ void deposit(const RayOut&, LaneCounters&) {}
    __device__ void flush_one(int) {}
    template <class H> __device__ void setup_paths(H&) {}
    __device__ long long path_first(long long) { return 0; }
    __device__ void path_done(long long, int) {}
    PointSourceFrame F;
    kt_emis_bins B;
    double log_dr;     // log(dr) evaluated by the host libm
    double* hist;      // global: double[5*Nr + KT_NCOUNTERS]
    int use_smem_hist;
    unsigned char* smem;

    __device__ double* block_hist() { return reinterpret_cast<double*>(smem); }
    __device__ void block_begin()
    {
        if (use_smem_hist) {
            double* hb = block_hist();
            for (int i = threadIdx.x; i < 5 * B.Nr; i += blockDim.x) hb[i] = 0.0;
        }
        __syncthreads();
    }
    // one atomic per non-empty bin per CTA
    __device__ void block_end()
    {
        __syncthreads();
        if (use_smem_hist) {
            const double* hb = block_hist();
            for (int i = threadIdx.x; i < 5 * B.Nr; i += blockDim.x) {
                const double v = hb[i];
                if (v != 0.0) atomicAdd(&hist[i], v);
            }
        }
    }
    __device__ double* counters_out() { return hist + 5 * (long long)B.Nr; }

    __device__ bool make(const TraceParams& P, long long idx, RayIn& in, LaneCounters&)
    {
        RaySeed sd;
        if (!pointsource_seed(F, P.a, idx, sd)) return false;
        in.t = sd.t; in.r = sd.r; in.th = sd.th; in.ph = sd.ph;
        in.k = sd.k; in.h = sd.h; in.Q = sd.Q;
        in.sr = sd.sr; in.sth = sd.sth;
        in.flips = 0; in.cross = 0; in.steps0 = 0; in.status0 = 0;
        in.pt = in.pr = in.pth = in.pph = 0; in.step0 = 0;
        // PointSource::redshift_start() -> Raytracer::redshift_start(velocity): pointsource.cpp:67-73
        in.emit = observed_energy_V(P, F.V, false, false, 0, sd.r, sd.th, sd.k, sd.h, sd.Q, sd.sr, sd.sth);
        return true;
    }
    __device__ void pack_in(Stage<ND_IN, NI_IN>& b, int slot, const RayIn& in)
    {
        b.d[0][slot] = in.k; b.d[1][slot] = in.h; b.d[2][slot] = in.Q; b.d[3][slot] = in.emit; b.d[4][slot] = in.step0;
        b.i[0][slot] = (int)(in.idx & 0xffffffffll); b.i[1][slot] = (int)(in.idx >> 32);
        b.i[2][slot] = (in.sr > 0 ? 1 : 0) | (in.sth > 0 ? 2 : 0);
    }
    __device__ void unpack_in(const Stage<ND_IN, NI_IN>& b, int slot, RayIn& in)
    {
        in.t = F.pos[0]; in.r = F.pos[1]; in.th = F.pos[2]; in.ph = F.pos[3];
        in.k = b.d[0][slot]; in.h = b.d[1][slot]; in.Q = b.d[2][slot]; in.emit = b.d[3][slot]; in.step0 = b.d[4][slot];
        in.idx = ((long long)b.i[1][slot] << 32) | (unsigned int)b.i[0][slot];
        const int sg = b.i[2][slot];
        in.sr = (sg & 1) ? 1 : -1; in.sth = (sg & 2) ? 1 : -1;
        in.flips = 0; in.cross = 0; in.steps0 = 0; in.status0 = 0;
        in.pt = in.pr = in.pth = in.pph = 0; in.alpha = in.beta = 0;
    }
    __device__ void pack_out(Stage<ND_OUT, NI_OUT>& b, int slot, const RayOut& o)
    {
        b.d[0][slot] = o.t; b.d[1][slot] = o.r; b.d[2][slot] = o.th;
        b.d[3][slot] = o.k; b.d[4][slot] = o.h; b.d[5][slot] = o.Q; b.d[6][slot] = o.emit;
        b.i[0][slot] = o.steps; b.i[1][slot] = o.status; b.i[2][slot] = (o.sr > 0 ? 1 : 0) | (o.sth > 0 ? 2 : 0);
    }
    __device__ void unpack_out(const Stage<ND_OUT, NI_OUT>& b, int slot, RayOut& o)
    {
        o.t = b.d[0][slot]; o.r = b.d[1][slot]; o.th = b.d[2][slot];
        o.k = b.d[3][slot]; o.h = b.d[4][slot]; o.Q = b.d[5][slot]; o.emit = b.d[6][slot];
        o.steps = b.i[0][slot]; o.status = b.i[1][slot];
        const int sg = b.i[2][slot];
        o.sr = (sg & 1) ? 1 : -1; o.sth = (sg & 2) ? 1 : -1;
    }
    // redshift + the app's binning loop: raytracer.cpp:480-600, emissivity.cpp:96-126, emissivity_rd.cpp:108-138
    __device__ void retire(const TraceParams& P, const RayOut& o, LaneCounters& c)
    {
        count_terminal(c, o.steps, o.status);
        const int st = final_steps(o.steps, o.status);
        if (!(st > 0)) return;
        double recv;
        if (B.redshift_mode == KT_REDSHIFT_DEST) recv = observed_energy_dest(P, false, o.r, o.th, o.k, o.h, o.Q, o.sr, o.sth);
        else recv = observed_energy_V(P, B.V, false, false, 0, o.r, o.th, o.k, o.h, o.Q, o.sr, o.sth);
        const double g = o.emit / recv;
        bool hit;
        if (B.hit_filter == KT_HIT_THETA) hit = o.th >= B.theta_lim - 1e-3;
        else hit = o.r * cos(o.th) < 1E-2;
        if (hit && g > 0 && o.r >= B.r_isco) {
            const int ir = B.logbin ? (int)(log(o.r / B.r_min) / log_dr) : (int)((o.r - B.r_min) / B.dr);
            if (ir >= 0 && ir < B.Nr) {
                double* hb = use_smem_hist ? block_hist() : hist;
                const int Nr = B.Nr;
                atomicAdd(&hb[ir], 1.0);
                atomicAdd(&hb[Nr + ir], 1 / (B.n_primary * pow(g, 1.0)));
                atomicAdd(&hb[2 * Nr + ir], 1 / pow(g, B.gamma));
                atomicAdd(&hb[3 * Nr + ir], g);
                atomicAdd(&hb[4 * Nr + ir], o.t);
            }
            c.disc++;
        }
    }
};
using EmisIO = EmisIOT<true>;

// =========================================================================================================
// IO policy 3: ImagePlane / ImagePlaneBundles -> image maps or compact per-ray results
//   IMAGE = true : imageplane_disc_image{,_rd,_isco}.cpp:109-161 (7 sum maps, global atomics: one ray per
//                  pixel when img_N == N, so contention is nil)
//   IMAGE = false: caustic_discplane.cpp:209-217 (trace + redshift; kt_ray_min per ray)
// =========================================================================================================
template <bool IMAGE>
struct PlaneIO {
    static constexpr int ND_IN = 7, NI_IN = 3, ND_OUT = 7, NI_OUT = 7;
    static constexpr bool DIRECT_RETIRE = false;
    static constexpr bool KEEP_MOM = false;
    static constexpr bool PATHS = false;
    static constexpr bool NEED_PHI = true;
    // image maps (imageplane_disc_image.cpp:122-161) read r, phi, t, the redshift and status & DEST; the compact per-ray
    // results of the caustic path carry status, rdot_flips and equatorial_crossings to the host
    static constexpr bool NEED_INFO = !IMAGE;
    static constexpr bool NEED_FAILED = !IMAGE;
    static constexpr bool COOP_RECORDS = false;
    __device__ void stage(const TraceParams&, long long) {}
    __device__ void deposit(const RayOut&, LaneCounters&) {}
    __device__ void flush_one(int) {}
    template <class H> __device__ void setup_paths(H&) {}
    __device__ long long path_first(long long) { return 0; }
    __device__ void path_done(long long, int) {}
    PlaneFrame F;
    kt_image_bins B;      // IMAGE only (redshift_mode / reverse / V also used by the ray variant)
    double* maps;         // IMAGE: double[7*npix + KT_NCOUNTERS]
    kt_ray_min* out;      // !IMAGE
    double* counters;     // !IMAGE
    int do_range_phi;
    unsigned char* smem;

    __device__ void block_begin() {}
    __device__ void block_end() {}
    __device__ double* counters_out() { return IMAGE ? maps + 7ll * B.img_Nx * B.img_Ny : counters; }

    __device__ bool make(const TraceParams& P, long long idx, RayIn& in, LaneCounters&)
    {
        double x, y;
        bool guard_b;
        if (!plane_slot(F, idx, x, y, guard_b)) return false;
        RaySeed sd;
        plane_seed(F, P.a, x, y, guard_b, sd, nullptr);
        in.t = 0; in.r = sd.r; in.th = sd.th; in.ph = sd.ph;
        in.k = sd.k; in.h = sd.h; in.Q = sd.Q;
        in.sr = sd.sr; in.sth = sd.sth;
        in.flips = 0; in.cross = 0; in.steps0 = 0; in.status0 = 0;
        in.pt = in.pr = in.pth = in.pph = 0; in.step0 = 0;
        // ImagePlane::redshift_start() -> Raytracer::redshift_start(0, true): imageplane.cpp:124-130
        in.emit = observed_energy_V(P, 0.0, true, false, 0, sd.r, sd.th, sd.k, sd.h, sd.Q, sd.sr, sd.sth);
        return true;
    }
    __device__ void pack_in(Stage<ND_IN, NI_IN>& b, int slot, const RayIn& in)
    {
        b.d[0][slot] = in.r; b.d[1][slot] = in.th; b.d[2][slot] = in.ph;
        b.d[3][slot] = in.h; b.d[4][slot] = in.Q; b.d[5][slot] = in.emit; b.d[6][slot] = in.step0;
        b.i[0][slot] = (int)(in.idx & 0xffffffffll); b.i[1][slot] = (int)(in.idx >> 32);
        b.i[2][slot] = (in.sr > 0 ? 1 : 0) | (in.sth > 0 ? 2 : 0);
    }
    __device__ void unpack_in(const Stage<ND_IN, NI_IN>& b, int slot, RayIn& in)
    {
        in.t = 0; in.r = b.d[0][slot]; in.th = b.d[1][slot]; in.ph = b.d[2][slot];
        in.k = 1; in.h = b.d[3][slot]; in.Q = b.d[4][slot]; in.emit = b.d[5][slot]; in.step0 = b.d[6][slot];
        in.idx = ((long long)b.i[1][slot] << 32) | (unsigned int)b.i[0][slot];
        const int sg = b.i[2][slot];
        in.sr = (sg & 1) ? 1 : -1; in.sth = (sg & 2) ? 1 : -1;
        in.flips = 0; in.cross = 0; in.steps0 = 0; in.status0 = 0;
        in.pt = in.pr = in.pth = in.pph = 0; in.alpha = in.beta = 0;
    }
    __device__ void pack_out(Stage<ND_OUT, NI_OUT>& b, int slot, const RayOut& o)
    {
        b.d[0][slot] = o.t; b.d[1][slot] = o.r; b.d[2][slot] = o.th; b.d[3][slot] = o.ph;
        b.d[4][slot] = o.h; b.d[5][slot] = o.Q; b.d[6][slot] = o.emit;
        b.i[0][slot] = (int)(o.idx & 0xffffffffll); b.i[1][slot] = (int)(o.idx >> 32);
        b.i[2][slot] = o.steps; b.i[3][slot] = o.status; b.i[4][slot] = (o.sr > 0 ? 1 : 0) | (o.sth > 0 ? 2 : 0);
        b.i[5][slot] = o.flips; b.i[6][slot] = o.cross;
    }
    __device__ void unpack_out(const Stage<ND_OUT, NI_OUT>& b, int slot, RayOut& o)
    {
        o.t = b.d[0][slot]; o.r = b.d[1][slot]; o.th = b.d[2][slot]; o.ph = b.d[3][slot];
        o.k = 1; o.h = b.d[4][slot]; o.Q = b.d[5][slot]; o.emit = b.d[6][slot];
        o.idx = ((long long)b.i[1][slot] << 32) | (unsigned int)b.i[0][slot];
        o.steps = b.i[2][slot]; o.status = b.i[3][slot];
        const int sg = b.i[4][slot];
        o.sr = (sg & 1) ? 1 : -1; o.sth = (sg & 2) ? 1 : -1;
        o.flips = b.i[5][slot]; o.cross = b.i[6][slot];
    }
    __device__ double powerlaw3(double r) const
    {
        // imageplane_disc_image.cpp:20-28
        if (r < B.rb1) return pow(r, -1 * B.q1);
        else if (r < B.rb2) return pow(B.rb1, B.q2 - B.q1) * pow(r, -1 * B.q2);
        else return pow(B.rb1, B.q2 - B.q1) * pow(B.rb2, B.q3 - B.q2) * pow(r, -1 * B.q3);
    }
    __device__ void retire(const TraceParams& P, const RayOut& o, LaneCounters& c)
    {
        count_terminal(c, o.steps, o.status);
        const int st = final_steps(o.steps, o.status);
        const bool rev = B.reverse != 0;
        double recv;
        if (B.redshift_mode == KT_REDSHIFT_DEST) recv = observed_energy_dest(P, rev, o.r, o.th, o.k, o.h, o.Q, o.sr, o.sth);
        else recv = observed_energy_V(P, B.V, rev, false, 0, o.r, o.th, o.k, o.h, o.Q, o.sr, o.sth);
        const double g = rev ? recv / o.emit : o.emit / recv;
        const double phi = do_range_phi ? wrap_phi(o.ph, st, -1 * KT_PI, KT_PI) : o.ph;
        if (!IMAGE) {
            kt_ray_min m;
            m.r = o.r; m.theta = o.th; m.phi = phi; m.t = o.t; m.redshift = g;
            m.steps = st; m.status = o.status; m.rdot_flips = o.flips; m.equatorial_crossings = o.cross;
            out[o.idx] = m;
            return;
        }
        if (!(st > 0)) return;
        bool hit;
        if (B.hit_filter == KT_HIT_DEST) hit = (o.status & KT_STATUS_DEST) && o.r < B.r_disc && g > 0;
        else if (B.hit_filter == KT_HIT_THETA) hit = o.th >= B.theta_lim - 1e-3 && o.r >= B.r_isco && o.r < B.r_disc && g > 0;
        else hit = o.r * cos(o.th) < 1E-2 && o.r >= B.r_isco && o.r < B.r_disc && g > 0;
        if (!hit) return;
        double x, y;
        bool guard_b;
        plane_slot(F, o.idx, x, y, guard_b);   // Ray::alpha / Ray::beta (imageplane.cpp:117-118)
        // bin-edge arithmetic exactly as imageplane_disc_image.cpp:134-135 (no FMA: plain sub and div)
        int ix = (int)(__ddiv_rn(__dsub_rn(x, B.x0), B.img_dx));
        int iy = (int)(__ddiv_rn(__dsub_rn(y, B.y0), B.img_dy));
        if (B.flip_image) iy = B.img_Ny - iy - 1;
        if (ix >= 0 && ix < B.img_Nx && iy >= 0 && iy < B.img_Ny) {
            const long long npix = (long long)B.img_Nx * B.img_Ny;
            const long long p = (long long)ix * B.img_Ny + iy;
            const double em = powerlaw3(o.r);
            atomicAdd(&maps[p], em / pow(g, 3.0));
            atomicAdd(&maps[npix + p], o.r);
            atomicAdd(&maps[2 * npix + p], phi);
            atomicAdd(&maps[3 * npix + p], 1. / g);
            atomicAdd(&maps[4 * npix + p], o.t);
            atomicAdd(&maps[5 * npix + p], em);
            atomicAdd(&maps[6 * npix + p], 1.0);
            c.disc++;
        }
    }
};

// =========================================================================================================
// One-thread-per-ray kernels on device-resident ray sets (the serial O(n) loops of the reference)
// =========================================================================================================
__global__ void k_rayset_reset(kt_ray* rays, long long n)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray z = {};
    z.steps = -1;
    rays[i] = z;
}

__global__ void k_init_pointsource(kt_ray* rays, long long n, TraceParams P, PointSourceFrame F)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray q = {};
    q.steps = -1;
    RaySeed sd;
    if (pointsource_seed(F, P.a, i, sd)) {
        q.t = sd.t; q.r = sd.r; q.theta = sd.th; q.phi = sd.ph;
        q.k = sd.k; q.h = sd.h; q.Q = sd.Q;
        q.rdot_sign = sd.sr; q.thetadot_sign = sd.sth;
        q.alpha = sd.alpha; q.beta = sd.beta;
        q.steps = 0;
    }
    rays[i] = q;
}

__global__ void k_init_imageplane(kt_ray* rays, long long n, TraceParams P, PlaneFrame F)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray q = {};
    q.steps = -1;
    double x, y;
    bool guard_b;
    if (plane_slot(F, i, x, y, guard_b)) {
        RaySeed sd;
        Mom p0;
        plane_seed(F, P.a, x, y, guard_b, sd, &p0);
        q.t = sd.t; q.r = sd.r; q.theta = sd.th; q.phi = sd.ph;
        q.pt = p0.pt; q.pr = p0.pr; q.ptheta = p0.pth; q.pphi = p0.pph;
        q.k = sd.k; q.h = sd.h; q.Q = sd.Q;
        q.rdot_sign = sd.sr; q.thetadot_sign = sd.sth;
        q.alpha = sd.alpha; q.beta = sd.beta;
        q.steps = 0;
    }
    rays[i] = q;
}

// redshift_start: raytracer.cpp:342-417 (V already resolved by the host if it was -1, see kt_api.cu)
__global__ void k_redshift_start(kt_ray* rays, long long n, TraceParams P, double V, int reverse, int projradius)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray& q = rays[i];
    q.emit = observed_energy_V(P, V, reverse != 0, projradius != 0, 0, q.r, q.theta, q.k, q.h, q.Q, q.rdot_sign,
                               q.thetadot_sign);
}

// redshift(V, ...): raytracer.cpp:420-447
__global__ void k_redshift_V(kt_ray* rays, long long n, TraceParams P, double V, int reverse, int projradius, int motion)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray& q = rays[i];
    const double recv = observed_energy_V(P, V, reverse != 0, projradius != 0, motion, q.r, q.theta, q.k, q.h, q.Q,
                                          q.rdot_sign, q.thetadot_sign);
    q.redshift = reverse ? recv / q.emit : q.emit / recv;
}

// redshift(dest, ...): raytracer.cpp:450-477
__global__ void k_redshift_dest(kt_ray* rays, long long n, TraceParams P, int reverse)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray& q = rays[i];
    const double recv = observed_energy_dest(P, reverse != 0, q.r, q.theta, q.k, q.h, q.Q, q.rdot_sign, q.thetadot_sign);
    q.redshift = reverse ? recv / q.emit : q.emit / recv;
}

__global__ void k_range_phi(kt_ray* rays, long long n, double lo, double hi)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    rays[i].phi = wrap_phi(rays[i].phi, rays[i].steps, lo, hi);
}

// calculate_momentum: raytracer.cpp:704-753
__global__ void k_calc_momentum(kt_ray* rays, long long n, TraceParams P)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    kt_ray& q = rays[i];
    const Mom m = momenta(P, q.k, q.h, q.Q, q.k * P.a, q.rdot_sign, q.thetadot_sign, q.r, q.theta);
    q.pt = m.pt; q.pr = m.pr; q.ptheta = m.pth; q.pphi = m.pph;
}

// the apps' binning loops over an already traced + redshifted ray set
__global__ void k_bin_emissivity(const kt_ray* rays, long long n, kt_emis_bins B, double log_dr, double* hist)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    LaneCounters c;
    c.clear();
    if (i < n) {
        const kt_ray q = rays[i];
        c.slots++;
        if (q.steps != -1 || q.status != 0) count_terminal(c, q.steps, q.status);
        if (q.steps > 0) {
            bool hit;
            if (B.hit_filter == KT_HIT_THETA) hit = q.theta >= B.theta_lim - 1e-3;
            else hit = q.r * cos(q.theta) < 1E-2;
            if (hit && q.redshift > 0 && q.r >= B.r_isco) {
                const int ir = B.logbin ? (int)(log(q.r / B.r_min) / log_dr) : (int)((q.r - B.r_min) / B.dr);
                if (ir >= 0 && ir < B.Nr) {
                    const int Nr = B.Nr;
                    atomicAdd(&hist[ir], 1.0);
                    atomicAdd(&hist[Nr + ir], 1 / (B.n_primary * pow(q.redshift, 1.0)));
                    atomicAdd(&hist[2 * Nr + ir], 1 / pow(q.redshift, B.gamma));
                    atomicAdd(&hist[3 * Nr + ir], q.redshift);
                    atomicAdd(&hist[4 * Nr + ir], q.t);
                }
                c.disc++;
            }
        }
    }
    flush_counters(c, hist + 5ll * B.Nr);
}

__global__ void k_bin_image(const kt_ray* rays, long long n, kt_image_bins B, double* maps)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    LaneCounters c;
    c.clear();
    const long long npix = (long long)B.img_Nx * B.img_Ny;
    if (i < n) {
        const kt_ray q = rays[i];
        c.slots++;
        if (q.steps != -1 || q.status != 0) count_terminal(c, q.steps, q.status);
        if (q.steps > 0) {
            const double g = q.redshift;
            bool hit;
            if (B.hit_filter == KT_HIT_DEST) hit = (q.status & KT_STATUS_DEST) && q.r < B.r_disc && g > 0;
            else if (B.hit_filter == KT_HIT_THETA) hit = q.theta >= B.theta_lim - 1e-3 && q.r >= B.r_isco && q.r < B.r_disc && g > 0;
            else hit = q.r * cos(q.theta) < 1E-2 && q.r >= B.r_isco && q.r < B.r_disc && g > 0;
            if (hit) {
                int ix = (int)(__ddiv_rn(__dsub_rn(q.alpha, B.x0), B.img_dx));
                int iy = (int)(__ddiv_rn(__dsub_rn(q.beta, B.y0), B.img_dy));
                if (B.flip_image) iy = B.img_Ny - iy - 1;
                if (ix >= 0 && ix < B.img_Nx && iy >= 0 && iy < B.img_Ny) {
                    const long long p = (long long)ix * B.img_Ny + iy;
                    double em;
                    if (q.r < B.rb1) em = pow(q.r, -1 * B.q1);
                    else if (q.r < B.rb2) em = pow(B.rb1, B.q2 - B.q1) * pow(q.r, -1 * B.q2);
                    else em = pow(B.rb1, B.q2 - B.q1) * pow(B.rb2, B.q3 - B.q2) * pow(q.r, -1 * B.q3);
                    atomicAdd(&maps[p], em / pow(g, 3.0));
                    atomicAdd(&maps[npix + p], q.r);
                    atomicAdd(&maps[2 * npix + p], q.phi);
                    atomicAdd(&maps[3 * npix + p], 1. / g);
                    atomicAdd(&maps[4 * npix + p], q.t);
                    atomicAdd(&maps[5 * npix + p], em);
                    atomicAdd(&maps[6 * npix + p], 1.0);
                    c.disc++;
                }
            }
        }
    }
    flush_counters(c, maps + 7 * npix);
}

// ---------------------------------------------------------------------------------------------------------
// caustic / critical-curve maps (caustic_discplane.cpp:146-493), one thread per pixel, on the compact per-ray results the
// plane kernel left in HBM.  Rays beyond n_live (surplus tail slots) read as unused (steps = -1).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ kt_ray_min caustic_ray(const kt_ray_min* rays, long long n_live, long long idx)
{
    if (idx < n_live) return rays[idx];
    kt_ray_min u = {};
    u.steps = -1;
    return u;
}
__device__ __forceinline__ bool caustic_valid_hit(const kt_ray_min& q, const kt_caustic_cfg& c)   // caustic_discplane.cpp:177-182
{
    return q.steps > 0 && q.r >= c.r_isco && q.r < c.r_disc && q.redshift > 0;
}
__device__ __forceinline__ void caustic_disc_xy(double phi_acc, double r, double& xd, double& yd)   // caustic_discplane.cpp:170-174
{
    const double phi_s = atan2(sin(phi_acc), cos(phi_acc));
    xd = r * cos(phi_s);
    yd = r * sin(phi_s);
}

// centre rays -> the seven per-pixel maps + the app's diagnostic counts: caustic_discplane.cpp:220-274 / 351-406
__global__ void k_caustic_centres(const kt_ray_min* rays, long long n_live, int plane_Ny, int per, kt_caustic_cfg c, CausticMaps m)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= (long long)c.img_Nx * c.img_Ny) return;
    const int ix = (int)(p / c.img_Ny), iy = (int)(p - (long long)ix * c.img_Ny);
    const kt_ray_min q = caustic_ray(rays, n_live, ((long long)ix * plane_Ny + iy) * per);
    if (caustic_valid_hit(q, c)) {
        const double phi_s = atan2(sin(q.phi), cos(q.phi));
        const int phi_ord = (int)(fabs(q.phi) / (2 * KT_PI)), r_ord = q.rdot_flips / 2;
        m.hit[p] = 1;
        m.radius[p] = q.r;
        m.phi[p] = phi_s;
        m.x_disc[p] = q.r * cos(phi_s);
        m.y_disc[p] = q.r * sin(phi_s);
        m.order[p] = phi_ord > r_ord ? phi_ord : r_ord;
        m.redshift[p] = q.redshift;
        atomicAdd(&m.cnt[0], 1ull);
    } else {
        m.hit[p] = 0;
        m.radius[p] = 0; m.phi[p] = 0; m.x_disc[p] = 0; m.y_disc[p] = 0; m.redshift[p] = 0;
        m.order[p] = -1;
    }
    if (q.steps > 0 && (q.r < c.r_isco || q.r >= c.r_disc || q.redshift <= 0)) atomicAdd(&m.cnt[4], 1ull);
    else if (q.steps <= 0 || !(q.status & KT_STATUS_DEST)) {
        if (q.status & KT_STATUS_HORIZON) atomicAdd(&m.cnt[1], 1ull);
        else if (q.status & KT_STATUS_RLIM) atomicAdd(&m.cnt[2], 1ull);
        else if (q.status & KT_STATUS_STEPLIM) atomicAdd(&m.cnt[3], 1ull);
        else atomicAdd(&m.cnt[5], 1ull);
    }
}

// Jacobian determinant of (x, y) -> (x_disc, y_disc): from the four satellite rays of the bundle
// (caustic_discplane.cpp:277-333) or from the four grid neighbours (caustic_discplane.cpp:409-447)
__global__ void k_caustic_jacobian(const kt_ray_min* rays, long long n_live, int plane_Ny, int bundles, double eps_x, double eps_y,
                                   kt_caustic_cfg c, CausticMaps m)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= (long long)c.img_Nx * c.img_Ny) return;
    const int Nx = c.img_Nx, Ny = c.img_Ny;
    const int ix = (int)(p / Ny), iy = (int)(p - (long long)ix * Ny);
    double det = __longlong_as_double(0x7ff8000000000000ll);   // quiet NaN: undefined
    double sgn = 0;
    bool defined = false;
    if (m.hit[p]) {
        if (bundles) {
            const long long base = ((long long)ix * plane_Ny + iy) * 5;
            const kt_ray_min rc = caustic_ray(rays, n_live, base), re = caustic_ray(rays, n_live, base + 1),
                             rw = caustic_ray(rays, n_live, base + 2), rn = caustic_ray(rays, n_live, base + 3),
                             rs = caustic_ray(rays, n_live, base + 4);
            if (caustic_valid_hit(re, c) && caustic_valid_hit(rw, c) && caustic_valid_hit(rn, c) && caustic_valid_hit(rs, c)) {
                const int flips = rc.rdot_flips;
                const double phi_c = rc.phi;
                const bool match = re.rdot_flips == flips && rw.rdot_flips == flips && rn.rdot_flips == flips &&
                                   rs.rdot_flips == flips && fabs(re.phi - phi_c) < KT_PI_2 && fabs(rw.phi - phi_c) < KT_PI_2 &&
                                   fabs(rn.phi - phi_c) < KT_PI_2 && fabs(rs.phi - phi_c) < KT_PI_2;
                if (!match) {
                    det = KT_CAUSTIC_SENTINEL;
                } else {
                    double xe, ye, xw, yw, xn, yn, xs, ys;
                    caustic_disc_xy(re.phi, re.r, xe, ye);
                    caustic_disc_xy(rw.phi, rw.r, xw, yw);
                    caustic_disc_xy(rn.phi, rn.r, xn, yn);
                    caustic_disc_xy(rs.phi, rs.r, xs, ys);
                    const double dxd_da = (xe - xw) / (2 * eps_x), dxd_db = (xn - xs) / (2 * eps_y);
                    const double dyd_da = (ye - yw) / (2 * eps_x), dyd_db = (yn - ys) / (2 * eps_y);
                    det = __dsub_rn(__dmul_rn(dxd_da, dyd_db), __dmul_rn(dxd_db, dyd_da));
                    defined = true;
                }
            }
        } else if (!(ix == 0 || ix == Nx - 1 || iy == 0 || iy == Ny - 1)) {
            const long long e = p + Ny, w = p - Ny, nn = p + 1, ss = p - 1;
            if (m.hit[e] && m.hit[w] && m.hit[nn] && m.hit[ss]) {
                const int ord = m.order[p];
                if (!(m.order[e] == ord && m.order[w] == ord && m.order[nn] == ord && m.order[ss] == ord)) {
                    det = KT_CAUSTIC_SENTINEL;
                } else {
                    const double dxd_dx = (m.x_disc[e] - m.x_disc[w]) / (2 * c.dx), dxd_dy = (m.x_disc[nn] - m.x_disc[ss]) / (2 * c.dy);
                    const double dyd_dx = (m.y_disc[e] - m.y_disc[w]) / (2 * c.dx), dyd_dy = (m.y_disc[nn] - m.y_disc[ss]) / (2 * c.dy);
                    det = __dsub_rn(__dmul_rn(dxd_dx, dyd_dy), __dmul_rn(dxd_dy, dyd_dx));
                    defined = true;
                }
            }
        }
    }
    if (defined) sgn = (det > 0) ? 1.0 : (det < 0) ? -1.0 : 0.0;
    m.det_J[p] = det;
    m.sign_J[p] = sgn;
    m.sign_snap[p] = sgn;
}

// suppression of alternating-sign pixels at geodesic branch boundaries, on the snapshot: caustic_discplane.cpp:467-492
__global__ void k_caustic_suppress(kt_caustic_cfg c, CausticMaps m)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= (long long)c.img_Nx * c.img_Ny) return;
    const int Nx = c.img_Nx, Ny = c.img_Ny;
    const int ix = (int)(p / Ny), iy = (int)(p - (long long)ix * Ny);
    const double s = m.sign_snap[p];
    if (s == 0.0) return;
    int n_same = 0, n_opp = 0;
    const int dix[4] = {-1, 1, 0, 0}, diy[4] = {0, 0, -1, 1};
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const int jx = ix + dix[d], jy = iy + diy[d];
        if (jx < 0 || jx >= Nx || jy < 0 || jy >= Ny) continue;
        const double sn = m.sign_snap[(long long)jx * Ny + jy];
        if (sn == 0.0) continue;
        if (sn * s > 0) ++n_same;
        else ++n_opp;
    }
    if (n_opp > n_same && n_opp >= 2) {
        m.det_J[p] = KT_CAUSTIC_SENTINEL;
        m.sign_J[p] = 0.0;
        atomicAdd(&m.cnt[6], 1ull);
    }
}

// FP64 FMA peak: 8 independent dependent-chains per thread, 4096 FMAs each
__global__ void k_dfma_peak(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;
}

// dependent-chain latency of DFMA (cycles per op for one warp alone): out[0] = cycles, out[1] = ops
__global__ void k_dfma_latency(double* out, int iters)
{
    double a = threadIdx.x * 1e-9;
    const double b = 1.0000001, c = 1e-7;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 32; u++) a = fma(a, b, c);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0); out[1] = (double)iters * 32.0; out[2] = a; }
}

#if KT_FAST
// developer self-test of the lean right-hand side against the reference expressions (momenta()) at random points:
// out[i] = max relative difference * 2^60 of {pt, pphi, |ptheta|, |pr|, sin^2, cos^2}
__global__ void k_selftest_lean(unsigned long long* out, unsigned long long seed, int per_thread, double a)
{
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    auto next = [&]() {
        x += 0x9E3779B97F4A7C15ull;
        unsigned long long z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto u01 = [&]() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); };
    TraceParams P;
    P.a = a; P.aa = a * a; P.two_a = 2 * a; P.two_a_a = 2 * a * a;
    unsigned long long worst[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < per_thread; i++) {
        RayState s;
        s.k = 1.0; s.h = 20.0 * (u01() - 0.5); s.Q = 300.0 * u01();
        s.ka = s.k * a; s.ka2 = s.ka * s.ka; s.h2 = s.h * s.h; s.ah = a * s.h; s.Kc = fma(s.h - s.ka, s.h - s.ka, s.Q);
        const double r = (i & 1) ? 1.2 + 30.0 * u01() : 30.0 + 500.0 * u01();
        const double th = 0.05 + 3.0 * u01();
        LeanVals v;
        lean_rhs<true>(P, s, r, th, v);
        // the reference's expressions (kerr.h:308-334) with the library's sincos, / and sqrt
        double sn, cs;
        sincos(th, &sn, &cs);
        Mom m;
        {
            const double sn2 = sn * sn, rr = r * r, ac = a * cs;
            const double rho2 = rr + ac * ac, del = rr - 2 * r + a * a, rho2del = rho2 * del;
            m.pt = ((rho2 * (rr + a * a) + 2 * a * a * r * sn2) * s.k - 2 * a * r * s.h) / rho2del;
            m.pph = (2 * a * r * sn2 * s.k + (rho2 - 2 * r) * s.h) / (sn2 * rho2del);
            const double thsq = (s.Q + (s.k * a * cs + s.h * cs / sn) * (s.k * a * cs - s.h * cs / sn)) / (rho2 * rho2);
            m.pth = sqrt(fabs(thsq));
            const double rsq = (s.k * m.pt - s.h * m.pph - rho2 * m.pth * m.pth) * del / rho2;
            m.pr = sqrt(fabs(rsq));
        }
        const double got[6] = {v.pt, v.pph, v.pth_abs, v.pr_abs, v.s2, 1.0 - v.s2};
        const double want[6] = {m.pt, m.pph, m.pth, m.pr, sn * sn, cs * cs};
        for (int j = 0; j < 6; j++) {
            const double e = fabs(got[j] - want[j]) / fmax(fabs(want[j]), 1e-300);
            // skip the neighbourhood of turning points, where both forms cancel: |p| tiny relative to its scale
            if ((j == 2 || j == 3) && fabs(want[j]) < 1e-3 * fabs(m.pt)) continue;
            const unsigned long long q = (unsigned long long)fmin(e * 1152921504606846976.0, 1.8e19);
            if (q > worst[j]) worst[j] = q;
        }
    }
    if (worst[5] > worst[4]) worst[4] = worst[5];
    for (int j = 0; j < 5; j++) atomicMax(&out[j], worst[j]);
}
#endif

// Self-test of xdiv / xsqrt (exact mode): random operands over the whole exponent range plus special values.
// out[0] = divisions where the guard accepted but the result differs from a / b   (must be 0)
// out[1] = square roots where the guard accepted but the result differs from sqrt(a) (must be 0)
// out[2] = divisions rejected by the guard, out[3] = square roots rejected, out[4] = samples
__global__ void k_selftest_divsqrt(unsigned long long* out, unsigned long long seed, int per_thread, int mode)
{
#if !KT_FAST
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    auto next = [&]() {
        x += 0x9E3779B97F4A7C15ull;
        unsigned long long z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    unsigned long long bad_d = 0, bad_s = 0, rej_d = 0, rej_s = 0;
    for (int i = 0; i < per_thread; i++) {
        unsigned long long ua = next(), ub = next();
        if (mode == 1) {            // moderate exponents (what the integrator sees): 2^-40 .. 2^40
            ua = (ua & 0x800fffffffffffffull) | ((unsigned long long)(1023 - 40 + (next() % 81)) << 52);
            ub = (ub & 0x800fffffffffffffull) | ((unsigned long long)(1023 - 40 + (next() % 81)) << 52);
        } else if (mode == 2 && (i & 7) == 0) {   // sprinkle special values
            const unsigned long long sp[8] = {0ull, 0x8000000000000000ull, 0x7ff0000000000000ull, 0xfff0000000000000ull,
                                              0x7ff8000000000000ull, 1ull, 0x000fffffffffffffull, 0x7fefffffffffffffull};
            if (i & 8) ua = sp[next() & 7]; else ub = sp[next() & 7];
        }
        if (mode == 4) {            // xdivc: constant divisors with host-style reciprocals, |a| over 2^-700 .. 2^700
            const double divisors[8] = {6.0, 100.0, 50.0, 3.0, 1000.0, 0.7, 1e-3, 12345.678};
            const double b = divisors[i & 7];
            const double y = 1.0 / b;
            const double a = __longlong_as_double((long long)((ua & 0x800fffffffffffffull) | ((unsigned long long)(1023 - 700 + (ub % 1401)) << 52)));
            bool ok = true;
            const double q = xdivc(a, b, y, ok);
            if (!ok) rej_d++;
            else if (__double_as_longlong(q) != __double_as_longlong(a / b)) bad_d++;
            continue;
        }
        if (mode == 6) {            // accuracy of the MUFU seeds: out[0] = max |1 - b rcp64h(b)| * 2^40, out[1] = max |1 - x rsq64h(x)^2| * 2^40
            const double b = __longlong_as_double((long long)((ua & 0x000fffffffffffffull) | (1023ull << 52)));          // [1, 2)
            const double x = __longlong_as_double((long long)((ub & 0x000fffffffffffffull) | ((1023ull + (i & 1)) << 52))); // [1, 4)
            double yb, yx;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yb) : "d"(b));
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yx) : "d"(x));
            const double eb = fabs(fma(-b, yb, 1.0)), ex = fabs(fma(-x * yx, yx, 1.0));
            const unsigned long long sb = (unsigned long long)(eb * 1099511627776.0), sx = (unsigned long long)(ex * 1099511627776.0);
            if (sb > bad_d) bad_d = sb;
            if (sx > bad_s) bad_s = sx;
            continue;
        }
        if (mode == 7) {            // kt_root5 against pow(x, 0.2) on [2^-40, 2^40]: out[0] = largest distance in ulps, out[1] = the step
                                    // factor's special cases that came out wrong (0, tiny, huge, NaN)
            const double x = __longlong_as_double((long long)((ua & 0x000fffffffffffffull) | ((unsigned long long)(1023 - 40 + (ub % 80)) << 52)));
            const long long d = __double_as_longlong(kt_root5(x)) - __double_as_longlong(pow(x, 0.2));
            const unsigned long long ad = (unsigned long long)(d < 0 ? -d : d);
            if (ad > bad_d) bad_d = ad;
            const double sp[6] = {0.0, 1e-300, 9.0e-13, 1e10, 1e300, __longlong_as_double(0x7ff8000000000000ll)};
            const double want[6] = {0.1, 0.1, 0.1, 5.0, 5.0, 5.0};
            if (rk45_step_factor(sp[i % 6]) != want[i % 6]) bad_s++;
            const double f = rk45_step_factor(x), g = std_max(0.1, std_min(5.0, 0.9 * pow(x, 0.2)));
            if (fabs(f - g) > 1e-15 * g) bad_s++;
            continue;
        }
        if (mode == 5) {            // ydiv / ysqrt / ydivc (range-propagated acceptance) over the whole domain their callers establish
            auto with_exp = [&](unsigned long long u, int lo, int hi) {
                const unsigned long long e = (unsigned long long)(1023 + lo) + next() % (unsigned long long)(hi - lo + 1);
                return __longlong_as_double((long long)((u & 0x800fffffffffffffull) | (e << 52)));
            };
            const double a = with_exp(ua, -840, 700), b = with_exp(ub, -333, 163);
            bool ok = true;
            const double q = ydiv(a, b, ok);
            if (!ok) rej_d++;
            else if (__double_as_longlong(q) != __double_as_longlong(a / b)) bad_d++;
            const double sa = fabs(with_exp(ua ^ ub, -500, 740));
            if (__double_as_longlong(ysqrt((i & 1) ? -sa : sa)) != __double_as_longlong(sqrt(sa))) bad_s++;   // ysqrt works on |x|
            const double divisors[8] = {6.0, 100.0, 50.0, 3.0, 1000.0, 0.7, 1e-3, 12345.678};
            const double c = divisors[i & 7];
            const double x = with_exp(ub ^ (ua >> 7), -530, 600);
            if (__double_as_longlong(ydivc(x, c, 1.0 / c)) != __double_as_longlong(x / c)) bad_s++;
            continue;
        }
        if (mode == 3) {            // kt_sincos_cuda vs sincos(): arguments a ray can have (|x| < 64) and a wide range (|x| < 2^30)
            const double u = (double)(ua >> 11) * (1.0 / 9007199254740992.0) - 0.5;      // [-0.5, 0.5)
            const double x = (i & 1) ? u * 128.0 : u * 2147483648.0;
            double s0, c0, s1, c1;
            sincos(x, &s0, &c0);
            kt_sincos_cuda(x, &s1, &c1);
            if (__double_as_longlong(s0) != __double_as_longlong(s1)) bad_d++;
            if (__double_as_longlong(c0) != __double_as_longlong(c1)) bad_s++;
            continue;
        }
        const double a = __longlong_as_double((long long)ua), b = __longlong_as_double((long long)ub);
        bool ok = true;
        const double q = xdiv(a, b, ok);
        const double qr = a / b;
        if (!ok) rej_d++;
        else if (__double_as_longlong(q) != __double_as_longlong(qr) && !(q != q && qr != qr)) bad_d++;
        bool oks = true;
        const double s = xsqrt(a, oks);
        const double sr = sqrt(a);
        if (!oks) rej_s++;
        else if (__double_as_longlong(s) != __double_as_longlong(sr) && !(s != s && sr != sr)) bad_s++;
    }
    if (mode == 6) { atomicMax(&out[0], bad_d); atomicMax(&out[1], bad_s); atomicAdd(&out[4], (unsigned long long)per_thread); return; }
    if (mode == 7) { atomicMax(&out[0], bad_d); atomicAdd(&out[1], bad_s); atomicAdd(&out[4], (unsigned long long)per_thread); return; }
    atomicAdd(&out[0], bad_d); atomicAdd(&out[1], bad_s); atomicAdd(&out[2], rej_d); atomicAdd(&out[3], rej_s);
    atomicAdd(&out[4], (unsigned long long)per_thread);
#endif
}

// =========================================================================================================
// Launchers
// =========================================================================================================
// threads per CTA of the persistent kernel.  Occupancy is set by registers per thread (65 536 per SM, allocated per warp in units
// of 8 per thread): with 128-thread CTAs the resident-warp count moves in steps of 4 (16, 20, 24 warps per SM at 128 / 96 / 80
// registers); developer builds with -DKT_TRACE_THREADS=32 -DKT_RK4_MINB=17..19 reach the steps in between.
#ifndef KT_TRACE_THREADS
#define KT_TRACE_THREADS 128
#endif
constexpr int TRACE_THREADS = KT_TRACE_THREADS;

template <int INTEG, class IO, int MINB>
static cudaError_t launch_trace(const LaunchCfg& cfg, const TraceParams& P, IO io, size_t extra_smem)
{
    auto kern = trace_persistent<INTEG, IO, TRACE_THREADS, MINB>;
    constexpr int WARPS = TRACE_THREADS / 32;
    const size_t smem = (sizeof(Stage<IO::ND_IN, IO::NI_IN>) + sizeof(Stage<IO::ND_OUT, IO::NI_OUT>) +
                         KT_NCOUNTERS * sizeof(unsigned long long)) * WARPS +
                        (IO::KEEP_MOM ? sizeof(double) * 4 * TRACE_THREADS : 0) + extra_smem;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cfg.carveout >= 0) {   // developer knob (KT_CARVEOUT, percent of the SM's shared-memory capacity); default: driver's choice
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cfg.carveout);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TRACE_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)cfg.sm_count * per_sm;
    // never launch more lanes than there is work
    const long long need = (P.n_work + TRACE_THREADS - 1) / TRACE_THREADS;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, TRACE_THREADS, smem, cfg.stream>>>(P, io);
    return cudaGetLastError();
}

template <class IO>
static cudaError_t launch_trace_rk4(const LaunchCfg& cfg, const TraceParams& P, IO io, size_t extra_smem)
{
#ifdef KT_DEV_MIN
    if constexpr (std::is_same<IO, EmisIOT<false>>::value) return launch_trace<KT_RK4, IO, KT_DEV_MIN>(cfg, P, io, extra_smem);
    else return cudaErrorNotSupported;
#elif defined(KT_RK4_MINB)
    return launch_trace<KT_RK4, IO, KT_RK4_MINB>(cfg, P, io, extra_smem);
#else
    const int minb = cfg.min_blocks;
    if (minb == 3) return launch_trace<KT_RK4, IO, 3>(cfg, P, io, extra_smem);
    if (minb == 4) return launch_trace<KT_RK4, IO, 4>(cfg, P, io, extra_smem);
    if (minb == 5) return launch_trace<KT_RK4, IO, 5>(cfg, P, io, extra_smem);
    if (minb == 6) return launch_trace<KT_RK4, IO, 6>(cfg, P, io, extra_smem);
    // both modes: 5 CTAs of 128 threads per SM (96 registers).  Round 2, exact mode, 1.5e7 rays: 4 / 5 / 6 -> 296 / 291 / 325 ms
    // (rk4_step_v3 needs ~90 registers: no local-memory traffic inside the loop at 96)
    return launch_trace<KT_RK4, IO, 5>(cfg, P, io, extra_smem);
#endif
}

template <class IO>
static cudaError_t launch_trace_any(const LaunchCfg& cfg, const TraceParams& P, IO io, size_t extra_smem)
{
#ifdef KT_DEV_MIN   // developer builds for SASS inspection: only the headline kernel (RK4, fused emissivity) is instantiated
    return cudaErrorNotSupported;
#else
    const int minb = cfg.min_blocks;
    if (P.integrator == KT_EULER) return launch_trace<KT_EULER, IO, 5>(cfg, P, io, extra_smem);
    // Defaults = measured best on B200 (1.5e7-ray lamppost runs through the fused kernel, profiles/README.md).  With the guard-free
    // step the kernel no longer pays for register spills with branch bubbles to hide: the spill-free occupancies win.
    //   RK4  exact: 3/4/5/6 CTAs of 128 threads per SM -> 358 / 346 / 385 / 385 ms ; fast: 4/5/6 -> 285 / 278 / 281 ms
    //   RK45 exact (2e6 rays): 3/4/5/6 -> 753 / 876 / 1092 / 1121 ms ; fast: 3/4/5 -> 509 / 521 / 658 ms
    if (P.integrator == KT_RK45) {
        if (minb == 4) return launch_trace<KT_RK45, IO, 4>(cfg, P, io, extra_smem);
        if (minb == 5) return launch_trace<KT_RK45, IO, 5>(cfg, P, io, extra_smem);
        return launch_trace<KT_RK45, IO, 3>(cfg, P, io, extra_smem);
    }
    return launch_trace_rk4(cfg, P, io, extra_smem);
#endif
}

static inline unsigned nblocks(long long n, int t) { return (unsigned)((n + t - 1) / t); }

static cudaError_t l_trace_rayset(const LaunchCfg& cfg, const TraceParams& P, kt_ray* rays, double* counters)
{
    RaysetIO io;
    io.rays = rays; io.counters = counters; io.smem = nullptr; io.staged = nullptr;
    return launch_trace_any(cfg, P, io, (size_t)RaysetIO::RECORD_STAGE_BYTES * (TRACE_THREADS / 32));
}

static cudaError_t l_trace_paths(const LaunchCfg& cfg, const TraceParams& P, kt_ray* rays, double* counters, const kt_path_opts& opts,
                                 long long* counts, const long long* offsets, double* points)
{
    PathIO io;
    io.rays = rays; io.counters = counters; io.smem = nullptr; io.staged = nullptr;
    io.opts = opts; io.counts = counts; io.offsets = offsets; io.points = points;
    // not a throughput path (ordered text output of a few thousand rays): one spill-free occupancy for all integrators
#ifdef KT_DEV_MIN
    return cudaErrorNotSupported;
#endif
    const size_t stage_bytes = (size_t)RaysetIO::RECORD_STAGE_BYTES * (TRACE_THREADS / 32);
    if (P.integrator == KT_EULER) return launch_trace<KT_EULER, PathIO, 3>(cfg, P, io, stage_bytes);
    if (P.integrator == KT_RK45) return launch_trace<KT_RK45, PathIO, 3>(cfg, P, io, stage_bytes);
    return launch_trace<KT_RK4, PathIO, 3>(cfg, P, io, stage_bytes);
}

static cudaError_t l_trace_emis(const LaunchCfg& cfg, const TraceParams& P, const PointSourceFrame& F, const kt_emis_bins& B,
                                double log_dr, double* hist)
{
    const size_t hist_bytes = (size_t)5 * B.Nr * sizeof(double);
    const bool smem_hist = hist_bytes <= 64 * 1024;
    // RK4 towards anything but a FlatPlaneDestination: nothing reads phi, so it is not integrated (EmisIOT<false>)
    if (P.integrator == KT_RK4 && P.stop.kind != KT_STOP_FLATPLANE) {
        EmisIOT<false> io;
        io.F = F; io.B = B; io.log_dr = log_dr; io.hist = hist; io.smem = nullptr;
        io.use_smem_hist = smem_hist;
        return launch_trace_rk4(cfg, P, io, smem_hist ? hist_bytes : 0);
    }
    EmisIO io;
    io.F = F; io.B = B; io.log_dr = log_dr; io.hist = hist; io.smem = nullptr;
    io.use_smem_hist = smem_hist;
    return launch_trace_any(cfg, P, io, smem_hist ? hist_bytes : 0);
}

static cudaError_t l_trace_image(const LaunchCfg& cfg, const TraceParams& P, const PlaneFrame& F, const kt_image_bins& B,
                                 int do_range_phi, double* maps)
{
    PlaneIO<true> io;
    io.F = F; io.B = B; io.maps = maps; io.out = nullptr; io.counters = nullptr; io.do_range_phi = do_range_phi; io.smem = nullptr;
    return launch_trace_any(cfg, P, io, 0);
}

static cudaError_t l_trace_planerays(const LaunchCfg& cfg, const TraceParams& P, const PlaneFrame& F, const kt_image_bins& B,
                                     int do_range_phi, kt_ray_min* out, double* counters)
{
    PlaneIO<false> io;
    io.F = F; io.B = B; io.maps = nullptr; io.out = out; io.counters = counters; io.do_range_phi = do_range_phi; io.smem = nullptr;
    return launch_trace_any(cfg, P, io, 0);
}

static cudaError_t l_rayset_reset(const LaunchCfg& cfg, kt_ray* rays, long long n)
{
    k_rayset_reset<<<nblocks(n, 256), 256, 0, cfg.stream>>>(rays, n);
    return cudaGetLastError();
}
static cudaError_t l_init_pointsource(const LaunchCfg& cfg, kt_ray* rays, long long n, const TraceParams& P, const PointSourceFrame& F)
{
    k_init_pointsource<<<nblocks(n, 128), 128, 0, cfg.stream>>>(rays, n, P, F);
    return cudaGetLastError();
}
static cudaError_t l_init_imageplane(const LaunchCfg& cfg, kt_ray* rays, long long n, const TraceParams& P, const PlaneFrame& F)
{
    k_init_imageplane<<<nblocks(n, 128), 128, 0, cfg.stream>>>(rays, n, P, F);
    return cudaGetLastError();
}
static cudaError_t l_redshift_start(const LaunchCfg& cfg, kt_ray* rays, long long n, const TraceParams& P, double V, int reverse, int projradius)
{
    k_redshift_start<<<nblocks(n, 128), 128, 0, cfg.stream>>>(rays, n, P, V, reverse, projradius);
    return cudaGetLastError();
}
static cudaError_t l_redshift_V(const LaunchCfg& cfg, kt_ray* rays, long long n, const TraceParams& P, double V, int reverse, int projradius, int motion)
{
    k_redshift_V<<<nblocks(n, 128), 128, 0, cfg.stream>>>(rays, n, P, V, reverse, projradius, motion);
    return cudaGetLastError();
}
static cudaError_t l_redshift_dest(const LaunchCfg& cfg, kt_ray* rays, long long n, const TraceParams& P, int reverse)
{
    k_redshift_dest<<<nblocks(n, 128), 128, 0, cfg.stream>>>(rays, n, P, reverse);
    return cudaGetLastError();
}
static cudaError_t l_range_phi(const LaunchCfg& cfg, kt_ray* rays, long long n, double lo, double hi)
{
    k_range_phi<<<nblocks(n, 256), 256, 0, cfg.stream>>>(rays, n, lo, hi);
    return cudaGetLastError();
}
static cudaError_t l_calc_momentum(const LaunchCfg& cfg, kt_ray* rays, long long n, const TraceParams& P)
{
    k_calc_momentum<<<nblocks(n, 128), 128, 0, cfg.stream>>>(rays, n, P);
    return cudaGetLastError();
}
static cudaError_t l_bin_emissivity(const LaunchCfg& cfg, const kt_ray* rays, long long n, const kt_emis_bins& B, double log_dr, double* hist)
{
    k_bin_emissivity<<<nblocks(n, 256), 256, 0, cfg.stream>>>(rays, n, B, log_dr, hist);
    return cudaGetLastError();
}
static cudaError_t l_bin_image(const LaunchCfg& cfg, const kt_ray* rays, long long n, const kt_image_bins& B, double* maps)
{
    k_bin_image<<<nblocks(n, 256), 256, 0, cfg.stream>>>(rays, n, B, maps);
    return cudaGetLastError();
}
static cudaError_t l_caustic_pass(const LaunchCfg& cfg, int pass, const kt_ray_min* rays, long long n_live, int plane_Ny, int bundles,
                                  double eps_x, double eps_y, const kt_caustic_cfg& c, const CausticMaps& m)
{
    const long long npix = (long long)c.img_Nx * c.img_Ny;
    if (pass == 0) k_caustic_centres<<<nblocks(npix, 256), 256, 0, cfg.stream>>>(rays, n_live, plane_Ny, bundles ? 5 : 1, c, m);
    else if (pass == 1) k_caustic_jacobian<<<nblocks(npix, 256), 256, 0, cfg.stream>>>(rays, n_live, plane_Ny, bundles, eps_x, eps_y, c, m);
    else k_caustic_suppress<<<nblocks(npix, 256), 256, 0, cfg.stream>>>(c, m);
    return cudaGetLastError();
}
static cudaError_t l_dfma_peak(const LaunchCfg& cfg, double* out, int blocks, int threads, int iters)
{
#if KT_FAST
    if (blocks == -7) { // developer self-test of the lean right-hand side (fast mode only)
        k_selftest_lean<<<cfg.sm_count * 4, 128, 0, cfg.stream>>>(reinterpret_cast<unsigned long long*>(out), 0x777ull, iters, threads > 0 ? 0.998 : -0.998);
        return cudaGetLastError();
    }
#endif
    if (blocks < -1) {  // self-test of xdiv / xsqrt: threads = sampling mode, iters = samples per thread
        k_selftest_divsqrt<<<cfg.sm_count * 8, 256, 0, cfg.stream>>>(reinterpret_cast<unsigned long long*>(out), 0x1234567ull + threads, iters, threads);
        return cudaGetLastError();
    }
    if (blocks < 0) {   // latency mode: one warp, dependent chain
        k_dfma_latency<<<1, 32, 0, cfg.stream>>>(out, iters);
        return cudaGetLastError();
    }
    k_dfma_peak<<<blocks, threads, 0, cfg.stream>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace KT_MODE_NS
}  // namespace kt

#if KT_FAST
#define KT_TABLE_NAME kt_kernel_table_fast
#else
#define KT_TABLE_NAME kt_kernel_table_exact
#endif

extern "C" const kt::KernelTable* KT_TABLE_NAME()
{
    static const kt::KernelTable table = {
        kt::KT_MODE_NS::l_trace_rayset, kt::KT_MODE_NS::l_trace_emis, kt::KT_MODE_NS::l_trace_image, kt::KT_MODE_NS::l_trace_planerays,
        kt::KT_MODE_NS::l_rayset_reset, kt::KT_MODE_NS::l_init_pointsource, kt::KT_MODE_NS::l_init_imageplane,
        kt::KT_MODE_NS::l_redshift_start, kt::KT_MODE_NS::l_redshift_V, kt::KT_MODE_NS::l_redshift_dest, kt::KT_MODE_NS::l_range_phi, kt::KT_MODE_NS::l_calc_momentum,
        kt::KT_MODE_NS::l_bin_emissivity, kt::KT_MODE_NS::l_bin_image, kt::KT_MODE_NS::l_dfma_peak,
        kt::KT_MODE_NS::l_caustic_pass, kt::KT_MODE_NS::l_trace_paths,
    };
    return &table;
}
