// kt_types.h -- POD parameter blocks shared by the host API (kt_api.cu) and the kernels (both arithmetic modes).
#pragma once
#include <cstdint>

#include "../../include/kerrtrace.h"

namespace kt {

// Per-launch constants (kernel argument -> constant bank).
struct TraceParams {
    kt_tracer tr;
    kt_stop stop;
    double rlim;
    int steplim;
    int integrator;
    // loop invariants pre-multiplied exactly as the reference's expressions associate: 2*a, (2*a)*a, a*a
    double a, two_a, two_a_a, aa;
    // x / precision, x / theta_precision and step / 6 divide by per-run constants: with y = RN(1/b) from the host's IEEE
    // division, q = RN(x*y), r = x - b*q (exact, one FMA), q' = RN(q + r*y) is the correctly rounded quotient (Markstein's
    // theorem; b's significand must not be all ones) -- 3 FP64 operations instead of 8 and no special-case frame.
    // divc_ok == 0 (odd divisors): the plain operator is used.  kt_device.cuh: xdivc
    double inv_precision, inv_theta_precision, inv_six;
    int divc_ok;
    // Host-evaluated facts about the run parameters for the exact-mode RK4 step (kt_trace.cuh: rk4_step_v3), whose ordering tests
    // run on bit patterns and therefore need to know the signs of the constants they compare with:
    //   hot_ok   : divc_ok, r_max > 0 and horizon > 0, all finite -- otherwise every step takes the plain out-of-line form
    //   th_mode  : the theta_max loop guard (raytracer.cpp:799): 0 none (theta_max == 0), 1 theta < theta_max, 2 theta > |theta_max|
    //   tcap_on  : max_tstep > 0 and maxtstep_rlim > 0 (raytracer.cpp:861);  phcap_on : max_phistep > 0 (:865)
    int hot_ok, th_mode, tcap_on, phcap_on;
    double cap_num_t, cap_num_ph;   // numerators of the two step caps: max_tstep / max_phistep, or 1 for a disabled cap (kept ordinary)
    double stuck_eps;   // KT_RUN_RETIRE_STUCK: retire ingoing RK45 rays with r - horizon < stuck_eps (0 = off)
    // work queue
    unsigned long long* queue;
    long long n_work;   // number of work ordinals this launch hands out
    // shard mapping: ordinal q -> ray index
    long long begin, end, chunk;
    int rank, nranks;
    // watchdog: when timeout_ns != 0 every warp polls %globaltimer and abandons the launch (setting *abort_flag)
    // once it has run longer than that -- a safety net so a bug can never wedge the GPU
    unsigned long long timeout_ns;
    int* abort_flag;
};

// Source-frame tetrad of a point source; depends only on the source position, so the HOST evaluates it once
// (kt_api.cu: make_ps_frame) with the reference's expressions (raytracer.cpp:632-650) and ships the numbers.
struct PointSourceFrame {
    double pos[4];
    double E;
    double et0, et3, e10, e13, e22, e31;
    double rho2, sn, cs, tn;  // rho^2, sin/cos/tan(theta) at the source
    double cosalpha0, dcosalpha, cosalphamax;
    double beta0, dbeta, betamax;
    double V;
    long long n_beta;
    long long n_cosalpha;
};

struct PlaneFrame {
    double D, incl, phi0;
    double x0, y0, dy;        // x uses dy as well (imageplane.cpp:43)
    double eps_x, eps_y;
    double sin_i, cos_i;      // host-evaluated sin/cos(incl)
    long long Nx, Ny;
    int bundles;
};

}  // namespace kt
