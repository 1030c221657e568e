/*
 * kerr_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see kerr_oracle.h for the rules and the parity pin).
 *
 * Restates, in plain C, what the reference computes on the hot path.  Citations are reference file:line.
 * Every floating-point expression keeps the reference's operand order and association so that a build
 * with -ffp-contract=off reproduces the reference bit for bit (checked by tests/test_oracle_vs_reference.py).
 * Comparisons are written in the reference's polarity so NaNs take the same branches.
 *
 * One deliberate deviation: the reference's RK45 retry loop never terminates when the error norm is NaN
 * (raytracer.cpp:1438-1541; SURVEY.md section 8c "oracle hazard").  Here such a ray is retired with
 * STEPLIM set and its step count negated, exactly as the CUDA path does, so neither can hang.
 */
#include "kerr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846
#define ORC_PI_2 1.57079632679489661923
#define ORC_DBL_MAX 1.7976931348623157e308

/* ------------------------------------------------------------------------------------------------ */
/* scalars: kerr.h:13-32                                                                              */
/* ------------------------------------------------------------------------------------------------ */
double orc_kerr_horizon(double a) { return 1 + 1 * sqrt((1 - a) * (1 + a)); }

double orc_kerr_isco(double a, int sign)
{
    /* kerr.h:29-31 keeps A and B in float and the return expression is evaluated in float */
    const float A = (float)(1. + pow(1. - a * a, 1. / 3.) * (pow(1. + a, 1. / 3.) + pow(1. - a, 1. / 3.)));
    const float AA = A * A;
    const float B = (float)sqrt(3. * a * a + (double)AA);
    const float lhs = 3 + B;
    const float p1 = 3 - A;
    const float twoB = 2 * B;
    const float p2 = (3 + A) + twoB;
    const float prod = p1 * p2;
    const float root = sqrtf(prod);
    const float signed_root = (float)sign * root;
    const float res = lhs - signed_root;
    return (double)res;
}

void orc_tracer_init(kt_tracer* tr, double spin, double precision)
{
    /* raytracer.cpp:13-37 */
    tr->spin = spin;
    tr->horizon = orc_kerr_horizon(spin);
    tr->precision = precision;
    tr->theta_precision = KT_THETA_PRECISION;
    tr->max_tstep = KT_MAXDT;
    tr->max_phistep = KT_MAXDPHI;
    tr->maxtstep_rlim = KT_MAXDT_RLIM;
    tr->rk45_tol = 1e-8;
}

int64_t orc_pointsource_count(const kt_pointsource* s, int32_t* n_cosalpha, int32_t* n_beta)
{
    /* pointsource.cpp:12,16-17: the product is formed in double and truncated once */
    const double fa = ((s->cosalphamax - s->cosalpha0) / s->dcosalpha) + 1;
    const double fb = ((s->betamax - s->beta0) / s->dbeta) + 1;
    if (n_cosalpha) *n_cosalpha = (int32_t)fa;
    if (n_beta) *n_beta = (int32_t)fb;
    return (int64_t)(int32_t)(fa * fb);
}

int64_t orc_imageplane_count(const kt_imageplane* p, int32_t* Nx, int32_t* Ny)
{
    /* imageplane.cpp:12-14; imageplane_bundles.h:146-150 */
    const double fx = ((p->xmax - p->x0) / p->dx) + 1;
    const double fy = ((p->ymax - p->y0) / p->dy) + 1;
    if (Nx) *Nx = (int32_t)fx;
    if (Ny) *Ny = (int32_t)fy;
    const int32_t slots = (int32_t)(fx * fy);
    return p->bundles ? (int64_t)(slots * 5) : (int64_t)slots;
}

void orc_rays_reset(kt_ray* rays, int64_t n)
{
    /* raytracer.cpp:43-49 leaves everything but steps/status uninitialised; we zero it */
    memset(rays, 0, (size_t)n * sizeof(kt_ray));
    for (int64_t i = 0; i < n; i++) rays[i].steps = -1;
}

/* ------------------------------------------------------------------------------------------------ */
/* photon momentum from the constants of motion: kerr.h:299-335                                       */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double pt, pr, pth, pph; } mom4;

static inline mom4 momenta(double k, double h, double Q, int sr, int sth, double r, double th, double a)
{
    mom4 m;
    const double sn = sin(th);
    const double cs = cos(th);
    const double sn2 = sn * sn;
    const double rho2 = r * r + (a * cs) * (a * cs);
    const double del = r * r - 2 * r + a * a;
    const double rho2del = rho2 * del;

    m.pt = (rho2 * (r * r + a * a) + 2 * a * a * r * sn2) * k - 2 * a * r * h;
    m.pt /= rho2del;

    m.pph = 2 * a * r * sn2 * k + (rho2 - 2 * r) * h;
    m.pph /= sn2 * rho2del;

    double thsq = Q + (k * a * cs + h * cs / sn) * (k * a * cs - h * cs / sn);
    thsq = thsq / (rho2 * rho2);
    m.pth = sqrt(fabs(thsq)) * sth;

    double rsq = k * m.pt - h * m.pph - rho2 * m.pth * m.pth;
    rsq = rsq * del / rho2;
    m.pr = sqrt(fabs(rsq)) * sr;
    return m;
}

void orc_momentum_from_consts(double* p4, double k, double h, double Q, int rdot_sign, int thetadot_sign, double r,
                              double theta, double a)
{
    mom4 m = momenta(k, h, Q, rdot_sign, thetadot_sign, r, theta, a);
    p4[0] = m.pt; p4[1] = m.pr; p4[2] = m.pth; p4[3] = m.pph;
}

/* ------------------------------------------------------------------------------------------------ */
/* sources                                                                                            */
/* ------------------------------------------------------------------------------------------------ */

/* Raytracer::calculate_constants, raytracer.cpp:625-676 */
static void constants_from_angles(kt_ray* q, double a, double alpha, double beta, double V, double E)
{
    const double r = q->r, th = q->theta;
    const double rho2 = r * r + (a * cos(th)) * (a * cos(th));
    const double del = r * r - 2 * r + a * a;
    const double sig2 = (r * r + a * a) * (r * r + a * a) - a * a * del * sin(th) * sin(th);

    const double e2nu = rho2 * del / sig2;
    const double e2psi = sig2 * sin(th) * sin(th) / rho2;
    const double omega = 2 * a * r / sig2;

    const double et0 = (1 / sqrt(e2nu)) / sqrt(1 - (V - omega) * (V - omega) * e2psi / e2nu);
    const double et3 = (1 / sqrt(e2nu)) * V / sqrt(1 - (V - omega) * (V - omega) * e2psi / e2nu);
    const double e10 = (V - omega) * sqrt(e2psi / e2nu) / sqrt(e2nu - (V - omega) * (V - omega) * e2psi);
    const double e13 = (1 / sqrt(e2nu * e2psi)) * (e2nu + V * omega * e2psi - omega * omega * e2psi) /
                       sqrt(e2nu - (V - omega) * (V - omega) * e2psi);
    const double e22 = -1 / sqrt(rho2);
    const double e31 = sqrt(del / rho2);

    const double n0 = E;
    const double n1 = E * sin(alpha) * cos(beta);
    const double n2 = E * sin(alpha) * sin(beta);
    const double n3 = E * cos(alpha);

    const double tdot = n0 * et0 + n1 * e10;
    const double phidot = n0 * et3 + n1 * e13;
    const double rdot = n3 * e31;
    const double thdot = n2 * e22;

    q->k = (1 - 2 * r / rho2) * tdot + (2 * a * r * sin(th) * sin(th) / rho2) * phidot;

    q->h = phidot * ((r * r + a * a) * (r * r + a * a * cos(th) * cos(th) - 2 * r) * sin(th) * sin(th) +
                     2 * a * a * r * sin(th) * sin(th) * sin(th) * sin(th));
    q->h = q->h - 2 * a * r * q->k * sin(th) * sin(th);
    q->h = q->h / (r * r + a * a * cos(th) * cos(th) - 2 * r);

    q->Q = rho2 * rho2 * thdot * thdot -
           (a * q->k * cos(th) + q->h / tan(th)) * (a * q->k * cos(th) - q->h / tan(th));

    q->rdot_sign = (rdot >= 0) ? 1 : -1;
    q->thetadot_sign = (thdot > 0) ? 1 : -1;
    q->rdot_flips = 0;
    q->equatorial_crossings = 0;
}

/* PointSource ctor + init_pointsource, pointsource.cpp:11-64 */
void orc_pointsource_init(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_pointsource* s)
{
    int32_t na, nb;
    orc_pointsource_count(s, &na, &nb);
    orc_rays_reset(rays, n);
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++) {
            const int64_t ix = (int64_t)i * nb + j;
            if (ix >= n) continue;
            kt_ray* q = &rays[ix];
            const double cosalpha = s->cosalpha0 + i * s->dcosalpha;
            const double beta = s->beta0 + j * s->dbeta;
            if (cosalpha >= s->cosalphamax || beta >= s->betamax) {
                q->steps = -1;
                continue;
            }
            const double alpha = acos(cosalpha);
            q->alpha = cosalpha; /* sic: pointsource.cpp:48 stores the cosine */
            q->beta = beta;
            q->t = s->pos[0];
            q->r = s->pos[1];
            q->theta = s->pos[2];
            q->phi = s->pos[3];
            q->pt = q->pr = q->ptheta = q->pphi = 0;
            q->steps = 0;
            constants_from_angles(q, tr->spin, alpha, beta, s->V, s->E);
        }
}

/* one image-plane ray: imageplane.cpp:45-118 and imageplane_bundles.h:60-128 (guard_b: the b>0 test of :113) */
static void plane_ray(kt_ray* q, double a, double D, double incl, double phi0, double x, double y, int guard_b)
{
    const double r = sqrt(D * D + x * x + y * y);
    const double th = acos((D * cos(incl) + y * sin(incl)) / r);
    const double ph = phi0 + atan2(x, D * sin(incl) - y * cos(incl));

    const double pr = D / r;
    const double pth = sin(acos(D / r)) / r;
    const double pph = x * sin(incl) / (x * x + (D * sin(incl) - y * cos(incl)) * (D * sin(incl) - y * cos(incl)));

    const double rho2 = r * r + (a * cos(th)) * (a * cos(th));
    const double del = r * r - 2 * r + a * a;
    const double sig2 = (r * r + a * a) * (r * r + a * a) - a * a * del * sin(th) * sin(th);
    const double e2nu = rho2 * del / sig2;
    const double e2psi = sig2 * sin(th) * sin(th) / rho2;
    const double omega = 2 * a * r / sig2;

    const double g00 = e2nu - omega * omega * e2psi;
    const double g03 = omega * e2psi;
    const double g11 = -rho2 / del;
    const double g22 = -rho2;
    const double g33 = -e2psi;

    const double A = g00;
    const double B = 2 * g03 * pph;
    const double C = g11 * pr * pr + g22 * pth * pth + g33 * pph * pph;
    double pt = (-B + sqrt(B * B - 4 * A * C)) / (2 * A);
    if (pt < 0) pt = (-B - sqrt(B * B - 4 * A * C)) / (2 * A);

    q->t = 0;
    q->r = r;
    q->theta = th;
    q->phi = ph;
    q->pt = pt;
    q->pr = pr;
    q->ptheta = pth;
    q->pphi = pph;
    /* calculate_constants_from_p (raytracer.cpp:679-701) only produces k,h,Q, all overwritten below */
    q->rdot_sign = -1;
    q->k = 1;

    const double b = sqrt(x * x + y * y);
    double bang = guard_b ? ((b > 0) ? asin(y / b) : 0) : asin(y / b);
    if (x < 0) bang = ORC_PI - bang;

    const double h = -1. * b * sin(incl) * cos(bang);
    const double lth = b * sin(bang);
    const double Q = (lth * lth) - (a * cos(th)) * (a * cos(th)) + ((h / tan(th))) * ((h / tan(th)));
    q->h = h;
    q->Q = Q;
    q->thetadot_sign = (lth >= 0) ? 1 : -1;
    q->steps = 0;
    q->alpha = x;
    q->beta = y;
}

void orc_imageplane_init(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_imageplane* p)
{
    int32_t Nx, Ny;
    orc_imageplane_count(p, &Nx, &Ny);
    orc_rays_reset(rays, n);
    const double a = tr->spin; /* already negated by the caller, imageplane.cpp:12 */
    const double incl = p->incl_deg * ORC_PI / 180;
    const double eps_x = p->eps_frac * p->dx, eps_y = p->eps_frac * p->dy;
    for (int i = 0; i < Nx; i++) {
        const double x = p->x0 + i * p->dy; /* sic: dy on both axes, imageplane.cpp:43 */
        for (int j = 0; j < Ny; j++) {
            const double y = p->y0 + j * p->dy;
            if (!p->bundles) {
                const int64_t ix = (int64_t)i * Ny + j;
                if (ix < n) plane_ray(&rays[ix], a, p->dist, incl, p->phi0, x, y, 0);
            } else {
                const int64_t base = ((int64_t)i * Ny + j) * 5;
                if (base + 4 >= n) continue;
                plane_ray(&rays[base + 0], a, p->dist, incl, p->phi0, x, y, 1);
                plane_ray(&rays[base + 1], a, p->dist, incl, p->phi0, x + eps_x, y, 1);
                plane_ray(&rays[base + 2], a, p->dist, incl, p->phi0, x - eps_x, y, 1);
                plane_ray(&rays[base + 3], a, p->dist, incl, p->phi0, x, y + eps_y, 1);
                plane_ray(&rays[base + 4], a, p->dist, incl, p->phi0, x, y - eps_y, 1);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* redshift: raytracer.cpp:342-600                                                                    */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double g00, g03, g11, g22, g33, e2nu, e2psi, omega; } metric_t;

static inline metric_t metric_at(double r, double th, double a)
{
    metric_t m;
    const double rho2 = r * r + (a * cos(th)) * (a * cos(th));
    const double del = r * r - 2 * r + a * a;
    const double sig2 = (r * r + a * a) * (r * r + a * a) - a * a * del * sin(th) * sin(th);
    m.e2nu = rho2 * del / sig2;
    m.e2psi = sig2 * sin(th) * sin(th) / rho2;
    m.omega = 2 * a * r / sig2;
    m.g00 = m.e2nu - m.omega * m.omega * m.e2psi;
    m.g03 = m.omega * m.e2psi;
    m.g11 = -rho2 / del;
    m.g22 = -rho2;
    m.g33 = -m.e2psi;
    return m;
}

/* sum_ij g[i][j] et[i] p[j] in the reference's i-major, j-minor order including the zero entries
 * (raytracer.cpp:412-415): adding +-0 products never changes a finite sum but NaN/inf propagate. */
static inline double contract(const metric_t* m, const double et[4], const double p[4])
{
    double g[16];
    for (int i = 0; i < 16; i++) g[i] = 0;
    g[0] = m->g00; g[3] = m->g03; g[12] = m->g03; g[5] = m->g11; g[10] = m->g22; g[15] = m->g33;
    double acc = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) acc += g[i * 4 + j] * et[i] * p[j];
    return acc;
}

void orc_redshift_start(kt_ray* rays, int64_t n, const kt_tracer* tr, double V, int reverse, int projradius)
{
    for (int64_t i = 0; i < n; i++) {
        kt_ray* q = &rays[i];
        const double a = reverse ? -1 * tr->spin : tr->spin;
        const metric_t m = metric_at(q->r, q->theta, a);
        /* by-value V is overwritten inside the loop (raytracer.cpp:391-394): only the first ray defines it */
        if (V == -1 && projradius)
            V = 1 / (a + q->r * sin(q->theta) * sqrt(q->r * sin(q->theta)));
        else if (V == -1)
            V = 1 / (a + q->r * sqrt(q->r));
        const double et[4] = {(1 / sqrt(m.e2nu)) / sqrt(1 - (V - m.omega) * (V - m.omega) * m.e2psi / m.e2nu), 0, 0,
                              (1 / sqrt(m.e2nu)) * V / sqrt(1 - (V - m.omega) * (V - m.omega) * m.e2psi / m.e2nu)};
        mom4 mm = momenta(q->k, q->h, q->Q, q->rdot_sign, q->thetadot_sign, q->r, q->theta, tr->spin);
        double p[4] = {mm.pt, mm.pr, mm.pth, mm.pph};
        if (reverse) { p[1] *= -1; p[2] *= -1; p[3] *= -1; }
        q->emit = contract(&m, et, p);
    }
}

/* ray_redshift(V, ...), raytracer.cpp:480-553 */
static double redshift_V(const kt_tracer* tr, double V, int reverse, int projradius, const kt_ray* q, int motion)
{
    const double r = q->r, th = q->theta, spin = tr->spin;
    const double a = reverse ? -1 * spin : spin;
    const metric_t m = metric_at(r, th, a);
    double et[4] = {0, 0, 0, 0};
    if (motion == 0) {
        if (V == -1 && projradius)
            V = 1 / (a + r * sin(th) * sqrt(r * sin(th)));
        else if (V == -1)
            V = 1 / (a + r * sqrt(r));
        et[0] = (1 / sqrt(m.e2nu)) / sqrt(1 - (V - m.omega) * (V - m.omega) * m.e2psi / m.e2nu);
        et[3] = (1 / sqrt(m.e2nu)) * V / sqrt(1 - (V - m.omega) * (V - m.omega) * m.e2psi / m.e2nu);
    } else if (motion == 1) {
        if (V < 0) V = fabs(V) * (r * r - 2 * r + spin + spin) / (r * r + spin * spin); /* sic: spin+spin */
        et[0] = 1. / sqrt(m.g00 + m.g11 * V * V);
        et[1] = V * et[0];
    }
    mom4 mm = momenta(q->k, q->h, q->Q, q->rdot_sign, q->thetadot_sign, r, th, spin);
    double p[4] = {mm.pt, mm.pr, mm.pth, mm.pph};
    if (reverse) { p[1] *= -1; p[2] *= -1; p[3] *= -1; }
    const double recv = contract(&m, et, p);
    return reverse ? recv / q->emit : q->emit / recv;
}

void orc_redshift(kt_ray* rays, int64_t n, const kt_tracer* tr, double V, int reverse, int projradius, int motion)
{
    for (int64_t i = 0; i < n; i++) rays[i].redshift = redshift_V(tr, V, reverse, projradius, &rays[i], motion);
}

/* RayDestination::four_velocity default (ray_destination.h:59-78) + ray_redshift(et, ...) (raytracer.cpp:556-600);
 * note: the stored (possibly negated) spin is used for both the metric and the Keplerian velocity */
void orc_redshift_dest(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_stop* stop, int reverse)
{
    (void)stop; /* none of the three concrete destinations overrides velocity(): V = -1 -> Keplerian */
    const double spin = tr->spin;
    for (int64_t i = 0; i < n; i++) {
        kt_ray* q = &rays[i];
        const double r = q->r, th = q->theta;
        double V = -1;
        const metric_t m = metric_at(r, th, spin);
        if (V == -1) V = 1 / (spin + r * sqrt(r));
        const double gam = 1 / sqrt(1 - (V - m.omega) * (V - m.omega) * m.e2psi / m.e2nu);
        const double et[4] = {gam / sqrt(m.e2nu), 0, 0, gam * V / sqrt(m.e2nu)};
        mom4 mm = momenta(q->k, q->h, q->Q, q->rdot_sign, q->thetadot_sign, r, th, spin);
        double p[4] = {mm.pt, mm.pr, mm.pth, mm.pph};
        if (reverse) { p[1] *= -1; p[2] *= -1; p[3] *= -1; }
        const double recv = contract(&m, et, p);
        q->redshift = reverse ? recv / q->emit : q->emit / recv;
    }
}

/* raytracer.cpp:603-622 */
void orc_range_phi(kt_ray* rays, int64_t n, double lo, double hi)
{
    for (int64_t i = 0; i < n; i++) {
        kt_ray* q = &rays[i];
        if (fabs(q->phi) > 1000 || q->phi != q->phi || !(q->steps > 0)) continue;
        while (q->phi >= hi) q->phi -= 2 * ORC_PI;
        while (q->phi < lo) q->phi += 2 * ORC_PI;
    }
}

/* raytracer.cpp:704-753 */
void orc_calculate_momentum(kt_ray* rays, int64_t n, const kt_tracer* tr)
{
    for (int64_t i = 0; i < n; i++) {
        kt_ray* q = &rays[i];
        mom4 m = momenta(q->k, q->h, q->Q, q->rdot_sign, q->thetadot_sign, q->r, q->theta, tr->spin);
        q->pt = m.pt; q->pr = m.pr; q->ptheta = m.pth; q->pphi = m.pph;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* stop surfaces: ray_destination.h:85-204                                                            */
/* ------------------------------------------------------------------------------------------------ */
static inline int theta_side(double theta_lim, double theta)
{
    if (theta_lim > 0) return theta >= theta_lim;
    if (theta_lim < 0) return theta <= -theta_lim;
    return 0;
}

static inline int dest_reached(const kt_stop* s, double r, double th, double ph, double th_prev)
{
    switch (s->kind) {
        case KT_STOP_FLATDISC: return theta_side(s->theta_lim, th);
        case KT_STOP_ISCODISC:
            if (r < s->r_isco) return 0;
            if (s->r_out > 0 && r > s->r_out) return 0;
            if (s->theta_lim > 0)
                return (th_prev < s->theta_lim && th >= s->theta_lim) || (th_prev > s->theta_lim && th <= s->theta_lim);
            if (s->theta_lim < 0) {
                const double tl = -s->theta_lim;
                return (th_prev > tl && th <= tl) || (th_prev < tl && th >= tl);
            }
            return 0;
        case KT_STOP_FLATPLANE:
            return r * (sin(th) * sin(s->incl) * cos(ph - s->phi0) + cos(th) * cos(s->incl)) <= -s->z_s;
        default: return 0;
    }
}

static inline double dest_step_limit(const kt_stop* s, double r, double th, double pth)
{
    if (s->kind == KT_STOP_FLATPLANE) return ORC_DBL_MAX;
    if (s->kind == KT_STOP_ISCODISC) {
        if (r < s->r_isco) return ORC_DBL_MAX;
        if (s->r_out > 0 && r > s->r_out) return ORC_DBL_MAX;
    }
    if (s->theta_lim > 0 && pth > 0 && th < s->theta_lim) return (s->theta_lim - th) / pth;
    if (s->theta_lim < 0 && pth < 0 && th > -s->theta_lim) return (-s->theta_lim - th) / pth;
    return ORC_DBL_MAX;
}

/* ------------------------------------------------------------------------------------------------ */
/* propagators                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* loop guard of the theta_max overloads: raytracer.cpp:799, 1362-1364 */
static inline int theta_inside(double thetalim, double th)
{
    return (thetalim > 0 && th < thetalim) || (thetalim < 0 && th > fabs(thetalim)) || thetalim == 0;
}

/* Ray-path dumping, the `outfile` blocks of the propagators (raytracer.cpp:293-312, 923-942, 1209-1228, 1552-1571,
 * 1850-1869): every write_step-th step inside the radial window writes one point; a ray that leaves the window after
 * its first point stops.  Points are (t, x, y, z) through cartesian() (kerr.h:40-56) or (t, r, theta, phi).
 * sink == NULL: no dumping.  pts == NULL: count only. */
typedef struct path_sink {
    const kt_path_opts* o;
    double* pts;     /* 4 doubles per point, or NULL */
    int64_t n;       /* points written so far */
    int started;
} path_sink;

static inline int path_point(path_sink* k, int steps, double t, double r, double th, double ph, double a)
{
    if (!k || (steps % k->o->write_step) != 0) return 0;
    if ((k->o->write_rmax < 0 || r < k->o->write_rmax) && (k->o->write_rmin < 0 || r > k->o->write_rmin)) {
        k->started = 1;
        if (k->pts) {
            double* p = k->pts + 4 * k->n;
            p[0] = t;
            if (k->o->write_cartesian) {
                p[1] = sqrt(r * r + a * a) * sin(th) * cos(ph);
                p[2] = sqrt(r * r + a * a) * sin(th) * sin(ph);
                p[3] = r * cos(th);
            } else {
                p[1] = r; p[2] = th; p[3] = ph;
            }
        }
        k->n++;
        return 0;
    }
    return k->started;
}

/* Euler: Raytracer::propagate, raytracer.cpp:129-340 (theta_max overload only: raytracer.cpp:983 asserts that Euler is
 * never combined with a RayDestination) */
static int trace_euler(kt_ray* q, const kt_tracer* tr, const kt_stop* stop, double rlim, int steplim, path_sink* sink)
{
    const double thetalim = stop->theta_lim;
    const double a = tr->spin, horizon = tr->horizon;
    const double k = q->k, h = q->h, Q = q->Q;
    double t = q->t, r = q->r, th = q->theta, ph = q->phi;
    double pt = q->pt, pr = q->pr, pth = q->ptheta, pph = q->pphi;
    int sr = q->rdot_sign, sth = q->thetadot_sign, flips = q->rdot_flips, cross = q->equatorial_crossings;
    int status = q->status;
    int r_seen_pos = 0, th_seen_pos = 1;
    int steps = 0;

    while (r < rlim && theta_inside(thetalim, th) && steps < steplim) {
        ++steps;
        const double sn = sin(th), cs = cos(th), sn2 = sn * sn;
        const double rho2 = r * r + (a * cs) * (a * cs);
        const double del = r * r - 2 * r + a * a;
        const double rho2del = rho2 * del;

        pt = (rho2 * (r * r + a * a) + 2 * a * a * r * sn2) * k - 2 * a * r * h;
        pt /= rho2del;
        pph = 2 * a * r * sn2 * k + (rho2 - 2 * r) * h;
        pph /= sn2 * rho2del;

        double thsq = Q + (k * a * cs + h * cs / sn) * (k * a * cs - h * cs / sn);
        thsq = thsq / (rho2 * rho2);
        if (thsq < 0 && th_seen_pos) {
            sth *= -1;
            th_seen_pos = 0;
            continue;
        }
        if (thsq >= 0) th_seen_pos = 1;
        pth = sqrt(fabs(thsq)) * sth;

        double rsq = k * pt - h * pph - rho2 * pth * pth;
        rsq = rsq * del / rho2;
        if (rsq <= 0 && r_seen_pos) {
            sr *= -1;
            r_seen_pos = 0;
            flips++;
        } else if (rsq > 0) {
            r_seen_pos = 1;
        }
        pr = sqrt(fabs(rsq)) * sr;

        double step = fabs((r - horizon) / pr) / tr->precision;
        if (step > fabs(th / pth) / tr->precision) step = fabs(th / pth) / tr->theta_precision;
        if (tr->max_tstep > 0 && r < tr->maxtstep_rlim && step > fabs(tr->max_tstep / pt)) step = fabs(tr->max_tstep / pt);
        if (tr->max_phistep > 0 && step > fabs(tr->max_phistep / pph)) step = fabs(tr->max_phistep / pph);
        if (step < KT_MIN_STEP) step = KT_MIN_STEP;
        if (rlim > 0 && r + pr * step > rlim) step = fabs((rlim - r) / pr);
        if (thetalim > 0 && th + pth * step > thetalim) step = fabs((thetalim - th) / pth);

        if (pt <= 0) status |= KT_STATUS_ERGO;
        if ((1 - 2 * r / rho2) * pt + (2 * a * r * sn2 / rho2) * pph < 0) status |= KT_STATUS_NEG_ENERGY;

        t += pt * step;
        r += pr * step;
        {
            const double tep = th;
            th += pth * step;
            if ((tep < ORC_PI_2 && th >= ORC_PI_2) || (tep > ORC_PI_2 && th <= ORC_PI_2)) ++cross;
        }
        ph += pph * step;
        if (th < 0) { th = -th; sth = -sth; ph += ORC_PI; }
        if (th > ORC_PI) { th = 2.0 * ORC_PI - th; sth = -sth; ph += ORC_PI; }

        if (r <= horizon) { status |= KT_STATUS_HORIZON; break; }
        if (path_point(sink, steps, t, r, th, ph, a)) break;
    }

    if (steps >= steplim) status |= KT_STATUS_STEPLIM;
    else if (r >= rlim) status |= KT_STATUS_RLIM;
    else if ((thetalim > 0 && th >= thetalim) || (thetalim < 0 && th <= fabs(thetalim))) status |= KT_STATUS_DEST;

    q->t = t; q->r = r; q->theta = th; q->phi = ph;
    q->pt = pt; q->pr = pr; q->ptheta = pth; q->pphi = pph;
    q->rdot_sign = sr; q->thetadot_sign = sth; q->rdot_flips = flips; q->equatorial_crossings = cross;
    q->status = status;
    if (steps > 0) q->steps += steps;
    if (status & KT_STATUS_STEPLIM) q->steps = -q->steps;
    return steps;
}

/* RK4: raytracer.cpp:755-970 (theta_max overload) and :1036-1254 (destination overload) */
static int trace_rk4(kt_ray* q, const kt_tracer* tr, const kt_stop* stop, double rlim, int steplim, path_sink* sink)
{
    const int use_dest = stop->kind != KT_STOP_THETALIM;
    const double thetalim = stop->theta_lim;
    const double a = tr->spin, horizon = tr->horizon;
    const double k = q->k, h = q->h, Q = q->Q;
    double t = q->t, r = q->r, th = q->theta, ph = q->phi;
    double pt = q->pt, pr = q->pr, pth = q->ptheta, pph = q->pphi;
    int sr = q->rdot_sign, sth = q->thetadot_sign, flips = q->rdot_flips, cross = q->equatorial_crossings;
    int status = q->status;
    int r_seen_pos = 0, th_seen_pos = 1;
    int steps = 0;

    while (r < rlim && (use_dest || theta_inside(thetalim, th)) && steps < steplim) {
        ++steps;
        const double sn = sin(th), cs = cos(th), sn2 = sn * sn;
        const double rho2 = r * r + (a * cs) * (a * cs);
        const double del = r * r - 2 * r + a * a;
        const double rho2del = rho2 * del;

        pt = (rho2 * (r * r + a * a) + 2 * a * a * r * sn2) * k - 2 * a * r * h;
        pt /= rho2del;
        pph = 2 * a * r * sn2 * k + (rho2 - 2 * r) * h;
        pph /= sn2 * rho2del;

        double thsq = Q + (k * a * cs + h * cs / sn) * (k * a * cs - h * cs / sn);
        thsq = thsq / (rho2 * rho2);
        if (thsq < 0 && th_seen_pos) { /* turning point in theta: a counted step that does not move */
            sth *= -1;
            th_seen_pos = 0;
            continue;
        }
        if (thsq >= 0) th_seen_pos = 1;
        pth = sqrt(fabs(thsq)) * sth;

        double rsq = k * pt - h * pph - rho2 * pth * pth;
        rsq = rsq * del / rho2;
        if (rsq <= 0 && r_seen_pos) {
            sr *= -1;
            r_seen_pos = 0;
            flips++;
        } else if (rsq > 0) {
            r_seen_pos = 1;
        }
        pr = sqrt(fabs(rsq)) * sr;

        const double pt1 = pt, pr1 = pr, pth1 = pth, pph1 = pph;

        double step = fabs((r - horizon) / pr1) / tr->precision;
        if (step > fabs(th / pth1) / tr->precision) step = fabs(th / pth1) / tr->theta_precision;
        if (tr->max_tstep > 0 && r < tr->maxtstep_rlim && step > fabs(tr->max_tstep / pt1)) step = fabs(tr->max_tstep / pt1);
        if (tr->max_phistep > 0 && step > fabs(tr->max_phistep / pph1)) step = fabs(tr->max_phistep / pph1);
        if (step < KT_MIN_STEP) step = KT_MIN_STEP;
        if (rlim > 0 && r + pr1 * step > rlim) step = fabs((rlim - r) / pr1);
        if (!use_dest && thetalim > 0 && th + pth1 * step > thetalim) step = fabs((thetalim - th) / pth1);

        if (pt1 <= 0) status |= KT_STATUS_ERGO;
        if ((1 - 2 * r / rho2) * pt1 + (2 * a * r * sn2 / rho2) * pph1 < 0) status |= KT_STATUS_NEG_ENERGY;

        const mom4 m2 = momenta(k, h, Q, sr, sth, r + (step / 2) * pr1, th + (step / 2) * pth1, a);
        const mom4 m3 = momenta(k, h, Q, sr, sth, r + (step / 2) * m2.pr, th + (step / 2) * m2.pth, a);
        const mom4 m4 = momenta(k, h, Q, sr, sth, r + step * m3.pr, th + step * m3.pth, a);

        const double th_prev = th;
        t += (step / 6) * (pt1 + 2 * m2.pt + 2 * m3.pt + m4.pt);
        r += (step / 6) * (pr1 + 2 * m2.pr + 2 * m3.pr + m4.pr);
        th += (step / 6) * (pth1 + 2 * m2.pth + 2 * m3.pth + m4.pth);
        if ((th_prev < ORC_PI_2 && th >= ORC_PI_2) || (th_prev > ORC_PI_2 && th <= ORC_PI_2)) ++cross;
        ph += (step / 6) * (pph1 + 2 * m2.pph + 2 * m3.pph + m4.pph);
        if (th < 0) { th = -th; sth = -sth; ph += ORC_PI; }
        if (th > ORC_PI) { th = 2.0 * ORC_PI - th; sth = -sth; ph += ORC_PI; }

        if (r <= horizon) { status |= KT_STATUS_HORIZON; break; }
        if (use_dest && dest_reached(stop, r, th, ph, th_prev)) { status |= KT_STATUS_DEST; break; }
        if (path_point(sink, steps, t, r, th, ph, a)) break;
    }

    if (steps >= steplim) status |= KT_STATUS_STEPLIM;
    else if (r >= rlim) status |= KT_STATUS_RLIM;
    else if (!use_dest && ((thetalim > 0 && th >= thetalim) || (thetalim < 0 && th <= fabs(thetalim)))) status |= KT_STATUS_DEST;

    q->t = t; q->r = r; q->theta = th; q->phi = ph;
    q->pt = pt; q->pr = pr; q->ptheta = pth; q->pphi = pph;
    q->rdot_sign = sr; q->thetadot_sign = sth; q->rdot_flips = flips; q->equatorial_crossings = cross;
    q->status = status;
    if (steps > 0) q->steps += steps;
    if (status & KT_STATUS_STEPLIM) q->steps = -q->steps;
    return steps;
}

/* DOPRI5 tableau: raytracer.cpp:1316-1330 */
#define A21 (1.0 / 5)
#define A31 (3.0 / 40)
#define A32 (9.0 / 40)
#define A41 (44.0 / 45)
#define A42 (-56.0 / 15)
#define A43 (32.0 / 9)
#define A51 (19372.0 / 6561)
#define A52 (-25360.0 / 2187)
#define A53 (64448.0 / 6561)
#define A54 (-212.0 / 729)
#define A61 (9017.0 / 3168)
#define A62 (-355.0 / 33)
#define A63 (46732.0 / 5247)
#define A64 (49.0 / 176)
#define A65 (-5103.0 / 18656)
#define B1 (35.0 / 384)
#define B3 (500.0 / 1113)
#define B4 (125.0 / 192)
#define B5 (-2187.0 / 6784)
#define B6 (11.0 / 84)
#define E1 (71.0 / 57600)
#define E3 (-71.0 / 16695)
#define E4 (71.0 / 1920)
#define E5 (-17253.0 / 339200)
#define E6 (22.0 / 525)
#define E7 (-1.0 / 40)

#define ORC_RK45_MAX_RETRIES 10000

static inline double dmax(double x, double y) { return (x < y) ? y : x; } /* std::max */
static inline double dmin(double x, double y) { return (y < x) ? y : x; } /* std::min */

/* RK45: raytracer.cpp:1260-1598 (theta_max overload) and :1600-1894 (destination overload) */
static int trace_rk45(kt_ray* q, const kt_tracer* tr, const kt_stop* stop, double rlim, int steplim, path_sink* sink)
{
    const int use_dest = stop->kind != KT_STOP_THETALIM;
    const double thetalim = stop->theta_lim;
    const double a = tr->spin, horizon = tr->horizon, tol = tr->rk45_tol;
    const double k = q->k, h = q->h, Q = q->Q;
    double t = q->t, r = q->r, th = q->theta, ph = q->phi;
    double pt, pr, pth, pph;
    int sr = q->rdot_sign, sth = q->thetadot_sign, flips = q->rdot_flips, cross = q->equatorial_crossings;
    int status = q->status;
    int r_seen_pos = 0, th_seen_pos = 1;
    int steps = 0;
    int poisoned = 0;
    double th_eq_prev = th;

    /* seed of the running step: raytracer.cpp:1341-1359 */
    {
        const double sn = sin(th), cs = cos(th), sn2 = sn * sn;
        const double rho2 = r * r + (a * cs) * (a * cs);
        const double del = r * r - 2 * r + a * a;
        pt = ((rho2 * (r * r + a * a) + 2 * a * a * r * sn2) * k - 2 * a * r * h) / (rho2 * del);
        pph = (2 * a * r * sn2 * k + (rho2 - 2 * r) * h) / (sn2 * rho2 * del);
        const double thsq = (Q + (k * a * cs + h * cs / sn) * (k * a * cs - h * cs / sn)) / (rho2 * rho2);
        pth = sqrt(fabs(thsq)) * sth;
        const double rsq = (k * pt - h * pph - rho2 * pth * pth) * del / rho2;
        pr = sqrt(fabs(rsq)) * sr;
    }
    double step = fabs((r - horizon) / pr) / tr->precision;
    if (fabs(pth) > 0 && step > fabs(th / pth) / tr->theta_precision) step = fabs(th / pth) / tr->theta_precision;
    if (tr->max_tstep > 0 && r < tr->maxtstep_rlim && step > fabs(tr->max_tstep / pt)) step = fabs(tr->max_tstep / pt);
    if (tr->max_phistep > 0 && step > fabs(tr->max_phistep / pph)) step = fabs(tr->max_phistep / pph);
    if (step < KT_MIN_STEP) step = KT_MIN_STEP;

    while (r < rlim && (use_dest || theta_inside(thetalim, th)) && steps < steplim) {
        ++steps;
        {
            const double sn = sin(th), cs = cos(th), sn2 = sn * sn;
            const double rho2 = r * r + (a * cs) * (a * cs);
            const double del = r * r - 2 * r + a * a;
            pt = ((rho2 * (r * r + a * a) + 2 * a * a * r * sn2) * k - 2 * a * r * h) / (rho2 * del);
            pph = (2 * a * r * sn2 * k + (rho2 - 2 * r) * h) / (sn2 * rho2 * del);
            const double thsq = (Q + (k * a * cs + h * cs / sn) * (k * a * cs - h * cs / sn)) / (rho2 * rho2);
            if (thsq < 0 && th_seen_pos) {
                sth *= -1;
                th_seen_pos = 0;
                continue;
            }
            if (thsq >= 0) th_seen_pos = 1;
            pth = sqrt(fabs(thsq)) * sth;
            const double rsq = (k * pt - h * pph - rho2 * pth * pth) * del / rho2;
            if (rsq <= 0 && r_seen_pos) {
                sr *= -1;
                r_seen_pos = 0;
                flips++;
            } else if (rsq > 0) {
                r_seen_pos = 1;
            }
            pr = sqrt(fabs(rsq)) * sr;
        }
        const double pt1 = pt, pr1 = pr, pth1 = pth, pph1 = pph;

        if (pt1 <= 0) status |= KT_STATUS_ERGO;
        {
            const double sn = sin(th), sn2 = sn * sn;
            const double rho2 = r * r + (a * cos(th)) * (a * cos(th));
            if ((1 - 2 * r / rho2) * pt1 + (2 * a * r * sn2 / rho2) * pph1 < 0) status |= KT_STATUS_NEG_ENERGY;
        }
        {
            double cap = fabs((r - horizon) / pr1) / tr->precision;
            if (tr->max_phistep > 0) {
                const double c = fabs(tr->max_phistep / pph1);
                if (c < cap) cap = c;
            }
            if (tr->max_tstep > 0 && r < tr->maxtstep_rlim) {
                const double c = fabs(tr->max_tstep / pt1);
                if (c < cap) cap = c;
            }
            if (step > cap) step = cap;
        }

        const double th_prev = th;
        int accepted = 0, retries = 0;
        while (!accepted) {
            double htry = step;
            int clamped = 0;
            if (use_dest) {
                if (rlim > 0 && r + pr1 * htry > rlim) { htry = fabs((rlim - r) / pr1); clamped = 1; }
                const double hd = dest_step_limit(stop, r, th, pth1);
                if (hd < htry) { htry = hd; clamped = 1; }
            } else if (thetalim > 0 && th + pth1 * htry > thetalim) {
                const double hth = fabs((thetalim - th) / pth1);
                if (hth < htry) { htry = hth; clamped = 1; }
            }

            const mom4 m2 = momenta(k, h, Q, sr, sth, r + htry * A21 * pr1, th + htry * A21 * pth1, a);
            const mom4 m3 = momenta(k, h, Q, sr, sth, r + htry * (A31 * pr1 + A32 * m2.pr),
                                    th + htry * (A31 * pth1 + A32 * m2.pth), a);
            const mom4 m4 = momenta(k, h, Q, sr, sth, r + htry * (A41 * pr1 + A42 * m2.pr + A43 * m3.pr),
                                    th + htry * (A41 * pth1 + A42 * m2.pth + A43 * m3.pth), a);
            const mom4 m5 = momenta(k, h, Q, sr, sth, r + htry * (A51 * pr1 + A52 * m2.pr + A53 * m3.pr + A54 * m4.pr),
                                    th + htry * (A51 * pth1 + A52 * m2.pth + A53 * m3.pth + A54 * m4.pth), a);
            const mom4 m6 = momenta(k, h, Q, sr, sth,
                                    r + htry * (A61 * pr1 + A62 * m2.pr + A63 * m3.pr + A64 * m4.pr + A65 * m5.pr),
                                    th + htry * (A61 * pth1 + A62 * m2.pth + A63 * m3.pth + A64 * m4.pth + A65 * m5.pth), a);

            double r_new = r + htry * (B1 * pr1 + B3 * m3.pr + B4 * m4.pr + B5 * m5.pr + B6 * m6.pr);
            double th_new = th + htry * (B1 * pth1 + B3 * m3.pth + B4 * m4.pth + B5 * m5.pth + B6 * m6.pth);
            double t_new = t + htry * (B1 * pt1 + B3 * m3.pt + B4 * m4.pt + B5 * m5.pt + B6 * m6.pt);
            double ph_new = ph + htry * (B1 * pph1 + B3 * m3.pph + B4 * m4.pph + B5 * m5.pph + B6 * m6.pph);
            /* polar reflection happens inside the trial and is not undone on rejection (raytracer.cpp:1498-1499) */
            if (th_new < 0) { th_new = -th_new; sth = -sth; ph_new += ORC_PI; }
            if (th_new > ORC_PI) { th_new = 2.0 * ORC_PI - th_new; sth = -sth; ph_new += ORC_PI; }

            const mom4 m7 = momenta(k, h, Q, sr, sth, r_new, th_new, a);

            const double err_r = htry * (E1 * pr1 + E3 * m3.pr + E4 * m4.pr + E5 * m5.pr + E6 * m6.pr + E7 * m7.pr);
            const double err_th = htry * (E1 * pth1 + E3 * m3.pth + E4 * m4.pth + E5 * m5.pth + E6 * m6.pth + E7 * m7.pth);
            const double sc_r = tol * (1.0 + dmax(fabs(r), fabs(r_new)));
            const double sc_th = tol * (1.0 + dmax(fabs(th), fabs(th_new)));
            const double err = sqrt(0.5 * ((err_r / sc_r) * (err_r / sc_r) + (err_th / sc_th) * (err_th / sc_th)));

            double fac = 0.9 * pow(1.0 / dmax(err, 1e-10), 0.2);
            fac = dmax(0.1, dmin(5.0, fac));
            const double step_new = htry * fac;

            if (err <= 1.0) {
                t = t_new; r = r_new; th = th_new; ph = ph_new;
                pt = m7.pt; pr = m7.pr; pth = m7.pth; pph = m7.pph;
                if (!clamped) step = dmax(step_new, KT_MIN_STEP);
                accepted = 1;
            } else {
                step = dmax(step_new, KT_MIN_STEP);
                if (step <= KT_MIN_STEP) {
                    t = t_new; r = r_new; th = th_new; ph = ph_new;
                    pt = m7.pt; pr = m7.pr; pth = m7.pth; pph = m7.pph;
                    accepted = 1;
                }
            }
            /* DEVIATION (documented above): a NaN error norm, or an absurd retry count, retires the ray */
            if (!accepted && (err != err || ++retries >= ORC_RK45_MAX_RETRIES)) { poisoned = 1; break; }
        }
        if (poisoned) break;

        if ((th_eq_prev < ORC_PI_2 && th >= ORC_PI_2) || (th_eq_prev > ORC_PI_2 && th <= ORC_PI_2)) ++cross;
        th_eq_prev = th;

        if (r <= horizon) { status |= KT_STATUS_HORIZON; break; }
        if (use_dest && dest_reached(stop, r, th, ph, th_prev)) { status |= KT_STATUS_DEST; break; }
        if (path_point(sink, steps, t, r, th, ph, a)) break;
    }

    if (steps >= steplim || poisoned) status |= KT_STATUS_STEPLIM;
    else if (r >= rlim) status |= KT_STATUS_RLIM;
    else if (!use_dest && ((thetalim > 0 && th >= thetalim) || (thetalim < 0 && th <= fabs(thetalim)))) status |= KT_STATUS_DEST;

    q->t = t; q->r = r; q->theta = th; q->phi = ph;
    q->pt = pt; q->pr = pr; q->ptheta = pth; q->pphi = pph;
    q->rdot_sign = sr; q->thetadot_sign = sth; q->rdot_flips = flips; q->equatorial_crossings = cross;
    q->status = status;
    if (steps > 0) q->steps += steps;
    if (status & KT_STATUS_STEPLIM) q->steps = -q->steps;
    return steps;
}

/* run_raytrace, parallel path: raytracer.cpp:63-127 / :972-1034 */
int64_t orc_trace(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_stop* stop, const kt_run* run, int nthreads)
{
    const int rk45 = run->integrator == KT_RK45;
    const int steplim = (run->steplim > 0) ? run->steplim : (rk45 ? KT_RK45_STEPLIM : KT_STEPLIM);
    int64_t traced = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : traced)
    for (int64_t i = 0; i < n; i++) {
        if (rays[i].steps < 0) continue;
        if (rays[i].steps >= steplim) continue;
        if (rk45) trace_rk45(&rays[i], tr, stop, run->r_max, steplim, NULL);
        else if (run->integrator == KT_EULER) trace_euler(&rays[i], tr, stop, run->r_max, steplim, NULL);
        else trace_rk4(&rays[i], tr, stop, run->r_max, steplim, NULL);
        traced++;
    }
    return traced;
}

/* run_raytrace, serial path with an outfile: raytracer.cpp:86-100 / :995-1008.  Rays are propagated in index order;
 * offsets[i]..offsets[i+1] are the points ray i wrote (a skipped ray writes none -- and no blank lines either).
 * points == NULL: count only (the rays are still propagated in place, so pass a scratch copy). */
int64_t orc_trace_paths(kt_ray* rays, int64_t n, const kt_tracer* tr, const kt_stop* stop, const kt_run* run,
                        const kt_path_opts* opts, int64_t* offsets, double* points)
{
    const int rk45 = run->integrator == KT_RK45;
    const int steplim = (run->steplim > 0) ? run->steplim : (rk45 ? KT_RK45_STEPLIM : KT_STEPLIM);
    int64_t total = 0;
    for (int64_t i = 0; i < n; i++) {
        offsets[i] = total;
        if (rays[i].steps < 0) continue;
        if (rays[i].steps >= steplim) continue;
        path_sink k = {opts, points ? points + 4 * total : NULL, 0, 0};
        if (rk45) trace_rk45(&rays[i], tr, stop, run->r_max, steplim, &k);
        else if (run->integrator == KT_EULER) trace_euler(&rays[i], tr, stop, run->r_max, steplim, &k);
        else trace_rk4(&rays[i], tr, stop, run->r_max, steplim, &k);
        total += k.n;
    }
    offsets[n] = total;
    return total;
}

/* ------------------------------------------------------------------------------------------------ */
/* reductions: the apps' serial loops                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* emissivity.cpp:96-126, emissivity_rd.cpp:108-138 */
void orc_bin_emissivity(const kt_ray* rays, int64_t n, const kt_emis_bins* b, kt_emis_out* out)
{
    for (int i = 0; i < b->Nr; i++) {
        out->N[i] = 0;
        out->flux[i] = out->emis[i] = out->redshift_sum[i] = out->time_sum[i] = 0;
    }
    memset(out->counters, 0, sizeof(out->counters));
    const double log_dr = log(b->dr);
    for (int64_t i = 0; i < n; i++) {
        const kt_ray* q = &rays[i];
        out->counters[KT_CNT_SLOTS]++;
        if (q->steps != -1 || q->status != 0) {
            out->counters[KT_CNT_TRACED]++;
            out->counters[KT_CNT_STEPS] += llabs((long long)q->steps);
            if (q->status & KT_STATUS_STEPLIM) out->counters[KT_CNT_STEPLIM]++;
            else if (q->status & KT_STATUS_RLIM) out->counters[KT_CNT_RLIM]++;
            else if (q->status & KT_STATUS_DEST) out->counters[KT_CNT_DEST]++;
            if (q->status & KT_STATUS_HORIZON) out->counters[KT_CNT_HORIZON]++;
        }
        if (!(q->steps > 0)) continue;
        int hit;
        if (b->hit_filter == KT_HIT_THETA) hit = q->theta >= b->theta_lim - 1e-3;
        else hit = q->r * cos(q->theta) < 1E-2;
        if (hit && q->redshift > 0 && q->r >= b->r_isco) {
            const int ir = b->logbin ? (int)(log(q->r / b->r_min) / log_dr) : (int)((q->r - b->r_min) / b->dr);
            if (ir >= 0 && ir < b->Nr) {
                ++out->N[ir];
                out->flux[ir] += 1 / (b->n_primary * pow(q->redshift, 1));
                out->emis[ir] += 1 / pow(q->redshift, b->gamma);
                out->redshift_sum[ir] += q->redshift;
                out->time_sum[ir] += q->t;
            }
            out->counters[KT_CNT_DISC]++;
        }
    }
}

/* powerlaw3: imageplane_disc_image.cpp:20-28 */
static double powerlaw3(double r, double q1, double rb1, double q2, double rb2, double q3)
{
    if (r < rb1) return pow(r, -1 * q1);
    else if (r < rb2) return pow(rb1, q2 - q1) * pow(r, -1 * q2);
    else return pow(rb1, q2 - q1) * pow(rb2, q3 - q2) * pow(r, -1 * q3);
}

/* imageplane_disc_image.cpp:122-161 (raw sums; the divisions of :165-174 are the caller's) */
void orc_bin_image(const kt_ray* rays, int64_t n, const kt_image_bins* b, kt_image_out* out)
{
    const int64_t npix = (int64_t)b->img_Nx * b->img_Ny;
    for (int64_t i = 0; i < npix; i++) {
        out->flux[i] = out->r[i] = out->phi[i] = out->enshift[i] = out->time[i] = out->emis[i] = 0;
        out->N[i] = 0;
    }
    memset(out->counters, 0, sizeof(out->counters));
    for (int64_t i = 0; i < n; i++) {
        const kt_ray* q = &rays[i];
        out->counters[KT_CNT_SLOTS]++;
        if (q->steps != -1 || q->status != 0) {
            out->counters[KT_CNT_TRACED]++;
            out->counters[KT_CNT_STEPS] += llabs((long long)q->steps);
            if (q->status & KT_STATUS_STEPLIM) out->counters[KT_CNT_STEPLIM]++;
            else if (q->status & KT_STATUS_RLIM) out->counters[KT_CNT_RLIM]++;
            else if (q->status & KT_STATUS_DEST) out->counters[KT_CNT_DEST]++;
            if (q->status & KT_STATUS_HORIZON) out->counters[KT_CNT_HORIZON]++;
        }
        if (!(q->steps > 0)) continue;
        int hit;
        if (b->hit_filter == KT_HIT_DEST) hit = (q->status & KT_STATUS_DEST) && q->r < b->r_disc && q->redshift > 0;
        else if (b->hit_filter == KT_HIT_THETA)
            hit = q->theta >= b->theta_lim - 1e-3 && q->r >= b->r_isco && q->r < b->r_disc && q->redshift > 0;
        else hit = q->r * cos(q->theta) < 1E-2 && q->r >= b->r_isco && q->r < b->r_disc && q->redshift > 0;
        if (hit) {
            int ix = (int)((q->alpha - b->x0) / b->img_dx);
            int iy = (int)((q->beta - b->y0) / b->img_dy);
            if (b->flip_image) iy = b->img_Ny - iy - 1;
            if (ix >= 0 && ix < b->img_Nx && iy >= 0 && iy < b->img_Ny) {
                const int64_t p = (int64_t)ix * b->img_Ny + iy;
                ++out->N[p];
                const double em = powerlaw3(q->r, b->q1, b->rb1, b->q2, b->rb2, b->q3);
                out->flux[p] += em / pow(q->redshift, 3);
                out->r[p] += q->r;
                out->phi[p] += q->phi;
                out->enshift[p] += 1. / q->redshift;
                out->time[p] += q->t;
                out->emis[p] += em;
                out->counters[KT_CNT_DISC]++;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* caustic / critical-curve maps: caustic_discplane.cpp:146-493                                       */
/* ------------------------------------------------------------------------------------------------ */
static inline int caustic_valid_hit(const kt_ray* q, const kt_caustic_cfg* c) /* caustic_discplane.cpp:177-182 */
{
    return q->steps > 0 && q->r >= c->r_isco && q->r < c->r_disc && q->redshift > 0;
}

static inline int caustic_order(const kt_ray* q) /* caustic_discplane.cpp:199-203 */
{
    const int phi_ord = (int)(fabs(q->phi) / (2 * ORC_PI));
    const int r_ord = q->rdot_flips / 2;
    return phi_ord > r_ord ? phi_ord : r_ord;
}

static inline void caustic_disc_xy(double phi_acc, double r, double* xd, double* yd) /* caustic_discplane.cpp:170-174 */
{
    const double phi_s = atan2(sin(phi_acc), cos(phi_acc));
    *xd = r * cos(phi_s);
    *yd = r * sin(phi_s);
}

void orc_caustic_maps(const kt_ray* rays, int64_t n, int32_t plane_Ny, int bundles, double eps_x, double eps_y,
                      const kt_caustic_cfg* c, kt_caustic_out* o)
{
    const int Nx = c->img_Nx, Ny = c->img_Ny;
    const int per = bundles ? 5 : 1;
    static const kt_ray unused = {.steps = -1};
#define RAY(ix, iy, sat) ((((int64_t)(ix) * plane_Ny + (iy)) * per + (sat)) < n ? &rays[((int64_t)(ix) * plane_Ny + (iy)) * per + (sat)] : &unused)
#define PIX(ix, iy) ((int64_t)(ix) * Ny + (iy))
    o->disc_count = o->cnt_horizon = o->cnt_rlim = o->cnt_steplim = o->cnt_outofrange = o->cnt_other = o->suppressed = 0;
    /* centre rays: caustic_discplane.cpp:220-253 (bundles) / :351-381 (grid) */
    for (int ix = 0; ix < Nx; ix++)
        for (int iy = 0; iy < Ny; iy++) {
            const kt_ray* q = RAY(ix, iy, 0);
            const int64_t p = PIX(ix, iy);
            if (caustic_valid_hit(q, c)) {
                const double phi_s = atan2(sin(q->phi), cos(q->phi));
                o->hit[p] = 1;
                o->radius[p] = q->r;
                o->phi[p] = phi_s;
                o->x_disc[p] = q->r * cos(phi_s);
                o->y_disc[p] = q->r * sin(phi_s);
                o->order[p] = caustic_order(q);
                o->redshift[p] = q->redshift;
                ++o->disc_count;
            } else {
                o->hit[p] = 0;
                o->radius[p] = o->phi[p] = o->x_disc[p] = o->y_disc[p] = o->redshift[p] = 0;
                o->order[p] = -1;
            }
            /* diagnostic counts: caustic_discplane.cpp:256-274 */
            if (q->steps > 0 && (q->r < c->r_isco || q->r >= c->r_disc || q->redshift <= 0)) ++o->cnt_outofrange;
            else if (q->steps <= 0 || !(q->status & KT_STATUS_DEST)) {
                if (q->status & KT_STATUS_HORIZON) ++o->cnt_horizon;
                else if (q->status & KT_STATUS_RLIM) ++o->cnt_rlim;
                else if (q->status & KT_STATUS_STEPLIM) ++o->cnt_steplim;
                else ++o->cnt_other;
            }
        }
    /* Jacobian */
    for (int ix = 0; ix < Nx; ix++)
        for (int iy = 0; iy < Ny; iy++) {
            const int64_t p = PIX(ix, iy);
            o->det_J[p] = NAN;
            o->sign_J[p] = 0;
            if (!o->hit[p]) continue;
            double det;
            if (bundles) { /* caustic_discplane.cpp:277-333 */
                const kt_ray *re = RAY(ix, iy, 1), *rw = RAY(ix, iy, 2), *rn = RAY(ix, iy, 3), *rs = RAY(ix, iy, 4);
                if (!caustic_valid_hit(re, c) || !caustic_valid_hit(rw, c) || !caustic_valid_hit(rn, c) || !caustic_valid_hit(rs, c))
                    continue;
                const kt_ray* rc = RAY(ix, iy, 0);
                const int flips = rc->rdot_flips;
                const double phi_c = rc->phi;
                const int match = re->rdot_flips == flips && rw->rdot_flips == flips && rn->rdot_flips == flips &&
                                  rs->rdot_flips == flips && fabs(re->phi - phi_c) < ORC_PI_2 && fabs(rw->phi - phi_c) < ORC_PI_2 &&
                                  fabs(rn->phi - phi_c) < ORC_PI_2 && fabs(rs->phi - phi_c) < ORC_PI_2;
                if (!match) { o->det_J[p] = KT_CAUSTIC_SENTINEL; continue; }
                double xe, ye, xw, yw, xn, yn, xs, ys;
                caustic_disc_xy(re->phi, re->r, &xe, &ye);
                caustic_disc_xy(rw->phi, rw->r, &xw, &yw);
                caustic_disc_xy(rn->phi, rn->r, &xn, &yn);
                caustic_disc_xy(rs->phi, rs->r, &xs, &ys);
                const double dxd_da = (xe - xw) / (2 * eps_x), dxd_db = (xn - xs) / (2 * eps_y);
                const double dyd_da = (ye - yw) / (2 * eps_x), dyd_db = (yn - ys) / (2 * eps_y);
                det = dxd_da * dyd_db - dxd_db * dyd_da;
            } else { /* caustic_discplane.cpp:409-447 */
                if (ix == 0 || ix == Nx - 1 || iy == 0 || iy == Ny - 1) continue;
                const int64_t e = PIX(ix + 1, iy), w = PIX(ix - 1, iy), nn = PIX(ix, iy + 1), ss = PIX(ix, iy - 1);
                if (!(o->hit[e] && o->hit[w] && o->hit[nn] && o->hit[ss])) continue;
                const int ord = o->order[p];
                if (!(o->order[e] == ord && o->order[w] == ord && o->order[nn] == ord && o->order[ss] == ord)) {
                    o->det_J[p] = KT_CAUSTIC_SENTINEL;
                    continue;
                }
                const double dxd_dx = (o->x_disc[e] - o->x_disc[w]) / (2 * c->dx), dxd_dy = (o->x_disc[nn] - o->x_disc[ss]) / (2 * c->dy);
                const double dyd_dx = (o->y_disc[e] - o->y_disc[w]) / (2 * c->dx), dyd_dy = (o->y_disc[nn] - o->y_disc[ss]) / (2 * c->dy);
                det = dxd_dx * dyd_dy - dxd_dy * dyd_dx;
            }
            o->det_J[p] = det;
            o->sign_J[p] = (det > 0) ? 1.0 : (det < 0) ? -1.0 : 0.0;
        }
    /* suppression of alternating-sign pixels on a snapshot of sign_J: caustic_discplane.cpp:467-492 */
    if (c->suppress) {
        double* snap = (double*)malloc(sizeof(double) * (size_t)Nx * Ny);
        memcpy(snap, o->sign_J, sizeof(double) * (size_t)Nx * Ny);
        static const int dix[4] = {-1, 1, 0, 0}, diy[4] = {0, 0, -1, 1};
        for (int ix = 0; ix < Nx; ix++)
            for (int iy = 0; iy < Ny; iy++) {
                const double s = snap[PIX(ix, iy)];
                if (s == 0.0) continue;
                int n_same = 0, n_opp = 0;
                for (int d = 0; d < 4; d++) {
                    const int jx = ix + dix[d], jy = iy + diy[d];
                    if (jx < 0 || jx >= Nx || jy < 0 || jy >= Ny) continue;
                    const double sn = snap[PIX(jx, jy)];
                    if (sn == 0.0) continue;
                    if (sn * s > 0) ++n_same;
                    else ++n_opp;
                }
                if (n_opp > n_same && n_opp >= 2) {
                    o->det_J[PIX(ix, iy)] = KT_CAUSTIC_SENTINEL;
                    o->sign_J[PIX(ix, iy)] = 0.0;
                    ++o->suppressed;
                }
            }
        free(snap);
    }
#undef RAY
#undef PIX
}
