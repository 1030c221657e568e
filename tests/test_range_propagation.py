"""CPU check of the interval bounds behind the exact mode's range-propagated acceptance (raytrace_cpu_b200/csrc/kt_trace.cuh:
rhs_head, rk4_step_v3; kt_device.cuh: ydiv).

The guard-free quotient sequence is valid when its divisor lies in 2^+-900, its quotient is a normal number and its numerator is
at least 2^-969 in magnitude.  The kernel tests only the quotient (2^-500 < |q| < 2^1017) and relies on bounds that follow from
three input checks: 2^-20 < |r| < 2^40, 2^-100 < |sin theta| <= 1, and k, h, Q, k a in {0} u [2^-100, 2^100), |a| in {0} u
[2^-20, 2^10].  Here the expressions of one right-hand-side evaluation are run in float64 over the corners of that input box
(and random interior points) and every claim of the comment block is asserted: nothing overflows, every divisor is 0 or inside
[2^-333, 2^163], and an accepted quotient implies an admissible numerator."""
import itertools

import numpy as np

L2 = np.log2


def evaluate(r, sn, cs, k, h, Q, a):
    """One right-hand-side evaluation in the reference's operand order (kerr.h:308-334); returns (numerator, divisor, quotient)
    triples of the five quotients and the arguments of the two roots."""
    with np.errstate(all="ignore"):
        sn2 = sn * sn
        rr = r * r
        ac = a * cs
        rho2 = rr + ac * ac
        dl = (rr - 2 * r) + a * a
        rho2del = rho2 * dl
        ptn = (rho2 * (rr + a * a) + 2 * a * a * r * sn2) * k - 2 * a * r * h
        pphn = 2 * a * r * sn2 * k + (rho2 - 2 * r) * h
        pphd = sn2 * rho2del
        hc = h * cs
        rho4 = rho2 * rho2
        pt = ptn / rho2del
        pph = pphn / pphd
        hcs = hc / sn
        ka = k * a
        num = Q + (ka * cs + hcs) * (ka * cs - hcs)
        thsq = num / rho4
        pth = np.sqrt(np.abs(thsq))
        rnum = (k * pt - h * pph - rho2 * pth * pth) * dl
        rsq = rnum / rho2
    return [(ptn, rho2del, pt), (pphn, pphd, pph), (hc, sn, hcs), (num, rho4, thsq), (rnum, rho2, rsq)], (thsq, rsq)


def accepted(q):
    return np.isfinite(q) and 2.0 ** -500 < abs(q) < 2.0 ** 1017


def check_point(r, sn, cs, k, h, Q, a):
    quotients, roots = evaluate(r, sn, cs, k, h, Q, a)
    chain_ok = True
    for num, den, q in quotients:
        if not chain_ok:
            break                              # the kernel's flag is already false: the step is redone with the plain operators
        assert np.isfinite(num) and np.isfinite(den), (num, den)           # no overflow anywhere before the quotient
        assert den == 0 or 2.0 ** -333 <= abs(den) <= 2.0 ** 163, den      # divisor bound of the comment block
        if accepted(q):
            assert abs(num) >= 2.0 ** -969, (num, den, q)                  # exact residual a - b q
            assert 2.0 ** -900 <= abs(den) <= 2.0 ** 900
        else:
            chain_ok = False
    if chain_ok:
        for x in roots:                        # root arguments are accepted quotients: inside the root sequence's domain
            assert 2.0 ** -969 <= abs(x) < 2.0 ** 1023


def test_corners_of_the_input_box():
    rs = [2.0 ** -20 * (1 + 2 ** -30), 1.0, 1.0632, 3.0, 2.0 ** 40 * (1 - 2 ** -30), -1.5, -(2.0 ** 39)]
    sns = [2.0 ** -100 * (1 + 2 ** -30), 1e-8, 0.5, 1.0, -1.0]
    consts = [0.0, 2.0 ** -100, -(2.0 ** -100), 1.0, -3.7, 2.0 ** 100 * (1 - 2 ** -30), -(2.0 ** 100) * (1 - 2 ** -30)]
    spins = [0.0, 2.0 ** -20, 0.998, -0.998, 2.0 ** 10]
    n = 0
    for r, sn, a in itertools.product(rs, sns, spins):
        cs_mag = np.sqrt(max(0.0, 1.0 - min(1.0, sn * sn)))
        for cs in (cs_mag, -cs_mag, 0.0, 6.123233995736766e-17):
            for k, h, Q in itertools.product([0.0, 2.0 ** -100, 1.0, 2.0 ** 100 * (1 - 2 ** -30)], consts, consts):
                check_point(r, sn, cs, k, h, Q, a)
                n += 1
    assert n > 100000


def test_random_points_of_the_input_box():
    rng = np.random.default_rng(20261019)
    for _ in range(40000):
        r = rng.choice([-1.0, 1.0]) * 2.0 ** rng.uniform(-20, 40)
        sn = rng.choice([-1.0, 1.0]) * 2.0 ** rng.uniform(-100, 0)
        cs = rng.choice([-1.0, 1.0]) * np.sqrt(max(0.0, 1 - min(1.0, sn * sn))) if rng.random() < 0.8 else rng.uniform(-1, 1)
        k, h, Q = (rng.choice([-1.0, 1.0]) * 2.0 ** rng.uniform(-100, 100) if rng.random() < 0.9 else 0.0 for _ in range(3))
        a = rng.choice([0.0, 0.998, -0.998, 2.0 ** rng.uniform(-20, 10)])
        check_point(r, sn, cs, k, h, Q, a)


def test_horizon_neighbourhood_where_delta_cancels():
    """r within a few ulps of the horizon: del = (r^2 - 2r) + a^2 cancels to a multiple of its operands' grid -- 0 or >= 2^-93."""
    a = 0.998
    hor = 1 + np.sqrt((1 - a) * (1 + a))
    r = hor
    for _ in range(200):
        dl = (r * r - 2 * r) + a * a
        assert dl == 0 or abs(dl) >= 2.0 ** -93
        check_point(r, 0.7, np.sqrt(1 - 0.49), 1.0, 2.0, 5.0, a)
        r = np.nextafter(r, 2.0)
