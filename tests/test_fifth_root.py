"""CPU check, in exact rational arithmetic, of the fifth root behind the RK45 step-size controller
(raytrace_cpu_b200/csrc/kt_device.cuh: kt_root5, which replaces pow(1 / err, 0.2) of raytracer.cpp:1517).

Every multiplication / FMA of the sequence is evaluated exactly (fractions.Fraction) and rounded once to double; the seed is a
model of the MUFU pair lg2 / ex2 in float (x^(-1/5) rounded to float after a relative perturbation of up to 2^-20, four times
the hardware's worst case); the result is compared with x^0.2 -- the fifth root computed to 400 bits times the factor for the
excess of the double 0.2 over 1/5, which reaches 2.8 ulp at the ends of the domain.  Cases: random arguments over the
whole clamped domain [2^-40, 2^40], the controller's usual range, and arguments within an ulp of an exact fifth power (the hard
cases for a final rounding).  The device code itself is compared with pow() by test_parity_gpu.py."""
import math
import random
import struct
from fractions import Fraction as Fr


def rn(x):
    return float(x)            # Fraction -> double is correctly rounded


def mul(a, b):
    return rn(Fr(a) * Fr(b))


def fma(a, b, c):
    return rn(Fr(a) * Fr(b) + Fr(c))


def to_float32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def ulp(y):
    return 2.0 ** (math.frexp(y)[1] - 53)


def iroot5(n):
    x = 1 << ((n.bit_length() + 4) // 5)
    while True:
        y = (4 * x + n // x ** 4) // 5
        if y >= x:
            return x
        x = y


def true_root(x, bits=400):
    f = Fr(x)
    return Fr(iroot5((f.numerator << (5 * bits)) // f.denominator), 1 << bits)


def root5(x, delta):
    w = float(to_float32(float(1 / true_root(x, 80)) * (1 + delta)))
    w2 = mul(w, w)
    w4 = mul(w2, w2)
    e = fma(-x, mul(w4, w), 1.0)
    w = fma(mul(w, 0.2), e, w)
    w2 = mul(w, w)
    w4 = mul(w2, w2)
    y = mul(x, w4)
    y2 = mul(y, y)
    res = fma(mul(y2, y2), y, -x)
    lf = to_float32(math.log2(x) * (1 + delta))          # the seed's float logarithm
    excess = mul(mul(7.69547959311662e-18, lf), y)
    return rn(Fr(y) + Fr(fma(-res, mul(0.2, w4), excess)))


EXCESS = Fr(0.2) - Fr(1, 5)       # the reference's exponent is the double 0.2: x^0.2 = x^(1/5) * exp(EXCESS ln x)


def true_power(x):
    return true_root(x) * (1 + EXCESS * Fr(math.log(x)))      # (the second-order term is 1e-33)


def test_fifth_root_sequence_is_within_one_ulp_of_the_power():
    rng = random.Random(20261019)
    worst = 0.0
    for i in range(4500):
        if i % 3 == 0:
            x = rng.uniform(1, 2) * 2.0 ** rng.randint(-40, 39)
        elif i % 3 == 1:
            x = 10 ** rng.uniform(-3, 10)
        else:
            b = rng.uniform(1, 2) * 2.0 ** rng.randint(-7, 7)
            x = rn(Fr(b) ** 5)
            x += rng.choice([-1, 0, 1]) * ulp(x)
        t = true_power(x)
        for delta in (-2.0 ** -20, 2.0 ** -20, rng.uniform(-1, 1) * 2.0 ** -20):
            worst = max(worst, float(abs(Fr(root5(x, delta)) - t) / Fr(ulp(float(t)))))
    assert worst < 1.0, worst
