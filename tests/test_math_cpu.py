"""The branch-free exp of the log-sum-exp kernels (autodiff_b200/csrc/adb_math.cuh: exp_nonpos) checked on the host: the
constants are parsed out of the header and the algorithm is replayed with EXACT fused multiply-adds (rational arithmetic,
one rounding per operation like the GPU's DFMA), then compared with a 60-digit reference."""
import math
import os
import re
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autodiff_b200", "csrc", "adb_math.cuh")


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _constants():
    src = open(HEADER).read()
    body = src[src.index("double exp_nonpos(double x)"):]
    body = body[:body.index("return x <")]
    log2e, magic = re.search(r"fma\(x, ([0-9.e+-]+), ([0-9.e+-]+)\)", body).groups()
    ln2 = re.findall(r"fma\(n, (-[0-9.e+-]+), \w+\)", body)
    p0 = re.search(r"double p = ([0-9.e+-]+);", body).group(1)
    coefs = re.findall(r"p = fma\(p, r, ([0-9.e+-]+)\);", body)
    return float(log2e), float(magic), [float(v) for v in ln2], float(p0), [float(c) for c in coefs]


def _exp_nonpos(x, k):
    log2e, magic, (ln2_hi, ln2_lo), p0, coefs = k
    if x < -700.0:
        return 0.0
    z = _fma(x, log2e, magic)
    n = z - magic
    r = _fma(n, ln2_hi, x)
    r = _fma(n, ln2_lo, r)
    p = p0
    for c in coefs:
        p = _fma(p, r, c)
    return math.ldexp(p, int(n))          # the kernel builds 2^n in the exponent field: exact for n >= -1022


def test_exp_nonpos_constants_are_a_degree_13_taylor_series():
    log2e, magic, ln2, p0, coefs = _constants()
    assert magic == 1.5 * 2.0 ** 52 and abs(log2e - math.log2(math.e)) < 1e-16
    assert len(ln2) == 2 and abs(-(ln2[0] + ln2[1]) - math.log(2.0)) < 1e-16
    taylor = [float(Fraction(1, math.factorial(j))) for j in range(13, -1, -1)]
    assert [p0] + coefs == taylor


def test_exp_nonpos_is_within_one_ulp_of_exp():
    getcontext().prec = 60
    k = _constants()
    rng = np.random.default_rng(0)
    xs = np.concatenate([-rng.random(3000) * 50, -rng.random(3000) * 700, -rng.random(1000) * 1e-3,
                         [0.0, -0.34657359027997264, -0.3465735902799727, -1e-300, -699.999, -math.log(2.0) * 0.5]])
    worst = Decimal(0)
    for x in xs:
        x = float(x)
        ref = Decimal(x).exp()
        got = Decimal(_exp_nonpos(x, k))
        worst = max(worst, abs(got - ref) / Decimal(math.ulp(float(ref))))
    assert worst < Decimal("0.85"), worst     # measured 0.805 ulp
    assert _exp_nonpos(-701.0, k) == 0.0 and _exp_nonpos(float("-inf"), k) == 0.0 and _exp_nonpos(0.0, k) == 1.0


# ------------------------------------------------------------------------------------------------ the 32-entry-table variant
def _tab_constants():
    src = open(HEADER).read()
    tab = src[src.index("exp_tab32[32] = {") + len("exp_tab32[32] = {"):]
    tab = [float(v) for v in tab[:tab.index("}")].replace("\n", " ").split(",")]
    body = src[src.index("double exp_nonpos_tab(double x)"):]
    body = body[:body.index("return x <")]
    scale, magic = re.search(r"fma\(x, ([0-9.e+-]+), ([0-9.e+-]+)\)", body).groups()
    ln2 = re.findall(r"fma\(n, (-[0-9.e+-]+), \w+\)", body)
    p0 = re.search(r"double p = ([0-9.e+-]+);", body).group(1)
    coefs = re.findall(r"p = fma\(p, r, ([0-9.e+-]+)\);", body)
    return tab, float(scale), float(magic), [float(v) for v in ln2], float(p0), [float(c) for c in coefs]


def _exp_nonpos_tab(x, k):
    tab, scale, magic, (c_hi, c_lo), p0, coefs = k
    if x < -700.0:
        return 0.0
    z = _fma(x, scale, magic)
    n = z - magic
    r = _fma(n, c_hi, x)
    r = _fma(n, c_lo, r)
    p = p0
    for c in coefs:
        p = _fma(p, r, c)
    i = int(n)
    v = float(Fraction(tab[i & 31]) * Fraction(p))          # one rounding: t * p
    return math.ldexp(v, i >> 5)                              # exact scaling by 2^q (normal range for x >= -700)


def test_exp_nonpos_tab_constants():
    getcontext().prec = 60
    tab, scale, magic, ln2, p0, coefs = _tab_constants()
    assert len(tab) == 32 and magic == 1.5 * 2.0 ** 52 and abs(scale - 32.0 / math.log(2.0)) < 1e-14
    for j, t in enumerate(tab):                              # correctly rounded 2^(j/32)
        exact = (Decimal(2).ln() * j / 32).exp()
        assert abs(Decimal(t) - exact) <= Decimal(math.ulp(t)) / 2, j
    assert len(ln2) == 2 and abs(Decimal(-ln2[0]) + Decimal(-ln2[1]) - Decimal(2).ln() / 32) < Decimal("1e-29")
    assert float(Fraction(-ln2[0]) * 65536) == -ln2[0] * 65536 and math.frexp(-ln2[0])[0] * 2 ** 36 % 1 == 0   # 36-bit head: n * hi is exact
    assert [p0] + coefs == [float(Fraction(1, math.factorial(j))) for j in range(6, -1, -1)]


def test_exp_nonpos_tab_is_within_two_ulp_of_exp():
    getcontext().prec = 60
    k = _tab_constants()
    rng = np.random.default_rng(1)
    xs = np.concatenate([-rng.random(3000) * 50, -rng.random(3000) * 700, -rng.random(1000) * 1e-3,
                         [0.0, -0.010830424696249145, -0.01083042469624915, -1e-300, -699.999, -math.log(2.0) / 64, -math.log(2.0) / 32]])
    worst = Decimal(0)
    for x in xs:
        x = float(x)
        ref = Decimal(x).exp()
        got = Decimal(_exp_nonpos_tab(x, k))
        worst = max(worst, abs(got - ref) / Decimal(math.ulp(float(ref))))
    assert worst < Decimal("2.0"), worst      # measured 1.76 ulp
    assert _exp_nonpos_tab(-701.0, k) == 0.0 and _exp_nonpos_tab(float("-inf"), k) == 0.0 and _exp_nonpos_tab(0.0, k) == 1.0
