"""CPU: the numerical claims behind the hand-written sqrt / exp of the sweeps (cglb_b200/csrc/common.cuh), checked in
50-digit decimal arithmetic with the constants parsed from the CUDA source, so that the test follows the kernels:
  * exp: r = -s - n ln2/2^TB with n = rint(-s 2^TB/ln2), e^r by the short polynomial, 2^(n/2^TB) from a table;
  * sqrt: MUFU.RSQ64H seed (relative error <= 2^-20, measured) + one third-order correction."""
import math
import os
import re
from decimal import Decimal, getcontext

getcontext().prec = 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "cglb_b200", "csrc", "common.cuh")).read()
LN2 = Decimal(2).ln()


def _const(pattern):
    m = re.search(pattern, SRC)
    assert m, pattern
    return [Decimal(g) for g in m.groups()]


def test_exp_reduction_constants_are_consistent():
    c6, c10 = _const(r"constexpr double C = \(TB == 6\) \? ([0-9.e+-]+) : ([0-9.e+-]+);")
    l6, l10 = _const(r"constexpr double L = \(TB == 6\) \? ([0-9.e+-]+) : ([0-9.e+-]+);")
    for tb, c, l in ((6, c6, l6), (10, c10, l10)):
        assert abs(c - Decimal(2) ** tb / LN2) / c < Decimal("2e-16")          # 2^TB / ln 2 to double precision
        assert abs(l - LN2 / Decimal(2) ** tb) / l < Decimal("2e-16")          # ln 2 / 2^TB
    # ln2/2^TB is ONE double (no hi/lo split): its representation error delta_L puts s * delta_L into r, i.e. a relative
    # error of s * 3.3e-17 on e^-s (1e-15 at s = 30) -- next to the s * 1.1e-16 .. 2.2e-16 that the rounding of s itself costs any
    # fp64 evaluation, and far inside the 1e-10 per-matvec tolerance
    assert abs(Decimal(float(l10)) - LN2 / Decimal(2) ** 10) / (LN2 / Decimal(2) ** 10) < Decimal("4e-17")
    assert abs(Decimal(float(l6)) - LN2 / Decimal(2) ** 6) / (LN2 / Decimal(2) ** 6) < Decimal("4e-17")
    # the magic number 1.5 * 2^52 makes `t - MAGIC` exact and leaves n in the low word for |n| < 2^31
    (magic,) = _const(r"const double MAGIC = ([0-9.]+);")
    assert magic == Decimal(3) * Decimal(2) ** 51


def _max_rel_err(poly, h, samples=2001):
    worst = Decimal(0)
    for i in range(samples):
        r = -h + 2 * h * Decimal(i) / Decimal(samples - 1)
        worst = max(worst, abs(poly(r) - r.exp()) / r.exp())
    return worst


def test_exp_polynomials_reach_double_precision_on_their_intervals():
    # TB = 10: p = c3 r + (1/2 + eps); p = p r + 1; e^r = 1 + r p      on |r| <= ln2 / 2^11 (n is the nearest integer)
    c3, c2a, c2b = _const(r"p = fma\(r, ([0-9.e+-]+), ([0-9.e+-]+) \+ ([0-9.e+-]+)\);")
    h10 = LN2 / Decimal(2) ** 11
    err10 = _max_rel_err(lambda r: 1 + r * ((c3 * r + (c2a + c2b)) * r + 1), h10)
    assert err10 < Decimal("1e-16"), err10
    # without the Chebyshev shift of the quadratic coefficient the truncation error would be r^4/24 = 5.4e-16
    err_plain = _max_rel_err(lambda r: 1 + r * ((c3 * r + c2a) * r + 1), h10)
    assert err_plain > 3 * err10
    # TB = 6: degree 5 in r after the leading 1 on |r| <= ln2 / 2^7
    c5, c4 = _const(r"p = fma\(r, ([0-9.e+-]+), ([0-9.e+-]+)\);\s*\n\s*p = fma\(p, r, 1\.6666")
    (c3b,) = _const(r"p = fma\(p, r, (1\.6666[0-9.e+-]+)\);")
    h6 = LN2 / Decimal(2) ** 7
    err6 = _max_rel_err(lambda r: 1 + r * ((((c5 * r + c4) * r + c3b) * r + Decimal("0.5")) * r + 1), h6)
    assert err6 < Decimal("1e-16"), err6


def test_sqrt_third_order_correction():
    # y = (1 + delta) / sqrt(q), |delta| <= 2^-20 (measured for rsqrt.approx.ftz.f64):
    # g = q y, e = 1 - g y, s = g + g e (1/2 + 3/8 e)  ->  relative error ~ (5/16) e^3 << 2^-53
    for q in (Decimal("1e-12"), Decimal("0.3"), Decimal(7), Decimal("480000")):
        for k in range(-8, 9):
            delta = Decimal(k) / 8 * Decimal(2) ** -20
            y = (1 + delta) / q.sqrt()
            g = q * y
            e = 1 - g * y
            s = g + g * (e * (Decimal("0.375") * e + Decimal("0.5")))
            assert abs(s - q.sqrt()) / q.sqrt() < Decimal("1e-17")


def test_clamp_keeps_the_exponent_field_in_range():
    # kappa() clamps s <= 693 (q <= 693^2 resp. 693) so that e^-s >= 2^-1000: the exponent-field add of fast_exp_neg
    # (res in [1, 2) scaled by 2^m, m = n >> TB >= -1000) stays in the normal range
    kern = open(os.path.join(ROOT, "cglb_b200", "csrc", "kmv_impl.cuh")).read()
    hi_m = int(re.search(r"kClampHiMatern = (0x[0-9a-fA-F]+);", kern).group(1), 16)
    hi_r = int(re.search(r"kClampHiRbf = (0x[0-9a-fA-F]+);", kern).group(1), 16)
    import struct
    qm = struct.unpack(">d", struct.pack(">II", hi_m, 0xFFFFFFFF))[0]      # largest double with this high word
    qr = struct.unpack(">d", struct.pack(">II", hi_r, 0xFFFFFFFF))[0]
    for s_max in (qm ** 0.5, qr):
        assert s_max < 693.01
        m = int(-s_max / 0.6931471805599453) - 1
        assert m >= -1001 and 1023 + m > 0


def _fma(a, b, c):
    """exact fused multiply-add in double precision (the DFMA of the kernels): one rounding of the exact a*b + c."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def test_kernel_map_end_to_end_in_emulated_fp64():
    """kappa(q) = (1 + s) e^-s, s = sqrt(q), step by step as common.cuh / kmv_impl.cuh compute it (exact FMA emulation,
    a 20-bit reciprocal-square-root seed with a zero low word, the 1024-entry table), against 50-digit arithmetic:
    the relative error stays below s * 2.6e-16 + 4e-16 over the whole clamped range. The s-proportional term is the
    rounding of s itself (the corrected sqrt is faithful, |delta s| < ulp(s) <= s * 2.2e-16, against s * 1.1e-16 for a
    correctly rounded one -- a term ANY fp64 evaluation that forms s as a double pays, torch's included) plus the representation error of the single-double ln2/2^TB (s * 3.3e-17)."""
    import math
    import random
    import struct
    c10 = float(_const(r"constexpr double C = \(TB == 6\) \? [0-9.e+-]+ : ([0-9.e+-]+);")[0])
    l10 = float(_const(r"constexpr double L = \(TB == 6\) \? [0-9.e+-]+ : ([0-9.e+-]+);")[0])
    c3, c2a, c2b = (float(v) for v in _const(r"p = fma\(r, ([0-9.e+-]+), ([0-9.e+-]+) \+ ([0-9.e+-]+)\);"))
    magic = 6755399441055744.0
    table = [float(Decimal(2) ** (Decimal(j) / 1024)) for j in range(1024)]      # correctly rounded 2^(j/1024)
    rnd = random.Random(0)
    worst = 0.0
    for _ in range(3000):
        q = math.exp(rnd.uniform(math.log(1e-8), math.log(480000.0)))
        # seed: relative error <= 2^-20, only the high word is produced (MUFU.RSQ64H), low word zero
        y = (1.0 + rnd.uniform(-1, 1) * 2.0 ** -20) / math.sqrt(q)
        y = struct.unpack(">d", struct.pack(">Q", struct.unpack(">Q", struct.pack(">d", y))[0] & 0xFFFFFFFF00000000))[0]
        g = q * y
        e = _fma(-g, y, 1.0)
        p = _fma(e, 0.375, 0.5)
        t = e * p
        s = _fma(g, t, g)
        assert abs(Decimal(s) - Decimal(q).sqrt()) < Decimal(math.ulp(s))          # faithful
        # exp(-s)
        tt = _fma(s, -c10, magic)
        n = struct.unpack(">q", struct.pack(">d", tt))[0] & 0xFFFFFFFF
        n = n - (1 << 32) if n >= (1 << 31) else n
        nf = tt - magic
        assert nf == float(n)
        r = _fma(nf, -l10, -s)
        pp = _fma(r, c3, c2a + c2b)
        pp = _fma(pp, r, 1.0)
        res = table[n & 1023] * _fma(r, pp, 1.0)
        ex = math.ldexp(res, n >> 10)
        kap = _fma(s, ex, ex)
        sq = Decimal(q).sqrt()
        ref = (1 + sq) * (-sq).exp()
        err = float(abs(Decimal(kap) - ref) / ref)
        assert err <= float(sq) * 2.6e-16 + 4e-16, (q, err)
        worst = max(worst, err / (float(sq) * 2.6e-16 + 4e-16))
    assert worst > 0.05          # the bound is not vacuous


def test_exponent_insert_through_the_pre_biased_table():
    """fast_exp_neg<10> (round 2): the 1024-entry table stores 2^(j/1024) with j << 10 subtracted from its high word
    (context.cu), and ONE integer multiply-add  hi = n * 2^10 + hi(T'[n & 1023])  forms 2^(n/1024) for n = (n >> 10) 2^10 + j
    without masking n.  Integer emulation of the two instructions against exact scaling, over the whole range the clamp allows
    (s <= 693 -> n >= -1023999), and the claim that scaling by 2^k before or after the multiplication gives the same bits."""
    import random
    import struct
    from decimal import Decimal, getcontext
    getcontext().prec = 50
    src = open(os.path.join(ROOT, "cglb_b200", "csrc", "context.cu")).read()
    assert "bits -= (unsigned long long)j << 42;" in src                     # the table construction this test mirrors
    table = [float(Decimal(2) ** (Decimal(j) / 1024)) for j in range(1024)]
    biased = []
    for j, t in enumerate(table):
        bits = struct.unpack("<Q", struct.pack("<d", t))[0]
        biased.append((bits - (j << 42)) & 0xFFFFFFFFFFFFFFFF)
    rng = random.Random(0)
    for n in [0, -1, -1023, -1024, -1025, -1023999, -512000] + [-rng.randrange(0, 1024000) for _ in range(2000)]:
        j = n & 1023                                                         # two's complement: the low 10 bits
        tb = biased[j]
        hi = ((n * 1024) + (tb >> 32)) & 0xFFFFFFFF                          # mad.lo.s32 hi, n, 1024, hi(T')
        val = struct.unpack("<d", struct.pack("<Q", (hi << 32) | (tb & 0xFFFFFFFF)))[0]
        k = (n - j) // 1024                                                  # n >> 10 (arithmetic)
        assert val == math.ldexp(table[j], k), (n, j, k)
        x = 1.0 + rng.random() * 2.0 ** -11                                  # 1 + r p of the polynomial
        assert math.ldexp(table[j], k) * x == math.ldexp(table[j] * x, k)    # scale first or last: same bits (no underflow)
