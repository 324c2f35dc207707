"""Independent pure-Python restatement of the reference's pointwise path (+ mpmath truth).

TEST INFRASTRUCTURE ONLY (same rule as sk_oracle.c).  Python floats are IEEE binary64 and
CPython never fuses a*b+c, so with the reference's operation order this module must agree
with the C oracle BIT FOR BIT; tests/test_oracle_crosscheck.py asserts that on small cases.
It also holds the high-precision (mpmath) truths used to pin the two things the reference
delegates to Julia Base: sinpi/cospi (grids) and hypot.

Citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from math import comb

import mpmath as mp

mp.mp.prec = 240

INTERIOR, ENDPOINT, INTERIOR2 = 0, 1, 2
NONE, BOUNDED_LINEAR, SEMI_INF, INF = 0, 1, 2, 3


# --------------------------------------------------------------------------- jets (derivatives.jl:105-134)
def seed(x, D):
    return [x] + [1.0 if i == 0 else 0.0 for i in range(D)]


def one(D):
    return [1.0] + [0.0] * D


def jadd(x, y):
    return [a + b for a, b in zip(x, y)]


def jsub(x, y):
    return [a - b for a, b in zip(x, y)]


def jscale(a, y):
    return [a * c for c in y]


def jmul(x, y):
    out = []
    for k in range(len(x)):
        s = (comb(k, 0) * x[0]) * y[k]
        for i in range(1, k + 1):
            s = s + (comb(k, i) * x[i]) * y[k - i]
        out.append(s)
    return out


# --------------------------------------------------------------------------- Julia Base pieces
def hypot_cr(x, y):
    """Correctly rounded hypot via mpmath (truth for Base.hypot's fma branch)."""
    if math.isinf(x) or math.isinf(y):
        return math.inf
    return float(mp.sqrt(mp.mpf(x) ** 2 + mp.mpf(y) ** 2))


def sinpi_cr(t):
    if t == 0.0:
        return t
    return float(mp.sin(mp.pi * mp.mpf(t)))


def cospi_cr(t):
    u = mp.mpf(1) / 2 - mp.mpf(t)
    if u == 0:
        return 0.0
    return float(mp.sin(mp.pi * u))


# --------------------------------------------------------------------------- transformations.jl
class Tr:
    def __init__(self, kind, a=0.0, b=1.0):
        self.kind = kind
        if kind == BOUNDED_LINEAR:            # :154-161
            if not (math.isfinite(a) and math.isfinite(b)):
                raise ValueError("DomainError")
            s, m = (b - a) / 2, (a + b) / 2
            if not s > 0:
                raise ValueError("DomainError")
            self.p0, self.p1 = m, s
        elif kind == SEMI_INF:                # :211-215
            if not math.isfinite(a) or not (math.isfinite(b) and b != 0):
                raise ValueError("DomainError")
            self.p0, self.p1 = float(a), float(b)
        elif kind == INF:                     # :285-289
            if not math.isfinite(a) or not (math.isfinite(b) and b > 0):
                raise ValueError("DomainError")
            self.p0, self.p1 = float(a), float(b)
        else:
            self.p0, self.p1 = 0.0, 1.0


def transform_from(t, x):
    if t.kind == BOUNDED_LINEAR:
        return x * t.p1 + t.p0                                  # :180
    if t.kind == SEMI_INF:
        if 1 - x == 0:
            return math.copysign(math.inf, t.p1 * (1 + x))
        return t.p0 + (t.p1 * (1 + x)) / (1 - x)                # :244
    if t.kind == INF:
        den = math.sqrt(1 - x * x)
        num = t.p1 * x
        if den == 0:
            return t.p0 + math.copysign(math.inf, num)
        return t.p0 + num / den                                 # :309
    return x


def _to_real(t, y):
    if t.kind == BOUNDED_LINEAR:
        return (y - t.p0) / t.p1                                # :185
    if t.kind == SEMI_INF:
        if y == math.inf or y == -math.inf:
            return 1.0                                          # :250-251
        z = y - t.p0
        return (z - t.p1) / (z + t.p1)
    if t.kind == INF:
        if math.isinf(y):
            return 1.0 if y > 0 else -1.0                       # :315-316
        z = y - t.p0
        return z / hypot_cr(z, t.p1)
    return y


def transform_to(t, yjet, ext2=False):
    D = len(yjet) - 1
    if t.kind == NONE:
        return list(yjet)
    if t.kind == BOUNDED_LINEAR:
        return [_to_real(t, yjet[0])] + [c / t.p1 for c in yjet[1:]]    # :188-195
    x0 = _to_real(t, yjet[0])
    if D == 0:
        return [x0]
    L = t.p1
    if t.kind == SEMI_INF:
        Q = (x0 - 1) * (x0 - 1)
        x1 = (yjet[1] * Q) / (2 * L)                            # :263-264
        if D == 1:
            return [x0, x1]
        if D > 2 or not ext2:
            raise NotImplementedError("derivative order")        # :266
        omx = 1 - x0
        xpp = -((omx * omx) * omx) / ((2 * L) * L)
        return [x0, x1, xpp * (yjet[1] * yjet[1]) + (Q / (2 * L)) * yjet[2]]
    Q = 1 - x0 * x0
    sQ = math.sqrt(Q)
    x1 = ((yjet[1] * Q) * sQ) / L                               # :328-330
    if D == 1:
        return [x0, x1]
    if D > 2 or not ext2:
        raise NotImplementedError("derivative order")            # :332
    xpp = -(((3 * x0) * Q) * Q) / (L * L)
    return [x0, x1, xpp * (yjet[1] * yjet[1]) + ((Q * sQ) / L) * yjet[2]]


# --------------------------------------------------------------------------- chebyshev.jl:56-70,88-108
def cheb_table(N, xjet):
    D = len(xjet) - 1
    out = [one(D)]
    fp, fpp = one(D), list(xjet)
    for _ in range(2, N + 1):
        f = jsub(jmul(jscale(2.0, xjet), fp), fpp)
        out.append(f)
        fpp, fp = fp, f
    return out


def gridpoint(kind, N, i):
    if kind == INTERIOR:
        return sinpi_cr((N - 2 * i + 1) / float(2 * N))
    if kind == ENDPOINT:
        return cospi_cr(1 / 2.0) if N == 1 else cospi_cr((N - i) / float(N - 1))
    return cospi_cr((N - i + 1) / float(N + 1))


# --------------------------------------------------------------------------- smolyak_traversal.jl
def total_length(kind, b):
    if kind == ENDPOINT:
        return 1 if b == 0 else (1 << b) + 1
    if kind == INTERIOR:
        return 3 ** b
    return (1 << (b + 1)) - 1


def block_length(kind, b):
    if kind == ENDPOINT:
        return b + 1 if b <= 1 else 1 << (b - 1)
    if kind == INTERIOR:
        return 1 if b == 0 else 2 * 3 ** (b - 1)
    return 1 << b


def naive_indices(kind, d, B, M):
    """test/test_smolyak_traversal.jl:74-90 — filter the full Cartesian product (column-major)."""
    import itertools
    M = min(M, B)
    m = total_length(kind, M)
    btab = [M] * m
    for b in range(M - 1, -1, -1):
        for i in range(total_length(kind, b)):
            btab[i] = b
    out = []
    for rev in itertools.product(range(1, m + 1), repeat=d):     # last varies fastest
        ix = rev[::-1]                                           # => first varies fastest
        if sum(btab[i - 1] for i in ix) <= B:
            out.append(ix)
    return out


def augment_coefficients(kind, d, B1, M1, B2, M2, theta1):
    """src/smolyak_api.jl:155-208 — PaddingIterator: one pass over the traversal of the larger basis; the smaller
    traversal advances only on a match and, once exhausted, never matches again (sentinel (0, 0, …))."""
    it1 = naive_indices(kind, d, B1, M1)
    it2 = naive_indices(kind, d, B2, M2)
    assert len(theta1) == len(it1)
    out, i = [], 0
    cur = it1[0] if it1 else None
    for ix in it2:
        if cur is not None and ix == cur:
            out.append(theta1[i])
            if i + 1 < len(it1):
                i += 1
                cur = it1[i]
            else:
                cur = None
        else:
            out.append(0.0 * theta1[0])
    return out


def shuffle_by_blocks(kind, b):
    """test/test_smolyak_traversal.jl:38-52 — block-by-block set differences of nested grids."""
    def g(bb):
        n = total_length(kind, bb)
        return [gridpoint(kind, n, i) for i in range(1, n + 1)]
    g0 = g(b)
    tol = math.sqrt(2.220446049250313e-16)

    def isin(x, grid):
        return any(abs(x - y) <= tol for y in grid)
    blocks = []
    for bb in range(b, 0, -1):
        gb, gbm1 = g(bb), g(bb - 1)
        blocks.append([i + 1 for i, x in enumerate(g0) if isin(x, gb) and not isin(x, gbm1)])
    blocks.append([(len(g0) + 1) // 2])
    return [i for blk in reversed(blocks) for i in blk]


# --------------------------------------------------------------------------- evaluation
def uni_lincomb(N, theta, y, tr=None, D=0, ext2=False):
    tr = tr or Tr(NONE)
    x = transform_to(tr, seed(y, D), ext2)
    tab = cheb_table(N, x)
    s = jscale(theta[0], tab[0])
    for k in range(1, N):
        s = jadd(s, jscale(theta[k], tab[k]))
    return s


def smolyak_lincomb(kind, d, B, M, theta, y, trs=None, spec=None, ext2=False):
    M = min(M, B)
    H = total_length(kind, M)
    idx = naive_indices(kind, d, B, M)
    trs = trs or [Tr(NONE)] * d
    if spec is None:
        tabs = [[c[0] for c in cheb_table(H, transform_to(trs[j], [y[j]]))] for j in range(d)]
        s = None
        for k, ix in enumerate(idx):
            b = tabs[0][ix[0] - 1]
            for j in range(1, d):
                b = b * tabs[j][ix[j] - 1]
            s = theta[k] * b if s is None else s + theta[k] * b
        return [s]
    deg = [max(q[j] for q in spec) for j in range(d)]
    tabs = [cheb_table(H, transform_to(trs[j], seed(y[j], deg[j]), ext2)) for j in range(d)]
    s = None
    for k, ix in enumerate(idx):
        term = []
        for q in spec:
            b = tabs[0][ix[0] - 1][q[0]]
            for j in range(1, d):
                b = b * tabs[j][ix[j] - 1][q[j]]
            term.append(theta[k] * b)
        s = term if s is None else [a + t for a, t in zip(s, term)]
    return s


# --------------------------------------------------------------------------- mpmath truths
def cheb_truth(n, x, r=0):
    """r-th derivative of T_n at x, exactly (polynomial evaluated in 240-bit arithmetic)."""
    x = mp.mpf(x)
    return mp.diff(lambda t: mp.chebyt(n, t), x, r) if r else mp.chebyt(n, x)


def transform_to_truth(t, y, r):
    """r-th derivative of the map y -> x (mpmath), for the 2nd-derivative extension."""
    A, L = mp.mpf(t.p0), mp.mpf(t.p1)
    if t.kind == BOUNDED_LINEAR:
        f = lambda v: (v - A) / L
    elif t.kind == SEMI_INF:
        f = lambda v: ((v - A) - L) / ((v - A) + L)
    elif t.kind == INF:
        f = lambda v: (v - A) / mp.sqrt((v - A) ** 2 + L ** 2)
    else:
        f = lambda v: v
    return mp.diff(f, mp.mpf(y), r) if r else f(mp.mpf(y))
