"""Generates the polynomial coefficient tables of rcppsmc_b200/csrc/fastmath.cuh (near-minimax: Chebyshev
interpolation at 60 digits, converted to the monomial basis, rounded to fp64) and reports the achieved error."""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    ys = [f(x) for x in xs]
    # Newton-free: solve Vandermonde in high precision (degree <= 13, 60 digits: fine)
    A = mp.matrix(n, n)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
    c = mp.lu_solve(A, mp.matrix(ys))
    return [c[i] for i in range(n)]


def max_rel_err(f, coefs, a, b, rel=True, m=4001):
    worst = mp.mpf(0)
    cd = [mp.mpf(float(c)) for c in coefs]
    for i in range(m):
        x = a + (b - a) * i / (m - 1)
        p = mp.polyval(cd[::-1], x)
        t = f(x)
        e = abs(p - t) / (abs(t) if rel and t != 0 else 1)
        worst = max(worst, e)
    return worst


def emit(name, coefs):
    body = ", ".join(f"{float(c)!r}" for c in coefs)
    print(f"static __constant__ double {name}[{len(coefs)}] = {{{body}}};")


ln2 = mp.log(2)
# exp(r), r in [-ln2/2, ln2/2], degree 11
c = cheb_fit(mp.exp, -ln2 / 2, ln2 / 2, 11)
emit("FM_EXP", c)
print("// exp rel err", mp.nstr(max_rel_err(mp.exp, c, -ln2 / 2, ln2 / 2), 3))
# atanh(s)/s as a polynomial in z = s^2, s in [0, 0.1716], degree 8
zmax = (mp.sqrt(2) - 1) ** 2 / (1 + mp.sqrt(2) - 1 + 1) ** 2  # placeholder, recomputed below
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
f_at = lambda z: mp.mpf(1) if z == 0 else mp.atanh(mp.sqrt(z)) / mp.sqrt(z)
c = cheb_fit(f_at, mp.mpf(0), smax ** 2 * mp.mpf("1.02"), 8)
emit("FM_ATANH", c)
print("// atanh(s)/s rel err", mp.nstr(max_rel_err(f_at, c, mp.mpf(0), smax ** 2), 3))
# sin(pi t)/t and cos(pi t) as polynomials in z = t^2, t in [0, 0.25]
f_s = lambda z: mp.pi if z == 0 else mp.sin(mp.pi * mp.sqrt(z)) / mp.sqrt(z)
f_c = lambda z: mp.cos(mp.pi * mp.sqrt(z))
c = cheb_fit(f_s, mp.mpf(0), mp.mpf("0.0625"), 7)
emit("FM_SINPI", c)
print("// sin(pi t)/t rel err", mp.nstr(max_rel_err(f_s, c, mp.mpf(0), mp.mpf("0.0625")), 3))
c = cheb_fit(f_c, mp.mpf(0), mp.mpf("0.0625"), 8)
emit("FM_COSPI", c)
print("// cos(pi t) rel err", mp.nstr(max_rel_err(f_c, c, mp.mpf(0), mp.mpf("0.0625")), 3))
print("// LN2_HI/LO", repr(float(mp.mpf(float(ln2)))), repr(float(ln2 - mp.mpf(float(ln2)))))
