"""Coefficients of the erfcx(x) = exp(x^2) erfc(x) approximation used by the charged-system pair kernel
(lumol_b200/csrc/pairs_lj2.cu, cq_erfcx): degree-14 polynomial in zz, where z = (x - 4) / (x + 4) and zz maps the
range of z for x in [0, XMAX] onto [-1, 1].  Chebyshev coefficients by discrete orthogonality at 600 nodes in 40-digit
arithmetic (mpmath), converted exactly to the monomial basis; prints the constants and the error of a float64 Horner
evaluation against mpmath.

    python tools/fit_erfcx.py
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 40
XMAX = 3.45
SHIFT = 4.0
DEGREE = 14
NODES = 600


def erfcx(x):
    return mp.exp(mp.mpf(x) ** 2) * mp.erfc(mp.mpf(x))


def main():
    zlo = (0 - SHIFT) / (0 + SHIFT)
    zhi = (XMAX - SHIFT) / (XMAX + SHIFT)

    def x_of(zz):
        z = (zz * (zhi - zlo) + (zhi + zlo)) / 2
        return SHIFT * (1 + z) / (1 - z)

    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / NODES) for k in range(NODES)]
    values = [erfcx(x_of(t)) for t in nodes]
    cheb = []
    for j in range(DEGREE + 1):
        total = sum(values[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / NODES) for k in range(NODES))
        cheb.append(2 * total / NODES if j > 0 else total / NODES)
    basis = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for n in range(2, DEGREE + 1):
        a = [mp.mpf(0)] + [2 * t for t in basis[n - 1]]
        b = basis[n - 2] + [mp.mpf(0)] * (len(a) - len(basis[n - 2]))
        basis.append([a[i] - b[i] for i in range(len(a))])
    mono = [mp.mpf(0)] * (DEGREE + 1)
    for j in range(DEGREE + 1):
        for i, t in enumerate(basis[j]):
            mono[i] += cheb[j] * t
    scale_a = 2 / (zhi - zlo)
    scale_b = -(zhi + zlo) / (zhi - zlo)
    print(f"// zz = z * {float(scale_a)!r} + {float(scale_b)!r}, z = (x - {SHIFT}) / (x + {SHIFT}), x in [0, {XMAX}]")
    print("constexpr double CQ_ERFCX[%d] = {" % (DEGREE + 1))
    for m in mono:
        print(f"    {float(m)!r},")
    print("};")
    xs = np.linspace(0.0, XMAX, 4001)
    coefficients = np.array([float(m) for m in mono])
    z = (xs - SHIFT) / (xs + SHIFT)
    zz = z * float(scale_a) + float(scale_b)
    p = np.zeros_like(zz)
    for m in coefficients[::-1]:
        p = p * zz + m
    reference = np.array([float(erfcx(x)) for x in xs])
    print("// float64 Horner: max relative error %.2e" % np.abs(p / reference - 1).max())


if __name__ == "__main__":
    main()
