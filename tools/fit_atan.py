#!/usr/bin/env python
"""Fits and measures the atan polynomial of gs_fused.cu (atan01): atan(a) ~ a + a*s*q(s), s = a*a, a in [0, 1].

q is fitted by iteratively re-weighted least squares on Chebyshev nodes (near-minimax), rounded to binary32, and the
total error (approximation + rounding) is measured by simulating the kernel's FMA Horner scheme in binary32 on 2e6
points.  Result for 8 coefficients: max |error| = 6.6e-8 rad.  The kernel's margins (7.5e-7 rad azimuth, 1e-6 rad
altitude, 4e-6 rad inclination angle) are sized >= 4x the budget: polynomial 6.6e-8 + quotient (2 ulp) 1.2e-7 +
fixed-point rounding."""
import numpy as np


def fit(n):
    k = np.arange(4000)
    a = 0.5 * (1 - np.cos(np.pi * (k + 0.5) / 4000))
    a = a[a > 1e-6]
    s = a * a
    target = (np.arctan(a) - a) / (a * s)
    V = np.vander(s, n, increasing=True)
    w = a * s
    coef = np.linalg.lstsq(V * w[:, None], target * w, rcond=None)[0]
    for _ in range(60):
        err = np.abs((V @ coef - target) * w)
        ww = w * (1 + 50 * (err / err.max()) ** 4)
        coef = np.linalg.lstsq(V * ww[:, None], target * ww, rcond=None)[0]
    return coef


def fma(a, b, c):
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


if __name__ == "__main__":
    for n in (7, 8, 9):
        cf = fit(n).astype(np.float32)
        a = np.linspace(0, 1, 2000001).astype(np.float32)
        s = (a * a).astype(np.float32)
        q = np.full_like(a, cf[-1])
        for k in range(n - 2, -1, -1):
            q = fma(q, s, cf[k])
        t = fma((a * s).astype(np.float32), q, a)
        err = np.abs(t.astype(np.float64) - np.arctan(a.astype(np.float64)))
        print(n, "max error %.3e" % err.max(), [float(x) for x in cf])
