#!/usr/bin/env python
"""Fit the fp32 polynomial kernels used by the CUDA path (sin/cos of pi*r on
|r|<=1/4, atan(t) in degrees on [0,1]) and report their fp32 Horner error.
Coefficients printed here are pasted into manytor_b200/csrc/mt_math.cuh."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def minimax_fit(f, lo, hi, deg, weight=None, iters=30):
    """Remez-ish: iteratively re-weighted least squares on a dense grid."""
    x = np.cos(np.linspace(0, np.pi, 20001))[::-1] * (hi - lo) / 2 + (hi + lo) / 2
    y = f(x)
    w = np.ones_like(x) if weight is None else weight(x)
    ww = np.ones_like(x)
    best = None
    for _ in range(iters):
        V = np.vander(x, deg + 1, increasing=True)
        coef, *_ = np.linalg.lstsq(V * (w * ww)[:, None], y * w * ww, rcond=None)
        err = np.abs((V @ coef - y) * w)
        if best is None or err.max() < best[0]:
            best = (err.max(), coef.copy())
        ww = ww * (1 + 4 * err / err.max()) ** 0.5
        ww /= ww.mean()
    return best[1], best[0]


def horner32(coef, s):
    r = np.full_like(s, f32(coef[-1]))
    for c in coef[-2::-1]:
        r = fma(r, s, np.full_like(s, f32(c)))
    return r


# ---- sin(pi r)/r and cos(pi r) as polynomials in u = r^2, |r| <= 0.25 ------
def sinc(u):
    r = np.sqrt(np.maximum(u, 1e-300))
    return np.where(u > 0, np.sin(np.pi * r) / r, np.pi)


cs, es = minimax_fit(sinc, 0.0, 0.0625, 3)
cc, ec = minimax_fit(lambda u: np.cos(np.pi * np.sqrt(u)), 0.0, 0.0625, 4)
print("sin coef (u^0..):", [float(f32(c)) for c in cs], "fit err", es)
print("cos coef (u^0..):", [float(f32(c)) for c in cc], "fit err", ec)
r = np.linspace(-0.25, 0.25, 2000001).astype(f32)
u = (r * r).astype(f32)
s32 = (horner32(cs, u) * r).astype(f32)
c32 = horner32(cc, u)
print("sin(pi r) fp32 max abs err", np.abs(s32 - np.sin(np.pi * r.astype(np.float64))).max())
print("cos(pi r) fp32 max abs err", np.abs(c32 - np.cos(np.pi * r.astype(np.float64))).max())

# ---- the same on |r| <= 0.5 (sincos_deg2's half-turn reduction: one sign flip instead of a quadrant fix-up) ----
cs2, es2 = minimax_fit(sinc, 0.0, 0.25, 4)
cc2, ec2 = minimax_fit(lambda u: np.cos(np.pi * np.sqrt(u)), 0.0, 0.25, 5)
print("half-turn sin coef (u^0..):", [float(f32(c)) for c in cs2], "fit err", es2)
print("half-turn cos coef (u^0..):", [float(f32(c)) for c in cc2], "fit err", ec2)
r = np.linspace(-0.5, 0.5, 4000001).astype(f32)
u = (r * r).astype(f32)
print("half-turn sin(pi r) fp32 max abs err", np.abs((horner32(cs2, u) * r).astype(f32) - np.sin(np.pi * r.astype(np.float64))).max())
print("half-turn cos(pi r) fp32 max abs err", np.abs(horner32(cc2, u) - np.cos(np.pi * r.astype(np.float64))).max())

# ---- atan(t)/t in degrees as polynomial in s = t^2, t in [0,1] --------------
def atd(s):
    t = np.sqrt(np.maximum(s, 1e-300))
    return np.where(s > 0, np.degrees(np.arctan(t)) / t, 180 / np.pi)


for deg in (7, 8, 9):
    ca, ea = minimax_fit(atd, 0.0, 1.0, deg)
    t = np.linspace(0, 1, 2000001).astype(f32)
    s = (t * t).astype(f32)
    a32 = (horner32(ca, s) * t).astype(f32)
    print(f"atan deg {deg}: coef", [float(f32(c)) for c in ca])
    print(f"   fit err {ea:.3e} deg; fp32 max abs err {np.abs(a32 - np.degrees(np.arctan(t.astype(np.float64)))).max():.3e} deg")
