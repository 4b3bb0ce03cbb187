"""Fit the polynomial used by the device GELU (csrc/afr_math.cuh).

gelu(x) = relu(x) - t*E*R(t),  gelu'(x) = 0.5 + sign(x)*(0.5 - E*R(t)) + x*E/sqrt(2pi)
with t = min(|x|, T), E = exp(-t^2/2), R(t) = Phi(-t)*exp(t^2/2) = erfcx(t/sqrt2)/2.
R is smooth and slowly varying, so a low-degree polynomial in v = 2t/T - 1 fits it; the
fit minimises max |max(1,t) * E * (R - Rhat)| (that weight bounds both the forward and the
derivative error).  Prints coefficients and the fp32-evaluated error versus exact erf GELU.
"""
import sys
import numpy as np
from scipy.optimize import linprog
from scipy.special import erfcx, erf

T = float(sys.argv[2]) if len(sys.argv) > 2 else 5.5
deg = int(sys.argv[1]) if len(sys.argv) > 1 else 9

t = np.linspace(0, T, 4001)
v = 2 * t / T - 1
E = np.exp(-0.5 * t * t)
R = 0.5 * erfcx(t / np.sqrt(2))
w = np.maximum(1, t) * E
A = np.polynomial.chebyshev.chebvander(v, deg)
# minimise e s.t. |w*(A c - R)| <= e
n = deg + 1
c_obj = np.zeros(n + 1); c_obj[-1] = 1
Aub = np.vstack([np.hstack([w[:, None] * A, -np.ones((len(t), 1))]),
                 np.hstack([-w[:, None] * A, -np.ones((len(t), 1))])])
bub = np.concatenate([w * R, -w * R])
res = linprog(c_obj, A_ub=Aub, b_ub=bub, bounds=[(None, None)] * n + [(0, None)], method="highs")
cheb = res.x[:n]
print("deg", deg, "T", T, "weighted minimax err", res.x[-1])
mono_v = np.polynomial.chebyshev.cheb2poly(cheb)          # monomial in v
# monomial in t: substitute v = a t - 1
P = np.polynomial.Polynomial(mono_v)(np.polynomial.Polynomial([-1.0, 2.0 / T]))
mono_t = P.coef
print("mono_v:", ", ".join(f"{c:.9e}f" for c in mono_v))
print("mono_t:", ", ".join(f"{c:.9e}f" for c in mono_t))

def eval32(coef, z):
    acc = np.full_like(z, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        acc = (acc * z + np.float32(c)).astype(np.float32)
    return acc

x = np.linspace(-8, 8, 400001).astype(np.float32)
tt = np.minimum(np.abs(x), np.float32(T)).astype(np.float32)
Ef = np.exp2((tt * tt * np.float32(-0.5 * np.log2(np.e))).astype(np.float32)).astype(np.float32)
exact = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
exact_d = 0.5 * (1 + erf(x.astype(np.float64) / np.sqrt(2))) + x.astype(np.float64) * np.exp(-0.5 * x.astype(np.float64) ** 2) / np.sqrt(2 * np.pi)
for name, coef, z in (("v", mono_v, (tt * np.float32(2 / T) - np.float32(1)).astype(np.float32)), ("t", mono_t, tt)):
    Rh = eval32(coef.astype(np.float32), z)
    g = (np.maximum(x, 0) - (tt * Ef).astype(np.float32) * Rh).astype(np.float32)
    ER = (Ef * Rh).astype(np.float32)
    gd = (np.float32(0.5) + np.sign(x) * (np.float32(0.5) - ER) + x * Ef * np.float32(1 / np.sqrt(2 * np.pi))).astype(np.float32)
    print(f"[{name}] fp32 eval: max abs err gelu {np.abs(g - exact).max():.3e}  gelu' {np.abs(gd - exact_d).max():.3e}")
