"""Alternative GELU form: gelu(x) = relu(x) - t * 2^P(t), t = min(|x|, T), P ~ log2(Phi(-t)).
Fit P by weighted minimax (weight = h(t) ln2, the sensitivity of h = t*2^P to P), report the
fp32-evaluated error of gelu against exact erf GELU."""
import sys
import numpy as np
from scipy.optimize import linprog
from scipy.special import log_ndtr, erf

deg = int(sys.argv[1]); T = float(sys.argv[2])
t = np.linspace(0, T, 6001)
P = log_ndtr(-t) / np.log(2)
h = t * np.exp2(P)
w = np.maximum(h, 1e-9) * np.log(2)
v = 2 * t / T - 1
A = np.polynomial.chebyshev.chebvander(v, deg)
n = deg + 1
c = np.zeros(n + 1); c[-1] = 1
Aub = np.vstack([np.hstack([w[:, None] * A, -np.ones((len(t), 1))]), np.hstack([-w[:, None] * A, -np.ones((len(t), 1))])])
bub = np.concatenate([w * P, -w * P])
res = linprog(c, A_ub=Aub, b_ub=bub, bounds=[(None, None)] * n + [(0, None)], method="highs")
cheb = res.x[:n]
mono_v = np.polynomial.chebyshev.cheb2poly(cheb)
mono_t = np.polynomial.Polynomial(mono_v)(np.polynomial.Polynomial([-1.0, 2.0 / T])).coef
print("deg", deg, "T", T, "lin. minimax", res.x[-1])
print("mono_t:", ", ".join(f"{c:.9e}f" for c in mono_t))
x = np.linspace(-9, 9, 600001).astype(np.float32)
tt = np.minimum(np.abs(x), np.float32(T)).astype(np.float32)
acc = np.full_like(tt, np.float32(mono_t[-1]))
for cc in mono_t[-2::-1]:
    acc = (acc * tt + np.float32(cc)).astype(np.float32)
E = np.exp2(acc.astype(np.float64)).astype(np.float32) * (1 + 2.0 ** -22)   # ex2.approx worst case
g = (np.maximum(x, 0) - tt * E).astype(np.float32)
xd = x.astype(np.float64)
exact = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
print("fp32 eval max abs err gelu: %.3e" % np.abs(g - exact).max())
