"""Minimax fit behind csrc/common.cuh::gelu_sig2:  gelu(x) ~= x * sigmoid(x (c1 + c3 x^2 + c5 x^4)).

    python tools/fit_gelu_sigmoid.py        # prints the coefficients, the constants the kernel uses and the fp32 error
"""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

x = np.linspace(-10, 10, 400001)
gelu = x * 0.5 * (1 + erf(x / np.sqrt(2)))


def err(c):
    x2 = x * x
    p = ((c[2] * x2 + c[1]) * x2 + c[0]) * x
    return np.abs(x / (1 + np.exp(-p)) - gelu).max()


c = np.array([1.5957, 0.0729, -0.0004])
for _ in range(10):
    c = minimize(err, c, method="Nelder-Mead", options=dict(xatol=1e-12, fatol=1e-14, maxiter=40000, maxfev=80000)).x
print("c1, c3, c5 =", list(c), " max |gelu error| (fp64) =", err(c))
L = np.float32(-1.4426950408889634)
C1, C3, C5 = [np.float32(v) * L for v in c]
print("kernel constants (times -log2 e):", float(C1), float(C3), float(C5))
xs = np.linspace(-30, 30, 600001).astype(np.float32)
xc = np.clip(xs, np.float32(-10), np.float32(10))
x2 = xc * xc
p = (x2 * C5 + C3).astype(np.float32)
p = (p * x2 + C1).astype(np.float32)
a = (p * xc).astype(np.float32)
r = (np.float32(1) / (np.exp2(a.astype(np.float64)).astype(np.float32) + np.float32(1))).astype(np.float32)
ref = xs.astype(np.float64) * 0.5 * (1 + erf(xs.astype(np.float64) / np.sqrt(2)))
print("fp32 evaluation with the +-10 clamp, x in [-30, 30]: max |error| =", np.abs((xs * r).astype(np.float32) - ref).max())
