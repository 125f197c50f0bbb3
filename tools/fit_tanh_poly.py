#!/usr/bin/env python
"""Coefficients of the FMA-pipe tanh of the tensor-core decoder (csrc/k_mlp_tc.cu, NVD_POLY_C*):
tanh(x) ~ sign(x) * u * Q(u), u = min(|x| / R, 1), Q of degree 5, fitted by Lawson's iteratively
re-weighted least squares (approaches the minimax fit) on u in (0, 1]; the error is then evaluated
with the kernel's fp32 arithmetic (one rounding per fused multiply-add) over |x| <= 9."""
import numpy as np


def lawson(V, y, iters=400):
    w = np.ones(len(y))
    for _ in range(iters):
        sw = np.sqrt(w)
        coef = np.linalg.lstsq(V * sw[:, None], y * sw, rcond=None)[0]
        err = np.abs(V @ coef - y)
        w = w * (err / err.max() + 1e-3)
        w /= w.sum()
    return coef, err.max()


def f32(v):
    return np.float32(v).astype(np.float64)


def evaluate(coef, R, x):
    u = np.minimum(f32(np.abs(f32(x)) * f32(1 / R)), 1.0)
    cs = [f32(k) for k in coef]
    p = np.full_like(u, cs[-1])
    for k in cs[-2::-1]:
        p = f32(p * u + k)
    return np.sign(x) * f32(p * u)


if __name__ == "__main__":
    x = np.linspace(-9, 9, 900001)
    for R in (3.25, 3.5, 3.75):
        u = np.linspace(0, 1, 8001)[1:]
        V = np.stack([u ** (i + 1) for i in range(6)], 1)
        coef, fit = lawson(V, np.tanh(R * u))
        err = np.abs(evaluate(coef, R, x) - np.tanh(x)).max()
        print(f"R = {R}: fit error {fit:.3e}, fp32 max |error| {err:.3e}, value at saturation {sum(f32(k) for k in coef):.6f}")
        print("   ", ", ".join(f"{k:.9e}f" for k in coef))
