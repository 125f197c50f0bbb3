"""Oracle: the training step of the feed-forward decoders in float64 (TEST INFRASTRUCTURE).

Restates what the reference's TensorFlow-1.x session computes per mini-batch:
  cost   mean softmax cross-entropy on the activated outputs (tf.nn.softmax_cross_entropy_with_logits,
         /root/reference/src/fnn_viterbi.py:549-552) or mean sigmoid cross-entropy over every element
         (fnn_viterbi_bitwise.py:439-441), plus beta * sum_l tf.nn.l2_loss(W_l) = beta * sum(W^2)/2 (:554-559)
  update tf.train.AdamOptimizer(learning_rate) (:562-564), TensorFlow's formulation:
         m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  lr_t = lr sqrt(1-b2^t)/(1-b1^t);  w -= lr_t m / (sqrt(v) + eps)
Gradients are written out by hand (no autograd) so that the CUDA trainer is checked against independent arithmetic.
"""
from __future__ import annotations

import numpy as np

ACT_IDENTITY, ACT_TANH, ACT_RELU = 0, 1, 2


def _act(z, a):
    return np.tanh(z) if a == ACT_TANH else np.maximum(z, 0) if a == ACT_RELU else z


def cost_and_grads(Ws, bs, acts, x, y, loss="softmax_ce", beta=0.0, l2_layers=()):
    """-> (cost, dWs, dbs) in float64."""
    A = [np.asarray(x, dtype=np.float64)]
    for W, b, a in zip(Ws, bs, acts):
        A.append(_act(A[-1] @ np.asarray(W, np.float64) + np.asarray(b, np.float64), a))
    out, y = A[-1], np.asarray(y, dtype=np.float64)
    B = out.shape[0]
    if loss == "softmax_ce":
        z = out - out.max(axis=1, keepdims=True)
        lse = np.log(np.exp(z).sum(axis=1, keepdims=True))
        cost = float((-(y * (z - lse)).sum(axis=1)).mean())
        d = (np.exp(z - lse) * y.sum(axis=1, keepdims=True) - y) / B
    elif loss == "sigmoid_ce":
        cost = float((np.maximum(out, 0) - out * y + np.log1p(np.exp(-np.abs(out)))).mean())
        d = (1.0 / (1.0 + np.exp(-out)) - y) / out.size
    else:
        raise ValueError(loss)
    dWs, dbs = [None] * len(Ws), [None] * len(Ws)
    for l in range(len(Ws) - 1, -1, -1):
        a = A[l + 1]
        if acts[l] == ACT_TANH:
            d = d * (1.0 - a * a)
        elif acts[l] == ACT_RELU:
            d = d * (a > 0)
        dWs[l] = A[l].T @ d
        dbs[l] = d.sum(axis=0)
        if l in l2_layers:
            dWs[l] = dWs[l] + beta * np.asarray(Ws[l], np.float64)
            cost += 0.5 * beta * float((np.asarray(Ws[l], np.float64) ** 2).sum())
        d = d @ np.asarray(Ws[l], np.float64).T
    return cost, dWs, dbs


class Adam:
    """TensorFlow-1.x AdamOptimizer arithmetic."""

    def __init__(self, params, lr=1e-4, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, epsilon, 0
        self.m = [np.zeros_like(np.asarray(p, np.float64)) for p in params]
        self.v = [np.zeros_like(np.asarray(p, np.float64)) for p in params]

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            out.append(np.asarray(p, np.float64) - lr_t * self.m[i] / (np.sqrt(self.v[i]) + self.eps))
        return out


def train(Ws, bs, acts, x, y, batch_size, steps, loss="softmax_ce", beta=0.0, l2_layers=(), lr=1e-4):
    """`steps` mini-batches in the reference's fixed order (fnn_viterbi.py:579-585) -> (Ws, bs, costs)."""
    Ws = [np.asarray(W, np.float64) for W in Ws]
    bs = [np.asarray(b, np.float64) for b in bs]
    opt = Adam(Ws + bs, lr=lr)
    n_batches = len(x) // batch_size
    costs = []
    for s in range(steps):
        i = s % n_batches
        xb, yb = x[i * batch_size:(i + 1) * batch_size], y[i * batch_size:(i + 1) * batch_size]
        c, dW, db = cost_and_grads(Ws, bs, acts, xb, yb, loss, beta, l2_layers)
        new = opt.step(Ws + bs, dW + db)
        Ws, bs = new[:len(Ws)], new[len(Ws):]
        costs.append(c)
    return Ws, bs, costs
