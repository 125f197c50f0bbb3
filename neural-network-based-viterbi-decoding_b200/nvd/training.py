"""Training loops of the drop-in modules (the step BEFORE the hot path).

The reference trains inside a TensorFlow-1.x session (/root/reference/src/fnn_viterbi.py:549-603,
fnn_viterbi_mixed_snr.py:722-785, fnn_viterbi_bitwise.py:439-497, rnn_viterbi.py:311-316).  The feed-forward
decoders train through libnvd's own training step (nvd_train_*, csrc/k_train.cu); the dropout variant and the
LSTM's MSE objective are restated with torch autograd (plumbing: CUDA when available, else CPU).  Either way the
trained parameters are handed to libnvd, whose kernels do every forward pass that produces a reported decision or BER."""
from __future__ import annotations

import numpy as np


def _torch():
    import torch
    return torch, torch.device("cuda" if torch.cuda.is_available() else "cpu")


def train_mlp(model, x, y_onehot, *, acts, loss="softmax_ce", beta=0.0, l2_layers=(0, 1, 2), lr=1e-4, epochs=1,
              batch_size=256, dropout_keep=None, display_step=1, log=print, epoch_callback=None, backend=None):
    """Train `model` in place; returns the per-epoch costs.  backend "native" (default when no dropout is asked for):
    libnvd's own training step, nvd_train_begin / nvd_train_step / nvd_train_end (csrc/k_train.cu) — forward, cost,
    backward, L2 and TensorFlow's Adam as CUDA kernels on the process-wide Simulator; backend "torch": the autograd
    restatement below (needed for dropout; CPU-capable, used by the tools that build weight fixtures)."""
    if backend is None:
        backend = "torch" if dropout_keep is not None else "native"
    if backend == "native":
        if dropout_keep is not None:
            raise ValueError("the native training step has no dropout; use backend='torch'")
        return train_mlp_native(model, x, y_onehot, acts=acts, loss=loss, beta=beta, l2_layers=l2_layers, lr=lr, epochs=epochs,
                                batch_size=batch_size, display_step=display_step, log=log, epoch_callback=epoch_callback)
    return train_mlp_torch(model, x, y_onehot, acts=acts, loss=loss, beta=beta, l2_layers=l2_layers, lr=lr, epochs=epochs,
                           batch_size=batch_size, dropout_keep=dropout_keep, display_step=display_step, log=log,
                           epoch_callback=epoch_callback)


def train_mlp_native(model, x, y_onehot, *, acts, loss="softmax_ce", beta=0.0, l2_layers=(0, 1, 2), lr=1e-4, epochs=1,
                     batch_size=256, display_step=1, log=print, epoch_callback=None, sim=None):
    """Mini-batch training on the GPU through the C ABI, in the reference's fixed batch order (fnn_viterbi.py:579-585).
    The whole training set and its one-hot targets are uploaded once; every step is three GEMMs per layer plus
    elementwise kernels; the cost of every step is read back (4 bytes) for the per-epoch average the reference prints."""
    import ctypes as C

    from . import _lib, runtime
    from .weights import MlpModel
    sim = sim or runtime.get_sim()
    m = MlpModel([np.ascontiguousarray(W, np.float32) for W in model.Ws], [np.ascontiguousarray(b, np.float32) for b in model.bs],
                 [int(a) for a in acts])
    h = sim.load_mlp(m)
    cfg = _lib.TrainConfig()
    cfg.struct_bytes = C.sizeof(_lib.TrainConfig)
    cfg.loss = _lib.LOSSES[loss]
    cfg.learning_rate, cfg.beta = float(lr), float(beta)
    cfg.l2_layer_mask = sum(1 << int(l) for l in l2_layers if int(l) < len(m.Ws))
    cfg.beta1, cfg.beta2, cfg.epsilon = 0.9, 0.999, 1e-8
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y_onehot, dtype=np.float32)
    n, n_in, n_out = x.shape[0], x.shape[1], y.shape[1]
    total_batch = int(n / batch_size)
    tr = C.c_int(-1)
    _lib.check(sim.lib.nvd_train_begin(sim.ctx, h, C.byref(cfg), batch_size, C.byref(tr)))
    d_x, d_y = sim.to_device(x[: total_batch * batch_size]), sim.to_device(y[: total_batch * batch_size])
    costs, cost = [], C.c_float(0.0)

    def read_back():
        Wp = (C.c_void_p * len(m.Ws))(*[W.ctypes.data for W in m.Ws])
        bp = (C.c_void_p * len(m.bs))(*[b.ctypes.data for b in m.bs])
        return Wp, bp

    try:
        for epoch in range(epochs):
            avg = 0.0
            for i in range(total_batch):
                _lib.check(sim.lib.nvd_train_step(sim.ctx, tr.value, d_x.ptr + i * batch_size * n_in * 4,
                                                  d_y.ptr + i * batch_size * n_out * 4, batch_size, C.byref(cost)))
                avg += cost.value / max(total_batch, 1)
            costs.append(avg)
            if display_step and epoch % display_step == 0:
                log(f"Epoch: {epoch + 1:04d} | cost = {avg:.9f}")
            if epoch_callback is not None:
                for l in range(len(m.Ws)):      # the callback evaluates / checkpoints the current parameters
                    _lib.check(sim.lib.nvd_memcpy_d2h(sim.ctx, m.Ws[l].ctypes.data, _device_param(sim, h, l, True), m.Ws[l].nbytes))
                    _lib.check(sim.lib.nvd_memcpy_d2h(sim.ctx, m.bs[l].ctypes.data, _device_param(sim, h, l, False), m.bs[l].nbytes))
                sim.sync()
                model.Ws, model.bs = [W.copy() for W in m.Ws], [b.copy() for b in m.bs]
                if epoch_callback(epoch, avg, model):
                    break
    finally:
        Wp, bp = read_back()
        _lib.check(sim.lib.nvd_train_end(sim.ctx, tr.value, Wp, bp))
        for b in (d_x, d_y):
            sim.free(b)
        _lib.check(sim.lib.nvd_mlp_free(sim.ctx, h))
    model.Ws, model.bs = [W.copy() for W in m.Ws], [b.copy() for b in m.bs]
    return costs


def _device_param(sim, h, layer, weight):
    import ctypes as C

    from . import _lib
    p = C.c_void_p()
    _lib.check(sim.lib.nvd_mlp_param(sim.ctx, h, layer, int(weight), C.byref(p)))
    return p


def train_mlp_torch(model, x, y_onehot, *, acts, loss="softmax_ce", beta=0.0, l2_layers=(0, 1, 2), lr=1e-4, epochs=1,
                    batch_size=256, dropout_keep=None, display_step=1, log=print, epoch_callback=None):
    """Mini-batch Adam over (x, y_onehot) in the reference's fixed batch order (no shuffling:
    fnn_viterbi.py:579-585).  cost = mean CE(out, y) + beta * sum_l l2_loss(W_l) with
    tf.nn.l2_loss(W) = sum(W^2)/2.  Updates `model` in place and returns the per-epoch costs.
    ONE optimizer lives across all epochs, as the reference's AdamOptimizer does (fnn_viterbi.py:562-603,
    production_wrapper.py:141-190).  `epoch_callback(epoch, cost, model)` runs after every epoch with `model`
    holding the current parameters (evaluation, checkpoints); returning True stops the training (early stopping)."""
    torch, dev = _torch()
    Ws = [torch.nn.Parameter(torch.from_numpy(np.array(W, dtype=np.float32)).to(dev)) for W in model.Ws]
    bs = [torch.nn.Parameter(torch.from_numpy(np.array(b, dtype=np.float32)).to(dev)) for b in model.bs]
    opt = torch.optim.Adam(Ws + bs, lr=lr)
    X = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
    Y = torch.from_numpy(np.ascontiguousarray(y_onehot, dtype=np.float32)).to(dev)
    n = X.shape[0]
    total_batch = int(n / batch_size)
    costs = []

    def act(h, a):
        return torch.tanh(h) if a == 1 else torch.relu(h) if a == 2 else h

    for epoch in range(epochs):
        avg = 0.0
        for i in range(total_batch):
            xb, yb = X[i * batch_size:(i + 1) * batch_size], Y[i * batch_size:(i + 1) * batch_size]
            h = xb
            for li, (W, b, a) in enumerate(zip(Ws, bs, acts)):
                h = act(h @ W + b, a)
                if dropout_keep is not None and li < len(Ws) - 1:
                    h = torch.nn.functional.dropout(h, p=1.0 - dropout_keep, training=True)
            if loss == "softmax_ce":
                ce = -(yb * torch.log_softmax(h, dim=1)).sum(1).mean()
            elif loss == "sigmoid_ce":
                ce = torch.nn.functional.binary_cross_entropy_with_logits(h, yb)
            else:
                raise ValueError(loss)
            reg = sum(0.5 * (Ws[l] ** 2).sum() for l in l2_layers if l < len(Ws))
            cost = ce + beta * reg
            opt.zero_grad(set_to_none=True)
            cost.backward()
            opt.step()
            avg += float(cost.detach()) / max(total_batch, 1)
        costs.append(avg)
        if display_step and epoch % display_step == 0:
            log(f"Epoch: {epoch + 1:04d} | cost = {avg:.9f}")
        if epoch_callback is not None:
            model.Ws = [W.detach().cpu().numpy() for W in Ws]
            model.bs = [b.detach().cpu().numpy() for b in bs]
            if epoch_callback(epoch, avg, model):
                break
    model.Ws = [W.detach().cpu().numpy() for W in Ws]
    model.bs = [b.detach().cpu().numpy() for b in bs]
    return costs


def train_lstm(params, enc_in, dec_in, dec_tgt, *, lr=0.01, batches=1, batch_size=100, log=print, log_every=0):
    """MSE(decoder logits, targets) with Adam (rnn_viterbi.py:311-316); TF LSTMCell equations."""
    torch, dev = _torch()
    P = {k: torch.nn.Parameter(torch.from_numpy(np.array(v, dtype=np.float32)).to(dev)) for k, v in params.items()}
    opt = torch.optim.Adam(list(P.values()), lr=lr)
    E = torch.from_numpy(np.ascontiguousarray(enc_in, dtype=np.float32)).to(dev)
    Din = torch.from_numpy(np.ascontiguousarray(dec_in, dtype=np.float32)).to(dev)
    Dt = torch.from_numpy(np.ascontiguousarray(dec_tgt, dtype=np.float32)).to(dev)
    n = E.shape[0]
    per_epoch = max(int(n / batch_size), 1)
    H = P["enc_b"].shape[0] // 4
    history = []

    def cell(x, c, h, W, b):
        z = torch.cat([x, h], 1) @ W + b
        i, j, f, o = z.chunk(4, 1)
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        return c, torch.sigmoid(o) * torch.tanh(c)

    for bi in range(batches):
        s = (bi % per_epoch) * batch_size
        e, di, dt = E[s:s + batch_size], Din[s:s + batch_size], Dt[s:s + batch_size]
        B = e.shape[0]
        c = torch.zeros(B, H, device=dev)
        h = torch.zeros(B, H, device=dev)
        for t in range(e.shape[1]):
            c, h = cell(e[:, t:t + 1], c, h, P["enc_W"], P["enc_b"])
        outs = []
        for t in range(di.shape[1]):
            c, h = cell(di[:, t:t + 1], c, h, P["dec_W"], P["dec_b"])
            outs.append(h @ P["lin_w"].reshape(H, 1) + P["lin_b"])
        loss = ((torch.cat(outs, 1) - dt) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        history.append(float(loss.detach()))
        if log_every and bi % log_every == 0:
            log(f"Epoch {bi // per_epoch + 1} | Loss: {history[-1]:.6f}")
    return {k: v.detach().cpu().numpy() for k, v in P.items()}, history
