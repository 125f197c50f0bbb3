"""Training loops of the drop-in modules (the step BEFORE the hot path).

The reference trains inside a TensorFlow-1.x session (/root/reference/src/fnn_viterbi.py:549-603,
fnn_viterbi_mixed_snr.py:722-785, fnn_viterbi_bitwise.py:439-497, rnn_viterbi.py:311-316).  Here the
same objectives are restated with torch autograd (plumbing: CUDA when available, else CPU); the
trained parameters are handed to libnvd, whose kernels do every forward pass that produces a
reported decision or BER.  Training is not part of the accelerated path and is kept simple."""
from __future__ import annotations

import numpy as np


def _torch():
    import torch
    return torch, torch.device("cuda" if torch.cuda.is_available() else "cpu")


def train_mlp(model, x, y_onehot, *, acts, loss="softmax_ce", beta=0.0, l2_layers=(0, 1, 2), lr=1e-4, epochs=1,
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
