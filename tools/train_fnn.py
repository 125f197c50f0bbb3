#!/usr/bin/env python
"""Train the reference FNN decoder (14-128-64-32-128, tanh x4) on the CPU and write the `.npz`
fixture neural-network-based-viterbi-decoding_b200/weights/fnn_133171_L7.npz.

Restates the reference's training objective (/root/reference/src/fnn_viterbi.py:549-564 and
fnn_viterbi_mixed_snr.py:722-739): softmax cross-entropy on the tanh OUTPUTS used as logits,
plus beta * (l2_loss(h1) + l2_loss(h2) + l2_loss(h3)) with tf.nn.l2_loss = sum(w^2)/2, Adam.
Data follow get_dataset's mixed-SNR training rows (fnn_viterbi_mixed_snr.py:430-449: every code
word at every SNR of -2..9 dB) but are drawn fresh each step instead of from a fixed 128k-row
set, and the step count / learning-rate schedule are ours (the reference's 1e-4 x 300 epochs
takes hours on a CPU).  Training is OFF the hot path; this script exists so
that BER curves are meaningful.  Usage: python tools/train_fnn.py [--code 133171|57] [--steps N]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))

from oracle import encoders, channel, mlp as omlp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="133171")
    ap.add_argument("--steps", type=int, default=40000)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--beta", type=float, default=1e-4)
    ap.add_argument("--logit-scale", type=float, default=1.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    torch.manual_seed(20170710)
    rng = np.random.default_rng(20170710)
    L = 7
    code = encoders.CODE_133171 if args.code == "133171" else encoders.CODE_57_REF
    book = torch.from_numpy(2.0 * encoders.conv_encode(encoders.generate_class_bits(L), code).astype(np.float32) - 1.0)
    sigmas = torch.from_numpy(channel.snr_db_to_sigma(np.arange(-2, 10)).astype(np.float32))

    dims = [14, 128, 64, 32, 128]
    Ws0, _ = omlp.init_fnn(dims, seed=20170710)
    Ws = [torch.nn.Parameter(torch.from_numpy(W.copy())) for W in Ws0]
    bs = [torch.nn.Parameter(torch.zeros(n)) for n in dims[1:]]
    opt = torch.optim.Adam(Ws + bs, lr=2e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=args.steps, eta_min=2e-5)

    def forward(x):
        h = x
        for W, b in zip(Ws, bs):
            h = torch.tanh(h @ W + b)
        return h

    for step in range(args.steps):
        cls = torch.randint(0, 128, (args.batch,))
        sig = sigmas[torch.randint(0, len(sigmas), (args.batch,))][:, None]
        x = book[cls] + sig * torch.randn(args.batch, 14)
        out = forward(x)
        loss = torch.nn.functional.cross_entropy(out * args.logit_scale, cls)
        reg = 0.5 * sum((W ** 2).sum() for W in Ws[:3])
        total = loss + args.beta * reg
        opt.zero_grad()
        total.backward()
        opt.step()
        sched.step()
        if step % 2000 == 0 or step == args.steps - 1:
            with torch.no_grad():
                acc = (out.argmax(1) == cls).float().mean().item()
            print(f"step {step:6d} loss {loss.item():.4f} batch-acc {acc:.4f}", flush=True)

    Wn = [W.detach().numpy().astype(np.float32) for W in Ws]
    bn = [b.detach().numpy().astype(np.float32) for b in bs]
    # quick BER table with numpy noise
    cb = encoders.generate_class_bits(L)
    for db in (0, 2, 4, 6):
        n = 1 << 18
        cls = rng.integers(0, 128, n)
        y = book.numpy()[cls] + float(channel.snr_db_to_sigma(db)) * rng.standard_normal((n, 14)).astype(np.float32)
        pred = omlp.argmax_first(omlp.forward(y, Wn, bn, [1, 1, 1, 1], dtype=np.float32))
        ber = np.mean(cb[pred] != cb[cls])
        print(f"  {db} dB: BER {ber:.3e} FER {np.mean(pred != cls):.3e}")

    from nvd import weights
    out = args.out or (weights.DEFAULT_FNN if args.code == "133171" else os.path.join(weights.WEIGHTS_DIR, "fnn_57_L7.npz"))
    weights.save_mlp(out, weights.MlpModel(Wn, bn, [1, 1, 1, 1]))
    print("wrote", out)


if __name__ == "__main__":
    main()
