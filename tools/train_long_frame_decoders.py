#!/usr/bin/env python
"""Train the two long-frame decoders on the CPU and write `.npz` fixtures under
neural-network-based-viterbi-decoding_b200/weights/ so that their Monte-Carlo BER curves are meaningful:

  bitwise : the wired per-bit predictor 1036-30-2, ReLU on both layers, sigmoid cross-entropy on the ReLU outputs
            (/root/reference/src/fnn_viterbi_bitwise.py:388-397, 439-448), random SNR per frame from -2..8 dB (:280-285),
            int32-truncated channel outputs (:274-285), start state 0, bit_position 0.
  rnn     : LSTM(20) encoder + LSTM(20) decoder + linear, MSE on the decoder logits (/root/reference/src/rnn_viterbi.py:311-316),
            teacher-forced inputs (:362-371), int32-truncated channel outputs, L = 100.

Training is OFF the hot path; step counts and learning-rate schedules are ours.
Usage: python tools/train_long_frame_decoders.py bitwise|rnn [--steps N]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from oracle import channel, encoders, lstm as olstm, mlp as omlp  # noqa: E402


def noisy_frames(rng, info, prefix, sigmas, trunc=True):
    B = info.shape[0]
    full = np.concatenate([np.tile(prefix, (B, 1)), info], axis=1) if len(prefix) else info
    coded = encoders.conv_encode(full, encoders.CODE_133171)
    sig = sigmas[rng.integers(0, len(sigmas), B)][:, None]
    y = (2.0 * coded - 1.0) + sig * rng.standard_normal(coded.shape)
    return (np.trunc(y) if trunc else y).astype(np.float32)


def train_bitwise(args):
    from nvd import weights
    rng = np.random.default_rng(20170711)
    torch.manual_seed(20170711)
    L, K = 512, 6
    prefix = encoders.extract_state_bits(0, K).astype(np.uint8)
    sigmas = channel.snr_db_to_sigma(np.arange(-2, 9))
    Ws0, _ = omlp.init_fnn([2 * (L + K), 30, 2], seed=5)
    W1 = torch.nn.Parameter(torch.from_numpy(Ws0[0].copy()))
    W2 = torch.nn.Parameter(torch.from_numpy((Ws0[1] * 0 + np.abs(rng.standard_normal((30, 2)) * 0.3)).astype(np.float32)))
    b1 = torch.nn.Parameter(torch.full((30,), 0.1))
    b2 = torch.nn.Parameter(torch.full((2,), 0.1))
    opt = torch.optim.Adam([W1, W2, b1, b2], lr=3e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=args.steps, eta_min=3e-5)
    for step in range(args.steps):
        info = rng.integers(0, 2, (args.batch, L)).astype(np.uint8)
        x = torch.from_numpy(noisy_frames(rng, info, prefix, sigmas))
        lab = torch.from_numpy(info[:, 0].astype(np.int64))
        out = torch.relu(torch.relu(x @ W1 + b1) @ W2 + b2)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.nn.functional.one_hot(lab, 2).float())
        loss = loss + 1e-4 * 0.5 * (W1 ** 2).sum()
        opt.zero_grad(); loss.backward(); opt.step(); sched.step()
        if step % 500 == 0 or step == args.steps - 1:
            acc = (out.argmax(1) == lab).float().mean().item()
            print(f"step {step:6d} loss {loss.item():.4f} batch-acc {acc:.4f}", flush=True)
    m = weights.MlpModel([W1.detach().numpy(), W2.detach().numpy()], [b1.detach().numpy(), b2.detach().numpy()], [2, 2])
    for db in (0, 2, 4, 6):
        info = rng.integers(0, 2, (1 << 14, L)).astype(np.uint8)
        y = noisy_frames(rng, info, prefix, channel.snr_db_to_sigma(np.array([db])))
        pred = omlp.argmax_first(omlp.forward(y, m.Ws, m.bs, m.acts, dtype=np.float32))
        print(f"  {db} dB: bit error rate {np.mean(pred != info[:, 0]):.4e}")
    out = os.path.join(weights.WEIGHTS_DIR, "bitwise_133171_L512_bit0.npz")
    weights.save_mlp(out, m)
    print("wrote", out)


def train_rnn(args):
    from nvd import weights
    rng = np.random.default_rng(20170712)
    torch.manual_seed(20170712)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    L, H = 100, 20
    sigmas = channel.snr_db_to_sigma(np.arange(0, 9))
    p0 = olstm.init_params(H, seed=12)
    P = {k: torch.nn.Parameter(torch.from_numpy(np.asarray(v, np.float32).copy())) for k, v in p0.items()}
    opt = torch.optim.Adam(list(P.values()), lr=1e-2)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=args.steps, eta_min=5e-4)

    def cell(x, c, h, W, b):
        z = torch.cat([x, h], 1) @ W + b
        i, j, f, o = z.chunk(4, 1)
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        return c, torch.sigmoid(o) * torch.tanh(c)

    for step in range(args.steps):
        info = rng.integers(0, 2, (args.batch, L)).astype(np.uint8)
        y = noisy_frames(rng, info, np.zeros(0, np.uint8), sigmas)
        _, dec_in, dec_tgt = olstm.prepare_batch(2.0 * info - 1.0, y, L)
        e, di, dt = (torch.from_numpy(np.asarray(a, np.float32)) for a in (y, dec_in, dec_tgt))
        B = e.shape[0]
        c = torch.zeros(B, H); h = torch.zeros(B, H)
        for t in range(e.shape[1]):
            c, h = cell(e[:, t:t + 1], c, h, P["enc_W"], P["enc_b"])
        outs = []
        for t in range(di.shape[1]):
            c, h = cell(di[:, t:t + 1], c, h, P["dec_W"], P["dec_b"])
            outs.append(h @ P["lin_w"].reshape(H, 1) + P["lin_b"])
        logits = torch.cat(outs, 1)
        loss = ((logits - dt) ** 2).mean()
        opt.zero_grad(); loss.backward()
        torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
        opt.step(); sched.step()
        if step % 50 == 0 or step == args.steps - 1:
            ber = ((logits[:, :L] > 0).numpy() != (info > 0)).mean()
            print(f"step {step:5d} loss {loss.item():.4f} batch BER {ber:.4f}", flush=True)
        if step % 200 == 199 or step == args.steps - 1:
            out = os.path.join(weights.WEIGHTS_DIR, "rnn_133171_L100.npz")
            np.savez(out, **{k: v.detach().numpy() for k, v in P.items()})
    print("wrote", out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["bitwise", "rnn"])
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    a = ap.parse_args()
    if a.which == "bitwise":
        a.steps, a.batch = a.steps or 6000, a.batch or 512
        train_bitwise(a)
    else:
        a.steps, a.batch = a.steps or 1200, a.batch or 128
        train_rnn(a)
