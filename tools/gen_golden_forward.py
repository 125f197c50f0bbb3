#!/usr/bin/env python
"""Generate tests/golden/reference_forward.npz by RUNNING the reference's own decoder graph
functions — multilayer_perceptron (src/fnn_viterbi.py:471-510, src/fnn_viterbi_mixed_snr.py:573-627),
build_network (src/fnn_viterbi_bitwise.py:326-399, src/fnn_viterbi_seq_predict.py:343-414) and
build_seq2seq_model (src/rnn_viterbi.py:235-325) — unmodified, from /root/reference, under the
numpy-backed `tensorflow` stand-in of tools/tf_numpy_shim (TensorFlow is not installable here).

Run in the build container only (the GPU box has no /root/reference):
    python tools/gen_golden_forward.py
Inputs are noisy code words made by the reference's own encoder / modulate_awgn with numpy's legacy
RNG seeded; parameters are the arrays tests/golden/_weights.py yields for the reference's variable
names, so the fixture stores inputs and outputs only.  Nothing is copied from the reference.
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"
OUT = os.path.join(HERE, "..", "tests", "golden", "reference_forward.npz")


def _stub_matplotlib():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))


def _load_rnn(tf):
    """rnn_viterbi.py raises NameError at import (`Optional`, :35): exec its defs from the AST."""
    tree = ast.parse(open(os.path.join(REF, "rnn_viterbi.py")).read())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))
            and n.name in ("RNNDecoderConfig", "encode_133171", "modulate_awgn", "prepare_batch", "build_seq2seq_model")]
    mod = types.ModuleType("rnn_viterbi_ast")
    from typing import Dict, Optional, Tuple
    mod.__dict__.update(np=np, tf=tf, Dict=Dict, Optional=Optional, Tuple=Tuple)
    exec(compile(ast.Module(body=keep, type_ignores=[]), "rnn_viterbi.py", "exec"), mod.__dict__)
    return mod


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "tf_numpy_shim"))
    _stub_matplotlib()
    import tensorflow as tf          # the numpy stand-in
    import fnn_viterbi as fv
    import fnn_viterbi_mixed_snr as fm
    import fnn_viterbi_bitwise as fb
    import fnn_viterbi_seq_predict as fs
    rv = _load_rnn(tf)

    g = {}
    BIAS = 0.05

    # ---- fnn_viterbi: initialize_weights + multilayer_perceptron + tf.argmax (:413-510, :606-607) ----
    cfg = fv.DecoderConfig()
    cb = fv.generate_class_bits(cfg.block_length)
    np.random.seed(101)
    labels = np.arange(512) % 128
    x = np.zeros((512, cfg.n_input), dtype=np.float32)
    for i, c in enumerate(labels):
        x[i, :] = fv.modulate_awgn(fv.encode_133171(cb[c]), cfg.sigma)       # float32 destination (:373,384)
    tf.reset(seed=11, bias_scale=BIAS)
    W, b = fv.initialize_weights(cfg)
    out = fv.multilayer_perceptron(x, W, b)
    g["fnn_seed"], g["fnn_bias_scale"] = np.int64(11), np.float64(BIAS)
    g["fnn_x"], g["fnn_labels"], g["fnn_out"] = x, labels, out
    g["fnn_pred"] = tf.argmax(out, 1)
    assert list(tf.VARIABLES) == ["weights_h1", "weights_h2", "weights_h3", "weights_out", "bias_h1", "bias_h2", "bias_h3", "bias_out"]

    # ---- fnn_viterbi_mixed_snr: both graph functions (:573-627), at every sigma of its table ----------
    mcfg = fm.MixedSNRDecoderConfig()
    np.random.seed(102)
    xm = np.zeros((len(mcfg.sigma_vec) * 32, mcfg.n_input), dtype=np.float32)
    lm = np.zeros(len(xm), dtype=np.int64)
    for i in range(len(xm)):
        lm[i] = (i * 37) % 128
        xm[i, :] = fm.modulate_awgn(fm.encode_133171(cb[lm[i]]), mcfg.sigma_vec[i // 32])
    tf.reset(seed=12, bias_scale=BIAS)
    Wm, bm = fm.initialize_weights(mcfg)
    om = fm.multilayer_perceptron(xm, Wm, bm)
    od = fm.multilayer_perceptron_with_dropout(xm, Wm, bm, 1.0)
    assert np.array_equal(om, od)
    g["mixed_seed"] = np.int64(12)
    g["mixed_x"], g["mixed_labels"], g["mixed_out"], g["mixed_pred"] = xm, lm, om, tf.argmax(om, 1)

    # ---- fnn_viterbi_bitwise: build_network (:326-399), int32-truncated inputs (:309-317), decision :504 ----
    bcfg = fb.BitwiseDecoderConfig()
    np.random.seed(103)
    nb = 96
    info_b = np.random.randint(0, 2, size=(nb, bcfg.block_length))
    xb = np.zeros((nb, bcfg.n_input), dtype=np.int32)
    for i in range(nb):
        coded = fb.encode_133171_with_state(info_b[i], bcfg.starting_state)
        xb[i, :] = fb.modulate_awgn(coded, bcfg.sigma_vec[2 + i % 9])        # int32 destination -> truncation
    tf.reset(seed=13, bias_scale=BIAS)
    ob, Wb, bb = fb.build_network(xb.astype(np.float32), bcfg)
    g["bitwise_seed"] = np.int64(13)
    g["bitwise_info"], g["bitwise_x"], g["bitwise_out"], g["bitwise_pred"] = info_b.astype(np.uint8), xb, ob, tf.argmax(ob, 1)
    g["bitwise_wout_shape"] = np.array(Wb["out"].shape)

    # ---- fnn_viterbi_seq_predict: build_network (:343-414) on stride-2 windows (:293-336) ---------------
    scfg = fs.SequentialDecoderConfig()
    np.random.seed(104)
    n_info = 400
    info_s = np.random.randint(0, 2, size=n_info)
    stream = np.zeros((1, 2 * n_info), dtype=np.int32)
    stream[0, :] = fs.modulate_awgn(fs.encode_133171(info_s), scfg.sigma_vec[scfg.test_snr_index])
    wins, labs = [], []
    for off in range(0, 2 * 96, 2):
        bx, by = fs.get_single_instance(scfg.window_length, stream, info_s.reshape(1, -1), off)
        wins.append(bx)
        labs.append(by)
    xs = np.asarray(wins, dtype=np.float32).reshape(len(wins), scfg.window_length)
    tf.reset(seed=14, bias_scale=BIAS)
    os_, Ws_, bs_ = fs.build_network(xs, scfg)
    g["seq_seed"] = np.int64(14)
    g["seq_stream"], g["seq_info"] = stream[0], info_s.astype(np.uint8)
    g["seq_x"], g["seq_labels_onehot"], g["seq_out"], g["seq_pred"] = xs, np.asarray(labs).reshape(len(labs), -1), os_, tf.argmax(os_, 1)

    # ---- rnn_viterbi: prepare_batch + build_seq2seq_model (:235-377); cell arithmetic = the stand-in's LSTMCell ----
    rcfg = rv.RNNDecoderConfig(block_length=100)
    np.random.seed(105)
    B = 12
    bits = np.random.randint(0, 2, size=(B, rcfg.block_length))
    enc = np.zeros((B, 2 * rcfg.block_length), dtype=np.int32)
    for i in range(B):
        enc[i, :] = rv.modulate_awgn(rv.encode_133171(bits[i]), rcfg.sigma_vec[rcfg.test_snr_index])   # int32 rows (:216-222)
    info_pm = (2.0 * bits - 1.0).astype(np.float32)
    batch = rv.prepare_batch(info_pm, enc.astype(np.float32), rcfg)
    # placeholders are rank 2 [T, B] and expanded inside the model (:253-275); prepare_batch yields [T, B, 1] (B-12)
    tf.reset(seed=15, bias_scale=BIAS, feeds={k: np.asarray(v)[:, :, 0] for k, v in batch.items()})
    model = rv.build_seq2seq_model(rcfg)
    g["rnn_seed"] = np.int64(15)
    g["rnn_bits"], g["rnn_enc"] = bits.astype(np.uint8), enc
    g["rnn_logits"] = tf.TAPS["decoder_logits"][:, :, 0].T.copy()            # [B, T]
    g["rnn_loss"] = np.float64(model["loss"])
    g["rnn_prediction"] = np.asarray(model["decoder_prediction"]).T.copy()   # argmax over a size-1 axis: all zero (B-12)
    assert list(tf.VARIABLES) == ["rnn/lstm_cell/kernel", "rnn/lstm_cell/bias", "plain_decoder/lstm_cell/kernel",
                                  "plain_decoder/lstm_cell/bias", "fully_connected/weights", "fully_connected/biases"]

    np.savez_compressed(OUT, **g)
    print("wrote", os.path.abspath(OUT), len(g), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
