#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz by RUNNING the live reference.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tools/gen_golden.py
The reference's numpy-only functions import unchanged once a stub `tensorflow` (only the
names its annotations touch) and a stub `matplotlib.pyplot` precede them on sys.path
(SURVEY.md §8c).  `rnn_viterbi.py` fails at import (`Optional` undefined at :35), so its
numpy functions are compiled from its AST instead.  Nothing is copied from the reference:
only inputs and the outputs it computed are stored.
"""
import ast
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_vectors.npz")


def _install_stubs():
    d = tempfile.mkdtemp(prefix="nvd_stub_")
    with open(os.path.join(d, "tensorflow.py"), "w") as f:
        f.write("class Tensor: pass\nclass Variable: pass\n")
    os.makedirs(os.path.join(d, "matplotlib"))
    open(os.path.join(d, "matplotlib", "__init__.py"), "w").close()
    open(os.path.join(d, "matplotlib", "pyplot.py"), "w").close()
    sys.path.insert(0, REF)
    sys.path.insert(0, d)


def _load_rnn_functions():
    """exec the numpy-only defs of rnn_viterbi.py from its AST (module import raises NameError)."""
    src = open(os.path.join(REF, "rnn_viterbi.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))
            and n.name in ("RNNDecoderConfig", "encode_133171", "modulate_awgn",
                           "generate_input_vectors", "generate_datasets", "prepare_batch")]
    mod = types.ModuleType("rnn_viterbi_ast")
    from typing import Dict, Optional, Tuple
    mod.__dict__.update(np=np, Dict=Dict, Optional=Optional, Tuple=Tuple)
    exec(compile(ast.Module(body=keep, type_ignores=[]), "rnn_viterbi.py", "exec"), mod.__dict__)
    return mod


def main():
    _install_stubs()
    import fnn_viterbi as fv
    import fnn_viterbi_mixed_snr as fm
    import fnn_viterbi_bitwise as fb
    import fnn_viterbi_seq_predict as fs
    rv = _load_rnn_functions()

    g = {}
    rng = np.random.default_rng(20170710)

    # ---- encoders (a4, a5, a6) -----------------------------------------------------
    cb7 = fv.generate_class_bits(7)
    g["class_bits_7"] = cb7
    g["class_bits_3"] = fv.generate_class_bits(3)
    g["enc133171_book7"] = np.stack([fv.encode_133171(r) for r in cb7])
    g["enc57_book7"] = np.stack([fv.encode_57(r) for r in cb7])
    for L in (6, 10, 33, 100, 518):
        x = rng.integers(0, 2, size=(8, L))
        g[f"enc_in_{L}"] = x.astype(np.uint8)
        g[f"enc133171_out_{L}"] = np.stack([fv.encode_133171(r) for r in x])
        g[f"enc57_out_{L}"] = np.stack([fv.encode_57(r) for r in x])
    xs = rng.integers(0, 2, size=20)
    g["state_in"] = xs.astype(np.uint8)
    g["state_out"] = np.stack([fb.encode_133171_with_state(xs, s) for s in range(64)])
    g["state_bits"] = np.stack([fb.extract_state_bits(s, 6) for s in range(64)])
    xf = rng.integers(0, 2, size=12).astype(np.float32)       # float inputs work (% 2)
    g["enc_float_in"] = xf
    g["enc_float_out"] = fv.encode_133171(xf)
    # identical copies in the other files
    x = rng.integers(0, 2, size=40)
    assert (fm.encode_133171(x) == fv.encode_133171(x)).all()
    assert (fs.encode_133171(x) == fv.encode_133171(x)).all()
    assert (rv.encode_133171(x) == fv.encode_133171(x)).all()
    assert (fm.encode_57(x) == fv.encode_57(x)).all()

    # ---- config constants (a1) -----------------------------------------------------
    g["sigma_2db"] = np.float64(fv.DecoderConfig().sigma)
    g["sigma_vec_mixed"] = fm.MixedSNRDecoderConfig().sigma_vec
    g["snr_db_vec_mixed"] = fm.MixedSNRDecoderConfig().snr_db_vec
    g["samples_per_snr_mixed"] = np.int64(fm.MixedSNRDecoderConfig().samples_per_snr)
    g["sigma_vec_bitwise"] = fb.BitwiseDecoderConfig().sigma_vec
    g["sigma_vec_rnn"] = rv.RNNDecoderConfig().sigma_vec
    c = fs.SequentialDecoderConfig()
    g["seq_total_test_instances"] = np.int64(c.total_test_instances)

    # ---- channel (a7) with the legacy numpy RNG seeded -----------------------------
    cw = g["enc133171_book7"][77]
    np.random.seed(1234)
    g["awgn_in"] = cw
    g["awgn_out_seed1234"] = fv.modulate_awgn(cw, 0.5617)
    np.random.seed(1234)
    row_i32 = np.zeros(14, dtype=np.int32)
    row_i32[:] = fb.modulate_awgn(cw, 0.8902)                  # int32 destination -> truncation
    g["awgn_out_i32_seed1234"] = row_i32
    np.random.seed(1234)
    g["awgn_out_f64_sigma0p8902"] = fb.modulate_awgn(cw, 0.8902)

    # ---- labels / metrics (a12-a14) ------------------------------------------------
    g["one_hot_3"] = fv.create_one_hot_labels(np.array([0, 2, 1]), 3)
    vals = np.array([0, 1, 5, 80, 127])
    g["ebs_in"] = vals
    g["ebs_out"] = np.stack([fm.extract_bits_single(int(v), 7) for v in vals])
    g["b2d_out"] = np.array([fm.binary_to_decimal(fm.extract_bits_single(int(v), 7)) for v in vals])
    a = rng.integers(0, 128, size=64)
    b = rng.integers(0, 128, size=64)
    g["ham_a"], g["ham_b"] = a, b
    g["ham_out"] = np.array([fm.calculate_hamming_distance_decimal(int(x), int(y), 7) for x, y in zip(a, b)])
    g["ber_out"] = np.float64(fm.calculate_ber(a, b, 7))
    g["acc_out"] = np.float64(fm.calculate_accuracy(a, b))

    # calculate_error on a small class-major set (prints -> captured)
    cfg = fv.DecoderConfig(block_length=4, samples_per_class_test=5)
    labels = np.arange(16).repeat(5)
    pred = labels.copy()
    flip = rng.choice(len(pred), size=17, replace=False)
    pred[flip] = rng.integers(0, 16, size=17)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fv.calculate_error(fv.create_one_hot_labels(labels, 16), pred, fv.generate_class_bits(4), cfg)
    g["calc_error_labels"] = labels
    g["calc_error_pred"] = pred
    g["calc_error_stdout"] = np.array(buf.getvalue())

    # ---- bitwise / seq / rnn data layout -------------------------------------------
    win = np.arange(40, dtype=np.float64).reshape(1, 40)
    info = (np.arange(20) % 2).reshape(1, 20)
    bx, by = fs.get_single_instance(8, win, info, 6)
    g["gsi_x"], g["gsi_y"] = bx, by
    rcfg = rv.RNNDecoderConfig(block_length=5)
    info_b = np.array([[1, -1, -1, 1, 1], [-1, -1, 1, 1, -1]], dtype=np.float32)
    enc_b = rng.standard_normal((2, 10)).astype(np.float32)
    pb = rv.prepare_batch(info_b, enc_b, rcfg)
    g["pb_info"], g["pb_enc"] = info_b, enc_b
    g["pb_encoder_inputs"] = pb["encoder_inputs"]
    g["pb_decoder_inputs"] = pb["decoder_inputs"]
    g["pb_decoder_targets"] = pb["decoder_targets"]

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **g)
    print("wrote", os.path.abspath(OUT), len(g), "arrays")


if __name__ == "__main__":
    main()
