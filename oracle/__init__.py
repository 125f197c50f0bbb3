"""CPU oracle for the Monte-Carlo decode path (TEST INFRASTRUCTURE — not the product).

This package is a numpy / torch-CPU restatement of the reference algorithm
(enessinanparildi/Neural-Network-Based-Viterbi-Decoding, pure Python) for the
path  info bits -> convolutional encoder -> BPSK+AWGN -> neural decoder forward
-> hard decision -> BER/FER.  Every function cites the reference file:line it
follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it, and only as the checker
or the timed CPU baseline; the product path (the CUDA library behind
`include/nvd.h`) never calls into it and fails loudly if the CUDA extension is
missing.

Parity pinning:
  * encoders, state bits, class-bit table, label packing, Hamming/BER counting,
    BPSK mapping, window extraction, LSTM batch layout: PINNED against the live
    reference code (imported under a stub `tensorflow`) by `tools/gen_golden.py`;
    vectors committed under `tests/golden/`.
  * decoder forward (MLP / LSTM arithmetic): the reference executes it inside
    TensorFlow 1.x, which is not in `/root/reference` nor installable here ->
    "parity unpinned" for that stage: the oracle restates the graph definition
    (`src/fnn_viterbi.py:495-508` etc.) and TF-1.x `LSTMCell` published semantics.
"""

from . import encoders, philox, channel, mlp, lstm, metrics, pipeline  # noqa: F401
