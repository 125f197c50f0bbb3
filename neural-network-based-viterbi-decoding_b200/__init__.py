"""B200-native Monte-Carlo decode path (encode -> BPSK+AWGN -> neural decode -> BER/FER).

Layout: csrc/ (CUDA kernels + the C ABI of include/nvd.h), nvd/ (ctypes host layer), src/ (drop-in
modules with the reference's names: fnn_viterbi, fnn_viterbi_mixed_snr, fnn_viterbi_bitwise,
fnn_viterbi_seq_predict, rnn_viterbi), production_wrapper.py + utils/.
The directory name is not a Python identifier; put it (and its src/) on sys.path, as
tests/conftest.py and __graft_entry__.py do.
"""
