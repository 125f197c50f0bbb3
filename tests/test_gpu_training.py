"""The native training step (nvd_train_begin / nvd_train_step / nvd_train_end, csrc/k_train.cu) against the float64
oracle with hand-written gradients and TensorFlow's Adam arithmetic (oracle/trainer.py): costs and parameters after several
mini-batches for the FNN (softmax cross-entropy on tanh outputs + L2, /root/reference/src/fnn_viterbi.py:549-564), the
per-bit predictor (sigmoid cross-entropy on ReLU outputs, src/fnn_viterbi_bitwise.py:439-448) and a ragged shape."""
import numpy as np
import pytest

from oracle import mlp as omlp, trainer as otrain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim():
    from nvd import montecarlo
    with montecarlo.Simulator(device=0, seed=1) as s:
        yield s


@pytest.mark.parametrize("dims,acts,loss,l2,batch", [([14, 128, 64, 32, 128], [1, 1, 1, 1], "softmax_ce", (0, 1, 2), 256),
                                                     ([1036, 30, 2], [2, 2], "sigmoid_ce", (0,), 100),
                                                     ([33, 17, 5], [1, 0], "softmax_ce", (1,), 37)])
def test_training_steps_match_oracle(sim, dims, acts, loss, l2, batch):
    from nvd import training, weights
    rng = np.random.default_rng(len(dims) + batch)
    Ws, bs = omlp.init_fnn(dims, seed=7, bias_scale=0.05)
    steps = 6
    x = rng.standard_normal((3 * batch, dims[0])).astype(np.float32)
    y = np.eye(dims[-1], dtype=np.float32)[rng.integers(0, dims[-1], 3 * batch)]
    lr, beta = 1e-3, 1e-2
    wantW, wantb, want_costs = otrain.train(Ws, bs, acts, x, y, batch, steps, loss=loss, beta=beta, l2_layers=l2, lr=lr)
    m = weights.MlpModel([W.copy() for W in Ws], [b.copy() for b in bs], acts)
    got_costs = training.train_mlp_native(m, x, y, acts=acts, loss=loss, beta=beta, l2_layers=l2, lr=lr, epochs=2, batch_size=batch,
                                          display_step=0, sim=sim)
    # epoch averages of the per-step costs (3 batches per epoch)
    assert np.allclose(got_costs, [np.mean(want_costs[:3]), np.mean(want_costs[3:])], rtol=2e-5, atol=1e-6)
    for l in range(len(Ws)):
        # six Adam steps of size lr each: compare the UPDATES, relative to the step size
        assert np.abs(m.Ws[l] - wantW[l]).max() <= 2e-3 * lr * steps + 1e-7, (l, np.abs(m.Ws[l] - wantW[l]).max())
        assert np.abs(m.bs[l] - wantb[l]).max() <= 2e-3 * lr * steps + 1e-7
        assert np.abs(m.Ws[l] - Ws[l]).max() > 0.5 * lr                      # the parameters did move


def test_training_callback_early_stop_and_image_rebuild(sim):
    """The per-epoch callback sees the current parameters and can stop the run; after nvd_train_end the model's
    tensor-core image follows the trained weights."""
    import ctypes as C
    from nvd import _lib, training, weights
    rng = np.random.default_rng(3)
    m = weights.init_mlp([14, 128, 64, 32, 128], ["tanh"] * 4, seed=2)
    x = rng.standard_normal((512, 14)).astype(np.float32)
    y = np.eye(128, dtype=np.float32)[rng.integers(0, 128, 512)]
    seen = []
    costs = training.train_mlp_native(m, x, y, acts=[1, 1, 1, 1], lr=1e-3, epochs=5, batch_size=128, display_step=0, sim=sim,
                                      epoch_callback=lambda e, c, mm: seen.append(mm.Ws[0].copy()) or e == 2)
    assert len(costs) == 3 and len(seen) == 3 and not np.array_equal(seen[0], seen[1])
    assert costs[2] < costs[0]
    assert np.array_equal(seen[-1], m.Ws[0])
    # in-place training of a loaded model: bf16 decisions afterwards come from the trained weights
    h = sim.load_mlp(m)
    cfg = _lib.TrainConfig(C.sizeof(_lib.TrainConfig), 0, 5e-2, 0.0, 0, 0.9, 0.999, 1e-8)
    tr = C.c_int(-1)
    _lib.check(sim.lib.nvd_train_begin(sim.ctx, h, C.byref(cfg), 128, C.byref(tr)))
    d_x, d_y = sim.to_device(x[:128]), sim.to_device(y[:128])
    before, _ = sim.decode(h, x[:128], "bf16", want_soft=False)
    for _ in range(30):
        _lib.check(sim.lib.nvd_train_step(sim.ctx, tr.value, d_x.ptr, d_y.ptr, 128, None))
    _lib.check(sim.lib.nvd_train_end(sim.ctx, tr.value, None, None))
    after_tc, _ = sim.decode(h, x[:128], "bf16", want_soft=False)
    after_fp, _ = sim.decode(h, x[:128], "fp32", want_soft=False)
    assert (after_tc == after_fp).mean() > 0.95                     # the tcgen05 image was rebuilt from the new weights
    assert (after_fp == y[:128].argmax(1)).mean() > (before == y[:128].argmax(1)).mean() + 0.2     # it learnt the batch
