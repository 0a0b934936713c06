"""The torch oracle (oracle/tf_ops.py, what every parity test compares the device against) checked
against the committed golden fixtures tests/golden/ops_v1.npz, which were produced by the
independent NumPy loop formulation and closed forms (tests/golden/make_golden.py).  The fixtures
are NOT TensorFlow outputs — TensorFlow cannot run here — so parity with TF stays "unpinned"."""
import os

import numpy as np
import pytest
import torch

from oracle import tf_ops

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    with np.load(os.path.join(HERE, "golden", "ops_v1.npz")) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


def T(a):
    return torch.tensor(np.asarray(a, dtype=np.float64))


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * (1 + np.max(np.abs(b)))


@pytest.mark.parametrize("tag", ["conv_same_s1", "conv_same_s2_even", "conv_7x7_s2", "conv_dilated",
                                 "conv_valid", "conv_1x1_s2"])
def test_conv2d(G, tag):
    k, s, d, same = [int(v) for v in G[tag + "/attrs"]]
    y = tf_ops.conv2d(T(G[tag + "/x"]), T(G[tag + "/w"]), (s, s), "SAME" if same else "VALID", (d, d))
    close(y.numpy(), G[tag + "/y"])


def test_depthwise_and_transposed(G):
    close(tf_ops.depthwise_conv2d(T(G["dwconv/x"]), T(G["dwconv/w"]), (2, 2), "SAME").numpy(), G["dwconv/y"])
    close(tf_ops.conv2d_transpose(T(G["tconv/x"]), T(G["tconv/w"]), (8, 8), (2, 2), "SAME").detach().numpy(), G["tconv/y"])


def test_batch_norm_train(G):
    y, m, v = tf_ops.fused_batch_norm_train(T(G["bn/x"]), T(G["bn/gamma"]), T(G["bn/beta"]), 1e-3)
    close(y.numpy(), G["bn/y"], 1e-9)
    close(m.numpy(), G["bn/mean"])
    close(v.numpy(), G["bn/var_unbiased"], 1e-9)


def test_pooling_and_argmax(G):
    x = T(G["pool/x"])
    close(tf_ops.max_pool(x, [3, 3], [2, 2], "SAME").numpy(), G["pool/max_3_2_same"])
    close(tf_ops.max_pool(x, [2, 2], [2, 2], "VALID").numpy(), G["pool/max_2_2_valid"])
    close(tf_ops.avg_pool(x, [3, 3], [2, 2], "SAME").numpy(), G["pool/avg_3_2_same"])
    idx = tf_ops.max_pool_argmax(T(G["argmax/x"]), [3, 3], [2, 2], "SAME").numpy()
    assert np.array_equal(idx.astype(np.int64), G["argmax/idx_3_2_same"])          # bit-exact, ties included


def test_resize_bilinear(G):
    x = T(G["resize/x"])
    close(tf_ops.resize_bilinear(x, (7, 9), align_corners=True).numpy(), G["resize/align_corners"])
    close(tf_ops.resize_bilinear(x, (8, 10)).numpy(), G["resize/legacy"])
    close(tf_ops.resize_bilinear(x, (8, 10), half_pixel_centers=True).numpy(), G["resize/half_pixel"])


def test_loss_and_update_rules(G):
    logits, labels = T(G["xent/logits"]), torch.tensor(G["xent/labels"])
    close(tf_ops.classification_loss(logits, labels, 5).numpy(), G["xent/loss_mean_ls0"])
    close(tf_ops.classification_loss(logits, labels, 5, label_smoothing=0.1).numpy(), G["xent/loss_mean_ls01"])
    w, a = tf_ops.nesterov_update(T(G["nesterov/w"]), T(G["nesterov/g"]), T(G["nesterov/accum"]), 0.05, 0.9)
    close(w.numpy(), G["nesterov/w_new"])
    close(a.numpy(), G["nesterov/accum_new"])
    w, ms, mom = tf_ops.rmsprop_update(T(G["nesterov/w"]), T(G["nesterov/g"]), T(G["rmsprop/ms"]),
                                       T(G["nesterov/accum"]), 0.05, 0.9, 0.9, 1e-3)
    close(w.numpy(), G["rmsprop/w_new"])
    close(ms.numpy(), G["rmsprop/ms_new"])
    close(mom.numpy(), G["rmsprop/mom_new"])
    got = [tf_ops.ema_decay(0.999, t) for t in (0, 1, 10, 100, 100000)]
    close(got, G["ema/decay_t"])
