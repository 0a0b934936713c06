"""Generates tests/golden/ops_v1.npz: inputs and float64 outputs of the layer ops on small seeded
tensors, computed with the INDEPENDENT NumPy loop formulation (oracle/np_ops.py) and closed forms
written out here — not with the torch oracle that the parity tests use, and not with the device
code.  The reference itself cannot produce vectors offline (TensorFlow 1.x is not installable), so
these fixtures pin the oracle against regressions and give the GPU tests a frozen target; they are
NOT TensorFlow outputs ("parity unpinned").

    python tests/golden/make_golden.py        # rewrites ops_v1.npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import np_ops  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    g = {}
    # convolutions: (tag, n, h, w, ci, co, k, stride, padding, dilation)
    for tag, n, h, w, ci, co, k, s, pad, d in [
            ("conv_same_s1", 2, 7, 9, 3, 5, 3, 1, "SAME", 1), ("conv_same_s2_even", 1, 8, 8, 4, 4, 3, 2, "SAME", 1),
            ("conv_7x7_s2", 1, 12, 12, 3, 4, 7, 2, "SAME", 1), ("conv_dilated", 1, 10, 10, 2, 2, 3, 1, "SAME", 2),
            ("conv_valid", 2, 8, 9, 3, 4, 3, 1, "VALID", 1), ("conv_1x1_s2", 1, 6, 6, 3, 4, 1, 2, "SAME", 1)]:
        x = rng.standard_normal((n, h, w, ci))
        wt = rng.standard_normal((k, k, ci, co))
        g[tag + "/x"], g[tag + "/w"] = x, wt
        g[tag + "/attrs"] = np.array([k, s, d, 1 if pad == "SAME" else 0])
        g[tag + "/y"] = np_ops.conv2d(x, wt, s, pad, d)
    x = rng.standard_normal((2, 9, 9, 6))
    wd = rng.standard_normal((3, 3, 6, 1))
    g["dwconv/x"], g["dwconv/w"], g["dwconv/y"] = x, wd, np_ops.depthwise_conv2d(x, wd, 2, "SAME")
    x = rng.standard_normal((2, 4, 4, 6))
    wt = rng.standard_normal((5, 5, 6, 3))            # stored [kh,kw,Cin,Cout] (convnet.py:2460)
    g["tconv/x"], g["tconv/w"], g["tconv/y"] = x, wt, np_ops.conv2d_transpose(x, wt, (8, 8), 2, "SAME")
    # batch norm (training): y, batch mean, UNBIASED variance
    x = rng.standard_normal((3, 5, 4, 6)) * 2 + 0.5
    gamma, beta = rng.uniform(0.5, 1.5, 6), rng.standard_normal(6)
    y, m, v = np_ops.batch_norm_train(x, gamma, beta, 1e-3)
    g["bn/x"], g["bn/gamma"], g["bn/beta"], g["bn/y"], g["bn/mean"], g["bn/var_unbiased"] = x, gamma, beta, y, m, v
    # pooling
    x = rng.standard_normal((2, 9, 11, 4))
    g["pool/x"] = x
    g["pool/max_3_2_same"], _ = np_ops.pool(x, 3, 2, "SAME", "max")
    g["pool/max_2_2_valid"], _ = np_ops.pool(x, 2, 2, "VALID", "max")
    g["pool/avg_3_2_same"] = np_ops.pool(x, 3, 2, "SAME", "avg")
    # argmax on quantised values (ties): FIRST maximum in row-major window order, flattened
    # (h*W + w)*C + c inside the image — written out here as explicit loops
    xq = rng.integers(0, 3, size=(1, 5, 6, 2)).astype(np.float64)
    k, s = 3, 2
    ho, pt = np_ops.same_pad(5, k, s, 1, "SAME")
    wo, pl = np_ops.same_pad(6, k, s, 1, "SAME")
    am = np.zeros((1, ho, wo, 2), np.int64)
    for p in range(ho):
        for q in range(wo):
            for c in range(2):
                best, idx = -np.inf, -1
                for r in range(k):
                    for t in range(k):
                        hh, ww = p * s - pt + r, q * s - pl + t
                        if 0 <= hh < 5 and 0 <= ww < 6 and xq[0, hh, ww, c] > best:
                            best, idx = xq[0, hh, ww, c], (hh * 6 + ww) * 2 + c
                am[0, p, q, c] = idx
    g["argmax/x"], g["argmax/idx_3_2_same"] = xq, am
    # bilinear resize, the three TF coordinate conventions
    x = rng.standard_normal((1, 4, 5, 3))
    g["resize/x"] = x
    g["resize/align_corners"] = np_ops.resize_bilinear(x, (7, 9), True, False)
    g["resize/legacy"] = np_ops.resize_bilinear(x, (8, 10), False, False)
    g["resize/half_pixel"] = np_ops.resize_bilinear(x, (8, 10), False, True)
    # softmax cross-entropy with label smoothing, mean over rows; a -1 label is an all-zero row
    logits = rng.standard_normal((6, 5))
    labels = np.array([0, 4, 2, -1, 1, 3])
    ls = 0.1
    onehot = np.zeros((6, 5))
    for i, l in enumerate(labels):
        if l >= 0:
            onehot[i, l] = 1.0
    valid = onehot.sum(1)
    sm = onehot * (1 - ls) + ls / 5 * valid[:, None]            # tf.losses-style smoothing of valid rows
    z = logits - logits.max(1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(1, keepdims=True))
    g["xent/logits"], g["xent/labels"] = logits, labels
    g["xent/loss_mean_ls0"] = np.mean(-(onehot * logp).sum(1))
    g["xent/loss_mean_ls01"] = np.mean(-(sm * logp).sum(1))
    # optimiser rules and schedules (closed forms of Appendix A.10-A.12)
    w, gr, acc = rng.standard_normal(7), rng.standard_normal(7), rng.standard_normal(7)
    a2 = 0.9 * acc + gr
    g["nesterov/w"], g["nesterov/g"], g["nesterov/accum"] = w, gr, acc
    g["nesterov/w_new"], g["nesterov/accum_new"] = w - 0.05 * (gr + 0.9 * a2), a2
    ms = np.abs(rng.standard_normal(7)) + 0.5
    ms2 = 0.9 * ms + 0.1 * gr * gr
    mom2 = 0.9 * acc + 0.05 * gr / np.sqrt(ms2 + 1e-3)
    g["rmsprop/ms"], g["rmsprop/w_new"], g["rmsprop/ms_new"], g["rmsprop/mom_new"] = ms, w - mom2, ms2, mom2
    g["ema/decay_t"] = np.array([min(0.999, (1.0 + t) / (10.0 + t)) for t in (0, 1, 10, 100, 100000)])
    np.savez_compressed(os.path.join(HERE, "ops_v1.npz"), **{k.replace("/", "|"): v for k, v in g.items()})
    print("wrote %d arrays" % len(g))


if __name__ == "__main__":
    main()
