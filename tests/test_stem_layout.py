"""CPU check of the gather-stem data layout (csrc/conv_tc.cu stem kernels, plan route "stem"):
the GEMM the kernels run — A row = receptive field gathered from the 4-channel image with the
filter row widened to kw+1 taps, K index (r*(kw+1) + s)*4 + c, weight stored [Kpad][Cout] through
Engine._to_storage — equals the oracle's conv2d (TF SAME padding) for the supported stems."""
import numpy as np
import pytest
import torch

from myconvnet_b200.engine import Engine
from myconvnet_b200.plan import Plan
from oracle import tf_ops


class _Var(object):
    def __init__(self, shape, storage_shape, rows):
        self.shape, self.storage_shape, self.storage_rows = shape, storage_shape, rows


@pytest.mark.parametrize("k,h,w,stride_h", [(7, 20, 22, 2), (3, 11, 14, 2), (7, 17, 16, 2), (3, 12, 12, 1)])
def test_stem_gemm_layout_equals_conv2d(k, h, w, stride_h):
    rng = np.random.default_rng(0)
    n, co = 2, 8
    x = rng.standard_normal((n, h, w, 3)).astype(np.float32)
    wt = rng.standard_normal((k, k, 3, co)).astype(np.float32)
    ho, pad_t, _ = tf_ops.same_pad(h, k, stride_h, 1, "SAME")
    wo, pad_l, _ = tf_ops.same_pad(w, k, 2, 1, "SAME")
    assert pad_l % 2 == 0 and w % 2 == 0                       # the route's eligibility conditions
    kpad = Plan.stem_kpad(k, k)
    rows = np.array([(r * (k + 1) + s) * 4 + c for r in range(k) for s in range(k) for c in range(3)])
    v = _Var(wt.shape, (kpad, co), rows)
    wst = Engine._to_storage(None, v, wt)
    assert wst.shape == (kpad, co) and np.count_nonzero(wst) == np.count_nonzero(wt)
    assert np.array_equal(Engine._from_storage(None, v, wst), wt)      # round trip
    # 4-channel image, then the gather exactly as stem_gather does it (16-byte = 2-pixel chunks)
    x4 = np.zeros((n, h, w, 4), np.float32)
    x4[..., :3] = x
    epr, cpr = (k + 1) * 4, (k + 1) // 2
    A = np.zeros((n, ho, wo, kpad), np.float32)
    for p in range(ho):
        for q in range(wo):
            w0, h0 = q * 2 - pad_l, p * stride_h - pad_t
            for r in range(k):
                hh = h0 + r
                if not 0 <= hh < h:
                    continue
                for j in range(cpr):
                    ww = w0 + 2 * j
                    if ww >= 0 and ww + 1 < w:
                        A[:, p, q, r * epr + j * 8:r * epr + j * 8 + 8] = x4[:, hh, ww:ww + 2, :].reshape(n, 8)
    y = A.reshape(-1, kpad) @ wst
    ref = tf_ops.conv2d(torch.tensor(x), torch.tensor(wt), (stride_h, 2), "SAME", (1, 1)).numpy()
    np.testing.assert_allclose(y.reshape(ref.shape), ref, rtol=1e-4, atol=1e-4)
    # wgrad of the same GEMM restricted to the real rows is the conv's weight gradient; the
    # widening tap's rows would be non-zero (they see real pixels) and must be masked by the kernel
    dy = rng.standard_normal(ref.shape).astype(np.float32)
    dW = A.reshape(-1, kpad).T @ dy.reshape(-1, co)
    xt, wtt = torch.tensor(x), torch.tensor(wt, requires_grad=True)
    tf_ops.conv2d(xt, wtt, (stride_h, 2), "SAME", (1, 1)).backward(torch.tensor(dy))
    np.testing.assert_allclose(dW[rows].reshape(wt.shape), wtt.grad.numpy(), rtol=1e-3, atol=1e-3)
    widen = np.array([(r * (k + 1) + k) * 4 + c for r in range(k) for c in range(3)])
    assert np.abs(dW[widen]).max() > 0
