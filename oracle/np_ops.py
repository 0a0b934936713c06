"""ORACLE cross-check (test infrastructure): a second, independent formulation of the TF op
semantics in plain NumPy loops, written without looking at oracle/tf_ops.py's torch calls.
tests/test_oracle.py requires the two to agree; small shapes only (pure Python loops)."""
import numpy as np


def same_pad(n, k, s, d, padding):
    eff = (k - 1) * d + 1
    if padding == "SAME":
        out = (n + s - 1) // s
        tot = max((out - 1) * s + eff - n, 0)
        return out, tot // 2
    return (n - eff) // s + 1, 0


def conv2d(x, w, stride, padding, dilation=1):
    n, h, ww, ci = x.shape
    kh, kw, _, co = w.shape
    ho, pt = same_pad(h, kh, stride, dilation, padding)
    wo, pl = same_pad(ww, kw, stride, dilation, padding)
    y = np.zeros((n, ho, wo, co), dtype=np.float64)
    for p in range(ho):
        for q in range(wo):
            for a in range(kh):
                for b in range(kw):
                    hh = p * stride + a * dilation - pt
                    wq = q * stride + b * dilation - pl
                    if 0 <= hh < h and 0 <= wq < ww:
                        y[:, p, q, :] += x[:, hh, wq, :].astype(np.float64) @ w[a, b].astype(np.float64)
    return y


def depthwise_conv2d(x, w, stride, padding):
    n, h, ww, c = x.shape
    kh, kw, _, m = w.shape
    ho, pt = same_pad(h, kh, stride, 1, padding)
    wo, pl = same_pad(ww, kw, stride, 1, padding)
    y = np.zeros((n, ho, wo, c * m), dtype=np.float64)
    for p in range(ho):
        for q in range(wo):
            for a in range(kh):
                for b in range(kw):
                    hh, wq = p * stride + a - pt, q * stride + b - pl
                    if 0 <= hh < h and 0 <= wq < ww:
                        for mm in range(m):
                            y[:, p, q, mm::m] += x[:, hh, wq, :] * w[a, b, :, mm]
    return y


def conv2d_transpose(x, w_stored, out_hw, stride, padding):
    """Scatter form: y[ih*s + a - pad, iw*s + b - pad, co] += x[ih,iw,ci]*W[a,b,ci,co]."""
    n, h, ww, ci = x.shape
    kh, kw, _, co = w_stored.shape
    _, pt = same_pad(out_hw[0], kh, stride, 1, padding)
    _, pl = same_pad(out_hw[1], kw, stride, 1, padding)
    y = np.zeros((n, out_hw[0], out_hw[1], co), dtype=np.float64)
    for ih in range(h):
        for iw in range(ww):
            for a in range(kh):
                for b in range(kw):
                    oh, ow = ih * stride + a - pt, iw * stride + b - pl
                    if 0 <= oh < out_hw[0] and 0 <= ow < out_hw[1]:
                        y[:, oh, ow, :] += x[:, ih, iw, :].astype(np.float64) @ w_stored[a, b].astype(np.float64)
    return y


def batch_norm_train(x, gamma, beta, eps):
    c = x.shape[-1]
    f = x.reshape(-1, c).astype(np.float64)
    mean = f.sum(0) / f.shape[0]
    var = (f * f).sum(0) / f.shape[0] - mean * mean
    y = (x - mean) / np.sqrt(var + eps) * gamma + beta
    return y, mean, var * f.shape[0] / (f.shape[0] - 1)


def pool(x, k, stride, padding, kind):
    n, h, ww, c = x.shape
    ho, pt = same_pad(h, k, stride, 1, padding)
    wo, pl = same_pad(ww, k, stride, 1, padding)
    y = np.zeros((n, ho, wo, c))
    arg = np.zeros((n, ho, wo, c), dtype=np.int64)
    for p in range(ho):
        for q in range(wo):
            vals, idx = [], []
            for a in range(k):
                for b in range(k):
                    hh, wq = p * stride + a - pt, q * stride + b - pl
                    if 0 <= hh < h and 0 <= wq < ww:
                        vals.append(x[:, hh, wq, :])
                        idx.append((hh * ww + wq) * c + np.arange(c))
            v = np.stack(vals, 0)
            if kind == "max":
                j = v.argmax(0)              # numpy argmax returns the FIRST maximum
                y[:, p, q, :] = v.max(0)
                arg[:, p, q, :] = np.stack(idx, 0)[j, np.arange(c)[None, :].repeat(n, 0)]
            else:
                y[:, p, q, :] = v.mean(0)
    return (y, arg) if kind == "max" else y


def resize_bilinear(x, out_hw, align_corners, half_pixel):
    n, h, w, c = x.shape
    y = np.zeros((n, out_hw[0], out_hw[1], c))

    def src(d, i, o):
        if align_corners:
            return d * (i - 1) / (o - 1) if o > 1 else 0.0
        if half_pixel:
            return (d + 0.5) * i / o - 0.5
        return d * i / o
    for p in range(out_hw[0]):
        sh = src(p, h, out_hw[0])
        h0 = int(np.floor(sh))
        fh = sh - h0
        h1 = min(int(np.ceil(sh)), h - 1)
        h0 = max(h0, 0)
        for q in range(out_hw[1]):
            sw = src(q, w, out_hw[1])
            w0 = int(np.floor(sw))
            fw = sw - w0
            w1 = min(int(np.ceil(sw)), w - 1)
            w0 = max(w0, 0)
            top = x[:, h0, w0] * (1 - fw) + x[:, h0, w1] * fw
            bot = x[:, h1, w0] * (1 - fw) + x[:, h1, w1] * fw
            y[:, p, q] = top * (1 - fh) + bot * fh
    return y
