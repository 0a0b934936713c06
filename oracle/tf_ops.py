"""ORACLE (test infrastructure, never shipped or imported by the product path).

CPU restatement, in torch fp32/fp64, of the TensorFlow 1.x op semantics behind the reference's
layer ops (reference convnet.py:1382-2577).  TensorFlow is a third-party dependency of the
reference that is NOT under /root/reference and cannot be installed here (no network, no TF 1.x
build for Python 3.12), and the reference ships no tests or golden vectors: PARITY IS UNPINNED
against TensorFlow itself.  What pins this file instead (tests/test_oracle_*.py):
  * a second, independent NumPy loop formulation of every op (oracle/np_ops.py),
  * analytic known-answer tests (SAME-pad offset probes, BN of constants, argmax ties),
  * fp64 finite-difference gradient checks,
  * structure KATs from the reference's own counters (25,557,032 parameters for ResNet-50).
Each function cites the reference call site it restates and the SURVEY.md Appendix A rule.
All tensors are NHWC, weights HWIO, like the reference.
"""
import math

import torch
import torch.nn.functional as F


def same_pad(in_size, k, stride, dilation, padding):
    """TF SAME/VALID rule (Appendix A.1): SAME out = ceil(in/stride), extra pad bottom/right."""
    eff = (k - 1) * dilation + 1
    if padding.upper() == "SAME":
        out = -(-in_size // stride)
        total = max((out - 1) * stride + eff - in_size, 0)
        return out, total // 2, total - total // 2
    out = -(-(in_size - eff + 1) // stride)
    return out, 0, 0


def _pad_nchw(x, k, stride, dilation, padding, value=0.0):
    h, w = x.shape[2], x.shape[3]
    _, pt, pb = same_pad(h, k[0], stride[0], dilation[0], padding)
    _, pl, pr = same_pad(w, k[1], stride[1], dilation[1], padding)
    if pt or pb or pl or pr:
        x = F.pad(x, (pl, pr, pt, pb), value=value)
    return x


def conv2d(x, w, stride=(1, 1), padding="SAME", dilation=(1, 1)):
    """tf.nn.conv2d (convnet.py:1659): NHWC x HWIO cross-correlation, asymmetric SAME padding."""
    xc = _pad_nchw(x.permute(0, 3, 1, 2), w.shape[:2], stride, dilation, padding)
    y = F.conv2d(xc, w.permute(3, 2, 0, 1), stride=tuple(stride), dilation=tuple(dilation))
    return y.permute(0, 2, 3, 1)


def depthwise_conv2d(x, w, stride=(1, 1), padding="SAME", dilation=(1, 1)):
    """tf.nn.depthwise_conv2d (convnet.py:1645): filter [kh,kw,C,M], output channel c*M+m."""
    kh, kw, c, m = w.shape
    xc = _pad_nchw(x.permute(0, 3, 1, 2), (kh, kw), stride, dilation, padding)
    wt = w.permute(2, 3, 0, 1).reshape(c * m, 1, kh, kw)
    y = F.conv2d(xc, wt, stride=tuple(stride), dilation=tuple(dilation), groups=c)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose(x, w_stored, out_hw, stride=(1, 1), padding="SAME", dilation=(1, 1)):
    """tf.nn.conv2d_transpose as used at convnet.py:2460-2463 (Appendix A.3).  w_stored is the
    variable [kh,kw,Cin,Cout]; the op is the input-gradient of conv2d(., W[kh,kw,Cout,Cin])
    whose input has spatial size out_hw.  Restated through autograd of the forward conv."""
    n = x.shape[0]
    kh, kw, cin, cout = w_stored.shape
    w_conv = w_stored.permute(0, 1, 3, 2)          # HWIO of the underlying conv: I=Cout, O=Cin
    with torch.enable_grad():
        probe = torch.zeros(n, out_hw[0], out_hw[1], cout, dtype=x.dtype, requires_grad=True)
        y = conv2d(probe, w_conv, stride, padding, dilation)
        if tuple(y.shape[1:3]) != tuple(x.shape[1:3]):
            raise ValueError("conv2d_transpose: output_shape inconsistent with input")
        (g,) = torch.autograd.grad(y, probe, grad_outputs=x, create_graph=True)
    return g


def dense(x, w, b=None):
    """tf.matmul(x, w) + b (convnet.py:1743,1755)."""
    y = x @ w
    return y if b is None else y + b


def fused_batch_norm_train(x, gamma, beta, eps):
    """tf.nn.fused_batch_norm(is_training=True) (convnet.py:1883; Appendix A.4): normalise with
    the biased variance; also returns the batch mean and the Bessel-corrected variance."""
    c = x.shape[-1]
    xf = x.reshape(-1, c)
    n = xf.shape[0]
    mean = xf.mean(0)
    var = ((xf - mean) ** 2).mean(0)
    y = (x - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y, mean, var * (n / max(n - 1, 1))


def fused_batch_norm_infer(x, gamma, beta, mean, var, eps):
    y = (x - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y


def max_pool(x, k, stride, padding="SAME"):
    """tf.nn.max_pool (convnet.py:1509; Appendix A.5): padding acts as -inf."""
    xc = _pad_nchw(x.permute(0, 3, 1, 2), k, stride, (1, 1), padding, value=float("-inf"))
    return F.max_pool2d(xc, tuple(k), tuple(stride)).permute(0, 2, 3, 1)


def max_pool_argmax(x, k, stride, padding="SAME"):
    """Flattened argmax (h*W + w)*C + c within the image, first maximum in row-major window
    order (TF CPU tie rule, Appendix A.5).  Plain loops: small inputs only."""
    n, h, w, c = x.shape
    ho, pt, _ = same_pad(h, k[0], stride[0], 1, padding)
    wo, pl, _ = same_pad(w, k[1], stride[1], 1, padding)
    out = torch.full((n, ho, wo, c), -1, dtype=torch.int64)
    best = torch.full((n, ho, wo, c), float("-inf"), dtype=x.dtype)
    cidx = torch.arange(c)
    for a in range(k[0]):
        for b in range(k[1]):
            for p in range(ho):
                hh = p * stride[0] + a - pt
                if hh < 0 or hh >= h:
                    continue
                for q in range(wo):
                    ww = q * stride[1] + b - pl
                    if ww < 0 or ww >= w:
                        continue
                    v = x[:, hh, ww, :]
                    upd = (v > best[:, p, q, :]) | (out[:, p, q, :] < 0)
                    best[:, p, q, :] = torch.where(upd, v, best[:, p, q, :])
                    out[:, p, q, :] = torch.where(upd, (hh * w + ww) * c + cidx, out[:, p, q, :])
    return out


def avg_pool(x, k, stride, padding="SAME"):
    """tf.nn.avg_pool (convnet.py:1548; Appendix A.6): SAME divides by the in-bounds count."""
    xc = x.permute(0, 3, 1, 2)
    ones = torch.ones_like(xc[:, :1])
    xs = F.avg_pool2d(_pad_nchw(xc, k, stride, (1, 1), padding), tuple(k), tuple(stride),
                      divisor_override=1)
    cnt = F.avg_pool2d(_pad_nchw(ones, k, stride, (1, 1), padding), tuple(k), tuple(stride),
                       divisor_override=1)
    return (xs / cnt).permute(0, 2, 3, 1)


def global_avg_pool(x, keepdims=False):
    """tf.reduce_mean(x, axis=[1,2]) (resnet_v1_5.py:73, efficientnet.py:108,183)."""
    return x.mean(dim=(1, 2), keepdim=keepdims)


def _resize_src(out_size, in_size, mode, dtype):
    d = torch.arange(out_size, dtype=dtype)
    if mode == "align_corners":
        s = d * ((in_size - 1) / (out_size - 1)) if out_size > 1 else d * 0
    elif mode == "half_pixel":
        s = (d + 0.5) * (in_size / out_size) - 0.5
    else:
        s = d * (in_size / out_size)
    fl = torch.floor(s)
    lo = fl.clamp(min=0).long()
    hi = torch.ceil(s).clamp(max=in_size - 1).long().clamp(min=0)
    lo = lo.clamp(max=in_size - 1)
    return lo, hi, (s - fl)


def resize_bilinear(x, out_hw, align_corners=False, half_pixel_centers=False):
    """tf.image.resize_bilinear (convnet.py:2397; Appendix A.7)."""
    mode = "align_corners" if align_corners else ("half_pixel" if half_pixel_centers else "legacy")
    h0, h1, fh = _resize_src(out_hw[0], x.shape[1], mode, x.dtype)
    w0, w1, fw = _resize_src(out_hw[1], x.shape[2], mode, x.dtype)
    fh = fh.view(1, -1, 1, 1)
    fw = fw.view(1, 1, -1, 1)
    top = x[:, h0][:, :, w0] + (x[:, h0][:, :, w1] - x[:, h0][:, :, w0]) * fw
    bot = x[:, h1][:, :, w0] + (x[:, h1][:, :, w1] - x[:, h1][:, :, w0]) * fw
    return top + (bot - top) * fh


def resize_nearest(x, out_hw, align_corners=False, half_pixel_centers=False):
    """tf.image.resize_nearest_neighbor (convnet.py:2393-2395; Appendix A.7): src = round(dst*
    (in-1)/(out-1)) with align_corners, floor((dst+0.5)*in/out) with half-pixel centres, else
    floor(dst*in/out); clamped to in-1.  fp32 coordinate arithmetic as in the TF kernel."""
    def src(out_size, in_size):
        d = torch.arange(out_size, dtype=torch.float32)
        if align_corners:
            sc = torch.tensor((in_size - 1) / (out_size - 1) if out_size > 1 else 0.0, dtype=torch.float32)
            s = torch.floor(d * sc + 0.5)          # roundf: half away from zero (coordinates are >= 0)
        elif half_pixel_centers:
            s = torch.floor((d + 0.5) * torch.tensor(in_size / out_size, dtype=torch.float32))
        else:
            s = torch.floor(d * torch.tensor(in_size / out_size, dtype=torch.float32))
        return s.long().clamp(0, in_size - 1)
    return x[:, src(out_hw[0], x.shape[1])][:, :, src(out_hw[1], x.shape[2])]


def group_norm(x, gamma, beta, num_groups, eps):
    """ConvNet.group_norm (convnet.py:1928-2013): per sample and group of C/G channels,
    tf.nn.moments over (H, W, C/G) (biased variance), (x-mean)/sqrt(var+eps)*gamma+beta."""
    n, c = x.shape[0], x.shape[-1]
    xg = x.reshape(n, -1, num_groups, c // num_groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = ((xg - mean) ** 2).mean(dim=(1, 3), keepdim=True)
    y = ((xg - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y


def weight_standardization(w):
    """weight_variable(weight_standardization=True) (convnet.py:1410-1419): per output channel
    (last axis) subtract the mean over all other axes, divide by (population std + 1e-5)."""
    axes = tuple(range(w.dim() - 1))
    c = w - w.mean(dim=axes, keepdim=True)
    std = torch.sqrt((c ** 2).mean(dim=axes, keepdim=True))
    return c / (std + 1e-5)


def activation(x, kind, alpha=None):
    """convnet.py:2514-2556 (Appendix A.13)."""
    kind = (kind or "none").lower()
    if kind == "relu":
        return torch.relu(x)
    if kind == "relu6":
        return torch.clamp(x, 0.0, 6.0)
    if kind in ("lrelu", "leaky_relu"):
        a = 0.2 if alpha is None else alpha
        return torch.maximum(x, a * x)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "sigmoid":
        return torch.sigmoid(x)
    if kind == "swish":
        return x * torch.sigmoid(x)
    if kind == "none":
        return x
    raise ValueError("Activation type of {} is not supported".format(kind))


def softmax_cross_entropy(logits, onehot):
    """tf.nn.softmax_cross_entropy_with_logits_v2 (convnet.py:600; Appendix A.9)."""
    return -(onehot * torch.log_softmax(logits, dim=-1)).sum(-1)


def sigmoid_cross_entropy(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits (gan.py:134-136; Appendix A.9)."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def classification_loss(logits, labels_int, num_classes, class_w=None, label_smoothing=0.0,
                        focal_gamma=0.0, sigmoid_focal_alpha=0.0, spatial_smoothing=False):
    """Data term of reference convnet.py:552-594: one-hot (label -1 -> zero row), valid mask
    |sum(Y)-1| < 1e-5, mean over ALL rows of w*valid*CE.  Label smoothing: uniform
    (convnet.py:603-607) or, spatial_smoothing, the 5x5 SAME average of the one-hot map
    (segmentation/segnet.py:116-121; labels_int is then [N,H,W]).  Focal factors of
    convnet.py:580-592: (1-p_true)^gamma (differentiated) and the stop-gradient sigmoid form."""
    onehot = torch.zeros(labels_int.shape + (num_classes,), dtype=logits.dtype)
    ok = (labels_int >= 0) & (labels_int < num_classes)
    idx = labels_int.clamp(min=0, max=num_classes - 1)
    onehot.scatter_(-1, idx.unsqueeze(-1), 1.0)
    onehot = onehot * ok.unsqueeze(-1).to(logits.dtype)
    w = torch.ones(num_classes, dtype=logits.dtype) if class_w is None else torch.as_tensor(class_w, dtype=logits.dtype)
    batch_w = (onehot * w).sum(-1)
    valid = ((onehot.sum(-1) - 1.0).abs() < 1e-5).to(logits.dtype)
    if label_smoothing > 0 and spatial_smoothing:
        labels = (1.0 - label_smoothing) * onehot + label_smoothing * avg_pool(onehot, (5, 5), (1, 1), "SAME")
    elif label_smoothing > 0:
        labels = onehot * (1.0 - label_smoothing) + label_smoothing / num_classes
    else:
        labels = onehot
    ce = softmax_cross_entropy(logits, labels)
    if focal_gamma > 0 or sigmoid_focal_alpha > 0:
        p_true = (onehot * torch.softmax(logits, dim=-1)).sum(-1)
        if focal_gamma > 0:
            ce = ce * torch.pow(1.0 - p_true, focal_gamma)
        if sigmoid_focal_alpha > 0:
            f = (1.0 - torch.sigmoid(sigmoid_focal_alpha * (p_true - 0.5))).detach()
            ce = ce * (f / (1.0 - torch.sigmoid(torch.tensor(-0.5 * sigmoid_focal_alpha, dtype=logits.dtype))))
    return (batch_w * valid * ce).mean()


def l2_loss(w):
    """tf.nn.l2_loss = sum(w^2)/2 (convnet.py:563)."""
    return 0.5 * (w ** 2).sum()


# ---------------------------------------------------------------- optimisers (Appendix A.11, A.12)
def ema_decay(decay, num_updates):
    return min(decay, (1.0 + num_updates) / (10.0 + num_updates))


def nesterov_update(w, g, accum, lr, momentum):
    """tf.train.MomentumOptimizer(use_nesterov=True) (optimizers.py:676)."""
    accum = momentum * accum + g
    return w - lr * (g + momentum * accum), accum


def rmsprop_update(w, g, ms, mom, lr, decay=0.9, momentum=0.9, eps=1e-3):
    """tf.train.RMSPropOptimizer (optimizers.py:690); ms starts at one."""
    ms = decay * ms + (1 - decay) * g * g
    mom = momentum * mom + lr * g / torch.sqrt(ms + eps)
    return w - mom, ms, mom


def adam_update(w, g, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-3):
    """tf.train.AdamOptimizer (optimizers.py:704); t counts from 1."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return w - lr_t * m / (torch.sqrt(v) + eps), m, v
