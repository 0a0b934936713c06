"""GPU parity of every libmcn entry point against the CPU oracle (oracle/tf_ops.py + autograd),
called through the C ABI exactly as the engine does.  fp32 kernels: rtol 1e-4-ish; bf16 kernels
vs the oracle on bf16-rounded inputs: rel-L2 1e-2."""
import numpy as np
import pytest
import torch

from oracle import tf_ops
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from myconvnet_b200 import lib
    lib.load()
    lib.ensure_workspace(64 << 20)      # reduction workspace of the deterministic kernels (mcn.h)
    return lib


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def tdt(code):
    return torch.float32 if code == 0 else torch.bfloat16


def desc_for(L, x_shape, w_shape, stride, dil, padding):
    n, h, w, ci = x_shape
    kh, kw, _, co = w_shape
    ho, pt, _ = tf_ops.same_pad(h, kh, stride, dil, padding)
    wo, pl, _ = tf_ops.same_pad(w, kw, stride, dil, padding)
    return L.ConvDescC(n, h, w, ci, co, kh, kw, stride, stride, dil, dil, pt, pl, ho, wo), ho, wo


CONV_CASES = [
    # n, h, w, ci, co, k, stride, dil, padding
    (2, 9, 11, 8, 16, 3, 1, 1, "SAME"),
    (2, 12, 12, 16, 8, 3, 2, 1, "SAME"),
    (1, 13, 10, 8, 8, 5, 2, 1, "SAME"),
    (2, 10, 10, 8, 24, 3, 1, 2, "SAME"),
    (2, 9, 9, 16, 16, 1, 2, 1, "SAME"),
    (2, 11, 11, 3, 16, 7, 2, 1, "SAME"),
    (2, 10, 9, 8, 8, 3, 1, 1, "VALID"),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("code", [0, 1])
def test_conv_direct(L, case, code):
    n, h, w, ci, co, k, s, dl, pad = case
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, h, w, ci)).astype(np.float32)
    wt = (rng.standard_normal((k, k, ci, co)) * 0.2).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    d, ho, wo = desc_for(L, x.shape, wt.shape, s, dl, pad)
    dy = rng.standard_normal((n, ho, wo, co)).astype(np.float32)
    dt = tdt(code)
    if code == 1:
        x, dy = [torch.tensor(a).bfloat16().float().numpy() for a in (x, dy)]
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    yref = tf_ops.conv2d(xt, wtt, (s, s), pad, (dl, dl)) + torch.tensor(b)
    yref.backward(torch.tensor(dy))
    lib = L.load()
    xd, wd, bd, dyd = dev(x, dt), dev(wt), dev(b), dev(dy, dt)
    y = torch.empty(n, ho, wo, co, device="cuda", dtype=dt)
    L.check(lib.mcn_conv2d_fprop_direct(d, code, xd.data_ptr(), 0, wd.data_ptr(), bd.data_ptr(),
                                        y.data_ptr(), None))
    dx = torch.empty_like(xd)
    L.check(lib.mcn_conv2d_dgrad_direct(d, code, dyd.data_ptr(), 0, wd.data_ptr(), dx.data_ptr(), None))
    dw = torch.zeros(k, k, ci, co, device="cuda")
    L.check(lib.mcn_conv2d_wgrad_direct(d, code, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), None))
    torch.cuda.synchronize()
    tol = 1e-5 if code == 0 else 6e-3
    assert rel_l2(y.float().cpu(), yref.detach()) < tol
    assert rel_l2(dx.float().cpu(), xt.grad) < tol
    assert rel_l2(dw.cpu(), wtt.grad) < 1e-5


TC_CASES = [
    (2, 14, 14, 64, 64, 3, 1, 1, "SAME"),
    (2, 28, 28, 64, 128, 3, 2, 1, "SAME"),
    (2, 15, 15, 64, 64, 3, 2, 1, "SAME"),
    (2, 16, 16, 128, 64, 1, 1, 1, "SAME"),
    (2, 28, 28, 64, 128, 1, 2, 1, "SAME"),
    (2, 16, 16, 64, 64, 3, 1, 2, "SAME"),
    (3, 16, 16, 64, 64, 5, 2, 1, "SAME"),
    (2, 12, 12, 72, 40, 3, 1, 1, "SAME"),
    (4, 7, 7, 128, 256, 3, 1, 1, "SAME"),
    # EfficientNet-B0 pointwise shapes: channel counts that are multiples of 8 but not of 64 (the
    # TMA-store epilogue clips the partial last 64-channel chunk; dgrad writes Cin = 16 / 24 / 96 ...)
    (2, 28, 28, 16, 96, 1, 1, 1, "SAME"),
    (2, 14, 14, 96, 24, 1, 1, 1, "SAME"),
    (3, 14, 14, 144, 40, 1, 1, 1, "SAME"),
    (2, 7, 7, 320, 1280, 1, 1, 1, "SAME"),
]


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_conv_tensor_core(L, case, mode):
    n, h, w, ci, co, k, s, dl, pad = case
    if mode == 0 and s > 1 and k > 1:
        pytest.skip("box-mode fprop/wgrad need stride 1 (strided convs use im2col mode)")
    rng = np.random.default_rng(1)
    r = lambda a: torch.tensor(a).bfloat16().float().numpy()
    x = r(rng.standard_normal((n, h, w, ci)).astype(np.float32))
    wt = r((rng.standard_normal((k, k, ci, co)) * 0.1).astype(np.float32))
    d, ho, wo = desc_for(L, x.shape, wt.shape, s, dl, pad)
    dy = r(rng.standard_normal((n, ho, wo, co)).astype(np.float32))
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    yref = tf_ops.conv2d(xt, wtt, (s, s), pad, (dl, dl))
    yref.backward(torch.tensor(dy))
    lib = L.load()
    xd, dyd = dev(x, torch.bfloat16), dev(dy, torch.bfloat16)
    wd = dev(wt)
    w_hwio = torch.empty(k * k, ci, co, device="cuda", dtype=torch.bfloat16)
    w_ohwi = torch.empty(k * k, co, ci, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_weight_prep(wd.data_ptr(), k * k, ci, co, w_hwio.data_ptr(), w_ohwi.data_ptr(), None))
    y = torch.empty(n, ho, wo, co, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_conv2d_fprop_tc(d, xd.data_ptr(), w_ohwi.data_ptr(), None, y.data_ptr(), 1, mode, 0, None))
    dx = torch.full((n, h, w, ci), 7.0, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_conv2d_dgrad_tc(d, dyd.data_ptr(), w_hwio.data_ptr(), dx.data_ptr(), 1, mode, 0, None))
    # accumulate-in-epilogue variants: out += op(.)
    base = torch.tensor(r(rng.standard_normal((n, h, w, ci)).astype(np.float32)))
    dx2 = base.cuda().bfloat16()
    L.check(lib.mcn_conv2d_dgrad_tc(d, dyd.data_ptr(), w_hwio.data_ptr(), dx2.data_ptr(), 1, mode, 1, None))
    ybase = torch.tensor(r(rng.standard_normal((n, ho, wo, co)).astype(np.float32)))
    y2 = ybase.cuda().bfloat16()
    L.check(lib.mcn_conv2d_fprop_tc(d, xd.data_ptr(), w_ohwi.data_ptr(), None, y2.data_ptr(), 1, mode, 1, None))
    dw = torch.zeros(k, k, ci, co, device="cuda")
    L.check(lib.mcn_conv2d_wgrad_tc(d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), mode, None))
    torch.cuda.synchronize()
    assert rel_l2(y.float().cpu(), yref.detach()) < 6e-3
    assert rel_l2(dx.float().cpu(), xt.grad) < 6e-3
    assert rel_l2(dw.cpu(), wtt.grad) < 1e-4
    assert rel_l2(dx2.float().cpu(), xt.grad + base) < 6e-3
    assert rel_l2(y2.float().cpu(), yref.detach() + ybase) < 6e-3


STATS_CASES = [
    # n, h, w, ci, co, k, stride, mode  (mode 2 = halo tiles where eligible)
    (2, 14, 14, 64, 64, 3, 1, 1),
    (3, 16, 16, 128, 256, 1, 1, 0),      # block_n 256
    (2, 28, 28, 64, 128, 3, 2, 1),
    (1, 56, 56, 64, 64, 3, 1, 2),        # halo kernel, weight-stationary
    (1, 62, 54, 64, 192, 3, 1, 2),       # halo kernel with ragged tiles; Cout 192 = 2 n-tiles of 128
    (5, 7, 7, 128, 512, 1, 1, 0),        # box tiles merged over the batch, rows of padding in the last tile
    (2, 12, 12, 64, 2048, 1, 1, 0),      # 8 n-tiles: per-CTA accumulator flushes between tiles
]


@pytest.mark.parametrize("case", STATS_CASES)
@pytest.mark.parametrize("bias", [False, True])
def test_conv_fused_bn_statistics(L, case, bias):
    """mcn_conv2d_fprop_tc_stats: same y as the plain kernel, and sums = per-channel sum / sum of
    squares of the STORED bf16 output (what the separate mcn_bn_stats pass would have read)."""
    n, h, w, ci, co, k, s, mode = case
    rng = np.random.default_rng(3)
    r = lambda a: torch.tensor(a).bfloat16().float().numpy()
    x = r(rng.standard_normal((n, h, w, ci)).astype(np.float32) + 0.3)
    wt = r((rng.standard_normal((k, k, ci, co)) * 0.1).astype(np.float32))
    b = rng.standard_normal(co).astype(np.float32) if bias else None
    d, ho, wo = desc_for(L, x.shape, wt.shape, s, 1, "SAME")
    lib = L.load()
    xd, wd = dev(x, torch.bfloat16), dev(wt)
    bd = dev(b) if bias else None
    w_hwio = torch.empty(k * k, ci, co, device="cuda", dtype=torch.bfloat16)
    w_ohwi = torch.empty(k * k, co, ci, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_weight_prep(wd.data_ptr(), k * k, ci, co, w_hwio.data_ptr(), w_ohwi.data_ptr(), None))
    y0 = torch.empty(n, ho, wo, co, device="cuda", dtype=torch.bfloat16)
    y1 = torch.full_like(y0, 3.0)
    bp = bd.data_ptr() if bias else None
    L.check(lib.mcn_conv2d_fprop_tc(d, xd.data_ptr(), w_ohwi.data_ptr(), bp, y0.data_ptr(), 1, mode, 0, None))
    sums = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_conv2d_fprop_tc_stats(d, xd.data_ptr(), w_ohwi.data_ptr(), bp, y1.data_ptr(), mode,
                                          sums.data_ptr(), None))
    ref = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_bn_stats(1, y0.data_ptr(), n * ho * wo, co, ref.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)                                   # bit-identical output
    yf = y0.double().reshape(-1, co)
    exact = torch.cat([yf.sum(0), (yf * yf).sum(0)]).cpu().numpy()
    scale = np.concatenate([np.abs(yf.cpu().numpy()).sum(0), (yf * yf).sum(0).cpu().numpy()]) + 1e-30
    assert np.max(np.abs(sums.cpu().numpy() - exact) / scale) < 2e-6      # fp32 partials, fp64 totals
    assert np.max(np.abs(ref.cpu().numpy() - exact) / scale) < 2e-6
    # the oracle's convolution agrees with the stored output
    yref = tf_ops.conv2d(torch.tensor(x), torch.tensor(wt), (s, s), "SAME", (1, 1))
    if bias:
        yref = yref + torch.tensor(b)
    assert rel_l2(y1.float().cpu(), yref) < 6e-3


STEM_CASES = [
    # n, h, w, k, stride_h, cout
    (2, 64, 64, 7, 2, 64),
    (3, 30, 34, 3, 2, 64),
    (1, 224, 224, 7, 2, 64),
    (5, 17, 22, 7, 2, 128),       # odd height, ragged last tile, two dy atoms
]


@pytest.mark.parametrize("case", STEM_CASES)
def test_stem_gather_convolution(L, case):
    """mcn_pad_rgb4 + mcn_stem_conv_fprop / _wgrad (RGB stem without an im2col matrix) vs the
    oracle's conv2d and its autograd weight gradient; fused BN statistics checked as well."""
    from myconvnet_b200.plan import Plan
    n, h, w, k, st, co = case
    rng = np.random.default_rng(5)
    r = lambda a: torch.tensor(a).bfloat16().float().numpy()
    x = r(rng.standard_normal((n, h, w, 3)).astype(np.float32))
    wt = r((rng.standard_normal((k, k, 3, co)) * 0.1).astype(np.float32))
    b = rng.standard_normal(co).astype(np.float32)
    d, ho, wo = desc_for(L, x.shape, wt.shape, st, 1, "SAME")
    d4 = L.ConvDescC(n, h, w, 4, co, k, k, st, st, 1, 1, d.pad_t, d.pad_l, ho, wo)
    lib = L.load()
    kpad = lib.mcn_stem_conv_kpad(d4)
    assert kpad == Plan.stem_kpad(k, k) and kpad > 0
    rows = np.array([(rr * (k + 1) + ss) * 4 + c for rr in range(k) for ss in range(k) for c in range(3)])
    wst = np.zeros((kpad, co), np.float32)
    wst[rows] = wt.reshape(-1, co)
    dy = r(rng.standard_normal((n, ho, wo, co)).astype(np.float32))
    xt = torch.tensor(x)
    wtt = torch.tensor(wt, requires_grad=True)
    yref = tf_ops.conv2d(xt, wtt, (st, st), "SAME", (1, 1)) + torch.tensor(b)
    yref.backward(torch.tensor(dy))
    xd, dyd, wd, bd = dev(x, torch.bfloat16), dev(dy, torch.bfloat16), dev(wst), dev(b)
    x4 = torch.full((n, h, w, 4), 9.0, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_pad_rgb4(xd.data_ptr(), n * h * w, x4.data_ptr(), None))
    w_t = torch.empty(co, kpad, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_weight_prep(wd.data_ptr(), 1, kpad, co, None, w_t.data_ptr(), None))
    y = torch.empty(n, ho, wo, co, device="cuda", dtype=torch.bfloat16)
    sums = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_stem_conv_fprop(d4, x4.data_ptr(), w_t.data_ptr(), bd.data_ptr(), y.data_ptr(),
                                    sums.data_ptr(), None))
    y2 = torch.empty_like(y)
    L.check(lib.mcn_stem_conv_fprop(d4, x4.data_ptr(), w_t.data_ptr(), bd.data_ptr(), y2.data_ptr(), None, None))
    dw = torch.zeros(kpad, co, device="cuda")
    L.check(lib.mcn_stem_conv_wgrad(d4, x4.data_ptr(), dyd.data_ptr(), dw.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(x4[..., :3], xd) and float(x4[..., 3].abs().max()) == 0.0
    assert rel_l2(y.float().cpu(), yref.detach()) < 6e-3
    assert torch.equal(y, y2)
    yf = y.double().reshape(-1, co)
    exact = torch.cat([yf.sum(0), (yf * yf).sum(0)]).cpu().numpy()
    scale = np.concatenate([np.abs(yf.cpu().numpy()).sum(0), (yf * yf).sum(0).cpu().numpy()]) + 1e-30
    assert np.max(np.abs(sums.cpu().numpy() - exact) / scale) < 2e-6
    dwh = dw.cpu().numpy()
    assert rel_l2(dwh[rows].reshape(k, k, 3, co), wtt.grad) < 1e-4
    other = np.setdiff1d(np.arange(kpad), rows)
    assert np.all(dwh[other] == 0.0)          # widening tap / 4th channel / padding rows carry no gradient


@pytest.mark.parametrize("code", [0, 1])
@pytest.mark.parametrize("C", [64, 20, 320])
@pytest.mark.parametrize("use_res", [False, True])
def test_bn_apply_from_sums_matches_finalize_then_apply(L, code, C, use_res):
    """mcn_bn_apply_stats == mcn_bn_finalize + mcn_bn_apply (saved statistics, moving statistics,
    output), and agrees with the oracle's fused_batch_norm."""
    lib = L.load()
    rng = np.random.default_rng(4)
    rows, eps, mom = 3 * 9 * 11, 1e-3, 0.9
    dt = tdt(code)
    q = (lambda a: torch.tensor(a).to(dt).float().numpy())
    x = q((rng.standard_normal((rows, C)) * 2 + 0.7).astype(np.float32))
    res = q(rng.standard_normal((rows, C)).astype(np.float32))
    gamma = rng.uniform(0.5, 1.5, C).astype(np.float32)
    beta = rng.standard_normal(C).astype(np.float32)
    mm = rng.standard_normal(C).astype(np.float32)
    mv = rng.uniform(0.5, 2, C).astype(np.float32)
    xd, rd, gd, bd = dev(x, dt), dev(res, dt), dev(gamma), dev(beta)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_bn_stats(code, xd.data_ptr(), rows, C, sums.data_ptr(), None))
    outs = []
    for fused in (False, True):
        mmd, mvd = dev(mm), dev(mv)
        save = torch.zeros(2 * C, device="cuda")
        y = torch.empty_like(xd)
        rp = rd.data_ptr() if use_res else None
        if fused:
            L.check(lib.mcn_bn_apply_stats(code, xd.data_ptr(), rows, C, sums.data_ptr(), float(rows), eps, mom,
                                           gd.data_ptr(), bd.data_ptr(), rp, 1, 0.0, y.data_ptr(),
                                           save.data_ptr(), save.data_ptr() + 4 * C, mmd.data_ptr(),
                                           mvd.data_ptr(), None))
        else:
            L.check(lib.mcn_bn_finalize(sums.data_ptr(), float(rows), C, eps, mom, save.data_ptr(),
                                        save.data_ptr() + 4 * C, mmd.data_ptr(), mvd.data_ptr(), None))
            L.check(lib.mcn_bn_apply(code, xd.data_ptr(), rows, C, save.data_ptr(), save.data_ptr() + 4 * C,
                                     gd.data_ptr(), bd.data_ptr(), rp, 1, 0.0, y.data_ptr(), None))
        torch.cuda.synchronize()
        outs.append((y.float().cpu().numpy(), save.cpu().numpy(), mmd.cpu().numpy(), mvd.cpu().numpy()))
    (y0, s0, m0, v0), (y1, s1, m1, v1) = outs
    assert rel_l2(s1, s0) < 1e-6 and rel_l2(m1, m0) < 1e-6 and rel_l2(v1, v0) < 1e-6
    assert rel_l2(y1, y0) < (1e-6 if code == 0 else 2e-3)
    yb, bm, bv = tf_ops.fused_batch_norm_train(torch.tensor(x), torch.tensor(gamma), torch.tensor(beta), eps)
    yref = torch.relu(yb + torch.tensor(res) if use_res else yb)
    assert rel_l2(y1, yref) < (2e-5 if code == 0 else 8e-3)
    assert rel_l2(s1[:C], bm) < 1e-5
    assert rel_l2(v1, mom * mv + (1 - mom) * bv.numpy()) < 1e-5


@pytest.mark.parametrize("code", [0, 1])
@pytest.mark.parametrize("C", [64, 24, 20, 320])
@pytest.mark.parametrize("variant", ["plain", "relu", "relu_res", "swish", "res_noact"])
def test_batch_norm(L, code, C, variant):
    lib = L.load()
    rng = np.random.default_rng(2)
    rows, eps, mom = 6 * 5 * 7, 1e-3, 0.9
    dt = tdt(code)
    q = (lambda a: torch.tensor(a).to(dt).float().numpy())
    x = q((rng.standard_normal((rows, C)) * 2 + 0.7).astype(np.float32))
    res = q(rng.standard_normal((rows, C)).astype(np.float32))
    dy = q(rng.standard_normal((rows, C)).astype(np.float32))
    gamma = rng.uniform(0.5, 1.5, C).astype(np.float32)
    beta = rng.standard_normal(C).astype(np.float32)
    mm = rng.standard_normal(C).astype(np.float32)
    mv = rng.uniform(0.5, 2, C).astype(np.float32)
    act = {"plain": 0, "relu": 1, "relu_res": 1, "swish": 6, "res_noact": 0}[variant]
    use_res = variant in ("relu_res", "res_noact")
    xt = torch.tensor(x, requires_grad=True)
    rt = torch.tensor(res, requires_grad=True)
    gt = torch.tensor(gamma, requires_grad=True)
    bt = torch.tensor(beta, requires_grad=True)
    yb, bm, bv = tf_ops.fused_batch_norm_train(xt, gt, bt, eps)
    pre = yb + rt if use_res else yb
    yref = tf_ops.activation(pre, {0: None, 1: "relu", 6: "swish"}[act])
    yref.backward(torch.tensor(dy))
    xd, rd, dyd = dev(x, dt), dev(res, dt), dev(dy, dt)
    gd, bd, mmd, mvd = dev(gamma), dev(beta), dev(mm), dev(mv)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    save = torch.empty(2 * C, device="cuda")
    y = torch.empty_like(xd)
    L.check(lib.mcn_bn_stats(code, xd.data_ptr(), rows, C, sums.data_ptr(), None))
    L.check(lib.mcn_bn_finalize(sums.data_ptr(), float(rows), C, eps, mom, save.data_ptr(),
                                save.data_ptr() + 4 * C, mmd.data_ptr(), mvd.data_ptr(), None))
    L.check(lib.mcn_bn_apply(code, xd.data_ptr(), rows, C, save.data_ptr(), save.data_ptr() + 4 * C,
                             gd.data_ptr(), bd.data_ptr(), rd.data_ptr() if use_res else None, act, 0.0,
                             y.data_ptr(), None))
    s1 = torch.zeros(C, device="cuda")
    s2 = torch.zeros(C, device="cuda")
    # derivative from the output when a residual is fused with an activation, else recompute
    yptr = y.data_ptr() if (use_res and act != 0) else None
    L.check(lib.mcn_bn_bwd_reduce(code, dyd.data_ptr(), xd.data_ptr(), yptr, rows, C, save.data_ptr(),
                                  save.data_ptr() + 4 * C, gd.data_ptr(), bd.data_ptr(), act, 0.0,
                                  s1.data_ptr(), s2.data_ptr(), None))
    dx = torch.empty_like(xd)
    dr = torch.empty_like(xd)
    L.check(lib.mcn_bn_bwd_apply(code, dyd.data_ptr(), xd.data_ptr(), yptr, rows, C, save.data_ptr(),
                                 save.data_ptr() + 4 * C, gd.data_ptr(), bd.data_ptr(), act, 0.0,
                                 s1.data_ptr(), s2.data_ptr(), float(rows), dx.data_ptr(),
                                 dr.data_ptr() if use_res else None, None))
    torch.cuda.synchronize()
    tol = 2e-5 if code == 0 else 8e-3
    assert rel_l2(save[:C].cpu(), bm.detach()) < 1e-5
    assert rel_l2(y.float().cpu(), yref.detach()) < tol
    assert rel_l2(mmd.cpu(), mom * mm + (1 - mom) * bm.detach().numpy()) < 1e-5
    assert rel_l2(mvd.cpu(), mom * mv + (1 - mom) * bv.detach().numpy()) < 1e-5
    assert rel_l2(dx.float().cpu(), xt.grad) < (1e-4 if code == 0 else 2e-2)
    assert rel_l2(s1.cpu(), bt.grad) < (1e-4 if code == 0 else 2e-2)
    assert rel_l2(s2.cpu(), gt.grad) < (1e-4 if code == 0 else 2e-2)
    if use_res:
        assert rel_l2(dr.float().cpu(), rt.grad) < (1e-5 if code == 0 else 1e-2)


@pytest.mark.parametrize("code", [0, 1])
def test_pools_gap_resize(L, code):
    lib = L.load()
    rng = np.random.default_rng(4)
    dt = tdt(code)
    n, h, w, c = 2, 9, 12, 16
    x = torch.tensor(rng.standard_normal((n, h, w, c)).astype(np.float32)).to(dt).float()
    xd = x.cuda().to(dt)
    tol = 1e-5 if code == 0 else 6e-3
    for k, s, pad in [(3, 2, "SAME"), (2, 2, "VALID"), (5, 1, "SAME")]:
        ho, pt, _ = tf_ops.same_pad(h, k, s, 1, pad)
        wo, pl, _ = tf_ops.same_pad(w, k, s, 1, pad)
        dy = torch.tensor(rng.standard_normal((n, ho, wo, c)).astype(np.float32)).to(dt).float()
        dyd = dy.cuda().to(dt)
        for kind in ("max", "avg"):
            xt = x.clone().requires_grad_(True)
            yref = (tf_ops.max_pool if kind == "max" else tf_ops.avg_pool)(xt, [k, k], [s, s], pad)
            yref.backward(dy)
            y = torch.empty(n, ho, wo, c, device="cuda", dtype=dt)
            dx = torch.empty_like(xd)
            if kind == "max":
                am = torch.empty(n, ho, wo, c, device="cuda", dtype=torch.int32)
                L.check(lib.mcn_maxpool_fwd(code, xd.data_ptr(), n, h, w, c, k, k, s, s, pt, pl, ho, wo,
                                            y.data_ptr(), am.data_ptr(), None))
                L.check(lib.mcn_maxpool_bwd(code, dyd.data_ptr(), am.data_ptr(), n, h, w, c, k, k, s, s,
                                            pt, pl, ho, wo, dx.data_ptr(), None))
            else:
                L.check(lib.mcn_avgpool_fwd(code, xd.data_ptr(), n, h, w, c, k, k, s, s, pt, pl, ho, wo,
                                            y.data_ptr(), None))
                L.check(lib.mcn_avgpool_bwd(code, dyd.data_ptr(), n, h, w, c, k, k, s, s, pt, pl, ho, wo,
                                            dx.data_ptr(), None))
            torch.cuda.synchronize()
            assert rel_l2(y.float().cpu(), yref.detach()) < tol, (kind, k, s, pad)
            assert rel_l2(dx.float().cpu(), xt.grad) < tol, (kind, k, s, pad)
    # global average pool
    xt = x.clone().requires_grad_(True)
    yref = tf_ops.global_avg_pool(xt)
    g = torch.tensor(rng.standard_normal((n, c)).astype(np.float32))
    yref.backward(g)
    y = torch.empty(n, c, device="cuda", dtype=dt)
    dx = torch.empty_like(xd)
    gd = g.cuda().to(dt)
    L.check(lib.mcn_gap_fwd(code, xd.data_ptr(), n, h * w, c, y.data_ptr(), code, None))
    L.check(lib.mcn_gap_bwd(code, gd.data_ptr(), code, n, h * w, c, dx.data_ptr(), None))
    torch.cuda.synchronize()
    assert rel_l2(y.float().cpu(), yref.detach()) < tol
    assert rel_l2(dx.float().cpu(), xt.grad) < (1e-5 if code == 0 else 1e-2)
    # bilinear resize, the three coordinate modes
    for mode, (ac, hp) in enumerate([(False, False), (True, False), (False, True)]):
        for (ho, wo) in [(18, 24), (13, 7), (9, 12)]:
            xt = x.clone().requires_grad_(True)
            yref = tf_ops.resize_bilinear(xt, [ho, wo], ac, hp)
            dy = torch.tensor(rng.standard_normal((n, ho, wo, c)).astype(np.float32)).to(dt).float()
            yref.backward(dy)
            y = torch.empty(n, ho, wo, c, device="cuda", dtype=dt)
            dx = torch.empty_like(xd)
            dyd = dy.cuda().to(dt)
            L.check(lib.mcn_resize_bilinear_fwd(code, xd.data_ptr(), n, h, w, c, ho, wo, mode, y.data_ptr(), None))
            L.check(lib.mcn_resize_bilinear_bwd(code, dyd.data_ptr(), n, h, w, c, ho, wo, mode, dx.data_ptr(), None))
            torch.cuda.synchronize()
            assert rel_l2(y.float().cpu(), yref.detach()) < tol, (mode, ho, wo)
            assert rel_l2(dx.float().cpu(), xt.grad) < (1e-5 if code == 0 else 1e-2), (mode, ho, wo)


@pytest.mark.parametrize("code", [0, 1])
@pytest.mark.parametrize("case", [(2, 9, 9, 16, 1, 3, 1), (2, 12, 12, 24, 1, 5, 2), (1, 8, 8, 8, 2, 3, 2)])
def test_depthwise(L, code, case):
    lib = L.load()
    n, h, w, c, mult, k, s = case
    rng = np.random.default_rng(5)
    dt = tdt(code)
    x = torch.tensor(rng.standard_normal((n, h, w, c)).astype(np.float32)).to(dt).float()
    wt = torch.tensor((rng.standard_normal((k, k, c, mult)) * 0.3).astype(np.float32))
    ho, pt, _ = tf_ops.same_pad(h, k, s, 1, "SAME")
    wo, pl, _ = tf_ops.same_pad(w, k, s, 1, "SAME")
    dy = torch.tensor(rng.standard_normal((n, ho, wo, c * mult)).astype(np.float32)).to(dt).float()
    xt, wtt = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    yref = tf_ops.depthwise_conv2d(xt, wtt, (s, s), "SAME", (1, 1))
    yref.backward(dy)
    d = L.ConvDescC(n, h, w, c, c, k, k, s, s, 1, 1, pt, pl, ho, wo)
    xd, dyd, wd = x.cuda().to(dt), dy.cuda().to(dt), wt.cuda()
    y = torch.empty(n, ho, wo, c * mult, device="cuda", dtype=dt)
    dx = torch.empty_like(xd)
    dw = torch.zeros_like(wd)
    L.check(lib.mcn_dwconv2d_fwd(d, mult, code, xd.data_ptr(), 0, wd.data_ptr(), y.data_ptr(), None))
    L.check(lib.mcn_dwconv2d_bwd_data(d, mult, code, dyd.data_ptr(), 0, wd.data_ptr(), dx.data_ptr(), None))
    L.check(lib.mcn_dwconv2d_bwd_filter(d, mult, code, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), None))
    torch.cuda.synchronize()
    tol = 1e-5 if code == 0 else 6e-3
    assert rel_l2(y.float().cpu(), yref.detach()) < tol
    assert rel_l2(dx.float().cpu(), xt.grad) < tol
    assert rel_l2(dw.cpu(), wtt.grad) < 1e-5


def test_softmax_xent_and_sigmoid_xent(L):
    lib = L.load()
    rng = np.random.default_rng(6)
    rows, C = 37, 21
    z = torch.tensor((rng.standard_normal((rows, C)) * 3).astype(np.float32), requires_grad=True)
    y = rng.integers(-1, C, size=rows).astype(np.int32)      # -1 rows are invalid
    cw = rng.uniform(0.5, 2.0, C).astype(np.float32)
    for ls in (0.0, 0.1):
        z.grad = None
        ref = tf_ops.classification_loss(z, torch.tensor(y).long(), C, cw, ls)
        ref.backward()
        zd, yd, cwd = z.detach().cuda(), dev(y), dev(cw)
        loss = torch.zeros(3, dtype=torch.int64, device="cuda")     # xsum accumulator
        dl = torch.empty(rows, C, device="cuda")
        pr = torch.empty(rows, C, device="cuda")
        L.check(lib.mcn_softmax_xent(zd.data_ptr(), yd.data_ptr(), rows, C, cwd.data_ptr(), ls, 0.0, 0.0, 0, 0,
                                     1.0 / rows, loss.data_ptr(), dl.data_ptr(), pr.data_ptr(), None))
        torch.cuda.synchronize()
        assert abs(L.xsum_value(loss.cpu().numpy()) / rows - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6
        assert rel_l2(dl.cpu(), z.grad) < 1e-5
        assert rel_l2(pr.cpu(), torch.softmax(z.detach(), -1)) < 1e-5
    x = torch.tensor(rng.standard_normal(50).astype(np.float32) * 4, requires_grad=True)
    for label in (0.0, 1.0, 0.9):
        x.grad = None
        ref = tf_ops.sigmoid_cross_entropy(x, torch.full_like(x, label)).mean()
        ref.backward()
        loss = torch.zeros(3, dtype=torch.int64, device="cuda")
        dl = torch.empty(50, device="cuda")
        L.check(lib.mcn_sigmoid_xent(x.detach().cuda().data_ptr(), 50, label, 1.0, 1.0 / 50, loss.data_ptr(),
                                     dl.data_ptr(), 0, None))
        torch.cuda.synchronize()
        assert abs(L.xsum_value(loss.cpu().numpy()) / 50 - ref.item()) < 1e-5
        assert rel_l2(dl.cpu(), x.grad) < 1e-5


def test_maxpool_argmax_bit_exact():
    """Pooling argmax indices are bit-exact against the oracle's first-max-in-window rule."""
    import ctypes
    from myconvnet_b200 import lib as L
    from oracle import tf_ops
    lib = L.load()
    rng = np.random.default_rng(3)
    # quantised values make ties frequent
    x = (rng.integers(0, 4, size=(2, 9, 11, 16)).astype(np.float32)) * 0.5
    xt = torch.from_numpy(x).cuda()
    for k, s, pad in [(3, 2, "SAME"), (2, 2, "VALID"), (3, 1, "SAME")]:
        ho, pt, _ = tf_ops.same_pad(9, k, s, 1, pad)
        wo, pl, _ = tf_ops.same_pad(11, k, s, 1, pad)
        y = torch.empty(2, ho, wo, 16, device="cuda")
        am = torch.empty(2, ho, wo, 16, dtype=torch.int32, device="cuda")
        L.check(lib.mcn_maxpool_fwd(0, xt.data_ptr(), 2, 9, 11, 16, k, k, s, s, pt, pl, ho, wo,
                                    y.data_ptr(), am.data_ptr(), None))
        torch.cuda.synchronize()
        ref_idx = tf_ops.max_pool_argmax(torch.from_numpy(x), [k, k], [s, s], pad).numpy()
        ref_val = tf_ops.max_pool(torch.from_numpy(x), [k, k], [s, s], pad).numpy()
        assert np.array_equal(am.cpu().numpy().astype(np.int64), ref_idx)
        assert np.array_equal(y.cpu().numpy(), ref_val)
        # compact (tap) form of the training step: same values, same argmax after expansion, and its
        # backward equals the int32-argmax backward bit for bit (fp32 and bf16)
        for code, dt in ((0, torch.float32), (1, torch.bfloat16)):
            xc = xt.to(dt)
            y2 = torch.empty(2, ho, wo, 16, device="cuda", dtype=dt)
            tap = torch.empty(2, ho, wo, 16, dtype=torch.uint8, device="cuda")
            am2 = torch.empty(2, ho, wo, 16, dtype=torch.int32, device="cuda")
            L.check(lib.mcn_maxpool_fwd_tap(code, xc.data_ptr(), 2, 9, 11, 16, k, k, s, s, pt, pl, ho, wo,
                                            y2.data_ptr(), tap.data_ptr(), None))
            L.check(lib.mcn_maxpool_tap_to_argmax(tap.data_ptr(), 2, 9, 11, 16, k, k, s, s, pt, pl, ho, wo,
                                                  am2.data_ptr(), None))
            gy = torch.randn(2, ho, wo, 16, device="cuda").to(dt)
            dx1, dx2 = torch.empty_like(xc), torch.empty_like(xc)
            L.check(lib.mcn_maxpool_bwd(code, gy.data_ptr(), am.data_ptr(), 2, 9, 11, 16, k, k, s, s, pt, pl, ho, wo,
                                        dx1.data_ptr(), None))
            L.check(lib.mcn_maxpool_bwd_tap(code, gy.data_ptr(), tap.data_ptr(), 2, 9, 11, 16, k, k, s, s, pt, pl,
                                            ho, wo, dx2.data_ptr(), None))
            torch.cuda.synchronize()
            assert np.array_equal(am2.cpu().numpy().astype(np.int64), ref_idx)
            assert torch.equal(y2.float(), y)
            assert torch.equal(dx1, dx2)


@pytest.mark.parametrize("strip_rows", [None, 1, 3, 100])
@pytest.mark.parametrize("shape", [(2, 9, 11, 16), (3, 12, 10, 8), (2, 16, 14, 24), (1, 7, 8, 8), (2, 112, 112, 64)])
def test_maxpool_3x3s2_strip_kernels(shape, strip_rows, monkeypatch):
    """The strip-walking 3x3 / stride-2 max-pool kernels (forward: one output column per thread, the
    shared input row carried in registers; backward: one 2-pixel input column pair per thread) give
    the oracle's values and first-max argmax bit for bit, and the int32-argmax backward's gradient
    bit for bit, for odd / even sizes (TF SAME puts the odd padding at the bottom / right) and any
    strip length."""
    from myconvnet_b200 import lib as L
    lib = L.load()
    if strip_rows is not None:
        monkeypatch.setenv("MCN_POOL_STRIP_ROWS", str(strip_rows))
    n, h, w, c = shape
    rng = np.random.default_rng(5)
    x = (rng.integers(0, 6, size=shape).astype(np.float32)) * 0.5 - 1.0       # frequent ties
    k, s = 3, 2
    ho, pt, _ = tf_ops.same_pad(h, k, s, 1, "SAME")
    wo, pl, _ = tf_ops.same_pad(w, k, s, 1, "SAME")
    ref_idx = tf_ops.max_pool_argmax(torch.from_numpy(x), [k, k], [s, s], "SAME").numpy()
    ref_val = tf_ops.max_pool(torch.from_numpy(x), [k, k], [s, s], "SAME")
    xt = torch.from_numpy(x).cuda()
    for code, dt in ((0, torch.float32), (1, torch.bfloat16)):
        xc = xt.to(dt)
        y = torch.empty(n, ho, wo, c, device="cuda", dtype=dt)
        tap = torch.empty(n, ho, wo, c, dtype=torch.uint8, device="cuda")
        am = torch.empty(n, ho, wo, c, dtype=torch.int32, device="cuda")
        L.check(lib.mcn_maxpool_fwd_tap(code, xc.data_ptr(), n, h, w, c, k, k, s, s, pt, pl, ho, wo,
                                        y.data_ptr(), tap.data_ptr(), None))
        L.check(lib.mcn_maxpool_tap_to_argmax(tap.data_ptr(), n, h, w, c, k, k, s, s, pt, pl, ho, wo,
                                              am.data_ptr(), None))
        gy = torch.randn(n, ho, wo, c, device="cuda").to(dt)
        dx1, dx2 = torch.empty_like(xc), torch.full_like(xc, 7.0)
        L.check(lib.mcn_maxpool_bwd(code, gy.data_ptr(), am.data_ptr(), n, h, w, c, k, k, s, s, pt, pl, ho, wo,
                                    dx1.data_ptr(), None))
        L.check(lib.mcn_maxpool_bwd_tap(code, gy.data_ptr(), tap.data_ptr(), n, h, w, c, k, k, s, s, pt, pl,
                                        ho, wo, dx2.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(am.cpu().numpy().astype(np.int64), ref_idx)
        assert torch.equal(y.float().cpu(), ref_val)          # the test values are exact in bf16
        assert torch.equal(dx1, dx2)
