"""GPU parity of the round-2 additions, through the C ABI / the engine, against the oracle:
the 23 distinct ResNet-50 convolution shapes of BASELINE config 1 at N = 32,
determinism of the reductions, the random train-time ops (dropout, stochastic depth) against the
same Philox stream regenerated on the host, nearest-neighbour resize, the focal / L1 / spatially
smoothed losses, the l1 / pseudo-Huber weight decay and frozen batch-norm."""
import numpy as np
import pytest
import torch

from oracle import philox, tf_ops
from tests.test_gpu_ops import desc_for
from tests.util import build_pair, rel_l2, synthetic_batch, worst

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from myconvnet_b200 import lib
    lib.load()
    lib.ensure_workspace(256 << 20)
    return lib


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def bf16_round(a):
    return torch.tensor(a).bfloat16().float().numpy()


# (H_in, Cin, k, stride, Cout): SURVEY Appendix B.1, minus the RGB stem (test_stem_gather_convolution)
R50_SHAPES = [
    (56, 64, 1, 1, 64), (56, 64, 1, 1, 256), (56, 64, 3, 1, 64), (56, 256, 1, 1, 64), (56, 256, 1, 1, 128),
    (56, 256, 1, 2, 512), (56, 128, 3, 2, 128),
    (28, 128, 1, 1, 512), (28, 128, 3, 1, 128), (28, 512, 1, 1, 128), (28, 512, 1, 1, 256),
    (28, 512, 1, 2, 1024), (28, 256, 3, 2, 256),
    (14, 256, 1, 1, 1024), (14, 256, 3, 1, 256), (14, 1024, 1, 1, 256), (14, 1024, 1, 1, 512),
    (14, 1024, 1, 2, 2048), (14, 512, 3, 2, 512),
    (7, 512, 1, 1, 2048), (7, 512, 3, 1, 512), (7, 2048, 1, 1, 512),
]


@pytest.mark.parametrize("shape", R50_SHAPES, ids=lambda s: "%dx%d_c%d_k%ds%d_o%d" % (s[0], s[0], s[1], s[2], s[3], s[4]))
def test_resnet50_conv_shapes_at_batch_32(L, shape):
    """fprop (+ fused statistics), dgrad and wgrad of every distinct ResNet-50 convolution at the
    size of BASELINE config 1 (N = 32), in the mode the planner uses (2: halo tiles where they pay,
    else im2col / box TMA), vs the oracle's conv2d and its autograd on bf16-rounded operands."""
    h, ci, k, s, co = shape
    n = 32
    lib = L.load()
    rng = np.random.default_rng(7)
    x = bf16_round(rng.standard_normal((n, h, h, ci)).astype(np.float32))
    wt = bf16_round((rng.standard_normal((k, k, ci, co)) * (1.0 / np.sqrt(k * k * ci))).astype(np.float32))
    ho, pt, _ = tf_ops.same_pad(h, k, s, 1, "SAME")
    d = L.ConvDescC(n, h, h, ci, co, k, k, s, s, 1, 1, pt, pt, ho, ho)
    dy = bf16_round(rng.standard_normal((n, ho, ho, co)).astype(np.float32))
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    yref = tf_ops.conv2d(xt, wtt, (s, s), "SAME", (1, 1))
    yref.backward(torch.tensor(dy))
    xd, dyd, wd = dev(x, torch.bfloat16), dev(dy, torch.bfloat16), dev(wt)
    w_hwio = torch.empty(k * k, ci, co, device="cuda", dtype=torch.bfloat16)
    w_ohwi = torch.empty(k * k, co, ci, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_weight_prep(wd.data_ptr(), k * k, ci, co, w_hwio.data_ptr(), w_ohwi.data_ptr(), None))
    y = torch.empty(n, ho, ho, co, device="cuda", dtype=torch.bfloat16)
    sums = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_conv2d_fprop_tc_stats(d, xd.data_ptr(), w_ohwi.data_ptr(), None, y.data_ptr(), 2,
                                          sums.data_ptr(), None))
    dx = torch.empty(n, h, h, ci, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_conv2d_dgrad_tc(d, dyd.data_ptr(), w_hwio.data_ptr(), dx.data_ptr(), 1, 2, 0, None))
    dws = []
    for _ in range(2):
        dw = torch.zeros(k, k, ci, co, device="cuda")
        L.check(lib.mcn_conv2d_wgrad_tc(d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 2, None))
        dws.append(dw)
    torch.cuda.synchronize()
    assert rel_l2(y.float().cpu(), yref.detach()) < 4e-3          # one bf16 rounding of the output
    assert rel_l2(dx.float().cpu(), xt.grad) < 4e-3
    assert rel_l2(dws[0].cpu(), wtt.grad) < 1e-4                  # fp32 accumulate over 32*Ho*Wo pixels
    assert torch.equal(dws[0], dws[1])                            # ordered split-K: bit-reproducible
    yf = y.double().reshape(-1, co)
    exact = torch.cat([yf.sum(0), (yf * yf).sum(0)]).cpu().numpy()
    scale = np.concatenate([np.abs(yf.cpu().numpy()).sum(0), (yf * yf).sum(0).cpu().numpy()]) + 1e-30
    assert np.max(np.abs(sums.cpu().numpy() - exact) / scale) < 2e-6


def test_workspace_is_required_and_left_clean(L):
    """Reductions fail loudly without a registered workspace and leave it zeroed (self-cleaning)."""
    lib = L.load()
    x = torch.randn(64, 256, device="cuda")
    sums = torch.zeros(512, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_set_workspace(None, 0))
    assert lib.mcn_bn_stats(0, x.data_ptr(), 64, 256, sums.data_ptr(), None) != 0
    assert b"workspace" in lib.mcn_last_error()
    ws = L.ensure_workspace(256 << 20)
    L.check(lib.mcn_bn_stats(0, x.data_ptr(), 64, 256, sums.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.allclose(sums[:256].float(), x.sum(0), atol=1e-4)
    assert int(ws[: 1 << 20].max()) == 0                    # counters and limbs are zero again


def test_xsum_accumulator_is_exact_and_order_independent(L):
    """The fixed-point accumulator behind every reduction: the loss kernel's sum over 20000 rows is
    the exactly rounded sum of its per-block partials whatever the block order, so two launches with
    different grids of the same rows agree to the last bit once the partials are the same — here:
    decode on the device (mcn_xsum_decode) == decode on the host, and catastrophic cancellation
    (1e8 + 1 - 1e8) survives."""
    lib = L.load()
    limbs = torch.zeros(3, dtype=torch.int64, device="cuda")
    logits = torch.zeros(1, 2, device="cuda")
    lab = torch.zeros(1, dtype=torch.int32, device="cuda")
    # CE of equal logits is ln 2 per row; weight it by +-1e8 and 1 through class weights
    total = 0.0
    for w in (1e8, 1.0, -1e8):
        cw = torch.tensor([w, w], dtype=torch.float32, device="cuda")
        L.check(lib.mcn_softmax_xent(logits.data_ptr(), lab.data_ptr(), 1, 2, cw.data_ptr(), 0.0, 0.0, 0.0, 0, 0,
                                     1.0, limbs.data_ptr(), None, None, None))
        total += w
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    L.check(lib.mcn_xsum_decode(limbs.data_ptr(), 1, None, out.data_ptr(), 0, None))
    torch.cuda.synchronize()
    ln2 = float(np.float32(np.log(np.float32(2.0))))
    assert abs(out.item() - ln2) < 1e-5 * ln2           # a float sum would have returned 0 or 8
    assert out.item() == L.xsum_value(limbs.cpu().numpy())


@pytest.mark.parametrize("code", [0, 1])
def test_dropout_and_stochastic_depth_use_the_philox_stream(L, code):
    lib = L.load()
    dt = torch.float32 if code == 0 else torch.bfloat16
    rng = np.random.default_rng(9)
    hp = torch.zeros(16, dtype=torch.float32)
    seed, step = 1234, 7
    hp.view(torch.int32)[12] = seed
    hp.view(torch.int32)[13] = step
    hpd = hp.cuda()
    n = 1003
    x = bf16_round(rng.standard_normal(n).astype(np.float32))
    xd = dev(x, dt)
    y = torch.empty_like(xd)
    rate, layer = 0.3, 5
    L.check(lib.mcn_dropout(code, xd.data_ptr(), n, rate, hpd.data_ptr(), layer, y.data_ptr(), None))
    keep = philox.dropout_keep(n, rate, seed, step, layer)
    ref = np.where(keep, x / np.float32(1.0 - rate), 0.0).astype(np.float32)
    assert 0.6 < keep.mean() < 0.8
    assert rel_l2(y.float().cpu().numpy(), ref) < (1e-6 if code == 0 else 4e-3)
    assert np.array_equal(y.float().cpu().numpy() != 0, keep & (x != 0))
    # stochastic depth: y = relu(a*s[n] + b); backward from the output
    N, per = 16, 24
    a = bf16_round(rng.standard_normal((N, per)).astype(np.float32))
    b = bf16_round(rng.standard_normal((N, per)).astype(np.float32))
    gy = bf16_round(rng.standard_normal((N, per)).astype(np.float32))
    s = philox.survive(N, rate, seed, step, layer).astype(np.float32)[:, None] / np.float32(1.0 - rate)
    at, bt = torch.tensor(a, requires_grad=True), torch.tensor(b, requires_grad=True)
    yref = torch.relu(at * torch.tensor(s) + bt)
    yref.backward(torch.tensor(gy))
    ad, bd, gd = dev(a, dt), dev(b, dt), dev(gy, dt)
    yd = torch.empty_like(ad)
    L.check(lib.mcn_sd_add_fwd(code, ad.data_ptr(), bd.data_ptr(), N, per, rate, hpd.data_ptr(), layer, 1, 0.0,
                               yd.data_ptr(), None))
    da, db = torch.empty_like(ad), torch.empty_like(ad)
    L.check(lib.mcn_sd_add_bwd(code, gd.data_ptr(), yd.data_ptr(), N, per, rate, hpd.data_ptr(), layer, 1, 0.0,
                               da.data_ptr(), db.data_ptr(), None))
    torch.cuda.synchronize()
    tol = 1e-6 if code == 0 else 6e-3
    assert rel_l2(yd.float().cpu(), yref.detach()) < tol
    if code == 0:
        assert rel_l2(da.cpu(), at.grad) < tol and rel_l2(db.cpu(), bt.grad) < tol
    assert 0 < (s == 0).sum() < N                                     # some samples dropped, some kept


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("size", [((5, 7), (10, 14)), ((6, 6), (15, 9)), ((8, 5), (8, 5))])
def test_resize_nearest(L, mode, size):
    lib = L.load()
    (h, w), (ho, wo) = size
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, h, w, 6)).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    ref = tf_ops.resize_nearest(xt, [ho, wo], mode == 1, mode == 2)
    gy = rng.standard_normal((2, ho, wo, 6)).astype(np.float32)
    ref.backward(torch.tensor(gy))
    xd, gd = dev(x), dev(gy)
    y = torch.empty(2, ho, wo, 6, device="cuda")
    dx = torch.empty_like(xd)
    L.check(lib.mcn_resize_nearest_fwd(0, xd.data_ptr(), 2, h, w, 6, ho, wo, mode, y.data_ptr(), None))
    L.check(lib.mcn_resize_nearest_bwd(0, gd.data_ptr(), 2, h, w, 6, ho, wo, mode, dx.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), ref.detach().numpy())
    assert rel_l2(dx.cpu(), xt.grad) < 1e-6


@pytest.mark.parametrize("opts", [dict(focal_gamma=2.0), dict(sigmoid_focal_alpha=4.0), dict(label_smoothing=0.1),
                                  dict(focal_gamma=1.5, sigmoid_focal_alpha=2.0, label_smoothing=0.05),
                                  dict(label_smoothing=0.2, spatial=True), dict(spatial=True, label_smoothing=0.1,
                                                                                focal_gamma=2.0)])
def test_loss_variants(L, opts):
    """Focal / sigmoid-focal factors (convnet.py:580-592), uniform and 5x5 spatial label smoothing
    (convnet.py:603-607, segmentation/segnet.py:116-121), class weights, ignored rows."""
    lib = L.load()
    rng = np.random.default_rng(12)
    spatial = opts.pop("spatial", False)
    C = 7
    shape = (2, 9, 11) if spatial else (53,)
    rows = int(np.prod(shape))
    z = torch.tensor((rng.standard_normal(shape + (C,)) * 2).astype(np.float32), requires_grad=True)
    y = rng.integers(-1, C, size=shape).astype(np.int32)
    cw = rng.uniform(0.5, 2.0, C).astype(np.float32)
    ref = tf_ops.classification_loss(z, torch.tensor(y).long(), C, cw, opts.get("label_smoothing", 0.0),
                                     focal_gamma=opts.get("focal_gamma", 0.0),
                                     sigmoid_focal_alpha=opts.get("sigmoid_focal_alpha", 0.0),
                                     spatial_smoothing=spatial)
    ref.backward()
    zd, yd, cwd = z.detach().cuda(), dev(y), dev(cw)
    loss = torch.zeros(3, dtype=torch.int64, device="cuda")
    dl = torch.empty(rows, C, device="cuda")
    L.check(lib.mcn_softmax_xent(zd.data_ptr(), yd.data_ptr(), rows, C, cwd.data_ptr(), opts.get("label_smoothing", 0.0),
                                 opts.get("focal_gamma", 0.0), opts.get("sigmoid_focal_alpha", 0.0),
                                 shape[1] if spatial else 0, shape[2] if spatial else 0, 1.0 / rows,
                                 loss.data_ptr(), dl.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert abs(L.xsum_value(loss.cpu().numpy()) / rows - ref.item()) < 2e-5 * abs(ref.item())
    assert rel_l2(dl.cpu().numpy().reshape(z.shape), z.grad) < 2e-5


@pytest.mark.parametrize("opts", [dict(), dict(label_smoothing=0.1), dict(label_smoothing=0.2, spatial=True),
                                  dict(spatial=True, label_smoothing=0.1, focal_gamma=2.0, sigmoid_focal_alpha=3.0),
                                  dict(spatial=True)])
@pytest.mark.parametrize("C", [21, 16, 5])
def test_loss_thread_per_row_kernel(L, opts, C, monkeypatch):
    """The thread-per-row softmax-CE kernel (C <= 47 and >= 4096 rows: segmentation heads) against the
    oracle and against the warp-per-row kernel: loss, gradient and probabilities, with class weights,
    ignored rows, uniform and 5x5 spatial label smoothing, focal factors, odd and even class counts."""
    lib = L.load()
    opts = dict(opts)
    rng = np.random.default_rng(13)
    spatial = opts.pop("spatial", False)
    shape = (2, 50, 47) if spatial else (4700,)
    rows = int(np.prod(shape))
    z = torch.tensor((rng.standard_normal(shape + (C,)) * 2).astype(np.float32), requires_grad=True)
    y = rng.integers(-1, C, size=shape).astype(np.int32)
    cw = rng.uniform(0.5, 2.0, C).astype(np.float32)
    ls = opts.get("label_smoothing", 0.0)
    ref = tf_ops.classification_loss(z, torch.tensor(y).long(), C, cw, ls, focal_gamma=opts.get("focal_gamma", 0.0),
                                     sigmoid_focal_alpha=opts.get("sigmoid_focal_alpha", 0.0),
                                     spatial_smoothing=spatial)
    ref.backward()
    zd, yd, cwd = z.detach().cuda(), dev(y), dev(cw)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MCN_XENT_ROWS", mode)
        loss = torch.zeros(3, dtype=torch.int64, device="cuda")
        dl = torch.full((rows, C), 7.0, device="cuda")
        pr = torch.full((rows, C), 7.0, device="cuda")
        args = (rows, C, cwd.data_ptr(), ls, opts.get("focal_gamma", 0.0), opts.get("sigmoid_focal_alpha", 0.0),
                shape[1] if spatial else 0, shape[2] if spatial else 0, 1.0 / rows)
        L.check(lib.mcn_softmax_xent(zd.data_ptr(), yd.data_ptr(), *args, loss.data_ptr(), dl.data_ptr(), None, None))
        L.check(lib.mcn_softmax_xent(zd.data_ptr(), None, *args, None, None, pr.data_ptr(), None))
        torch.cuda.synchronize()
        res[mode] = (L.xsum_value(loss.cpu().numpy()) / rows, dl.cpu().numpy(), pr.cpu().numpy())
    for mode in ("1", "0"):
        lv, dl, pr = res[mode]
        assert abs(lv - ref.item()) < 2e-5 * abs(ref.item()), mode
        assert rel_l2(dl.reshape(z.shape), z.grad) < 2e-5, mode
        assert rel_l2(pr.reshape(z.shape), torch.softmax(z.detach(), -1)) < 1e-6, mode
    assert rel_l2(res["1"][1], res["0"][1]) < 1e-6 and abs(res["1"][0] - res["0"][0]) < 1e-6 * abs(res["0"][0])


SMALL_SHAPE, SMALL_BATCH, SMALL_NCLS = [128, 128, 3], 16, 16     # block_4 BN over 256 values per channel


def _small(dtype="f32", **kw):
    return build_pair("models/resnet_v1_5.py", "ResNet50", SMALL_SHAPE, SMALL_NCLS, SMALL_BATCH, dtype,
                      base_learning_rate=0.05, **kw)


def _one_step(pm, om, vals, X, Y, **okw):
    from myconvnet_b200.engine import Engine
    from oracle.step import OracleTrainer
    from tests.test_gpu_resnet import relu_pattern
    eng = Engine(pm, **okw)
    eng.set_variables(vals)
    loss_dev = eng.train_step(X, Y)
    om.forced_relu_masks = relu_pattern(eng, pm, vals)
    tr = OracleTrainer(om, **okw)
    loss_ref = tr.step(X, Y)
    new = eng.get_variables()
    uerr = {k: rel_l2(new[k] - vals[k], om.vars[k].detach().numpy() - vals[k]) for k in vals
            if np.linalg.norm(om.vars[k].detach().numpy() - vals[k]) > 1e-9}
    return eng, tr, loss_dev, loss_ref, uerr


@pytest.mark.parametrize("kw", [dict(l1_reg=1e-5), dict(base_weight_decay=1e-3, l1_weight_decay=True),
                                dict(base_weight_decay=1e-3, huber_decay_delta=0.05),
                                dict(focal_loss_factor=2.0, label_smoothing=0.1)],
                         ids=["l1_reg", "l1_weight_decay", "huber_weight_decay", "focal"])
def test_regularisers_and_loss_options_in_a_step(have_reference_models, kw):
    pm, om, vals = _small(**kw)
    X, Y = synthetic_batch(SMALL_BATCH, SMALL_SHAPE, SMALL_NCLS)
    eng, tr, a, b, uerr = _one_step(pm, om, vals, X, Y)
    assert abs(a - b) <= 2e-5 * abs(b), (a, b)
    assert worst(uerr, 1)[0][1] <= 3e-3, worst(uerr)


@pytest.mark.parametrize("dtype,tol", [("f32", 3e-3), ("bf16", 0.3)])
def test_stochastic_depth_and_dropout_in_a_step(have_reference_models, dtype, tol):
    """initial/final_drop_rate (resnet_v1_5.py:56-58) and dropout_rate (resnet_v1_5.py:75) > 0: the
    step matches the oracle drawing the same masks; a second step draws different ones; inference
    ignores them.  The fp32 run is the parity test (3e-3 on every update); the bf16 run checks the
    same masks reach the tensor-core plan — whole-step bf16 updates are only a sanity bound, see
    tests/test_gpu_resnet.py for why (the kernels themselves are held to 6e-3 in
    test_dropout_and_stochastic_depth_use_the_philox_stream)."""
    kw = dict(initial_drop_rate=0.1, final_drop_rate=0.4, dropout_rate=0.3)
    return _sd_body(dtype, tol, kw)


def _sd_body(dtype, tol, kw):
    pm, om, vals = _small(dtype, **kw)
    assert any(n.op == "sd_add" for n in pm.graph.nodes) and any(n.op == "dropout" for n in pm.graph.nodes)
    X, Y = synthetic_batch(SMALL_BATCH, SMALL_SHAPE, SMALL_NCLS)
    eng, tr, a, b, uerr = _one_step(pm, om, vals, X, Y)
    assert abs(a - b) <= (2e-5 if dtype == "f32" else 2e-2) * abs(b), (a, b)
    assert worst(uerr, 1)[0][1] <= tol, worst(uerr)
    # the masks are a function of the step index: same step -> same loss bit for bit, next step ->
    # other masks -> another loss from the SAME weights; inference ignores them
    eng.set_variables(vals)
    l0, l0b = eng.train_step(X, Y, update=False), eng.train_step(X, Y, update=False)
    eng.global_step = 1
    l1 = eng.train_step(X, Y, update=False)
    assert l0 == l0b and abs(l1 - l0) > 1e-4, (l0, l0b, l1)
    p1, p2 = eng.predict(X), eng.predict(X)
    assert np.array_equal(p1, p2)


def test_frozen_blocks_use_moving_statistics(have_reference_models):
    """blocks_to_train (convnet.py:1384-1389, 1781-1795, 1916-1924): variables of the other blocks
    get no update, their batch-norm layers normalise with the stored moving statistics (and leave
    them untouched) while training; gradients still flow through them to earlier trainable blocks."""
    pm, om, vals = _small(blocks_to_train=[1, 3, None])
    rng = np.random.default_rng(5)
    for k in vals:                      # non-trivial stored statistics
        if k.endswith("/mu"):
            vals[k] = rng.normal(0, 0.3, vals[k].shape).astype(np.float32)
        if k.endswith("/sigma"):
            vals[k] = rng.uniform(0.5, 2.0, vals[k].shape).astype(np.float32)
    om.set_variables(vals)
    X, Y = synthetic_batch(SMALL_BATCH, SMALL_SHAPE, SMALL_NCLS)
    eng, tr, a, b, uerr = _one_step(pm, om, vals, X, Y)
    assert abs(a - b) <= 2e-5 * abs(b), (a, b)
    assert worst(uerr, 1)[0][1] <= 3e-3, worst(uerr)
    new = eng.get_variables()
    frozen = [k for k in vals if k.startswith(("block_0/", "block_2/", "block_4/"))]
    assert frozen and all(np.array_equal(new[k], vals[k]) for k in frozen)
    assert any(k.startswith("block_1/") for k in uerr) and any(k.startswith("block_3/") for k in uerr)


# EfficientNet-B0 depthwise layers (H_in, C, k, stride): SURVEY Appendix B.2
B0_DW = [(112, 32, 3, 1), (112, 96, 3, 2), (56, 144, 3, 1), (56, 144, 5, 2), (28, 240, 5, 1), (28, 240, 3, 2),
         (14, 480, 3, 1), (14, 480, 5, 1), (14, 672, 5, 1), (14, 672, 5, 2), (7, 1152, 5, 1), (7, 1152, 3, 1)]


@pytest.mark.parametrize("code", [0, 1])
@pytest.mark.parametrize("case", B0_DW, ids=lambda c: "%dx%d_c%d_k%ds%d" % (c[0], c[0], c[1], c[2], c[3]))
def test_efficientnet_depthwise_shapes(L, code, case):
    """The register-tiled depthwise kernels (csrc/dwconv.cu) at every EfficientNet-B0 depthwise shape
    (batch 4): forward, backward-data, backward-filter vs tf.nn.depthwise_conv2d's restatement and
    its autograd; the filter gradient is bit-reproducible (ordered slices, no atomics)."""
    lib = L.load()
    h, c, k, s = case
    n = 4
    dt = torch.float32 if code == 0 else torch.bfloat16
    rng = np.random.default_rng(21)
    rd = bf16_round if code == 1 else (lambda a: a)
    x = rd(rng.standard_normal((n, h, h, c)).astype(np.float32))
    wt = (rng.standard_normal((k, k, c, 1)) * 0.3).astype(np.float32)
    ho, pt, _ = tf_ops.same_pad(h, k, s, 1, "SAME")
    d = L.ConvDescC(n, h, h, c, c, k, k, s, s, 1, 1, pt, pt, ho, ho)
    dy = rd(rng.standard_normal((n, ho, ho, c)).astype(np.float32))
    xt, wtt = torch.tensor(x, requires_grad=True), torch.tensor(wt, requires_grad=True)
    yref = tf_ops.depthwise_conv2d(xt, wtt, (s, s), "SAME", (1, 1))
    yref.backward(torch.tensor(dy))
    xd, dyd, wd = dev(x, dt), dev(dy, dt), dev(wt)
    y = torch.empty(n, ho, ho, c, device="cuda", dtype=dt)
    dx = torch.empty_like(xd)
    L.check(lib.mcn_dwconv2d_fwd(d, 1, code, xd.data_ptr(), 0, wd.data_ptr(), y.data_ptr(), None))
    L.check(lib.mcn_dwconv2d_bwd_data(d, 1, code, dyd.data_ptr(), 0, wd.data_ptr(), dx.data_ptr(), None))
    dws = []
    for _ in range(2):
        dw = torch.zeros(k, k, c, 1, device="cuda")
        L.check(lib.mcn_dwconv2d_bwd_filter(d, 1, code, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), None))
        dws.append(dw)
    torch.cuda.synchronize()
    tol = 2e-6 if code == 0 else 4e-3
    assert rel_l2(y.float().cpu(), yref.detach()) < tol
    assert rel_l2(dx.float().cpu(), xt.grad) < tol
    assert rel_l2(dws[0].cpu(), wtt.grad) < 2e-5
    assert torch.equal(dws[0], dws[1])


# ------------------------------------------------------------------ group norm / weight standardisation
@pytest.mark.parametrize("code", [0, 1])
@pytest.mark.parametrize("case", [(3, (7, 5), 64, 32), (2, (), 96, 8), (4, (9, 9), 40, 5)],
                         ids=["hw35_c64_g32", "dense_c96_g8", "hw81_c40_g5"])
def test_group_norm_kernels(L, code, case):
    """mcn_gn_fwd / mcn_gn_bwd against the oracle's group_norm (reference convnet.py:1928-2013) and its
    autograd: output, input gradient, dgamma, dbeta; a second backward gives bit-identical results."""
    n, hw, c, g = case
    rng = np.random.default_rng(5)
    shape = (n,) + tuple(hw) + (c,)
    x = rng.standard_normal(shape).astype(np.float32) * 1.5 + 0.3
    dy = rng.standard_normal(shape).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, c).astype(np.float32)
    beta = rng.standard_normal(c).astype(np.float32)
    dt = torch.float32 if code == 0 else torch.bfloat16
    if code == 1:
        x, dy = bf16_round(x), bf16_round(dy)
    xt, gt, bt = (torch.tensor(a, requires_grad=True) for a in (x, gamma, beta))
    yref = tf_ops.group_norm(xt, gt, bt, g, 1e-3)
    yref.backward(torch.tensor(dy))
    lib = L.load()
    HW = int(np.prod(hw)) if hw else 1
    xd, dyd, gd, bd = dev(x, dt), dev(dy, dt), dev(gamma), dev(beta)
    y = torch.empty_like(xd)
    save = torch.zeros(n * g * 2, device="cuda")
    L.check(lib.mcn_gn_fwd(code, xd.data_ptr(), n, HW, c, g, 1e-3, gd.data_ptr(), bd.data_ptr(), y.data_ptr(),
                           save.data_ptr(), None))
    outs = []
    for _ in range(2):
        dx = torch.empty_like(xd)
        dgam, dbet = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        scratch = torch.zeros(2 * n * g + 2 * n * c, device="cuda")
        L.check(lib.mcn_gn_bwd(code, dyd.data_ptr(), xd.data_ptr(), n, HW, c, g, gd.data_ptr(), save.data_ptr(),
                               scratch.data_ptr(), dx.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), None))
        outs.append((dx, dgam, dbet))
    torch.cuda.synchronize()
    tol = 2e-5 if code == 0 else 6e-3
    assert rel_l2(y.float().cpu(), yref.detach()) < tol
    assert rel_l2(outs[0][0].float().cpu(), xt.grad) < tol
    assert rel_l2(outs[0][1].cpu(), gt.grad) < 2e-5 and rel_l2(outs[0][2].cpu(), bt.grad) < 2e-5
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))


@pytest.mark.parametrize("wshape", [(3, 3, 16, 24), (1, 1, 64, 256), (72, 10), (5, 5, 8, 1)])
def test_weight_standardisation_kernels(L, wshape):
    """mcn_ws_fwd / mcn_ws_bwd against the oracle's restatement of convnet.py:1410-1419 and its autograd
    (the statistic is per LAST axis, population std + 1e-5)."""
    rng = np.random.default_rng(6)
    w = (rng.standard_normal(wshape) * 0.1 + 0.02).astype(np.float32)
    g = rng.standard_normal(wshape).astype(np.float32)
    wt = torch.tensor(w, requires_grad=True)
    ref = tf_ops.weight_standardization(wt)
    ref.backward(torch.tensor(g))
    lib = L.load()
    cols = wshape[-1]
    rows = w.size // cols
    wd, gd = dev(w), dev(g)
    ws = torch.empty_like(wd)
    stats = torch.zeros(2 * cols, device="cuda")
    L.check(lib.mcn_ws_fwd(wd.data_ptr(), rows, cols, 1e-5, ws.data_ptr(), stats.data_ptr(), None))
    base = rng.standard_normal(wshape).astype(np.float32)
    grad = dev(base)
    L.check(lib.mcn_ws_bwd(gd.data_ptr(), wd.data_ptr(), stats.data_ptr(), rows, cols, 1e-5, grad.data_ptr(), None))
    torch.cuda.synchronize()
    assert rel_l2(ws.cpu(), ref.detach()) < 2e-6
    assert rel_l2(grad.cpu().numpy() - base, wt.grad) < 2e-5        # accumulates into the gradient


@pytest.mark.parametrize("dtype,tol", [("f32", 3e-3), ("bf16", 0.35)])
def test_wsgn_resnet_step(have_reference_models, dtype, tol):
    """models/resnet_v1_5_wsgn.py, unchanged: every convolution runs on standardised weights
    (tensor-core routes on per-step bf16 copies, the RGB stem on the CUDA-core route) and every
    normalisation is a group norm.  One optimiser step against the oracle: loss and every update
    (fp32 is the parity run; bf16 checks the tensor-core plan end to end, see test_gpu_resnet.py)."""
    shape, batch = ([64, 64, 3], 8) if dtype == "f32" else (SMALL_SHAPE, SMALL_BATCH)
    pm, om, vals = build_pair("models/resnet_v1_5_wsgn.py", "ResNet50", shape, SMALL_NCLS, batch, dtype,
                              base_learning_rate=0.05)
    hist = {}
    from myconvnet_b200.plan import Plan
    for l in Plan(pm.graph).fwd:
        hist[l.fn] = hist.get(l.fn, 0) + 1
    assert hist["mcn_ws_fwd"] == 53 and hist["mcn_gn_fwd"] == 53 and "mcn_bn_apply_stats" not in hist
    X, Y = synthetic_batch(batch, shape, SMALL_NCLS)
    eng, tr, a, b, uerr = _one_step(pm, om, vals, X, Y)
    assert abs(a - b) <= (2e-5 if dtype == "f32" else 2e-2) * abs(b), (a, b)
    assert len(uerr) >= 150 and worst(uerr, 1)[0][1] <= tol, worst(uerr)
    # inference runs the same standardisation on the EMA weights
    p = eng.predict(X)
    assert p.shape == (batch, SMALL_NCLS) and np.isfinite(p).all() and np.allclose(p.sum(-1), 1.0, atol=1e-3)


# ------------------------------------------------------------------ BN backward sums in the dgrad epilogue
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("case", [(8, 28, 128, 256, 1), (4, 56, 64, 64, 3), (8, 14, 256, 256, 3), (3, 20, 64, 128, 3)],
                         ids=["1x1_28_128<-256", "3x3_halo_56_64<-64", "3x3_im2col_14_256<-256", "3x3_ragged_20_64<-128"])
def test_dgrad_with_fused_bn_backward_sums(L, case, act):
    """mcn_conv2d_dgrad_tc_bnred + mcn_bn_bwd_finalize against the separate launches they replace
    (mcn_conv2d_dgrad_tc, then mcn_bn_bwd_reduce on its output): dx bit-identical, the two sum vectors
    equal up to fp32 summation order (the fused ones are exact sums of the same products), twice."""
    n, hw, ci, co, k = case
    lib = L.load()
    L.ensure_workspace(64 << 20)
    rng = np.random.default_rng(11)
    d, ho, wo = desc_for(L, (n, hw, hw, ci), (k, k, ci, co), 1, 1, "SAME")
    dy = dev(rng.standard_normal((n, ho, wo, co)).astype(np.float32), torch.bfloat16)
    w_hwio = dev((rng.standard_normal((k * k, ci, co)) * 0.05).astype(np.float32), torch.bfloat16)
    x = dev(rng.standard_normal((n, hw, hw, ci)).astype(np.float32) * 2 + 0.5, torch.bfloat16)   # the BN input
    mean = dev(rng.standard_normal(ci).astype(np.float32) * 0.3)
    invstd = dev(rng.uniform(0.5, 1.5, ci).astype(np.float32))
    gamma = dev(rng.uniform(0.5, 1.5, ci).astype(np.float32))
    beta = dev(rng.standard_normal(ci).astype(np.float32) * 0.3)
    dx_ref = torch.empty(n, hw, hw, ci, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mcn_conv2d_dgrad_tc(d, dy.data_ptr(), w_hwio.data_ptr(), dx_ref.data_ptr(), 1, 2, 0, None))
    s_ref = torch.zeros(2, ci, device="cuda")
    L.check(lib.mcn_bn_bwd_reduce(1, dx_ref.data_ptr(), x.data_ptr(), None, n * hw * hw, ci, mean.data_ptr(),
                                  invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, 0.0,
                                  s_ref[0].data_ptr(), s_ref[1].data_ptr(), None))
    assert lib.mcn_conv2d_dgrad_bnred_supported(d, 2) == 1
    outs = []
    for _ in range(2):
        dx = torch.full_like(dx_ref, 3.0)
        sums = torch.zeros(2 * ci, device="cuda", dtype=torch.float64)
        L.check(lib.mcn_conv2d_dgrad_tc_bnred(d, dy.data_ptr(), w_hwio.data_ptr(), dx.data_ptr(), 2, x.data_ptr(),
                                              mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                              beta.data_ptr(), act, sums.data_ptr(), None, None, None))
        s = torch.zeros(2, ci, device="cuda")
        L.check(lib.mcn_bn_bwd_finalize(sums.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ci,
                                        s[0].data_ptr(), s[1].data_ptr(), None))
        # the variant whose last block writes the final sums itself: same values, bit for bit
        dx3 = torch.full_like(dx_ref, 5.0)
        sums3 = torch.zeros(2 * ci, device="cuda", dtype=torch.float64)
        s3 = torch.zeros(2, ci, device="cuda")
        L.check(lib.mcn_conv2d_dgrad_tc_bnred(d, dy.data_ptr(), w_hwio.data_ptr(), dx3.data_ptr(), 2, x.data_ptr(),
                                              mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                              beta.data_ptr(), act, sums3.data_ptr(), s3[0].data_ptr(),
                                              s3[1].data_ptr(), None))
        torch.cuda.synchronize()
        assert torch.equal(dx3, dx) and torch.equal(s3, s)
        outs.append((dx, s))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], dx_ref)
    scale = float(s_ref.abs().max())
    assert float((outs[0][1] - s_ref).abs().max()) <= 2e-5 * scale + 1e-4, (outs[0][1] - s_ref).abs().max()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


# ------------------------------------------------------------------ ReLU bit mask of the residual BN layers
@pytest.mark.parametrize("shape", [(4, 14, 14, 256), (2, 7, 9, 64), (3, 5, 5, 2048), (1, 3, 3, 8), (2, 9, 7, 96),
                                   (2, 5, 5, 1152)])
def test_bn_relu_bit_mask_variants_equal_the_output_reading_ones(L, shape):
    """mcn_bn_apply_stats_mask writes the same y / saved statistics as mcn_bn_apply_stats plus one bit
    per element (y > 0); mcn_bn_bwd_reduce_mask / mcn_bn_bwd_apply_mask reading that mask give exactly
    what mcn_bn_bwd_reduce / mcn_bn_bwd_apply give reading y (bf16, fused residual, ReLU)."""
    lib = L.load()
    n, h, w, c = shape
    rows = n * h * w
    rng = np.random.default_rng(21)
    x = dev(rng.standard_normal(shape).astype(np.float32), torch.bfloat16)
    res = dev(rng.standard_normal(shape).astype(np.float32), torch.bfloat16)
    gy = dev(rng.standard_normal(shape).astype(np.float32), torch.bfloat16)
    gamma = dev(rng.uniform(0.5, 1.5, c).astype(np.float32))
    beta = dev(rng.standard_normal(c).astype(np.float32) * 0.2)
    sums = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
    L.check(lib.mcn_bn_stats(1, x.data_ptr(), rows, c, sums.data_ptr(), None))
    outs = []
    for use_mask in (False, True):
        y = torch.empty_like(x)
        save = torch.zeros(2, c, device="cuda")
        mm, mv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        mask = torch.full(((rows * c // 8 + 3) // 4 * 4,), 0xAA, device="cuda", dtype=torch.uint8)
        if use_mask:
            L.check(lib.mcn_bn_apply_stats_mask(1, x.data_ptr(), rows, c, sums.data_ptr(), float(rows), 1e-3, 0.9,
                                                gamma.data_ptr(), beta.data_ptr(), res.data_ptr(), 1, 0.0,
                                                y.data_ptr(), mask.data_ptr(), save[0].data_ptr(),
                                                save[1].data_ptr(), mm.data_ptr(), mv.data_ptr(), None))
        else:
            L.check(lib.mcn_bn_apply_stats(1, x.data_ptr(), rows, c, sums.data_ptr(), float(rows), 1e-3, 0.9,
                                           gamma.data_ptr(), beta.data_ptr(), res.data_ptr(), 1, 0.0, y.data_ptr(),
                                           save[0].data_ptr(), save[1].data_ptr(), mm.data_ptr(), mv.data_ptr(),
                                           None))
        s = torch.zeros(2, c, device="cuda")
        dx, dres = torch.empty_like(x), torch.empty_like(x)
        if use_mask:
            L.check(lib.mcn_bn_bwd_reduce_mask(1, gy.data_ptr(), x.data_ptr(), mask.data_ptr(), rows, c,
                                               save[0].data_ptr(), save[1].data_ptr(), s[0].data_ptr(),
                                               s[1].data_ptr(), None))
            L.check(lib.mcn_bn_bwd_apply_mask(1, gy.data_ptr(), x.data_ptr(), mask.data_ptr(), rows, c,
                                              save[0].data_ptr(), save[1].data_ptr(), gamma.data_ptr(),
                                              s[0].data_ptr(), s[1].data_ptr(), float(rows), dx.data_ptr(),
                                              dres.data_ptr(), None))
        else:
            L.check(lib.mcn_bn_bwd_reduce(1, gy.data_ptr(), x.data_ptr(), y.data_ptr(), rows, c, save[0].data_ptr(),
                                          save[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, 0.0,
                                          s[0].data_ptr(), s[1].data_ptr(), None))
            L.check(lib.mcn_bn_bwd_apply(1, gy.data_ptr(), x.data_ptr(), y.data_ptr(), rows, c, save[0].data_ptr(),
                                         save[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, 0.0,
                                         s[0].data_ptr(), s[1].data_ptr(), float(rows), dx.data_ptr(),
                                         dres.data_ptr(), None))
        torch.cuda.synchronize()
        outs.append((y, save, mm, mv, s, dx, dres, mask))
    a, b = outs
    for i in range(7):
        assert torch.equal(a[i], b[i]), i
    # the mask itself: bit e & 7 of byte e >> 3 <=> y[e] > 0
    bits = (b[0].float().reshape(-1) > 0).to(torch.uint8).reshape(-1, 8)
    weights = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], device="cuda", dtype=torch.uint8)
    assert torch.equal((bits * weights).sum(1).to(torch.uint8), b[7][:rows * c // 8])


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("case", [(2, 9, 7, 16, 36, 28, torch.bfloat16), (1, 8, 8, 21, 32, 32, torch.float32),
                                  (2, 12, 10, 8, 5, 7, torch.float32), (1, 5, 6, 24, 17, 11, torch.bfloat16)])
def test_resize_backward_table_kernel_equals_candidate_scan(L, case, mode, monkeypatch):
    """The table-driven bilinear-resize backward (per-axis weight tables in shared memory, 16-byte channel
    vectors where C allows) is bit-identical to the coordinate-per-candidate kernel it replaces, for the
    three coordinate conventions, up- and down-sampling, vector and scalar channel counts."""
    lib = L.load()
    n, h, w, c, ho, wo, dt = case
    code = 0 if dt == torch.float32 else 1
    rng = np.random.default_rng(31)
    dy = dev(rng.standard_normal((n, ho, wo, c)).astype(np.float32), dt)
    outs = []
    for tab in ("1", "0"):
        monkeypatch.setenv("MCN_RESIZE_TABLE", tab)
        dx = torch.full((n, h, w, c), 3.0, device="cuda", dtype=dt)
        L.check(lib.mcn_resize_bilinear_bwd(code, dy.data_ptr(), n, h, w, c, ho, wo, mode, dx.data_ptr(), None))
        torch.cuda.synchronize()
        outs.append(dx)
    assert torch.equal(outs[0], outs[1])
    # and the vectorised forward agrees with the adjoint identity <resize(x), dy> == <x, resize_bwd(dy)>
    x = dev(rng.standard_normal((n, h, w, c)).astype(np.float32), dt)
    y = torch.empty(n, ho, wo, c, device="cuda", dtype=dt)
    L.check(lib.mcn_resize_bilinear_fwd(code, x.data_ptr(), n, h, w, c, ho, wo, mode, y.data_ptr(), None))
    torch.cuda.synchronize()
    lhs = float((y.double() * dy.double()).sum())
    rhs = float((x.double() * outs[0].double()).sum())
    # bf16: both sides carry the rounding of ~numel stored values of unit scale
    tol = 8e-3 * float(np.sqrt(y.numel() + x.numel())) if code else 1e-4 * (abs(lhs) + abs(rhs) + 1.0)
    assert abs(lhs - rhs) <= tol
