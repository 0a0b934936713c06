"""GPU parity of the other north-star model files, loaded UNCHANGED: EfficientNet-B0 (depthwise
conv, SE, swish) and DeepLabv3+ on the dilated ResNet (atrous conv, ASPP concat, bilinear
upsampling, per-pixel loss).  fp32 path: tight activation/loss check; bf16 path: tensor-core
kernels, loss and loss curve within 3 %.  See tests/test_gpu_resnet.py for the tolerance rationale."""
import numpy as np
import pytest

from myconvnet_b200 import loader
from myconvnet_b200.engine import draw_initial_value
from tests.util import layerwise_backward_errors, layerwise_forward_errors, rel_l2, worst

pytestmark = pytest.mark.gpu


def _pair(rel_path, cls, shape, ncls, batch, dtype, oracle_facade, **kw):
    pm = getattr(loader.load_reference_model(rel_path, loader.product_facade()), cls)(
        shape, ncls, batch_size=batch, compute_dtype=dtype, **kw)
    rng = np.random.default_rng(0)
    vals = {v.name: draw_initial_value(v, rng) for v in pm.graph.vars.values()}
    for k in vals:
        if k.endswith("gamma") and not vals[k].any():
            vals[k] = np.full_like(vals[k], 0.5)
    om = getattr(loader.load_reference_model(rel_path, oracle_facade), cls)(
        shape, ncls, oracle_round_bf16=(dtype == "bf16"), **kw)
    om.set_variables(vals)
    return pm, om, vals


def _oracle_facade():
    from oracle import ref_convnet, ref_segnet
    return {"convnet": ref_convnet, "segmentation.segnet": ref_segnet}


def _check(pm, om, vals, X, Y, dtype, steps=3, curve_tol=None):
    from myconvnet_b200.engine import Engine
    from oracle.step import OracleTrainer
    taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k != "pred"}
    eng = Engine(pm, keep=list(taps.values()))
    eng.set_variables(vals)
    loss_dev = eng.train_step(X, Y, update=False)
    tr = OracleTrainer(om, batch_size=len(X))
    tr.step(X, Y, update=False)
    assert abs(loss_dev - float(om.data_loss)) <= (1e-3 if dtype == "f32" else 3e-2) * abs(float(om.data_loss)) + 1e-4
    if dtype == "f32":
        bad = [(k, rel_l2(eng.fetch(t), om.d[k].t.detach().numpy())) for k, t in taps.items()
               if om.d.get(k) is not None]
        bad = [b for b in bad if b[1] > 5e-4]
        assert not bad, bad[:6]
    grads = eng.get_gradients()
    assert all(np.isfinite(g).all() for g in grads.values())
    # the production (fused) plan, layer by layer with teacher forcing, forward and backward: every
    # conv / depthwise conv / dense / BN(+activation) / pool node against the oracle op at the
    # device's own tensors (tests/util.py) — one layer's arithmetic per comparison
    eng3 = Engine(pm, keep_grads=True)
    eng3.set_variables(vals)
    eng3.train_step(X, Y, update=False)
    ferr = layerwise_forward_errors(eng3, pm)
    berr = layerwise_backward_errors(eng3, pm)
    assert ferr and worst(ferr, 1)[0][1] <= (5e-5 if dtype == "f32" else 6e-3), worst(ferr)
    assert berr and worst(berr, 1)[0][1] <= (2e-4 if dtype == "f32" else 1e-2), worst(berr)
    del eng3
    # fused plan over a few optimiser steps
    eng2 = Engine(pm)
    eng2.set_variables(vals)
    tr2 = OracleTrainer(om, batch_size=len(X))
    om.set_variables(vals)
    tol = curve_tol or (1e-2 if dtype == "f32" else 4e-2)
    for _ in range(steps):
        a, b = eng2.train_step(X, Y), tr2.step(X, Y)
        assert np.isfinite(a) and abs(a - b) <= tol * abs(b) + tol, (a, b)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_efficientnet_b0(have_reference_models, dtype):
    shape, ncls, batch = [64, 64, 3], 10, 8
    pm, om, vals = _pair("models/efficientnet.py", "EfficientNetB0", shape, ncls, batch, dtype,
                         _oracle_facade(), base_learning_rate=0.05)
    rng = np.random.default_rng(1)
    X = rng.uniform(size=[batch] + shape).astype(np.float32)
    Y = rng.integers(0, ncls, size=batch).astype(np.int32)
    _check(pm, om, vals, X, Y, dtype)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_deeplabv3plus(have_reference_models, dtype):
    shape, ncls, batch = [64, 64, 3], 5, 2
    pm, om, vals = _pair("models/deeplabv3plus.py", "DeepLabV3PlusResNet", shape, ncls, batch, dtype,
                         _oracle_facade(), base_learning_rate=0.05)
    rng = np.random.default_rng(2)
    X = rng.uniform(size=[batch] + shape).astype(np.float32)
    Y = rng.integers(0, ncls + 1, size=[batch] + shape[:2]).astype(np.int32)    # 0 = ignore
    # batch of 2 with BN over 4x4 maps: a single flipped ReLU moves the loss by ~1 % after a step
    _check(pm, om, vals, X, Y, dtype, steps=2, curve_tol=4e-2)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_dcgan_simultaneous_update(have_reference_models, dtype):
    """models/dcgan.py unchanged: D(real)/G/D(fake) with shared D variables, two losses, D and G
    gradients from the same forward, one Adam update (optimizers_gan.py:56-58,110-120)."""
    from myconvnet_b200.engine import Engine
    from oracle import ref_convnet, ref_gan
    from oracle.step import OracleTrainer
    shape, latent, batch = [64, 64, 3], 100, 8
    kw = dict(base_learning_rate=5e-4 * 32, momentum=0.5, generator_scaling_factor=2.0, label_smoothing=0.1)
    facade = {"convnet": ref_convnet, "generative.gan": ref_gan}
    pm, om, vals = _pair("models/dcgan.py", "DCGAN", shape, latent, batch, dtype, facade, **kw)
    rng = np.random.default_rng(3)
    X = rng.uniform(size=[batch] + shape).astype(np.float32)
    Z = rng.uniform(-1, 1, size=(batch, latent)).astype(np.float32)
    assert pm.generate.shape == (batch, 64, 64, 3)           # 4x4 seed
    eng = Engine(pm, optimizer="adam", keep=[pm.generate, pm.logits_real, pm.logits_fake])
    eng.set_variables(vals)
    tr = OracleTrainer(om, optimizer="adam", batch_size=batch)
    tol = 2e-3 if dtype == "f32" else 5e-2
    # gradients of both passes on the first (identical-weights) step
    eng.train_step(X, Z, update=False)
    tr.step(X, Z, update=False)
    assert rel_l2(eng.fetch(pm.generate), om.d["generate"].t.detach().numpy()) < (1e-4 if dtype == "f32" else 3e-2)
    ld, lg = eng.last_losses
    assert abs(ld - tr.last_losses[0]) < tol * abs(tr.last_losses[0]) + tol
    assert abs(lg - tr.last_losses[1]) < tol * abs(tr.last_losses[1]) + tol
    if dtype == "f32":
        grads = eng.get_gradients()
        errs = {k: rel_l2(grads[k], g.numpy()) for k, g in tr.grads.items() if float(g.norm()) > 1e-7}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        assert worst[0][1] < 8e-2, worst       # LeakyReLU/ReLU kink sensitivity, see test_gpu_resnet
    for _ in range(3):
        eng.train_step(X, Z)
        tr.step(X, Z)
        ld, lg = eng.last_losses
        assert np.isfinite(ld) and np.isfinite(lg)
        assert abs(ld - tr.last_losses[0]) < 3 * tol * abs(tr.last_losses[0]) + 3 * tol, (ld, tr.last_losses)
        assert abs(lg - tr.last_losses[1]) < 3 * tol * abs(tr.last_losses[1]) + 3 * tol, (lg, tr.last_losses)
