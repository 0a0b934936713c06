"""GPU parity: the reference's models/resnet_v1_5.py, unchanged, on the B200 engine vs the CPU
oracle — per-layer activations through the model's own d[...] taps, weight gradients, loss, and a
short loss curve (SURVEY.md 8c pins 4).

Tolerances (stated, with their reason):
  fp32 activations  rel-L2 <= 2e-4 per tap, loss 1e-3 relative.
  fp32 gradients    rel-L2 <= 8e-2 per variable.  A ReLU network's gradient is discontinuous in
      its pre-activations: two fp32 implementations whose activations agree to 3e-5 still disagree
      on the sign of a handful of near-zero pre-activations, and each flipped mask bit changes the
      gradient of a 32-pixel-per-channel layer by ~1 %.  tests/test_oracle.py shows the SAME 1-2 %
      between the fp32 and fp64 oracle, so this is the floor for any model-level comparison; the
      kernels themselves are held to 1e-4..1e-5 on gradients in tests/test_gpu_ops.py.
  bf16 activations  rel-L2 <= 6e-2 up to block_2 and <= 0.35 after (two bf16 pipelines drift apart
      by an ulp-sized random walk amplified by small-batch BN; the oracle rounds at the same
      points but accumulates in a different order), loss 3 %, 4-step loss curve 3 %."""
import numpy as np
import pytest
import torch

from tests.util import build_pair, rel_l2, synthetic_batch

pytestmark = pytest.mark.gpu

SHAPE = [64, 64, 3]
NCLS = 16
BATCH = 8


def _engine(pm, vals, keep=(), **kw):
    from myconvnet_b200.engine import Engine
    eng = Engine(pm, keep=keep, **kw)
    eng.set_variables(vals)
    return eng


@pytest.mark.parametrize("dtype,tol_act,tol_grad", [("f32", 2e-4, 8e-2), ("bf16", 6e-2, None)])
def test_resnet50_layers_and_grads(have_reference_models, dtype, tol_act, tol_grad):
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, dtype)
    X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
    taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k not in ("pred",)}
    eng = _engine(pm, vals, keep=list(taps.values()))
    loss_dev = eng.train_step(X, Y, update=False)
    tr = OracleTrainer(om)
    tr.step(X, Y, update=False)
    loss_ref = float(om.data_loss)     # without an optimiser step the device reports the data term
    bad = []
    for k, t in taps.items():
        ref = om.d[k].t.detach().numpy()
        got = eng.fetch(t)
        e = rel_l2(got, ref)
        late = dtype == "bf16" and (k.startswith("block_3") or k.startswith("block_4") or k.startswith("logits"))
        if e > (0.35 if late else tol_act):
            bad.append((k, e))
    assert not bad, "activation mismatches: %s" % bad[:8]
    assert abs(loss_dev - loss_ref) <= 1e-3 * abs(loss_ref) + (1e-4 if dtype == "f32" else 3e-2)
    grads = eng.get_gradients()
    assert all(np.isfinite(g).all() for g in grads.values())
    if tol_grad is None:
        return
    gbad = []
    for k, g in tr.grads.items():
        ref = g.numpy()
        if k.endswith("weights"):
            ref = ref - 0.0   # oracle grads include the L2 term; device folds it into the optimiser
            ref = ref - om._parameters.get("l2_reg", 1e-4) * vals[k]
        e = rel_l2(grads[k], ref)
        if e > tol_grad and np.linalg.norm(ref) > 1e-6:
            gbad.append((k, e))
    assert not gbad, "gradient mismatches: %s" % gbad[:8]


@pytest.mark.parametrize("dtype,tol", [("f32", 5e-3), ("bf16", 3e-2)])
def test_resnet50_fused_loss_curve(have_reference_models, dtype, tol):
    """Fused plan (BN+ReLU+residual, dense+cast, conv-epilogue statistics, gather stem) over
    several optimiser steps vs the oracle.

    Tolerance: `tol` on the first step (pure forward), widening linearly with the step index.
    Reason (measured, scripts/debug_fold.py): with 8 images the late BN layers normalise over a
    few dozen values, and a 1-ulp difference of one invstd flips enough ReLU mask bits to move
    BN-parameter gradients by ~1 % in ONE step (the step-3 loss by 1.2 % in fp32) — two correct
    implementations separate at that rate, so a fixed band over all steps would test rounding luck.
    A systematic error (wrong update order, missing momentum, stale statistics) shows up as tens
    of percent by step 2 and is still caught."""
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, dtype,
                              base_learning_rate=0.05)
    X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
    eng = _engine(pm, vals)
    tr = OracleTrainer(om, base_learning_rate=0.05)
    dev, ref = [], []
    for _ in range(4):
        dev.append(eng.train_step(X, Y))
        ref.append(tr.step(X, Y))
    assert np.all(np.isfinite(dev))
    for k, (a, b) in enumerate(zip(dev, ref)):
        band = tol * (1 + k)
        assert abs(a - b) <= band * abs(b) + band, (k, dev, ref)
    assert dev[-1] < 0.6 * dev[0] and ref[-1] < 0.6 * ref[0]          # both actually train
    # moving statistics and EMA shadows follow the reference update order
    v_dev = eng.get_variables()
    for k in ("block_0/conv_0/bn/mu", "block_0/conv_0/bn/sigma"):
        assert rel_l2(v_dev[k], om.vars[k].numpy()) < 5 * tol, k
    e_dev = eng.get_variables(ema=True)
    k = "block_None/logits/weights"
    assert rel_l2(e_dev[k], tr.ema[k].numpy()) < 5 * tol


def test_trainer_drives_the_engine_with_the_reference_schedule(have_reference_models):
    """Step driver: Trainer feeds shuffled in-memory batches and the warm-up /
    cosine multiplier of the reference loop (optimizers.py:608-632) to the device engine; the
    oracle trainer stepped with the same batches and multipliers gives the same losses."""
    from myconvnet_b200.trainer import Trainer
    from oracle import schedule
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, "f32",
                              base_learning_rate=0.05)
    rng = np.random.default_rng(11)
    n = 3 * BATCH + 3                                   # 4 steps per epoch, the 4th (partial) is skipped
    X = rng.uniform(size=[n] + SHAPE).astype(np.float32)
    Y = rng.integers(0, NCLS, size=n).astype(np.int32)
    eng = _engine(pm, vals)
    seen = []
    orig = eng.train_step

    def spy(Xb, Yb, lr_multiplier=1.0, fetch_loss=True):
        seen.append((Xb.copy(), Yb.copy(), lr_multiplier))
        return orig(Xb, Yb, lr_multiplier=lr_multiplier, fetch_loss=fetch_loss)
    eng.train_step = spy
    tr = Trainer(eng, n, num_epochs=2, seed=3, learning_warmup_epochs=0.5,
                 learning_rate_decay_method="cosine", learning_rate_decay_params=(0,))
    dev = tr.fit(X, Y, num_steps=5)
    want = schedule.multipliers(n, BATCH, 2, 0.5, "cosine", (0,))
    assert tr.steps_per_epoch == 4 and len(dev) == 4 and tr.curr_step == 5          # step 3 had no full batch
    assert np.allclose([m for _, _, m in seen], [want[0], want[1], want[2], want[4]], rtol=1e-12)
    ot = OracleTrainer(om, base_learning_rate=0.05)
    ref = [ot.step(xb, yb, lr_multiplier=m) for xb, yb, m in seen]
    for k, (a, b) in enumerate(zip(dev, ref)):
        band = 5e-3 * (1 + k)
        assert abs(a - b) <= band * abs(b) + band, (k, dev, ref)


def test_gradient_clipping_matches_clip_by_global_norm(have_reference_models):
    """gradient_threshold (optimizers.py:112-113, tf.clip_by_global_norm on the gradient of the full
    loss incl. the L2 term): the device step (mcn_grad_sqnorm + clip factor inside mcn_opt_step)
    moves the weights like the oracle's clipped step — and half as far as an unclipped one."""
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, "f32",
                              base_learning_rate=0.05)
    X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
    probe = OracleTrainer(om, base_learning_rate=0.05, gradient_threshold=1e12)
    probe.step(X, Y, update=False)
    thr = 0.5 * probe.grad_norm                          # the clip halves every gradient
    assert thr > 0
    key = "block_2/res_1/conv_1/weights"
    w0 = vals[key].astype(np.float64)
    deltas = {}
    for name, t in (("clipped", thr), ("free", None)):
        eng = _engine(pm, vals, gradient_threshold=t)
        eng.train_step(X, Y)
        deltas[name] = eng.get_variables()[key].astype(np.float64) - w0
    om.set_variables(vals)
    ot = OracleTrainer(om, base_learning_rate=0.05, gradient_threshold=thr)
    ot.step(X, Y)
    d_ref = om.vars[key].detach().numpy().astype(np.float64) - w0
    assert rel_l2(deltas["clipped"], d_ref) < 0.1                  # ReLU-kink floor of model-level gradients
    ratio = np.linalg.norm(deltas["clipped"]) / np.linalg.norm(deltas["free"])
    assert 0.49 < ratio < 0.51, ratio
    assert abs(np.linalg.norm(deltas["clipped"]) / np.linalg.norm(d_ref) - 1.0) < 0.02


def test_predict_uses_ema_shadows_and_moving_statistics(have_reference_models):
    """ConvNet.predict semantics (reference convnet.py:609-665, 1406, 1872-1876): after a few
    training steps the inference pass runs on the EMA shadows with BN in inference mode."""
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, "f32",
                              base_learning_rate=0.05)
    X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
    eng = _engine(pm, vals)
    tr = OracleTrainer(om, base_learning_rate=0.05)
    for _ in range(3):
        eng.train_step(X, Y)
        tr.step(X, Y)
    p_dev = eng.predict(X)
    p_ref = tr.predict(X, Y)
    assert p_dev.shape == p_ref.shape == (BATCH, NCLS)
    assert np.allclose(p_dev.sum(-1), 1.0, atol=1e-4)
    assert rel_l2(p_dev, p_ref) < 2e-2
    # and it differs from a training-mode forward (batch statistics, raw weights)
    eng.forward(X, Y)
    assert rel_l2(eng.fetch(pm.pred), p_dev) > 1e-3


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
