"""GPU parity of the reference's models/resnet_v1_5.py, unchanged, on the B200 engine vs the CPU
oracle (SURVEY.md 8c pins 4) — built so that every comparison is well conditioned:

* SINGLE steps from identical state.  Before every step the engine is re-synchronised with the
  oracle (variables, optimiser slots, EMA shadows, step counter: tests/util.sync_engine_from_oracle)
  and loss, every gradient and every updated variable of that one step are compared.  Free-running
  trajectories are kept only as a loose sanity check (they measure ReLU-mask luck, see below).
* BASELINE shapes: config 1 exactly (ResNet-v1.5-50, 224x224x3, batch 32, 1000 classes), fp32 and
  bf16 — block_4 batch-norm then normalises over 1568 values per channel instead of the 32 of a
  64x64 / batch-8 toy.
* The oracle is evaluated on the DEVICE's ReLU pattern for gradient comparisons
  (oracle ConvNet.forced_relu_masks).  Measured at config 1 (profiles/r02_parity_config1.txt):
  fp32 activations agree to 3e-5 unforced, yet gradients only to 1.4e-2, and in bf16 (activations
  0.2 rel-L2 in block_4) the gradients decorrelate completely (rel-L2 ~ 1) — each flipped mask bit
  moves the gradient by a finite amount, ~sqrt(fraction of flipped units) in rel-L2, whatever the
  kernels do.  With the pattern forced the fp32 difference is arithmetic only (2e-3 on every
  gradient and update).  The pattern itself is pinned by the unforced fp32 activation check
  (every ReLU output is a tap).
* bf16 is checked LAYER BY LAYER with teacher forcing, forward and backward, on the production
  (fused) plan: every convolution / BN(+residual+ReLU) / pool / dense node is re-evaluated and
  differentiated by the oracle on the device's own inputs and upstream gradient
  (tests/util.layerwise_*_errors), so each comparison sees one layer's arithmetic — one bf16
  rounding, 5e-3 — instead of two separately rounded 50-layer pipelines, which drift apart
  chaotically (measured with the pattern forced: 0.2 rel-L2 in block_4 activations, 3e-2..0.2 on
  whole-step gradients; scripts/debug_sd.py).  Whole-step bf16 numbers are kept as loose sanity
  bounds (loss 1 %, updates 0.3).
* Everything is bit-reproducible (no floating-point atomics), so a tolerance that holds once holds
  on every B200."""
import numpy as np
import pytest
import torch

from tests.util import (build_pair, layerwise_backward_errors, layerwise_forward_errors, rel_l2,
                        synthetic_batch, sync_engine_from_oracle, worst)

pytestmark = pytest.mark.gpu

SHAPE = [128, 128, 3]        # block_4 BN over 4*4*16 = 256 values per channel
NCLS = 16
BATCH = 16


def _engine(pm, vals, keep=(), **kw):
    from myconvnet_b200.engine import Engine
    eng = Engine(pm, keep=keep, **kw)
    eng.set_variables(vals)
    return eng


def relu_pattern(eng, pm, variables=None):
    """The device's activation pattern: output > 0 of every ReLU of the graph, in call order."""
    return [eng.fetch(n.outputs[0]) > 0 for n in pm.graph.nodes if n.op == "act" and n.attrs["act"] == 1]


def data_grads(tr, vals, l2):
    """Oracle gradients of the data term (the device folds the L2 term into the optimiser)."""
    return {k: g.numpy() - (l2 * vals[k] if k.endswith("weights") else 0.0) for k, g in tr.grads.items()}


def grad_errors(dev, ref):
    """rel-L2 per tensor; tensors whose reference gradient is numerically zero are skipped."""
    gmax = max(float(np.linalg.norm(g)) for g in ref.values())
    return {k: rel_l2(dev[k], g) for k, g in ref.items() if np.linalg.norm(g) > 1e-7 * gmax}


# ------------------------------------------------------------------ BASELINE config 1, one step
@pytest.mark.parametrize("dtype,tol_act,tol_grad,tol_upd", [("f32", 1e-4, 2e-3, 2e-3), ("bf16", 0.3, 0.3, 0.3)])
def test_resnet50_config1_single_step(have_reference_models, dtype, tol_act, tol_grad, tol_upd):
    """BASELINE.json configs[0]: ResNet-v1.5-50, synthetic 224x224x3, batch 32, one training step.
    Fused plan (conv-epilogue statistics, BN+ReLU+residual, gather stem) vs the oracle on the
    device's ReLU pattern: loss, all 161 gradients, all updated variables, moving statistics, EMA;
    then the unfused plan with every d[...] tap materialised: activations layer by layer, and the
    112^2 -> 56^2 max-pool argmax bit-exact on the device's own input."""
    from oracle import tf_ops
    from oracle.step import OracleTrainer
    shape, ncls, batch = [224, 224, 3], 1000, 32
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", shape, ncls, batch, dtype)
    X, Y = synthetic_batch(batch, shape, ncls)
    l2 = om._parameters.get("l2_reg", 1e-4)
    # ---- fused plan, one optimiser step
    eng = _engine(pm, vals)
    loss_dev = eng.train_step(X, Y)
    om.forced_relu_masks = relu_pattern(eng, pm, vals)
    tr = OracleTrainer(om)
    loss_ref = tr.step(X, Y)
    assert abs(loss_dev - loss_ref) <= (1e-5 if dtype == "f32" else 1e-2) * abs(loss_ref), (loss_dev, loss_ref)
    gerr = grad_errors(eng.get_gradients(), data_grads(tr, vals, l2))
    assert len(gerr) >= 150 and worst(gerr, 1)[0][1] <= tol_grad, worst(gerr)
    new = eng.get_variables()
    uerr = {k: rel_l2(new[k] - vals[k], om.vars[k].detach().numpy() - vals[k]) for k in vals
            if np.linalg.norm(om.vars[k].detach().numpy() - vals[k]) > 1e-9}
    assert len(uerr) >= 250 and worst(uerr, 1)[0][1] <= tol_upd, worst(uerr)      # 161 trainable + moving statistics
    ema = eng.get_variables(ema=True)
    eerr = {k: rel_l2(ema[k], tr.ema[k].numpy()) for k in vals}
    assert worst(eerr, 1)[0][1] <= (1e-5 if dtype == "f32" else 5e-2), worst(eerr)
    # ---- a second run from the same state is bit-identical (no floating-point atomics anywhere)
    eng.set_variables(vals)
    assert eng.train_step(X, Y) == loss_dev
    again = eng.get_variables()
    assert all(np.array_equal(again[k], new[k]) for k in new)
    # ---- the production (fused) plan, layer by layer with teacher forcing: every conv / BN(+residual
    # +ReLU) / pool / dense node re-evaluated by the oracle on the device's own inputs — one layer's
    # arithmetic per comparison (bf16: a single rounding of the output, 2^-9 relative)
    eng.set_variables(vals)
    eng.train_step(X, Y, update=False)         # the forward must belong to the weights being read back
    lerr = layerwise_forward_errors(eng, pm)
    assert len(lerr) >= 108 and worst(lerr, 1)[0][1] <= (2e-5 if dtype == "f32" else 5e-3), worst(lerr)
    del eng
    # ---- and the backward pass of the same fused plan, layer by layer: all 161 parameter gradients
    # and every single-consumer input gradient against the oracle's vjp at the device's own tensors
    eng = _engine(pm, vals, keep_grads=True)
    eng.train_step(X, Y, update=False)
    berr = layerwise_backward_errors(eng, pm)
    n_par = sum(1 for k in berr if k.endswith(("/dw", "/db", "/dgamma", "/dbeta")))
    assert n_par == 161 and len(berr) >= 250, (n_par, len(berr))
    assert worst(berr, 1)[0][1] <= (1e-4 if dtype == "f32" else 8e-3), worst(berr)
    del eng
    # ---- unfused plan: every tap of the model's dict, layer by layer
    om.set_variables(vals)
    taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k != "pred"}
    eng = _engine(pm, vals, keep=list(taps.values()))
    eng.train_step(X, Y, update=False)
    if dtype == "f32":
        om.forced_relu_masks = None          # the fp32 pattern check is UNFORCED: it pins the masks
    else:
        om.forced_relu_masks = relu_pattern(eng, pm, vals)
    OracleTrainer(om).step(X, Y, update=False)
    aerr = {k: rel_l2(eng.fetch(t), om.d[k].t.detach().numpy()) for k, t in taps.items()}
    assert len(aerr) >= 170 and worst(aerr, 1)[0][1] <= tol_act, worst(aerr)
    node = [n for n in pm.graph.nodes if n.op == "max_pool"][0]
    xin = eng.fetch(node.inputs[0])[:4]
    am = eng.maxpool_argmax(node)[:4].astype(np.int64)
    assert np.array_equal(am, tf_ops.max_pool_argmax(torch.from_numpy(xin), [3, 3], [2, 2], "SAME").numpy())


# ------------------------------------------------------------------ re-synchronised steps
@pytest.mark.parametrize("dtype,opt,tol", [("f32", "nesterov", 2e-3), ("bf16", "nesterov", 0.3),
                                           ("f32", "rmsprop", 2e-3), ("f32", "adam", 5e-3)])
def test_resynchronised_steps(have_reference_models, dtype, opt, tol):
    """Three optimiser steps, each from the oracle's exact state (weights, optimiser slots, EMA,
    step counter): loss and every updated variable per step.  Covers the device rules of Nesterov
    momentum, RMSProp (mean-square slot starts at one) and Adam (bias correction by step) with
    non-trivial slots, the EMA warm-up d_t = min(decay, (1+t)/(10+t)) and the moving statistics."""
    from oracle.step import OracleTrainer
    kw = dict(base_learning_rate=0.05, base_weight_decay=1e-4) if opt == "nesterov" else \
        dict(base_learning_rate=0.002, momentum=0.9)
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, dtype, **kw)
    eng = _engine(pm, vals, optimizer=opt)
    tr = OracleTrainer(om, optimizer=opt, **kw)
    for step in range(3):
        X, Y = synthetic_batch(BATCH, SHAPE, NCLS, seed=10 + step)
        before = {k: v.detach().numpy().copy() for k, v in om.vars.items()}
        sync_engine_from_oracle(eng, tr)
        loss_dev = eng.train_step(X, Y, lr_multiplier=1.0 - 0.25 * step)
        om.forced_relu_masks = relu_pattern(eng, pm, before)
        loss_ref = tr.step(X, Y, lr_multiplier=1.0 - 0.25 * step)
        assert abs(loss_dev - loss_ref) <= (2e-5 if dtype == "f32" else 1e-2) * abs(loss_ref), (step, loss_dev, loss_ref)
        new = eng.get_variables()
        uerr = {k: rel_l2(new[k] - before[k], om.vars[k].detach().numpy() - before[k]) for k in before
                if np.linalg.norm(om.vars[k].detach().numpy() - before[k]) > 1e-9}
        assert worst(uerr, 1)[0][1] <= tol, (step, worst(uerr))
        ema = eng.get_variables(ema=True)
        assert max(rel_l2(ema[k], tr.ema[k].numpy()) for k in before) <= (1e-5 if dtype == "f32" else 5e-2)
        slots = eng.get_optimizer_state()
        from tests.util import oracle_slots
        serr = {k: rel_l2(slots[k], v) for k, v in oracle_slots(tr).items() if np.linalg.norm(v) > 1e-12}
        assert worst(serr, 1)[0][1] <= tol, (step, worst(serr))


def test_free_running_curve_is_sane(have_reference_models):
    """Sanity only: four free-running bf16 steps train on both sides and stay within 10 % of each
    other.  (Divergence of free-running trajectories measures mask luck, not correctness: the
    per-step comparisons above are the parity tests.)"""
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, "bf16",
                              base_learning_rate=0.05)
    X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
    eng = _engine(pm, vals)
    tr = OracleTrainer(om, base_learning_rate=0.05)
    dev = [eng.train_step(X, Y) for _ in range(4)]
    ref = [tr.step(X, Y) for _ in range(4)]
    assert np.all(np.isfinite(dev))
    assert dev[-1] < 0.8 * dev[0] and ref[-1] < 0.8 * ref[0], (dev, ref)
    assert all(abs(a - b) <= 0.1 * abs(b) for a, b in zip(dev, ref)), (dev, ref)


def test_trainer_drives_the_engine_with_the_reference_schedule(have_reference_models):
    """Step driver: Trainer feeds shuffled in-memory batches and the warm-up /
    cosine multiplier of the reference loop (optimizers.py:608-632) to the device engine.  Every
    step is compared with the oracle trainer from a re-synchronised state (same batch, same
    multiplier)."""
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
    ot = OracleTrainer(om, base_learning_rate=0.05)
    seen, pairs = [], []
    orig = eng.train_step

    def spy(Xb, Yb, lr_multiplier=1.0, fetch_loss=True):
        seen.append(lr_multiplier)
        sync_engine_from_oracle(eng, ot)
        pre = {k: v.detach().numpy().copy() for k, v in om.vars.items()}
        a = orig(Xb, Yb, lr_multiplier=lr_multiplier, fetch_loss=True)
        om.forced_relu_masks = relu_pattern(eng, pm, pre)
        pairs.append((a, ot.step(Xb, Yb, lr_multiplier=lr_multiplier)))
        return a
    eng.train_step = spy
    tr = Trainer(eng, n, num_epochs=2, seed=3, learning_warmup_epochs=0.5,
                 learning_rate_decay_method="cosine", learning_rate_decay_params=(0,))
    dev = tr.fit(X, Y, num_steps=5)
    want = schedule.multipliers(n, BATCH, 2, 0.5, "cosine", (0,))
    assert tr.steps_per_epoch == 4 and len(dev) == 4 and tr.curr_step == 5          # step 3 had no full batch
    assert np.allclose(seen, [want[0], want[1], want[2], want[4]], rtol=1e-12)
    for a, b in pairs:
        assert abs(a - b) <= 2e-5 * abs(b), pairs
    assert eng.global_step == ot.global_step == 4


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
        masks = relu_pattern(eng, pm, vals)
    om.set_variables(vals)
    om.forced_relu_masks = masks
    ot = OracleTrainer(om, base_learning_rate=0.05, gradient_threshold=thr)
    ot.step(X, Y)
    d_ref = om.vars[key].detach().numpy().astype(np.float64) - w0
    assert rel_l2(deltas["clipped"], d_ref) < 2e-3
    ratio = np.linalg.norm(deltas["clipped"]) / np.linalg.norm(deltas["free"])
    assert 0.499 < ratio < 0.501, ratio


def test_predict_uses_ema_shadows_and_moving_statistics(have_reference_models):
    """ConvNet.predict semantics (reference convnet.py:609-665, 1406, 1872-1876): after a few
    training steps the inference pass runs on the EMA shadows with BN in inference mode.  The
    oracle predicts from the DEVICE's shadows, so the comparison is of the inference pass alone."""
    from oracle.step import OracleTrainer
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, "f32",
                              base_learning_rate=0.05)
    X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
    eng = _engine(pm, vals)
    for _ in range(3):
        eng.train_step(X, Y)
    tr = OracleTrainer(om, base_learning_rate=0.05)
    tr.ema = {k: torch.from_numpy(v) for k, v in eng.get_variables(ema=True).items()}
    p_dev = eng.predict(X)
    p_ref = tr.predict(X, Y)
    assert p_dev.shape == p_ref.shape == (BATCH, NCLS)
    assert np.allclose(p_dev.sum(-1), 1.0, atol=1e-4)
    assert rel_l2(p_dev, p_ref) < 1e-4
    # and it differs from a training-mode forward (batch statistics, raw weights)
    eng.forward(X, Y)
    assert rel_l2(eng.fetch(pm.pred), p_dev) > 1e-3


def test_checkpoint_round_trip_resumes_the_run(have_reference_models, tmp_path):
    """Engine.save_checkpoint -> a fresh engine -> load_checkpoint continues bit-identically: same
    next-step loss and variables as the engine that never stopped (variables, EMA shadows,
    optimiser slots and global_step are all restored; SURVEY 8f row 4, optimizers.py:312), and
    predict() from the restored shadows agrees exactly."""
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", [64, 64, 3], NCLS, 8, "bf16",
                              base_learning_rate=0.05)
    X, Y = synthetic_batch(8, [64, 64, 3], NCLS)
    a = _engine(pm, vals)
    for _ in range(2):
        a.train_step(X, Y)
    path = str(tmp_path / "ck.npz")
    a.save_checkpoint(path)
    b = _engine(pm, {k: np.zeros_like(v) for k, v in vals.items()})
    missing = b.load_checkpoint(path)
    assert not missing and b.global_step == a.global_step == 2
    assert np.array_equal(a.predict(X), b.predict(X))
    la, lb = a.train_step(X, Y), b.train_step(X, Y)
    assert la == lb
    va, vb = a.get_variables(), b.get_variables()
    assert all(np.array_equal(va[k], vb[k]) for k in va)
    # a transfer-style load starts the optimiser from scratch
    c = _engine(pm, vals)
    c.load_checkpoint(path, resume=False)
    assert c.global_step == 0 and not any(np.any(v) for v in c.get_optimizer_state().values())
