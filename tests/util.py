"""Shared helpers for the parity tests: build the same model on the device engine and on the CPU
oracle with identical weights and inputs."""
import numpy as np

from myconvnet_b200 import convnet as product_convnet
from myconvnet_b200 import loader
from myconvnet_b200.engine import draw_initial_value


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def build_pair(rel_path, cls_name, input_shape, num_classes, batch, dtype, seed=0, gamma_fill=0.5,
               **kwargs):
    """Returns (product_model, oracle_model, values): same reference model file on both facades."""
    from oracle import ref_convnet
    pmod = loader.load_reference_model(rel_path, loader_product())
    omod = loader.load_reference_model(rel_path, {"convnet": ref_convnet})
    pm = getattr(pmod, cls_name)(input_shape, num_classes, batch_size=batch, compute_dtype=dtype, **kwargs)
    rng = np.random.default_rng(seed)
    vals = {v.name: draw_initial_value(v, rng) for v in pm.graph.vars.values()}
    for k in vals:
        # zero_scale_init gammas would silence the residual branch
        if k.endswith("gamma") and not vals[k].any():
            vals[k] = np.full_like(vals[k], gamma_fill)
    om = getattr(omod, cls_name)(input_shape, num_classes, oracle_round_bf16=(dtype == "bf16"), **kwargs)
    om.set_variables(vals)
    return pm, om, vals


def loader_product():
    return {"convnet": product_convnet}


def synthetic_batch(batch, input_shape, num_classes, seed=1):
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(batch,) + tuple(input_shape)).astype(np.float32)
    Y = rng.integers(0, num_classes, size=batch).astype(np.int32)
    return X, Y


# optimiser-slot naming: oracle state key -> TF slot name per optimiser (Engine.SLOT_NAMES)
_SLOTS = {"nesterov": {"a": "Momentum"}, "momentum": {"a": "Momentum"}, "sgd": {"a": "Momentum"},
          "rmsprop": {"mom": "RMSProp_1", "ms": "RMSProp"}, "adam": {"m": "Adam", "v": "Adam_1"}}


def oracle_slots(tr):
    """OracleTrainer.state -> {<var>/<TF slot name>: numpy} (what Engine.set_optimizer_state takes)."""
    names = _SLOTS[tr.kind]
    return {k + "/" + names[s]: t.detach().numpy() for k, st in tr.state.items() for s, t in st.items()}


def sync_engine_from_oracle(eng, tr):
    """Re-synchronise the device engine with the oracle trainer before a step: variables, optimiser
    slots, EMA shadows and the step counter.  Parity tests compare ONE step from identical state
    (loss, every gradient, every updated variable) instead of free-running trajectories, whose
    divergence measures ReLU-mask luck rather than correctness."""
    m = tr.model
    eng.set_variables({k: v.detach().numpy() for k, v in m.vars.items()}, reset_state=False)
    eng.set_optimizer_state(oracle_slots(tr), global_step=tr.global_step)
    if tr.ema is not None:
        eng.set_ema({k: v.detach().numpy() for k, v in tr.ema.items()})


def worst(errs, n=5):
    return sorted(errs.items(), key=lambda kv: -kv[1])[:n]


def layerwise_forward_errors(eng, pm):
    """Teacher-forced layer-by-layer check of the LAST forward pass of `eng`, on whatever plan it
    runs (the fused production plan included): every convolution, depthwise convolution, dense
    layer, batch-norm (with its fused residual / activation), max-pool and global-average-pool node
    is re-evaluated by the oracle op on the DEVICE's own input tensors and compared with the
    device's output.  Each comparison therefore sees one layer's arithmetic (one bf16 rounding in
    bf16 mode) instead of the accumulated, chaotically amplified difference of two separately
    rounded pipelines.  Returns {node scope / op: rel-L2}."""
    import torch
    from oracle import tf_ops as ops
    rd = (lambda t: t.bfloat16().float()) if pm.graph.compute_dtype == "bf16" else (lambda t: t)
    variables = {k: torch.from_numpy(v) for k, v in eng.get_variables().items()}
    errs = {}

    def fetch(t):
        return torch.from_numpy(np.asarray(eng.fetch(t), dtype=np.float32))
    for node in pm.graph.nodes:
        if node.attrs.get("fused_into") is not None:
            continue
        a = node.attrs
        if node.op in ("conv2d", "dwconv2d"):
            x = fetch(node.inputs[0])
            w = rd(variables[node.vars["w"].name])
            pad = "SAME" if (a["pad"][0] or a["pad"][1] or node.outputs[0].shape[1] * a["s"][0] >= x.shape[1]) else "VALID"
            f = ops.depthwise_conv2d if node.op == "dwconv2d" else ops.conv2d
            y = f(x, w, a["s"], pad, a["d"])
            if "b" in node.vars:
                y = y + variables[node.vars["b"].name]
            errs[node.scope + "/" + node.op] = rel_l2(fetch(node.outputs[0]).numpy(), y.numpy())
        elif node.op == "dense":
            x = fetch(node.inputs[0])
            y = ops.dense(x, rd(variables[node.vars["w"].name]),
                          variables[node.vars["b"].name] if "b" in node.vars else None)
            errs[node.scope + "/dense"] = rel_l2(fetch(a.get("final", node.outputs[0])).numpy(), y.numpy())
        elif node.op == "bn" and a.get("update", True):
            x = fetch(node.inputs[0])
            g = variables[node.vars["gamma"].name] if "gamma" in node.vars else None
            b = variables[node.vars["beta"].name] if "beta" in node.vars else None
            y, _, _ = ops.fused_batch_norm_train(x, g, b, a["eps"])
            if a.get("residual") is not None:
                y = y + fetch(a["residual"])
            code = a.get("act", 0)
            kinds = {0: None, 1: "relu", 2: "relu6", 3: "lrelu", 4: "tanh", 5: "sigmoid", 6: "swish"}
            if code:
                y = ops.activation(y, kinds[code], a.get("alpha") or None)
            errs[node.scope + "/bn"] = rel_l2(fetch(a.get("final", node.outputs[0])).numpy(), y.numpy())
        elif node.op == "max_pool":
            x = fetch(node.inputs[0])
            pad = "SAME" if node.outputs[0].shape[1] * a["s"][0] >= x.shape[1] else "VALID"
            errs[node.scope + "/max_pool"] = rel_l2(fetch(node.outputs[0]).numpy(), ops.max_pool(x, a["k"], a["s"], pad).numpy())
        elif node.op == "gap":
            x = fetch(node.inputs[0])
            errs["gap:%d" % node.id] = rel_l2(fetch(node.outputs[0]).numpy().reshape(x.shape[0], -1),
                                              x.mean(dim=(1, 2)).numpy())
    return errs


def layerwise_backward_errors(eng, pm):
    """Teacher-forced layer-by-layer check of the LAST backward pass of `eng`
    (Engine(keep_grads=True), after train_step(update=False)): for every convolution, dense layer,
    batch-norm (+ fused residual / activation), max-pool and global-average-pool node the oracle
    op is differentiated (torch autograd) at the DEVICE's own input tensors with the DEVICE's own
    upstream gradient; its parameter gradients are compared with the device's (all of them) and its
    input gradient with the device's whenever that input has no other consumer.  Returns
    {node scope / what: rel-L2}."""
    import torch
    from oracle import tf_ops as ops
    rd = (lambda t: t.bfloat16().float()) if pm.graph.compute_dtype == "bf16" else (lambda t: t)
    variables = {k: torch.from_numpy(v) for k, v in eng.get_variables().items()}
    dev_grads = eng.get_gradients()
    uses = {}
    for node in pm.graph.nodes:
        for v in node.vars.values():
            uses[v.name] = uses.get(v.name, 0) + 1
    kinds = {0: None, 1: "relu", 2: "relu6", 3: "lrelu", 4: "tanh", 5: "sigmoid", 6: "swish"}
    errs = {}

    def fetch(t, grad=False):
        a = eng.fetch_grad(t) if grad else eng.fetch(t)
        return None if a is None else torch.from_numpy(np.asarray(a, dtype=np.float32))

    def leaf(t):
        return t.clone().requires_grad_(True)

    def sole_consumer(t, node):
        cons = [c.attrs.get("fused_into") or c for c in t.consumers]
        return len(cons) == 1 and cons[0] is node

    def cmp_var(node, key, grad_t, tag, floor=0.0):
        """floor: natural scale of the sum (||upstream gradient||_F for the per-channel sums dbeta /
        dgamma / dbias).  When the true value cancels to rounding noise (a BN whose shift is removed
        by a following BN has dbeta == 0 up to 1e-10) the error is measured against 1 % of that
        scale instead of against the noise."""
        v = node.vars.get(key)
        if v is not None and v.trainable and uses[v.name] == 1 and v.name in dev_grads:
            ref = grad_t.numpy().astype(np.float64)
            den = max(float(np.linalg.norm(ref)), 0.01 * floor)
            if den > 1e-30:
                errs[node.scope + "/" + tag] = float(np.linalg.norm(dev_grads[v.name].astype(np.float64) - ref) / den)

    for node in pm.graph.nodes:
        if node.attrs.get("fused_into") is not None:
            continue
        a = node.attrs
        if node.op in ("conv2d", "dwconv2d", "dense"):
            out = a.get("final", node.outputs[0])
            gy = fetch(out, grad=True)
            if gy is None:
                continue
            x = leaf(fetch(node.inputs[0]))
            w = leaf(rd(variables[node.vars["w"].name]))
            b = leaf(variables[node.vars["b"].name]) if "b" in node.vars else None
            if node.op == "dense":
                y = ops.dense(x, w, b)
            else:
                pad = "SAME" if (a["pad"][0] or a["pad"][1] or node.outputs[0].shape[1] * a["s"][0] >= x.shape[1]) else "VALID"
                f = ops.depthwise_conv2d if node.op == "dwconv2d" else ops.conv2d
                y = f(x, w, a["s"], pad, a["d"])
                if b is not None:
                    y = y + b
            y.backward(gy)
            cmp_var(node, "w", w.grad, "dw")
            if b is not None:
                cmp_var(node, "b", b.grad, "db", floor=float(gy.norm()))
            gx = fetch(node.inputs[0], grad=True)
            if gx is not None and sole_consumer(node.inputs[0], node) and float(x.grad.norm()) > 1e-12:
                errs[node.scope + "/dx"] = rel_l2(gx.numpy(), x.grad.numpy())
        elif node.op == "bn" and a.get("update", True):
            final = a.get("final", node.outputs[0])
            gy = fetch(final, grad=True)
            if gy is None:
                continue
            x = leaf(fetch(node.inputs[0]))
            g = leaf(variables[node.vars["gamma"].name]) if "gamma" in node.vars else None
            be = leaf(variables[node.vars["beta"].name]) if "beta" in node.vars else None
            y, _, _ = ops.fused_batch_norm_train(x, g, be, a["eps"])
            r = None
            if a.get("residual") is not None:
                r = leaf(fetch(a["residual"]))
                y = y + r
            if a.get("act", 0) == 1:
                # ReLU on the DEVICE's pattern: a pre-activation within one ulp of zero may round to
                # the other side in the oracle's summation order, and one flipped element moves a
                # per-channel sum over 1e5 rows by ~1e-3 relative — mask luck, not arithmetic
                y = y * (fetch(final) > 0).to(y.dtype)
            elif a.get("act", 0):
                y = ops.activation(y, kinds[a["act"]], a.get("alpha") or None)
            y.backward(gy)
            if g is not None:
                cmp_var(node, "gamma", g.grad, "dgamma", floor=float(gy.norm()))
            if be is not None:
                cmp_var(node, "beta", be.grad, "dbeta", floor=float(gy.norm()))
            gx = fetch(node.inputs[0], grad=True)
            if gx is not None and sole_consumer(node.inputs[0], node):
                errs[node.scope + "/bn_dx"] = rel_l2(gx.numpy(), x.grad.numpy())
            if r is not None:
                gr = fetch(a["residual"], grad=True)
                if gr is not None and sole_consumer(a["residual"], node):
                    errs[node.scope + "/bn_dres"] = rel_l2(gr.numpy(), r.grad.numpy())
        elif node.op in ("max_pool", "gap"):
            gy = fetch(node.outputs[0], grad=True)
            gx = fetch(node.inputs[0], grad=True)
            if gy is None or gx is None or not sole_consumer(node.inputs[0], node):
                continue
            x = leaf(fetch(node.inputs[0]))
            if node.op == "gap":
                y = x.mean(dim=(1, 2)).reshape(gy.shape)
            else:
                pad = "SAME" if node.outputs[0].shape[1] * a["s"][0] >= x.shape[1] else "VALID"
                y = ops.max_pool(x, a["k"], a["s"], pad)
            y.backward(gy)
            errs["%s:%d/dx" % (node.op, node.id)] = rel_l2(gx.numpy(), x.grad.numpy())
    return errs
