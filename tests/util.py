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
