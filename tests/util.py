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
