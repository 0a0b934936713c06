"""Variable import / export with the reference's names.

The engine exposes variables under the reference's own names (`block_1/res_0/conv_0/weights`,
`.../bn/{mu,sigma,gamma,beta}`, reference convnet.py:1392,1805-1854) in the reference's layouts, and
their EMA shadows, which TensorFlow stores as `<name>/ExponentialMovingAverage`.  This module
writes / reads them as `.npz` (TensorFlow's checkpoint reader is not available offline) and
provides the name mapping to the public TF-slim ResNet-v1 checkpoints that the reference's
models/init_from_checkpoint.py:13-36 uses, so weights converted elsewhere can be dropped in."""
import re

import numpy as np

EMA_SUFFIX = "/ExponentialMovingAverage"
_BN = {"mu": "moving_mean", "sigma": "moving_variance", "gamma": "gamma", "beta": "beta"}


def slim_resnet_v1_name(name, depth=50):
    """`block_2/res_1/conv_0/bn/gamma` -> `resnet_v1_50/block2/unit_2/bottleneck_v1/conv1/BatchNorm/gamma`.
    Follows the rules of reference models/init_from_checkpoint.py:13-120 (blocks keep their number,
    units and convolutions count from 1, `conv_skip` is `shortcut`, block_0 is the stem `conv1`,
    block_None the `logits`).  Returns None for names the slim checkpoints do not contain."""
    ema = name.endswith(EMA_SUFFIX)
    parts = (name[:-len(EMA_SUFFIX)] if ema else name).split("/")
    leaf = parts[-1]
    out = ["resnet_v1_%d" % depth]
    tail = []
    if len(parts) >= 2 and parts[-2] == "bn":
        if leaf not in _BN:
            return None
        tail = ["BatchNorm", _BN[leaf]]
        parts = parts[:-2]
    elif leaf in ("weights", "biases"):
        tail = [leaf]
        parts = parts[:-1]
    else:
        return None
    m = re.fullmatch(r"block_(\w+)", parts[0])
    if m is None:
        return None
    block = m.group(1)
    if block == "0":
        c = re.fullmatch(r"conv_(\d+)", parts[1]) if len(parts) == 2 else None
        if c is None:
            return None
        out.append("conv%d" % (int(c.group(1)) + 1))
    elif block == "None":
        if parts[1:] != ["logits"]:
            return None
        out.append("logits")
    else:
        u = re.fullmatch(r"res_(\d+)", parts[1]) if len(parts) == 3 else None
        c = re.fullmatch(r"conv_(\w+)", parts[2]) if len(parts) == 3 else None
        if u is None or c is None:
            return None
        out += ["block%d" % int(block), "unit_%d" % (int(u.group(1)) + 1), "bottleneck_v1"]
        out.append("shortcut" if c.group(1) == "skip" else "conv%d" % (int(c.group(1)) + 1))
    return "/".join(out + tail) + (EMA_SUFFIX if ema else "")


def flatten(variables, ema=None):
    """{name: array} (+ shadows) -> one dict with TensorFlow's shadow naming."""
    out = {k: np.asarray(v) for k, v in variables.items()}
    for k, v in (ema or {}).items():
        out[k + EMA_SUFFIX] = np.asarray(v)
    return out


def split(flat):
    """Inverse of flatten: (variables, ema)."""
    var, ema = {}, {}
    for k, v in flat.items():
        if k.endswith(EMA_SUFFIX):
            ema[k[:-len(EMA_SUFFIX)]] = np.asarray(v)
        else:
            var[k] = np.asarray(v)
    return var, ema


EXTRA_PREFIX = "__state__/"      # optimiser slots, global_step: kept apart from the variables


def save_npz(path, variables, ema=None, naming="reference", depth=50, extra=None):
    """naming='reference' keeps the reference's names; 'slim' writes the TF-slim ResNet-v1 keys
    (variables without a slim counterpart are kept under their own name).  `extra`: optimiser
    slots (`<var>/Momentum`, ...), `global_step` — what tf.train.Saver would also have saved."""
    flat = flatten(variables, ema)
    if naming == "slim":
        flat = {(slim_resnet_v1_name(k, depth) or k): v for k, v in flat.items()}
    elif naming != "reference":
        raise ValueError("naming must be 'reference' or 'slim'")
    for k, v in (extra or {}).items():
        flat[EXTRA_PREFIX + k] = np.asarray(v)
    np.savez(path, **{k.replace("/", "|"): v for k, v in flat.items()})    # '/' is not a valid zip member name on every OS


def load_extra(path):
    """The `extra` dict of save_npz ({} for files without one)."""
    with np.load(path) as z:
        return {k.replace("|", "/")[len(EXTRA_PREFIX):]: z[k] for k in z.files
                if k.replace("|", "/").startswith(EXTRA_PREFIX)}


def load_npz(path, expected=None, naming="reference", depth=50, prefer_ema=False):
    """Returns (variables, ema).  With `expected` ({name: shape}, e.g. from the model) the keys are
    translated back from slim names when naming='slim', arrays whose shape does not match are
    skipped (as init_from_checkpoint.py:126-140 does) except [1,1,C,K] 1x1-conv logits, which are
    reshaped to the dense [C,K]; prefer_ema loads a shadow into the variable itself when present
    (load_moving_average=True there)."""
    with np.load(path) as z:
        flat = {k.replace("|", "/"): z[k] for k in z.files}
    flat = {k: v for k, v in flat.items() if not k.startswith(EXTRA_PREFIX)}
    if expected is None:
        return split(flat)
    var, ema = {}, {}
    for name, shape in expected.items():
        key = slim_resnet_v1_name(name, depth) if naming == "slim" else name
        if key is None:
            continue
        cands = [key + EMA_SUFFIX, key] if prefer_ema else [key]
        for ck in cands:
            if ck in flat:
                a = np.asarray(flat[ck])
                if a.ndim == 4 and len(shape) == 2 and a.shape[:2] == (1, 1) and a.shape[2:] == tuple(shape):
                    a = a.reshape(shape)
                if a.shape == tuple(shape):
                    var[name] = a
                    break
        if key + EMA_SUFFIX in flat and np.asarray(flat[key + EMA_SUFFIX]).shape == tuple(shape):
            ema[name] = np.asarray(flat[key + EMA_SUFFIX])
    return var, ema
