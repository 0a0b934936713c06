"""Minimal stand-in for ``tensorflow.compat.v1`` — only what the reference model files touch.

This is NOT TensorFlow.  The five north-star model files (models/resnet_v1_5.py,
resnet_v1_5_dilated.py, efficientnet.py, deeplabv3plus.py, dcgan.py) call a dozen ``tf.*``
symbols; each is forwarded to a method of the tensor object it is given
(``x._tf_reduce_mean(...)``), so the same shim serves any engine whose tensors implement that
protocol.  Variable scopes are a plain name stack shared by whoever is building a model.
"""
import contextlib
import types

# ------------------------------------------------------------------ variable scopes
_scope_stack = []


class _ScopeState(object):
    """What tf.get_variable_scope() returns: the current name and the reuse flag."""

    def __init__(self):
        self.reuse = False

    @property
    def name(self):
        return "/".join(_scope_stack)

    def reuse_variables(self):
        self.reuse = True


_state = _ScopeState()


def reset_scopes():
    del _scope_stack[:]
    _state.reuse = False


def current_scope():
    return "/".join(_scope_stack)


@contextlib.contextmanager
def variable_scope(name_or_scope, *args, **kwargs):
    """tf.variable_scope(name): push a path component.  Passing the scope object returned by
    get_variable_scope() (reference convnet.py:430) re-enters the current scope unchanged."""
    if isinstance(name_or_scope, _ScopeState) or name_or_scope is None or name_or_scope == "":
        yield _state
        return
    parts = [p for p in str(name_or_scope).split("/") if p]
    _scope_stack.extend(parts)
    try:
        yield _state
    finally:
        for _ in parts:
            _scope_stack.pop()


def get_variable_scope():
    return _state


@contextlib.contextmanager
def name_scope(name, *args, **kwargs):
    yield


@contextlib.contextmanager
def device(name):
    yield


# ------------------------------------------------------------------ initializers
class Initializer(object):
    """Descriptor of a TF initializer; engines draw the numbers."""

    def __init__(self, kind, scale=1.0, mode="fan_in", distribution="truncated_normal", value=0.0):
        self.kind = kind
        self.scale = scale
        self.mode = mode
        self.distribution = distribution
        self.value = value

    def __repr__(self):
        return "Initializer(%s, scale=%g, mode=%s, dist=%s)" % (
            self.kind, self.scale, self.mode, self.distribution)


def _he_normal(seed=None):
    return Initializer("variance_scaling", scale=2.0, mode="fan_in", distribution="truncated_normal")


def _variance_scaling(scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None,
                      dtype=None):
    if distribution == "normal":
        distribution = "truncated_normal"
    return Initializer("variance_scaling", scale=scale, mode=mode, distribution=distribution)


def _zeros(dtype=None):
    return Initializer("constant", value=0.0)


def _ones(dtype=None):
    return Initializer("constant", value=1.0)


initializers = types.SimpleNamespace(he_normal=_he_normal, variance_scaling=_variance_scaling,
                                     zeros=_zeros, ones=_ones)
zeros_initializer = _zeros
ones_initializer = _ones


# ------------------------------------------------------------------ tensor ops (protocol dispatch)
def is_tensor(x):
    return hasattr(x, "_tf_is_tensor")


def reduce_mean(x, axis=None, keepdims=False, name=None):
    return x._tf_reduce_mean(axis, keepdims)


def concat(values, axis, name=None):
    return values[0]._tf_concat(list(values), axis)


def reshape(x, shape, name=None):
    return x._tf_reshape(list(shape))


def transpose(x, perm=None, name=None):
    return x._tf_transpose(perm)


def stop_gradient(x, name=None):
    return x._tf_stop_gradient()


def identity(x, name=None):
    return x


def _softmax(x, axis=-1, name=None):
    return x._tf_softmax()


def _sigmoid(x, name=None):
    return x._tf_activation("sigmoid", None)


def _dropout(x, rate=0.0, name=None, **kwargs):
    return x._tf_dropout(rate)


def _relu(x, name=None):
    return x._tf_activation("relu", None)


nn = types.SimpleNamespace(softmax=_softmax, sigmoid=_sigmoid, dropout=_dropout, relu=_relu)
math = types.SimpleNamespace(sigmoid=_sigmoid)
float32 = "f32"
float16 = "f16"
bfloat16 = "bf16"
int32 = "i32"
