"""Static layer graph recorded while a model file runs against the ConvNet facade.

The reference builds a TensorFlow graph once and then runs it (graph mode); this host does the
same: model code sees symbolic ``Tensor`` objects with static shapes, every layer call records a
``Node``, and ``plan.py`` later turns the node list into a fixed sequence of kernel launches.
Nothing here touches a GPU.
"""
import collections

import numpy as np


def same_pad(in_size, k, stride, dilation, padding):
    """TF padding rule: returns (out, pad_before, pad_after).
    SAME: out = ceil(in/stride), extra padding goes after (bottom/right)."""
    eff = (k - 1) * dilation + 1
    if padding.upper() == "SAME":
        out = -(-in_size // stride)
        total = max((out - 1) * stride + eff - in_size, 0)
        return out, total // 2, total - total // 2
    if padding.upper() == "VALID":
        out = -(-(in_size - eff + 1) // stride)
        return out, 0, 0
    raise ValueError("Padding of {} is not supported".format(padding))


class Shape(list):
    """What Tensor.get_shape() returns: a list of ints with TensorShape's as_list()."""

    def as_list(self):
        return list(self)

    def __getitem__(self, i):
        r = list.__getitem__(self, i)
        return Shape(r) if isinstance(i, slice) else r


class Var(object):
    """A model variable (reference weight_variable/bias_variable/BN variables,
    convnet.py:1382-1462, 1805-1870)."""

    def __init__(self, name, shape, init, trainable, kind, block, storage_shape=None):
        self.name = name
        self.shape = tuple(int(s) for s in shape)          # logical shape, reference layout
        self.storage_shape = tuple(storage_shape or self.shape)  # device layout (may be padded)
        self.init = init                # tfshim.Initializer
        self.trainable = trainable
        self.kind = kind                # 'weight' | 'bias' | 'norm' | 'stat'
        self.block = block
        self.needs_bf16 = False         # tensor-core operand copies wanted
        self.needs_bf16_t = False
        self.gemm_dims = None           # (taps, cin, cout) of the storage layout
        self.storage_rows = None        # storage row of each logical leading index (gather-stem layout)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def storage_size(self):
        return int(np.prod(self.storage_shape))

    def __repr__(self):
        return "Var(%s %s %s)" % (self.name, self.shape, self.kind)


class Tensor(object):
    _tf_is_tensor = True

    def __init__(self, graph, shape, dtype, node=None, name=None):
        self.graph = graph
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype              # 'f32' | 'bf16' | 'i32'
        self.node = node
        self.name = name
        self.consumers = []

    # ---- protocol used by model files
    def get_shape(self):
        return Shape(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    def __repr__(self):
        return "Tensor(%s %s %s)" % (self.name, self.shape, self.dtype)

    def __mul__(self, other):
        if isinstance(other, Tensor):
            return self.graph.mul(self, other)
        return self.graph.scale(self, float(other))

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self.graph.scale(self, 1.0 / float(other))

    def __add__(self, other):
        if isinstance(other, Tensor):
            return self.graph.add(self, other)
        return self.graph.offset(self, float(other))

    __radd__ = __add__

    # ---- tf shim protocol
    def _tf_reduce_mean(self, axis, keepdims):
        return self.graph.reduce_mean(self, axis, keepdims)

    def _tf_concat(self, values, axis):
        return self.graph.concat(values, axis)

    def _tf_reshape(self, shape):
        return self.graph.reshape(self, shape)

    def _tf_transpose(self, perm):
        raise NotImplementedError("tf.transpose is only reached with channel_first=True, which "
                                  "this backend does not support (NHWC only)")

    def _tf_stop_gradient(self):
        return self.graph.unary("stop_gradient", self)

    def _tf_softmax(self):
        return self.graph.unary("softmax", self)

    def _tf_activation(self, kind, alpha):
        return self.graph.activation(self, kind, alpha)

    def _tf_dropout(self, rate):
        rate = float(rate)
        if rate == 0.0:
            return self  # tf.nn.dropout(rate=0) is the identity
        if not 0.0 < rate < 1.0:
            raise ValueError("dropout rate must be in [0, 1)")
        return self.graph._add("dropout", [self], [self.shape], [self.dtype],
                               {"rate": rate, "layer": self.graph.next_random_layer()}).outputs[0]


class Node(object):
    def __init__(self, op, inputs, attrs, scope):
        self.op = op
        self.inputs = list(inputs)
        self.attrs = dict(attrs)
        self.scope = scope
        self.outputs = []
        self.vars = {}
        self.id = -1

    def __repr__(self):
        return "Node(%d %s %s)" % (self.id, self.op, self.scope)


ACT_CODES = {None: 0, "none": 0, "relu": 1, "relu6": 2, "lrelu": 3, "leaky_relu": 3, "tanh": 4,
             "sigmoid": 5, "swish": 6}


class Graph(object):
    def __init__(self, compute_dtype="bf16"):
        self.compute_dtype = compute_dtype
        self.nodes = []
        self.vars = collections.OrderedDict()
        self.inputs = collections.OrderedDict()
        self.losses = []                # scalar loss tensors (node op 'softmax_xent' ...)
        self.flops = 0
        self.collections = collections.defaultdict(list)

    def next_random_layer(self):
        """Id of a random op (dropout / stochastic depth): part of its Philox key."""
        self.random_layers = getattr(self, "random_layers", 0) + 1
        return self.random_layers

    # ---- construction helpers
    def _add(self, op, inputs, out_shapes, out_dtypes, attrs=None, scope=""):
        node = Node(op, inputs, attrs or {}, scope)
        node.id = len(self.nodes)
        self.nodes.append(node)
        for t in inputs:
            t.consumers.append(node)
        for i, (s, dt) in enumerate(zip(out_shapes, out_dtypes)):
            node.outputs.append(Tensor(self, s, dt, node, "%s/%s:%d" % (scope, op, i)))
        return node

    def placeholder(self, name, shape, dtype):
        node = self._add("input", [], [shape], [dtype], {"name": name}, name)
        self.inputs[name] = node.outputs[0]
        return node.outputs[0]

    def get_var(self, name, shape, init, trainable, kind, block, storage_shape=None):
        """tf.get_variable with reuse: same scope path -> same variable (convnet.py:475, gan.py:64)."""
        if name in self.vars:
            v = self.vars[name]
            if tuple(shape) != v.shape:
                raise ValueError("variable %s reused with shape %s != %s" % (name, shape, v.shape))
            return v, False
        v = Var(name, shape, init, trainable, kind, block, storage_shape)
        self.vars[name] = v
        return v, True

    # ---- element-wise / structural ops reachable from model files through the shim
    def unary(self, op, x, attrs=None):
        return self._add(op, [x], [x.shape], [x.dtype], attrs).outputs[0]

    def activation(self, x, kind, alpha=None):
        kind = (kind or "none").lower()
        if kind not in ACT_CODES:
            raise ValueError("Activation type of {} is not supported".format(kind))
        if ACT_CODES[kind] == 0:
            return x
        if kind in ("lrelu", "leaky_relu") and alpha is None:
            alpha = 0.2
        return self._add("act", [x], [x.shape], [x.dtype],
                         {"act": ACT_CODES[kind], "alpha": float(alpha or 0.0)}).outputs[0]

    def add(self, a, b):
        if a.shape != b.shape:
            raise ValueError("add: shape mismatch %s vs %s" % (a.shape, b.shape))
        return self._add("add", [a, b], [a.shape], [a.dtype]).outputs[0]

    def mul(self, a, b):
        # x * se_mask with mask [N,1,1,C] (efficientnet.py:163)
        if a.shape == b.shape and len(a.shape) == 4 and a.shape[1:3] == (1, 1):
            pass
        if len(b.shape) == 4 and b.shape[1:3] == (1, 1) and b.shape[0] == a.shape[0] \
                and b.shape[3] == a.shape[3]:
            return self._add("scale_bcast", [a, b], [a.shape], [a.dtype]).outputs[0]
        if len(a.shape) == 4 and a.shape[1:3] == (1, 1) and a.shape[0] == b.shape[0] \
                and a.shape[3] == b.shape[3]:
            return self._add("scale_bcast", [b, a], [b.shape], [b.dtype]).outputs[0]
        raise NotImplementedError("tensor*tensor is only supported as x[N,H,W,C]*mask[N,1,1,C]")

    def scale(self, x, s):
        return self._add("affine", [x], [x.shape], [x.dtype], {"scale": s, "offset": 0.0}).outputs[0]

    def offset(self, x, o):
        return self._add("affine", [x], [x.shape], [x.dtype], {"scale": 1.0, "offset": o}).outputs[0]

    def reduce_mean(self, x, axis, keepdims):
        axis = sorted(a % len(x.shape) for a in (axis if isinstance(axis, (list, tuple)) else [axis]))
        if len(x.shape) != 4 or axis != [1, 2]:
            raise NotImplementedError("reduce_mean is only supported over the H,W axes of NHWC")
        n, h, w, c = x.shape
        shape = (n, 1, 1, c) if keepdims else (n, c)
        return self._add("gap", [x], [shape], [x.dtype], {"keepdims": bool(keepdims)}).outputs[0]

    def concat(self, values, axis):
        r = len(values[0].shape)
        if axis % r != r - 1:
            raise NotImplementedError("concat is only supported along the channel axis")
        lead = values[0].shape[:-1]
        for v in values:
            if v.shape[:-1] != lead:
                raise ValueError("concat: leading shape mismatch")
        c = sum(v.shape[-1] for v in values)
        return self._add("concat", values, [lead + (c,)], [values[0].dtype]).outputs[0]

    def reshape(self, x, shape):
        shape = list(shape)
        known = int(np.prod([s for s in shape if s != -1]))
        if -1 in shape:
            shape[shape.index(-1)] = x.size // known
        if int(np.prod(shape)) != x.size:
            raise ValueError("reshape: %s -> %s" % (x.shape, shape))
        return self._add("reshape", [x], [tuple(shape)], [x.dtype]).outputs[0]

    def add_to_collection(self, name, value):
        self.collections[name].append(value)

    def get_collection(self, name):
        return list(self.collections.get(name, []))
