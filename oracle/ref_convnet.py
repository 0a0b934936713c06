"""ORACLE (test infrastructure): eager CPU ConvNet with the reference's layer-op API.

Runs the reference's model files UNCHANGED (through myconvnet_b200.loader with this module as the
``convnet`` base) on torch-CPU tensors, using oracle.tf_ops for every op and torch autograd for
gradients.  It restates reference convnet.py:425-607 (tower build, loss) and :1382-2556 (layer ops)
for ONE tower in training mode.  The product never imports this module; only tests/, smoke() and
bench.py's cpu_baseline leg do.  Parity vs TensorFlow itself is unpinned (see oracle/tf_ops.py).
"""
from contextlib import nullcontext

import numpy as np
import torch

from myconvnet_b200 import tfshim as tf
from . import tf_ops as ops


def _pair(v):
    if not isinstance(v, (list, tuple)):
        return [v, v]
    if len(v) == 1:
        return [v[0], v[0]]
    return list(v)


class OShape(list):
    def as_list(self):
        return list(self)

    def __getitem__(self, i):
        r = list.__getitem__(self, i)
        return OShape(r) if isinstance(i, slice) else r


class OTensor(object):
    """torch tensor wrapper implementing the protocol model files use."""
    _tf_is_tensor = True

    def __init__(self, t):
        self.t = t

    def get_shape(self):
        return OShape(self.t.shape)

    def __mul__(self, o):
        return OTensor(self.t * (o.t if isinstance(o, OTensor) else o))

    __rmul__ = __mul__

    def __add__(self, o):
        return OTensor(self.t + (o.t if isinstance(o, OTensor) else o))

    __radd__ = __add__

    def __truediv__(self, o):
        return OTensor(self.t / o)

    def _tf_reduce_mean(self, axis, keepdims):
        return OTensor(self.t.mean(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis,
                                   keepdim=keepdims))

    def _tf_concat(self, values, axis):
        return OTensor(torch.cat([v.t for v in values], dim=axis))

    def _tf_reshape(self, shape):
        return OTensor(self.t.reshape(shape))

    def _tf_transpose(self, perm):
        return OTensor(self.t.permute(*perm))

    def _tf_stop_gradient(self):
        return OTensor(self.t.detach())

    def _tf_softmax(self):
        return OTensor(torch.softmax(self.t, dim=-1))

    def _tf_activation(self, kind, alpha):
        return OTensor(ops.activation(self.t, kind, alpha))

    def _tf_dropout(self, rate):
        # tf.nn.dropout(x, rate): keep u >= rate, scale 1/(1-rate); the keep mask is the device's
        # counter-based one (oracle/philox.py), off at inference
        rate = float(rate)
        m = _CURRENT[0]
        if rate == 0.0:
            return self
        layer = m._next_random_layer()
        if not getattr(m, "is_train", True):
            return self
        from . import philox
        keep = philox.dropout_keep(self.t.numel(), rate, m.random_seed, m.random_step, layer)
        mask = torch.from_numpy(keep.reshape(tuple(self.t.shape))).to(self.t.dtype) / (1.0 - rate)
        return OTensor(m._q(self.t * mask))


_CURRENT = [None]      # the model whose forward is running (tensor-level ops need its RNG context)


class ConvNet(object):
    """Eager oracle: constructing the model does NOT run it; call forward(X, Y)."""

    def __init__(self, input_shape, num_classes, loss_weights=None, session=None, model_scope=None,
                 companion_networks=None, next_elements=None, backbone_only=False, auto_build=True,
                 **kwargs):
        self._block_list = []
        self._curr_block = None
        self._input_size = list(input_shape)
        self._num_classes = num_classes
        self._loss_weights = loss_weights
        self._model_scope = model_scope
        self._backbone_only = backbone_only
        self._parameters = kwargs
        self.channel_first = False
        self.dtype = torch.float64 if kwargs.get("oracle_fp64", False) else torch.float32
        # emulate the device's bf16 storage of activations/weights (round-to-nearest-even)
        self.round_bf16 = bool(kwargs.get("oracle_round_bf16", False))
        self._blocks_to_train = kwargs.get("blocks_to_train", None)
        self._update_batch_norm = kwargs.get("update_batch_norm", None)
        self._moving_average_decay = kwargs.get("moving_average_momentum", kwargs.get("moving_average_decay", 0.99))
        self._batch_norm_decay = kwargs.get("batch_norm_momentum", kwargs.get("batch_norm_decay", 0.99))
        self._feature_reduction = kwargs.get("feature_reduction_factor", 0)
        self.dropout_rate_features = float(kwargs.get("dropout_rate", 0.0)) if kwargs.get("dropout_features", True) else 0.0
        self.image_mean = float(kwargs.get("image_mean", 0.5)) if kwargs.get("zero_center", True) else 0.0
        self.scale_factor = float(kwargs.get("scale_factor", 2.0))
        self.vars = {}          # name -> torch tensor (requires_grad for trainable)
        self.var_meta = {}      # name -> dict(kind, trainable, block, init, shape)
        self.collections = {}
        self.bn_updates = {}    # name -> new moving value computed this forward
        # random train-time ops: same (seed, step, layer) -> same masks as the device (philox.py)
        self.random_seed = int(kwargs.get("random_seed", 0))
        self.random_step = 0
        self._random_layers = 0
        # teacher-forced ReLU pattern: list of boolean arrays in call order (see _relu)
        self.forced_relu_masks = None
        self._relu_calls = 0
        self._flops = 0
        self._params = 0
        self._init_params(**kwargs)

    # ---- bookkeeping identical in spirit to the reference
    def __setattr__(self, key, value):
        if key == "_curr_block":
            self.__dict__[key] = value
            if value not in self._block_list:
                self._block_list.append(value)
        else:
            super(ConvNet, self).__setattr__(key, value)

    def _init_params(self, **kwargs):
        pass

    @property
    def input_size(self):
        return self._input_size

    @property
    def num_classes(self):
        return self._num_classes

    @property
    def loss_weights(self):
        return self._loss_weights

    @property
    def backbone_only(self):
        return self._backbone_only

    @property
    def blocks_to_train(self):
        return self._blocks_to_train

    @property
    def update_batch_norm(self):
        return self._update_batch_norm

    @property
    def batch_norm_decay(self):
        return self._batch_norm_decay

    @property
    def moving_average_decay(self):
        return self._moving_average_decay

    @property
    def feature_reduction(self):
        return self._feature_reduction

    @property
    def block_list(self):
        return tuple(self._block_list)

    @property
    def num_blocks(self):
        return len([b for b in self._block_list if b is not None
                    and self.collections.get("block_{}/variables".format(b))])

    def add_to_collection(self, name, v):
        self.collections.setdefault(name, []).append(v)

    def get_collection(self, name):
        return list(self.collections.get(name, []))

    # ---- variables
    def _trainable(self):
        return self.blocks_to_train is None or self._curr_block in self.blocks_to_train

    def _get_var(self, name, shape, init, kind, trainable=None):
        full = tf.current_scope() + "/" + name if tf.current_scope() else name
        if trainable is None:
            trainable = self._trainable()
        if full not in self.var_meta:
            self.var_meta[full] = dict(kind=kind, trainable=trainable, block=self._curr_block,
                                       init=init, shape=tuple(int(s) for s in shape))
        coll = "block_{}/variables".format(self.var_meta[full]["block"])
        if full not in self.collections.get(coll, ()):
            self.add_to_collection(coll, full)
        if full not in self.vars:
            raise KeyError("oracle variable %s has no value: call set_variables first" % full)
        return self.vars[full]

    def set_variables(self, values):
        """{name: numpy array in reference layout}."""
        for k, v in values.items():
            t = torch.tensor(np.asarray(v), dtype=self.dtype)
            self.vars[k] = t

    def _next_random_layer(self):
        self._random_layers += 1
        return self._random_layers

    def _relu(self, t):
        """max(x, 0) — or, with forced_relu_masks, x*mask with the DEVICE's activation pattern.
        A ReLU network's gradient is discontinuous in its pre-activations: two correct
        implementations whose activations differ by rounding flip a few masks, and every flip moves
        gradients by a finite amount (measured at BASELINE config 1, fp32: activations agree to
        3e-5 and gradients only to 1e-2; in bf16 the patterns decorrelate the gradients
        completely).  Evaluating the oracle ON the device's pattern removes that noise floor, so
        the comparison measures the arithmetic and tolerances can be tight.  The pattern itself is
        checked separately (unforced fp32 activations)."""
        if self.forced_relu_masks is None:
            return torch.relu(t)
        m = self.forced_relu_masks[self._relu_calls]
        self._relu_calls += 1
        return t * torch.from_numpy(np.ascontiguousarray(m)).to(t.dtype).reshape(t.shape)

    def _q(self, t):
        if self.round_bf16:
            # straight-through bf16 rounding: forward sees the rounded value, gradient passes
            return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()
        return t

    # ---- forward for one batch
    def forward(self, X, Y):
        """X: float [N,H,W,C] in [0,1]; Y: int labels.  Returns the loss; fills self.d."""
        tf.reset_scopes()
        _CURRENT[0] = self
        self._random_layers = 0
        self._relu_calls = 0
        self._block_list = []
        self.collections = {}       # every forward is a fresh build (block registry included)
        self.bn_updates = {}
        for name, t in self.vars.items():
            t.requires_grad_(self.var_meta.get(name, {}).get("trainable", True)
                             and self.var_meta.get(name, {}).get("kind", "weight") != "stat")
            t.grad = None
        x = torch.as_tensor(np.asarray(X), dtype=self.dtype)
        self.X = OTensor(self._q((x - self.image_mean) * self.scale_factor))
        self.Y = torch.as_tensor(np.asarray(Y)).long()
        self._curr_block = None
        self.d = self._build_model()
        self.logits = self.d["logits"].t.to(self.dtype)
        self.pred = self.d["pred"].t
        return self._build_loss(**self._parameters)

    def _build_loss(self, **kwargs):
        l1_factor = kwargs.get("l1_reg", 0.0)
        l2_factor = kwargs.get("l2_reg", 1e-4)
        ls = kwargs.get("label_smoothing", 0.0)
        data = ops.classification_loss(self.logits, self.Y, self.num_classes, self.loss_weights, ls,
                                       focal_gamma=kwargs.get("focal_loss_factor", 0.0),
                                       sigmoid_focal_alpha=kwargs.get("sigmoid_focal_loss_factor", 0.0),
                                       spatial_smoothing=self._spatial_label_smoothing)
        names = [n for n, m in self.var_meta.items() if m["kind"] == "weight"]
        if kwargs.get("bias_norm_decay", False):
            names += [n for n, m in self.var_meta.items() if m["kind"] in ("bias", "norm")]
        reg = sum(ops.l2_loss(self.vars[n]) for n in names) * l2_factor if l2_factor > 0 else 0.0
        if l1_factor > 0:
            reg = reg + l1_factor * sum(self.vars[n].abs().sum() for n in names)
        self.data_loss = data
        return data + reg

    _spatial_label_smoothing = False

    # ---- layer ops (signatures of reference convnet.py)
    def max_pool(self, x, side_l, stride, padding="SAME"):
        return OTensor(self._q(ops.max_pool(x.t, _pair(side_l), _pair(stride), padding)))

    def avg_pool(self, x, side_l, stride, padding="SAME"):
        return OTensor(self._q(ops.avg_pool(x.t, _pair(side_l), _pair(stride), padding)))

    def pooling_layer(self, x, kernel, stride, padding="SAME", pooling_type="AVG"):
        if pooling_type.lower() == "avg":
            return self.avg_pool(x, kernel, stride, padding)
        if pooling_type.lower() == "max":
            return self.max_pool(x, kernel, stride, padding)
        raise ValueError("Pooling type of {} is not supported".format(pooling_type))

    def conv_layer(self, x, kernel, stride, out_channels=None, padding="SAME", biased=True,
                   depthwise=False, scope=None, dilation=(1, 1), ws=False,
                   kernel_paddings=((0, 0), (0, 0)), weight_initializer=tf.initializers.he_normal(),
                   bias_initializer=tf.initializers.zeros(), verbose=False):
        kernel, stride, dilation = _pair(kernel), _pair(stride), _pair(dilation)
        cin = x.t.shape[-1]
        if out_channels is None:
            out_channels = cin
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            if depthwise:
                mult = max(out_channels // cin, 1)
                w = self._q(self._ws(self._get_var("weights", [kernel[0], kernel[1], cin, mult], weight_initializer, "weight"), ws))
                y = ops.depthwise_conv2d(x.t, w, stride, padding, dilation)
                out_c = cin * mult
            else:
                w = self._q(self._ws(self._get_var("weights", [kernel[0], kernel[1], cin, out_channels], weight_initializer, "weight"), ws))
                y = ops.conv2d(x.t, w, stride, padding, dilation)
                out_c = out_channels
            if biased:
                y = y + self._get_var("biases", [out_c], bias_initializer, "bias")
        return OTensor(self._q(y))

    def conv_bn_act(self, x, kernel, stride, out_channels=None, padding="SAME", biased=False,
                    depthwise=False, scope=None, dilation=(1, 1), ws=False,
                    kernel_paddings=((0, 0), (0, 0)), weight_initializer=tf.initializers.he_normal(),
                    bias_initializer=tf.initializers.zeros(), scale=True, shift=True,
                    zero_scale_init=False, epsilon=1e-3, act_type="relu", act_params=None, verbose=False):
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            x = self.conv_layer(x, kernel, stride, out_channels, padding=padding, biased=biased,
                                depthwise=depthwise, dilation=dilation, ws=ws,
                                weight_initializer=weight_initializer, bias_initializer=bias_initializer)
            x = self.batch_norm(x, scale=scale, shift=shift, zero_scale_init=zero_scale_init, epsilon=epsilon)
            x = self.activation(x, activation_type=act_type, params=act_params)
        return x

    def transposed_conv_layer(self, x, kernel, stride, out_channels, padding="SAME", biased=True,
                              output_shape=None, dilation=(1, 1), scope=None,
                              weight_initializer=tf.initializers.he_normal(),
                              bias_initializer=tf.initializers.zeros(), ws=False, verbose=False):
        kernel, stride, dilation = _pair(kernel), _pair(stride), _pair(dilation)
        n, h, w_, cin = x.t.shape
        if output_shape is None:
            if padding.lower() == "valid":
                out_hw = [h * stride[0] - kernel[0] + 1, w_ * stride[1] - kernel[1] + 1]
            else:
                out_hw = [h * stride[0], w_ * stride[1]]
        else:
            out_hw = list(output_shape[1:3])
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            w = self._q(self._ws(self._get_var("weights", [kernel[0], kernel[1], cin, out_channels], weight_initializer, "weight"), ws))
            y = ops.conv2d_transpose(x.t, w, out_hw, stride, padding, dilation)
            if biased:
                y = y + self._get_var("biases", [out_channels], bias_initializer, "bias")
        return OTensor(self._q(y))

    def fc_layer(self, x, out_dim, biased=True, scope=None, ws=False,
                 weight_initializer=tf.initializers.he_normal(),
                 bias_initializer=tf.initializers.zeros(), verbose=False):
        in_dim = int(x.t.shape[-1])
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            w = self._q(self._ws(self._get_var("weights", [in_dim, out_dim], weight_initializer, "weight"), ws))
            b = self._get_var("biases", [out_dim], bias_initializer, "bias") if biased else None
            y = ops.dense(x.t, w, b)
        return OTensor(y)   # logits stay fp32 (device epilogue writes fp32)

    @staticmethod
    def _ws(w, enabled):
        return ops.weight_standardization(w) if enabled else w

    def normalization(self, x, norm_type="batch", norm_param=None, scale=True, shift=True,
                      zero_scale_init=False, epsilon=1e-3, scope="norm"):
        if norm_type is None:
            return x
        if norm_type.lower() == "batch":
            return self.batch_norm(x, scale=scale, shift=shift, zero_scale_init=zero_scale_init,
                                   epsilon=epsilon, scope=scope)
        if norm_type.lower() == "group":
            return self.group_norm(x, num_groups=32 if norm_param is None else norm_param, scale=scale,
                                   shift=shift, zero_scale_init=zero_scale_init, epsilon=epsilon, scope=scope)
        raise NotImplementedError("oracle: norm_type %s" % norm_type)

    def group_norm(self, x, num_groups=32, scale=True, shift=True, zero_scale_init=False, epsilon=1e-3,
                   scope="gn"):
        """convnet.py:1928-2013."""
        trainable = self._trainable()
        c = x.t.shape[-1]
        assert c // num_groups * num_groups == c, \
            "Number of channels must be a multiple of num_groups ({})".format(num_groups)
        with tf.variable_scope(scope):
            gamma = self._get_var("gamma", [c], tf.zeros_initializer() if zero_scale_init else tf.ones_initializer(),
                                  "norm", trainable=trainable) if scale else None
            beta = self._get_var("beta", [c], tf.zeros_initializer(), "norm", trainable=trainable) if shift else None
            y = ops.group_norm(x.t, gamma, beta, num_groups, epsilon)
        return OTensor(y)   # rounding happens after the fused activation, as on device

    def batch_norm(self, x, scale=True, shift=True, zero_scale_init=False, epsilon=1e-3, scope="bn"):
        """convnet.py:1780-1926, training branch; moving statistics recorded in self.bn_updates."""
        if isinstance(self.update_batch_norm, bool):
            update = self.update_batch_norm
        else:
            update = self._trainable()
        trainable = self._trainable()
        c = x.t.shape[-1]
        with tf.variable_scope(scope):
            sc = tf.current_scope()
            mu = self._get_var("mu", [c], tf.zeros_initializer(), "stat", trainable=False)
            sigma = self._get_var("sigma", [c], tf.ones_initializer(), "stat", trainable=False)
            gamma = self._get_var("gamma", [c], tf.zeros_initializer() if zero_scale_init else tf.ones_initializer(),
                                  "norm", trainable=trainable) if scale else None
            beta = self._get_var("beta", [c], tf.zeros_initializer(), "norm", trainable=trainable) if shift else None
            if not getattr(self, "is_train", True) or not update:
                # is_train=False: moving statistics (their EMA shadows, loaded by the trainer).
                # update=False (blocks outside blocks_to_train / update_batch_norm=False): the
                # reference calls fused_batch_norm(is_training=False, mean=mu, variance=sigma) even
                # while training (convnet.py:1916-1924) — frozen layers normalise with the stored
                # moving statistics and have no batch-statistics terms in their backward pass
                return OTensor(ops.fused_batch_norm_infer(x.t, gamma, beta, mu.detach(), sigma.detach(), epsilon))
            y, bm, bv = ops.fused_batch_norm_train(x.t, gamma, beta, epsilon)
            if update:
                m = self.batch_norm_decay
                mu_prev = self.bn_updates.get(sc + "/mu", mu)
                sigma_prev = self.bn_updates.get(sc + "/sigma", sigma)
                self.bn_updates[sc + "/mu"] = (m * mu_prev + (1 - m) * bm).detach()
                self.bn_updates[sc + "/sigma"] = (m * sigma_prev + (1 - m) * bv).detach()
        return OTensor(y)   # rounding happens after the fused activation/residual, as on device

    def upsampling_2d_layer(self, x, scale=2, out_shape=None, align_corners=False,
                            force_unaligned=False, upsampling_method="bilinear", name="upsampling"):
        if out_shape is None:
            out_shape = [x.t.shape[1] * scale, x.t.shape[2] * scale]
        if force_unaligned:
            ac, hp = False, False
        else:
            ac, hp = align_corners, not align_corners
        if upsampling_method.lower() in ("nearest", "nearest_neighbor"):
            return OTensor(ops.resize_nearest(x.t, [int(s) for s in out_shape], ac, hp))
        if upsampling_method.lower() != "bilinear":
            raise ValueError("Upsampling method of {} is not supported".format(upsampling_method))
        return OTensor(self._q(ops.resize_bilinear(x.t, [int(s) for s in out_shape], ac, hp)))

    def stochastic_depth(self, x, skip, drop_rate=0.0, name="drop"):
        """convnet.py:2500-2512: x*survived + skip, survived[n] = (u_n >= rate)/(1-rate) per sample
        while training; the uniform numbers are the device's (oracle/philox.py)."""
        if drop_rate > 0.0:
            layer = self._next_random_layer()
            if not getattr(self, "is_train", True):
                return x + skip
            from . import philox
            n = x.t.shape[0]
            keep = philox.survive(n, drop_rate, self.random_seed, self.random_step, layer)
            s = torch.from_numpy(keep).to(x.t.dtype).reshape([n] + [1] * (x.t.dim() - 1)) / (1.0 - drop_rate)
            return OTensor(x.t * s + skip.t)
        return x + skip

    def activation(self, x, activation_type="relu", params=None):
        if activation_type is None:
            return x
        if activation_type.lower() == "relu":
            return OTensor(self._q(self._relu(x.t)))
        return OTensor(self._q(ops.activation(x.t, activation_type, params)))

    def relu(self, x, name="relu"):
        return OTensor(self._q(self._relu(x.t)))

    def relu6(self, x, name="relu6"):
        return OTensor(self._q(ops.activation(x.t, "relu6")))

    def lrelu(self, x, alpha=None, name="lrelu"):
        return OTensor(self._q(ops.activation(x.t, "lrelu", alpha)))

    def tanh(self, x, name="tanh"):
        return OTensor(self._q(torch.tanh(x.t)))

    def sigmoid(self, x, name=None):
        return OTensor(self._q(torch.sigmoid(x.t)))

    def swish(self, x, name="swish"):
        return OTensor(self._q(ops.activation(x.t, "swish")))
