"""ConvNet facade: the model-definition API of the reference, recording a static layer graph.

Mirrors the method surface model files use (reference convnet.py:1382-2577: weight_variable ...
swish; constructor and attribute protocol of convnet.py:14-424) so that the reference's own
models/resnet_v1_5.py, efficientnet.py, deeplabv3plus.py and dcgan.py import and build unchanged
(see loader.py).  Layer calls do no arithmetic: they append nodes to ``self.graph``; plan.py
lowers the graph to launches of the sm_100a kernels in libmcn.so.  There is no CPU execution path.

Differences by design: one process per GPU, so a model instance builds
ONE tower; variables are replicated, not on a parameter device; BN statistics are synchronised
across ranks instead of chained tower after tower.
"""
from contextlib import nullcontext

import numpy as np

from . import tfshim as tf
from .graph import Graph, same_pad


def _pair(v):
    if not isinstance(v, (list, tuple)):
        return [v, v]
    if len(v) == 1:
        return [v[0], v[0]]
    return list(v)


class ConvNet(object):
    def __init__(self, input_shape, num_classes, loss_weights=None, session=None, model_scope=None,
                 companion_networks=None, next_elements=None, backbone_only=False, auto_build=True,
                 **kwargs):
        self._block_list = []
        self._curr_block = None
        assert len(input_shape) == 3, 'input_size must contain 3D size'
        self._input_size = list(input_shape)
        self._num_classes = num_classes
        self._loss_weights = loss_weights
        self._model_scope = model_scope
        self._backbone_only = backbone_only
        self._parameters = kwargs

        # compute dtype: the reference's half_precision flag selects fp16 (convnet.py:63); on
        # B200 the reduced-precision type is bf16.  `compute_dtype` overrides explicitly.
        default_dt = 'bf16' if kwargs.get('half_precision', False) else 'f32'
        self._dtype = kwargs.get('compute_dtype', default_dt)
        if self._dtype not in ('f32', 'bf16'):
            raise ValueError('compute_dtype must be f32 or bf16')
        if kwargs.get('channel_first', False):
            raise NotImplementedError('channel_first=True (NCHW) is not supported: the B200 kernels '
                                      'are NHWC only (reference default, convnet.py:64)')
        if kwargs.get('dropout_weights', False):
            raise NotImplementedError('dropout_weights is not supported')
        self._channel_first = False
        self._argmax_output = kwargs.get('argmax_output', False)
        self._batch_size = int(kwargs.get('batch_size', 32))   # per-device (tower) batch

        self._num_devices = 1
        self._compute_device = 'gpu'
        self._device_offset = 0
        self._param_device = '/gpu:0'
        self._curr_device = 0

        self._padded_size = np.round(np.array(self.input_size[0:2]) * (1.0 + kwargs.get('zero_pad_ratio', 0.0)))
        self.pad_value = kwargs.get('pad_value', 0.5)
        self._dropout_weights = False
        self._dropout_features = kwargs.get('dropout_features', True)
        self._blocks_to_train = kwargs.get('blocks_to_train', None)
        self._update_batch_norm = kwargs.get('update_batch_norm', None)
        self._moving_average_decay = kwargs.get('moving_average_momentum',
                                                kwargs.get('moving_average_decay', 0.99))
        self._batch_norm_decay = kwargs.get('batch_norm_momentum', kwargs.get('batch_norm_decay', 0.99))
        self._feature_reduction = kwargs.get('feature_reduction_factor', 0)

        self._flops = 0
        self._params = 0
        self._nodes = 0
        self._layer_info = []
        self.dicts = []
        self.losses = []
        self.valid_masks = []
        self.graph = Graph(self._dtype)
        self._reuse = False

        if auto_build:
            self.build()

    # ------------------------------------------------------------------ build
    def build(self):
        kwargs = self._parameters
        tf.reset_scopes()
        # Train-time constants the reference holds in tf.cond/placeholders (convnet.py:146-175).
        self.is_train = True
        self.dropout_rate = float(kwargs.get('dropout_rate', 0.0))
        self.dropout_rate_weights = 0.0
        self.dropout_rate_features = self.dropout_rate if self._dropout_features else 0.0
        self.image_mean = float(kwargs.get('image_mean', 0.5)) if kwargs.get('zero_center', True) else 0.0
        self.scale_factor = float(kwargs.get('scale_factor', 2.0))
        with tf.variable_scope(self.model_scope) if self.model_scope is not None else nullcontext():
            self._init_params(**kwargs)
            self._init_model(**kwargs)
        self._flops = int(self._flops)
        self._params = int(self._params)
        self._nodes = int(self._nodes)
        for blk in self.block_list:
            if not self.get_collection('block_{}/variables'.format(blk)):
                self._block_list.remove(blk)
        if kwargs.get('verbose', False):
            print('\n# FLOPs : {:-15,}\n# Params: {:-15,}\n# Nodes : {:-15,}\n'.format(
                self.flops, self.params, self.nodes))

    def __setattr__(self, key, value):
        if key == '_curr_block':
            self.__dict__[key] = value
            if value not in self._block_list:
                self._block_list.append(value)
        elif key == '_num_blocks':
            raise KeyError('Cannot set _num_blocks manually.')
        else:
            super(ConvNet, self).__setattr__(key, value)

    def _init_params(self, **kwargs):
        pass

    def _build_model(self):
        raise NotImplementedError

    def _make_inputs(self):
        """Network input: images arrive fp32 NHWC in [0,1]; the net sees (X - mean)*scale cast to
        the compute dtype (reference convnet.py:452,466,471; zero-pad/crop are no-ops at ratio 0
        with image size == input size)."""
        n = self._batch_size
        h, w, c = self.input_size
        if kwargs_get(self._parameters, 'zero_pad_ratio', 0.0) != 0.0:
            raise NotImplementedError('zero_pad_ratio != 0 is not supported (input pipeline is out of scope)')
        # images of the data set may be larger than the network input: centre crop (no augmentation),
        # reference convnet.py:453-456,1137-1149.  `input_dtype='u8'` takes raw uint8 images (they are
        # divided by 255 on the device) — a quarter of the host->device bytes of fp32 batches.
        hi, wi = kwargs_get(self._parameters, 'image_size', (h, w, c))[:2]
        if hi < h or wi < w:
            raise ValueError('image_size %s is smaller than the network input %s' % ((hi, wi), (h, w)))
        in_dt = kwargs_get(self._parameters, 'input_dtype', 'f32')
        if in_dt not in ('f32', 'u8'):
            raise ValueError("input_dtype must be 'f32' or 'u8'")
        x_in = self.graph.placeholder('X', (n, int(hi), int(wi), c), in_dt)
        node = self.graph._add('input_prep', [x_in], [(n, h, w, c)], [self._dtype],
                               {'mean': self.image_mean, 'scale': self.scale_factor})
        return x_in, node.outputs[0]

    def _init_model(self, **kwargs):
        """One tower (reference convnet.py:431-499 builds one per device in-process)."""
        self._curr_device = 0
        self._curr_block = None
        self.X_in, self.X = self._make_inputs()
        # labels: float class index, NaN -> -1 fake label, one-hot of -1 is the zero row
        # (convnet.py:441-449).  The device takes int32 class indices.
        self.Y = self.graph.placeholder('Y', (self._batch_size,), 'i32')
        self.d = self._build_model()
        self._reuse = True
        if not self.backbone_only:
            self.logits = self._to_f32(self.d['logits'])
            self.d['logits'] = self.logits
            self.pred = self.d['pred']
            self.losses.append(self._build_loss(**kwargs))
            self.loss = self.losses[0]
        self.dicts.append(self.d)

    def _to_f32(self, x):
        if x.dtype == 'f32':
            return x
        return self.graph._add('cast', [x], [x.shape], ['f32']).outputs[0]

    def _build_loss(self, **kwargs):
        """softmax CE averaged over ALL rows, weighted and masked, + L2/L1 regularisers
        (reference convnet.py:528-597)."""
        l1_factor = kwargs.get('l1_reg', 0e-8)
        l2_factor = kwargs.get('l2_reg', 1e-4)
        ls_factor = kwargs.get('label_smoothing', 0.0)
        w = self.loss_weights
        w = None if w is None else np.array(w, dtype=np.float32)
        logits = self.logits
        rows = int(np.prod(logits.shape[:-1]))
        # the smoothing rule is the task base's _label_smoothing (convnet.py:603-607 uniform;
        # segmentation/segnet.py:116-121 5x5 spatial average of the one-hot map)
        seg_hw = self._label_smoothing_map() if ls_factor > 0.0 else None
        node = self.graph._add('softmax_xent', [logits, self.Y], [()], ['f32'],
                               {'class_weights': w, 'label_smoothing': float(ls_factor),
                                'seg_hw': seg_hw,
                                'focal_gamma': float(kwargs.get('focal_loss_factor', 0.0)),
                                'sigmoid_focal_alpha': float(kwargs.get('sigmoid_focal_loss_factor', 0.0)),
                                'rows': rows, 'l2': float(l2_factor), 'l1': float(l1_factor),
                                'bias_norm_decay': bool(kwargs.get('bias_norm_decay', False))})
        loss = node.outputs[0]
        self.graph.losses.append(loss)
        return loss

    def _label_smoothing_map(self):
        """None: labels*(1-ls) + ls/num_classes (convnet.py:603-607).  Task bases whose smoothing is
        spatial return the (H, W) of their label maps."""
        return None

    # ------------------------------------------------------------------ properties
    @property
    def name(self):
        return 'ConvNet'

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
    def model_scope(self):
        return self._model_scope

    @property
    def backbone_only(self):
        return self._backbone_only

    @property
    def dtype(self):
        return self._dtype

    @property
    def channel_first(self):
        return self._channel_first

    @property
    def argmax_output(self):
        return self._argmax_output

    @property
    def num_devices(self):
        return self._num_devices

    @property
    def compute_device(self):
        return self._compute_device

    @property
    def device_offset(self):
        return self._device_offset

    @property
    def param_device(self):
        return self._param_device

    @property
    def block_list(self):
        return tuple(self._block_list)

    @property
    def num_blocks(self):
        # Number of integer blocks that already own a variable.  The reference counts every
        # registered block including None (convnet.py:239-248), which makes dcgan.py:46 compute a
        # 1x1 seed; counting integer blocks with variables restores the 4x4 seed / 64x64 output
        # the GAN optimiser also assumes.
        return len([b for b in self._block_list
                    if b is not None and self.get_collection('block_{}/variables'.format(b))])

    @property
    def flops(self):
        return self._flops

    @property
    def params(self):
        return self._params

    @property
    def nodes(self):
        return self._nodes

    @property
    def layer_info(self):
        return self._layer_info

    @property
    def dropout_weights(self):
        return self._dropout_weights

    @property
    def dropout_features(self):
        return self._dropout_features

    @property
    def blocks_to_train(self):
        return self._blocks_to_train

    @property
    def update_batch_norm(self):
        return self._update_batch_norm

    @property
    def moving_average_decay(self):
        return self._moving_average_decay

    @property
    def batch_norm_decay(self):
        return self._batch_norm_decay

    @property
    def feature_reduction(self):
        return self._feature_reduction

    def add_to_collection(self, name, tensor):
        if self.model_scope is not None:
            name = str(self.model_scope) + '/' + name
        self.graph.add_to_collection(name, tensor)

    def get_collection(self, key):
        if self.model_scope is not None:
            key = str(self.model_scope) + '/' + key
        return self.graph.get_collection(key)

    def close(self):
        pass

    # ------------------------------------------------------------------ variables
    def _trainable(self):
        return self.blocks_to_train is None or self._curr_block in self.blocks_to_train

    def _get_var(self, name, shape, init, kind, trainable=None, storage_shape=None):
        full = tf.current_scope() + '/' + name if tf.current_scope() else name
        if trainable is None:
            trainable = self._trainable()
        v, created = self.graph.get_var(full, shape, init, trainable, kind, self._curr_block,
                                        storage_shape)
        if created:
            coll = {'weight': 'weight_variables', 'bias': 'bias_variables',
                    'norm': 'norm_variables', 'stat': 'norm_statistics'}[kind]
            self.add_to_collection(coll, v)
            self.add_to_collection('block_{}/variables'.format(self._curr_block), v)
            self.add_to_collection('block_{}/{}'.format(self._curr_block, coll), v)
        return v, created

    def weight_variable(self, shape, initializer=tf.initializers.he_normal(),
                        weight_standardization=False, paddings=((0, 0), (0, 0)), name='weights'):
        if any(p != 0 for pp in paddings for p in pp):
            raise NotImplementedError('kernel_paddings are not supported')
        v, _ = self._get_var(name, shape, initializer, 'weight')
        # weight standardisation (reference convnet.py:1410-1419) happens inside the graph: the layer
        # that asked for it flags its node (attrs['ws']) and the plan standardises per step
        self._ws_requested = bool(weight_standardization)
        return v

    def bias_variable(self, shape, initializer=tf.initializers.zeros(), name='biases'):
        if not isinstance(shape, (list, tuple)):
            shape = [shape]
        v, _ = self._get_var(name, shape, initializer, 'bias')
        return v

    def _count(self, name, shape, flops, params, nodes, created):
        self._flops += flops
        self._nodes += nodes
        self._layer_info.append({'name': name, 'shape': shape, 'flops': int(flops),
                                 'params': int(params), 'nodes': int(nodes)})
        if created:
            self._params += params

    # ------------------------------------------------------------------ pooling
    def pooling_layer(self, x, kernel, stride, padding='SAME', pooling_type='AVG'):
        if pooling_type.lower() == 'avg':
            return self.avg_pool(x, kernel, stride, padding=padding)
        elif pooling_type.lower() == 'max':
            return self.max_pool(x, kernel, stride, padding=padding)
        else:
            raise ValueError('Pooling type of {} is not supported'.format(pooling_type))

    def _pool(self, kind, x, side_l, stride, padding):
        side_l, stride = _pair(side_l), _pair(stride)
        n, h, w, c = x.shape
        ho, pt, _ = same_pad(h, side_l[0], stride[0], 1, padding)
        wo, pl, _ = same_pad(w, side_l[1], stride[1], 1, padding)
        flops = side_l[0] * side_l[1] * ho * wo * c
        self._count(tf.current_scope() + '/' + kind, [None, ho, wo, c], flops, 0, ho * wo * c, False)
        node = self.graph._add(kind, [x], [(n, ho, wo, c)], [x.dtype],
                               {'k': side_l, 's': stride, 'pad': (pt, pl)}, tf.current_scope())
        return node.outputs[0]

    def max_pool(self, x, side_l, stride, padding='SAME'):
        return self._pool('max_pool', x, side_l, stride, padding)

    def avg_pool(self, x, side_l, stride, padding='SAME'):
        return self._pool('avg_pool', x, side_l, stride, padding)

    # ------------------------------------------------------------------ convolution / dense
    def conv_bn_act(self, x, kernel, stride, out_channels=None, padding='SAME', biased=False,
                    depthwise=False, scope=None, dilation=(1, 1), ws=False,
                    kernel_paddings=((0, 0), (0, 0)), weight_initializer=tf.initializers.he_normal(),
                    bias_initializer=tf.initializers.zeros(), scale=True, shift=True,
                    zero_scale_init=False, epsilon=1e-3, act_type='relu', act_params=None,
                    verbose=False):
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            x = self.conv_layer(x, kernel, stride, out_channels, padding=padding, biased=biased,
                                depthwise=depthwise, dilation=dilation, ws=ws,
                                kernel_paddings=kernel_paddings, weight_initializer=weight_initializer,
                                bias_initializer=bias_initializer)
            x = self.batch_norm(x, scale=scale, shift=shift, zero_scale_init=zero_scale_init,
                                epsilon=epsilon)
            x = self.activation(x, activation_type=act_type, params=act_params)
        return x

    def conv_layer(self, x, kernel, stride, out_channels=None, padding='SAME', biased=True,
                   depthwise=False, scope=None, dilation=(1, 1), ws=False,
                   kernel_paddings=((0, 0), (0, 0)), weight_initializer=tf.initializers.he_normal(),
                   bias_initializer=tf.initializers.zeros(), verbose=False):
        kernel, stride, dilation = _pair(kernel), _pair(stride), _pair(dilation)
        n, h, w, in_channels = x.shape
        ho, pt, _ = same_pad(h, kernel[0], stride[0], dilation[0], padding)
        wo, pl, _ = same_pad(w, kernel[1], stride[1], dilation[1], padding)
        if out_channels is None:
            out_channels = in_channels
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            sc = tf.current_scope()
            if depthwise:
                mult = max(out_channels // in_channels, 1)
                out_c = in_channels * mult
                wshape = [kernel[0], kernel[1], in_channels, mult]
                flops = ho * wo * kernel[0] * kernel[1] * in_channels * mult
                op = 'dwconv2d'
            else:
                mult = 0
                out_c = out_channels
                wshape = [kernel[0], kernel[1], in_channels, out_channels]
                flops = ho * wo * kernel[0] * kernel[1] * in_channels * out_channels
                op = 'conv2d'
            params = int(np.prod(wshape))
            weights = self.weight_variable(wshape, initializer=weight_initializer,
                                           weight_standardization=ws, paddings=kernel_paddings)
            created = weights.name not in self.__dict__.setdefault('_counted', set())
            self._counted.add(weights.name)
            attrs = {'k': kernel, 's': stride, 'd': dilation, 'pad': (pt, pl), 'mult': mult,
                     'biased': bool(biased)}
            attrs['ws'] = bool(ws)
            node = self.graph._add(op, [x], [(n, ho, wo, out_c)], [x.dtype], attrs, sc)
            node.vars['w'] = weights
            if biased:
                node.vars['b'] = self.bias_variable(out_c, initializer=bias_initializer)
                flops += ho * wo * out_c
                params += out_c
            self._count(sc, [None, ho, wo, out_c], flops, params, ho * wo * out_c, created)
        return node.outputs[0]

    def transposed_conv_layer(self, x, kernel, stride, out_channels, padding='SAME', biased=True,
                              output_shape=None, dilation=(1, 1), scope=None,
                              weight_initializer=tf.initializers.he_normal(),
                              bias_initializer=tf.initializers.zeros(), ws=False, verbose=False):
        kernel, stride, dilation = _pair(kernel), _pair(stride), _pair(dilation)
        n, h, w, in_channels = x.shape
        if output_shape is None:
            if padding.lower() == 'valid':
                out_hw = [h * stride[0] - kernel[0] + 1, w * stride[1] - kernel[1] + 1]
            else:
                out_hw = [h * stride[0], w * stride[1]]
        else:
            out_hw = list(output_shape[1:3])
        # conv2d_transpose == input-gradient of the conv that maps out_hw -> (h, w)
        ho, pt, _ = same_pad(out_hw[0], kernel[0], stride[0], dilation[0], padding)
        wo, pl, _ = same_pad(out_hw[1], kernel[1], stride[1], dilation[1], padding)
        if (ho, wo) != (h, w):
            raise ValueError('transposed conv: output shape %s inconsistent with input %s' % (out_hw, (h, w)))
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            sc = tf.current_scope()
            weights = self.weight_variable([kernel[0], kernel[1], in_channels, out_channels],
                                           initializer=weight_initializer, weight_standardization=ws)
            created = weights.name not in self.__dict__.setdefault('_counted', set())
            self._counted.add(weights.name)
            flops = out_hw[0] * out_hw[1] * kernel[0] * kernel[1] * in_channels * out_channels
            params = kernel[0] * kernel[1] * in_channels * out_channels
            attrs = {'k': kernel, 's': stride, 'd': dilation, 'pad': (pt, pl), 'biased': bool(biased),
                     'ws': bool(ws)}
            node = self.graph._add('conv2d_transpose', [x], [(n, out_hw[0], out_hw[1], out_channels)],
                                   [x.dtype], attrs, sc)
            node.vars['w'] = weights
            if biased:
                node.vars['b'] = self.bias_variable(out_channels, initializer=bias_initializer)
                flops += out_hw[0] * out_hw[1] * out_channels
                params += out_channels
            self._count(sc, [None] + out_hw + [out_channels], flops, params,
                        out_hw[0] * out_hw[1] * out_channels, created)
        return node.outputs[0]

    def fc_layer(self, x, out_dim, biased=True, scope=None, ws=False,
                 weight_initializer=tf.initializers.he_normal(),
                 bias_initializer=tf.initializers.zeros(), verbose=False):
        in_dim = int(x.get_shape()[-1])
        if len(x.shape) != 2:
            raise ValueError('fc_layer expects a [N, in_dim] tensor')
        with tf.variable_scope(scope) if scope is not None else nullcontext():
            sc = tf.current_scope()
            weights = self.weight_variable([in_dim, out_dim], initializer=weight_initializer,
                                           weight_standardization=ws)
            created = weights.name not in self.__dict__.setdefault('_counted', set())
            self._counted.add(weights.name)
            flops = in_dim * out_dim
            params = in_dim * out_dim
            node = self.graph._add('dense', [x], [(x.shape[0], out_dim)], [x.dtype],
                                   {'biased': bool(biased), 'ws': bool(ws)}, sc)
            node.vars['w'] = weights
            if biased:
                node.vars['b'] = self.bias_variable(out_dim, initializer=bias_initializer)
                flops += out_dim
                params += out_dim
            self._count(sc, [None, out_dim], flops, params, out_dim, created)
        return node.outputs[0]

    # ------------------------------------------------------------------ normalisation
    def normalization(self, x, norm_type='batch', norm_param=None, scale=True, shift=True,
                      zero_scale_init=False, epsilon=1e-3, scope='norm'):
        supported_types = ['batch', 'group', 'grouped_batch']
        if norm_type is None:
            return x
        elif norm_type.lower() == 'batch':
            return self.batch_norm(x, scale=scale, shift=shift, zero_scale_init=zero_scale_init,
                                   epsilon=epsilon, scope=scope)
        elif norm_type.lower() == 'group':
            return self.group_norm(x, num_groups=32 if norm_param is None else norm_param, scale=scale,
                                   shift=shift, zero_scale_init=zero_scale_init, epsilon=epsilon, scope=scope)
        elif norm_type.lower() == 'grouped_batch':
            # "Experimental grouped batch normalization" of the reference (convnet.py:2175): no model
            # file uses it; out of scope
            raise NotImplementedError('norm_type grouped_batch (experimental in the reference) is not supported')
        else:
            raise ValueError('Normalization type of {} is not supported. Supported types: {}'
                             .format(norm_type, supported_types))

    def batch_norm(self, x, scale=True, shift=True, zero_scale_init=False, epsilon=1e-3, scope='bn'):
        if isinstance(self.update_batch_norm, bool):
            update = self.update_batch_norm
        else:
            update = self._trainable()
        trainable = self._trainable()
        c = x.shape[-1]
        hw = int(np.prod(x.shape[1:-1]))
        with tf.variable_scope(scope):
            sc = tf.current_scope()
            mu, _ = self._get_var('mu', [c], tf.zeros_initializer(), 'stat', trainable=False)
            sigma, _ = self._get_var('sigma', [c], tf.ones_initializer(), 'stat', trainable=False)
            gamma = beta = None
            if scale:
                init = tf.zeros_initializer() if zero_scale_init else tf.ones_initializer()
                gamma, created = self._get_var('gamma', [c], init, 'norm', trainable=trainable)
                if created:
                    self._params += c
            if shift:
                beta, created = self._get_var('beta', [c], tf.zeros_initializer(), 'norm',
                                              trainable=trainable)
                if created:
                    self._params += c
                self._flops += hw * c
            node = self.graph._add('bn', [x], [x.shape], [x.dtype],
                                   {'eps': float(epsilon), 'momentum': float(self.batch_norm_decay),
                                    'update': bool(update), 'act': 0, 'alpha': 0.0}, sc)
            node.vars.update({'mu': mu, 'sigma': sigma})
            if gamma is not None:
                node.vars['gamma'] = gamma
            if beta is not None:
                node.vars['beta'] = beta
        return node.outputs[0]

    def group_norm(self, x, num_groups=32, scale=True, shift=True, zero_scale_init=False, epsilon=1e-3,
                   scope='gn'):
        """Reference convnet.py:1928-2013: per-sample moments over H x W x (C / num_groups), then the
        per-channel affine; gamma / beta are ordinary trainable norm variables (no moving statistics)."""
        trainable = self._trainable()
        c = x.shape[-1]
        hw = int(np.prod(x.shape[1:-1])) if len(x.shape) > 2 else 1
        assert c // num_groups * num_groups == c, \
            'Number of channels must be a multiple of num_groups ({})'.format(num_groups)
        with tf.variable_scope(scope):
            sc = tf.current_scope()
            gamma = beta = None
            if scale:
                init = tf.zeros_initializer() if zero_scale_init else tf.ones_initializer()
                gamma, created = self._get_var('gamma', [c], init, 'norm', trainable=trainable)
                if created:
                    self._params += c
                self._flops += hw * c
            if shift:
                beta, created = self._get_var('beta', [c], tf.zeros_initializer(), 'norm', trainable=trainable)
                if created:
                    self._params += c
                self._flops += hw * c
            node = self.graph._add('gn', [x], [x.shape], [x.dtype],
                                   {'eps': float(epsilon), 'groups': int(num_groups)}, sc)
            if gamma is not None:
                node.vars['gamma'] = gamma
            if beta is not None:
                node.vars['beta'] = beta
        return node.outputs[0]

    # ------------------------------------------------------------------ resize
    def upsampling_2d_layer(self, x, scale=2, out_shape=None, align_corners=False,
                            force_unaligned=False, upsampling_method='bilinear', name='upsampling'):
        in_shape = x.get_shape()
        if out_shape is None:
            out_shape = [in_shape[1] * scale, in_shape[2] * scale]
        out_shape = [int(s) for s in out_shape]
        if force_unaligned:
            mode = 0
        else:
            mode = 1 if align_corners else 2
        if upsampling_method.lower() == 'bilinear':
            node = self.graph._add('resize_bilinear', [x],
                                   [(x.shape[0], out_shape[0], out_shape[1], x.shape[3])], [x.dtype],
                                   {'mode': mode}, tf.current_scope() + '/' + name)
            return node.outputs[0]
        elif upsampling_method.lower() in ('nearest', 'nearest_neighbor'):
            node = self.graph._add('resize_nearest', [x],
                                   [(x.shape[0], out_shape[0], out_shape[1], x.shape[3])], [x.dtype],
                                   {'mode': mode}, tf.current_scope() + '/' + name)
            return node.outputs[0]
        raise ValueError('Upsampling method of {} is not supported'.format(upsampling_method))

    # ------------------------------------------------------------------ residual / activations
    def stochastic_depth(self, x, skip, drop_rate=0.0, name='drop'):
        # x*survived + skip with a per-sample Bernoulli keep (reference convnet.py:2500-2512); the
        # keep mask is drawn on the device (csrc/dropout.cu), inference adds the branches unchanged
        if drop_rate > 0.0:
            if x.shape != skip.shape:
                raise ValueError('stochastic_depth: shape mismatch %s vs %s' % (x.shape, skip.shape))
            node = self.graph._add('sd_add', [x, skip], [x.shape], [x.dtype],
                                   {'rate': float(drop_rate), 'layer': self.graph.next_random_layer()},
                                   tf.current_scope() + '/' + name)
            return node.outputs[0]
        return x + skip

    def activation(self, x, activation_type='relu', params=None):
        supported_types = ['relu', 'relu6', 'lrelu', 'tanh', 'sigmoid', 'swish']
        if activation_type is None:
            return x
        act = activation_type.lower()
        if act in ('relu', 'relu6', 'lrelu', 'leaky_relu', 'tanh', 'sigmoid', 'swish'):
            return self.graph.activation(x, act, params)
        raise ValueError('Activation type of {} is not supported. Supported types: {}'
                         .format(activation_type, supported_types))

    def relu(self, x, name='relu'):
        return self.graph.activation(x, 'relu')

    def relu6(self, x, name='relu6'):
        return self.graph.activation(x, 'relu6')

    def lrelu(self, x, alpha=None, name='lrelu'):
        return self.graph.activation(x, 'lrelu', 0.2 if alpha is None else alpha)

    def tanh(self, x, name='tanh'):
        return self.graph.activation(x, 'tanh')

    def sigmoid(self, x, name=None):
        return self.graph.activation(x, 'sigmoid')

    def swish(self, x, name='swish'):
        return self.graph.activation(x, 'swish')


def kwargs_get(d, key, default):
    return d.get(key, default) if d else default
