"""Model builders written directly against the ConvNet facade.

The north-star models are the reference's own files, loaded unchanged by loader.py.  This module
only provides a self-contained ResNet-v1.5-50 definition (same layers, scopes and variable names
as reference models/resnet_v1_5.py:8-209 produces) so that bench.py and smoke() still run on a
machine where the reference files were never staged.  tests/test_graph.py asserts that it builds
the identical graph (ops, shapes, variable names) as the reference file.
"""
from . import tfshim as tf
from .convnet import ConvNet


class ResNet50(ConvNet):
    channels = [64, 256, 512, 1024, 2048]
    units = [None, 3, 4, 6, 3]
    strides = [2, 1, 2, 2, 2]

    def _bottleneck(self, x, stride, out_channels, d, name):
        in_channels = x.get_shape()[-1]
        with tf.variable_scope(name):
            if in_channels == out_channels:
                skip = self.max_pool(x, stride, stride, padding='VALID') if stride > 1 else x
            else:
                with tf.variable_scope('conv_skip'):
                    skip = self.conv_layer(x, 1, stride, out_channels, padding='SAME', biased=False)
                    skip = self.normalization(skip, scope='bn')
            d[name + '/branch'] = skip
            specs = [(1, 1, out_channels // 4, True, False), (3, stride, out_channels // 4, True, False),
                     (1, 1, out_channels, False, True)]
            for i, (k, s, c, act, zero) in enumerate(specs):
                with tf.variable_scope('conv_%d' % i):
                    x = self.conv_layer(x, k, s, c, padding='SAME', biased=False)
                    d['%s/conv_%d' % (name, i)] = x
                    x = self.normalization(x, scope='bn', zero_scale_init=zero)
                    d['%s/conv_%d/bn' % (name, i)] = x
                    if act:
                        x = self.relu(x)
                        d['%s/conv_%d/relu' % (name, i)] = x
            x = self.relu(self.stochastic_depth(x, skip))
            d[name] = x
        return x

    def _build_model(self):
        d = {}
        self._curr_block = 0
        with tf.variable_scope('block_0'):
            with tf.variable_scope('conv_0'):
                x = self.conv_layer(self.X, 7, 2, 64, padding='SAME', biased=False)
                d['block_0/conv_0'] = x
                x = self.normalization(x, scope='bn')
                d['block_0/conv_0/bn'] = x
                x = self.relu(x)
                d['block_0/conv_0/relu'] = x
                x = self.max_pool(x, 3, 2, padding='SAME')
                d['block_0/conv_0/maxpool'] = x
            d['block_0'] = x
        for i in range(1, 5):
            self._curr_block = i
            for j in range(self.units[i]):
                x = self._bottleneck(x, self.strides[i] if j == 0 else 1, self.channels[i], d,
                                     'block_%d/res_%d' % (i, j))
            d['block_%d' % i] = x
        if not self.backbone_only:
            self._curr_block = None
            with tf.variable_scope('block_None'):
                with tf.variable_scope('logits'):
                    x = tf.reduce_mean(x, axis=[1, 2])
                    d['logits/avgpool'] = x
                    x = tf.nn.dropout(x, rate=self.dropout_rate_features)
                    x = self.fc_layer(x, self.num_classes)
                    d['logits'] = x
                    d['pred'] = tf.nn.softmax(x)
        return d


def resnet50(input_shape, num_classes, prefer_reference=True, **kwargs):
    """ResNet-v1.5-50 on the facade: the reference's file when staged, else the builder above."""
    from . import convnet, loader
    if prefer_reference and loader.reference_root() is not None:
        mod = loader.load_reference_model('models/resnet_v1_5.py', {'convnet': convnet})
        return mod.ResNet50(input_shape, num_classes, **kwargs), 'reference models/resnet_v1_5.py (unchanged)'
    return ResNet50(input_shape, num_classes, **kwargs), 'myconvnet_b200.zoo.ResNet50'
