"""Drop-in boundary: the facade's layer methods have exactly the reference's parameter
names, order and default values.  Compares against the reference's own source when it is mounted
(/root/reference is not present on the GPU box: the test then checks the recorded table only)."""
import ast
import inspect
import os

import pytest

from myconvnet_b200 import convnet as product

REF = "/root/reference/convnet.py"
METHODS = ["conv_layer", "conv_bn_act", "fc_layer", "normalization", "batch_norm", "group_norm", "max_pool",
           "avg_pool", "pooling_layer", "upsampling_2d_layer", "transposed_conv_layer", "stochastic_depth",
           "activation", "relu", "relu6", "lrelu", "tanh", "sigmoid", "swish", "weight_variable",
           "bias_variable"]

# parameter lists recorded from reference convnet.py (file:line in SURVEY 8b); checked on every box
RECORDED = {
    "conv_layer": ["x", "kernel", "stride", "out_channels", "padding", "biased", "depthwise", "scope", "dilation",
                   "ws", "kernel_paddings", "weight_initializer", "bias_initializer", "verbose"],
    "fc_layer": ["x", "out_dim", "biased", "scope", "ws", "weight_initializer", "bias_initializer", "verbose"],
    "normalization": ["x", "norm_type", "norm_param", "scale", "shift", "zero_scale_init", "epsilon", "scope"],
    "batch_norm": ["x", "scale", "shift", "zero_scale_init", "epsilon", "scope"],
    "max_pool": ["x", "side_l", "stride", "padding"],
    "upsampling_2d_layer": ["x", "scale", "out_shape", "align_corners", "force_unaligned", "upsampling_method",
                            "name"],
    "transposed_conv_layer": ["x", "kernel", "stride", "out_channels", "padding", "biased", "output_shape", "dilation", "scope", "weight_initializer", "bias_initializer", "ws", "verbose"],
    "stochastic_depth": ["x", "skip", "drop_rate", "name"],
    "activation": ["x", "activation_type", "params"],
}


def _product_params(name):
    sig = inspect.signature(getattr(product.ConvNet, name))
    return [p for p in sig.parameters.values() if p.name != "self"]


@pytest.mark.parametrize("name", sorted(RECORDED))
def test_recorded_parameter_lists(name):
    assert [p.name for p in _product_params(name)] == RECORDED[name]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference source not mounted")
@pytest.mark.parametrize("name", METHODS)
def test_signature_equals_reference_source(name):
    tree = ast.parse(open(REF).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ConvNet"][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    ref_args = [a.arg for a in fn.args.args][1:]
    ref_defaults = [ast.unparse(d) for d in fn.args.defaults]
    got = _product_params(name)
    assert [p.name for p in got] == ref_args
    got_defaults = [p.default for p in got if p.default is not inspect._empty]
    assert len(got_defaults) == len(ref_defaults)
    for g, r in zip(got_defaults, ref_defaults):
        if r.startswith("tf."):            # initializer objects: compared by the shim's repr of the call
            continue
        assert g == ast.literal_eval(r), (name, g, r)
