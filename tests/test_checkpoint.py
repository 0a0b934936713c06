"""Checkpoint naming and .npz round trips (reference models/init_from_checkpoint.py)."""
import numpy as np

from myconvnet_b200 import checkpoint as ck


def test_slim_resnet_names_follow_the_reference_mapping():
    f = ck.slim_resnet_v1_name
    assert f("block_0/conv_0/weights") == "resnet_v1_50/conv1/weights"
    assert f("block_0/conv_0/bn/mu") == "resnet_v1_50/conv1/BatchNorm/moving_mean"
    assert f("block_1/res_0/conv_0/weights") == "resnet_v1_50/block1/unit_1/bottleneck_v1/conv1/weights"
    assert f("block_2/res_1/conv_0/bn/gamma") == "resnet_v1_50/block2/unit_2/bottleneck_v1/conv1/BatchNorm/gamma"
    assert f("block_3/res_5/conv_2/bn/sigma") == \
        "resnet_v1_50/block3/unit_6/bottleneck_v1/conv3/BatchNorm/moving_variance"
    assert f("block_4/res_0/conv_skip/weights") == "resnet_v1_50/block4/unit_1/bottleneck_v1/shortcut/weights"
    assert f("block_4/res_0/conv_skip/bn/beta") == "resnet_v1_50/block4/unit_1/bottleneck_v1/shortcut/BatchNorm/beta"
    assert f("block_None/logits/biases", depth=101) == "resnet_v1_101/logits/biases"
    assert f("block_1/res_0/conv_0/weights" + ck.EMA_SUFFIX).endswith("conv1/weights/ExponentialMovingAverage")
    assert f("global_step") is None and f("block_1/res_0/conv_0/bn/unknown") is None


def test_every_resnet50_variable_has_a_distinct_slim_name():
    from myconvnet_b200.zoo import ResNet50
    m = ResNet50([64, 64, 3], 10, batch_size=2, compute_dtype="f32")
    names = [v.name for v in m.graph.vars.values()]
    slim = [ck.slim_resnet_v1_name(n) for n in names]
    assert all(s is not None for s in slim) and len(set(slim)) == len(names) == 267


def test_npz_round_trip_with_shadows_and_slim_import(tmp_path):
    rng = np.random.default_rng(0)
    shapes = {"block_0/conv_0/weights": (7, 7, 3, 64), "block_0/conv_0/bn/mu": (64,),
              "block_1/res_0/conv_skip/weights": (1, 1, 64, 256), "block_None/logits/weights": (2048, 10)}
    var = {k: rng.standard_normal(s).astype(np.float32) for k, s in shapes.items()}
    ema = {k: v * 0.5 for k, v in var.items()}
    p = str(tmp_path / "ref.npz")
    ck.save_npz(p, var, ema)
    v2, e2 = ck.load_npz(p)
    assert set(v2) == set(var) and all(np.array_equal(v2[k], var[k]) and np.array_equal(e2[k], ema[k]) for k in var)
    # slim naming, a 1x1-conv logits tensor and a shape mismatch, shadows preferred
    ps = str(tmp_path / "slim.npz")
    slim_var = dict(var)
    slim_var["block_None/logits/weights"] = var["block_None/logits/weights"].reshape(1, 1, 2048, 10)
    slim_var["block_0/conv_0/bn/mu"] = np.zeros(32, np.float32)                    # wrong shape: skipped
    ck.save_npz(ps, slim_var, {"block_0/conv_0/weights": ema["block_0/conv_0/weights"]}, naming="slim")
    with np.load(ps) as z:
        assert "resnet_v1_50|conv1|weights" in z.files and "resnet_v1_50|logits|weights" in z.files
    v3, e3 = ck.load_npz(ps, expected=shapes, naming="slim", prefer_ema=True)
    assert "block_0/conv_0/bn/mu" not in v3
    assert np.array_equal(v3["block_0/conv_0/weights"], ema["block_0/conv_0/weights"])     # the shadow wins
    assert np.array_equal(v3["block_None/logits/weights"], var["block_None/logits/weights"])
    assert np.array_equal(v3["block_1/res_0/conv_skip/weights"], var["block_1/res_0/conv_skip/weights"])
    assert set(e3) == {"block_0/conv_0/weights"}
