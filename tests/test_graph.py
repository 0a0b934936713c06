"""Host logic without a GPU: the reference model files build unchanged on the facade, the
structure KATs of SURVEY.md 8c hold, and the planner fuses / differentiates as designed."""
import numpy as np
import pytest

from myconvnet_b200 import convnet, loader
from myconvnet_b200.plan import Plan


@pytest.fixture(scope="module")
def r50(have_reference_models):
    mod = loader.load_reference_model("models/resnet_v1_5.py", {"convnet": convnet})
    return mod.ResNet50([224, 224, 3], 1000, batch_size=256, compute_dtype="bf16")


def test_resnet50_structure_kats(r50):
    g = r50.graph
    assert r50.params == 25557032                       # reference's own parameter counter
    convs = [n for n in g.nodes if n.op == "conv2d"]
    bns = [n for n in g.nodes if n.op == "bn"]
    assert len(convs) == 53 and len(bns) == 53
    macs = sum(n.outputs[0].shape[1] * n.outputs[0].shape[2] * int(np.prod(n.vars["w"].shape)) for n in convs)
    macs += 2048 * 1000
    assert macs == 4089184256                           # fwd MAC/img
    assert len(g.vars) == 267 and sum(v.trainable for v in g.vars.values()) == 161
    names = set(g.vars)
    for k in ("block_0/conv_0/weights", "block_0/conv_0/bn/mu", "block_1/res_0/conv_skip/bn/gamma",
              "block_4/res_2/conv_2/bn/sigma", "block_None/logits/weights", "block_None/logits/biases"):
        assert k in names, k
    # TF SAME asymmetry on the strided layers
    stem = convs[0]
    assert stem.attrs["pad"] == (2, 2) and stem.outputs[0].shape == (256, 112, 112, 64)
    s2 = [n for n in convs if n.attrs["s"] == [2, 2] and n.attrs["k"] == [3, 3]]
    assert all(n.attrs["pad"] == (0, 0) for n in s2) and len(s2) == 3
    assert r50.block_list == (None, 0, 1, 2, 3, 4) and r50.num_blocks == 5


def test_native_builder_matches_reference_graph(r50):
    from myconvnet_b200 import zoo
    m = zoo.ResNet50([224, 224, 3], 1000, batch_size=256, compute_dtype="bf16")
    a = [(n.op, tuple(o.shape for o in n.outputs), tuple(sorted(v.name for v in n.vars.values()))) for n in m.graph.nodes]
    b = [(n.op, tuple(o.shape for o in n.outputs), tuple(sorted(v.name for v in n.vars.values()))) for n in r50.graph.nodes]
    assert a == b
    assert set(m.d) == set(r50.d)


def test_plan_fusion_and_launch_counts(r50):
    p = Plan(r50.graph)
    h = p.launch_histogram()
    # convs with a deep reduction take the BN statistics in their epilogue; the store-bound 1x1
    # expansions keep the separate statistics pass; finalize is folded into the apply kernel
    assert h["mcn_stem_conv_fprop"] == 1 and h["mcn_stem_conv_wgrad"] == 1 and h["mcn_pad_rgb4"] == 1
    assert h["mcn_conv2d_fprop_tc_stats"] + h["mcn_conv2d_fprop_tc"] == 53          # 52 convs + dense
    # every BN layer's statistics ride on its producing conv's epilogue (+1: the stem fuses too);
    # the round-1 rule (MCN_FUSE_STATS_RULE=k) left the store-bound 1x1 expansions on mcn_bn_stats
    assert h["mcn_conv2d_fprop_tc_stats"] + 1 == 53 and "mcn_bn_stats" not in h
    assert "mcn_bn_finalize" not in h
    assert h["mcn_conv2d_wgrad_tc"] == 53
    # no dgrad into the images; the stride-1 dgrads that produce the whole gradient of a BN+ReLU
    # output (conv_1 / conv_2 of every unit, minus the three stride-2 3x3) also take that layer's
    # backward sums in their epilogue, which replaces its reduction pass
    assert h["mcn_conv2d_dgrad_tc"] + h["mcn_conv2d_dgrad_tc_bnred"] == 53
    # (their last block also writes the final sums: no mcn_bn_bwd_finalize launch)
    assert h["mcn_conv2d_dgrad_tc_bnred"] == 29 and "mcn_bn_bwd_finalize" not in h
    # the 16 layers with a fused residual leave a ReLU bit mask in the forward pass; their backward
    # passes read it instead of the output tensor
    assert h["mcn_bn_bwd_reduce_mask"] == 16 == h["mcn_bn_apply_stats_mask"] == h["mcn_bn_bwd_apply_mask"]
    assert h["mcn_bn_bwd_reduce"] == 53 - 29 - 16
    assert h["mcn_bn_apply_stats"] + h["mcn_bn_apply_stats_mask"] == 53
    assert "mcn_act_fwd" not in h and "mcn_add_act_fwd" not in h
    # the unfused plan keeps the separate statistics pass
    hu = Plan(r50.graph, fuse_bn_stats=False).launch_histogram()
    assert hu["mcn_conv2d_fprop_tc"] == 53 and hu["mcn_bn_stats"] == 53 and hu["mcn_bn_apply_stats"] + hu["mcn_bn_apply_stats_mask"] == 53
    assert "mcn_accumulate" not in h                    # multi-consumer gradients add in the dgrad epilogue
    fused = [n for n in r50.graph.nodes if n.op == "bn"]
    assert sum(n.attrs["residual"] is not None for n in fused) == 16
    assert sum(n.attrs["act"] == 1 for n in fused) == 49
    # arena fits comfortably in 180 GB and regions do not overlap
    spans = sorted(p.region_span.values())
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    assert p.arena_bytes < 40e9
    offs = sorted((b.offset, b.offset + b.nbytes) for b in p.bufs if b.region != "temp")
    assert all(offs[i][1] <= offs[i + 1][0] for i in range(len(offs) - 1))
    # stem: the RGB input is gathered from a 4-channel copy of the image; the weight is stored
    # [Kpad=256][64] with row (r*8 + s)*4 + c (8th tap and 4th channel zero)
    stem = [n for n in r50.graph.nodes if n.op == "conv2d"][0]
    assert stem.attrs["route"] == "stem" and stem.attrs["kpad"] == 256
    w = stem.vars["w"]
    assert w.storage_shape == (256, 64) and len(w.storage_rows) == 147
    assert w.storage_rows[3] == 4 and w.storage_rows[21] == 32 and w.storage_rows.max() == 6 * 32 + 6 * 4 + 2
    # with the gather stem disabled the explicit im2col route is used (K padded to 152)
    import os
    os.environ["MCN_STEM_GATHER"] = "0"
    try:
        p2 = Plan(r50.graph)
        assert stem.attrs["route"] == "im2col" and stem.attrs["kpad"] == 152
        assert stem.vars["w"].storage_shape == (152, 64) and p2.launch_histogram()["mcn_im2col"] == 1
    finally:
        del os.environ["MCN_STEM_GATHER"]
        Plan(r50.graph)          # restore the default layout on the shared fixture


def test_plan_keeps_taps_unfused(r50):
    taps = [t for k, t in r50.d.items() if k.endswith("/bn")]
    p = Plan(r50.graph, keep=taps)
    assert p.launch_histogram().get("mcn_act_fwd", 0) > 0
    for t in taps:
        assert t in p.tbuf


def test_sync_bn_plan_has_collective_points(r50):
    p = Plan(r50.graph, world_size=8)
    assert len([a for a in p.allreduce_points if a[0] == "f"]) == 53
    assert len([a for a in p.allreduce_points if a[0] == "b"]) == 53


def test_gradient_buckets_become_ready_in_reverse_layer_order(r50):
    """The flat gradient buffer mirrors the variable order (stem first), backward runs last layer
    first: bucket k's all-reduce can start after launch ready[k], long before backward ends."""
    p = Plan(r50.graph, world_size=8)
    sched = p.grad_bucket_schedule(4 * 1024 * 1024)
    assert sched[0][0] == 0 and sched[-1][1] == p.n_train
    assert all(a[1] == b[0] for a, b in zip(sched, sched[1:]))
    ready = [r for _, _, r in sched]
    assert ready == sorted(ready, reverse=True) and ready[-1] < 10 and ready[0] == len(p.bwd) - 1
    # the dense layer's gradients (last bucket) come from the first few backward launches
    tags = [p.bwd[i].tag for i in range(ready[-1] + 1)]
    assert any("logits" in t for t in tags)
    # every launch that writes into the gradient buffer is at or before its bucket's ready index
    for li, l in enumerate(p.bwd):
        for a in l.args:
            if hasattr(a, "buf") and a.buf is p.b_grad:
                e = a.off // 4
                r = [rd for s0, e0, rd in sched if s0 <= e < e0][0]
                assert li <= r, (li, r, l.tag)


def test_fp32_config_uses_exact_path(have_reference_models):
    mod = loader.load_reference_model("models/resnet_v1_5.py", {"convnet": convnet})
    m = mod.ResNet50([224, 224, 3], 1000, batch_size=32)        # BASELINE config 1: fp32, batch 32
    h = Plan(m.graph).launch_histogram()
    assert "mcn_conv2d_fprop_tc" not in h and h["mcn_conv2d_fprop_direct"] == 54


def test_facade_errors_mirror_reference():
    m = convnet.ConvNet([8, 8, 3], 4, auto_build=False, batch_size=2)
    x = m.graph.placeholder("x", (2, 8, 8, 4), "f32")
    with pytest.raises(ValueError):
        m.pooling_layer(x, 2, 2, pooling_type="median")           # convnet.py:1470
    with pytest.raises(ValueError):
        m.normalization(x, norm_type="layer")                     # convnet.py:1776
    with pytest.raises(ValueError):
        m.activation(x, activation_type="gelu")                   # convnet.py:2533
    with pytest.raises(ValueError):
        m.upsampling_2d_layer(x, upsampling_method="bicubic")     # convnet.py:2400
    with pytest.raises(NotImplementedError):
        convnet.ConvNet([8, 8, 3], 4, channel_first=True, batch_size=2)
    assert m.relu(x).shape == (2, 8, 8, 4) and m.activation(x, None) is x


def test_engine_refuses_to_run_without_cuda(r50):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from myconvnet_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU execution"):
        Engine(r50)


def test_other_north_star_models_build_and_plan(have_reference_models):
    """efficientnet.py, deeplabv3plus.py and dcgan.py import unchanged and lower to launch lists."""
    fac = loader.product_facade()
    b0 = loader.load_reference_model("models/efficientnet.py", fac).EfficientNetB0(
        [224, 224, 3], 1000, batch_size=8, compute_dtype="bf16")
    assert b0.params == 5288548                       # 5.25 M non-BN + BN gamma/beta
    h = Plan(b0.graph).launch_histogram()
    assert h["mcn_dwconv2d_fwd"] == 16 and h["mcn_scale_bcast_fwd"] == 16 and h["mcn_bn_apply_stats"] == 49
    assert h["mcn_bn_stats"] >= 16                      # depthwise outputs keep the separate statistics pass
    assert "block_1/mbconv_0/se_mask/conv_0/biases" in b0.graph.vars
    dl = loader.load_reference_model("models/deeplabv3plus.py", fac).DeepLabV3PlusResNet(
        [512, 512, 3], 21, batch_size=2, compute_dtype="bf16")
    dil = sorted({tuple(n.attrs["d"]) for n in dl.graph.nodes if n.op == "conv2d"})
    assert (6, 6) in dil and (12, 12) in dil and (18, 18) in dil and (2, 2) in dil and (8, 8) in dil
    assert dl.logits.shape == (2, 512, 512, 21) and dl.Y.shape == (2, 512, 512)
    hp = Plan(dl.graph).launch_histogram()
    assert hp["mcn_resize_bilinear_fwd"] == 2 and hp["mcn_copy_channels"] > 0
    gan = loader.load_reference_model("models/dcgan.py", fac).DCGAN(
        [64, 64, 3], 100, batch_size=128, compute_dtype="bf16")
    assert gan.generate.shape == (128, 64, 64, 3)       # 4x4 seed: SURVEY Appendix D.1
    vd, vg = gan.gan_variable_split()
    assert {v.name.split("/")[0] for v in vd} == {"block_0", "block_1", "block_2", "block_3", "block_None"}
    assert {v.name.split("/")[0] for v in vg} == {"block_%d" % i for i in range(4, 9)}
    p = Plan(gan.graph)
    assert p.launch_histogram()["mcn_sigmoid_xent"] == 3
    # D convolutions are differentiated twice (real + fake) for D's loss, and a third time without
    # weight gradients for G's loss
    tags = [l.tag for l in p.bwd]
    assert sum(t.endswith("discriminator_unit/conv_0/wgrad") for t in tags) == 8


def test_inference_plan_uses_ema_and_moving_statistics(r50):
    p = Plan(r50.graph)
    names = [l.fn for l in p.inf]
    assert names.count("mcn_bn_infer") == 53 and "mcn_bn_stats" not in names
    assert names.count("mcn_conv2d_fprop_tc") == 53 and names.count("mcn_stem_conv_fprop") == 1
    assert "mcn_conv2d_fprop_tc_stats" not in names
    # every weight operand of the inference list comes from the EMA copies
    ema_ptrs = [a for l in p.inf for a in l.args if hasattr(a, "buf") and a.buf in (p.b_ema, p.b_bf16_ema)]
    raw_ptrs = [a for l in p.inf for a in l.args if hasattr(a, "buf") and a.buf in (p.b_param, p.b_bf16)]
    assert ema_ptrs and not raw_ptrs


def test_every_launch_pointer_lies_inside_its_buffer(r50):
    """Structural check of the planned launch lists (train, backward, inference): every pointer
    argument addresses a byte inside the buffer it names, buffers of the arena do not overlap, and
    every fused-statistics conv writes into the sums buffer of the BN that reads it."""
    from myconvnet_b200.plan import Ptr
    for world in (1, 8):
        p = Plan(r50.graph, world_size=world)
        for l in p.fwd + p.bwd + p.inf:
            for a in l.args:
                if isinstance(a, Ptr):
                    assert 0 <= a.off < max(a.buf.nbytes, 1), (l.fn, l.tag, a.buf.name, a.off, a.buf.nbytes)
                    assert a.buf.offset is not None and a.buf.offset % 256 == 0
        spans = sorted((b.offset, b.offset + b.nbytes, b.name) for b in p.bufs if b.nbytes > 0)
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
        sums_of_apply = {}
        for l in p.fwd:
            if l.fn in ("mcn_bn_apply_stats", "mcn_bn_apply_stats_mask"):
                sums_of_apply[l.tag.rsplit("/bn/", 1)[0]] = l.args[4].buf
        fused = [l for l in p.fwd if l.fn in ("mcn_conv2d_fprop_tc_stats", "mcn_stem_conv_fprop")]
        assert len(fused) == 53
        for l in fused:
            sums = l.args[6] if l.fn == "mcn_conv2d_fprop_tc_stats" else l.args[5]
            assert sums.buf is sums_of_apply[l.tag] and sums.buf.region == "zero"
        # synchronised BN: one exchange before every apply and one before every backward apply
        if world > 1:
            for phase, idx, ptr, nbytes, dt, srcs in p.allreduce_points:
                nxt = (p.fwd if phase == "f" else p.bwd)[idx]
                assert nxt.fn in (("mcn_bn_apply_stats", "mcn_bn_apply_stats_mask") if phase == "f"
                                  else ("mcn_bn_bwd_apply", "mcn_bn_bwd_apply_mask")), nxt.fn
                assert any(isinstance(a, Ptr) and a.buf is ptr.buf and a.off == ptr.off for a in nxt.args)


def test_wsgn_model_plans_standardised_weights_and_group_norm():
    """models/resnet_v1_5_wsgn.py (unchanged): every convolution's weight is standardised per step into
    node-owned buffers (the tensor-core routes then run on per-step bf16 copies), the weight gradient
    is folded back onto the raw weight, and every normalisation is a group norm."""
    from myconvnet_b200 import loader
    if loader.reference_root() is None:
        pytest.skip("reference model files not staged")
    from tests.util import build_pair
    pm, _, _ = build_pair("models/resnet_v1_5_wsgn.py", "ResNet50", [64, 64, 3], 16, 8, "bf16")
    p = Plan(pm.graph)
    f, b = {}, {}
    for l in p.fwd:
        f[l.fn] = f.get(l.fn, 0) + 1
    for l in p.bwd:
        b[l.fn] = b.get(l.fn, 0) + 1
    assert f["mcn_ws_fwd"] == 53 and f["mcn_weight_prep"] == 52 and f["mcn_gn_fwd"] == 53
    assert f["mcn_conv2d_fprop_direct"] == 1           # the RGB stem: plain storage layout, CUDA cores
    assert b["mcn_ws_bwd"] == 53 and b["mcn_gn_bwd"] == 53 and "mcn_bn_bwd_reduce" not in b
    # the standardised operand, not the raw bf16 copy, feeds the tensor-core launches
    raw = {id(p.b_bf16)}
    for l in p.fwd:
        if l.fn == "mcn_conv2d_fprop_tc" and "logits" not in l.tag:
            assert id(l.args[2].buf) not in raw, l.tag
    # inference standardises the EMA weights the same way
    assert sum(1 for l in p.inf if l.fn == "mcn_ws_fwd") == 53
