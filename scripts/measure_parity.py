"""Prints the parity error tables of BASELINE config 1 (ResNet-v1.5-50, 224x224x3, batch 32, one
training step) for fp32 and bf16: per-tap activations, loss, all gradients, updated variables,
run-to-run bit identity, max-pool argmax.  The tolerances in tests/test_gpu_resnet.py come from this
output (profiles/r02_parity_config1.txt)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.util import build_pair, rel_l2, synthetic_batch, sync_engine_from_oracle, worst  # noqa: E402


def main():
    from myconvnet_b200.engine import Engine
    from myconvnet_b200.plan import Ptr
    from oracle.step import OracleTrainer
    from oracle import tf_ops
    shape, ncls, batch = [224, 224, 3], 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 32
    for dtype in ("f32", "bf16"):
        t0 = time.time()
        pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", shape, ncls, batch, dtype)
        X, Y = synthetic_batch(batch, shape, ncls)
        taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k != "pred"}
        eng = Engine(pm, keep=list(taps.values()))
        eng.set_variables(vals)
        loss_dev = eng.train_step(X, Y, update=False)
        tr = OracleTrainer(om)
        t1 = time.time()
        tr.step(X, Y, update=False)
        print("[%s] oracle fwd+bwd %.1f s, build+device %.1f s" % (dtype, time.time() - t1, t1 - t0))
        loss_ref = float(om.data_loss.detach())
        print("[%s] loss device %.7f oracle %.7f rel %.2e" % (dtype, loss_dev, loss_ref, abs(loss_dev - loss_ref) / loss_ref))
        aerr = {k: rel_l2(eng.fetch(t), om.d[k].t.detach().numpy()) for k, t in taps.items()}
        print("[%s] activations: %d taps, worst %s" % (dtype, len(aerr), worst(aerr)))
        print("[%s] activations by stage: %s" % (dtype, {b: "%.2e" % max(v for k, v in aerr.items() if k.startswith(b))
                                                       for b in ("block_0", "block_1", "block_2", "block_3", "block_4", "logits")}))
        grads = eng.get_gradients()
        l2 = om._parameters.get("l2_reg", 1e-4)
        gerr = {}
        for k, g in tr.grads.items():
            ref = g.numpy() - (l2 * vals[k] if k.endswith("weights") else 0.0)
            if np.linalg.norm(ref) > 1e-9:
                gerr[k] = rel_l2(grads[k], ref)
        print("[%s] gradients: %d tensors, worst %s, median %.2e" % (dtype, len(gerr), worst(gerr), float(np.median(list(gerr.values())))))
        # max-pool argmax on the device's own pool input
        node = [n for n in pm.graph.nodes if n.op == "max_pool"][0]
        xin = eng.fetch(node.inputs[0])
        am = eng.maxpool_argmax(node)
        t2 = time.time()
        ref_idx = tf_ops.max_pool_argmax(torch.from_numpy(xin[:4]), [3, 3], [2, 2], "SAME").numpy()
        print("[%s] argmax bit-exact on 4 images: %s (%.1f s), ties in input: %d" % (
            dtype, np.array_equal(am[:4].astype(np.int64), ref_idx), time.time() - t2, 0))
        del eng
        # fused plan: one full step from identical state, twice (bit identity), vs the oracle
        eng2 = Engine(pm)
        outs = []
        for rep in range(2):
            eng2.set_variables(vals)
            outs.append((eng2.train_step(X, Y), eng2.get_variables(), eng2.get_gradients()))
        same = outs[0][0] == outs[1][0] and all(np.array_equal(outs[0][1][k], outs[1][1][k]) for k in outs[0][1]) \
            and all(np.array_equal(outs[0][2][k], outs[1][2][k]) for k in outs[0][2])
        print("[%s] fused step run twice from the same state: bit-identical = %s (loss %r)" % (dtype, same, outs[0][0]))
        om.set_variables(vals)
        tr2 = OracleTrainer(om)
        lref = tr2.step(X, Y)
        print("[%s] fused loss device %.7f oracle %.7f" % (dtype, outs[0][0], lref))
        verr = {}
        uerr = {}
        for k, v in outs[0][1].items():
            ref = om.vars[k].detach().numpy()
            verr[k] = rel_l2(v, ref)
            du, dr = v - vals[k], ref - vals[k]
            if np.linalg.norm(dr) > 1e-12:
                uerr[k] = rel_l2(du, dr)
        print("[%s] updated variables worst %s" % (dtype, worst(verr)))
        print("[%s] update deltas (w1-w0) worst %s median %.2e" % (dtype, worst(uerr), float(np.median(list(uerr.values())))))
        gerr2 = {}
        for k, g in tr2.grads.items():
            ref = g.numpy() - (l2 * vals[k] if k.endswith("weights") else 0.0)
            if np.linalg.norm(ref) > 1e-9:
                gerr2[k] = rel_l2(outs[0][2][k], ref)
        print("[%s] fused-plan gradients worst %s median %.2e" % (dtype, worst(gerr2), float(np.median(list(gerr2.values())))))
        del eng2
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
