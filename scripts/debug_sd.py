"""Where does the bf16 stochastic-depth step leave the oracle?  Per-tap activation errors."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests.util import build_pair, rel_l2, synthetic_batch, worst  # noqa: E402
from tests.test_gpu_resnet import relu_pattern  # noqa: E402
from myconvnet_b200.engine import Engine  # noqa: E402
from oracle.step import OracleTrainer  # noqa: E402

for kw in (dict(initial_drop_rate=0.1, final_drop_rate=0.4, dropout_rate=0.3), dict(initial_drop_rate=0.1, final_drop_rate=0.4),
           dict(dropout_rate=0.3), dict()):
    pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", [128, 128, 3], 16, 16, "bf16", base_learning_rate=0.05, **kw)
    X, Y = synthetic_batch(16, [128, 128, 3], 16)
    taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k != "pred"}
    eng = Engine(pm, keep=list(taps.values()))
    eng.set_variables(vals)
    eng.train_step(X, Y, update=False)
    om.forced_relu_masks = relu_pattern(eng, pm)
    OracleTrainer(om).step(X, Y, update=False)
    aerr = {k: rel_l2(eng.fetch(t), om.d[k].t.detach().numpy()) for k, t in taps.items()}
    print(kw)
    for k in list(taps)[:6] + [k for k in taps if k.startswith("block_1/res_0") or k.startswith("block_4/res_2") or k.startswith("logits")]:
        print("   %-34s %.3e" % (k, aerr[k]))
    print("   worst", worst(aerr, 3))
