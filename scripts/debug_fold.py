"""A/B of the folded BN finalize: same model, same weights, two engines; compares gradients and
post-step variables after one step, then the loss after the second."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import build_pair, rel_l2, synthetic_batch
from myconvnet_b200.engine import Engine

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", [64, 64, 3], 16, 8, dtype, base_learning_rate=0.05)
X, Y = synthetic_batch(8, [64, 64, 3], 16)
out = {}
for fold in ("0", "1"):
    os.environ["MCN_BN_FOLD_FINALIZE"] = fold
    eng = Engine(pm)
    eng.set_variables(vals)
    l1 = eng.train_step(X, Y)
    g = eng.get_gradients()
    v = eng.get_variables()
    l2 = eng.train_step(X, Y)
    out[fold] = (l1, l2, g, v)
    print("fold", fold, "losses", l1, l2)
g0, g1 = out["0"][2], out["1"][2]
v0, v1 = out["0"][3], out["1"][3]
bad = sorted(((rel_l2(g1[k], g0[k]), k) for k in g0), reverse=True)[:12]
print("largest gradient differences:", bad)
badv = sorted(((rel_l2(v1[k], v0[k]), k) for k in v0), reverse=True)[:12]
print("largest variable differences after step 1:", badv)
