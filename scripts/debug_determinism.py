"""Diagnostic: run the same step repeatedly at a given shape and report the first activations /
gradients / tensor gradients that change between runs (bitwise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import build_pair, synthetic_batch
from myconvnet_b200.engine import Engine

dtype = sys.argv[1]
hw, batch, ncls = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
runs_n = int(sys.argv[5]) if len(sys.argv) > 5 else 4
SHAPE = [hw, hw, 3]
pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, ncls, batch, dtype)
X, Y = synthetic_batch(batch, SHAPE, ncls)
eng = Engine(pm, keep_grads=True)
eng.set_variables(vals)
acts = [t for n in pm.graph.nodes for t in n.outputs if t in eng.plan.tbuf]
runs = []
for r in range(runs_n):
    loss = eng.train_step(X, Y, update=False)
    a = {("act", t.node.scope, t.node.op, t.id if hasattr(t, "id") else 0): eng.fetch(t) for t in acts}
    tg = {}
    for t in acts:
        g = eng.fetch_grad(t)
        if g is not None:
            tg[("tgrad", t.node.scope, t.node.op, 0)] = g
    pg = {("pgrad", k, "", 0): v for k, v in eng.get_gradients().items()}
    runs.append((loss, a, tg, pg))
print("losses", [r[0] for r in runs])
for idx, label in ((1, "activations"), (2, "tensor gradients (backward order = reverse)"), (3, "parameter gradients")):
    print("==", label)
    keys = list(runs[0][idx].keys())
    if idx == 2:
        keys = keys[::-1]
    shown = 0
    for k in keys:
        base = runs[0][idx][k]
        nd = max(int((r[idx][k] != base).sum()) for r in runs[1:])
        if nd:
            print("  %-70s differing %d / %d" % (k[1] + " " + k[2], nd, base.size))
            shown += 1
            if shown >= 12:
                break
print("done")
