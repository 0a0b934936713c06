"""Diagnostic: run the same step repeatedly and report gradient entries that change between runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import build_pair, synthetic_batch
from myconvnet_b200.engine import Engine

dtype = sys.argv[1]
keep = sys.argv[2] == "keep"
SHAPE, NCLS, BATCH = [64, 64, 3], 16, 8
pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, dtype)
X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k not in ("pred",)}
eng = Engine(pm, keep=list(taps.values()) if keep else ())
eng.set_variables(vals)
runs = []
for r in range(6):
    eng.train_step(X, Y, update=False)
    runs.append(eng.get_gradients())
names = list(runs[0].keys())
for k in names:
    base = runs[0][k].ravel()
    worst = 0.0
    nvar = 0
    for r in runs[1:]:
        d = np.abs(r[k].ravel() - base)
        worst = max(worst, float(d.max() / (np.abs(base).max() + 1e-20)))
        nvar = max(nvar, int((d > 1e-4 * np.abs(base).max()).sum()))
    if worst > 1e-4:
        print("%-50s max rel change %.3g  entries varying %d / %d" % (k, worst, nvar, base.size))
print("done")
