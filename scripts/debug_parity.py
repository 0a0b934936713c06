"""Diagnostic: per-layer activation / gradient errors of the device engine vs the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import build_pair, rel_l2, synthetic_batch
from myconvnet_b200.engine import Engine
from oracle.step import OracleTrainer

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
fused = len(sys.argv) > 2 and sys.argv[2] == "fused"
SHAPE, NCLS, BATCH = [64, 64, 3], 16, 8
pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, dtype)
X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k not in ("pred",)}
eng = Engine(pm, keep=() if fused else list(taps.values()))
eng.set_variables(vals)
loss_dev = eng.train_step(X, Y, update=False)
tr = OracleTrainer(om)
tr.step(X, Y, update=False)
print("loss dev %.6f ref %.6f" % (loss_dev, float(om.data_loss)))
if not fused:
    errs = [(rel_l2(eng.fetch(t), om.d[k].t.detach().numpy()), k) for k, t in taps.items()]
    print("activations in graph order:")
    for e, k in errs:
        print("   %.3e %s" % (e, k))
grads = eng.get_gradients()
gerrs = []
for k, g in tr.grads.items():
    ref = g.numpy()
    if k.endswith("weights"):
        ref = ref - 1e-4 * vals[k]
    gerrs.append((rel_l2(grads[k], ref), k, float(np.linalg.norm(ref)), float(np.linalg.norm(grads[k]))))
print("gradients in variable order (err, name, |ref|, |dev|):")
for e in gerrs:
    print("   %.4g %s %.4g %.4g" % e)
print("median grad err %.4g" % np.median([e[0] for e in gerrs]))
for name in sys.argv[3:]:
    ref = tr.grads[name].numpy().ravel()
    got = grads[name].ravel()
    if name.endswith("weights"):
        ref = ref - 1e-4 * vals[name].ravel()
    d = np.abs(got - ref)
    idx = np.argsort(-d)[:12]
    print(name, "n", d.size, "nbad(>1e-4*max)", int((d > 1e-4 * np.abs(ref).max()).sum()))
    for i in idx:
        print("   idx %d got %.6g ref %.6g" % (i, got[i], ref[i]))
