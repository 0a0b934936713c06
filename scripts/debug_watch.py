"""Diagnostic: run the step launch by launch and report which launches change one gradient slot."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import build_pair, synthetic_batch
from myconvnet_b200.engine import Engine
from myconvnet_b200.plan import Ptr
from myconvnet_b200 import lib as L

name, idx = sys.argv[1], int(sys.argv[2])
SHAPE, NCLS, BATCH = [64, 64, 3], 16, 8
pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, "f32")
X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k not in ("pred",)}
eng = Engine(pm, keep=list(taps.values()))
eng.set_variables(vals)
eng.load_inputs(X=X, Y=Y)
eng._set_hyper(1.0)
v = [v for v in eng.plan.trainable if v.name == name][0]
slot = eng._var_view(eng.plan.b_grad, v)
L.check(eng.lib.mcn_fill_f32(eng._zero_ptr, eng._zero_n, 0.0, None))
last = 0.0
for phase, launches in (("f", eng._fwd), ("b", eng._bwd)):
    for i, (fn, args, fname, tag) in enumerate(launches):
        L.check(fn(*args, None), fname)
        torch.cuda.synchronize()
        cur = float(slot[idx])
        if cur != last:
            print("%s[%d] %s [%s] changed slot: %.8g -> %.8g" % (phase, i, fname, tag, last, cur))
            last = cur

# ---- compare the gradient tensor entering a given backward launch with the oracle's
import torch
from oracle.step import OracleTrainer
from tests.util import rel_l2
tap = sys.argv[3] if len(sys.argv) > 3 else None
if tap:
    loss = om.forward(X, Y)
    (gref,) = torch.autograd.grad(loss, om.d[tap].t)
    gref = gref.numpy().ravel()
    # re-run backward up to the launch that consumes this gradient
    L.check(eng.lib.mcn_fill_f32(eng._zero_ptr, eng._zero_n, 0.0, None))
    for fn, args, fname, tag in eng._fwd:
        L.check(fn(*args, None), fname)
    target = tap.replace("/bn", "/bn/bwd_reduce") if tap.endswith("/bn") else None
    for i, (fn, args, fname, tag) in enumerate(eng._bwd):
        if tag == target:
            addr = args[1]
            off = addr - eng.base + eng._skew
            got = eng.arena[off:off + gref.size * 4].view(torch.float32).cpu().numpy()
            d = np.abs(got - gref)
            bad = np.nonzero(d > 1e-5 * np.abs(gref).max())[0]
            print("grad into", tag, "rel_l2", rel_l2(got, gref), "nbad", bad.size, "first bad", bad[:20])
            for j in bad[:10]:
                print("    idx %d (pixel %d ch %d) got %.6g ref %.6g" % (j, j // 512, j % 512, got[j], gref[j]))
            print("launches before it:", [(k, eng._bwd[k][2], eng._bwd[k][3]) for k in range(max(0, i - 4), i)])
            print("temp offset of gy:", addr - eng.base - eng.plan.temp_buf.offset, "bytes", gref.size * 4)
            break
        L.check(fn(*args, None), fname)
        torch.cuda.synchronize()
