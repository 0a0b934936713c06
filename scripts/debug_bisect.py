"""Diagnostic: replay the backward launch list several times over a fixed forward and report the
first launch after which the (temp + grads) memory image differs between passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import build_pair, synthetic_batch
from myconvnet_b200.engine import Engine
from myconvnet_b200 import lib as L

dtype = sys.argv[1]
SHAPE, NCLS, BATCH = [64, 64, 3], 16, 8
pm, om, vals = build_pair("models/resnet_v1_5.py", "ResNet50", SHAPE, NCLS, BATCH, dtype)
X, Y = synthetic_batch(BATCH, SHAPE, NCLS)
taps = {k: t for k, t in pm.d.items() if hasattr(t, "shape") and k not in ("pred",)}
eng = Engine(pm, keep=list(taps.values()))
eng.set_variables(vals)
eng.load_inputs(X=X, Y=Y)
eng._set_hyper(1.0)
p = eng.plan
t0, t1 = p.region_span["temp"]
z0, z1 = p.region_span["zero"]
def image():
    a = eng.arena[eng._skew + t0: eng._skew + t1].clone()
    b = eng.arena[eng._skew + z0: eng._skew + z1].clone()
    return a, b
snaps = []
for rep in range(4):
    L.check(eng.lib.mcn_fill_f32(eng._zero_ptr, eng._zero_n, 0.0, None))
    eng.arena[eng._skew + t0: eng._skew + t1].zero_()
    for fn, args, fname, tag in eng._fwd:
        L.check(fn(*args, None), fname)
    cur = []
    for i, (fn, args, fname, tag) in enumerate(eng._bwd):
        L.check(fn(*args, None), fname)
        torch.cuda.synchronize()
        cur.append(image())
    snaps.append(cur)

def fdiff(a, b):
    a = a.view(torch.float32); b = b.view(torch.float32)
    d = (a - b).abs()
    scale = torch.maximum(a.abs(), b.abs()).clamp(min=1e-6)
    bad = (d > 1e-3 * scale) & (d > 1e-7)
    bad &= torch.isfinite(a) & torch.isfinite(b)
    return bad
for (r0, r1) in ((1, 2), (2, 3)):
    for i in range(len(eng._bwd)):
        da = fdiff(snaps[r0][i][0], snaps[r1][i][0])
        g0, g1 = snaps[r0][i][1], snaps[r1][i][1]
        ng = eng.plan.n_train * 4
        db = fdiff(g0[:ng], g1[:ng])
        if da.any() or db.any():
            fn, args, fname, tag = eng._bwd[i]
            ia = torch.nonzero(da).flatten()[:8].tolist()
            ib = torch.nonzero(db).flatten()[:8].tolist()
            print("pass %d vs %d: first difference after b[%d] %s [%s]; temp float idx %s (n=%d) grad float idx %s (n=%d)"
                  % (r0, r1, i, fname, tag, ia, int(da.sum()), ib, int(db.sum())))
            print("  ptr args rel. to temp base:", [a - (eng.base + t0) for a in args if isinstance(a, int) and a > 1 << 30])
            a = snaps[r0][i][0].view(torch.float32); b = snaps[r1][i][0].view(torch.float32)
            for j in ia[:4]:
                print("   temp[%d] = %.6g vs %.6g" % (j, a[j].item(), b[j].item()))
            print("  previous launches:", [(k, eng._bwd[k][2], eng._bwd[k][3]) for k in range(max(0, i - 3), i + 1)])
            break
    else:
        print("pass %d vs %d identical within tolerance" % (r0, r1))
