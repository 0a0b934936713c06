import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from myconvnet_b200 import loader
from myconvnet_b200.engine import Engine, draw_initial_value

mod = loader.load_reference_model("models/deeplabv3plus.py", loader.product_facade())
shape, ncls, batch = [64, 64, 3], 5, 2
pm = mod.DeepLabV3PlusResNet(shape, ncls, batch_size=batch, compute_dtype="f32", base_learning_rate=0.05)
rng = np.random.default_rng(0)
vals = {v.name: draw_initial_value(v, rng) for v in pm.graph.vars.values()}
for k in vals:
    if k.endswith("gamma") and not vals[k].any():
        vals[k] = np.full_like(vals[k], 0.5)
rng = np.random.default_rng(2)
X = rng.uniform(size=[batch] + shape).astype(np.float32)
Y = rng.integers(0, ncls + 1, size=[batch] + shape[:2]).astype(np.int32)
eng = Engine(pm, keep_grads=True)
eng.set_variables(vals)
eng.train_step(X, Y, update=False)
g = eng.get_gradients()
for node in pm.graph.nodes:
    if node.op == "bn" and "aspp" in node.scope:
        final = node.attrs.get("final", node.outputs[0])
        gy = eng.fetch_grad(final)
        x = eng.fetch(node.inputs[0])
        print(node.scope, "act", node.attrs["act"], "res", node.attrs["residual"] is not None, "final is out", final is node.outputs[0],
              "consumers", [c.op for c in final.consumers], "x", x.shape)
        if gy is None:
            print("  no grad"); continue
        sg = gy.reshape(-1, gy.shape[-1]).sum(0)
        print("  sum gy       ", sg[:4])
        print("  device dbeta ", g[node.vars["beta"].name][:4])
        xh = (x - x.reshape(-1, x.shape[-1]).mean(0)) / np.sqrt(x.reshape(-1, x.shape[-1]).var(0) + node.attrs["eps"])
        print("  sum gy*xhat  ", (gy * xh).reshape(-1, gy.shape[-1]).sum(0)[:4])
        print("  device dgamma", g[node.vars["gamma"].name][:4])
