"""ORACLE (test infrastructure): one training step with the reference's update semantics.

Restates SURVEY.md 3.2 / reference optimizers.py:89-177 for one tower on CPU:
  loss -> gradients -> [EMA shadows from PRE-step values with num_updates = global_step before
  the increment; BN moving statistics assigned] -> apply_gradients -> decoupled weight decay.
lr = base_lr * batch/256 * multiplier (optimizers.py:46,57).  Data-parallel replicas with
synchronised BN are numerically one tower at the global batch, which is what this computes.
"""
import torch

from . import tf_ops as ops


class OracleTrainer(object):
    def __init__(self, model, optimizer="nesterov", batch_size=None, **kwargs):
        self.model = model
        self.kind = optimizer.lower()
        kw = dict(model._parameters)
        kw.update(kwargs)
        self.kw = kw
        self.batch_size = batch_size
        self.base_lr = float(kw.get("base_learning_rate", 0.1))
        self.momentum = float(kw.get("momentum", 0.9))
        self.base_wd = float(kw.get("base_weight_decay", 0.0))
        self.wd_sched = bool(kw.get("weight_decay_scheduling", True))
        self.bias_norm_decay = bool(kw.get("bias_norm_decay", False))
        self.global_step = 0
        self.state = {}
        self.ema = None
        self.grads = {}

    def step(self, X, Y, lr_multiplier=1.0, update=True):
        m = self.model
        batch = self.batch_size or len(X)
        m.random_step = self.global_step        # same (seed, step) as the device's random ops
        loss = m.forward(X, Y)
        if self.ema is None:
            self.ema = {k: v.detach().clone() for k, v in m.vars.items()}
        if isinstance(loss, tuple):
            # GAN: d(loss_d)/d(theta_D) and d(gsf*loss_g)/d(theta_G) from the same forward
            # (optimizers_gan.py:56-58)
            loss_d, loss_g = loss
            vd, vg = m.variable_split()
            td = [k for k in vd if m.var_meta[k]["trainable"] and m.var_meta[k]["kind"] != "stat"]
            tg = [k for k in vg if m.var_meta[k]["trainable"] and m.var_meta[k]["kind"] != "stat"]
            gsf = float(self.kw.get("generator_scaling_factor", 1.0))
            gd = torch.autograd.grad(loss_d, [m.vars[k] for k in td], allow_unused=True, retain_graph=True)
            gg = torch.autograd.grad(loss_g * gsf, [m.vars[k] for k in tg], allow_unused=True)
            train = td + tg
            grads = list(gd) + list(gg)
            self.last_losses = (float(loss_d.detach()), float(loss_g.detach()))
            loss = loss_d
        else:
            train = [k for k, meta in m.var_meta.items() if meta["trainable"] and meta["kind"] != "stat"]
            grads = torch.autograd.grad(loss, [m.vars[k] for k in train], allow_unused=True)
        self.grads = {k: (g if g is not None else torch.zeros_like(m.vars[k])) for k, g in zip(train, grads)}
        thr = self.kw.get("gradient_threshold", None)
        if thr is not None:
            # tf.clip_by_global_norm (optimizers.py:112-113): g * clip / max(global_norm, clip)
            gn = torch.sqrt(sum((g.double() ** 2).sum() for g in self.grads.values()))
            self.grad_norm = float(gn)
            scale = float(thr) / max(float(gn), float(thr))
            self.grads = {k: g * scale for k, g in self.grads.items()}
        if not update:
            return float(loss.detach())
        lr = self.base_lr * batch / 256.0 * lr_multiplier
        d_t = ops.ema_decay(m.moving_average_decay, self.global_step)
        with torch.no_grad():
            # BN moving statistics are assigned, then the EMA shadows are updated: shadows of the
            # trainable variables see their PRE-step values (control deps optimizers.py:159,175);
            # the shadow of a moving statistic and the statistic's own assign are two unordered
            # update ops in the reference (convnet.py:1869-1914 — a race, like SURVEY Appendix D.9),
            # resolved here as "assign first": the shadow tracks the value inference would read.
            for k, v in m.bn_updates.items():
                m.vars[k] = v.clone()
            for k, v in m.vars.items():
                self.ema[k] = self.ema[k] - (1.0 - d_t) * (self.ema[k] - v)
            t = self.global_step + 1
            for k in train:
                w, g = m.vars[k].detach(), self.grads[k]
                st = self.state.setdefault(k, {})
                if self.kind in ("nesterov", "momentum", "sgd"):
                    w, st["a"] = ops.nesterov_update(w, g, st.get("a", torch.zeros_like(w)), lr, self.momentum)
                elif self.kind == "rmsprop":
                    w, st["ms"], st["mom"] = ops.rmsprop_update(
                        w, g, st.get("ms", torch.ones_like(w)), st.get("mom", torch.zeros_like(w)), lr,
                        0.9, self.momentum, 1e-3)
                elif self.kind == "adam":
                    w, st["m"], st["v"] = ops.adam_update(
                        w, g, st.get("m", torch.zeros_like(w)), st.get("v", torch.zeros_like(w)), lr, t,
                        self.momentum, 0.999, 1e-3)
                else:
                    raise ValueError(self.kind)
                meta = m.var_meta[k]
                decayed = meta["kind"] == "weight" or (self.bias_norm_decay and meta["kind"] in ("bias", "norm"))
                if self.base_wd > 0 and decayed:
                    # decoupled weight decay and its l1 / pseudo-Huber forms (optimizers.py:163-172)
                    wd = self.base_wd * batch / 256.0 * (lr_multiplier if self.wd_sched else 1.0)
                    delta = self.kw.get("huber_decay_delta", None)
                    if delta is not None:
                        w = w - wd * w / torch.sqrt(1 + (w / delta) ** 2)
                    elif self.kw.get("l1_weight_decay", False):
                        w = w - wd * torch.sign(w)
                    else:
                        w = w - wd * w
                m.vars[k] = w.clone()
        self.global_step += 1
        return float(loss.detach())

    def predict(self, X, Y=None):
        """reference ConvNet.predict: is_train=False -> EMA shadows of all variables, BN inference."""
        m = self.model
        saved = dict(m.vars)
        ema = self.ema if self.ema is not None else saved
        try:
            m.vars = {k: ema[k].detach().clone() for k in saved}
            m.is_train = False
            import numpy as np
            with torch.no_grad():
                m.forward(X, Y if Y is not None else np.zeros(len(X), dtype=np.int64))
            return m.pred.detach().numpy()
        finally:
            m.vars = saved
            m.is_train = True
