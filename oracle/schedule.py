"""ORACLE (test infrastructure): the reference's learning-rate multiplier, step by step.

Restates Optimizer._update_learning_rate (reference optimizers.py:608-632) with its loop state
(curr_step, curr_epoch, curr_multiplier; _reset at optimizers.py:65-70, epoch increment at :545,
steps_per_epoch = ceil(train_size / batch_size) at :182).  Kept as the stateful loop the reference
runs so the product's closed-form schedule (myconvnet_b200/trainer.py) is checked against an
independent formulation."""
import numpy as np


def multipliers(train_size, batch_size, num_epochs, warmup_epochs=1.0, method=None, params=(0.94, 2)):
    """List of the multiplier used by every step of a full run."""
    steps_per_epoch = int(np.ceil(train_size / batch_size))
    num_steps = steps_per_epoch * num_epochs
    curr_step, curr_epoch, mult = 0, 1, 1.0
    out = []
    for i in range(num_steps):
        warmup_steps = np.around(warmup_epochs * steps_per_epoch)
        if curr_step < warmup_steps:
            mult = (curr_step + 1) / warmup_steps
        elif method is not None:
            m = method.lower()
            if m == "step":
                mult = 1.0
                for n in range(len(params) - 1):
                    mult *= np.power(params[0], np.maximum(np.sign(curr_epoch - params[n + 1]), 0.0))
            elif m == "exponential":
                mult = params[0] ** ((curr_step - warmup_steps) / steps_per_epoch / params[1])
            elif m in ("poly", "polynomial"):
                power = params[0] if isinstance(params, (list, tuple)) else params
                total = steps_per_epoch * num_epochs - warmup_steps
                mult = (1 - (curr_step - warmup_steps) / total) ** power
            else:
                anneal = params[0] if isinstance(params, (list, tuple)) else params
                anneal = 0 if anneal is None else int(anneal)
                total = steps_per_epoch * num_epochs - warmup_steps
                prog = ((anneal + 1) * (curr_step - warmup_steps) / total) % 1.0
                mult = 0.5 * (1 + np.cos(prog * np.pi))
        out.append(float(mult))
        curr_step += 1
        if (i + 1) % steps_per_epoch == 0:
            curr_epoch += 1
    return out
