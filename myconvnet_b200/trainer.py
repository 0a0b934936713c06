"""Step driver: the reference's training loop shape on the B200 engine.

Mirrors Optimizer.train (reference optimizers.py:403-422): every step first derives the
learning-rate multiplier from (step, epoch) — warm-up then step / exponential / polynomial /
cosine-with-restarts decay, optimizers.py:608-632 — and runs one optimisation step with it.  The
multiplier is a device-side hyper-parameter, so the captured CUDA graph is replayed unchanged.
Batches come from memory (the reference's DataSet(from_memory=True) path, dataset.py:46,134):
contiguous shards of a shuffled index per rank, the last partial batch of an epoch is skipped as
the reference does (optimizers.py:411-414)."""
import math

import numpy as np


class LearningRateSchedule(object):
    """multiplier(step) as a pure function of the global step (0-based)."""

    def __init__(self, steps_per_epoch, num_epochs, warmup_epochs=1.0, method=None, params=(0.94, 2)):
        self.steps_per_epoch = int(steps_per_epoch)
        self.num_epochs = int(num_epochs)
        self.warmup_steps = float(np.around(warmup_epochs * self.steps_per_epoch))
        self.method = None if method is None else method.lower()
        self.params = params
        if self.method not in (None, "step", "exponential", "poly", "polynomial", "cosine"):
            self.method = "cosine"        # the reference treats every other name as cosine (optimizers.py:627)

    def __call__(self, step):
        w = self.warmup_steps
        if step < w:
            return (step + 1) / w
        if self.method is None:
            return 1.0                    # the multiplier keeps its last warm-up value, which is 1
        p = self.params
        if self.method == "step":
            epoch = 1 + step // self.steps_per_epoch          # epochs count from 1 (optimizers.py:67)
            passed = sum(1 for boundary in p[1:] if epoch > boundary)
            return float(p[0]) ** passed
        if self.method == "exponential":
            return float(p[0]) ** ((step - w) / self.steps_per_epoch / p[1])
        total = self.steps_per_epoch * self.num_epochs - w
        first = p[0] if isinstance(p, (list, tuple)) else p
        if self.method in ("poly", "polynomial"):
            return (1.0 - (step - w) / total) ** first
        restarts = 0 if first is None else int(first)
        progress = math.fmod((restarts + 1) * (step - w) / total, 1.0)
        return 0.5 * (1.0 + math.cos(progress * math.pi))


class Trainer(object):
    """Runs epochs of engine.train_step over an in-memory data set."""

    def __init__(self, engine, train_size, num_epochs=None, seed=0, shuffle=True, **kwargs):
        kw = dict(getattr(engine, "kw", {}))
        kw.update(kwargs)
        self.engine = engine
        self.batch = int(engine.batch)
        self.world = int(getattr(engine, "world", 1))
        self.rank = int(getattr(engine, "rank", 0))
        self.global_batch = self.batch * self.world
        self.train_size = int(train_size)
        self.num_epochs = int(num_epochs if num_epochs is not None else kw.get("num_epochs", 100))
        self.steps_per_epoch = int(math.ceil(self.train_size / self.global_batch))
        self.schedule = LearningRateSchedule(
            self.steps_per_epoch, self.num_epochs,
            kw.get("learning_warmup_epochs", kw.get("learning_warmup_epoch", 1.0)),
            kw.get("learning_rate_decay_method", None), kw.get("learning_rate_decay_params", (0.94, 2)))
        self.shuffle = shuffle
        self.rng = np.random.default_rng(seed)          # same seed on every rank: same permutation
        self.curr_step = 0
        self.history = []

    @property
    def curr_epoch(self):
        return 1 + self.curr_step // self.steps_per_epoch

    def _epoch_batches(self):
        order = self.rng.permutation(self.train_size) if self.shuffle else np.arange(self.train_size)
        for s in range(self.steps_per_epoch):
            idx = order[s * self.global_batch:(s + 1) * self.global_batch]
            if len(idx) < self.global_batch:
                yield None                 # the reference ignores the last partial batch
            else:
                yield idx[self.rank * self.batch:(self.rank + 1) * self.batch]

    def fit(self, X, Y, num_steps=None, fetch_loss=True, callback=None):
        """X: [train_size, H, W, C] float32 in [0,1]; Y: [train_size] labels.  Returns the losses."""
        total = self.steps_per_epoch * self.num_epochs
        limit = total if num_steps is None else min(total, self.curr_step + int(num_steps))
        losses = []
        while self.curr_step < limit:
            skip = self.curr_step % self.steps_per_epoch       # resuming inside an epoch
            for k, idx in enumerate(self._epoch_batches()):
                if k < skip:
                    continue
                if self.curr_step >= limit:
                    break
                mult = self.schedule(self.curr_step)
                if idx is not None:
                    loss = self.engine.train_step(np.ascontiguousarray(X[idx]), np.ascontiguousarray(Y[idx]),
                                                  lr_multiplier=mult, fetch_loss=fetch_loss)
                    losses.append(loss)
                    self.history.append((self.curr_step, self.curr_epoch, mult, loss))
                    if callback is not None:
                        callback(self)
                self.curr_step += 1
        return losses

    def fit_dataset(self, dataset, num_steps=None, fetch_loss=True):
        """Same loop over a DataSet(from_memory=True) (dataset.py): the data set's own shard order
        and labels (NaN = fake label), this trainer's learning-rate schedule."""
        total = self.steps_per_epoch * self.num_epochs
        limit = total if num_steps is None else min(total, self.curr_step + int(num_steps))
        losses = []
        while self.curr_step < limit:
            stepped = False
            for Xb, Yb in dataset.shard_batches(self.rank):
                if self.curr_step >= limit:
                    break
                mult = self.schedule(self.curr_step)
                loss = self.engine.train_step(np.ascontiguousarray(Xb), np.ascontiguousarray(Yb),
                                              lr_multiplier=mult, fetch_loss=fetch_loss)
                losses.append(loss)
                self.history.append((self.curr_step, self.curr_epoch, mult, loss))
                self.curr_step += 1
                stepped = True
            if not stepped:
                break
        return losses
