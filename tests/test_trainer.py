"""Step driver: the product's closed-form learning-rate schedule against the
oracle's restatement of the reference loop (optimizers.py:608-632), and the epoch / shard
bookkeeping of Trainer with a recording stand-in for the engine (no GPU)."""
import numpy as np
import pytest

from myconvnet_b200.trainer import LearningRateSchedule, Trainer
from oracle import schedule as ref


CASES = [
    dict(method=None, params=(0.94, 2), warmup=1.0),
    dict(method="step", params=(0.1, 3, 5, 7), warmup=0.5),
    dict(method="exponential", params=(0.94, 2), warmup=1.0),
    dict(method="poly", params=(2.0,), warmup=0.0),
    dict(method="polynomial", params=0.9, warmup=1.5),
    dict(method="cosine", params=(2,), warmup=1.0),
    dict(method="cosine", params=None, warmup=0.25),
    dict(method="anything-else-is-cosine", params=(0,), warmup=1.0),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("train_size,batch", [(1000, 64), (257, 32), (64, 64)])
def test_learning_rate_schedule_matches_reference_loop(case, train_size, batch):
    epochs = 9
    want = ref.multipliers(train_size, batch, epochs, case["warmup"], case["method"], case["params"])
    spe = int(np.ceil(train_size / batch))
    sch = LearningRateSchedule(spe, epochs, case["warmup"], case["method"], case["params"])
    got = [sch(s) for s in range(spe * epochs)]
    assert len(got) == len(want)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)


class _RecordingEngine(object):
    def __init__(self, batch, world=1, rank=0):
        self.batch, self.world, self.rank, self.kw = batch, world, rank, {}
        self.calls = []

    def train_step(self, X, Y, lr_multiplier=1.0, fetch_loss=True):
        self.calls.append((X[:, 0, 0, 0].copy(), Y.copy(), lr_multiplier))
        return float(len(self.calls))


def test_trainer_shards_epochs_and_skips_the_partial_batch():
    n, b, world = 100, 8, 2                       # global batch 16 -> 7 steps/epoch, the 7th is partial
    X = np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1) * np.ones((1, 2, 2, 3), np.float32)
    Y = np.arange(n, dtype=np.int32)
    runs = []
    for rank in range(world):
        eng = _RecordingEngine(b, world, rank)
        tr = Trainer(eng, n, num_epochs=3, seed=5, learning_warmup_epochs=1.0,
                     learning_rate_decay_method="cosine", learning_rate_decay_params=(0,))
        losses = tr.fit(X, Y)
        assert tr.steps_per_epoch == 7 and tr.curr_step == 21 and tr.curr_epoch == 4
        assert len(losses) == 18                  # 6 full batches per epoch
        runs.append(eng.calls)
    want = ref.multipliers(n, b * world, 3, 1.0, "cosine", (0,))
    used = [m for s, m in enumerate(want) if s % 7 != 6]
    for calls in runs:
        np.testing.assert_allclose([c[2] for c in calls], used, rtol=1e-12)
        assert all(np.array_equal(c[0].astype(np.int32), c[1]) for c in calls)      # images follow labels
    for a, c in zip(*runs):
        assert not set(a[1]) & set(c[1])          # ranks see disjoint shards of the same global batch
    for e in range(3):                            # every epoch covers 96 distinct samples
        seen = np.concatenate([np.concatenate([runs[r][e * 6 + s][1] for r in range(world)]) for s in range(6)])
        assert len(set(seen.tolist())) == 96


def test_trainer_resumes_from_a_step_budget():
    eng = _RecordingEngine(4)
    tr = Trainer(eng, 40, num_epochs=2, shuffle=False, learning_warmup_epochs=0.0)
    X = np.zeros((40, 2, 2, 3), np.float32)
    Y = np.arange(40, dtype=np.int32)
    assert len(tr.fit(X, Y, num_steps=7)) == 7 and tr.curr_step == 7
    assert [int(c[1][0]) for c in eng.calls] == [0, 4, 8, 12, 16, 20, 24]
    assert len(tr.fit(X, Y, num_steps=5)) == 5 and tr.curr_step == 12          # continues inside epoch 1
    assert [int(c[1][0]) for c in eng.calls[7:]] == [28, 32, 36, 0, 4]
