"""Data-parallel plumbing: gradient buckets and statistics all-reduce over torch.distributed.

Replaces the reference's parameter-server averaging (optimizers.py:117-147: concat + reduce_mean
on the parameter device) and its tower-after-tower BN moving-statistics chain
(convnet.py:1898-1914).  One process per GPU; NCCL over NVLink on the device, gloo in CPU tests.
The flat gradient buffer mirrors the parameter layout, so a bucket is a slice — no packing copy.
"""
import torch.distributed as dist


def bucket_ranges(n_elems, bucket_elems, align=64):
    """Split [0, n_elems) into contiguous buckets of about bucket_elems (aligned starts)."""
    bucket_elems = max(align, (int(bucket_elems) + align - 1) // align * align)
    out = []
    s = 0
    while s < n_elems:
        e = min(s + bucket_elems, n_elems)
        out.append((s, e))
        s = e
    return out


def allreduce_sum_flat(flat, bucket_elems, group=None, reverse=True):
    """Sum-all-reduce a flat tensor bucket by bucket.  reverse=True starts with the END of the
    buffer: gradients of the last layers are produced first by the backward pass, so their
    buckets can be launched first when overlapped."""
    ranges = bucket_ranges(flat.numel(), bucket_elems)
    if reverse:
        ranges = ranges[::-1]
    works = [dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM, group=group, async_op=True) for s, e in ranges]
    for w in works:
        w.wait()
    return len(ranges)


class BucketOverlap(object):
    """Starts each gradient bucket's all-reduce as soon as the backward pass has produced it.

    schedule: [(start, end, ready_idx)] from Plan.grad_bucket_schedule — ready_idx is the index of
    the last backward launch that writes into the slice (-1: none).  The engine calls
    after_launch(i) behind every backward launch and finish() before the optimiser; with
    overlap=False everything is exchanged in finish() (last buckets first).  Replaces the
    reference's gather-all-towers-then-average (optimizers.py:117-147)."""

    def __init__(self, flat, schedule, group=None, overlap=True):
        self.flat = flat
        self.group = group
        self.ready = {}
        self.tail = []
        for s0, e0, r in schedule:
            if overlap and r >= 0:
                self.ready.setdefault(int(r), []).append((s0, e0))
            else:
                self.tail.append((s0, e0))
        self.works = []

    def _start(self, s0, e0):
        self.works.append(dist.all_reduce(self.flat[s0:e0], op=dist.ReduceOp.SUM, group=self.group,
                                          async_op=True))

    def after_launch(self, i):
        for s0, e0 in self.ready.get(i, ()):
            self._start(s0, e0)

    def finish(self):
        for s0, e0 in self.tail[::-1]:
            self._start(s0, e0)
        for w in self.works:
            w.wait()
        n = len(self.works)
        self.works = []
        return n

    @property
    def n_overlapped(self):
        return sum(len(v) for v in self.ready.values())


def allreduce_stats(t, group=None):
    """Sum-all-reduce one BN statistics vector ([sum x, sum x^2] or [sum dz, sum dz*xhat])."""
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
