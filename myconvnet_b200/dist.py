"""Data-parallel plumbing: gradient buckets and statistics all-reduce over torch.distributed.

Replaces the reference's parameter-server averaging (optimizers.py:117-147: concat + reduce_mean
on the parameter device) and its tower-after-tower BN moving-statistics chain
(convnet.py:1898-1914).  One process per GPU; NCCL over NVLink on the device, gloo in CPU tests.
The flat gradient buffer mirrors the parameter layout, so a bucket is a slice — no packing copy.
"""
import torch
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


def allreduce_stats(t, group=None):
    """Sum-all-reduce one BN statistics vector ([sum x, sum x^2] or [sum dz, sum dz*xhat])."""
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
