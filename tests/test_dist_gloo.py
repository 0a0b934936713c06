"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucketed gradient averaging
equals the reference's concat+reduce_mean over towers (optimizers.py:137-138), and summed per-rank
BN statistics reproduce the single-tower statistics at the global batch (synchronised BN)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from myconvnet_b200.dist import allreduce_stats, allreduce_sum_flat, bucket_ranges
    from oracle import tf_ops
    rng = np.random.default_rng(100 + rank)
    n = 10007
    g = torch.tensor(rng.standard_normal(n).astype(np.float32))
    mine = g.clone()
    nb = allreduce_sum_flat(g, 1024)
    g /= world
    # BN statistics of this rank's shard of a global batch
    full = np.random.default_rng(7).standard_normal((8, 3, 3, 5)).astype(np.float64)
    shard = torch.tensor(full[rank * 4:(rank + 1) * 4])
    st = torch.cat([shard.reshape(-1, 5).sum(0), (shard.reshape(-1, 5) ** 2).sum(0)])
    allreduce_stats(st)
    cnt = full.size // 5
    mean = st[:5] / cnt
    var = st[5:] / cnt - mean ** 2
    _, m_ref, v_ref = tf_ops.fused_batch_norm_train(torch.tensor(full), None, None, 1e-3)
    # gradient buckets started DURING a (simulated) backward pass: slices become final in reverse
    # order, each bucket's all-reduce starts right after the launch that completes it
    from myconvnet_b200.dist import BucketOverlap
    flat = torch.zeros(5000)
    sched = [(0, 2048, 9), (2048, 4096, 5), (4096, 5000, 1)]          # (start, end, ready launch)
    writes = {1: (4096, 5000), 3: (3000, 4096), 5: (2048, 3000), 7: (1000, 2048), 9: (0, 1000)}
    started = []
    for overlap in (True, False):
        flat.zero_()
        bo = BucketOverlap(flat, sched, overlap=overlap)
        for launch in range(10):
            if launch in writes:
                a, b = writes[launch]
                flat[a:b] = torch.arange(a, b, dtype=torch.float32) * (rank + 1)
            before = len(bo.works)
            bo.after_launch(launch)
            started.append((overlap, launch, len(bo.works) - before))
        nfin = bo.finish()
        expect = torch.arange(5000, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(flat, expect), "bucket overlap (overlap=%s) wrong sum" % overlap
        assert nfin == 3 and bo.n_overlapped == (3 if overlap else 0)
    assert [x for x in started if x[0] and x[2]] == [(True, 1, 1), (True, 5, 1), (True, 9, 1)]
    # the ENGINE's own multi-GPU glue (Engine._setup_grad_buckets / _run / _allreduce_grads) driven
    # over a real plan with no-op launches: every trainable gradient ends up summed over the ranks
    from myconvnet_b200.engine import Engine
    from myconvnet_b200.plan import Plan
    from myconvnet_b200.zoo import ResNet50
    pm = ResNet50([32, 32, 3], 10, batch_size=2, compute_dtype="bf16")

    class FakeEngine(object):
        pass
    fe = FakeEngine()
    fe.plan = Plan(pm.graph, world_size=world)
    fe.kw, fe.pg, fe.world, fe.rank = {"bucket_elems": 1 << 20}, None, world, rank
    flat_store = torch.full((fe.plan.n_train,), float(rank + 1))
    fe.view = lambda ptr, n, dt: flat_store[:n]
    fe._ar, fe._peer = {"f": {}, "b": {}}, None
    Engine._setup_grad_buckets(fe)
    assert len(fe._bucket_ready) >= 3 and not fe._bucket_tail
    launches = [((lambda *a: 0), (), l.fn, l.tag) for l in fe.plan.bwd]
    Engine._run(fe, launches, "b", None)
    assert len(fe._buckets.works) == sum(len(v) for v in fe._bucket_ready.values())
    Engine._allreduce_grads(fe)
    assert torch.equal(flat_store, torch.full_like(flat_store, float(sum(r + 1 for r in range(world)))))
    # synchronised-BN exchange points through the same _run loop, NCCL/gloo fallback branch (no peer
    # mapping on CPU): forward sums in place, backward gathers its two source segments first
    store = {}

    def view(ptr, count, dt):
        key = (id(ptr.buf), ptr.off, count, dt)
        if key not in store:
            store[key] = torch.full((count,), float(rank + 1), dtype=dt)
        return store[key]
    fe.view = view
    fe._ar = {"f": {}, "b": {}}
    for k, (phase, idx, ptr, nbytes, dt, srcs) in enumerate(fe.plan.allreduce_points):
        tdt = torch.float64 if dt == "f64" else torch.float32
        cnt_k = nbytes // (8 if dt == "f64" else 4)
        seg = None if srcs is None else (view(srcs[0], srcs[1], tdt), view(srcs[2], srcs[3], tdt))
        if seg is not None:
            seg[1].mul_(10.0)                     # make the two segments distinguishable
        fe._ar[phase].setdefault(idx, []).append((view(ptr, cnt_k, tdt), k, seg))
    fe._bucket_ready = {}
    Engine._run(fe, [((lambda *a: 0), (), l.fn, l.tag) for l in fe.plan.fwd], "f", None)
    Engine._run(fe, [((lambda *a: 0), (), l.fn, l.tag) for l in fe.plan.bwd], "b", None)
    tot = float(sum(r + 1 for r in range(world)))
    n_f = n_b = 0
    for phase in ("f", "b"):
        for lst in fe._ar[phase].values():
            for t, k, seg in lst:
                if seg is None:
                    assert torch.equal(t, torch.full_like(t, tot))
                    n_f += 1
                else:
                    c = seg[0].numel()
                    assert torch.equal(t[:c], torch.full_like(t[:c], tot))
                    assert torch.equal(t[c:], torch.full_like(t[c:], 10.0 * tot))
                    n_b += 1
    assert n_f == 53 and n_b == 53
    q.put((rank, mine.numpy(), g.numpy(), nb, float((mean - m_ref).abs().max()),
           float((var * cnt / (cnt - 1) - v_ref).abs().max()), bucket_ranges(n, 1024)[-1]))
    dist.destroy_process_group()


def test_bucketed_gradient_mean_and_sync_bn_statistics():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.mean(np.stack([r[1] for r in res]), axis=0)          # concat + reduce_mean over towers
    for r in res:
        assert np.allclose(r[2], expect, atol=1e-6)
        assert r[3] == 10 and r[4] < 1e-12 and r[5] < 1e-12
        assert r[6] == (9216, 10007)
