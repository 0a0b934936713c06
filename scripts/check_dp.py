"""Multi-GPU check (launch with torchrun, one rank per GPU):
  1. mcn_peer_allreduce (NVLink mailbox all-reduce) against the exact sum,
  2. n-GPU data-parallel ResNet-50 steps (synchronised BN through the peer all-reduce, gradient
     buckets overlapped with backward) == ONE device at the global batch (SURVEY 8e parity target),
     eager and CUDA-graph replay.
Prints one PASS/FAIL line per check on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myconvnet_b200 import lib as L                      # noqa: E402
from myconvnet_b200.engine import Engine, draw_initial_value   # noqa: E402
from myconvnet_b200.zoo import resnet50                  # noqa: E402


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    dist.init_process_group("nccl")
    dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
    shape, ncls, b = [64, 64, 3], 16, 8
    ok = True

    def report(name, good, detail=""):
        nonlocal ok
        ok = ok and good
        if rank == 0:
            print("%s %s %s" % ("PASS" if good else "FAIL", name, detail), flush=True)

    pm, _ = resnet50(shape, ncls, batch_size=b, compute_dtype=dtype)
    rng = np.random.default_rng(0)
    vals = {v.name: draw_initial_value(v, rng) for v in pm.graph.vars.values()}
    for k in vals:
        if k.endswith("gamma") and not vals[k].any():
            vals[k] = np.full_like(vals[k], 0.5)
    Xg = rng.uniform(size=[b * world] + shape).astype(np.float32)
    Yg = rng.integers(0, ncls, size=b * world).astype(np.int32)
    X, Y = Xg[rank * b:(rank + 1) * b], Yg[rank * b:(rank + 1) * b]

    for graph in (False, True):
        eng = Engine(pm, world_size=world, rank=rank, process_group=dist.group.WORLD,
                     use_cuda_graph=graph, base_learning_rate=0.05)
        eng.set_variables(vals)
        if not graph:
            report("peer memory mapped", eng._peer is not None)
            if eng._peer is not None:
                # 1. the collective itself, on the first BN layer's mailbox
                pr = eng._peer
                pts = eng.plan.allreduce_points
                # the first BN layer's mailbox (one block) and the largest fp64 one (several blocks share the vector)
                big = max((k for k in range(len(pts)) if pts[k][4] == "f64"), key=lambda k: pts[k][3])
                for k in (0, big):
                    n = pts[k][3] // 8
                    t = torch.arange(n, dtype=torch.float64, device="cuda") * (rank + 1) + 0.25 * rank
                    for rep in range(3):
                        u = t.clone() + rep
                        L.check(eng.lib.mcn_peer_allreduce(pr["peers"], pr["mail"][k], pr["stride"][k], pr["flag"][k],
                                                           pr["ctr"] + 16 * k, 1, u.data_ptr(), n, None, 0, u.data_ptr(),
                                                           rank, world, torch.cuda.current_stream().cuda_stream))
                        torch.cuda.synchronize()
                        ref = sum(torch.arange(n, dtype=torch.float64) * (r + 1) + 0.25 * r + rep for r in range(world))
                        report("peer all-reduce n=%d rep %d" % (n, rep), bool(torch.equal(u.cpu(), ref)))
        steps = 4 if graph else 2
        losses = [eng.train_step(X, Y) for _ in range(steps)]
        v_dp = eng.get_variables()
        lt = torch.tensor(losses, dtype=torch.float64, device="cuda")
        dist.all_reduce(lt)
        lt /= world
        if rank == 0:
            pm1, _ = resnet50(shape, ncls, batch_size=b * world, compute_dtype=dtype)
            e1 = Engine(pm1, base_learning_rate=0.05)
            e1.set_variables(vals)
            l1 = [e1.train_step(Xg, Yg) for _ in range(steps)]
            v1 = e1.get_variables()
            worst = max((rel_l2(v_dp[k], v1[k]), k) for k in v1 if "weights" in k or k.endswith("gamma"))
            tol = 2e-3 if dtype == "f32" else 6e-2
            # the L2 term is identical on every rank; the data term averages over ranks
            report("dp%d == 1 device at the global batch (%s, graph=%s)" % (world, dtype, graph),
                   worst[0] < tol and abs(float(lt[0]) - l1[0]) < 1e-3 * abs(l1[0]) + (1e-4 if dtype == "f32" else 3e-2),
                   "worst var %s %.2e; losses dp %s vs single %s" % (worst[1], worst[0],
                                                                   ["%.5f" % x for x in lt.tolist()],
                                                                   ["%.5f" % x for x in l1]))
            del e1
        # all ranks hold identical variables after the steps
        chk = torch.tensor([float(np.sum(v_dp["block_None/logits/weights"].astype(np.float64)))], dtype=torch.float64, device="cuda")
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        report("replicas identical (graph=%s)" % graph, float(hi - lo) == 0.0)
        del eng
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("ALL PASS" if ok else "SOME FAILED", flush=True)
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
