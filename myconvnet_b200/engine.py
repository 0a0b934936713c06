"""Executes a Plan on one B200: arena allocation, variable initialisation, the training step.

Training-step semantics follow the reference (optimizers.py:89-177):
forward -> backward -> gradient mean over replicas -> [EMA shadows on pre-step values; BN
moving statistics] -> optimiser update (L2 folded in) -> decoupled weight decay, with
lr = base_lr * global_batch/256 * multiplier (optimizers.py:46,57).  Replicas are processes
(one per GPU); the gradient mean is an NCCL all-reduce over flat buckets; BN statistics are
all-reduced per layer (synchronised BN) instead of the reference's tower-after-tower chain.

torch is used for device memory, streams, CUDA-graph capture and torch.distributed only: every
arithmetic kernel on the step is a libmcn launch.
"""
import collections
import ctypes
import os
import sys

import numpy as np
import torch

from . import lib as _lib
from .plan import ConvDesc, Plan, Ptr

OPT_KINDS = {"nesterov": 0, "momentum": 0, "sgd": 0, "rmsprop": 1, "adam": 2}


def draw_initial_value(var, rng):
    """Initial value of a variable in its reference layout."""
    init = var.init
    shape = var.shape
    if init.kind == "constant":
        return np.full(shape, init.value, dtype=np.float32)
    if len(shape) >= 2:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    else:
        fan_in = fan_out = shape[0]
    fan = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": 0.5 * (fan_in + fan_out)}[init.mode]
    scale = init.scale / max(1.0, fan)
    if init.distribution == "uniform":
        lim = np.sqrt(3.0 * scale)
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)
    std = np.sqrt(scale) / 0.87962566103423978
    v = rng.standard_normal(size=shape)
    bad = np.abs(v) > 2.0
    while bad.any():                      # truncated normal: resample beyond 2 sigma
        v[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(v) > 2.0
    return (v * std).astype(np.float32)


class Engine(object):
    def __init__(self, model, optimizer="nesterov", keep=(), seed=0, conv_mode=2, device=None,
                 world_size=1, rank=0, process_group=None, use_cuda_graph=False,
                 fetch_pred=True, keep_grads=False, **kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("myconvnet_b200.Engine needs a CUDA device: there is no CPU execution "
                               "path (plans can be inspected on CPU through myconvnet_b200.plan.Plan)")
        self.lib = _lib.load()
        self.model = model
        self.graph = model.graph
        self.kw = dict(model._parameters)
        self.kw.update(kwargs)
        self.world = int(world_size)
        self.rank = int(rank)
        self.pg = process_group
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.opt_kind = OPT_KINDS[optimizer.lower()]
        self.plan = Plan(self.graph, world_size=self.world, keep=keep, conv_mode=conv_mode,
                         fetch_pred=fetch_pred, loss_scale=float(self.kw.get("loss_scaling_factor", 1.0)),
                         keep_grads=keep_grads)
        p = self.plan
        self.arena = torch.empty(p.arena_bytes + 1024, dtype=torch.uint8, device=self.device)
        base = self.arena.data_ptr()
        self.base = (base + 255) // 256 * 256
        self._skew = self.base - base
        self.arena.zero_()
        self._descs = {}
        self._resolve()
        self.global_step = 0
        self.batch = model._batch_size
        self.global_batch = self.batch * self.world
        self.base_lr = float(self.kw.get("base_learning_rate", 0.1))
        self.momentum = float(self.kw.get("momentum", 0.9))
        self.ema_decay = float(model.moving_average_decay)
        # GAN losses carry no L2 term (gan.py overrides _build_loss)
        self.l2 = float(self.kw.get("l2_reg", 1e-4)) if getattr(model, "uses_l2", True) else 0.0
        self.bias_norm_decay = bool(self.kw.get("bias_norm_decay", False))
        self.base_wd = float(self.kw.get("base_weight_decay", 0.0)) * self.global_batch / 256
        self.wd_scheduling = bool(self.kw.get("weight_decay_scheduling", True))
        # decoupled weight-decay form (optimizers.py:163-172): w -= wd*w | wd*sign(w) | pseudo-Huber
        self.huber_delta = self.kw.get("huber_decay_delta", None)
        self.wd_form = 2 if self.huber_delta is not None else (1 if self.kw.get("l1_weight_decay", False) else 0)
        self.l1 = float(self.kw.get("l1_reg", 0.0)) if getattr(model, "uses_l2", True) else 0.0
        # tf.clip_by_global_norm on the gradient of the full loss (optimizers.py:112-113); with
        # several GPUs it is applied to the rank-averaged gradient (one tower at the global batch)
        gt = self.kw.get("gradient_threshold", None)
        self.grad_threshold = None if gt is None else float(gt)
        self.random_seed = int(self.kw.get("random_seed", seed))
        self._pin_rings = {}
        self._loss_queue = collections.deque()     # enqueue_loss_read() -> pop_loss(), at most 3 pending
        self.hp_dev = self.view(Ptr(p.b_hp), 16, torch.float32)
        self._ws = _lib.ensure_workspace(self._workspace_bytes(), self.device)
        self._build_opt_table()
        self.init_variables(seed)
        self._graph_exec = None
        self._eager_steps = 0
        self.use_cuda_graph = use_cuda_graph
        self._inputs = {name: self.tensor_view(t) for name, t in self.graph.inputs.items()}
        self._pinned = {}
        self._bucket_ready = {}
        self._peer = None
        if self.world > 1:
            self._setup_grad_buckets()
            self._setup_peer_comm()

    def _workspace_bytes(self):
        """Reduction workspace (include/mcn.h, mcn_set_workspace): the fixed limb area plus the
        split-K slices the largest wgrad of this plan would like."""
        need = int(self.lib.mcn_workspace_min_bytes())
        cap = 1 << 30
        for l in self.plan.bwd:
            if l.fn == "mcn_conv2d_wgrad_tc":
                need = max(need, int(self.lib.mcn_conv2d_wgrad_workspace_bytes(self._carg(l.args[0]), l.args[4], 0)))
            elif l.fn == "mcn_stem_conv_wgrad":
                need = max(need, int(self.lib.mcn_conv2d_wgrad_workspace_bytes(self._carg(l.args[0]), 0, 1)))
            elif l.fn in ("mcn_conv2d_wgrad_direct", "mcn_dwconv2d_bwd_filter"):
                d = l.args[0]
                n = d.kh * d.kw * d.Cin * (l.args[1] if l.fn == "mcn_dwconv2d_bwd_filter" else d.Cout)
                need = max(need, int(self.lib.mcn_workspace_min_bytes()) + min(64 * n * 4 + 4096, cap))
        return min(need, int(self.lib.mcn_workspace_min_bytes()) + cap)

    # ------------------------------------------------------------------ memory views
    def addr(self, ptr):
        return self.base + ptr.buf.offset + ptr.off

    def view(self, ptr, count, dtype):
        esz = torch.empty(0, dtype=dtype).element_size()
        start = self._skew + ptr.buf.offset + ptr.off
        return self.arena[start:start + count * esz].view(dtype)

    def tensor_view(self, t, stored_dtype=None):
        dt = stored_dtype or self.plan._logits_dtype(t)
        tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "i32": torch.int32, "u8": torch.uint8}[dt]
        return self.view(self.plan.tbuf[t], t.size, tdt).view(*t.shape) if t.shape else \
            self.view(self.plan.tbuf[t], 1, tdt)

    def fetch(self, t):
        """Copy a graph tensor's current value to host as float32 numpy."""
        v = self.tensor_view(t)
        return v.float().cpu().numpy() if v.dtype != torch.int32 else v.cpu().numpy()

    def fetch_grad(self, t):
        """Gradient of the loss w.r.t. graph tensor t after the last backward pass, as float32 numpy
        (Engine(keep_grads=True) only: gradient buffers are then never recycled).  None when the
        tensor received no gradient."""
        p = self.plan.grad_ptr.get(t)
        if p is None:
            return None
        dt = self.plan._logits_dtype(t)
        tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[dt]
        return self.view(p, t.size, tdt).view(*t.shape).float().cpu().numpy()

    # ------------------------------------------------------------------ launch resolution
    def _carg(self, a):
        if isinstance(a, Ptr):
            return self.addr(a)
        if isinstance(a, ConvDesc):
            k = a.key()
            if k not in self._descs:
                self._descs[k] = _lib.ConvDescC(*k)
            return ctypes.byref(self._descs[k])
        return a

    def _resolve(self):
        def conv(launches):
            out = []
            for l in launches:
                fn = getattr(self.lib, l.fn)
                out.append((fn, tuple(self._carg(a) for a in l.args), l.fn, l.tag))
            return out
        self._fwd = conv(self.plan.fwd)
        self._bwd = conv(self.plan.bwd)
        self._inf = conv(self.plan.inf)
        # collective points, keyed by launch index
        self._ar = {"f": {}, "b": {}}
        for k, (phase, idx, ptr, nbytes, dt, srcs) in enumerate(self.plan.allreduce_points):
            tdt = torch.float64 if dt == "f64" else torch.float32
            n = nbytes // (8 if dt == "f64" else 4)
            seg = None
            if srcs is not None:
                s0, n0, s1, n1 = srcs
                seg = (self.view(s0, n0, tdt), self.view(s1, n1, tdt))
            self._ar[phase].setdefault(idx, []).append((self.view(ptr, n, tdt), k, seg))
        z0, z1 = self.plan.region_span["zero"]
        self._zero_ptr = self.base + z0
        self._zero_n = (z1 - z0) // 4

    def _run(self, launches, phase, stream):
        ar = self._ar[phase]
        check = _lib.check
        ready = self._bucket_ready if phase == "b" else {}
        if not ar and not ready:
            for fn, args, name, tag in launches:
                rc = fn(*args, stream)
                if rc:
                    check(rc, name + " [" + tag + "]")
            return
        import torch.distributed as dist
        peer = self._peer
        for i, (fn, args, name, tag) in enumerate(launches):
            if i in ar:
                for t, k, seg in ar[i]:
                    if peer is not None:
                        # one-shot exchange over NVLink peer memory (csrc/comm.cu)
                        if seg is None:
                            a0, n0, a1, n1 = t.data_ptr(), t.numel(), None, 0
                        else:
                            a0, n0, a1, n1 = seg[0].data_ptr(), seg[0].numel(), seg[1].data_ptr(), seg[1].numel()
                        check(self.lib.mcn_peer_allreduce(
                            peer["peers"], peer["mail"][k], peer["stride"][k], peer["flag"][k], peer["ctr"] + 16 * k,
                            1 if t.dtype == torch.float64 else 0, a0, n0, a1, n1, t.data_ptr(), self.rank,
                            self.world, stream), "peer_allreduce")
                    else:
                        if seg is not None:
                            t[:seg[0].numel()].copy_(seg[0])
                            t[seg[0].numel():].copy_(seg[1])
                        dist.all_reduce(t, group=self.pg)
            rc = fn(*args, stream)
            if rc:
                check(rc, name + " [" + tag + "]")
            if i in ready:
                # every gradient of these buckets is final: exchange them while backward continues
                self._buckets.after_launch(i)

    # ------------------------------------------------------------------ multi-GPU set-up
    def _setup_grad_buckets(self):
        # 8 M-element buckets (4 for ResNet-50).  Measured, ms per step on 2 / 8 B200 (ResNet-50 b256):
        # 4 M buckets overlapped 21.06 / 20.89, 8 M overlapped 20.87 / -, one bucket after backward
        # 20.91 / 20.73 — NCCL's CTAs take SMs from the persistent conv grids, so overlap hides about
        # what it costs; fewer, larger buckets keep the overlap and most of the difference.
        from .dist import BucketOverlap
        p = self.plan
        self._flat_grads = self.view(Ptr(p.b_grad), p.n_train, torch.float32)
        self._buckets = BucketOverlap(
            self._flat_grads, p.grad_bucket_schedule(int(self.kw.get("bucket_elems",
                                                                    int(os.environ.get("MCN_BUCKET_ELEMS", 8 * 1024 * 1024))))),
            group=self.pg, overlap=os.environ.get("MCN_OVERLAP_GRADS", "1") != "0")
        self._bucket_ready = self._buckets.ready
        self._bucket_tail = self._buckets.tail

    def _setup_peer_comm(self):
        """Symmetric-memory mailboxes for the synchronised-BN exchanges (mcn_peer_allreduce).  Falls
        back to NCCL all-reduces when peer mapping is not available (MCN_PEER_ALLREDUCE=0 forces it)."""
        pts = self.plan.allreduce_points
        if not pts or os.environ.get("MCN_PEER_ALLREDUCE", "1") == "0":
            return
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            mail, stride, off = [], [], 0
            for _, _, _, nbytes, _, _ in pts:
                mail.append(off)
                # tagged 8-byte words (csrc/comm.cu, "LL" protocol): a slot is twice the payload
                stride.append((self.world * nbytes * 2 + 255) // 256 * 256)
                off += 2 * stride[-1]                      # two mailboxes per point (sequence parity)
            flag = [off + 512 * k for k in range(len(pts))]       # [world] uint64 per point, world <= 64
            off += 512 * len(pts)
            buf = symm.empty(off, dtype=torch.uint8, device=self.device)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            group = self.pg if self.pg is not None else dist.group.WORLD
            hdl = symm.rendezvous(buf, group)
            ctr = torch.zeros(2 * len(pts), dtype=torch.int64, device=self.device)   # [sequence, ticket] per point
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)                          # every region is zeroed before any push
            self._peer = {"buf": buf, "hdl": hdl, "ctr_t": ctr, "ctr": ctr.data_ptr(),
                          "peers": int(hdl.buffer_ptrs_dev), "mail": mail, "stride": stride, "flag": flag}
        except Exception as e:                                     # noqa: BLE001
            sys.stderr.write("[myconvnet_b200] peer-memory all-reduce unavailable (%s: %s); "
                             "using NCCL for the BN statistics\n" % (type(e).__name__, e))
            self._peer = None

    # ------------------------------------------------------------------ variables
    def _var_view(self, buf, v, n=None):
        return self.view(Ptr(buf, self.plan.var_off[v] * 4), n or v.storage_size, torch.float32)

    def _to_storage(self, v, value):
        """reference layout -> device storage layout (im2col-padded stems)."""
        value = np.asarray(value, dtype=np.float32).reshape(v.shape)
        if v.storage_shape != v.shape:
            kpad, co = v.storage_shape
            flat = value.reshape(-1, co)
            out = np.zeros((kpad, co), dtype=np.float32)
            rows = getattr(v, "storage_rows", None)
            if rows is not None:
                out[rows] = flat            # gather-stem layout: widened filter rows, 4-channel pixels
            else:
                out[:flat.shape[0]] = flat
            return out
        return value

    def _from_storage(self, v, arr):
        arr = np.asarray(arr).reshape(v.storage_shape)
        if v.storage_shape != v.shape:
            rows = getattr(v, "storage_rows", None)
            if rows is not None:
                return arr[rows].reshape(v.shape).copy()
            k = int(np.prod(v.shape[:-1]))
            return arr[:k].reshape(v.shape).copy()
        return arr.reshape(v.shape).copy()

    def init_variables(self, seed=0):
        rng = np.random.default_rng(seed)
        vals = {v.name: draw_initial_value(v, rng) for v in self.graph.vars.values()}
        self.set_variables(vals, reset_state=True)

    def set_variables(self, values, reset_state=True):
        """values: {name: array in reference layout}.  EMA shadows are (re)initialised to the
        value (tf ExponentialMovingAverage initialises the shadow to the variable)."""
        p = self.plan
        for v in p.all_vars:
            if v.name not in values:
                continue
            arr = torch.from_numpy(self._to_storage(v, values[v.name]).reshape(-1)).to(self.device)
            self._var_view(p.b_param, v).copy_(arr)
            if reset_state:
                self._var_view(p.b_ema, v).copy_(arr)
        if reset_state:
            self.view(Ptr(p.b_mom), max(p.n_train, 1), torch.float32).zero_()
            # RMSProp's mean-square accumulator starts at one
            self.view(Ptr(p.b_v), max(p.n_train, 1), torch.float32).fill_(1.0 if self.opt_kind == 1 else 0.0)
            self.global_step = 0
        self.refresh_operand_copies()
        torch.cuda.synchronize(self.device)

    def refresh_operand_copies(self, ema=False):
        """bf16 tensor-core operand copies of the fp32 master weights (or of their EMA shadows)."""
        p = self.plan
        st = torch.cuda.current_stream(self.device).cuda_stream
        p.phase = "infer" if ema else "train"
        try:
            for v in p.all_vars:
                if v.needs_bf16 or v.needs_bf16_t:
                    taps, ci, co = v.gemm_dims
                    _lib.check(self.lib.mcn_weight_prep(
                        self.addr(p.pvar(v)), taps, ci, co,
                        self.addr(p.pbf16(v)) if v.needs_bf16 else None,
                        self.addr(p.pbf16t(v)) if v.needs_bf16_t else None, st), "weight_prep")
        finally:
            p.phase = "train"

    def predict(self, X, fetch="pred"):
        """Inference as the reference's ConvNet.predict (convnet.py:609-665): is_train=False, so
        every variable is replaced by its EMA shadow (convnet.py:1406) and batch-norm uses the
        (shadowed) moving statistics.  Returns the model's `pred` tensor as float32 numpy."""
        self.load_inputs(X=X)
        self.refresh_operand_copies(ema=True)
        _lib.use_workspace(self._ws)
        st = torch.cuda.current_stream(self.device).cuda_stream
        for fn, args, name, tag in self._inf:
            rc = fn(*args, st)
            if rc:
                _lib.check(rc, name + " [" + tag + "]")
        t = getattr(self.model, fetch) if isinstance(fetch, str) else fetch
        return self.fetch(t)

    def maxpool_argmax(self, node):
        """TF-convention argmax (h*W + w)*C + c of a max_pool node of the last forward pass, int32
        numpy [N,Ho,Wo,C] (tf.nn.max_pool_with_argmax without the batch term)."""
        x, y = node.inputs[0], node.outputs[0]
        a = node.attrs
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        if not a.get("tap_form"):
            return self.view(Ptr(a["argmax"]), y.size, torch.int32).cpu().numpy().reshape(y.shape)
        out = torch.empty(y.size, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.mcn_maxpool_tap_to_argmax(self.addr(Ptr(a["argmax"])), n, h, w, c, a["k"][0], a["k"][1],
                                                      a["s"][0], a["s"][1], a["pad"][0], a["pad"][1], ho, wo,
                                                      out.data_ptr(), st), "maxpool_tap_to_argmax")
        return out.cpu().numpy().reshape(y.shape)

    def get_variables(self, ema=False):
        p = self.plan
        buf = p.b_ema if ema else p.b_param
        return {v.name: self._from_storage(v, self._var_view(buf, v).cpu().numpy()) for v in p.all_vars}

    # TF slot names of the three optimisers (tf.train.MomentumOptimizer / RMSPropOptimizer /
    # AdamOptimizer), as tf.train.Saver writes them next to the variables (optimizers.py:312)
    SLOT_NAMES = {0: ("Momentum", None), 1: ("RMSProp_1", "RMSProp"), 2: ("Adam", "Adam_1")}

    def get_optimizer_state(self):
        """{<var>/<slot>: array in reference layout} for the optimiser slots (momentum / RMSProp
        mean-square + momentum / Adam moments)."""
        p = self.plan
        m_name, v_name = self.SLOT_NAMES[self.opt_kind]
        out = {}
        for v in p.trainable:
            out[v.name + "/" + m_name] = self._from_storage(v, self._var_view(p.b_mom, v).cpu().numpy())
            if v_name is not None:
                out[v.name + "/" + v_name] = self._from_storage(v, self._var_view(p.b_v, v).cpu().numpy())
        return out

    def set_optimizer_state(self, state, global_step=None):
        """Inverse of get_optimizer_state; slots that are not in `state` keep their value."""
        p = self.plan
        m_name, v_name = self.SLOT_NAMES[self.opt_kind]
        for v in p.trainable:
            for buf, slot in ((p.b_mom, m_name), (p.b_v, v_name)):
                key = None if slot is None else v.name + "/" + slot
                if key is not None and key in state:
                    arr = torch.from_numpy(self._to_storage(v, state[key]).reshape(-1)).to(self.device)
                    self._var_view(buf, v).copy_(arr)
        if global_step is not None:
            self.global_step = int(global_step)
        torch.cuda.synchronize(self.device)

    def set_ema(self, values):
        """EMA shadows from {name: array in reference layout}."""
        p = self.plan
        for v in p.all_vars:
            if v.name in values:
                arr = torch.from_numpy(self._to_storage(v, values[v.name]).reshape(-1)).to(self.device)
                self._var_view(p.b_ema, v).copy_(arr)
        torch.cuda.synchronize(self.device)

    def save_checkpoint(self, path, naming="reference", trainer=None):
        """Everything tf.train.Saver keeps for a resumable run (reference optimizers.py:312 saves all
        global variables): the variables and their EMA shadows under the reference's names
        (checkpoint.py), the optimiser slots under TF's slot names, global_step and, when a
        Trainer is passed, its step counter; .npz."""
        from . import checkpoint
        extra = dict(self.get_optimizer_state())
        extra["global_step"] = np.asarray(self.global_step, dtype=np.int64)
        if trainer is not None:
            extra["trainer/curr_step"] = np.asarray(trainer.curr_step, dtype=np.int64)
        checkpoint.save_npz(path, self.get_variables(), self.get_variables(ema=True), naming=naming,
                            extra=extra)

    def load_checkpoint(self, path, naming="reference", prefer_ema=False, resume=True, trainer=None):
        """Loads what matches by name and shape; returns the names that were not found.  The EMA
        shadows restart from the loaded values unless the file carries its own.  resume=True also
        restores optimiser slots and global_step when the file has them (a run continues where it
        stopped: EMA warm-up, Adam bias correction and momentum are not reset); resume=False is a
        transfer-style load that starts the optimiser from scratch."""
        from . import checkpoint
        expected = {v.name: v.shape for v in self.plan.all_vars}
        var, ema = checkpoint.load_npz(path, expected=expected, naming=naming, prefer_ema=prefer_ema)
        self.set_variables(var, reset_state=True)
        self.set_ema(ema)
        if resume:
            extra = checkpoint.load_extra(path)
            self.set_optimizer_state(extra, global_step=extra.get("global_step"))
            if trainer is not None and "trainer/curr_step" in extra:
                trainer.curr_step = int(extra["trainer/curr_step"])
        return sorted(set(expected) - set(var))

    def get_gradients(self):
        p = self.plan
        return {v.name: self._from_storage(v, self._var_view(p.b_grad, v).cpu().numpy())
                for v in p.trainable}

    def _build_opt_table(self):
        p = self.plan
        n = len(p.all_vars)
        table = (_lib.OptTensorC * n)()
        max_n = 1
        for i, v in enumerate(p.all_vars):
            e = table[i]
            off = p.var_off[v] * 4
            e.w = self.base + p.b_param.offset + off
            e.ema = self.base + p.b_ema.offset + off
            if v.trainable:
                e.g = self.base + p.b_grad.offset + off
                e.m = self.base + p.b_mom.offset + off
                e.v = self.base + p.b_v.offset + off
            e.n = v.storage_size
            if v.needs_bf16:
                e.w_bf16 = self.addr(p.pbf16(v))
            if v.needs_bf16_t:
                e.w_bf16_t = self.addr(p.pbf16t(v))
            if v.gemm_dims:
                e.taps, e.cin, e.cout = v.gemm_dims
            else:
                e.taps, e.cin, e.cout = 1, 1, max(v.storage_size, 1)
            decayed = v.kind == "weight" or (self.bias_norm_decay and v.kind in ("bias", "norm"))
            # the regularisation LOSS sums over every weight variable, frozen ones included
            # (convnet.py:535-563 takes the whole 'weight_variables' collection); gradients and the
            # decoupled decay only touch trainable ones (g == NULL skips both in the kernel)
            e.l2 = self.l2 if decayed else 0.0
            e.l1 = self.l1 if decayed else 0.0
            e.wd = self.base_wd if (decayed and v.trainable) else 0.0
            max_n = max(max_n, v.storage_size)
        raw = bytes(table)
        self.opt_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        self.opt_n = n
        self.opt_max_n = max_n

    # ------------------------------------------------------------------ the step
    def _set_hyper(self, lr_multiplier):
        t = self.global_step
        lr = self.base_lr * self.global_batch / 256.0 * lr_multiplier
        d_t = min(self.ema_decay, (1.0 + t) / (10.0 + t))
        b1, b2 = self.momentum, (0.9 if self.opt_kind == 1 else 0.999)
        adam_lr = lr * np.sqrt(1.0 - b2 ** (t + 1)) / (1.0 - b1 ** (t + 1)) if self.opt_kind == 2 else lr
        hp = self._pinned_slot("hp", (16,), torch.float32, slots=4)
        hp.zero_()
        hp[0], hp[1], hp[2], hp[3] = lr, b1, b2, 1e-3
        hp[4], hp[5] = d_t, adam_lr
        hp[6] = 1.0 / (self.world * self.plan.loss_scale)
        hp[7] = lr_multiplier if self.wd_scheduling else 1.0
        hp[8] = self.grad_threshold if self.grad_threshold is not None else 0.0
        hp[9] = float(self.wd_form)
        hp[10] = float(self.huber_delta) if self.huber_delta is not None else 1.0
        # seed / step of the random train-time ops (dropout, stochastic depth): uint32 bit patterns
        hpi = hp.view(torch.int32)
        hpi[12] = int(np.int32(np.uint32(self.random_seed & 0xFFFFFFFF)))
        hpi[13] = int(np.int32(np.uint32(t & 0xFFFFFFFF)))
        self.hp_dev.copy_(hp, non_blocking=True)
        self._pinned_done("hp")

    def _pinned_slot(self, key, shape, dtype, slots=2):
        """Pinned host staging buffer for an asynchronous H2D copy.  A slot is only rewritten after
        the copy that last read it has completed (event recorded by _pinned_done): with
        fetch_loss=False or CUDA-graph replay the host runs ahead of the GPU, and rewriting a
        buffer an earlier copy_(non_blocking=True) is still reading would hand step N the
        hyper-parameters or batch of step N+k."""
        ring = self._pin_rings.setdefault(key, {"bufs": [], "events": [], "next": 0})
        if not ring["bufs"] or tuple(ring["bufs"][0].shape) != tuple(shape) or ring["bufs"][0].dtype != dtype:
            ring["bufs"] = [torch.empty(tuple(shape), dtype=dtype).pin_memory() for _ in range(slots)]
            ring["events"] = [None] * slots
            ring["next"] = 0
        i = ring["next"]
        ring["next"] = (i + 1) % len(ring["bufs"])
        if ring["events"][i] is not None:
            ring["events"][i].synchronize()
        ring["cur"] = i
        return ring["bufs"][i]

    def _pinned_done(self, key, stream=None):
        ring = self._pin_rings[key]
        ev = torch.cuda.Event()
        ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
        ring["events"][ring["cur"]] = ev

    def load_inputs(self, **arrays):
        """Host (numpy / pinned torch) -> device input buffers, asynchronously on the current
        stream.  Returns the bytes copied."""
        nbytes = 0
        off = int(getattr(self.model, "label_offset", 0))
        for name, arr in arrays.items():
            dst = self._inputs[name]
            if name == "Y" and off:
                # task bases that take labels with 0 = ignore (segnet.py:50) store class-1 on device
                arr = (np.asarray(arr) if isinstance(arr, np.ndarray) else arr.numpy()).astype(np.int32) - np.int32(off)
            if dst.dtype == torch.uint8 and (arr.dtype != np.uint8 if isinstance(arr, np.ndarray) else arr.dtype != torch.uint8):
                raise TypeError("this model was built with input_dtype='u8': pass raw uint8 images")
            staged = isinstance(arr, np.ndarray)
            if staged:
                pin = self._pinned_slot("in:" + name, dst.shape, dst.dtype)
                pin.copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(dst.dtype).view(dst.shape))
                arr = pin
            dst.copy_(arr.view(dst.shape), non_blocking=True)
            if staged:
                self._pinned_done("in:" + name)
            nbytes += dst.numel() * dst.element_size()
        return nbytes

    def prefetch_inputs(self, **arrays):
        """Start copying the NEXT step's host batch (pinned memory) into device staging buffers on a
        side stream, so the PCIe transfer overlaps the current step's kernels.  The following
        train_step(prefetched=True) moves staging -> input buffers with a device-to-device copy."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staging = {n: torch.empty_like(t) for n, t in self._inputs.items()}
            self._staged = torch.cuda.Event()
            self._consumed = None
        if self._consumed is not None:
            self._copy_stream.wait_event(self._consumed)     # staging is free again
        off = int(getattr(self.model, "label_offset", 0))
        nbytes = 0
        with torch.cuda.stream(self._copy_stream):
            for name, arr in arrays.items():
                dst = self._staging[name]
                if isinstance(arr, np.ndarray) or (name == "Y" and off):
                    a = np.asarray(arr) if isinstance(arr, np.ndarray) else arr.numpy()
                    if name == "Y" and off:
                        a = a.astype(np.int32) - np.int32(off)
                    pin = self._pinned_slot("pre:" + name, dst.shape, dst.dtype)
                    pin.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(dst.dtype).view(dst.shape))
                    arr = pin
                    dst.copy_(arr.view(dst.shape), non_blocking=True)
                    self._pinned_done("pre:" + name, self._copy_stream)
                else:
                    dst.copy_(arr.view(dst.shape), non_blocking=True)
                nbytes += dst.numel() * dst.element_size()
            self._staged.record(self._copy_stream)
        return nbytes

    def _consume_prefetched(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged)
        for name, dst in self._inputs.items():
            dst.copy_(self._staging[name], non_blocking=True)
        self._consumed = torch.cuda.Event()
        self._consumed.record(cur)

    def _step_body(self, backward=True, update=True):
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.use_workspace(self._ws)
        _lib.check(self.lib.mcn_fill_f32(self._zero_ptr, self._zero_n, 0.0, st), "zero")
        self._run(self._fwd, "f", st)
        if not backward:
            return
        self._run(self._bwd, "b", st)
        if self.world > 1:
            self._allreduce_grads()
        if update:
            p = self.plan
            gnorm = None
            if self.grad_threshold is not None:
                gnorm = self.addr(Ptr(p.b_gnorm))
                _lib.check(self.lib.mcn_grad_sqnorm(self.opt_table.data_ptr(), self.opt_n, self.opt_max_n,
                                                    self.addr(Ptr(p.b_hp)), gnorm, st), "grad_sqnorm")
            _lib.check(self.lib.mcn_opt_step(self.opt_kind, self.opt_table.data_ptr(), self.opt_n,
                                             self.opt_max_n, self.addr(Ptr(p.b_hp)),
                                             self.addr(p.loss_slots["l2"]) if "l2" in p.loss_slots else None,
                                             gnorm, st), "opt_step")

    def _allreduce_grads(self):
        """Buckets whose gradients were final early are already in flight (started from _run);
        start the rest and wait for all of them before the optimiser reads the buffer."""
        self._buckets.finish()

    def train_step(self, X=None, Y=None, lr_multiplier=1.0, fetch_loss=True, update=True,
                   prefetched=False):
        """One optimisation step.  X: fp32 NHWC images in [0,1]; Y: int32 labels (-1 = none).
        Returns the reference's loss value (data term + L2 term) when fetch_loss.
        prefetched=True consumes the batch uploaded by prefetch_inputs()."""
        if prefetched:
            self._consume_prefetched()
        elif X is not None:
            self.load_inputs(X=X, Y=Y)
        self._set_hyper(lr_multiplier)
        if self.use_cuda_graph and update and self._eager_steps >= 1:
            if self._graph_exec is None:
                self._capture(update)
            self._graph_exec.replay()
        else:
            # the first step always runs eagerly: lazy driver/attribute initialisation must not
            # happen inside a stream capture
            self._step_body(update=update)
            self._eager_steps += 1
        if update:
            self.global_step += 1
        return self.read_loss() if fetch_loss else None

    def _capture(self, update):
        # warm-up launch outside capture so lazy CUDA/driver initialisation is done
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                self._step_body(update=update)
        torch.cuda.current_stream(self.device).wait_stream(s)
        self._graph_exec = g

    def forward(self, X=None, Y=None):
        if X is not None:
            self.load_inputs(X=X, Y=Y)
        self._step_body(backward=False)

    def read_loss(self):
        p = self.plan
        if "loss" not in p.loss_slots:
            return None
        node = self.graph.losses[0].node
        # the slots are exact fixed-point accumulators (xsum limbs), decoded on the host
        if "loss_g" in p.loss_slots:
            v = self.view(p.loss_slots["loss"], 9, torch.int64).cpu().numpy().reshape(3, 3)
            d, g = _lib.xsum_value(v[0]), _lib.xsum_value(v[1])
            self.last_losses = (d / node.attrs["rows"], g / node.attrs["rows"])
            return self.last_losses[0]
        v = self.view(p.loss_slots["loss"], 6, torch.int64).cpu().numpy().reshape(2, 3)
        data = _lib.xsum_value(v[0]) / node.attrs["rows"]
        return data + _lib.xsum_value(v[1])

    def enqueue_loss_read(self):
        """Asynchronous form of read_loss(): copies this step's loss accumulators into a pinned ring
        slot on the current stream (before the next step's zero-fill, by stream order) and returns at
        once; pop_loss() later waits for that copy only.  An input pipeline reads the loss of step
        i while step i+1 is already running, so the device never idles on the host's read-back."""
        p = self.plan
        if "loss" not in p.loss_slots:
            return
        n = 9 if "loss_g" in p.loss_slots else 6
        if len(self._loss_queue) >= 3:
            # the ring has four slots and a slot is only protected against the COPY still being in flight,
            # not against the host not having read it yet
            raise RuntimeError("enqueue_loss_read: at most 3 losses may be pending; call pop_loss() first")
        pin = self._pinned_slot("loss", (n,), torch.int64, slots=4)
        pin.copy_(self.view(p.loss_slots["loss"], n, torch.int64), non_blocking=True)
        self._pinned_done("loss")
        ring = self._pin_rings["loss"]
        self._loss_queue.append((pin, ring["events"][ring["cur"]]))

    def pop_loss(self):
        """The oldest loss queued by enqueue_loss_read() (blocks until its copy has landed)."""
        pin, ev = self._loss_queue.popleft()
        ev.synchronize()
        node = self.graph.losses[0].node
        v = pin.numpy()
        if v.size == 9:
            v = v.reshape(3, 3)
            self.last_losses = (_lib.xsum_value(v[0]) / node.attrs["rows"], _lib.xsum_value(v[1]) / node.attrs["rows"])
            return self.last_losses[0]
        v = v.reshape(2, 3)
        return _lib.xsum_value(v[0]) / node.attrs["rows"] + _lib.xsum_value(v[1])

    def launches_per_step(self):
        return 1 + len(self._fwd) + len(self._bwd) + 1
