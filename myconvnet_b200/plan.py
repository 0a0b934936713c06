"""Lower a recorded layer graph to a fixed launch sequence over one device arena.

Input: graph.Graph built by the ConvNet facade.  Output: a ``Plan`` — buffer table with byte
offsets into a single arena, forward launches, backward launches (reverse-mode differentiation
done here, op by op), and the optimiser table.  Everything is decided ahead of time (shapes are
static, as in the reference's TF graph mode), so a training step is a flat list of C-ABI calls
that can be captured into a CUDA graph.  Planning runs without a GPU; tests inspect plans on CPU.

Fusion decided here (SURVEY.md 2.2 "new sm_100a kernel" column):
  conv -> BN -> act                     : BN statistics + one apply pass with the activation
  conv -> BN -> (+ skip) -> act         : same pass also adds the residual
  dense -> cast(f32)                    : GEMM epilogue writes fp32 logits
  softmax-xent forward + gradient (+ softmax probabilities) in one kernel
"""
import collections
import os

import numpy as np

ALIGN = 256
# halo_eligible() of csrc/conv_tc.cu (kHaloMinEff, kHaloMinHw): keep in step
HALO_MIN_EFF = 0.85
HALO_MIN_HW = 48 * 48
DT_SIZE = {"f32": 4, "bf16": 2, "i32": 4, "f64": 8, "u8": 1}
DT_CODE = {"f32": 0, "bf16": 1, "u8": 2}


class Buf(object):
    __slots__ = ("name", "nbytes", "offset", "region")

    def __init__(self, name, nbytes, region):
        self.name = name
        self.nbytes = int(nbytes)
        self.offset = None
        self.region = region


class Ptr(object):
    """Placeholder for arena_base + buf.offset + off, resolved by the engine."""
    __slots__ = ("buf", "off")

    def __init__(self, buf, off=0):
        self.buf = buf
        self.off = int(off)

    def __add__(self, n):
        return Ptr(self.buf, self.off + int(n))


NULL = None


class Launch(object):
    __slots__ = ("fn", "args", "tag")

    def __init__(self, fn, args, tag=""):
        self.fn = fn
        self.args = args
        self.tag = tag


class ConvDesc(object):
    """Mirror of mcn_conv_desc (include/mcn.h)."""
    FIELDS = ("N", "H", "W", "Cin", "Cout", "kh", "kw", "sh", "sw", "dh", "dw", "pad_t", "pad_l",
              "Ho", "Wo")

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, int(kw[f]))

    def key(self):
        return tuple(getattr(self, f) for f in self.FIELDS)


def _align(n, a=ALIGN):
    return (n + a - 1) // a * a


class TempAllocator(object):
    """First-fit free-list allocator simulated at plan time for backward temporaries."""

    def __init__(self):
        self.free = []       # (offset, size)
        self.top = 0
        self.peak = 0

    def alloc(self, nbytes):
        nbytes = _align(nbytes)
        for i, (off, size) in enumerate(self.free):
            if size >= nbytes:
                if size == nbytes:
                    self.free.pop(i)
                else:
                    self.free[i] = (off + nbytes, size - nbytes)
                return off, nbytes
        off = self.top
        self.top += nbytes
        self.peak = max(self.peak, self.top)
        return off, nbytes

    def release(self, off, nbytes):
        self.free.append((off, nbytes))
        self.free.sort()
        merged = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        if merged and merged[-1][0] + merged[-1][1] == self.top:
            self.top = merged[-1][0]
            merged.pop()
        self.free = merged


class Plan(object):
    def __init__(self, graph, world_size=1, keep=(), conv_mode=2, fetch_pred=True,
                 sync_bn=True, loss_scale=1.0, fuse_bn_stats=None, keep_grads=False):
        self.graph = graph
        if fuse_bn_stats is None:
            fuse_bn_stats = os.environ.get("MCN_FUSE_BN_STATS", "1") != "0"
        self.fuse_bn_stats = bool(fuse_bn_stats)
        self.world = int(world_size)
        self.keep = set(keep)            # tensors that must stay materialised (parity taps)
        # parity aid: gradient buffers are never recycled and their addresses are recorded, so a
        # test can read the gradient of ANY tensor after a step (layer-wise backward checks)
        self.keep_grads = bool(keep_grads)
        self.grad_ptr = {}
        self.conv_mode = conv_mode       # 2 = halo tiles where they pay off else im2col TMA, 1 = im2col, 0 = box
        self.fetch_pred = fetch_pred
        self.sync_bn = sync_bn and self.world > 1
        self.loss_scale = loss_scale
        self.cdt = graph.compute_dtype
        self.csz = DT_SIZE[self.cdt]
        self.ccode = DT_CODE[self.cdt]
        self.bufs = []
        self.fwd = []
        self.bwd = []
        self.inf = []                    # inference (is_train=False) launches: EMA weights, BN moving stats
        self.phase = "train"
        self._node_bufs = {}             # (node id, key) -> Buf: per-PLAN cache (a graph may be re-planned)
        self.tbuf = {}                   # Tensor -> Ptr
        self.conv_descs = {}
        self.bn_layers = []
        self.allreduce_points = []       # (phase, launch index, dst Ptr, nbytes, dtype, (src0, n0, src1, n1) | None = in place)
        self.temp = TempAllocator()
        self.temp_buf = Buf("temp", 0, "temp")
        self.bufs.append(self.temp_buf)
        self.loss_slots = {}
        self._layout_vars()
        self._fuse()
        self._emit_forward()
        self._emit_backward()
        self._emit_inference()
        self.temp_buf.nbytes = self.temp.peak
        self._assign_offsets()

    # ------------------------------------------------------------------ buffers
    def node_buf(self, node, key, name, nbytes, region="act"):
        """Buffer owned by a node, created once per plan (training and inference lists share it)."""
        k = (node.id, key)
        if k not in self._node_bufs:
            self._node_bufs[k] = self.new_buf(name, nbytes, region)
        return self._node_bufs[k]

    def new_buf(self, name, nbytes, region="act"):
        b = Buf(name, max(int(nbytes), 16), region)
        self.bufs.append(b)
        return b

    def _assign_offsets(self):
        order = ["param", "zero", "state", "bf16", "act", "temp"]
        off = 0
        self.region_span = {}
        for region in order:
            start = off
            for b in self.bufs:
                if b.region == region:
                    b.offset = off
                    off += _align(b.nbytes)
            self.region_span[region] = (start, off)
        self.arena_bytes = off

    # ------------------------------------------------------------------ variables
    def _layout_vars(self):
        """All trainable variables contiguous (grads mirror them 1:1 -> flat buckets for the
        gradient all-reduce), then the non-trainable ones.  Conv weights whose Cin is not a
        multiple of 8 (RGB stems) are stored as a zero-padded [kpad, Cout] im2col matrix."""
        g = self.graph
        self._mark_operand_copies()
        tr = [v for v in g.vars.values() if v.trainable]
        nt = [v for v in g.vars.values() if not v.trainable]
        self.var_off = {}
        off = 0
        for v in tr + nt:
            self.var_off[v] = off
            off += _align(v.storage_size * 4, 64) // 4   # keep 64-byte alignment per tensor
        self.n_param = off
        self.n_train = (self.var_off[nt[0]] if nt else off) if tr else 0
        self.trainable = tr
        self.all_vars = tr + nt
        self.b_param = self.new_buf("params_f32", self.n_param * 4, "param")
        self.b_ema = self.new_buf("ema_f32", self.n_param * 4, "param")
        self.b_grad = self.new_buf("grads_f32", max(self.n_train, 1) * 4, "zero")
        self.b_mom = self.new_buf("opt_m", max(self.n_train, 1) * 4, "state")
        self.b_v = self.new_buf("opt_v", max(self.n_train, 1) * 4, "state")
        self.b_hp = self.new_buf("hyper_params", 64, "state")
        self.b_gnorm = self.new_buf("grad_sqnorm", 32, "zero")      # xsum accumulator: squared global gradient norm (clipping)
        # bf16 operand copies
        self.bf16_off = {}
        self.bf16t_off = {}
        n = 0
        for v in self.all_vars:
            if v.needs_bf16:
                self.bf16_off[v] = n
                n += _align(v.storage_size * 2, 256)
            if v.needs_bf16_t:
                self.bf16t_off[v] = n
                n += _align(v.storage_size * 2, 256)
        self.b_bf16 = self.new_buf("weights_bf16", n, "bf16")
        self.b_bf16_ema = self.new_buf("weights_bf16_ema", n, "bf16")

    def _mark_operand_copies(self):
        for node in self.graph.nodes:
            if node.op in ("conv2d", "dense", "conv2d_transpose") and self.cdt == "bf16":
                w = node.vars["w"]
                route = self._conv_route(node)
                node.attrs["route"] = route
                if route in ("tc", "mixed"):
                    w.needs_bf16 = w.needs_bf16_t = True
                    if node.op == "dense":
                        w.gemm_dims = (1, w.shape[0], w.shape[1])
                    else:
                        kh, kw, ci, co = w.shape
                        w.gemm_dims = (kh * kw, ci, co)
                elif route == "im2col":
                    kh, kw, ci, co = w.shape
                    kpad = _align(kh * kw * ci, 8)
                    w.storage_shape = (kpad, co)
                    w.storage_rows = None
                    w.gemm_dims = (1, kpad, co)
                    w.needs_bf16 = w.needs_bf16_t = True
                    node.attrs["kpad"] = kpad
                elif route == "stem":
                    kh, kw, ci, co = w.shape
                    kpad, rows = self._stem_geometry(node)
                    w.storage_shape = (kpad, co)
                    w.storage_rows = rows           # storage row of every reference [r, s, c]
                    w.gemm_dims = (1, kpad, co)
                    w.needs_bf16_t = True           # [Cout][Kpad]: the fprop operand
                    node.attrs["kpad"] = kpad
            elif node.op in ("conv2d", "dense", "conv2d_transpose"):
                node.attrs["route"] = "direct"

    def _conv_route(self, node):
        """tc: TMA/tcgen05 implicit GEMM; im2col: explicit im2col + tc GEMM (Cin % 8 != 0, fprop
        and wgrad only); direct: CUDA-core kernel."""
        w = node.vars["w"]
        if node.op == "dense":
            ci, co = w.shape
            return "tc" if ci % 8 == 0 and co % 8 == 0 else "direct"
        kh, kw, ci, co = w.shape
        if node.op == "conv2d_transpose":
            # runs as dgrad: GEMM K = input channels (ci), N = output channels (co, masked if odd);
            # "mixed": forward on tcgen05, backward on the CUDA-core kernels (co % 8 != 0: RGB out)
            if ci % 8 == 0 and kh * kw <= 52:
                return "tc" if co % 8 == 0 else "mixed"
            return "direct"
        if ci % 8 == 0 and co % 8 == 0 and kh * kw <= 52:
            return "tc"
        if node.attrs.get("ws"):
            return "direct"      # the stem / im2col routes store the weight permuted and padded
        if self._stem_geometry(node) is not None:
            return "stem"
        if co % 8 == 0 and kh * kw <= 52:
            return "im2col"
        return "direct"

    @staticmethod
    def stem_kpad(kh, kw):
        """Rows of the stem weight storage [Kpad][Cout]: k = (r*(kw+1) + s)*4 + c, padded to a
        multiple of 128 (mirrors mcn_stem_conv_kpad in csrc/conv_tc.cu)."""
        epr = (kw + 1) * 4
        rpc = 64 // epr
        kc = (kh + rpc - 1) // rpc
        return (kc + 1) // 2 * 2 * 64

    def _stem_geometry(self, node):
        """RGB stems the gather kernels handle (mcn_stem_conv_fprop): 3x3 / 7x7, stride 2 along W,
        unit dilation, even image width and left pad, Cout % 64 == 0, fed by the network input
        (no gradient into the image).  Returns (kpad, row index of every [r, s, c]) or None."""
        if os.environ.get("MCN_STEM_GATHER", "1") == "0" or node.op != "conv2d":
            return None
        w = node.vars["w"]
        kh, kw, ci, co = w.shape
        a = node.attrs
        x = node.inputs[0]
        if ci != 3 or (kh, kw) not in ((3, 3), (7, 7)) or co % 64 != 0 or co > 256:
            return None
        if a["s"][1] != 2 or tuple(a["d"]) != (1, 1) or a["pad"][1] % 2 or x.shape[2] % 2:
            return None
        if x.node is not None and x.node.op not in ("input", "input_prep"):
            return None
        if x.size % 4:       # mcn_pad_rgb4 works on whole quads of pixels plus a scalar tail: any size is fine
            pass
        rows = np.array([(r * (kw + 1) + s2) * 4 + c for r in range(kh) for s2 in range(kw) for c in range(ci)],
                        dtype=np.int64)
        return self.stem_kpad(kh, kw), rows

    def _needs_input_grad(self, node):
        return self._tensor_needs_grad(node.inputs[0])

    def _var_trains(self, v):
        """Is v updated by the backward pass being emitted?  (GANs run one pass per loss, each
        with its own variable subset, optimizers_gan.py:56-58.)"""
        ps = self.__dict__.get("cur_pass")
        return v.trainable and (ps is None or ps["train"] is None or v in ps["train"])

    def _tensor_needs_grad(self, t):
        cache = self.__dict__.setdefault("_ng_cache", {})
        if t in cache:
            return cache[t]
        n = t.node
        ps = self.__dict__.get("cur_pass")
        if n is None or n.op in ("input", "stop_gradient") or (ps is not None and t in ps["stop"]):
            r = False
        else:
            r = any(self._var_trains(v) for v in n.vars.values()) or \
                any(self._tensor_needs_grad(i) for i in n.inputs)
        cache[t] = r
        return r

    def pvar(self, v):
        # inference reads the EMA shadows of every variable (tf.cond(is_train, v, v_ema),
        # reference convnet.py:1406,1872-1876)
        return Ptr(self.b_ema if self.phase == "infer" else self.b_param, self.var_off[v] * 4)

    def pgrad(self, v):
        return Ptr(self.b_grad, self.var_off[v] * 4)

    def pbf16(self, v):
        return Ptr(self.b_bf16_ema if self.phase == "infer" else self.b_bf16, self.bf16_off[v])

    def pbf16t(self, v):
        return Ptr(self.b_bf16_ema if self.phase == "infer" else self.b_bf16, self.bf16t_off[v])

    def _bnred_target(self, conv):
        """The batch-norm node whose backward reduction can ride on this conv's dgrad epilogue
        (mcn_conv2d_dgrad_tc_bnred), or None.  Conditions: the conv's input is the final output of a
        BN(+ReLU) layer with this conv as its only consumer (so this dgrad IS the BN output's whole
        gradient and is written, not accumulated), training statistics, no fused residual, bf16, and a
        geometry the TMA-store epilogue covers (mirrors mcn_conv2d_dgrad_bnred_supported)."""
        if self.cdt != "bf16" or os.environ.get("MCN_FUSE_BN_BWD", "1") == "0" \
                or os.environ.get("MCN_TMA_STORE", "1") == "0":
            return None
        x = conv.inputs[0]
        prod = x.node
        if prod is None or x in self.keep or len(x.consumers) != 1 or x in self.g or self.keep_grads:
            return None
        bn = prod.attrs.get("fused_into") if prod.op == "act" else prod
        if bn is None or bn.op != "bn" or bn.attrs.get("final") is not x or bn.attrs.get("residual") is not None:
            return None
        if not bn.attrs.get("update", True) or bn.attrs.get("act", 0) not in (0, 1) or "save" not in bn.attrs:
            return None
        if not self._tensor_needs_grad(x):
            return None
        a = conv.attrs
        kh, kw, ci, co = conv.vars["w"].shape
        if tuple(a["s"]) != (1, 1) or ci % 64 or co % 8 or bn.inputs[0].shape != x.shape:
            return None
        if x.size // ci >= 2 ** 31:
            return None
        pointwise = kh == 1 and kw == 1
        return bn if (pointwise or (self.conv_mode >= 1 and co % 64 == 0) or self._halo_ok(conv)) else None

    def _halo_ok(self, conv):
        """halo_eligible() of csrc/conv_tc.cu for the dgrad of `conv` (dy is the halo-fed tensor)."""
        if self.conv_mode != 2:
            return False
        kh, kw, ci, co = conv.vars["w"].shape
        a = conv.attrs
        n, h, w_, _ = conv.inputs[0].shape
        if tuple(a["s"]) != (1, 1) or (kh == 1 and kw == 1) or co % 64 or kh * kw > 52:
            return False
        hwb, hhb = 8 + (kw - 1) * a["d"][1], 16 + (kh - 1) * a["d"][0]
        if hwb > 256 or hhb > 256 or hwb * hhb * 128 > 40 * 1024:
            return False
        eff = (h * w_) / (((h + 15) // 16 * 16) * ((w_ + 7) // 8 * 8))
        return eff >= float(os.environ.get("MCN_HALO_MIN_EFF", HALO_MIN_EFF)) \
            and h * w_ >= int(os.environ.get("MCN_HALO_MIN_HW", HALO_MIN_HW))

    # ------------------------------------------------------------------ weight standardisation
    # convnet.py:1410-1419: w' = (w - mean_o) / (std_o + 1e-5) per output channel, inside the graph.
    # A node whose weight is standardised computes w' (fp32 + the bf16 operand copies) into buffers
    # of its own at the start of its forward pass, runs on those, collects d(loss)/dw' in a buffer
    # of its own and folds it back onto the raw weight's gradient with mcn_ws_bwd.
    WS_EPS = 1e-5

    def _ws_dims(self, node):
        w = node.vars["w"]
        cols = w.shape[-1]
        return w.size // cols, cols

    def _ws_prepare(self, node):
        if not node.attrs.get("ws"):
            return
        w = node.vars["w"]
        assert tuple(w.storage_shape) == tuple(w.shape), "weight standardisation needs the plain storage layout"
        rows, cols = self._ws_dims(node)
        ws = {"f32": self.node_buf(node, "ws_f32", "ws_f32:%s" % node.scope, w.size * 4),
              "stats": self.node_buf(node, "ws_stats", "ws_stats:%s" % node.scope, 2 * cols * 4)}
        self.L("f", "mcn_ws_fwd", self.pvar(w), rows, cols, self.WS_EPS, Ptr(ws["f32"]), Ptr(ws["stats"]),
               tag=node.scope + "/ws")
        if w.needs_bf16 or w.needs_bf16_t:
            taps, ci, co = w.gemm_dims
            if w.needs_bf16:
                ws["bf16"] = self.node_buf(node, "ws_bf16", "ws_bf16:%s" % node.scope, w.size * 2)
            if w.needs_bf16_t:
                ws["bf16t"] = self.node_buf(node, "ws_bf16t", "ws_bf16t:%s" % node.scope, w.size * 2)
            self.L("f", "mcn_weight_prep", Ptr(ws["f32"]), taps, ci, co,
                   Ptr(ws["bf16"]) if "bf16" in ws else NULL, Ptr(ws["bf16t"]) if "bf16t" in ws else NULL,
                   tag=node.scope + "/ws_prep")
        node.attrs["ws_buf"] = ws

    def _w_f32(self, node):
        return Ptr(node.attrs["ws_buf"]["f32"]) if node.attrs.get("ws") else self.pvar(node.vars["w"])

    def _w_bf16(self, node):
        return Ptr(node.attrs["ws_buf"]["bf16"]) if node.attrs.get("ws") else self.pbf16(node.vars["w"])

    def _w_bf16t(self, node):
        return Ptr(node.attrs["ws_buf"]["bf16t"]) if node.attrs.get("ws") else self.pbf16t(node.vars["w"])

    def _w_grad(self, node):
        """Where the weight-gradient kernels accumulate: the variable's gradient, or (standardised
        weight) a zeroed buffer of the node's own that _ws_finish folds back."""
        if not node.attrs.get("ws"):
            return self.pgrad(node.vars["w"])
        w = node.vars["w"]
        buf = self.node_buf(node, "ws_grad", "ws_grad:%s" % node.scope, w.size * 4)
        node.attrs["ws_buf"]["grad"] = buf
        self.L("b", "mcn_fill_f32", Ptr(buf), w.size, 0.0, tag="zero")
        return Ptr(buf)

    def _ws_finish(self, node):
        ws = node.attrs.get("ws_buf") if node.attrs.get("ws") else None
        if not ws or "grad" not in ws:
            return
        rows, cols = self._ws_dims(node)
        w = node.vars["w"]
        self.L("b", "mcn_ws_bwd", Ptr(ws["grad"]), self.pvar(w), Ptr(ws["stats"]), rows, cols, self.WS_EPS,
               self.pgrad(w), tag=node.scope + "/ws_bwd")
        del ws["grad"]

    # ------------------------------------------------------------------ fusion
    def _single_consumer(self, t):
        return t.consumers[0] if len(t.consumers) == 1 and t not in self.keep else None

    def _fuse(self):
        for node in self.graph.nodes:
            node.attrs["fused_into"] = None          # a graph may be planned more than once
            node.attrs["bn_stats_node"] = None
        for node in self.graph.nodes:
            if node.op == "bn":
                # conv -> BN: the statistics are taken in the tensor-core conv's epilogue
                prod = node.inputs[0].node
                node.attrs["stats_in_conv"] = False
                if (self.fuse_bn_stats and self.cdt == "bf16" and prod is not None
                        and node.attrs["update"]
                        and prod.op == "conv2d" and prod.attrs.get("route") in ("tc", "im2col", "stem")
                        and prod.attrs["bn_stats_node"] is None
                        and node.inputs[0].shape[-1] % 64 == 0
                        and self._stats_fusion_pays(prod)):
                    prod.attrs["bn_stats_node"] = node
                    node.attrs["stats_in_conv"] = True
                node.attrs["act"], node.attrs["alpha"] = 0, 0.0
                node.attrs["residual"] = None
                node.attrs["final"] = node.outputs[0]
                cur = node.outputs[0]
                c = self._single_consumer(cur)
                if c is not None and c.op == "add":
                    other = c.inputs[1] if c.inputs[0] is cur else c.inputs[0]
                    # the residual must already exist when this BN runs
                    if other.node is not None and other.node.id < node.id and other is not cur:
                        node.attrs["residual"] = other
                        c.attrs["fused_into"] = node
                        cur = c.outputs[0]
                        node.attrs["final"] = cur
                        c = self._single_consumer(cur)
                if c is not None and c.op == "act":
                    # with a fused residual the derivative must come from the output: relu family
                    if node.attrs["residual"] is None or c.attrs["act"] in (1, 2, 3):
                        node.attrs["act"] = c.attrs["act"]
                        node.attrs["alpha"] = c.attrs["alpha"]
                        c.attrs["fused_into"] = node
                        node.attrs["final"] = c.outputs[0]
            elif node.op == "dense":
                # dense -> cast(f32): the GEMM epilogue writes fp32 logits.  A softmax consumer
                # of the pre-cast logits (d['pred']) reads the fp32 values instead.
                node.attrs["final"] = node.outputs[0]
                out = node.outputs[0]
                cons = [c for c in out.consumers if c.op != "softmax"]
                if len(cons) == 1 and cons[0].op == "cast" and out not in self.keep \
                        and node.attrs.get("route") == "tc":
                    cons[0].attrs["fused_into"] = node
                    node.attrs["final"] = cons[0].outputs[0]
            elif node.op in ("add", "sd_add"):
                node.attrs["act"] = 0
                node.attrs["alpha"] = 0.0
                node.attrs["final"] = node.outputs[0]
        # add -> act for adds that were not absorbed by a BN
        for node in self.graph.nodes:
            if node.op in ("add", "sd_add") and node.attrs["fused_into"] is None:
                c = self._single_consumer(node.outputs[0])
                if c is not None and c.op == "act" and c.attrs["act"] in (1, 2, 3) \
                        and c.attrs["fused_into"] is None:
                    node.attrs["act"] = c.attrs["act"]
                    node.attrs["alpha"] = c.attrs["alpha"]
                    c.attrs["fused_into"] = node
                    node.attrs["final"] = c.outputs[0]

    # ------------------------------------------------------------------ helpers
    def tensor_ptr(self, t):
        return self.tbuf[t]

    def alloc_act(self, t, dtype=None):
        if t in self.tbuf:              # inference reuses the training forward's buffers
            return self.tbuf[t]
        dt = dtype or t.dtype
        b = self.new_buf("act:%s" % t.name, t.size * DT_SIZE[dt], "act")
        p = Ptr(b)
        self.tbuf[t] = p
        return p

    def conv_desc(self, **kw):
        d = ConvDesc(**kw)
        return self.conv_descs.setdefault(d.key(), d)

    def L(self, phase, fn, *args, **kw):
        if self.phase == "infer":
            self.inf.append(Launch(fn, args, kw.get("tag", "")))
            return
        (self.fwd if phase == "f" else self.bwd).append(Launch(fn, args, kw.get("tag", "")))

    def talloc(self, nbytes):
        off, size = self.temp.alloc(nbytes)
        return Ptr(self.temp_buf, off), (off, size)

    def tfree(self, handle):
        if os.environ.get("MCN_NO_TEMP_REUSE"):   # debugging aid: never recycle temporaries
            return
        self.temp.release(*handle)

    # ------------------------------------------------------------------ forward emission
    def _emit_inference(self):
        """Forward-only launch list for is_train=False (reference predict(), convnet.py:609-665)."""
        self.phase = "infer"
        self._emit_forward()
        self.phase = "train"

    def _emit_forward(self):
        g = self.graph
        # scratch that must be zero at step start: BN fp64 sums, loss accumulators
        for node in g.nodes:
            if node.attrs.get("fused_into") is not None:
                continue
            getattr(self, "_f_" + node.op)(node)
        for node in g.nodes:
            if node.op == "softmax" and node.attrs.get("standalone"):
                x = node.inputs[0]
                src = self.tbuf[x]
                c = x.shape[-1]
                if self._logits_dtype(x) != "f32":
                    # use the fp32 copy the loss reads when there is one, else make one
                    f32 = [cn.outputs[0] for cn in x.consumers
                           if cn.op == "cast" and cn.outputs[0].dtype == "f32" and cn.outputs[0] in self.tbuf]
                    if f32:
                        src = self.tbuf[f32[0]]
                    else:
                        b = self.node_buf(node, "f32_copy", "softmax_in_f32", x.size * 4)
                        src = Ptr(b)
                        self.L("f", "mcn_cast", DT_CODE[x.dtype], self.tbuf[x], 0, src, x.size, tag="softmax_cast")
                self.L("f", "mcn_softmax_xent", src, NULL, x.size // c, c, NULL, 0.0, 0.0, 0.0, 0, 0, 0.0,
                       NULL, NULL, self.tbuf[node.outputs[0]], tag="softmax")

    def _logits_dtype(self, t):
        """dtype of the data actually stored for t (a fused dense writes fp32 under a bf16 name)."""
        n = t.node
        if n is not None and n.op == "dense" and n.attrs.get("final") is not t:
            return n.attrs["final"].dtype
        if n is not None and n.op == "softmax":
            return "f32"          # probabilities are always written in fp32 (mcn_softmax_xent)
        return t.dtype

    def _f_input(self, node):
        t = node.outputs[0]
        self.alloc_act(t)

    def _f_input_prep(self, node):
        x, y = node.inputs[0], node.outputs[0]
        py = self.alloc_act(y)
        n, hi, wi, c = x.shape
        _, h, w, _ = y.shape
        self.L("f", "mcn_input_prep", self.tbuf[x], DT_CODE[x.dtype], n, hi, wi, h, w, c, node.attrs["mean"],
               node.attrs["scale"], DT_CODE[y.dtype], py, tag="input_prep")

    def _conv_geometry(self, node):
        x, y = node.inputs[0], node.outputs[0]
        a = node.attrs
        n, h, w, ci = x.shape
        _, ho, wo, co = y.shape
        return self.conv_desc(N=n, H=h, W=w, Cin=ci, Cout=co, kh=a["k"][0], kw=a["k"][1],
                              sh=a["s"][0], sw=a["s"][1], dh=a["d"][0], dw=a["d"][1],
                              pad_t=a["pad"][0], pad_l=a["pad"][1], Ho=ho, Wo=wo)

    def _f_conv2d(self, node):
        x, y = node.inputs[0], node.outputs[0]
        w = node.vars["w"]
        self._ws_prepare(node)
        b = node.vars.get("b")
        d = self._conv_geometry(node)
        node.attrs["desc"] = d
        py = self.alloc_act(y)
        route = node.attrs["route"]
        pb = self.pvar(b) if b is not None else NULL
        bn = node.attrs.get("bn_stats_node") if self.phase == "train" else None
        psums = Ptr(self._bn_sums_buf(bn)) if bn is not None else None
        if route == "tc":
            if bn is not None:
                self.L("f", "mcn_conv2d_fprop_tc_stats", d, self.tbuf[x], self._w_bf16t(node), pb, py,
                       self.conv_mode, psums, tag=node.scope)
            else:
                self.L("f", "mcn_conv2d_fprop_tc", d, self.tbuf[x], self._w_bf16t(node), pb, py, self.ccode,
                       self.conv_mode, 0, tag=node.scope)
        elif route == "stem":
            # 4-channel copy of the image, then the gather convolution (no im2col matrix)
            x4 = self.node_buf(node, "x4", "rgb4:%s" % node.scope, d.N * d.H * d.W * 4 * 2)
            node.attrs["x4"] = x4
            d4 = self.conv_desc(**dict({f: getattr(d, f) for f in ConvDesc.FIELDS}, Cin=4))
            node.attrs["desc4"] = d4
            self.L("f", "mcn_pad_rgb4", self.tbuf[x], d.N * d.H * d.W, Ptr(x4), tag=node.scope + "/rgb4")
            self.L("f", "mcn_stem_conv_fprop", d4, Ptr(x4), self._w_bf16t(node), pb, py,
                   psums if bn is not None else NULL, tag=node.scope)
        elif route == "im2col":
            kpad = node.attrs["kpad"]
            m = d.N * d.Ho * d.Wo
            col = self.node_buf(node, "col", "im2col:%s" % node.scope, m * kpad * 2)
            node.attrs["col"] = col
            gd = self.conv_desc(N=1, H=1, W=m, Cin=kpad, Cout=d.Cout, kh=1, kw=1, sh=1, sw=1, dh=1,
                                dw=1, pad_t=0, pad_l=0, Ho=1, Wo=m)
            node.attrs["gemm_desc"] = gd
            self.L("f", "mcn_im2col", d, DT_CODE[x.dtype], self.tbuf[x], Ptr(col), kpad,
                   tag=node.scope + "/im2col")
            if bn is not None:
                self.L("f", "mcn_conv2d_fprop_tc_stats", gd, Ptr(col), self._w_bf16t(node), pb, py, 0, psums,
                       tag=node.scope)
            else:
                self.L("f", "mcn_conv2d_fprop_tc", gd, Ptr(col), self._w_bf16t(node), pb, py, self.ccode, 0, 0,
                       tag=node.scope)
        else:
            self.L("f", "mcn_conv2d_fprop_direct", d, self.ccode, self.tbuf[x], 0, self._w_f32(node), pb, py,
                   tag=node.scope)            # fp32 master (or standardised) weights

    @staticmethod
    def _stats_fusion_pays(conv):
        """The epilogue statistics cost ~0.5-1.4k cycles per 128x64 output chunk.  That hides behind
        the MMAs when the reduction is deep (K = kh*kw*Cin) and shows when the conv is epilogue /
        store bound (1x1 expansions with small K).  Rule fitted on the per-layer A/B profile of
        ResNet-50 at batch 256 (profiles/r01_fused_stats_ab.txt)."""
        kh, kw, ci, co = conv.vars["w"].shape
        k = kh * kw * ci
        # Every eligible conv since the statistics ride on the TMA-store staging tile and the
        # intermediate flushes take no ticket (ResNet-50 b256: 19.66 -> 19.34 ms per step; with the
        # round-1 epilogue only K >= 512 or (K >= 128 and Cout <= 64) paid: MCN_FUSE_STATS_RULE=k).
        mode = os.environ.get("MCN_FUSE_STATS_RULE", "all")
        if mode == "all":
            return True
        return k >= 512 or (k >= 128 and co <= 64)

    def _bn_sums_buf(self, bn_node):
        c = bn_node.inputs[0].shape[-1]
        return self.node_buf(bn_node, "sums", "bn_sums:%s" % bn_node.scope, 2 * c * 8, "zero")

    def _f_dwconv2d(self, node):
        x, y = node.inputs[0], node.outputs[0]
        w = node.vars["w"]
        self._ws_prepare(node)
        d = self._conv_geometry(node)
        # depthwise: desc.Cout is unused by the kernels; keep the input channel count
        d = self.conv_desc(**dict({f: getattr(d, f) for f in ConvDesc.FIELDS}, Cout=d.Cin))
        node.attrs["desc"] = d
        py = self.alloc_act(y)
        self.L("f", "mcn_dwconv2d_fwd", d, node.attrs["mult"], self.ccode, self.tbuf[x], 0,
               self._w_f32(node), py, tag=node.scope)
        if "b" in node.vars:
            self.L("f", "mcn_bias_add", self.ccode, py, y.size // y.shape[-1], y.shape[-1],
                   self.pvar(node.vars["b"]), tag=node.scope + "/bias")

    def _f_conv2d_transpose(self, node):
        # forward of conv2d_transpose == dgrad of the conv mapping y-shaped -> x-shaped tensors
        x, y = node.inputs[0], node.outputs[0]
        w = node.vars["w"]   # stored [kh,kw,Cin_t,Cout_t] (reference convnet.py:2460)
        self._ws_prepare(node)
        a = node.attrs
        n, h, wd, ci = x.shape
        _, ho, wo, co = y.shape
        # underlying conv: input = y (co channels), output = x (ci channels); its HWIO weight is
        # [kh,kw,co,ci] = the stored weight with the last two axes swapped, i.e. the stored
        # layout read as O-H-W-I... handled by passing the "transposed" bf16 copy to dgrad.
        d = self.conv_desc(N=n, H=ho, W=wo, Cin=co, Cout=ci, kh=a["k"][0], kw=a["k"][1],
                           sh=a["s"][0], sw=a["s"][1], dh=a["d"][0], dw=a["d"][1],
                           pad_t=a["pad"][0], pad_l=a["pad"][1], Ho=h, Wo=wd)
        node.attrs["desc"] = d
        py = self.alloc_act(y)
        if node.attrs["route"] in ("tc", "mixed"):
            # dgrad wants W_conv[tap][Cin_conv=co][Cout_conv=ci] = stored[tap][ci][co]^T = bf16_t copy
            self.L("f", "mcn_conv2d_dgrad_tc", d, self.tbuf[x], self._w_bf16t(node), py, self.ccode,
                   self.conv_mode, 0, tag=node.scope)
        else:
            # exact / odd-channel path on CUDA cores: the underlying conv's HWIO weight
            # [tap][co][ci] is rebuilt from the stored [tap][ci][co] master every step (tiny)
            kh, kw, ci_t, co_t = w.shape
            wT = self.node_buf(node, "wT", "tconv_wT:%s" % node.scope, w.size * 4)
            node.attrs["wT"] = wT
            self.L("f", "mcn_fill_f32", Ptr(wT), w.size, 0.0, tag="zero")
            self.L("f", "mcn_transpose_add_f32", self._w_f32(node), kh * kw, ci_t, co_t, Ptr(wT),
                   tag=node.scope + "/wT")
            self.L("f", "mcn_conv2d_dgrad_direct", d, self.ccode, self.tbuf[x], 0, Ptr(wT), py,
                   tag=node.scope)
        if "b" in node.vars:
            self.L("f", "mcn_bias_add", self.ccode, py, y.size // co, co, self.pvar(node.vars["b"]),
                   tag=node.scope + "/bias")

    def _f_dense(self, node):
        x = node.inputs[0]
        y = node.attrs["final"]
        w = node.vars["w"]
        self._ws_prepare(node)
        b = node.vars.get("b")
        n, ci = x.shape
        co = w.shape[1]
        d = self.conv_desc(N=1, H=1, W=n, Cin=ci, Cout=co, kh=1, kw=1, sh=1, sw=1, dh=1, dw=1,
                           pad_t=0, pad_l=0, Ho=1, Wo=n)
        node.attrs["desc"] = d
        py = self.alloc_act(y)
        if y is not node.outputs[0]:
            self.tbuf[node.outputs[0]] = py
        pb = self.pvar(b) if b is not None else NULL
        if node.attrs["route"] == "tc":
            self.L("f", "mcn_conv2d_fprop_tc", d, self.tbuf[x], self._w_bf16t(node), pb, py,
                   DT_CODE[y.dtype], 0, 0, tag=node.scope)
        else:
            self.L("f", "mcn_conv2d_fprop_direct", d, self.ccode, self.tbuf[x], 0, self._w_f32(node), pb,
                   py, tag=node.scope)

    def _f_cast(self, node):
        x, y = node.inputs[0], node.outputs[0]
        if x.dtype == y.dtype:
            self.tbuf[y] = self.tbuf[x]
            return
        py = self.alloc_act(y)
        self.L("f", "mcn_cast", DT_CODE[x.dtype], self.tbuf[x], DT_CODE[y.dtype], py, x.size,
               tag="cast")

    def _f_bn(self, node):
        x = node.inputs[0]
        y = node.attrs["final"]
        c = x.shape[-1]
        rows = x.size // c
        v = node.vars
        if self.phase == "infer":
            py = self.alloc_act(y)
            for o in node.outputs:
                self.tbuf.setdefault(o, py)
            res = node.attrs["residual"]
            self.L("f", "mcn_bn_infer", self.ccode, self.tbuf[x], rows, c, self.pvar(v["mu"]),
                   self.pvar(v["sigma"]), node.attrs["eps"],
                   self.pvar(v["gamma"]) if "gamma" in v else NULL,
                   self.pvar(v["beta"]) if "beta" in v else NULL,
                   self.tbuf[res] if res is not None else NULL, node.attrs["act"], node.attrs["alpha"], py,
                   tag=node.scope + "/infer")
            return
        sums = self._bn_sums_buf(node)
        save = self.new_buf("bn_save:%s" % node.scope, 2 * c * 4, "state")
        node.attrs["save"] = save
        node.attrs["rows"] = rows
        py = self.alloc_act(y)
        for o in node.outputs:
            self.tbuf.setdefault(o, py)
        pg = self.pvar(v["gamma"]) if "gamma" in v else NULL
        pbeta = self.pvar(v["beta"]) if "beta" in v else NULL
        res = node.attrs["residual"]
        pres = self.tbuf[res] if res is not None else NULL
        if not node.attrs["update"]:
            # frozen layer: normalise with the stored moving statistics (reference
            # convnet.py:1916-1924, fused_batch_norm(is_training=False) in the training graph)
            self.L("f", "mcn_bn_frozen_stats", self.pvar(v["mu"]), self.pvar(v["sigma"]), c, node.attrs["eps"],
                   Ptr(save), Ptr(save, c * 4), tag=node.scope + "/frozen_stats")
            self.L("f", "mcn_bn_apply", self.ccode, self.tbuf[x], rows, c, Ptr(save), Ptr(save, c * 4), pg,
                   pbeta, pres, node.attrs["act"], node.attrs["alpha"], py, tag=node.scope + "/apply")
            self.bn_layers.append(node)
            return
        if not node.attrs.get("stats_in_conv"):
            self.L("f", "mcn_bn_stats", self.ccode, self.tbuf[x], rows, c, Ptr(sums), tag=node.scope + "/stats")
        if self.sync_bn:
            self.allreduce_points.append(("f", len(self.fwd), Ptr(sums), 2 * c * 8, "f64", None))
        upd = node.attrs["update"]
        if os.environ.get("MCN_BN_FOLD_FINALIZE", "1") == "0":     # A/B switch: separate finalize launch
            self.L("f", "mcn_bn_finalize", Ptr(sums), float(rows * (self.world if self.sync_bn else 1)), c,
                   node.attrs["eps"], node.attrs["momentum"], Ptr(save), Ptr(save, c * 4),
                   self.pvar(v["mu"]) if upd else NULL, self.pvar(v["sigma"]) if upd else NULL,
                   tag=node.scope + "/finalize")
            self.L("f", "mcn_bn_apply", self.ccode, self.tbuf[x], rows, c, Ptr(save), Ptr(save, c * 4), pg,
                   pbeta, pres, node.attrs["act"], node.attrs["alpha"], py, tag=node.scope + "/apply")
            self.bn_layers.append(node)
            return
        # finalize (mean / invstd / moving statistics) happens in the apply kernel's prologue
        cv = c // 8
        if (self.cdt == "bf16" and res is not None and node.attrs["act"] == 1 and c % 8 == 0 and cv <= 256
                and os.environ.get("MCN_BN_RELU_MASK", "1") != "0"
                and os.environ.get("MCN_BN_PIPE", "1") != "0"):
            # residual fused: backward needs the sign of the OUTPUT; the apply pass leaves it as one bit
            # per element, so neither backward pass re-reads y (mcn_bn_bwd_*_mask)
            mask = self.node_buf(node, "relu_mask", "bn_relu_mask:%s" % node.scope, (rows * cv + 3) // 4 * 4)
            node.attrs["relu_mask"] = mask
            self.L("f", "mcn_bn_apply_stats_mask", self.ccode, self.tbuf[x], rows, c, Ptr(sums),
                   float(rows * (self.world if self.sync_bn else 1)), node.attrs["eps"], node.attrs["momentum"],
                   pg, pbeta, pres, node.attrs["act"], node.attrs["alpha"], py, Ptr(mask), Ptr(save),
                   Ptr(save, c * 4), self.pvar(v["mu"]) if upd else NULL, self.pvar(v["sigma"]) if upd else NULL,
                   tag=node.scope + "/apply")
            self.bn_layers.append(node)
            return
        self.L("f", "mcn_bn_apply_stats", self.ccode, self.tbuf[x], rows, c, Ptr(sums),
               float(rows * (self.world if self.sync_bn else 1)), node.attrs["eps"], node.attrs["momentum"],
               pg, pbeta, pres, node.attrs["act"], node.attrs["alpha"], py, Ptr(save), Ptr(save, c * 4),
               self.pvar(v["mu"]) if upd else NULL, self.pvar(v["sigma"]) if upd else NULL,
               tag=node.scope + "/apply")
        self.bn_layers.append(node)

    def _f_act(self, node):
        x, y = node.inputs[0], node.outputs[0]
        py = self.alloc_act(y)
        self.L("f", "mcn_act_fwd", DT_CODE[x.dtype], self.tbuf[x], x.size, node.attrs["act"],
               node.attrs["alpha"], py, tag="act")

    def _f_add(self, node):
        a, b = node.inputs
        y = node.attrs["final"]
        py = self.alloc_act(y)
        self.tbuf.setdefault(node.outputs[0], py)
        self.L("f", "mcn_add_act_fwd", DT_CODE[a.dtype], self.tbuf[a], self.tbuf[b], a.size,
               node.attrs["act"], node.attrs["alpha"], py, tag="add")

    def _f_sd_add(self, node):
        # stochastic depth (convnet.py:2500-2512): train y = act(a*survived[n] + b); inference a + b
        a, b = node.inputs
        y = node.attrs["final"]
        py = self.alloc_act(y)
        self.tbuf.setdefault(node.outputs[0], py)
        if self.phase == "infer":
            self.L("f", "mcn_add_act_fwd", DT_CODE[a.dtype], self.tbuf[a], self.tbuf[b], a.size,
                   node.attrs["act"], node.attrs["alpha"], py, tag="add")
            return
        n = a.shape[0]
        self.L("f", "mcn_sd_add_fwd", DT_CODE[a.dtype], self.tbuf[a], self.tbuf[b], n, a.size // n,
               node.attrs["rate"], Ptr(self.b_hp), node.attrs["layer"], node.attrs["act"],
               node.attrs["alpha"], py, tag=node.scope)

    def _b_sd_add(self, node, gy):
        a, b = node.inputs
        y = node.attrs["final"]
        esz = DT_SIZE[a.dtype]
        n = a.shape[0]
        need_a, need_b = self._tensor_needs_grad(a), self._tensor_needs_grad(b)
        pa, ha = self.talloc(a.size * esz) if need_a else (NULL, None)
        pb, hb = self.talloc(b.size * esz) if need_b else (NULL, None)
        if need_a or need_b:
            self.L("b", "mcn_sd_add_bwd", DT_CODE[a.dtype], gy, self.tbuf[y], n, a.size // n,
                   node.attrs["rate"], Ptr(self.b_hp), node.attrs["layer"], node.attrs["act"],
                   node.attrs["alpha"], pa, pb, tag=node.scope + "/bwd")
        for t, p, h in ((a, pa, ha), (b, pb, hb)):
            if h is None:
                continue
            if t not in self.g:
                self.g[t] = (p, h)          # first contribution: hand the buffer over
            else:
                self.L("b", "mcn_accumulate", DT_CODE[t.dtype], self.g[t][0], p, t.size, tag="grad_accumulate")
                self.tfree(h)

    def _f_dropout(self, node):
        x, y = node.inputs[0], node.outputs[0]
        py = self.alloc_act(y)
        if self.phase == "infer":      # is_train False -> rate 0: identity
            self.L("f", "mcn_cast", DT_CODE[x.dtype], self.tbuf[x], DT_CODE[x.dtype], py, x.size, tag="dropout_off")
            return
        self.L("f", "mcn_dropout", DT_CODE[x.dtype], self.tbuf[x], x.size, node.attrs["rate"], Ptr(self.b_hp),
               node.attrs["layer"], py, tag="dropout")

    def _b_dropout(self, node, gy):
        x = node.inputs[0]
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_dropout", DT_CODE[x.dtype], gy, x.size, node.attrs["rate"],
                                         Ptr(self.b_hp), node.attrs["layer"], p, tag="dropout_bwd"))

    def _f_resize_nearest(self, node):
        x, y = node.inputs[0], node.outputs[0]
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        py = self.alloc_act(y)
        self.L("f", "mcn_resize_nearest_fwd", DT_CODE[x.dtype], self.tbuf[x], n, h, w, c, ho, wo,
               node.attrs["mode"], py, tag="resize_nearest")

    def _b_resize_nearest(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_resize_nearest_bwd", DT_CODE[x.dtype], gy, n, h, w, c,
                                         ho, wo, node.attrs["mode"], p, tag="resize_nearest_bwd"))

    def _f_scale_bcast(self, node):
        x, m = node.inputs
        y = node.outputs[0]
        n, h, w, c = x.shape
        py = self.alloc_act(y)
        self.L("f", "mcn_scale_bcast_fwd", DT_CODE[x.dtype], self.tbuf[x], self.tbuf[m], n, h * w, c,
               py, tag="se_scale")

    def _f_max_pool(self, node):
        x, y = node.inputs[0], node.outputs[0]
        a = node.attrs
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        py = self.alloc_act(y)
        # compact form: the winning tap of the window as one byte (the int32 TF argmax is twice
        # the size of a bf16 tensor); expanded on demand by mcn_maxpool_tap_to_argmax
        vec = 16 // DT_SIZE[x.dtype]
        a["tap_form"] = (c % vec == 0 and a["k"][0] * a["k"][1] <= 255 and x.size // 4 < 2 ** 31)
        if a["tap_form"]:
            arg = self.node_buf(node, "tap", "pool_tap:%s" % node.scope, y.size)
            node.attrs["argmax"] = arg
            self.L("f", "mcn_maxpool_fwd_tap", DT_CODE[x.dtype], self.tbuf[x], n, h, w, c, a["k"][0], a["k"][1],
                   a["s"][0], a["s"][1], a["pad"][0], a["pad"][1], ho, wo, py, Ptr(arg), tag="max_pool")
            return
        arg = self.node_buf(node, "argmax", "argmax:%s" % node.scope, y.size * 4)
        node.attrs["argmax"] = arg
        self.L("f", "mcn_maxpool_fwd", DT_CODE[x.dtype], self.tbuf[x], n, h, w, c, a["k"][0], a["k"][1],
               a["s"][0], a["s"][1], a["pad"][0], a["pad"][1], ho, wo, py, Ptr(arg), tag="max_pool")

    def _f_avg_pool(self, node):
        x, y = node.inputs[0], node.outputs[0]
        a = node.attrs
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        py = self.alloc_act(y)
        self.L("f", "mcn_avgpool_fwd", DT_CODE[x.dtype], self.tbuf[x], n, h, w, c, a["k"][0], a["k"][1],
               a["s"][0], a["s"][1], a["pad"][0], a["pad"][1], ho, wo, py, tag="avg_pool")

    def _f_gap(self, node):
        x, y = node.inputs[0], node.outputs[0]
        n, h, w, c = x.shape
        py = self.alloc_act(y)
        self.L("f", "mcn_gap_fwd", DT_CODE[x.dtype], self.tbuf[x], n, h * w, c, py, DT_CODE[y.dtype],
               tag="gap")

    def _f_concat(self, node):
        y = node.outputs[0]
        py = self.alloc_act(y)
        ctot = y.shape[-1]
        rows = y.size // ctot
        off = 0
        for t in node.inputs:
            c = t.shape[-1]
            self.L("f", "mcn_copy_channels", DT_CODE[t.dtype], self.tbuf[t], rows, c, 0, py, ctot, off,
                   c, 0, tag="concat")
            off += c

    def _f_reshape(self, node):
        self.tbuf[node.outputs[0]] = self.tbuf[node.inputs[0]]

    def _f_stop_gradient(self, node):
        self.tbuf[node.outputs[0]] = self.tbuf[node.inputs[0]]

    def _f_affine(self, node):
        raise NotImplementedError("scalar affine ops on tensors are not lowered yet")

    def _f_resize_bilinear(self, node):
        x, y = node.inputs[0], node.outputs[0]
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        py = self.alloc_act(y)
        self.L("f", "mcn_resize_bilinear_fwd", DT_CODE[x.dtype], self.tbuf[x], n, h, w, c, ho, wo,
               node.attrs["mode"], py, tag="resize")

    def _f_softmax(self, node):
        # probabilities come for free from the fused loss kernel when it exists
        x, y = node.inputs[0], node.outputs[0]
        if y not in self.tbuf:
            b = self.new_buf("probs", x.size * 4, "act")
            self.tbuf[y] = Ptr(b)
        node.attrs["standalone"] = True

    def _f_softmax_xent(self, node):
        logits, labels = node.inputs
        a = node.attrs
        c = logits.shape[-1]
        rows = a["rows"]
        if self.phase == "infer":
            return          # predictions come from the softmax node (emitted stand-alone below)
        # loss accumulators are exact fixed-point sums (xsum: 3 int64 limbs each, include/mcn.h)
        loss = self.new_buf("loss", 2 * 24, "zero")
        self.loss_slots["loss"] = Ptr(loss)
        self.loss_slots["l2"] = Ptr(loss, 24)
        self.tbuf[node.outputs[0]] = Ptr(loss)
        dlog = self.new_buf("dlogits", logits.size * 4, "act")
        node.attrs["dlogits"] = dlog
        probs = NULL
        src = logits.node.inputs[0] if logits.node is not None and logits.node.op == "cast" else logits
        for cns in list(logits.consumers) + list(src.consumers):
            if cns.op == "softmax" and self.fetch_pred:
                probs = self.tbuf.get(cns.outputs[0])
                if probs is None:
                    pb = self.new_buf("probs", logits.size * 4, "act")
                    probs = Ptr(pb)
                    self.tbuf[cns.outputs[0]] = probs
                cns.attrs["standalone"] = False
        cw = NULL
        if a["class_weights"] is not None:
            cwb = self.new_buf("class_weights", c * 4, "state")
            node.attrs["cw_buf"] = cwb
            cw = Ptr(cwb)
        # loss = mean over rows (convnet.py:594); gradient seeded with loss_scale / rows
        seg_h, seg_w = a.get("seg_hw") or (0, 0)
        self.L("f", "mcn_softmax_xent", self.tbuf[logits], self.tbuf[labels], rows, c, cw,
               a["label_smoothing"], a.get("focal_gamma", 0.0), a.get("sigmoid_focal_alpha", 0.0),
               seg_h, seg_w, self.loss_scale / rows, Ptr(loss), Ptr(dlog), probs, tag="softmax_xent")

    def _f_gan_loss(self, node):
        if self.phase == "infer":
            return
        lr_, lf_ = node.inputs
        a = node.attrs
        n = a["rows"]
        loss = self.new_buf("gan_loss", 3 * 24, "zero")
        self.loss_slots["loss"] = Ptr(loss)
        self.loss_slots["loss_g"] = Ptr(loss, 24)
        self.loss_slots["l2"] = Ptr(loss, 48)
        self.tbuf[node.outputs[0]] = Ptr(loss)
        self.tbuf[node.outputs[1]] = Ptr(loss, 24)
        bufs = [self.new_buf("dlogits_%s" % k, n * 4, "act") for k in ("real", "fake_d", "fake_g")]
        a["dlogits"] = bufs
        w0, w1 = a["w"]
        ls = a["label_smoothing"]
        gs = self.loss_scale / n
        # loss_d = mean(w1*sigCE(1-ls, D(x)) + w0*sigCE(0, D(G(z)))); loss_g = mean(w0*sigCE(1-ls, D(G(z))))
        self.L("f", "mcn_sigmoid_xent", self.tbuf[lr_], n, 1.0 - ls, w1, gs, Ptr(loss), Ptr(bufs[0]), 0,
               tag="gan_loss/real")
        self.L("f", "mcn_sigmoid_xent", self.tbuf[lf_], n, 0.0, w0, gs, Ptr(loss), Ptr(bufs[1]), 0,
               tag="gan_loss/fake_d")
        self.L("f", "mcn_sigmoid_xent", self.tbuf[lf_], n, 1.0 - ls, w0, gs * a["generator_scaling_factor"],
               Ptr(loss, 24), Ptr(bufs[2]), 0, tag="gan_loss/fake_g")

    def _b_gan_loss(self, node):
        lr_, lf_ = node.inputs
        b = node.attrs["dlogits"]
        if self.cur_pass["id"] == 0:
            self.g[lr_] = (Ptr(b[0]), None)
            self.g[lf_] = (Ptr(b[1]), None)
        else:
            self.g[lf_] = (Ptr(b[2]), None)

    # ------------------------------------------------------------------ backward emission
    def _backward_passes(self):
        gl = [n for n in self.graph.nodes if n.op == "gan_loss"]
        if not gl:
            return [{"id": 0, "train": None, "stop": set()}]
        model_vars = list(self.graph.vars.values())
        node = gl[0]
        g_blocks = node.attrs["g_blocks"]
        vd = set(v for v in model_vars if v.block not in g_blocks)
        vg = set(v for v in model_vars if v.block in g_blocks)
        return [{"id": 0, "train": vd, "stop": {node.attrs["generate"]}},
                {"id": 1, "train": vg, "stop": set()}]

    def _emit_backward(self):
        g = self.graph
        uses = collections.Counter()
        for node in g.nodes:
            if node.op == "bn":
                for k in ("gamma", "beta"):
                    if k in node.vars:
                        uses[node.vars[k]] += 1
        self._bn_var_uses = uses
        for ps in self._backward_passes():
            self.cur_pass = ps
            self._ng_cache = {}
            self.g = {}          # Tensor -> (Ptr, handle or None)
            for node in reversed(g.nodes):
                if node.attrs.get("fused_into") is not None:
                    continue
                fn = getattr(self, "_b_" + node.op, None)
                if fn is None:
                    continue
                if node.op in ("softmax_xent", "gan_loss"):
                    fn(node)
                    continue
                out = node.attrs.get("final", node.outputs[0] if node.outputs else None)
                if out is None or out not in self.g:
                    continue
                fn(node, self.g[out][0])
                self._release_grad(out)
            for t in list(self.g):
                self._release_grad(t)
        self.cur_pass = None
        self._ng_cache = {}

    def _release_grad(self, t):
        p, h = self.g.pop(t)
        if self.keep_grads:
            self.grad_ptr.setdefault(t, p)
            return
        if h is not None:
            self.tfree(h)

    def contribute(self, t, nbytes, emit, dtype=None, emit_acc=None):
        """Route a gradient contribution for tensor t: the first one writes the gradient buffer;
        later ones are added in the producing kernel's epilogue when it can (emit_acc), else go
        through a temporary and a separate accumulate pass."""
        if not self._tensor_needs_grad(t):
            return
        dt = dtype or t.dtype
        if t not in self.g:
            p, h = self.talloc(nbytes)
            self.g[t] = (p, h)
            emit(p)
        elif emit_acc is not None:
            emit_acc(self.g[t][0])
        else:
            p, h = self.talloc(nbytes)
            emit(p)
            self.L("b", "mcn_accumulate", DT_CODE[dt], self.g[t][0], p, nbytes // DT_SIZE[dt],
                   tag="grad_accumulate")
            self.tfree(h)

    def _b_softmax_xent(self, node):
        logits = node.inputs[0]
        # the fused kernel already produced dlogits (fp32)
        self.g[logits] = (Ptr(node.attrs["dlogits"]), None)

    def _b_cast(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_cast", DT_CODE[y.dtype], gy, DT_CODE[x.dtype], p,
                                         x.size, tag="cast_bwd"))

    def _b_dense(self, node, gy):
        x = node.inputs[0]
        y = node.attrs["final"]
        w = node.vars["w"]
        d = node.attrs["desc"]
        n, ci = x.shape
        co = w.shape[1]
        if "b" in node.vars and self._var_trains(node.vars["b"]):
            self.L("b", "mcn_bias_grad", DT_CODE[y.dtype], gy, n, co, self.pgrad(node.vars["b"]),
                   tag=node.scope + "/dbias")
        if node.attrs["route"] == "tc":
            gyb = gy
            h = None
            if y.dtype != "bf16":
                gyb, h = self.talloc(n * co * 2)
                self.L("b", "mcn_cast", DT_CODE[y.dtype], gy, 1, gyb, n * co, tag="dlogits_bf16")
            if self._var_trains(w):
                self.L("b", "mcn_conv2d_wgrad_tc", d, self.tbuf[x], gyb, self._w_grad(node), 0,
                       tag=node.scope + "/wgrad")
            self.contribute(x, x.size * 2,
                            lambda p: self.L("b", "mcn_conv2d_dgrad_tc", d, gyb, self._w_bf16(node), p, 1, 0, 0,
                                             tag=node.scope + "/dgrad"),
                            emit_acc=lambda p: self.L("b", "mcn_conv2d_dgrad_tc", d, gyb, self._w_bf16(node), p,
                                                      1, 0, 1, tag=node.scope + "/dgrad+"))
            if h is not None:
                self.tfree(h)
        else:
            if self._var_trains(w):
                self.L("b", "mcn_conv2d_wgrad_direct", d, self.ccode, self.tbuf[x], gy, self._w_grad(node),
                       tag=node.scope + "/wgrad")
            self.contribute(x, x.size * self.csz,
                            lambda p: self.L("b", "mcn_conv2d_dgrad_direct", d, self.ccode, gy, 0,
                                             self._w_f32(node), p, tag=node.scope + "/dgrad"))
        self._ws_finish(node)

    def _b_conv2d(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        w = node.vars["w"]
        d = node.attrs["desc"]
        route = node.attrs["route"]
        if "b" in node.vars and self._var_trains(node.vars["b"]):
            self.L("b", "mcn_bias_grad", self.ccode, gy, y.size // d.Cout, d.Cout,
                   self.pgrad(node.vars["b"]), tag=node.scope + "/dbias")
        if route == "tc":
            if self._var_trains(w):
                self.L("b", "mcn_conv2d_wgrad_tc", d, self.tbuf[x], gy, self._w_grad(node), self.conv_mode,
                       tag=node.scope + "/wgrad")
            bn = self._bnred_target(node)

            def emit_dgrad(p):
                if bn is None:
                    self.L("b", "mcn_conv2d_dgrad_tc", d, gy, self._w_bf16(node), p, 1, self.conv_mode, 0,
                           tag=node.scope + "/dgrad")
                    return
                # the BN layer feeding this conv gets its backward sums from this dgrad's epilogue
                bx = bn.inputs[0]
                cb = bx.shape[-1]
                sums = self.node_buf(bn, "bwd_sums", "bn_bwd_sums:%s" % bn.scope, 2 * cb * 8, "zero")
                bn.attrs["bwd_sums"] = sums
                bv, sv = bn.vars, bn.attrs["save"]
                # when the layer's sums go straight into dbeta / dgamma (the usual case) the dgrad's last
                # block writes the final values itself: no finalize launch
                o1 = o2 = NULL
                if self._bn_sums_direct(bn) and os.environ.get("MCN_BN_BWD_FINALIZE_LAUNCH", "0") != "1":
                    o1, o2 = self.pgrad(bv["beta"]), self.pgrad(bv["gamma"])
                    bn.attrs["bwd_sums_final"] = True
                self.L("b", "mcn_conv2d_dgrad_tc_bnred", d, gy, self._w_bf16(node), p, self.conv_mode,
                       self.tbuf[bx], Ptr(sv), Ptr(sv, cb * 4),
                       self.pvar(bv["gamma"]) if "gamma" in bv else NULL,
                       self.pvar(bv["beta"]) if "beta" in bv else NULL, bn.attrs["act"], Ptr(sums), o1, o2,
                       tag=node.scope + "/dgrad+bn_bwd_sums")
            self.contribute(x, x.size * 2, emit_dgrad,
                            emit_acc=lambda p: self.L("b", "mcn_conv2d_dgrad_tc", d, gy, self._w_bf16(node), p, 1,
                                                      self.conv_mode, 1, tag=node.scope + "/dgrad+"))
        elif route == "stem":
            if self._var_trains(w):
                self.L("b", "mcn_stem_conv_wgrad", node.attrs["desc4"], Ptr(node.attrs["x4"]), gy,
                       self._w_grad(node), tag=node.scope + "/wgrad")
            assert not self._needs_input_grad(node), "stem route is only chosen for network inputs"
        elif route == "im2col":
            if self._var_trains(w):
                self.L("b", "mcn_conv2d_wgrad_tc", node.attrs["gemm_desc"], Ptr(node.attrs["col"]), gy,
                       self._w_grad(node), 0, tag=node.scope + "/wgrad")
            # the padded [kpad, Cout] storage is [tap][Cin][Cout] for its first rows = dgrad's B operand
            self.contribute(x, x.size * 2,
                            lambda p: self.L("b", "mcn_conv2d_dgrad_tc", d, gy, self._w_bf16(node), p, 1,
                                             self.conv_mode, 0, tag=node.scope + "/dgrad"),
                            emit_acc=lambda p: self.L("b", "mcn_conv2d_dgrad_tc", d, gy, self._w_bf16(node), p, 1,
                                                      self.conv_mode, 1, tag=node.scope + "/dgrad+"))
        else:
            if self._var_trains(w):
                self.L("b", "mcn_conv2d_wgrad_direct", d, self.ccode, self.tbuf[x], gy, self._w_grad(node),
                       tag=node.scope + "/wgrad")
            self.contribute(x, x.size * self.csz,
                            lambda p: self.L("b", "mcn_conv2d_dgrad_direct", d, self.ccode, gy, 0,
                                             self._w_f32(node), p, tag=node.scope + "/dgrad"))
        self._ws_finish(node)

    def _b_dwconv2d(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        w = node.vars["w"]
        d = node.attrs["desc"]
        mult = node.attrs["mult"]
        if "b" in node.vars and self._var_trains(node.vars["b"]):
            self.L("b", "mcn_bias_grad", self.ccode, gy, y.size // y.shape[-1], y.shape[-1],
                   self.pgrad(node.vars["b"]), tag=node.scope + "/dbias")
        if self._var_trains(w):
            self.L("b", "mcn_dwconv2d_bwd_filter", d, mult, self.ccode, self.tbuf[x], gy, self._w_grad(node),
                   tag=node.scope + "/dw_wgrad")
        self.contribute(x, x.size * self.csz,
                        lambda p: self.L("b", "mcn_dwconv2d_bwd_data", d, mult, self.ccode, gy, 0,
                                         self._w_f32(node), p, tag=node.scope + "/dw_dgrad"))
        self._ws_finish(node)

    def _b_conv2d_transpose(self, node, gy):
        # y = dgrad_conv(x): dL/dx = fprop_conv(gy); dL/dW_conv[tap][co][ci] = wgrad_conv(a=gy, dy=x)
        x, y = node.inputs[0], node.outputs[0]
        w = node.vars["w"]
        d = node.attrs["desc"]
        if "b" in node.vars and self._var_trains(node.vars["b"]):
            self.L("b", "mcn_bias_grad", self.ccode, gy, y.size // y.shape[-1], y.shape[-1],
                   self.pgrad(node.vars["b"]), tag=node.scope + "/dbias")
        mixed = node.attrs["route"] in ("mixed", "direct")
        if node.attrs["route"] == "direct":
            kh, kw, ci_t, co_t = w.shape
            w.gemm_dims = (kh * kw, ci_t, co_t)
        if self._var_trains(w):
            # wgrad writes [tap][Cin_conv=co][Cout_conv=ci]; the variable is stored [tap][ci][co]
            tmp, h = self.talloc(w.storage_size * 4)
            self.L("b", "mcn_fill_f32", tmp, w.storage_size, 0.0, tag="zero")
            if mixed:
                self.L("b", "mcn_conv2d_wgrad_direct", d, self.ccode, gy, self.tbuf[x], tmp,
                       tag=node.scope + "/wgrad")
            else:
                self.L("b", "mcn_conv2d_wgrad_tc", d, gy, self.tbuf[x], tmp, self.conv_mode,
                       tag=node.scope + "/wgrad")
            taps, ci, co = w.gemm_dims
            self.L("b", "mcn_transpose_add_f32", tmp, taps, co, ci, self._w_grad(node), tag="wgrad_T")
            self.tfree(h)
        if mixed:
            # conv HWIO [tap][Cin_conv=co][Cout_conv=ci] is the transposed bf16 copy
            if node.attrs["route"] == "direct":
                self.contribute(x, x.size * self.csz,
                                lambda p: self.L("b", "mcn_conv2d_fprop_direct", d, self.ccode, gy, 0,
                                                 Ptr(node.attrs["wT"]), NULL, p, tag=node.scope + "/dgrad"))
            else:
                self.contribute(x, x.size * 2,
                                lambda p: self.L("b", "mcn_conv2d_fprop_direct", d, self.ccode, gy, 1,
                                                 self._w_bf16t(node), NULL, p, tag=node.scope + "/dgrad"))
            self._ws_finish(node)
            return
        # fprop of the underlying conv needs W_conv as [tap][Cout_conv=ci][Cin_conv=co] = stored layout
        self.contribute(x, x.size * 2,
                        lambda p: self.L("b", "mcn_conv2d_fprop_tc", d, gy, self._w_bf16(node), NULL, p, 1,
                                         self.conv_mode, 0, tag=node.scope + "/dgrad"),
                        emit_acc=lambda p: self.L("b", "mcn_conv2d_fprop_tc", d, gy, self._w_bf16(node), NULL, p,
                                                  1, self.conv_mode, 1, tag=node.scope + "/dgrad+"))
        self._ws_finish(node)

    def _bn_sums_direct(self, node):
        """The layer's local backward sums double as dbeta / dgamma (both trained, used by this node only)."""
        v = node.vars
        return ("beta" in v and self._var_trains(v["beta"]) and "gamma" in v
                and self._var_trains(v["gamma"]) and self._bn_var_uses[v["beta"]] == 1
                and self._bn_var_uses[v["gamma"]] == 1)

    def _b_bn(self, node, gy):
        x = node.inputs[0]
        y = node.attrs["final"]
        c = x.shape[-1]
        rows = node.attrs["rows"]
        v = node.vars
        save = node.attrs["save"]
        act = node.attrs["act"]
        res = node.attrs["residual"]
        pg = self.pvar(v["gamma"]) if "gamma" in v else NULL
        pbeta = self.pvar(v["beta"]) if "beta" in v else NULL
        # derivative from the output when a residual was fused (or for relu-family generally it is
        # cheaper to recompute from x, which is read anyway)
        py = self.tbuf[y] if (res is not None and act != 0) else NULL
        # local sums double as dbeta / dgamma; without a trainable beta/gamma they go to scratch
        scratch = None
        direct = self._bn_sums_direct(node)
        if direct:
            s1, s2 = self.pgrad(v["beta"]), self.pgrad(v["gamma"])
        else:
            # variables shared by several BN nodes (D(real)/D(fake)) or not trained in this pass:
            # this node's own sums go to scratch and are added to the gradients afterwards
            sp, scratch = self.talloc(2 * c * 4)
            self.L("b", "mcn_fill_f32", sp, 2 * c, 0.0, tag="zero")
            s1, s2 = sp, sp + c * 4
        fused = node.attrs.pop("bwd_sums", None)
        mask = node.attrs.get("relu_mask")      # bit mask of y > 0 written by the forward apply pass
        if fused is not None and node.attrs.pop("bwd_sums_final", False):
            pass        # ... and its last block already wrote the final sums into s1 / s2
        elif fused is not None:
            # the dgrad that produced gy already took sum dz / sum dz*x in its epilogue
            self.L("b", "mcn_bn_bwd_finalize", Ptr(fused), Ptr(save), Ptr(save, c * 4), c, s1, s2,
                   tag=node.scope + "/bwd_finalize")
        elif mask is not None:
            self.L("b", "mcn_bn_bwd_reduce_mask", self.ccode, gy, self.tbuf[x], Ptr(mask), rows, c, Ptr(save),
                   Ptr(save, c * 4), s1, s2, tag=node.scope + "/bwd_reduce")
        else:
            self.L("b", "mcn_bn_bwd_reduce", self.ccode, gy, self.tbuf[x], py, rows, c, Ptr(save),
                   Ptr(save, c * 4), pg, pbeta, act, node.attrs["alpha"], s1, s2, tag=node.scope + "/bwd_reduce")
        g1, g2, gh = s1, s2, None
        count = float(rows)
        if not node.attrs["update"]:
            # frozen statistics: dx = dz*gamma*invstd — the batch-statistics terms vanish (zero sums)
            gp, gh = self.talloc(2 * c * 4)
            self.L("b", "mcn_fill_f32", gp, 2 * c, 0.0, tag="zero")
            g1, g2 = gp, gp + c * 4
        elif self.sync_bn:
            # global sums go to a scratch vector: the LOCAL sums stay in place as dbeta / dgamma
            # (they are averaged over ranks with the other gradients).  The exchange gathers its
            # two source segments itself (mcn_peer_allreduce src0/src1), no staging copies.
            gp, gh = self.talloc(2 * c * 4)
            self.allreduce_points.append(("b", len(self.bwd), gp, 2 * c * 4, "f32", (s1, c, s2, c)))
            g1, g2 = gp, gp + c * 4
            count = float(rows * self.world)
        need_x = self._tensor_needs_grad(x)
        need_r = res is not None and self._tensor_needs_grad(res)
        if need_x or need_r:
            esz = self.csz
            # residual gradient: first contribution can be written in place by the kernel
            pres, res_tmp = NULL, None
            if need_r:
                if res not in self.g:
                    p, h = self.talloc(res.size * esz)
                    self.g[res] = (p, h)
                    pres = p
                else:
                    pres, res_tmp = self.talloc(res.size * esz)
            def bwd_apply(p):
                if mask is not None:
                    self.L("b", "mcn_bn_bwd_apply_mask", self.ccode, gy, self.tbuf[x], Ptr(mask), rows, c,
                           Ptr(save), Ptr(save, c * 4), pg, g1, g2, count, p, pres, tag=node.scope + "/bwd_apply")
                else:
                    self.L("b", "mcn_bn_bwd_apply", self.ccode, gy, self.tbuf[x], py, rows, c, Ptr(save),
                           Ptr(save, c * 4), pg, pbeta, act, node.attrs["alpha"], g1, g2, count, p, pres,
                           tag=node.scope + "/bwd_apply")
            if not need_x:
                dxp, dxh = self.talloc(x.size * esz)
                bwd_apply(dxp)
                self.tfree(dxh)
            if need_x:
                self.contribute(x, x.size * esz, bwd_apply)
            if res_tmp is not None:
                self.L("b", "mcn_accumulate", self.ccode, self.g[res][0], pres, res.size,
                       tag="grad_accumulate")
                self.tfree(res_tmp)
        if gh is not None:
            self.tfree(gh)
        if scratch is not None:
            if "beta" in v and self._var_trains(v["beta"]):
                self.L("b", "mcn_accumulate", 0, self.pgrad(v["beta"]), s1, c, tag="dbeta+=")
            if "gamma" in v and self._var_trains(v["gamma"]):
                self.L("b", "mcn_accumulate", 0, self.pgrad(v["gamma"]), s2, c, tag="dgamma+=")
            self.tfree(scratch)

    # ------------------------------------------------------------------ group normalisation
    def _gn_dims(self, node):
        x = node.inputs[0]
        n, c = x.shape[0], x.shape[-1]
        return n, x.size // (n * c), c, node.attrs["groups"]

    def _f_gn(self, node):
        x, y = node.inputs[0], node.outputs[0]
        n, hw, c, g = self._gn_dims(node)
        v = node.vars
        save = self.node_buf(node, "save", "gn_save:%s" % node.scope, n * g * 2 * 4)
        node.attrs["save"] = save
        py = self.alloc_act(y)
        self.L("f", "mcn_gn_fwd", DT_CODE[x.dtype], self.tbuf[x], n, hw, c, g, node.attrs["eps"],
               self.pvar(v["gamma"]) if "gamma" in v else NULL, self.pvar(v["beta"]) if "beta" in v else NULL,
               py, Ptr(save), tag=node.scope)

    def _b_gn(self, node, gy):
        x = node.inputs[0]
        n, hw, c, g = self._gn_dims(node)
        v = node.vars
        sp, sh = self.talloc((2 * n * g + 2 * n * c) * 4)
        dgam = self.pgrad(v["gamma"]) if "gamma" in v and self._var_trains(v["gamma"]) else NULL
        dbet = self.pgrad(v["beta"]) if "beta" in v and self._var_trains(v["beta"]) else NULL
        pg = self.pvar(v["gamma"]) if "gamma" in v else NULL
        done = []

        def emit(p, params=True):
            self.L("b", "mcn_gn_bwd", DT_CODE[x.dtype], gy, self.tbuf[x], n, hw, c, g, pg,
                   Ptr(node.attrs["save"]), sp, p, dgam if params else NULL, dbet if params else NULL,
                   tag=node.scope + "/bwd")
            done.append(1)
        self.contribute(x, x.size * DT_SIZE[x.dtype], emit)
        if not done and (dgam is not NULL or dbet is not NULL):
            emit(NULL)                     # no gradient into x: parameter gradients only
        self.tfree(sh)

    def _b_act(self, node, gy):
        x = node.inputs[0]
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_act_bwd", DT_CODE[x.dtype], gy, self.tbuf[x], x.size,
                                         node.attrs["act"], node.attrs["alpha"], p, tag="act_bwd"))

    def _b_add(self, node, gy):
        a, b = node.inputs
        y = node.attrs["final"]
        esz = DT_SIZE[a.dtype]
        if node.attrs["act"] != 0:
            dz, h = self.talloc(a.size * esz)
            self.L("b", "mcn_add_act_bwd", DT_CODE[a.dtype], gy, self.tbuf[y], a.size, node.attrs["act"],
                   node.attrs["alpha"], dz, tag="add_act_bwd")
        else:
            dz, h = gy, None
        for t in (a, b):
            self.contribute(t, t.size * esz,
                            lambda p: self.L("b", "mcn_cast", DT_CODE[t.dtype], dz, DT_CODE[t.dtype], p,
                                             t.size, tag="grad_copy"))
        if h is not None:
            self.tfree(h)

    def _b_scale_bcast(self, node, gy):
        x, m = node.inputs
        n, hh, w, c = x.shape
        dm32, h32 = self.talloc(n * c * 4)
        state = {}

        def emit(p):
            self.L("b", "mcn_scale_bcast_bwd", DT_CODE[x.dtype], gy, self.tbuf[x], self.tbuf[m], n,
                   hh * w, c, p, dm32, tag="se_scale_bwd")
            state["done"] = True
        self.contribute(x, x.size * DT_SIZE[x.dtype], emit)
        if state.get("done"):
            self.contribute(m, m.size * DT_SIZE[m.dtype],
                            lambda p: self.L("b", "mcn_cast", 0, dm32, DT_CODE[m.dtype], p, n * c,
                                             tag="se_dm_cast"))
        self.tfree(h32)

    def _b_max_pool(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        a = node.attrs
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        fn = "mcn_maxpool_bwd_tap" if a.get("tap_form") else "mcn_maxpool_bwd"
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", fn, DT_CODE[x.dtype], gy, Ptr(a["argmax"]), n,
                                         h, w, c, a["k"][0], a["k"][1], a["s"][0], a["s"][1], a["pad"][0],
                                         a["pad"][1], ho, wo, p, tag="max_pool_bwd"))

    def _b_avg_pool(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        a = node.attrs
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_avgpool_bwd", DT_CODE[x.dtype], gy, n, h, w, c,
                                         a["k"][0], a["k"][1], a["s"][0], a["s"][1], a["pad"][0],
                                         a["pad"][1], ho, wo, p, tag="avg_pool_bwd"))

    def _b_gap(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        n, h, w, c = x.shape
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_gap_bwd", DT_CODE[x.dtype], gy, DT_CODE[y.dtype], n,
                                         h * w, c, p, tag="gap_bwd"))

    def _b_concat(self, node, gy):
        y = node.outputs[0]
        ctot = y.shape[-1]
        rows = y.size // ctot
        off = 0
        for t in node.inputs:
            c = t.shape[-1]
            o = off
            self.contribute(t, t.size * DT_SIZE[t.dtype],
                            lambda p: self.L("b", "mcn_copy_channels", DT_CODE[t.dtype], gy, rows, ctot, o,
                                             p, c, 0, c, 0, tag="concat_bwd"))
            off += c

    def _b_reshape(self, node, gy):
        x = node.inputs[0]
        if x not in self.g and self._tensor_needs_grad(x):
            # a reshape is a view: hand the gradient buffer over instead of copying
            self.g[x] = self.g[node.outputs[0]]
            self.g[node.outputs[0]] = (gy, None)
            return
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_cast", DT_CODE[x.dtype], gy, DT_CODE[x.dtype], p, x.size,
                                         tag="grad_copy"))

    def _b_resize_bilinear(self, node, gy):
        x, y = node.inputs[0], node.outputs[0]
        n, h, w, c = x.shape
        _, ho, wo, _ = y.shape
        self.contribute(x, x.size * DT_SIZE[x.dtype],
                        lambda p: self.L("b", "mcn_resize_bilinear_bwd", DT_CODE[x.dtype], gy, n, h, w, c,
                                         ho, wo, node.attrs["mode"], p, tag="resize_bwd"))

    # ------------------------------------------------------------------ gradient buckets
    def grad_bucket_schedule(self, bucket_elems):
        """Slices of the flat gradient buffer and, for each, the index of the LAST backward launch
        that writes into it: [(start_elem, end_elem, ready_idx)], ready_idx = -1 when no launch
        does.  The engine starts a bucket's all-reduce right after launch `ready_idx`, so the
        exchange of the last layers' gradients overlaps the rest of the backward pass (replaces
        the reference's gather-everything-then-average, optimizers.py:117-147)."""
        import bisect
        from .dist import bucket_ranges
        starts = [self.var_off[v] for v in self.trainable]
        ends = [self.var_off[v] + v.storage_size for v in self.trainable]
        last = [-1] * len(starts)
        for li, l in enumerate(self.bwd):
            for a in l.args:
                if isinstance(a, Ptr) and a.buf is self.b_grad:
                    e = a.off // 4
                    vi = bisect.bisect_right(starts, e) - 1
                    if 0 <= vi < len(starts) and e < ends[vi]:
                        last[vi] = max(last[vi], li)
        out = []
        for s0, e0 in bucket_ranges(self.n_train, bucket_elems):
            lo = max(0, bisect.bisect_right(starts, s0) - 1)
            ready = -1
            for vi in range(lo, len(starts)):
                if starts[vi] >= e0:
                    break
                if ends[vi] > s0:
                    ready = max(ready, last[vi])
            out.append((s0, e0, ready))
        return out

    # ------------------------------------------------------------------ summaries
    def launch_histogram(self):
        h = collections.Counter()
        for l in self.fwd + self.bwd:
            h[l.fn] += 1
        return dict(h)
