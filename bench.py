"""Benchmark of the hot path: ResNet-v1.5-50 data-parallel training step on synthetic ImageNet-shaped
data (BASELINE.json configs[1]: batch 256 per GPU, bf16, synchronised BN across GPUs).

  python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  `value` times the step with inputs resident in HBM; `e2e` times
the same step through the public API with HOST input buffers (pinned) copied in and the loss read
back every step.  `roofline` describes the kernel class that takes the largest share of the step,
timed live with CUDA events on the launching stream.  `cpu_baseline` is the CPU restatement of the
reference's num_gpus=0 path (TensorFlow cannot be installed here: see oracle/tf_ops.py) on a
bounded sample (batch 32, fp32), timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T0 = time.time()

MODEL = "ResNet-v1.5-50"
IMG = [224, 224, 3]
NCLS = 1000

# BASELINE.json configs.  The default ("r50") is configs[1], the one the headline metric is quoted
# on; the others are measured with --config and recorded under profiles/ (same JSON schema).
CONFIGS = {
    "r50": "ResNet-v1.5-50, synthetic 224x224x3, 1000 classes, batch 256/GPU, bf16 (BASELINE configs[1])",
    "r50_fp32_b32": "ResNet-v1.5-50, synthetic 224x224x3, batch 32, fp32 (GPU leg of BASELINE configs[0])",
    "effnet_b0": "EfficientNet-B0, synthetic 224x224x3, 1000 classes, batch 256/GPU, bf16 (BASELINE configs[2])",
    "deeplab_r50_512": "DeepLabv3+ on dilated ResNet-v1.5-50 (OS16), synthetic 512x512x3, 21 classes, "
                       "batch 16/GPU, bf16 (BASELINE configs[3])",
    "dcgan_64": "DCGAN 64x64x3, latent 100, batch 128, bf16, D and G updated from one forward "
                "(BASELINE configs[4]; reference semantics, optimizers_gan.py:56-58)",
}


def build_workload(args, np, torch):
    """(model, source, X, Y, optimizer, dtype, description) of the selected BASELINE config."""
    from myconvnet_b200 import loader
    from myconvnet_b200.zoo import resnet50
    cfg = args.config
    rng = np.random.default_rng(1234 + int(os.environ.get("RANK", "0")))
    u8 = lambda shape: torch.from_numpy(rng.integers(0, 256, size=shape, dtype=np.uint8)).pin_memory()  # noqa: E731
    if cfg in ("r50", "r50_fp32_b32"):
        dt = "bf16" if cfg == "r50" else "f32"
        batch = args.batch or (256 if cfg == "r50" else 32)
        # images cross the host boundary as raw uint8 NHWC (what an image data set holds); the device
        # prologue divides by 255, zero-centres, scales and casts (mcn_input_prep, convnet.py:449-471)
        model, src = resnet50(IMG, NCLS, batch_size=batch, compute_dtype=dt, input_dtype="u8")
        X = u8([batch] + IMG)
        Y = torch.from_numpy(rng.integers(0, NCLS, size=batch).astype(np.int32)).pin_memory()
        return model, src, X, Y, "nesterov", dt, batch
    if loader.reference_root() is None:
        raise RuntimeError("config %s needs the reference model files (scripts/stage_reference.py)" % cfg)
    fac = loader.product_facade()
    if cfg == "effnet_b0":
        batch = args.batch or 256
        mod = loader.load_reference_model("models/efficientnet.py", fac)
        model = mod.EfficientNetB0(IMG, NCLS, batch_size=batch, compute_dtype="bf16", input_dtype="u8")
        X = u8([batch] + IMG)
        Y = torch.from_numpy(rng.integers(0, NCLS, size=batch).astype(np.int32)).pin_memory()
        return model, "reference models/efficientnet.py (unchanged)", X, Y, "rmsprop", "bf16", batch
    if cfg == "deeplab_r50_512":
        batch = args.batch or 16
        mod = loader.load_reference_model("models/deeplabv3plus.py", fac)

        class DeepLabV3PlusResNet50(mod.DeepLabV3PlusResNet):
            # the file hard-imports ResNet101OS16; the BASELINE config names the ResNet-50 depth
            # (res_units of resnet_v1_5_dilated.py:10 with the OS16 strides of :157-162)
            def _init_params(self, **kwargs):
                mod.DeepLabV3PlusResNet._init_params(self, **kwargs)
                self.res_units = [None, 3, 4, 6, 3]
        shape = [512, 512, 3]
        model = DeepLabV3PlusResNet50(shape, 21, batch_size=batch, compute_dtype="bf16", input_dtype="u8")
        X = u8([batch] + shape)
        Y = torch.from_numpy(rng.integers(0, 22, size=[batch, 512, 512]).astype(np.int32)).pin_memory()
        return model, "reference models/deeplabv3plus.py (unchanged; ResNet-50 depth)", X, Y, "nesterov", "bf16", batch
    if cfg == "dcgan_64":
        batch = args.batch or 128
        mod = loader.load_reference_model("models/dcgan.py", fac)
        shape = [64, 64, 3]
        model = mod.DCGAN(shape, 100, batch_size=batch, compute_dtype="bf16", input_dtype="u8",
                          base_learning_rate=5e-4, momentum=0.5, generator_scaling_factor=2.0)
        X = u8([batch] + shape)
        Y = torch.from_numpy(rng.uniform(-1, 1, size=(batch, 100)).astype(np.float32)).pin_memory()
        return model, "reference models/dcgan.py (unchanged)", X, Y, "adam", "bf16", batch
    raise ValueError("unknown --config %s (one of %s)" % (cfg, ", ".join(CONFIGS)))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference_step_rate(steps, warmup, batch=32):
    """Times the oracle restatement of the reference step (fwd + bwd + Nesterov update, fp32) on
    the host cores.  Returns (img/s, seconds per step, threads)."""
    import numpy as np
    import torch
    from myconvnet_b200 import loader
    from myconvnet_b200.engine import draw_initial_value
    from myconvnet_b200.zoo import resnet50
    from oracle import ref_convnet
    from oracle.step import OracleTrainer

    pm, _ = resnet50(IMG, NCLS, batch_size=batch, compute_dtype="f32")
    rng = np.random.default_rng(0)
    vals = {v.name: draw_initial_value(v, rng) for v in pm.graph.vars.values()}
    if loader.reference_root() is not None:
        om = loader.load_reference_model("models/resnet_v1_5.py", {"convnet": ref_convnet}).ResNet50(IMG, NCLS)
    else:
        from myconvnet_b200 import zoo
        om = type("OracleResNet50", (ref_convnet.ConvNet,),
                  {"_build_model": zoo.ResNet50._build_model, "_bottleneck": zoo.ResNet50._bottleneck,
                   "channels": zoo.ResNet50.channels, "units": zoo.ResNet50.units,
                   "strides": zoo.ResNet50.strides})(IMG, NCLS)
    om.set_variables(vals)
    tr = OracleTrainer(om)
    X = rng.uniform(size=[batch] + IMG).astype(np.float32)
    Y = rng.integers(0, NCLS, size=batch)
    for _ in range(warmup):
        tr.step(X, Y)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(X, Y)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm uses every host core
    import torch
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    steps, warmup = max(1, args.steps), max(1, args.warmup)     # ~1.5 s per batch-32 step on 16 cores
    rate, dt, threads = cpu_reference_step_rate(steps, warmup)
    line = {
        "impl": "reference", "metric": "train_images_per_sec", "value": rate, "unit": "img/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s train step, synthetic 224x224x3" % MODEL,
                   "sample": "batch 32 per step on the host cores (bounded sample of the batch-256 workload)"},
        "cpu_baseline": {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of batch 32, fp32, torch-CPU restatement of the reference "
                                   "num_gpus=0 path (TensorFlow unavailable offline)" % steps},
        "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ roofline helpers
def launch_work(name, args, esz):
    """Algorithmic (bytes, flops) of one libmcn launch from its resolved argument tuple — the
    per-unit figures of SURVEY 8(d) / DESIGN.md section 4 times the units the launch processes."""
    if name == "mcn_conv2d_dgrad_tc_bnred":
        # the dgrad's own bytes + one read of the BN input (the reduction pass it replaces read 2 tensors)
        d = args[0]._obj
        m = d.N * d.Ho * d.Wo
        flops = 2.0 * m * d.kh * d.kw * d.Cin * d.Cout
        byt = esz * (2 * d.N * d.H * d.W * d.Cin + m * d.Cout) + esz * d.kh * d.kw * d.Cin * d.Cout
        return byt, flops
    if name in ("mcn_conv2d_fprop_tc", "mcn_conv2d_fprop_tc_stats", "mcn_conv2d_dgrad_tc",
                "mcn_conv2d_wgrad_tc", "mcn_conv2d_fprop_direct", "mcn_conv2d_dgrad_direct",
                "mcn_conv2d_wgrad_direct"):
        d = args[0]._obj
        m = d.N * d.Ho * d.Wo
        flops = 2.0 * m * d.kh * d.kw * d.Cin * d.Cout
        byt = esz * (d.N * d.H * d.W * d.Cin + m * d.Cout) + esz * d.kh * d.kw * d.Cin * d.Cout
        return byt, flops
    if name in ("mcn_stem_conv_fprop", "mcn_stem_conv_wgrad"):
        d = args[0]._obj
        m = d.N * d.Ho * d.Wo
        flops = 2.0 * m * d.kh * d.kw * 3 * d.Cout            # the real taps / channels
        byt = esz * (d.N * d.H * d.W * 4 + m * d.Cout)
        return byt, flops
    if name in ("mcn_dwconv2d_fwd", "mcn_dwconv2d_bwd_data", "mcn_dwconv2d_bwd_filter"):
        d = args[0]._obj
        return esz * (d.N * d.H * d.W * d.Cin + d.N * d.Ho * d.Wo * d.Cin), 0.0       # (|x| + |y|) * s
    if name == "mcn_bn_stats":
        return args[2] * args[3] * esz, 0.0
    if name == "mcn_bn_apply":
        n = args[2] * args[3]
        return n * esz * (2 + (1 if args[8] else 0)), 0.0
    if name == "mcn_bn_apply_stats":
        n = args[2] * args[3]
        return n * esz * (2 + (1 if args[10] else 0)), 0.0
    if name == "mcn_bn_apply_stats_mask":          # x + residual in, y + one mask bit per element out
        n = args[2] * args[3]
        return n * (esz * 3 + 0.125), 0.0
    if name == "mcn_bn_bwd_reduce_mask":           # dy, x and the bit mask
        n = args[4] * args[5]
        return n * (esz * 2 + 0.125), 0.0
    if name == "mcn_bn_bwd_apply_mask":            # dy, x, mask in; dx (+ residual gradient) out
        n = args[4] * args[5]
        return n * (esz * (3 + (1 if args[13] else 0)) + 0.125), 0.0
    if name == "mcn_bn_bwd_reduce":
        n = args[4] * args[5]
        return n * esz * (2 + (1 if args[3] else 0)), 0.0
    if name == "mcn_bn_bwd_apply":
        n = args[4] * args[5]
        return n * esz * (3 + (1 if args[3] else 0) + (1 if args[16] else 0)), 0.0
    if name in ("mcn_maxpool_fwd_tap", "mcn_maxpool_bwd_tap"):
        # fwd: x in [N,H,W,C], y + one tap byte out; bwd: dy + tap in, dx out — same bytes
        off = 2 if name == "mcn_maxpool_fwd_tap" else 3
        n, h, w, c = args[off:off + 4]
        ho, wo = args[off + 10], args[off + 11]
        return n * c * (h * w * esz + ho * wo * (esz + 1)), 0.0
    if name == "mcn_gap_fwd":
        return args[2] * args[3] * args[4] * esz, 0.0
    if name == "mcn_gap_bwd":
        return args[3] * args[4] * args[5] * esz, 0.0
    if name in ("mcn_scale_bcast_fwd", "mcn_scale_bcast_bwd"):
        off = 3 if name == "mcn_scale_bcast_fwd" else 4
        n = args[off] * args[off + 1] * args[off + 2]
        return n * esz * (2 if name == "mcn_scale_bcast_fwd" else 3), 0.0
    if name in ("mcn_act_fwd", "mcn_add_act_fwd"):
        return args[2 if name == "mcn_act_fwd" else 3] * esz * (2 if name == "mcn_act_fwd" else 3), 0.0
    return 0.0, 0.0


# libmcn entry point -> the CUDA kernel (family) it launches: launches are grouped by KERNEL, so the
# fprop and dgrad calls of gemm_conv_kernel / halo_conv_kernel form one class
KERNEL_OF = {
    "mcn_conv2d_fprop_tc": "conv_tc (gemm_conv_kernel / halo_conv_kernel: fprop + dgrad)",
    "mcn_conv2d_fprop_tc_stats": "conv_tc (gemm_conv_kernel / halo_conv_kernel: fprop + dgrad)",
    "mcn_conv2d_dgrad_tc": "conv_tc (gemm_conv_kernel / halo_conv_kernel: fprop + dgrad)",
    "mcn_conv2d_dgrad_tc_bnred": "conv_tc (gemm_conv_kernel / halo_conv_kernel: fprop + dgrad)",
    "mcn_conv2d_wgrad_tc": "wgrad_tc (wgrad_kernel / wgrad_halo_kernel + splitk_reduce)",
    "mcn_stem_conv_fprop": "stem (stem_fprop_kernel / stem_wgrad_kernel)",
    "mcn_stem_conv_wgrad": "stem (stem_fprop_kernel / stem_wgrad_kernel)",
    "mcn_bn_apply_stats": "bn_apply", "mcn_bn_apply": "bn_apply", "mcn_bn_apply_stats_mask": "bn_apply",
    "mcn_bn_bwd_reduce_mask": "bn_bwd_reduce", "mcn_bn_bwd_apply_mask": "bn_bwd_apply",
    "mcn_maxpool_fwd_tap": "maxpool", "mcn_maxpool_bwd_tap": "maxpool",
    "mcn_dwconv2d_fwd": "dwconv", "mcn_dwconv2d_bwd_data": "dwconv", "mcn_dwconv2d_bwd_filter": "dwconv",
    "mcn_conv2d_fprop_direct": "conv_direct (igemm_kernel)", "mcn_conv2d_dgrad_direct": "conv_direct (igemm_kernel)",
    "mcn_conv2d_wgrad_direct": "conv_direct (igemm_kernel)",
}


def kernel_class(name, tag):
    return KERNEL_OF.get(name, name.replace("mcn_", ""))


def profile_step(eng):
    """Times every launch of one step with CUDA events (eager, serialised on the launching stream).
    Returns {class: [ms, bytes, flops, launches]}."""
    import torch
    from myconvnet_b200 import lib as L
    st = torch.cuda.current_stream().cuda_stream
    torch.cuda.nvtx.range_push("mcn_profiled_step")      # ncu --nvtx --nvtx-include "mcn_profiled_step/"
    L.check(eng.lib.mcn_fill_f32(eng._zero_ptr, eng._zero_n, 0.0, st))
    recs = []
    for launches in (eng._fwd, eng._bwd):
        for fn, args, name, tag in launches:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(fn(*args, st), name)
            e1.record()
            recs.append((name, tag, args, e0, e1))
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    esz = 2 if eng.plan.cdt == "bf16" else 4
    out = {}
    detail = []
    for name, tag, args, e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        byt, fl = launch_work(name, args, esz)
        detail.append((kernel_class(name, tag), tag, ms, byt, fl))
        c = out.setdefault(kernel_class(name, tag), [0.0, 0.0, 0.0, 0])
        c[0] += ms
        c[1] += byt
        c[2] += fl
        c[3] += 1
    profile_step.detail = detail
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--config", default="r50", choices=sorted(CONFIGS), help="BASELINE config to run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel-class table here")
    ap.add_argument("--conv-mode", type=int, default=None, help="0 box, 1 im2col, 2 halo where eligible")
    ap.add_argument("--stall-timeout", type=int, default=120,
                    help="abort if no benchmark stage completes within this many seconds")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from myconvnet_b200 import lib as L
    from myconvnet_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def stage(msg):
        sys.stderr.write("[bench rank %d %.1fs] %s\n" % (rank, time.time() - T0, msg))
        sys.stderr.flush()

    # A collective that never completes must not hang the driver: a watchdog ends the process
    # with a diagnostic if no stage is reached for --stall-timeout seconds.
    last_progress = [time.time()]

    def watchdog():
        while True:
            time.sleep(5)
            if time.time() - last_progress[0] > args.stall_timeout:
                sys.stderr.write("[bench rank %d] no progress for %ds: aborting\n" % (rank, args.stall_timeout))
                sys.stderr.flush()
                os._exit(3)
    threading.Thread(target=watchdog, daemon=True).start()
    _stage = stage

    def stage(msg):  # noqa: F811
        last_progress[0] = time.time()
        _stage(msg)

    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import datetime
        # NVLS (in-switch multicast) needs a healthy fabric-manager/IMEX setup; the plain NVLink
        # ring/tree paths do not.  Opt in with MCN_NCCL_NVLS=1.
        # NCCL prints its version banner (NCCL_DEBUG >= VERSION) on STDOUT, next to the JSON line:
        # send NCCL's own log to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("MCN_NCCL_NVLS", "0") != "1":
            os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        import torch.distributed as dist
        stage("init_process_group(nccl), world %d" % world)
        dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=args.stall_timeout))
        pg = dist.group.WORLD
        t = torch.ones(1, device="cuda")
        dist.all_reduce(t)
        torch.cuda.synchronize()
        stage("first all-reduce done (%d)" % int(t.item()))
    warmup = max(args.warmup, 3)
    stage("building model and engine")
    model, model_src, X, Y, optimizer, cdt, args.batch = build_workload(args, np, torch)
    eng = Engine(model, optimizer=optimizer, world_size=world, rank=rank, process_group=pg,
                 use_cuda_graph=(not args.no_graph), seed=0, fetch_pred=False,
                 **({"conv_mode": args.conv_mode} if args.conv_mode is not None else {}))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    stage("warm-up (%d steps; step 2 captures the CUDA graph)" % warmup)
    eng.load_inputs(X=X, Y=Y)
    for i in range(warmup):
        eng.train_step(fetch_loss=False)
        torch.cuda.synchronize()
        stage("warm-up step %d done" % i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.train_step(fetch_loss=False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    stage("timed region done: %.2f ms/step" % ms)
    sampler.stop_flag = True
    if rank == 0:
        # a query still in flight (the first nvidia-smi on a fresh box takes seconds) must not overlap
        # the end-to-end region: it measurably slows kernel launches (33 vs 23 ms/step observed)
        sampler.join(timeout=10)
    loss = eng.read_loss()
    eager_launches = eng.launches_per_step()
    # ---- end-to-end: every step's batch comes from pinned HOST memory (H2D inside the timed
    # region) and the loss is read back (D2H) every step.  The upload of batch i+1 runs on a copy
    # stream while step i computes (Engine.prefetch_inputs), as an input pipeline would do.
    # warm the pipeline itself first: the pinned staging rings of prefetch_inputs / enqueue_loss_read
    # are allocated on first use (cudaHostAlloc of 38 MB takes milliseconds and synchronises)
    eng.prefetch_inputs(X=X, Y=Y)
    for _ in range(3):
        eng._consume_prefetched()
        eng.prefetch_inputs(X=X, Y=Y)
        eng.train_step(fetch_loss=False)
        eng.enqueue_loss_read()
        eng.pop_loss()
    eng._consume_prefetched()
    barrier()
    t0 = time.perf_counter()
    eng.prefetch_inputs(X=X, Y=Y)
    e2e_losses = []
    for i in range(args.steps):
        eng._consume_prefetched()
        if i + 1 < args.steps:
            eng.prefetch_inputs(X=X, Y=Y)
        eng.train_step(fetch_loss=False)
        eng.enqueue_loss_read()                  # D2H of this step's loss, asynchronous
        if i > 0:
            e2e_losses.append(eng.pop_loss())    # the loss of step i-1, read while step i runs
    e2e_losses.append(eng.pop_loss())
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    stage("end-to-end region done: %.2f ms/step" % ms_e2e)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    dp = None
    if world > 1 and args.config == "r50":
        # numeric parity of the N-rank step against one device at the global batch (small model)
        from myconvnet_b200.dp_check import dp_parity
        dp = dp_parity(rank, world, dtype="bf16", graph=True)
        stage("data-parallel parity check done")
    if world > 1:
        # Tearing the NCCL communicator down while captured graphs still reference it can block;
        # peers leave right after the last collective and rank 0 exits with os._exit below.
        torch.cuda.synchronize()
        stage("collectives done")
    if rank != 0:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    # ---- roofline of the dominant kernel class (live CUDA-event timing of every launch)
    hbm_peak, tc_peak, peak_src = measured_peaks()
    prof = profile_step(eng)
    prof = profile_step(eng)
    total_ms = sum(v[0] for v in prof.values())
    esz = 2 if eng.plan.cdt == "bf16" else 4

    def roof_of(cname, cms, cbytes, cflops, cn):
        """achieved = algorithmic bytes (flops) of the class / its measured time; the bound is the one
        that takes longer at the measured peaks."""
        tensor_bound = cflops > 0 and (cflops / (tc_peak * 1e12)) > (cbytes / (hbm_peak * 1e9))
        if tensor_bound:
            ach = cflops / (cms * 1e-3) / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak}
        else:
            ach = cbytes / (cms * 1e-3) / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak}
        r.update({"kernel": cname, "launches_per_step": cn, "share_of_step": cms / total_ms,
                  "ms_per_step": cms})
        return r
    # per-LAUNCH roofline time = max(bytes / HBM peak, flops / tensor peak): what the class would take
    # if every launch sat on its own roof (a class mixes HBM-bound 1x1 and tensor-bound 3x3 layers)
    per_launch = {}
    for cname, tag, lms, lb, lf in profile_step.detail:
        t_roof = max(lb / (hbm_peak * 1e9), lf / (tc_peak * 1e12)) * 1e3
        per_launch[cname] = per_launch.get(cname, 0.0) + t_roof
    by_kernel = []
    for cname, (cms, cbytes, cflops, cn) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        if cms / total_ms < 0.005 or (cbytes == 0 and cflops == 0):
            continue
        r = roof_of(cname, cms, cbytes, cflops, cn)
        r["frac_of_per_launch_roofline"] = per_launch.get(cname, 0.0) / cms if cms > 0 else None
        by_kernel.append(r)
    top = max(prof.items(), key=lambda kv: kv[1][0])
    cname, (cms, cbytes, cflops, cn) = top
    roof = roof_of(cname, cms, cbytes, cflops, cn)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel class, from the
        # committed ncu capture of the same command (scripts/ncu_profile.sh)
        t = json.load(open(tpath)).get("classes", {}).get(cname.split(" ")[0])
        if t:
            traffic = t["dram_bytes_per_launch"]
    roof.update({"peak_source": peak_src, "traffic": traffic,
                 "note": "dominant CUDA kernel by time share (fprop and dgrad are the same kernel); "
                         "achieved = algorithmic flops or bytes of all its launches in one step / their "
                         "CUDA-event time; peak = sustained bf16 GEMM / copy bandwidth of MEASURED_PEAKS.json"})
    table = {k: {"ms": v[0], "GB": v[1] / 1e9, "TFLOP": v[2] / 1e12, "launches": v[3],
                 "GB/s": (v[1] / 1e9) / (v[0] * 1e-3) if v[0] > 0 else 0,
                 "TFLOP/s": (v[2] / 1e12) / (v[0] * 1e-3) if v[0] > 0 else 0}
             for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    if args.profile_json:
        os.makedirs(os.path.dirname(args.profile_json) or ".", exist_ok=True)
        json.dump({"serialized_step_ms": total_ms, "graph_step_ms": ms, "classes": table,
                   "launches": [{"class": c, "tag": t, "us": m * 1e3, "GB/s": (b / 1e9) / (m * 1e-3) if m > 0 else 0,
                                 "TFLOP/s": (f / 1e12) / (m * 1e-3) if m > 0 else 0}
                                for c, t, m, b, f in profile_step.detail]},
                  open(args.profile_json, "w"), indent=1)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, threads = cpu_reference_step_rate(steps=2, warmup=1)
        cpu = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
               "sample": "2 steps of batch 32 fp32 (oracle restatement of the reference num_gpus=0 path; "
                         "TensorFlow unavailable offline), %.1f s/step" % dt}
    gb = args.batch * world
    line = {
        "metric": "train_images_per_sec", "value": gb / (ms * 1e-3), "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": cdt, "data": "synthetic",
        "config": {"workload": "%s: one training step (forward + backward + %s update with L2 / EMA), "
                               "fp32 master weights, sync-BN across GPUs" % (CONFIGS[args.config], optimizer),
                   "name": args.config,
                   "global_batch": gb, "parallelism": "dp%d" % world, "model_source": model_src,
                   "input": "uint8 NHWC images from pinned host memory, normalised on the device",
                   "l2_flush": "not needed: one step touches %.1f GB of HBM per GPU (>> 126 MB L2)"
                               % (eng.plan.arena_bytes / 1e9),
                   "cuda_graph": bool(eng.use_cuda_graph),
                   "sync_bn_exchange": ("none (1 GPU)" if world == 1 else
                                        "one-shot all-reduce kernel over NVLink peer memory (csrc/comm.cu)"
                                        if eng._peer is not None else "ncclAllReduce per layer"),
                   "grad_allreduce": ("none (1 GPU)" if world == 1 else
                                      "NCCL, %d buckets started during backward + %d after"
                                      % (sum(len(v) for v in eng._bucket_ready.values()), len(eng._bucket_tail)))},
        "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(X.numel() * X.element_size() + Y.numel() * Y.element_size() + 64),
                "d2h_bytes_per_step": 48,
                "pipeline": "batch i+1 uploads on a copy stream while step i runs; the loss of step i is "
                            "copied to pinned memory behind the step and read by the host while step "
                            "i+1 runs (Engine.enqueue_loss_read / pop_loss): every step's loss is read",
                "losses_read": len(e2e_losses), "last_loss": e2e_losses[-1] if e2e_losses else None},
        "gpu_launches": eager_launches * args.steps,
        "launches_counted_by_library": L.launch_count() - launches0,
        "final_loss": loss,
        "dp_parity": dp,
        "roofline": roof,
        "roofline_by_kernel": by_kernel,
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
    }
    print(json.dumps(line))
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


if __name__ == "__main__":
    main()
