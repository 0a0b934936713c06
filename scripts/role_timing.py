"""Per-role stall breakdown of the tensor-core conv kernels on ResNet-50 batch-256 shapes.
Run with the instrumented library:  MCN_ROLE_TIMING=1 python -m myconvnet_b200.build   (here), then
MCN_LIB=myconvnet_b200/libmcn_timing.so python scripts/role_timing.py > profiles/rNN_role_timing.txt
Columns are per-CTA averages in cycles (include/mcn.h, mcn_debug_role_cycles)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myconvnet_b200 import lib as L

lib = L.load()
L.ensure_workspace(256 << 20)
N = int(os.environ.get("BATCH", "256"))
# name, op (f fprop, d dgrad, d+ accumulate, w wgrad), H, Cin, Cout, k, stride, stats
CASES = [
    ("b1 conv_0 1x1 256->64", "f", 56, 256, 64, 1, 1, 1),
    ("b1 conv_1 3x3 64->64 (halo)", "f", 56, 64, 64, 3, 1, 1),
    ("b1 conv_1 3x3 64->64 (halo)", "d", 56, 64, 64, 3, 1, 0),
    ("b1 conv_2 1x1 64->256", "f", 56, 64, 256, 1, 1, 0),
    ("b1 conv_2 1x1 64->256", "d", 56, 64, 256, 1, 1, 0),
    ("b1 conv_0 1x1 256->64", "d+", 56, 256, 64, 1, 1, 0),
    ("b2 conv_1 3x3 128->128", "f", 28, 128, 128, 3, 1, 1),
    ("b2 conv_2 1x1 128->512", "f", 28, 128, 512, 1, 1, 0),
    ("b3 conv_0 1x1 1024->256", "f", 14, 1024, 256, 1, 1, 1),
    ("b3 conv_1 3x3 256->256", "f", 14, 256, 256, 3, 1, 1),
    ("b3 conv_1 3x3 256->256", "d", 14, 256, 256, 3, 1, 0),
    ("b3 conv_2 1x1 256->1024", "f", 14, 256, 1024, 1, 1, 0),
    ("b3 conv_0 1x1 1024->256", "d+", 14, 1024, 256, 1, 1, 0),
    ("b4 conv_1 3x3 512->512", "f", 7, 512, 512, 3, 1, 1),
    ("b4 conv_2 1x1 512->2048", "f", 7, 512, 2048, 1, 1, 0),
    ("b4 conv_0 1x1 2048->512", "f", 7, 2048, 512, 1, 1, 1),
    ("b1 conv_1 3x3 64->64", "w", 56, 64, 64, 3, 1, 0),
    ("b1 conv_2 1x1 64->256", "w", 56, 64, 256, 1, 1, 0),
    ("b2 conv_1 3x3 128->128", "w", 28, 128, 128, 3, 1, 0),
    ("b3 conv_0 1x1 1024->256", "w", 14, 1024, 256, 1, 1, 0),
    ("b3 conv_1 3x3 256->256", "w", 14, 256, 256, 3, 1, 0),
    ("b3 conv_2 1x1 256->1024", "w", 14, 256, 1024, 1, 1, 0),
    ("b4 conv_0 1x1 2048->512", "w", 7, 2048, 512, 1, 1, 0),
    ("b4 conv_1 3x3 512->512", "w", 7, 512, 512, 3, 1, 0),
]
if os.environ.get("ONLY_STEM") == "1":
    CASES = []
SLOTS = ["prod_wait_empty", "mma_wait_full", "mma_wait_acc", "epi_wait_acc", "epi_total", "cta_life", "ctas", "mma_total"]


def cycles(reset):
    buf = (ctypes.c_ulonglong * 16)()
    L.check(lib.mcn_debug_role_cycles(ctypes.cast(buf, ctypes.c_void_p), 1 if reset else 0))
    return list(buf)


print("%-34s %-3s %8s | %s" % ("case", "op", "us", " ".join("%9s" % s[:9] for s in SLOTS)))
for name, op, hw, ci, co, k, s, stats in CASES:
    pad = (k - 1) // 2
    ho = (hw + s - 1) // s
    d = L.ConvDescC(N, hw, hw, ci, co, k, k, s, s, 1, 1, pad, pad, ho, ho)
    x = torch.randn(N, hw, hw, ci, device="cuda").bfloat16()
    dy = torch.randn(N, ho, ho, co, device="cuda").bfloat16()
    w_hwio = (torch.randn(k * k, ci, co, device="cuda") * 0.05).bfloat16()
    w_ohwi = w_hwio.transpose(1, 2).contiguous()
    y = torch.empty(N, ho, ho, co, device="cuda", dtype=torch.bfloat16)
    dx = torch.zeros(N, hw, hw, ci, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(k * k, ci, co, device="cuda")
    sums = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        if op == "f" and stats:
            L.check(lib.mcn_conv2d_fprop_tc_stats(d, x.data_ptr(), w_ohwi.data_ptr(), None, y.data_ptr(), 2,
                                                  sums.data_ptr(), st))
        elif op == "f":
            L.check(lib.mcn_conv2d_fprop_tc(d, x.data_ptr(), w_ohwi.data_ptr(), None, y.data_ptr(), 1, 2, 0, st))
        elif op in ("d", "d+"):
            L.check(lib.mcn_conv2d_dgrad_tc(d, dy.data_ptr(), w_hwio.data_ptr(), dx.data_ptr(), 1, 2,
                                            1 if op == "d+" else 0, st))
        else:
            L.check(lib.mcn_conv2d_wgrad_tc(d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), 2, st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    cycles(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    c = cycles(True)
    n = max(1, c[6])
    per = [v / n for v in c[:8]]
    per[6] = c[6] / iters
    print("%-34s %-3s %8.1f | %s" % (name, op, us, " ".join("%9.0f" % v for v in per)))
    del x, dy, y, dx

# ---- the RGB stem (stem_fprop_kernel): 7x7 stride 2, 224^2 x 4 -> 112^2 x 64
if os.environ.get("STEM", "1") == "1":
    n, h, co, k = N, 224, 64, 7
    d4 = L.ConvDescC(n, h, h, 4, co, k, k, 2, 2, 1, 1, 2, 2, 112, 112)
    kpad = lib.mcn_stem_conv_kpad(d4)
    x4 = torch.randn(n, h, h, 4, device="cuda").bfloat16()
    w_t = (torch.randn(co, kpad, device="cuda") * 0.05).bfloat16()
    y = torch.empty(n, 112, 112, co, device="cuda", dtype=torch.bfloat16)
    sums = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    for with_stats in (1, 0):
        run = lambda: L.check(lib.mcn_stem_conv_fprop(d4, x4.data_ptr(), w_t.data_ptr(), None, y.data_ptr(),
                                                      sums.data_ptr() if with_stats else None, st))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        cycles(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        c = cycles(True)
        nc = max(1, c[6])
        names = ["prod_wait_empty", "mma_wait_full", "mma_wait_acc", "epi_wait_acc", "epi_total", "mma_life", "ctas",
                 "mma_total", "prod_wait_copies", "prod_life"]
        print("stem fprop 7x7 s2 (stats=%d) %8.1f us | %s" % (with_stats, e0.elapsed_time(e1) * 100,
              " ".join("%s=%.0f" % (nm, (c[i] / nc) if i != 6 else c[i] / 10) for i, nm in enumerate(names))))
