"""ctypes binding of libmcn.so (the C ABI declared in include/mcn.h).

There is no fallback: if the library is missing or a symbol cannot be bound, importing code fails
loudly.  The signature table below is checked against the header by tests/test_abi.py.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCN_LIB") or os.path.join(HERE, "libmcn.so")   # MCN_LIB: the instrumented build (build.py)


class ConvDescC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("N", "H", "W", "Cin", "Cout", "kh", "kw", "sh", "sw", "dh", "dw", "pad_t", "pad_l",
                 "Ho", "Wo")]


class OptTensorC(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p),
                ("v", ctypes.c_void_p), ("ema", ctypes.c_void_p), ("w_bf16", ctypes.c_void_p),
                ("w_bf16_t", ctypes.c_void_p), ("n", ctypes.c_longlong), ("taps", ctypes.c_int),
                ("cin", ctypes.c_int), ("cout", ctypes.c_int), ("l2", ctypes.c_float),
                ("wd", ctypes.c_float), ("l1", ctypes.c_float)]


_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float,
      "d": ctypes.c_double, "D": ctypes.POINTER(ConvDescC)}

# name -> argument codes WITHOUT the trailing stream (every entry ends with a void* stream)
SIGNATURES = {
    "mcn_conv2d_fprop_tc": "Dppppiii",
    "mcn_conv2d_fprop_tc_stats": "Dppppip",
    "mcn_conv2d_dgrad_tc": "Dpppiii",
    "mcn_conv2d_wgrad_tc": "Dpppi",
    "mcn_conv2d_fprop_direct": "Dipippp",
    "mcn_conv2d_dgrad_direct": "Dipipp",
    "mcn_conv2d_wgrad_direct": "Dippp",
    "mcn_dwconv2d_fwd": "Diipipp",
    "mcn_dwconv2d_bwd_data": "Diipipp",
    "mcn_dwconv2d_bwd_filter": "Diippp",
    "mcn_stem_conv_fprop": "Dppppp",
    "mcn_stem_conv_wgrad": "Dppp",
    "mcn_pad_rgb4": "plp",
    "mcn_weight_prep": "piiipp",
    "mcn_im2col": "Dippi",
    "mcn_bn_stats": "iplip",
    "mcn_bn_finalize": "pdiffpppp",
    "mcn_bn_frozen_stats": "ppifpp",
    "mcn_bn_apply": "iplipppppifp",
    "mcn_bn_apply_stats": "iplipdffpppifppppp",
    "mcn_bn_infer": "iplippfpppifp",
    "mcn_bn_bwd_reduce": "ippplippppifpp",
    "mcn_bn_apply_stats_mask": "iplipdffpppifpppppp",
    "mcn_bn_bwd_reduce_mask": "ippplipppp",
    "mcn_bn_bwd_apply_mask": "ippplipppppdpp",
    "mcn_bn_bwd_apply": "ippplippppifppdpp",
    "mcn_maxpool_fwd": "ipiiiiiiiiiiiipp",
    "mcn_maxpool_bwd": "ippiiiiiiiiiiiip",
    "mcn_maxpool_fwd_tap": "ipiiiiiiiiiiiipp",
    "mcn_maxpool_bwd_tap": "ippiiiiiiiiiiiip",
    "mcn_maxpool_tap_to_argmax": "piiiiiiiiiiiip",
    "mcn_avgpool_fwd": "ipiiiiiiiiiiiip",
    "mcn_avgpool_bwd": "ipiiiiiiiiiiiip",
    "mcn_gap_fwd": "ipiiipi",
    "mcn_gap_bwd": "ipiiiip",
    "mcn_act_fwd": "iplifp",
    "mcn_act_bwd": "ipplifp",
    "mcn_add_act_fwd": "ipplifp",
    "mcn_add_act_bwd": "ipplifp",
    "mcn_accumulate": "ippl",
    "mcn_scale_bcast_fwd": "ippiiip",
    "mcn_scale_bcast_bwd": "ipppiiipp",
    "mcn_bias_add": "iplip",
    "mcn_bias_grad": "iplip",
    "mcn_cast": "ipipl",
    "mcn_input_prep": "piiiiiiiffip",
    "mcn_copy_channels": "ipliipiiii",
    "mcn_resize_bilinear_fwd": "ipiiiiiiip",
    "mcn_resize_bilinear_bwd": "ipiiiiiiip",
    "mcn_resize_nearest_fwd": "ipiiiiiiip",
    "mcn_resize_nearest_bwd": "ipiiiiiiip",
    "mcn_dropout": "iplfpip",
    "mcn_sd_add_fwd": "ippilfpiifp",
    "mcn_sd_add_bwd": "ippilfpiifpp",
    "mcn_softmax_xent": "pplipfffiifppp",
    "mcn_sigmoid_xent": "plfffppi",
    "mcn_opt_step": "ipilppp",
    "mcn_grad_sqnorm": "pilpp",
    "mcn_transpose_add_f32": "piiip",
    "mcn_peer_allreduce": "plllpipipipii",
    "mcn_xsum_decode": "pippi",
    "mcn_conv2d_dgrad_tc_bnred": "Dpppipppppippp",
    "mcn_bn_bwd_finalize": "pppipp",
    "mcn_gn_fwd": "ipiliifpppp",
    "mcn_gn_bwd": "ippiliipppppp",
    "mcn_ws_fwd": "piifpp",
    "mcn_ws_bwd": "pppiifp",
    "mcn_fill_f32": "plf",
    "mcn_scale_f32": "plf",
}

_lib = None


def load():
    """Load libmcn.so and bind every entry point.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmcn.so not found at %s — build it with `python -m myconvnet_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, codes in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.argtypes = [_T[c] for c in codes] + [ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.mcn_last_error.restype = ctypes.c_char_p
    lib.mcn_last_error.argtypes = []
    lib.mcn_version.restype = ctypes.c_int
    lib.mcn_launch_count.restype = ctypes.c_longlong
    lib.mcn_stem_conv_kpad.restype = ctypes.c_int
    lib.mcn_stem_conv_kpad.argtypes = [ctypes.POINTER(ConvDescC)]
    lib.mcn_debug_role_cycles.restype = ctypes.c_int
    lib.mcn_debug_role_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.mcn_conv2d_dgrad_bnred_supported.restype = ctypes.c_int
    lib.mcn_conv2d_dgrad_bnred_supported.argtypes = [ctypes.POINTER(ConvDescC), ctypes.c_int]
    lib.mcn_set_workspace.restype = ctypes.c_int
    lib.mcn_set_workspace.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
    lib.mcn_workspace_min_bytes.restype = ctypes.c_longlong
    lib.mcn_workspace_min_bytes.argtypes = []
    lib.mcn_conv2d_wgrad_workspace_bytes.restype = ctypes.c_longlong
    lib.mcn_conv2d_wgrad_workspace_bytes.argtypes = [ctypes.POINTER(ConvDescC), ctypes.c_int, ctypes.c_int]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libmcn %s failed (%d): %s" % (what, rc, load().mcn_last_error().decode()))


def launch_count():
    return int(load().mcn_launch_count())


def xsum_value(limbs):
    """Host decode of xsum accumulators (include/mcn.h): limbs is an int64 array [3][n] (or [3]);
    exact Python-integer arithmetic, returned as float(s)."""
    import numpy as np
    a = np.asarray(limbs, dtype=np.int64).reshape(3, -1)
    out = [float((int(l0) + (int(l1) << 40) + (int(l2) << 80)) / (1 << 80)) if abs(int(l2)) < (1 << 61)
           else float("nan") for l0, l1, l2 in zip(a[0], a[1], a[2])]
    return out[0] if len(out) == 1 else np.array(out)


_workspaces = {}      # device index -> list of workspace tensors, newest last (none is ever freed:
                      # captured CUDA graphs keep raw pointers into the one they were recorded with)


def ensure_workspace(nbytes=0, device=None):
    """Allocate (once per device, growing on demand), zero and register the reduction workspace
    every deterministic kernel needs (mcn_set_workspace).  Returns the torch tensor holding it.
    One workspace per device is shared by everything in the process that runs on the current
    stream (launches sharing it must be stream-ordered)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = load()
    need = max(int(nbytes), int(lib.mcn_workspace_min_bytes()))
    held = _workspaces.setdefault(dev.index, [])
    if not held or held[-1].numel() < need + 256:
        torch.cuda.synchronize(dev)
        held.append(torch.zeros(need + 256, dtype=torch.uint8, device=dev))
    use_workspace(held[-1])
    return held[-1]


def use_workspace(t):
    """Register an existing (zeroed) uint8 tensor as the current device's workspace."""
    import torch
    base = (t.data_ptr() + 255) // 256 * 256
    with torch.cuda.device(t.device):
        check(load().mcn_set_workspace(base, t.numel() - (base - t.data_ptr())), "set_workspace")
