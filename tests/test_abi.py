"""The C-ABI library loads and exports exactly what include/mcn.h declares; lib.py's ctypes
signature table matches the header's parameter lists (no compute calls: runs without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from myconvnet_b200 import build
    return build.build()


def header_decls():
    src = open(os.path.join(ROOT, "include", "mcn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(mcn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        decls[m.group(2)] = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
    return decls


def code_of(arg):
    if "mcn_conv_desc" in arg:
        return "D"
    if "*" in arg:
        return "p"
    if arg.startswith("long long"):
        return "l"
    if arg.startswith("float"):
        return "f"
    if arg.startswith("double"):
        return "d"
    if arg.startswith("int"):
        return "i"
    raise AssertionError("unknown parameter type: " + arg)


def test_exports_match_header(built_lib):
    decls = header_decls()
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib]).decode()
    exported = set(re.findall(r" T (mcn_\w+)", out))
    assert set(decls) == exported, (set(decls) ^ exported)


def test_ctypes_table_matches_header(built_lib):
    from myconvnet_b200 import lib
    decls = header_decls()
    for name, codes in lib.SIGNATURES.items():
        args = decls[name]
        assert args[-1].replace(" ", "") == "void*stream", name
        got = "".join(code_of(a) for a in args[:-1])
        assert got == codes, "%s: header %s vs lib.py %s" % (name, got, codes)
    missing = set(decls) - set(lib.SIGNATURES) - {"mcn_last_error", "mcn_version", "mcn_launch_count",
                                                        "mcn_stem_conv_kpad", "mcn_set_workspace",
                                                        "mcn_workspace_min_bytes",
                                                        "mcn_conv2d_wgrad_workspace_bytes", "mcn_debug_role_cycles",
                                                        "mcn_conv2d_dgrad_bnred_supported"}
    assert not missing, missing
    L = lib.load()
    assert L.mcn_version() >= 100
    assert ctypes.sizeof(lib.OptTensorC) == 88 and ctypes.sizeof(lib.ConvDescC) == 60


def test_sass_is_blackwell_native(built_lib):
    """tcgen05 / TMA evidence in the shipped binary (B200_PROFILING.md mnemonics)."""
    sass = subprocess.check_output(["cuobjdump", "-sass", built_lib]).decode()
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTMALDG.4D.IM2COL", "UTMASTG", "UTMAREDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass      # no legacy mma.sync path
    # determinism: no floating-point atomic / reduction instruction anywhere in the library (the
    # exact accumulators of xsum.cuh use 64-bit integer atomics; bf16 accumulation is TMA reduce-add)
    import re
    assert not re.search(r"(ATOMG|REDG|ATOM|RED)\.E\.ADD\.F(16|32|64)", sass)
