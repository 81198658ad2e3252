"""The C-ABI shared library loads without a GPU and exports exactly the entry points include/cstp_b200.h declares;
argument validation fails loudly with CSTP_EINVAL + a message (no compute is attempted here)."""
import ctypes as C
import os
import re

import pytest

from cstp_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cstp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cstp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(L.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.cstp_version() >= 100
    assert lib.cstp_launch_count() >= 0


def test_struct_sizes_match_header_limits():
    assert (L.CSTP_MAX_AMAPS, L.CSTP_MAX_TAPS, L.CSTP_MAX_MCHUNKS) == (4, 32, 96)
    assert C.sizeof(L.Tensor5) == 8 + 5 * 4 + 4 + 4 * 8      # ptr, dims, padding, strides
    assert C.sizeof(L.Tap) == 20 and C.sizeof(L.MChunk) == 20


def test_invalid_arguments_fail_loudly():
    lib = L.load()
    d = L.ConvDesc()          # all zero: n_amaps == 0 is invalid
    h = C.c_void_p()
    rc = lib.cstp_conv_plan_create(C.byref(d), C.byref(h))
    assert rc == -1
    assert b"invalid argument" in lib.cstp_last_error()
    with pytest.raises(L.CstpError):
        L.check(rc)
    w = L.WgradDesc()
    assert lib.cstp_wgrad_plan_create(C.byref(w), C.byref(h)) == -1
    assert lib.cstp_ema_update(None, None, 0, 0.5, 0.5, None) == -1
    assert lib.cstp_ntxent(None, 3, 8, 0.1, 1, None, None, None, 0, None) == -1


def test_product_path_rejects_cpu_tensors():
    import torch
    from cstp_b200 import ops
    from cstp_b200.loss.NTXent import NTXentLoss
    with pytest.raises(L.CstpError):
        ops.conv_fwd_plan(torch.zeros(1, 1, 1, 128, 16, dtype=torch.bfloat16), torch.zeros(16, 64, dtype=torch.bfloat16),
                          torch.zeros(1, 1, 1, 128, 16, dtype=torch.bfloat16), ops.ConvGeom((1, 1, 1)))
    with pytest.raises(L.CstpError):
        NTXentLoss("cpu", 4, 0.1, True)(torch.randn(4, 8), torch.randn(4, 8))


def test_bn_sync_exchange_argument_checks():
    lib = L.load()
    assert lib.cstp_bn_sync_buffer_bytes(8, 4, 16384) == 4 * 8 * 16384 * 4 + 4 * 8 * 4
    assert lib.cstp_bn_sync_buffer_bytes(0, 4, 16) == -1
    assert lib.cstp_bn_sync_exchange(None, 16, None, 2, 0, 4, 64, 1, None, None, None, None) == -1
    assert b"invalid argument" in lib.cstp_last_error()


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under cstp_b200/ (nor the native arm of bench.py) may import it."""
    pkg = os.path.join(ROOT, "cstp_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
