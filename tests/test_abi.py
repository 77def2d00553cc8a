"""The C-ABI library loads without a GPU and exports every symbol include/hoomd_b200.h declares; the
product fails loudly (no CPU fallback) when there is no device; the host-only helpers work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "hoomd_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from hoomd_blue_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 60
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/hoomd_b200.h but not exported"
    assert L.hb2_abi_version() == 2


def test_ctypes_binding_covers_header():
    """hoomd_blue_b200/_lib.py declares argtypes for every entry point that takes arguments."""
    from hoomd_blue_b200 import _lib
    L = _lib.lib()
    for n in declared_symbols():
        if n in ("hb2_last_error", "hb2_abi_version", "hb2_launch_count"):
            continue
        assert getattr(L, n).argtypes is not None, f"{n} has no ctypes signature"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from hoomd_blue_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    st = L.hb2_system_create(C.byref(h), 0, 1, 0)
    assert st == _lib.ERR_NO_DEVICE
    assert b"no CPU path" in L.hb2_last_error()
    import hoomd_blue_b200 as hoomd
    with pytest.raises(RuntimeError):
        hoomd.device.CPU()


def test_host_helpers():
    from hoomd_blue_b200 import _lib
    L = _lib.lib()
    b = _lib.Box()
    assert L.hb2_box_make(C.byref(b), 4.0, 6.0, 8.0) == 0
    assert list(b.hi) == [2.0, 3.0, 4.0] and list(b.lo) == [-2.0, -3.0, -4.0]
    assert L.hb2_box_make(C.byref(b), -1.0, 1.0, 1.0) == _lib.ERR_INVALID
    # LJ packing: lj1 = 4 eps sigma^12, lj2 = 4 eps sigma^6 (hoomd/md/EvaluatorPairLJ.h:119-140)
    user = (C.c_double * 2)(1.5, 0.8)
    out = (C.c_float * _lib.PARAM_FLOATS)()
    assert L.hb2_pack_params(_lib.EVAL["lj"], user, 2, out) == 0
    assert np.allclose(list(out)[:2], [4 * 1.5 * 0.8 ** 12, 4 * 1.5 * 0.8 ** 6], rtol=1e-6)
    # Mie packs four: prefactor eps sigma^n, prefactor eps sigma^m, n, m (hoomd/md/EvaluatorPairMie.h:71-82)
    user4 = (C.c_double * 4)(1.0, 1.0, 12.0, 6.0)
    assert L.hb2_pack_params(_lib.EVAL["mie"], user4, 4, out) == 0
    assert np.allclose(list(out)[:4], [4.0, 4.0, 12.0, 6.0], rtol=1e-6)
    assert L.hb2_pack_params(_lib.EVAL["mie"], user4, 3, out) == _lib.ERR_INVALID
    assert L.hb2_pack_params(_lib.EVAL["lj"], user, 1, out) == _lib.ERR_INVALID
