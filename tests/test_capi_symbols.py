"""CPU tests of the drop-in boundary: libsem2d_b200.so loads without a GPU, exports every function
include/sem2d_b200.h declares (and the ctypes table binds exactly that set), and refuses to run a
problem when there is no CUDA device -- there is no CPU fallback to fall into."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sem2dpack_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sem2d_b200.h")


def _declared_in_header():
    with open(HEADER) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(s2d_[a-z0-9_]+)\s*\(", txt))
    names -= {"s2d_exchange_fn"}
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = _declared_in_header()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sem2d_b200.h but not exported"


def test_ctypes_table_matches_header():
    assert sorted(capi.declared_symbols()) == _declared_in_header()


def test_version_string():
    assert b"sm_100a" in capi.lib().s2d_version()


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device behaviour")
def test_no_device_no_fallback():
    """s2d_create / s2d_cart_create must fail with S2D_ENODEV on a machine without a GPU."""
    L = capi.lib()
    ngll = 3
    ibool = np.arange(1, 10, dtype=np.int32)
    H = np.zeros(9)
    rmass = np.ones(9)
    sch = capi.Scheme(0, 1e-3, 0.0, 0.5, 1.0)
    h = C.c_void_p()
    rc = L.s2d_create(C.byref(h), ngll, 1, 1, 9, ibool.ctypes.data, H.ctypes.data, rmass.ctypes.data, 8,
                      C.byref(sch), -1)
    assert rc == -2 and not h.value
    d = capi.CartDesc()
    d.ngll, d.ndof, d.nx, d.nz = 5, 2, 4, 4
    d.x0, d.x1, d.z0, d.z1 = 0.0, 1.0, 0.0, 1.0
    d.rho, d.cp, d.cs = 1.0, 2.0, 1.0
    d.precision = 8
    d.scheme = sch
    d.courant = 0.5
    d.device = -1
    assert L.s2d_cart_create(C.byref(h), C.byref(d)) == -2


def test_argument_validation_precedes_device_use():
    L = capi.lib()
    h = C.c_void_p()
    sch = capi.Scheme(0, 1e-3, 0.0, 0.5, 1.0)
    assert L.s2d_create(C.byref(h), 2, 1, 1, 9, None, None, None, 8, C.byref(sch), -1) == -1  # S2D_EINVAL
    assert L.s2d_destroy(None) == -1
    assert L.s2d_step(None, 1, None, None) == -1


def test_header_is_plain_c_and_structs_match_the_bindings(tmp_path):
    """include/sem2d_b200.h is the C-ABI: it must compile as C99 (no C++ or torch types) and its structs must
    have the layout the ctypes bindings (and the Fortran bind(C) types of INTEGRATION.md) assume."""
    import ctypes as C
    import subprocess
    from sem2dpack_b200 import capi
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include "sem2d_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu\\n", sizeof(s2d_scheme), sizeof(s2d_cart_desc), '
                   'sizeof(s2d_dynflt_desc)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes == [C.sizeof(capi.Scheme), C.sizeof(capi.CartDesc), C.sizeof(capi.DynfltDesc)]
