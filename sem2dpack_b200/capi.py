"""ctypes binding of include/sem2d_b200.h -- the same entry points the ISO_C_BINDING shim binds.

There is no CPU fallback: `lib()` raises if the CUDA library has not been built, and every engine
call raises `S2DError` (the text the Fortran shim would pass to IO_abort, SRC/stdio.f90:205-214)
when the C-ABI returns a non-zero status.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# S2D_LIB_PATH: a differently tuned build of the same library (kernel experiments only)
LIB_PATH = os.environ.get("S2D_LIB_PATH") or os.path.join(_HERE, "lib", "libsem2d_b200.so")
_LIB = None

S2D_ASM_PATCH, S2D_ASM_COLOR, S2D_ASM_ATOMIC = 0, 1, 2
LEAPFROG, NEWMARK, HHT_ALPHA, SYMPLECTIC = 0, 1, 2, 3

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


class S2DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sem2d_b200 error {code}: {msg}")
        self.code = code


MAX_STAGES = 8  # S2D_MAX_STAGES


class Scheme(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dt", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
                ("alpha", C.c_double), ("nstages", C.c_int32), ("coa", C.c_double * (MAX_STAGES + 1)),
                ("cob", C.c_double * MAX_STAGES)]


class DynfltDesc(C.Structure):
    _fields_ = [
        ("np", C.c_int32), ("node1", _PI), ("node2", _PI), ("n1", _PD), ("B", _PD), ("invM1", _PD),
        ("invM2", _PD), ("Z", _PD), ("T0", _PD), ("cohesion", _PD), ("coord", _PD), ("V0", _PD),
        ("CoefA2V", C.c_double), ("CoefA2D", C.c_double), ("allow_opening", C.c_int32),
        ("swf_kind", C.c_int32), ("swf_healing", C.c_int32),
        ("swf_dc", _PD), ("swf_mus", _PD), ("swf_mud", _PD), ("swf_p", _PD), ("swf_alpha", _PD), ("swf_theta", _PD),
        ("rsf_kind", C.c_int32),
        ("rsf_dc", _PD), ("rsf_mus", _PD), ("rsf_a", _PD), ("rsf_b", _PD), ("rsf_Vstar", _PD), ("rsf_theta", _PD),
        ("rsf_Vc", _PD),
        ("twf_kind", C.c_int32),
        ("twf_X", C.c_double), ("twf_Z", C.c_double), ("twf_mus", C.c_double), ("twf_mud", C.c_double),
        ("twf_mu0", C.c_double), ("twf_L", C.c_double), ("twf_V", C.c_double), ("twf_T", C.c_double),
        ("twf_Dc", C.c_double),
        ("normal_kind", C.c_int32), ("normal_T", C.c_double), ("normal_L", C.c_double), ("normal_V", C.c_double),
        ("oix1", C.c_int32), ("oixn", C.c_int32), ("oixd", C.c_int32), ("oit", C.c_int32), ("oitd", C.c_int32),
        ("nt_max", C.c_int32),
    ]


HALO_IPC_BYTES = 192  # S2D_HALO_IPC_BYTES


class CartDesc(C.Structure):
    _fields_ = [
        ("ngll", C.c_int32), ("ndof", C.c_int32), ("nx", C.c_int32), ("nz", C.c_int32), ("ezflt", C.c_int32),
        ("x0", C.c_double), ("x1", C.c_double), ("z0", C.c_double), ("z1", C.c_double),
        ("seed", C.c_uint64), ("ix0", C.c_int64), ("iz0", C.c_int64),
        ("rho", C.c_double), ("cp", C.c_double), ("cs", C.c_double),
        ("precision", C.c_int32), ("scheme", Scheme), ("courant", C.c_double), ("device", C.c_int32),
        ("halo_left", C.c_int32), ("halo_right", C.c_int32), ("coef_mode", C.c_int32), ("renumber", C.c_int32),
    ]


# every symbol include/sem2d_b200.h declares (checked by tests/test_capi_symbols.py)
_SIGS = {
    "s2d_create": [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_int32, C.POINTER(Scheme), C.c_int32],
    "s2d_destroy": [C.c_void_p],
    "s2d_set_elastic": [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32],
    "s2d_set_kv": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_set_mass": [C.c_void_p, C.c_void_p],
    "s2d_add_abso": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                     C.c_void_p, C.c_void_p],
    "s2d_add_dirneu": [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_add_dynflt": [C.c_void_p, C.POINTER(DynfltDesc), C.POINTER(C.c_int32)],
    "s2d_add_force": [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_add_periodic": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_cart_add_periodic": [C.c_void_p, C.c_int32, C.c_int32],
    "s2d_add_moment": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_add_receivers": [C.c_void_p, C.c_int32, C.c_char, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                          C.c_void_p, C.c_void_p],
    "s2d_commit": [C.c_void_p, C.c_int32],
    "s2d_set_fields": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_get_fields": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_step": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_compute_fint": [C.c_void_p, C.c_void_p],
    "s2d_get_it": [C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_get_seis": [C.c_void_p, C.c_void_p],
    "s2d_get_seis_row": [C.c_void_p, C.c_int32, C.c_void_p],
    "s2d_get_fault": [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_get_fault_state": [C.c_void_p, C.c_int32] + [C.c_void_p] * 7,
    "s2d_progress": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "s2d_energy": [C.c_void_p, C.POINTER(C.c_double)],
    "s2d_energy_w25d": [C.c_void_p, C.POINTER(C.c_double)],
    "s2d_get_coloring": [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p],
    "s2d_time_fint": [C.c_void_p, C.c_int32, C.POINTER(C.c_float)],
    "s2d_time_steps": [C.c_void_p, C.c_int32, C.POINTER(C.c_float)],
    "s2d_kernel_ms": [C.c_void_p, C.POINTER(C.c_float)],
    "s2d_rcm_box": [C.c_int32, C.c_int32, C.c_void_p],
    "s2d_kernel_route": [C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_detect_structured": [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                              C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_time_phases": [C.c_void_p, C.c_int32, C.c_void_p],
    "s2d_launch_count": [C.c_void_p, C.POINTER(C.c_int64)],
    "s2d_stream": [C.c_void_p, C.POINTER(C.c_void_p)],
    "s2d_cart_create": [C.POINTER(C.c_void_p), C.POINTER(CartDesc)],
    "s2d_cart_add_abso": [C.c_void_p, C.c_int32, C.c_int32],
    "s2d_cart_add_fault_swf": [C.c_void_p] + [C.c_double] * 8 + [C.c_int32] * 3 + [C.POINTER(C.c_int32)],
    "s2d_cart_add_force": [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_cart_add_receivers": [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_char,
                               C.c_int32, C.c_int32],
    "s2d_cart_fault_info": [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_cart_fault_nodes": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p],
    "s2d_cart_add_dynflt": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(DynfltDesc), C.POINTER(C.c_int32)],
    "s2d_cart_add_dirneu": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32],
    "s2d_cart_set_kv": [C.c_void_p, C.c_void_p],
    "s2d_cart_add_receivers_interp": [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_char,
                                      C.c_int32, C.c_int32],
    "s2d_cart_add_moment": [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int32)],
    "s2d_cart_receiver_info": [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p],
    "s2d_cart_set_material": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_cart_set_kv_elems": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_cart_set_w25d": [C.c_void_p, C.c_double],
    "s2d_cart_set_plastic": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_cart_get_plastic_strain": [C.c_void_p, C.c_void_p],
    "s2d_cart_set_damage": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "s2d_cart_get_damage_state": [C.c_void_p, C.c_void_p],
    "s2d_cart_set_visco": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_cart_info": [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)],
    "s2d_cart_set_dt": [C.c_void_p, C.c_double],
    "s2d_cart_snapshot_elem": [C.c_void_p, C.c_char, C.c_void_p],
    "s2d_cart_get_gll": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_cart_get": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_cart_fill_fields": [C.c_void_p, C.c_uint64, C.c_double, C.c_double],
    "s2d_cart_get_window": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_halo_info": [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p],
    "s2d_halo_set_exchange": [C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_halo_set_peers": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "s2d_halo_peer_buffers": [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)],
    "s2d_halo_ipc_export": [C.c_void_p, C.c_void_p],
    "s2d_halo_ipc_open": [C.c_void_p, C.c_void_p, C.c_void_p],
}
_STR_FUNCS = ("s2d_last_error", "s2d_version")


def lib():
    """Load libsem2d_b200.so; raise if it is missing (there is no fallback path)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). sem2dpack_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.s2d_last_error.argtypes = [C.c_void_p]
        L.s2d_last_error.restype = C.c_char_p
        L.s2d_version.argtypes = []
        L.s2d_version.restype = C.c_char_p
        _LIB = L
    return _LIB


def declared_symbols():
    return sorted(list(_SIGS) + list(_STR_FUNCS))


def _f64(x):
    return None if x is None else np.ascontiguousarray(x, dtype=np.float64)


def _i32(x):
    return None if x is None else np.ascontiguousarray(x, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pd(a):
    return None if a is None else a.ctypes.data_as(_PD)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_PI)
