"""ctypes access to the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = {}


def build_oracle(force=False, variant="parity"):
    """variant "parity": -O2 -ffp-contract=off (what the parity tests compare against); "o3": -O3 -march=x86-64-v3
    (bench.py's CPU legs only: BASELINE.md section 4's optimisation level)"""
    so = os.path.join(ORACLE_DIR, "_build", "liboracle.so" if variant == "parity" else "liboracle_o3.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in
            ("oracle_capi.cpp", "sem2d_oracle.hpp", "gll.hpp", "rcm.hpp", "parinp.hpp")]
    stale = (not os.path.exists(so)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib(variant="parity"):
    if variant not in _LIB:
        L = C.CDLL(build_oracle(variant=variant))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.c_ulonglong, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.orc_create_at.restype = C.c_void_p
        L.orc_create_at.argtypes = [C.c_char_p, C.c_ulonglong, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                    C.c_char_p, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.orc_time_solve.restype = C.c_double
        L.orc_time_solve.argtypes = [C.c_void_p, C.c_int]
        L.orc_compute_fint.argtypes = [C.c_void_p]
        L.orc_set_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_snapshot.argtypes = [C.c_void_p, C.c_int]
        L.orc_energy_EW.restype = C.c_double
        L.orc_energy_EW.argtypes = [C.c_void_p]
        L.orc_energy_Ek.restype = C.c_double
        L.orc_energy_Ek.argtypes = [C.c_void_p]
        L.orc_get_int.restype = C.c_longlong
        L.orc_get_int.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_double.restype = C.c_double
        L.orc_get_double.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_stf.restype = C.c_double
        L.orc_stf.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_array.restype = C.c_longlong
        L.orc_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p]
        L.orc_gll.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_rcm.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_hash_u.restype = C.c_double
        L.orc_hash_u.argtypes = [C.c_ulonglong] * 4
        _LIB[variant] = L
    return _LIB[variant]


_DT = {b"i": np.int32, b"d": np.float64, b"f": np.float32}


class Oracle:
    """One SEM2DPACK problem built from Par.inp text and advanced on the CPU."""

    def __init__(self, parinp_text, synthetic_seed=0, renumber=True, kd_force_kd1=False, lattice_origin=(0, 0),
                 variant="parity"):
        """lattice_origin: GLL lattice coordinates of this mesh's corner inside a larger synthetic mesh (the hash
        medium is a function of global lattice coordinates), for windowed parity checks"""
        self.L = lib(variant)
        err = C.create_string_buffer(512)
        self.h = self.L.orc_create_at(parinp_text.encode(), synthetic_seed, int(renumber), int(kd_force_kd1),
                                      int(lattice_origin[0]), int(lattice_origin[1]), err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())

    @classmethod
    def from_file(cls, path, **kw):
        with open(path) as f:
            return cls(f.read(), **kw)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def i(self, name):
        v = self.L.orc_get_int(self.h, name.encode())
        if v == -999:
            raise KeyError(name)
        return int(v)

    def f(self, name):
        v = self.L.orc_get_double(self.h, name.encode())
        if v != v:
            raise KeyError(name)
        return float(v)

    def arr(self, name, copy=True):
        p = C.c_void_p()
        dt = C.create_string_buffer(2)
        n = self.L.orc_array(self.h, name.encode(), C.byref(p), dt)
        if n < 0:
            raise KeyError(name)
        dtype = _DT[dt.value[:1]]
        if n == 0:
            return np.zeros(0, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype)
        return a.copy() if copy else a

    def step(self, n=1):
        err = C.create_string_buffer(512)
        if self.L.orc_step(self.h, n, err, 512) != 0:
            raise RuntimeError(err.value.decode())

    def time_solve(self, n):
        return self.L.orc_time_solve(self.h, n)

    def compute_fint(self):
        self.L.orc_compute_fint(self.h)
        return self.arr("fint")

    def snapshot(self, what):
        """PLOT_FIELD's element-wise field 'E' | 'S' | 'd' | 'c': float32 (ncomp, nelem, ngll, ngll)"""
        self.L.orc_snapshot(self.h, ord(what))
        n = self.i("ngll")
        return self.arr("snap").reshape(-1, self.i("nelem"), n, n)

    def set_fields(self, d=None, v=None, a=None):
        def p(x):
            if x is None:
                return None
            x = np.ascontiguousarray(x, dtype=np.float64)
            keep.append(x)
            return x.ctypes.data
        keep = []
        self.L.orc_set_fields(self.h, p(d), p(v), p(a))

    def stf(self, i, t):
        return self.L.orc_stf(self.h, i, t)

    def seis(self):
        """(nt, nx, ndof) float32 seismograms, as REC_write would dump them (receivers.f90:351-392)."""
        nt, nx, nd = self.i("rec.nt"), self.i("rec.nx"), self.i("ndof")
        return self.arr("rec.sis").reshape(nd, nx, nt).transpose(2, 1, 0)


def gll(n):
    x = np.zeros(n)
    w = np.zeros(n)
    H = np.zeros((n, n))
    lib().orc_gll(n, x.ctypes.data, w.ctypes.data, H.ctypes.data)
    return x, w, H.T.copy()  # H[ip, ix] (column-major in C buffer)


def rcm(nx, nz):
    p = np.zeros(nx * nz, np.int32)
    lib().orc_rcm(nx, nz, p.ctypes.data)
    return p


REFERENCE_EXAMPLES = "/root/reference/EXAMPLES"
