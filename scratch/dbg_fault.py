import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import window
from sem2dpack_b200 import CartEngine
SEED = 20261017
H = 100.0
nx, nz, nsteps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fill = (777, 1e-3, 1.0)
ez = nz // 2
e = CartEngine(5, 2, nx, nz, (0.0, nx * H), (0.0, nz * H), ezflt=ez, seed=SEED, scheme_kind=0, courant=0.5)
fid = e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * H / 2, 1537.0, oixd=1, oitd=1, nt_max=nsteps)
for s in (1, 2, 3, 4):
    e.add_abso_side(s, False)
e.commit()
e.fill_fields(*fill)
w = window.Window(0, 0, nx, nz, nx, nz, ez, e.dt, nsteps, SEED, fill)
o = w.o
np_f = o.i("bc.0.np")
import ctypes as C
cnt = C.c_int32(); 
co = np.empty(2 * np_f); T0 = np.empty(2 * np_f); B = np.empty(np_f)
e._ck(e.L.s2d_cart_fault_info(e.h, fid, C.byref(cnt), co.ctypes.data_as(C.c_void_p), T0.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p)))
print("np", cnt.value, np_f)
for name, got, ref in (("coord", co, o.arr("bc.0.coord")), ("T0", T0, o.arr("bc.0.T0")), ("B", B, o.arr("bc.0.B")[:np_f])):
    dif = np.abs(got - ref)
    print(name, "maxdiff", dif.max(), "at", int(dif.argmax()), "n differing", int((dif > 1e-9 * np.abs(ref).max()).sum()))
bad = np.nonzero(np.abs(T0 - o.arr("bc.0.T0")) > 1.0)[0]
print("T0 differing idx", bad[:20], "engine", T0[bad[:6]], "oracle", o.arr("bc.0.T0")[bad[:6]], "x", co[2 * (bad[:6] % np_f)])
e.step(nsteps, None)
o.step(nsteps)
d, v = e.get_window(0, 0, w.LXw, w.LZw)
npoin = o.i("npoin")
ov = o.arr("v")
for c in range(2):
    ref = ov[c * npoin + w.lat]
    dif = np.abs(v[c] - ref)
    idx = np.dstack(np.unravel_index(np.argsort(dif.ravel())[::-1][:8], dif.shape))[0]
    print("comp", c, "max", dif.max(), "scale", np.abs(ref).max(), "worst (gz,gx):", idx.tolist())
st = e.fault_state(fid, np_f)
for k in ("D", "V", "T", "MU", "sigma"):
    ref = o.arr("bc.0." + k)
    dif = np.abs(st[k] - ref)
    print(k, "maxdiff", dif.max(), "scale", np.abs(ref).max(), "at", int(dif.argmax()))
