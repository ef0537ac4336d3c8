import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import harness, orc
from harness import Rig, rel_l2
from sem2dpack_b200.engine import detect_structured
o = orc.Oracle(harness.deck("tpv3"))
n = o.i("npoin") * 2
rng = np.random.default_rng(8)
d0 = rng.standard_normal(n)
o.set_fields(d0, np.zeros(n))
ref = o.compute_fint()
box = detect_structured(6, o.arr("ibool"), o.i("npoin"))
r = Rig(orc.Oracle(harness.deck("tpv3")))
r.e.set_fields(d0, np.zeros(n))
for rep in range(3):
    f = r.e.compute_fint()
    dif = np.abs(f - ref)
    npn = o.i("npoin")
    bad = np.nonzero(dif > 1e-9 * np.abs(ref).max())[0]
    nodes = bad % npn
    print("rep", rep, "route", r.e.route(), "rel", rel_l2(f, ref), "nbad", bad.size, "finite", np.isfinite(f).all(),
          "gx", (box["gx"][nodes].min(), box["gx"][nodes].max()) if bad.size else None, "gz", (box["gz"][nodes].min(), box["gz"][nodes].max()) if bad.size else None)
kv = o.arr("kv_elem"); print("nkv", kv.size, "nelem", o.i("nelem"), "eta finite", np.isfinite(o.arr("kv_eta")).all(), "eta max", o.arr("kv_eta").max(), "min", o.arr("kv_eta").min())
