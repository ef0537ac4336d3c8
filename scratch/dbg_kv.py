import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import harness, orc
from harness import Rig, rel_l2
from sem2dpack_b200.engine import detect_structured
o = orc.Oracle(harness.deck("tpv3"))
n = o.i("npoin") * 2
rng = np.random.default_rng(8)
d0, v0 = rng.standard_normal(n), rng.standard_normal(n)
o.set_fields(d0, v0)
ref = o.compute_fint()
box = detect_structured(6, o.arr("ibool"), o.i("npoin"))
for mode in ("v-random", "v-zero"):
    vv = v0 if mode == "v-random" else np.zeros(n)
    o.set_fields(d0, vv)
    ref = o.compute_fint()
    for flag in ("1", "0"):
        os.environ["S2D_ROUTE_STRIP"] = flag
        r = Rig(orc.Oracle(harness.deck("tpv3")))
        r.e.set_fields(d0, vv)
        f = r.e.compute_fint()
        dif = np.abs(f - ref)
        npn = o.i("npoin")
        bad = np.nonzero(dif > 1e-9 * np.abs(ref).max())[0]
        print(mode, "route", r.e.route(), "rel", rel_l2(f, ref), "nbad", bad.size)
        if bad.size:
            nodes = bad % npn
            print("  gx range", box["gx"][nodes].min(), box["gx"][nodes].max(), "gz range", box["gz"][nodes].min(), box["gz"][nodes].max())
            print("  gz hist", np.unique(box["gz"][nodes], return_counts=True))
            print("  gx mod 25 hist", np.unique(box["gx"][nodes] % 25, return_counts=True))
        r.close()
