import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from sem2dpack_b200 import CartEngine
SEED = 20261017
ngll, nx, nz, ezflt = 5, 19, 13, 6
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
xl, zl = (0.0, nx * 100.0), (0.0, nz * 130.0)
outs = {}
for rep in range(2):
  for m in (0, 1):
    e = CartEngine(ngll, 2, nx, nz, xl, zl, ezflt=ezflt, seed=SEED, coef_mode=m)
    d = np.random.default_rng(11).standard_normal(e.npoin * 2)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side)
    e.commit()
    e.set_fields(d * 1e-3, d)
    f0 = e.compute_fint()
    e.step(nsteps, None)
    outs[(rep, m)] = (f0,) + tuple(e.get_fields())
    e.close()
names = ["f0", "d", "v", "a"]
for k in range(4):
    a, b = outs[(0, 0)][k], outs[(0, 1)][k]
    print(names[k], "compact vs full: maxdiff", np.abs(a - b).max(), "scale", np.abs(a).max(), "nbad", (a != b).sum(),
          "| repeat compact", (outs[(0, 0)][k] != outs[(1, 0)][k]).sum(), "repeat full", (outs[(0, 1)][k] != outs[(1, 1)][k]).sum())
