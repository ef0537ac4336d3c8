"""Plastic box for the ncu capture of the PLAST instantiation (profiles/run_ncu_r2_variants.sh): n x n elements, NGLL 5,
the 2.5D_plastic material, absorbing sides, seeded state.  usage: bench_plastic.py n amp_d amp_v"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sem2dpack_b200 import CartEngine
n = int(sys.argv[1]); h = 100.0
e = CartEngine(5, 2, n, n, (0.0, n * h), (0.0, n * h), ezflt=n // 2, seed=0, rho=2705.0, cp=5770.0, cs=3330.0)
e.set_plastic([[9.669501e6, 30.0, 0.0561, -4.162378e-04, -4.162378e-04, 3.504802e-04]], np.ones(n * n, np.int32))
for s in (1, 2, 3, 4):
    e.add_abso_side(s)
e.commit()
e.fill_fields(7, float(sys.argv[2]), float(sys.argv[3]))
e.time_steps(3)
ms = e.time_steps(10)
print("n", n, "ms/step %.3f" % (ms / 10), "G DOF/s %.2f" % (e.npoin * 2 * 10 / ms / 1e6), "kernel_ms %.3f" % e.kernel_ms())
