import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import window
from sem2dpack_b200 import CartEngine
SEED = 20261017
H = 100.0
def run(nx, nz, seg, nsteps, fill=(777, 1e-3, 1.0), fault=True, abso=True, src=True, scheme=0):
    os.environ["S2D_SEG"] = str(seg)
    ez = nz // 2 if fault else 0
    e = CartEngine(5, 2, nx, nz, (0.0, nx * H), (0.0, nz * H), ezflt=ez, seed=SEED, scheme_kind=scheme, courant=0.5)
    if fault:
        e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * H / 2, 1537.0, oixd=1, oitd=1, nt_max=nsteps)
    if abso:
        for s in (1, 2, 3, 4):
            e.add_abso_side(s, False)
    sp = (0.37 * nx * H, 0.61 * nz * H) if src else None
    if src:
        e.add_force_at(sp[0], sp[1], [-0.5, 0.8660254037844386])
    e.commit()
    if fill:
        e.fill_fields(*fill)
    w = window.Window(0, 0, nx, nz, nx, nz, ez, e.dt, nsteps, SEED, fill or (1, 0.0, 0.0), scheme="leapfrog" if scheme == 0 else "newmark", src=sp)
    if not abso:
        pass
    e.step(nsteps, w.stf_table(nsteps) if src else None)
    ed, ev, nn = w.compare(e)
    print(f"nx {nx} nz {nz} seg {seg} steps {nsteps} fill {bool(fill)} fault {fault} abso {abso} src {src}: ed {ed:.3e} ev {ev:.3e}", flush=True)
    w.close(); e.close()
for args in [(24, 16, 32, 50), (24, 16, 4, 50), (24, 80, 32, 50), (100, 16, 32, 50), (100, 80, 32, 50), (256, 256, 32, 1), (256, 256, 32, 10), (256, 256, 32, 50),
             (256, 256, 32, 200)]:
    run(*args)
run(256, 256, 32, 200, fill=None)
run(256, 256, 32, 200, fault=False)
run(256, 256, 32, 200, src=False)
