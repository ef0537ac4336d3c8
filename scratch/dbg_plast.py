import sys, os
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import numpy as np, orc, harness
from sem2dpack_b200 import CartEngine
ngll, nx, nz, ezflt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 0
os.environ["S2D_SEG"] = sys.argv[4]
h = 100.0
coh, phi, Tv, e0 = float(os.environ.get('COH','2.0e6')), 30.0, 0.02, tuple(float(os.environ.get('E0S','1'))*q for q in (-4.0e-4, -3.0e-4, 2.5e-4))
L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='plastic', verbose='0000', ItInfo=1000 /",
     "&MESH_DEF method='CARTESIAN' /",
     f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz} /",
     "&MATERIAL tag=1, kind='PLAST' /",
     f"&MAT_PLASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0, phi={phi}d0, coh={coh}d0, Tv={Tv}d0, e0={e0[0]}d0,{e0[1]}d0,{e0[2]}d0 /",
     "&TIME NbSteps=10, courant=0.5d0, kind='leapfrog' /"]
o = orc.Oracle("\n".join(L) + "\n", renumber=False)
e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=0, rho=2670.0, cp=6000.0, cs=3464.0)
e.set_dt(o.f("dt"))
e.set_plastic([[coh, phi, Tv, *e0]], np.ones(nx * nz, np.int32))
e.commit()
rng = np.random.default_rng(1)
d = 1e-3 * rng.standard_normal(e.npoin * 2)
e.set_fields(d, d); o.set_fields(d, d)
ref = o.compute_fint(); got = e.compute_fint()
n = e.npoin
co = o.arr("coord").reshape(n, 2)
err = np.abs(got - ref).reshape(2, n).max(axis=0)
bad = np.where(err > 1e-9 * np.abs(ref).max())[0]
print("nbad", bad.size, "of", n, "maxerr", err.max() / np.abs(ref).max())
xs = np.unique(np.round(co[bad, 0] / h, 3)); zs = np.unique(np.round(co[bad, 1] / h, 3))
print("x (elements):", xs[:40]); print("z (elements):", zs[:40])
ep_ref = o.arr("pl_ep").reshape(nx * nz, 3, ngll, ngll); ep = e.plastic_strain()
de = np.abs(ep - ep_ref).reshape(nz, nx, -1).max(axis=2)
print("bad ep elements (iz, ix):", np.argwhere(de > 1e-12 * np.abs(ep_ref).max())[:40].tolist())
ix = np.round(co[:, 0] / h, 3); iz = np.round(co[:, 1] / h, 3)
sel = bad[:60]
for b in sel[:24]:
    print("x %.3f z %.3f err %.2e %.2e ref %.3e %.3e" % (ix[b], iz[b], got[b] - ref[b], got[n + b] - ref[n + b], ref[b], ref[n + b]))
