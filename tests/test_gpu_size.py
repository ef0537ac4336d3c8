"""Parity of the fused strip path (the benchmark path) against the oracle AT SIZE (SURVEY.md 8d names 256^2 and
1024^2; VERDICT r1 asks for the full-size mesh too).

  * 256^2 (200 steps) and 1024^2 (20 steps): the whole mesh, every node, fault + ABSORB + source, from a seeded
    non-trivial state (s2d_cart_fill_fields) -- relative L2 <= 1e-10 on d and v.
  * 8192^2 (the benchmark mesh, npoin*ndof > 2^31): windows of the mesh run by the oracle with the same medium,
    state, fault and absorbing sides; compared on the nodes whose domain of dependence stays inside the window
    (tests/window.py).  The windows sit on the fault's nucleation patch, in the top-right corner (largest node
    indices, beyond 2^31 for the second component) and in the bottom-left corner.
"""
import numpy as np
import pytest

import window
from sem2dpack_b200 import CartEngine

pytestmark = pytest.mark.gpu
SEED = 20261017
FILL = (777, 1.0e-3, 1.0)
H = window.H
HALF_NUC = 1537.0   # not a multiple of the GLL spacing: no node sits on the patch edge (harness.nuc_radius)


def _bench_engine(nx, nz, nsteps, scheme_kind=0, src=None, coef_mode=0):
    e = CartEngine(5, 2, nx, nz, (0.0, nx * H), (0.0, nz * H), ezflt=nz // 2, seed=SEED, scheme_kind=scheme_kind,
                   courant=0.5, coef_mode=coef_mode)
    e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * H / 2, HALF_NUC, oixd=1, oitd=1, nt_max=nsteps)
    for s in (1, 2, 3, 4):
        e.add_abso_side(s, False)
    if src is not None:
        e.add_force_at(src[0], src[1], [-0.5, 0.8660254037844386])
    e.commit()
    e.fill_fields(*FILL)
    return e


@pytest.mark.parametrize("n,nsteps,scheme", [(256, 200, "leapfrog"), (256, 100, "newmark"), (1024, 20, "leapfrog")])
def test_whole_mesh_vs_oracle_at_size(n, nsteps, scheme):
    src = (0.37 * n * H, 0.61 * n * H)
    e = _bench_engine(n, n, nsteps, 0 if scheme == "leapfrog" else 1, src)
    w = window.Window(0, 0, n, n, n, n, n // 2, e.dt, nsteps, SEED, FILL, scheme=scheme, src=src, half_nuc=HALF_NUC)
    assert w.o.i("npoin") == e.npoin
    e.step(nsteps, w.stf_table(nsteps))
    ed, ev, nn = w.compare(e)
    assert nn == e.npoin
    assert ed <= 1e-10 and ev <= 1e-10, (ed, ev)
    # the fault has slipped inside the nucleation patch (the run exercises the friction solve)
    st = e.fault_state(0, w.o.i("bc.0.np"))
    assert np.abs(st["D"]).max() > 0.0
    assert np.allclose(st["D"], w.o.arr("bc.0.D"), rtol=0, atol=1e-10 * max(np.abs(st["D"]).max(), 1e-30))
    w.close()
    e.close()


def test_benchmark_mesh_windows_vs_oracle():
    """the 8192 x 8192 mesh of BASELINE.json configs[4] itself, 6 steps, three windows"""
    N, k = 8192, 6
    ez = N // 2
    src = (N * H / 2 + 330.0, ez * H + 710.0)
    e = _bench_engine(N, N, k, 0, src)
    assert e.npoin * 2 > 2 ** 31
    kw = dict(half_nuc=HALF_NUC)
    wins = [window.Window(ez - 24, ez - 24, 48, 48, N, N, ez, e.dt, k, SEED, FILL, src=src, half_nuc=HALF_NUC),     # fault + source
            window.Window(N - 40, N - 40, 40, 40, N, N, ez, e.dt, k, SEED, FILL, **kw),                 # top-right corner
            window.Window(0, 0, 40, 40, N, N, ez, e.dt, k, SEED, FILL, **kw),                           # bottom-left corner
            window.Window(2037, ez - 20, 44, 40, N, N, ez, e.dt, k, SEED, FILL, **kw)]                  # fault away from the patch
    e.step(k, wins[0].stf_table(k))
    for w in wins:
        ed, ev, nn = w.compare(e)
        assert nn > 10000
        assert ed <= 1e-10 and ev <= 1e-10, (w.X0, w.Z0, ed, ev)
        w.close()
    vmax, dmax = e.progress()
    assert np.isfinite(vmax) and vmax > 0.1
    e.close()
