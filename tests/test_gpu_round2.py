"""Round-2 regression tests (ADVICE r1 + VERDICT r1 rows a20, a9): kinetic energy against the oracle on both
engines, sources that share nodes, stepping past nt_max, restoring accelerations on their own."""
import numpy as np
import pytest

import harness
import orc
from harness import Rig, rel_l2
from sem2dpack_b200 import CartEngine, Engine

pytestmark = pytest.mark.gpu
SEED = 20261017


@pytest.mark.parametrize("name", ["lamb", "ratestate"])
def test_energy_generic_engine(name):
    """energy_compute (energy.f90:49-106): E_k = 1/2 sum(w rho v.v) over the elements == 1/2 sum(M v.v) over the
    nodes with the assembled mass of MAT_MASS_init; s2d_energy after a run vs the oracle's element loop"""
    o = orc.Oracle(harness.deck(name))
    r = Rig(o)
    r.e.set_mass(o.arr("mass"))
    n = 250
    o.step(n)
    r.step(n)
    L = o.L
    ek_ref = L.orc_energy_Ek(o.h)
    ek = r.e.energy()
    assert ek_ref > 0
    assert abs(ek - ek_ref) <= 1e-10 * ek_ref, (ek, ek_ref)
    r.close()


@pytest.mark.parametrize("set_mass", [False, True])
def test_energy_builder_engine(set_mass):
    """builder-made engines keep their fields on the lattice: the mass must be there too (set_mass
    uploaded it in the caller's order).  Without s2d_set_mass the builder's own assembled mass is used."""
    nx, nz, ez, nsteps = 24, 16, 8, 120
    o = orc.Oracle(harness.cart_deck(nx, nz, ezflt=ez, nsteps=nsteps), synthetic_seed=SEED, renumber=False)
    e = CartEngine(5, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=ez, seed=SEED, scheme_kind=0, courant=0.5)
    e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * 50.0, harness.nuc_radius(nx), nt_max=nsteps)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side, False)
    e.add_force_at(0.37 * nx * 100.0, 0.61 * nz * 100.0, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    if set_mass:
        e.set_mass(o.arr("mass"))
    e.commit()
    e.step(nsteps, np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)]))
    o.step(nsteps)
    ek_ref = o.L.orc_energy_Ek(o.h)
    ek = e.energy()
    assert ek_ref > 0 and abs(ek - ek_ref) <= 1e-10 * ek_ref, (ek, ek_ref)
    e.close()
    o.close()


def test_sources_sharing_a_node_add_up_in_order():
    """SO_add (src_gen.f90:290-317) adds the sources one after the other; two point forces on one node and a
    moment source whose terms overlap them must neither race nor depend on the launch"""
    o = orc.Oracle(harness.deck("lamb"))
    ngll, npoin = o.i("ngll"), o.i("npoin")
    ib = o.arr("ibool").reshape(-1, ngll, ngll)
    node = int(ib[5, 3, 4])
    rng = np.random.default_rng(4)
    mnodes = np.concatenate([ib[5, 3, :], ib[5, :, 4], ib[5, 3, :]]).astype(np.int32)   # the node appears three times
    mcoef = rng.standard_normal((2, mnodes.size))
    nsteps = 40
    amp = rng.standard_normal((nsteps, 3))
    outs = []
    for rep in range(2):
        e = Engine(ngll, 2, o.arr("ibool"), o.arr("H"), o.arr("rmass"), 0, o.f("dt"))
        e.set_elastic(o.i("nelast"), o.arr("a"), o.arr("elem2set"), False)
        e.add_force(node, [0.3, -1.1])
        e.add_moment(mnodes, mcoef.ravel())
        e.add_force(node, [2.0, 0.7])
        e.commit()
        e.step(nsteps, amp)
        outs.append(e.get_fields())
        e.close()
    for x, y in zip(*outs):
        assert np.array_equal(x, y)          # deterministic
    # the same load as independent single-node forces (one per term): equal up to the order of the additions
    e = Engine(ngll, 2, o.arr("ibool"), o.arr("H"), o.arr("rmass"), 0, o.f("dt"))
    e.set_elastic(o.i("nelast"), o.arr("a"), o.arr("elem2set"), False)
    nodes_all = [node] + [int(n) for n in mnodes] + [node]
    dirs = [[0.3, -1.1]] + [[mcoef[0, t], mcoef[1, t]] for t in range(mnodes.size)] + [[2.0, 0.7]]
    cols = [0] + [1] * mnodes.size + [2]
    for nd, dr in zip(nodes_all, dirs):
        e.add_force(nd, dr)
    e.commit()
    e.step(nsteps, amp[:, cols])
    ref = e.get_fields()
    e.close()
    for x, y in zip(outs[0], ref):
        assert np.abs(y).max() > 0
        assert rel_l2(x, y) <= 1e-13
    o.close()


def test_stepping_past_nt_max_keeps_the_fault_history_in_bounds():
    """BC_DYNFLT_write's potency table holds nt_max + 1 lines; further steps are not recorded and
    s2d_get_fault never copies more than that"""
    nx, nz, ez, nt_max = 12, 8, 4, 10
    e = CartEngine(5, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=ez, seed=SEED, scheme_kind=0, courant=0.5)
    fid = e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * 50.0, 337.0, nt_max=nt_max)
    e.commit()
    e.fill_fields(3, 1e-3, 1.0)
    e.step(nt_max + 25, None)
    rec, pot = e.fault(fid, nx * 4 + 1)
    assert pot.shape[0] == nt_max + 1
    assert rec.shape[0] <= nt_max + 2
    assert np.isfinite(pot).all() and np.isfinite(rec).all()
    e.close()


def test_restoring_accel_alone_invalidates_the_cached_prediction():
    """fused explicit Newmark caches d[n+1] = d + dt v + dt^2/2 a; s2d_set_fields(NULL, NULL, accel) must drop it"""
    nx, nz = 14, 9
    es = [CartEngine(5, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), seed=SEED, scheme_kind=1, courant=0.5)
          for _ in range(2)]
    for e in es:
        e.commit()
        e.fill_fields(9, 1e-3, 1.0)
        e.step(5, None)
    d, v, a = es[0].get_fields()
    a2 = a * 1.5
    es[0].set_fields(d, v, a2)          # everything at once
    es[1].set_fields(None, None, a2)    # only the accelerations (d and v are already there)
    outs = []
    for e in es:
        e.step(7, None)
        outs.append(e.get_fields())
        e.close()
    for x, y in zip(*outs):
        assert np.array_equal(x, y)
