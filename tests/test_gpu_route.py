"""Kernel routing of the generic C-ABI (VERDICT r1 "what's weak" 1): what a Fortran host hands over through
s2d_create / s2d_set_elastic -- RCM-ordered ibool, per-element coefficient blocks, boundary tables in its own
node numbering -- must reach the z-marching strip kernel when the mesh is a MESH_CART box, with the same results
as the any-mesh patch kernel and the oracle."""
import numpy as np
import pytest

import harness
import orc
from harness import Rig, rel_l2

pytestmark = pytest.mark.gpu
SEED = 20261017


def _run(deck, nsteps, seed=0, chunk=97):
    o = orc.Oracle(deck, synthetic_seed=seed)
    r = Rig(o)
    done = 0
    while done < nsteps:
        n = min(chunk, nsteps - done)
        o.step(n)
        r.step(n)
        done += n
    return o, r


@pytest.mark.parametrize("name,nsteps,tol", [("testsh", 600, 1e-10), ("lamb", 600, 1e-10), ("ratestate", 400, 1e-9),
                                             ("inabox", 500, 1e-10)])
def test_reference_decks_take_the_strip_kernel(name, nsteps, tol):
    o, r = _run(harness.deck(name), nsteps)
    assert r.e.route() == 1
    d, v, a = r.e.get_fields()
    for nm, got in (("d", d), ("v", v), ("acc", a)):
        assert rel_l2(got, o.arr(nm)) <= tol, (nm, rel_l2(got, o.arr(nm)))
    if o.i("rec.present"):
        s_ref, s_got = o.seis(), r.e.seis()
        assert np.abs(s_got - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    for fid, ibc, np_, onx in r.faults:
        st = r.e.fault_state(fid, np_)
        for k in ("D", "V", "T", "MU"):
            assert rel_l2(st[k], o.arr(f"bc.{ibc}.{k}")) <= tol, k
    r.close()


@pytest.mark.parametrize("scheme,stacey,nx,nz,ezflt", [("leapfrog", False, 40, 24, 16), ("newmark", True, 26, 21, 9),
                                                      ("leapfrog", False, 100, 70, 35)])
def test_synthetic_family_rcm_order_through_the_generic_api(scheme, stacey, nx, nz, ezflt):
    """heterogeneous medium (one a(5,5,6) block per element), two-sided fault on the split-node row, RCM element
    order: the route north_star describes"""
    nsteps = 300
    o, r = _run(harness.cart_deck(nx, nz, ezflt=ezflt, scheme=scheme, stacey=stacey, nsteps=nsteps), nsteps, seed=SEED)
    assert r.e.route() == 1
    d, v, a = r.e.get_fields()
    for nm, got in (("d", d), ("v", v), ("acc", a)):
        assert rel_l2(got, o.arr(nm)) <= 1e-10, nm
    s_ref, s_got = o.seis(), r.e.seis()
    assert np.abs(s_got - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    fid, ibc, np_, onx = r.faults[0]
    st = r.e.fault_state(fid, np_)
    for k, floor in (("D", 1e-3), ("V", 1e-3), ("T", 1.0), ("MU", 1e-3)):   # a locked fault slips at rounding level
        ref = o.arr(f"bc.{ibc}.{k}")
        assert np.abs(st[k] - ref).max() <= 1e-10 * max(np.abs(ref).max(), floor), k
    if nx >= 40:
        assert np.abs(st["D"]).max() > 1e-3
    rec, pot = r.e.fault(fid, onx)
    ref = o.arr(f"bc.{ibc}.out").reshape(-1, 6, onx)
    for c in range(6):
        assert np.abs(rec[:, c] - ref[:, c]).max() <= 2e-7 * max(np.abs(ref[:, c]).max(), 1e-30)
    r.close()


def test_routing_can_be_switched_off_and_both_kernels_agree(monkeypatch):
    deck = harness.cart_deck(30, 20, ezflt=10, nsteps=150)
    outs = []
    for flag, route in (("1", 1), ("0", 0)):
        monkeypatch.setenv("S2D_ROUTE_STRIP", flag)
        o, r = _run(deck, 150, seed=SEED)
        assert r.e.route() == route
        d, v, a = r.e.get_fields()
        assert rel_l2(d, o.arr("d")) <= 1e-10
        f = r.e.compute_fint()
        assert rel_l2(f, o.compute_fint()) <= 1e-12
        nc, col = r.e.coloring()       # the colouring stays available whatever kernel runs
        assert nc >= 4 and col.size == o.i("nelem")
        outs.append((d, v, a))
        r.close()
    for x, y in zip(*outs):
        assert rel_l2(x, y) <= 1e-11


def test_kelvin_voigt_elements_take_the_strip_kernel():
    """EXAMPLES/TestFlt2D_SCEC_TPV3_inplane: ELAST + KV elements (eta(ngll,ngll) per element, mat_kelvin_voigt.f90:137-150),
    Newmark, one-sided fault: k_elem_strip<KV> forms d + eta*v element by element"""
    o, r = _run(harness.deck("tpv3"), 500)
    assert r.e.route() == 1
    d, v, a = r.e.get_fields()
    for nm, got in (("d", d), ("v", v), ("acc", a)):
        assert rel_l2(got, o.arr(nm)) <= 1e-10, nm
    rng = np.random.default_rng(8)
    n = o.i("npoin") * 2
    d0, v0 = rng.standard_normal(n), rng.standard_normal(n)
    r.e.set_fields(d0, v0)
    o.set_fields(d0, v0)
    assert rel_l2(r.e.compute_fint(), o.compute_fint()) <= 1e-13
    r.close()


def test_general_planes_keep_the_any_mesh_kernel():
    """nelast = 10 (curved-mesh planes, mat_elastic.f90:344-358) is not the strip kernel's flat form"""
    from sem2dpack_b200 import Engine
    o = orc.Oracle(harness.cart_deck(9, 7, ngll=5, ndof=2, nrec=0, src=False), synthetic_seed=SEED)
    ag = np.zeros((o.i("ncoefsets"), 10, 25))
    ag[:, :6] = o.arr("a").reshape(o.i("ncoefsets"), 6, 25)
    e = Engine(5, 2, o.arr("ibool"), o.arr("H"), o.arr("rmass"), 0, o.f("dt"))
    e.set_elastic(10, ag, o.arr("elem2set"), False)
    e.commit()
    assert e.route() == 0
    e.close()
    o.close()


def test_fields_energy_and_set_fields_in_the_callers_numbering():
    """a routed handle still speaks the caller's node numbering: set_fields / get_fields round trip, s2d_energy
    with the caller's mass"""
    o = orc.Oracle(harness.cart_deck(19, 13, ezflt=6, nsteps=50), synthetic_seed=SEED)
    r = Rig(o)
    assert r.e.route() == 1
    rng = np.random.default_rng(5)
    n = o.i("npoin") * 2
    d0, v0 = rng.standard_normal(n) * 1e-3, rng.standard_normal(n)
    r.e.set_fields(d0, v0)
    o.set_fields(d0, v0)
    d1, v1, _ = r.e.get_fields()
    assert np.array_equal(d1, d0) and np.array_equal(v1, v0)
    r.e.set_mass(o.arr("mass"))
    assert abs(r.e.energy() - o.L.orc_energy_Ek(o.h)) <= 1e-12 * o.L.orc_energy_Ek(o.h)
    o.step(50)
    r.step(50)
    d, v, _ = r.e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10 and rel_l2(v, o.arr("v")) <= 1e-10
    r.close()


@pytest.mark.parametrize("name,nsteps,route_flag", [("inplane25d", 300, "1"), ("kvfz", 300, "1"), ("kvfz", 200, "0")])
def test_finite_seismogenic_width_decks(name, nsteps, route_flag, monkeypatch):
    """&GENERAL W (2.5D): MAT_ELAST_init_25D / MAT_ELAST_add_25D_f (mat_elastic.f90:363-383,447-459): every element
    force gets - beta*d, with the Kelvin-Voigt-modified d on KV elements (mat_gen.f90:435-440).  EXAMPLES/2.5D_inplane
    (P-SV, SWF + TWF fault, leapfrog) and EXAMPLES/Kelvin_Visco_FZ (SH, two tags, KV layer with ETAxDT=F, Newmark),
    on the strip kernel and on the any-mesh patch kernel.  No reference artefact pins these decks: oracle parity."""
    monkeypatch.setenv("S2D_ROUTE_STRIP", route_flag)
    o, r = _run(harness.deck(name), nsteps)
    assert o.arr("beta25d").size > 0 and r.e.route() == int(route_flag)
    d, v, a = r.e.get_fields()
    assert np.abs(o.arr("d")).max() > 0
    for nm, got in (("d", d), ("v", v), ("acc", a)):
        assert rel_l2(got, o.arr(nm)) <= 1e-10, (nm, rel_l2(got, o.arr(nm)))
    rng = np.random.default_rng(12)
    n = o.i("npoin") * o.i("ndof")
    d0, v0 = rng.standard_normal(n), rng.standard_normal(n)
    r.e.set_fields(d0, v0)
    o.set_fields(d0, v0)
    assert rel_l2(r.e.compute_fint(), o.compute_fint()) <= 1e-13
    r.close()


@pytest.mark.parametrize("scheme", ["newmark", "leapfrog"])
@pytest.mark.parametrize("kv_fused", ["1", "0"])
def test_kelvin_voigt_inside_the_fused_step(scheme, kv_fused, monkeypatch):
    """MAT_KV_add_etav (mat_kelvin_voigt.f90:137-150; solver.f90:293-295): the element force is taken from
    d + eta*v with the PREDICTED velocity.  On the strip kernel the node update rides in the same launch (v and,
    for Newmark, a are double-buffered because neighbours read them); S2D_KV_FUSED=0 keeps the separate
    predictor / corrector passes.  TPV3 deck (P-SV, NGLL 6, KV layer, one-sided SWF fault, ABSORB + DIRNEU)."""
    monkeypatch.setenv("S2D_KV_FUSED", kv_fused)
    deck = harness.deck("tpv3")
    if scheme == "leapfrog":
        deck = deck.replace("kind='newmark'", "kind='leapfrog'")
        assert "kind='leapfrog'" in deck

    def check():
        d, v, a = r.e.get_fields()
        assert np.abs(o.arr("d")).max() > 0
        for nm, got in (("d", d), ("v", v), ("acc", a)):
            assert rel_l2(got, o.arr(nm)) <= 1e-10, (nm, rel_l2(got, o.arr(nm)))
        for fid, ibc, np_, onx in r.faults:
            st = r.e.fault_state(fid, np_)
            for k in ("D", "V", "T"):
                ref = o.arr(f"bc.{ibc}.{k}")
                assert np.abs(st[k] - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-12), k

    o, r = _run(deck, 650)
    assert r.e.route() == 1 and o.i("nkv") > 0 and r.faults
    check()
    n_launch = r.e.launch_count()
    o.step(40)   # a second call continues from the swapped buffers
    r.step(40)
    check()
    if kv_fused == "1":
        assert (r.e.launch_count() - n_launch) <= 7 * 40  # no separate predictor / corrector passes
    else:
        assert (r.e.launch_count() - n_launch) > 7 * 40
    r.close()
