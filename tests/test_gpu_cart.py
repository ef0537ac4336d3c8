"""Structured (MESH_CART) device-side builder against the oracle run with OPT_RENUMBER=.false.
(natural element order): bit-exact ibool, coefficient planes / inverse mass to rounding, and the
whole time loop of the synthetic benchmark family at test size."""
import numpy as np
import pytest

import harness
import orc
from harness import rel_l2
from sem2dpack_b200 import CartEngine

pytestmark = pytest.mark.gpu
SEED = 20261017


def _oracle(nx, nz, **kw):
    return orc.Oracle(harness.cart_deck(nx, nz, **kw), synthetic_seed=kw.pop("seed", SEED), renumber=False)


@pytest.mark.parametrize("ngll,nx,nz,ezflt", [(5, 8, 8, 0), (5, 19, 13, 0), (5, 19, 21, 8), (5, 17, 16, 5),
                                             (6, 11, 9, 4), (9, 7, 8, 3), (3, 30, 30, 15), (4, 5, 3, 1)])
def test_ibool_bit_exact(ngll, nx, nz, ezflt):
    o = orc.Oracle(harness.cart_deck(nx, nz, ngll=ngll, ezflt=ezflt, nrec=0, src=False, abso=(), fault=None),
                   renumber=False)
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=ezflt, rho=2670.0, cp=6000.0, cs=3464.0)
    assert e.npoin == o.i("npoin")
    ib, _, _, co = e.get_tables(ibool=True, rmass=False, coord=True)
    assert np.array_equal(ib, o.arr("ibool"))
    assert np.abs(co - o.arr("coord")).max() <= 1e-9
    e.close()
    o.close()


@pytest.mark.parametrize("ndof", [1, 2])
@pytest.mark.parametrize("seed", [0, SEED])
def test_operator_tables(ndof, seed):
    nx, nz = 19, 13
    o = orc.Oracle(harness.cart_deck(nx, nz, ndof=ndof, nrec=0, src=False, abso=()), synthetic_seed=seed, renumber=False)
    e = CartEngine(5, ndof, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), seed=seed, rho=2670.0, cp=6000.0, cs=3464.0)
    assert abs(e.dt - o.f("dt")) <= 1e-13 * e.dt
    _, a, rm, _ = e.get_tables(ibool=False, a=(seed != 0), rmass=True)
    assert rel_l2(rm, o.arr("rmass")) <= 1e-13
    if seed != 0:
        assert rel_l2(a, o.arr("a")) <= 1e-13
    d = np.random.default_rng(3).standard_normal(e.npoin * ndof)
    e.commit()
    e.set_fields(d, d)
    o.set_fields(d, d)
    assert rel_l2(e.compute_fint(), o.compute_fint()) <= 1e-12
    e.close()
    o.close()


@pytest.mark.parametrize("scheme,stacey,nx,nz,ezflt", [("leapfrog", False, 24, 16, 8), ("newmark", True, 19, 21, 9),
                                                      ("leapfrog", False, 40, 24, 16)])
def test_synthetic_benchmark_family(scheme, stacey, nx, nz, ezflt):
    """heterogeneous P-SV box, two-sided SWF fault on tags 5,6 with a nucleation patch, absorbing
    sides 1-4, force source, receivers: 300 steps, builder-made engine vs oracle."""
    nsteps = 300
    h = 100.0
    o = orc.Oracle(harness.cart_deck(nx, nz, ezflt=ezflt, scheme=scheme, stacey=stacey, nsteps=nsteps),
                   synthetic_seed=SEED, renumber=False)
    kind = 0 if scheme == "leapfrog" else 1
    e = CartEngine(5, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=SEED, scheme_kind=kind, courant=0.5)
    assert abs(e.dt - o.f("dt")) <= 1e-13 * e.dt
    # same order as the deck: DYNFLT first, then ABSORB 1..4 (bc_gen.f90:229-246)
    fid = e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * h / 2, harness.nuc_radius(nx, h), nt_max=nsteps)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side, stacey)
    e.add_force_at(0.37 * nx * h, 0.61 * nz * h, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.add_receiver_line(8, (0.1 * nx * h, 0.3 * nz * h), (0.9 * nx * h, 0.8 * nz * h), "V", 1, nsteps + 1)
    e.commit()
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    e.step(nsteps, tab)
    o.step(nsteps)
    d, v, a = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10
    assert rel_l2(v, o.arr("v")) <= 1e-10
    assert rel_l2(a, o.arr("acc")) <= 1e-10
    s_ref, s_got = o.seis(), e.seis()
    assert np.abs(s_got - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    np_f = o.i("bc.0.np")
    st = e.fault_state(fid, np_f)
    for k in ("D", "V", "T", "MU"):
        assert rel_l2(st[k], o.arr("bc.0." + k)) <= 1e-10, k
    assert np.abs(st["D"]).max() > 1e-3
    rec, pot = e.fault(fid, np_f)
    assert rec.shape[0] == nsteps + 1
    ref = o.arr("bc.0.out").reshape(-1, 6, np_f)
    for c in range(6):
        assert np.abs(rec[:, c] - ref[:, c]).max() <= 2e-7 * max(np.abs(ref[:, c]).max(), 1e-30)
    e.close()
    o.close()


@pytest.mark.parametrize("ngll,nx,nz,ezflt", [(5, 19, 13, 6), (6, 11, 9, 4), (9, 7, 8, 0)])
def test_compact_coefficients_equal_stored_planes(ngll, nx, nz, ezflt):
    """coef_mode 0 keeps (lambda, mu) per GLL point; coef_mode 1 stores the six planes of MAT_ELAST_init_a
    (mat_elastic.f90:334-340,355-357).  The planes s2d_cart_get reports for the compact mode are formed with the
    reference's sequence of roundings and are BITWISE those of the stored mode; the compact kernel itself folds the
    constant metric factors of the flat grid into its derivative matrices (S2D_COMPACT_FOLD, strip_kernels.cuh), so
    forces and a whole run agree with the stored-plane mode to rounding (<= 1e-13), not bit for bit."""
    xl, zl = (0.0, nx * 100.0), (0.0, nz * 130.0)   # non-square elements: KD2's a4*(..) shortcut differs from KD1
    es = [CartEngine(ngll, 2, nx, nz, xl, zl, ezflt=ezflt, seed=SEED, coef_mode=m) for m in (0, 1)]
    planes = [e.get_tables(ibool=False, a=True, rmass=False)[1] for e in es]
    assert np.array_equal(planes[0], planes[1])
    d = np.random.default_rng(11).standard_normal(es[0].npoin * 2)
    out = []
    for e in es:
        for side in (1, 2, 3, 4):
            e.add_abso_side(side)
        e.commit()
        e.set_fields(d * 1e-3, d)
        f0 = e.compute_fint()
        e.step(40, None)
        out.append((f0,) + tuple(e.get_fields()))
        e.close()
    for x, y in zip(*out):
        assert np.abs(y).max() > 0
        assert rel_l2(x, y) <= 1e-13


@pytest.mark.parametrize("coef_mode", [0, 1])
def test_accel_policy_and_modes(coef_mode, monkeypatch):
    """fields%accel is materialised on the last step of every s2d_step call (S2D_STORE_ACCEL=2, the
    default) -- the same values as when it is written on every step, for either coefficient mode."""
    nx, nz, nsteps = 24, 16, 60
    o = orc.Oracle(harness.cart_deck(nx, nz, ezflt=8, nsteps=nsteps), synthetic_seed=SEED, renumber=False)
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    o.step(nsteps)
    res = []
    for policy in ("2", "1"):
        monkeypatch.setenv("S2D_STORE_ACCEL", policy)
        e = CartEngine(5, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=8, seed=SEED, coef_mode=coef_mode)
        e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * 50.0, harness.nuc_radius(nx), nt_max=nsteps)
        for side in (1, 2, 3, 4):
            e.add_abso_side(side)
        e.add_force_at(0.37 * nx * 100, 0.61 * nz * 100, [o.f("src.0.dir1"), o.f("src.0.dir2")])
        e.commit()
        e.step(25, tab[:25])
        e.step(nsteps - 25, tab[25:])
        res.append(e.get_fields())
        e.close()
    for x, y in zip(*res):
        assert np.array_equal(x, y)
    for x, k in zip(res[0], ("d", "v", "acc")):
        assert rel_l2(x, o.arr(k)) <= 1e-10, k
    o.close()


def test_fp32_builder():
    nx, nz, nsteps = 24, 16, 100
    o = orc.Oracle(harness.cart_deck(nx, nz, nsteps=nsteps, abso=(1, 2, 3, 4)), synthetic_seed=SEED, renumber=False)
    e = CartEngine(5, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), seed=SEED, precision=4)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side)
    e.add_force_at(0.37 * nx * 100, 0.61 * nz * 100, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.commit()
    e.step(nsteps, np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)]))
    o.step(nsteps)
    d, v, _ = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-5
    assert rel_l2(v, o.arr("v")) <= 1e-5
    e.close()


@pytest.mark.parametrize("ngll,ndof,nx,nz,ezflt,seg", [
    (5, 2, 13, 11, 0, 3),     # 3 strips (last one 1 element wide), 4 bands
    (5, 2, 12, 12, 5, 2),     # fault: bands 2+1 below, 4 above; nx multiple of the strip width
    (5, 1, 20, 9, 4, 4),      # SH
    (6, 2, 11, 10, 3, 3),     # 5 elements per strip
    (9, 2, 7, 6, 2, 2),       # 3 elements per strip
    (3, 2, 23, 8, 4, 5),      # 10 elements per strip
    (4, 1, 9, 7, 0, 2),
    (7, 2, 5, 5, 2, 1),       # one element row per band: every row boundary is a band halo
    (8, 1, 6, 4, 0, 32),
    (10, 2, 4, 5, 3, 2),
    (5, 2, 6, 40, 17, 32),    # single strip, tall
])
def test_strip_kernel_fint(ngll, ndof, nx, nz, ezflt, seg, monkeypatch):
    """compute_Fint of the z-marching strip kernel against the oracle on random fields, over strip /
    band decompositions that exercise every halo case (right-edge columns, band-top rows, their
    corners, the duplicated fault row, a narrower last strip)."""
    monkeypatch.setenv("S2D_SEG", str(seg))
    o = orc.Oracle(harness.cart_deck(nx, nz, ngll=ngll, ndof=ndof, ezflt=ezflt, nrec=0, src=False, abso=(), fault=None),
                   synthetic_seed=SEED, renumber=False)
    e = CartEngine(ngll, ndof, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=ezflt, seed=SEED)
    assert e.npoin == o.i("npoin")
    e.commit()
    rng = np.random.default_rng(ngll * 100 + nx)
    d = rng.standard_normal(e.npoin * ndof)
    e.set_fields(d, d)
    o.set_fields(d, d)
    ref = o.compute_fint()
    got = e.compute_fint()
    assert rel_l2(got, ref) <= 1e-13, rel_l2(got, ref)
    # fields round-trip through the lattice permutation unchanged
    dd, vv, _ = e.get_fields()
    assert np.array_equal(dd, d) and np.array_equal(vv, d)
    # deterministic: bitwise identical on a second evaluation
    assert np.array_equal(e.compute_fint(), got)
    e.close()
    o.close()


@pytest.mark.parametrize("ngll,ndof,nx,nz,ezflt,seg,scheme", [
    (5, 2, 27, 11, 0, 3, "leapfrog"),     # two full groups + a 1-strip remainder, 4 bands, no fault
    (5, 2, 50, 12, 5, 2, "newmark"),      # three groups (one partial), bands of 2 rows on both sides of the fault
    (5, 1, 31, 9, 4, 4, "leapfrog"),      # SH
    (5, 1, 26, 7, 3, 1, "newmark"),       # SH, one element row per band: every row boundary is shared
    (6, 2, 23, 10, 3, 3, "newmark"),      # 5 elements per strip (the TPV3 order)
    (9, 2, 14, 6, 2, 2, "leapfrog"),      # 3 elements per strip (the Lamb order)
    (3, 2, 47, 8, 4, 5, "leapfrog"),      # 10 elements per strip
    (5, 2, 6, 40, 17, 32, "newmark"),     # a single strip, tall
    (4, 2, 33, 5, 2, 64, "leapfrog"),     # band taller than the mesh
])
def test_fused_step_decompositions(ngll, ndof, nx, nz, ezflt, seg, scheme, monkeypatch):
    """the fused leapfrog / explicit Newmark step over strip, group and band decompositions that exercise
    every hand-over: strips of one CTA, columns finished by the second CTA to arrive, rows shared by
    two bands, their corners, deferred boundary rows and columns, the duplicated fault row"""
    monkeypatch.setenv("S2D_SEG", str(seg))
    nsteps, h = 80, 100.0
    o = orc.Oracle(harness.cart_deck(nx, nz, ngll=ngll, ndof=ndof, ezflt=ezflt, scheme=scheme, nsteps=nsteps, nrec=0,
                                     fault="swf" if ezflt else None),
                   synthetic_seed=SEED, renumber=False)
    e = CartEngine(ngll, ndof, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=SEED,
                   scheme_kind=0 if scheme == "leapfrog" else 1, courant=0.5)
    if ezflt:
        e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * h / 2, harness.nuc_radius(nx, h), nt_max=nsteps)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side, False)
    e.add_force_at(0.37 * nx * h, 0.61 * nz * h, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.commit()
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    e.step(30, tab[:30])
    e.step(nsteps - 30, tab[30:])
    o.step(nsteps)
    d, v, a = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10 and rel_l2(v, o.arr("v")) <= 1e-10 and rel_l2(a, o.arr("acc")) <= 1e-10
    assert np.abs(d).max() > 0
    e.close()
    o.close()


@pytest.mark.parametrize("field", ["A", "D"])
@pytest.mark.parametrize("scheme", ["leapfrog", "newmark"])
def test_fused_step_receiver_fields(field, scheme):
    """stations that record accelerations force the fused step to materialise them on every step (they are
    otherwise written on the last step of a call only); displacement stations read the buffer that holds
    d[n] after the two displacement buffers have swapped"""
    nx, nz, nsteps, h = 24, 16, 120, 100.0
    deck = harness.cart_deck(nx, nz, ezflt=8, scheme=scheme, nsteps=nsteps).replace("field='V'", f"field='{field}'")
    o = orc.Oracle(deck, synthetic_seed=SEED, renumber=False)
    e = CartEngine(5, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=8, seed=SEED,
                   scheme_kind=0 if scheme == "leapfrog" else 1, courant=0.5)
    e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * h / 2, harness.nuc_radius(nx, h), nt_max=nsteps)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side, False)
    e.add_force_at(0.37 * nx * h, 0.61 * nz * h, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.add_receiver_line(8, (0.1 * nx * h, 0.3 * nz * h), (0.9 * nx * h, 0.8 * nz * h), field, 1, nsteps + 1)
    e.commit()
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    e.step(nsteps, tab)
    o.step(nsteps)
    s_ref, s_got = o.seis(), e.seis()
    assert np.abs(s_ref).max() > 0
    assert np.abs(s_got - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    e.close()
    o.close()


@pytest.mark.parametrize("ngll,nx,nz,ezflt", [(5, 19, 13, 6), (6, 11, 9, 0), (3, 30, 7, 3), (9, 7, 8, 4), (5, 64, 40, 20)])
def test_rcm_order_ibool_bit_exact_and_fields_in_that_numbering(ngll, nx, nz, ezflt):
    """s2d_cart_desc.renumber = 1: OPT_RENUMBER = .true., the reference's DEFAULT (SRC/constants.f90:10-15).  Elements in
    the reverse Cuthill-McKee order of MESH_STRUCTURED_renumber (mesh_structured.f90:204-269, rcm.f90) and GLL nodes
    numbered by SE_init_numbering in that order: ibool bit for bit against the oracle's restatement, coordinates, and
    a run whose fields come back in that numbering."""
    deck = harness.cart_deck(nx, nz, ngll=ngll, ezflt=ezflt, nsteps=60, fault=None, nrec=0)
    o = orc.Oracle(deck, synthetic_seed=SEED, renumber=True)
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=ezflt, seed=SEED, scheme_kind=0,
                   courant=0.5, renumber=True)
    assert e.npoin == o.i("npoin")
    ib, a, rm, co = e.get_tables(ibool=True, a=True, rmass=True, coord=True)
    assert np.array_equal(ib, o.arr("ibool"))
    assert np.abs(co - o.arr("coord")).max() <= 1e-9
    assert rel_l2(a, o.arr("a")) <= 1e-13          # per-element planes in RCM element order
    for side in (1, 2, 3, 4):
        e.add_abso_side(side, False)
    e.add_force_at(0.37 * nx * 100.0, 0.61 * nz * 100.0, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.commit()
    rng = np.random.default_rng(2)
    n = e.npoin * 2
    d0, v0 = rng.standard_normal(n) * 1e-3, rng.standard_normal(n)
    e.set_fields(d0, v0)
    o.set_fields(d0, v0)
    nsteps = 60
    e.step(nsteps, np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)]))
    o.step(nsteps)
    d, v, _ = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10 and rel_l2(v, o.arr("v")) <= 1e-10
    e.close()
    o.close()


@pytest.mark.parametrize("ndof", [1, 2])
def test_snapshot_elem_fields(ndof):
    """s2d_cart_snapshot_elem on random fields, heterogeneous medium, natural element order, split-node row"""
    nx, nz, ez = 13, 9, 4
    o = orc.Oracle(harness.cart_deck(nx, nz, ndof=ndof, ezflt=ez, nrec=0, src=False, abso=(), fault=None), synthetic_seed=SEED,
                   renumber=False)
    e = CartEngine(5, ndof, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=ez, seed=SEED)
    e.commit()
    rng = np.random.default_rng(4)
    n = e.npoin * ndof
    d0, v0 = rng.standard_normal(n), rng.standard_normal(n)
    e.set_fields(d0, v0)
    o.set_fields(d0, v0)
    for what in ("E", "S") + (("d", "c") if ndof == 2 else ()):
        got, ref = e.snapshot_elem(what), o.snapshot(what)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max(), what
    e.close()
    o.close()


def test_accelerations_on_demand(monkeypatch):
    """The fused leapfrog step does not write the accelerations of the nodes it advances (8 B/DOF per s2d_step call
    saved): s2d_get_fields / s2d_cart_get_window form them from one force evaluation of d[n] (Engine::ensure_accel),
    also when the caller replaces the displacement first.  Same values as with S2D_ACCEL_LAZY=0 (stored by the last
    step of the call) and as the oracle's; one s2d_step call per step, as the e2e leg of bench.py does it."""
    nx, nz, nsteps = 43, 14, 40
    monkeypatch.setenv("S2D_SEG", "3")
    res = {}
    for lazy in ("1", "0"):
        monkeypatch.setenv("S2D_ACCEL_LAZY", lazy)
        o = orc.Oracle(harness.cart_deck(nx, nz, ezflt=6, nsteps=nsteps + 1), synthetic_seed=SEED, renumber=False)
        tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps + 1)])
        e = CartEngine(5, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), ezflt=6, seed=SEED)
        e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * 50.0, harness.nuc_radius(nx), nt_max=nsteps + 1)
        for side in (1, 2, 3, 4):
            e.add_abso_side(side)
        e.add_force_at(0.37 * nx * 100, 0.61 * nz * 100, [o.f("src.0.dir1"), o.f("src.0.dir2")])
        e.commit()
        n0 = e.launch_count()
        for k in range(nsteps):
            e.step(1, tab[k:k + 1])
        per_step = (e.launch_count() - n0) / nsteps
        o.step(nsteps)
        win = e.get_window(3, 2, 150, 40, a=True)[2]          # accelerations of a lattice window, formed on demand
        d, v, a = e.get_fields()
        for nm, got in (("d", d), ("v", v), ("acc", a)):
            assert rel_l2(got, o.arr(nm)) <= 1e-10, (lazy, nm, rel_l2(got, o.arr(nm)))
        # a new displacement does not lose the accelerations of the step before it
        e.step(1, tab[nsteps:])
        o.step(1)
        e.set_fields(np.zeros_like(d), None)
        assert rel_l2(e.get_fields()[2], o.arr("acc")) <= 1e-10
        res[lazy] = (a, win, per_step)
        e.close()
        o.close()
    scale = np.abs(res["0"][0]).max()
    assert np.abs(res["1"][0] - res["0"][0]).max() <= 1e-13 * scale
    assert np.abs(res["1"][1] - res["0"][1]).max() <= 1e-13 * scale
    assert res["1"][2] <= 6.1 and res["0"][2] <= 6.1   # 6 per step + the predictor of the very first one


@pytest.mark.parametrize("ngll,nx,nz,ezflt,seg", [(5, 27, 11, 0, 3), (5, 50, 12, 5, 2), (6, 23, 10, 3, 3), (4, 9, 7, 0, 32),
                                                  (3, 47, 8, 4, 5)])
def test_plastic_force_evaluations_and_plastic_strain(ngll, nx, nz, ezflt, seg, monkeypatch):
    """Coulomb plasticity in the strip kernel (s2d_cart_set_plastic): MAT_strain_PSV -> MAT_PLAST_stress(update) ->
    MAT_forces (mat_gen.f90:445-449,752-775,834-866; mat_plastic.f90:281-387).  Three successive force evaluations on
    growing random displacements -- each advances the plastic strain of every element GLL point -- against the
    oracle: forces after each, the plastic strain at the end, over strip / band decompositions with every halo case."""
    monkeypatch.setenv("S2D_SEG", str(seg))
    h = 100.0
    coh, phi, Tv, e0 = 2.0e6, 30.0, 0.02, (-4.0e-4, -3.0e-4, 2.5e-4)
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='plastic', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz}" + (f", ezflt={ezflt}" if ezflt else "") + " /",
         "&MATERIAL tag=1, kind='PLAST' /",
         f"&MAT_PLASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0, phi={phi}d0, coh={coh}d0, Tv={Tv}d0, e0={e0[0]}d0,{e0[1]}d0,{e0[2]}d0 /",
         "&TIME NbSteps=10, courant=0.5d0, kind='leapfrog' /"]
    o = orc.Oracle("\n".join(L) + "\n", renumber=False)
    assert o.i("npl") == nx * nz
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=0, rho=2670.0, cp=6000.0, cs=3464.0)
    assert e.npoin == o.i("npoin") and abs(e.dt - o.f("dt")) <= 1e-12 * e.dt
    e.set_dt(o.f("dt"))     # vp_factor = 1 - exp(-dt/Tv) from the same dt
    e.set_plastic([[coh, phi, Tv, *e0]], np.ones(nx * nz, np.int32))
    e.commit()
    _plastic_evaluations(e, o, ngll, nx, nz, np.arange(nx * nz))
    e.close()
    o.close()


def _plastic_evaluations(e, o, ngll, nx, nz, plastic_elems):
    rng = np.random.default_rng(ngll * 100 + nx)
    base = rng.standard_normal(e.npoin * 2)
    for amp in (1e-3, 3e-2, 1e-1):   # metres over 100 m elements: from barely yielding to deep in the plastic range
        d = amp * base
        e.set_fields(d, d)
        o.set_fields(d, d)
        ref = o.compute_fint()
        got = e.compute_fint()
        assert rel_l2(got, ref) <= 1e-12, (amp, rel_l2(got, ref))
    ep_ref = np.zeros((nx * nz, 3, ngll, ngll))
    ep_ref[plastic_elems] = o.arr("pl_ep").reshape(-1, 3, ngll, ngll)
    ep = e.plastic_strain()
    assert np.abs(ep_ref).max() > 1e-5
    assert np.abs(ep - ep_ref).max() <= 1e-12 * np.abs(ep_ref).max()


def test_plastic_and_elastic_elements_in_one_box(monkeypatch):
    """two tags: tag 1 plastic, tag 2 (fztag: the element rows next to the fault) elastic with other wave speeds.  The
    elastic elements take the same strain -> stress -> force kernel with material set 0 (never yields); the reference
    evaluates them with MAT_ELAST_f -- equal to rounding."""
    monkeypatch.setenv("S2D_SEG", "3")
    ngll, nx, nz, ezflt, h = 5, 31, 12, 6, 100.0
    coh, phi, Tv, e0 = 2.0e6, 25.0, 0.0, (-4.0e-4, -3.0e-4, 2.5e-4)   # Tv = 0: vp_factor = 1 (mat_plastic.f90:181-185)
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='plastic', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz}, ezflt={ezflt}, fztag=2, fznz=2 /",
         "&MATERIAL tag=1, kind='PLAST' /",
         f"&MAT_PLASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0, phi={phi}d0, coh={coh}d0, Tv={Tv}d0, e0={e0[0]}d0,{e0[1]}d0,{e0[2]}d0 /",
         "&MATERIAL tag=2, kind='ELAST' /",
         "&MAT_ELASTIC rho=2500.d0, cp=5000.d0, cs=2900.d0 /",
         "&TIME NbSteps=10, courant=0.5d0, kind='leapfrog' /"]
    o = orc.Oracle("\n".join(L) + "\n", renumber=False)
    tag = np.ones((nz, nx), np.int32)
    tag[ezflt - 2:ezflt + 2] = 2
    assert o.i("npl") == int((tag == 1).sum())
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=0, rho=2670.0, cp=6000.0, cs=3464.0)
    one = np.ones((nx * nz, ngll, ngll))
    t = tag.ravel()[:, None, None]
    e.set_material(np.where(t == 1, 2670.0, 2500.0) * one, np.where(t == 1, 6000.0, 5000.0) * one, np.where(t == 1, 3464.0, 2900.0) * one)
    e.set_dt(o.f("dt"))
    e.set_plastic([[coh, phi, Tv, *e0]], (tag.ravel() == 1).astype(np.int32))
    e.commit()
    _plastic_evaluations(e, o, ngll, nx, nz, np.flatnonzero(tag.ravel() == 1))
    e.close()
    o.close()


@pytest.mark.parametrize("ngll,nx,nz,ezflt,seg,nbody", [(5, 27, 11, 0, 3, 3), (6, 23, 10, 3, 3, 5), (4, 9, 7, 0, 32, 1), (3, 47, 8, 4, 5, 8)])
def test_visco_elastic_force_evaluations(ngll, nx, nz, ezflt, seg, nbody, monkeypatch):
    """Visco-elasticity in the strip kernel (s2d_cart_set_visco): MAT_strain_PSV -> MAT_VISCO_stress -> MAT_forces
    (mat_gen.f90:451-457, mat_visco.f90:206-248).  Every force evaluation relaxes the Nbody memory variables of each
    element GLL point towards the strain of the previous evaluation and keeps the new strain, so four successive
    evaluations on different displacements test the whole state; theta, wbody and the unrelaxed moduli come from the
    oracle's get_attenuation (mat_visco.f90:251-340), as a Fortran host would hand them over."""
    monkeypatch.setenv("S2D_SEG", str(seg))
    h = 100.0
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='visco', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz}" + (f", ezflt={ezflt}" if ezflt else "") + " /",
         "&MATERIAL tag=1, kind='VISCO' /",
         f"&MAT_VISCO rho=2000d0, cp=3000d0, cs=2000d0, QP=30d0, QS=20d0, Nbody={nbody}, fmin=1.8d0, fmax=180d0 /",
         "&TIME NbSteps=10, courant=0.5d0, kind='leapfrog' /"]
    o = orc.Oracle("\n".join(L) + "\n", renumber=False)
    assert o.i("nvs") == nx * nz
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=0, rho=2000.0, cp=3000.0, cs=2000.0)
    e.set_dt(o.f("dt"))
    e.set_visco([nbody], o.arr("mat.1.moduli"), [o.arr("mat.1.wbody")], [o.arr("mat.1.theta").reshape(3, nbody).T],
                np.ones(nx * nz, np.int32))
    e.commit()
    rng = np.random.default_rng(ngll * 100 + nx)
    for k in range(4):
        d = rng.standard_normal(e.npoin * 2) * 1e-3
        e.set_fields(d, d)
        o.set_fields(d, d)
        ref = o.compute_fint()
        got = e.compute_fint()
        assert rel_l2(got, ref) <= 1e-12, (k, rel_l2(got, ref))
    assert np.abs(o.arr("vs_el")).max() > 0
    e.close()
    o.close()


@pytest.mark.parametrize("ngll,nx,nz,ezflt,seg,beta,alpha0", [(5, 27, 11, 0, 3, 0.0, 0.0), (6, 23, 10, 3, 3, 0.0, 0.1),
                                                            (4, 9, 7, 0, 32, 0.5, 0.05), (3, 47, 8, 4, 5, 0.0, 0.0)])
def test_damage_rheology_force_evaluations_and_state(ngll, nx, nz, ezflt, seg, beta, alpha0, monkeypatch):
    """Damage rheology in the strip kernel (s2d_cart_set_damage): MAT_strain_PSV -> MAT_DMG_stress(update, dt) ->
    MAT_forces (mat_gen.f90:451-457, mat_damage.f90:337-491).  Successive force evaluations on growing random
    displacements over the deck's prestrain: forces after each, damage variable and plastic strain at the end."""
    monkeypatch.setenv("S2D_SEG", str(seg))
    h = 100.0
    e0, ep0 = (-1.487381e-3, -1.708729e-4, 3.5e-4), (1e-5, -2e-5, 3e-5) if alpha0 else (0.0, 0.0, 0.0)
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='damage', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz}" + (f", ezflt={ezflt}" if ezflt else "") + " /",
         "&MATERIAL tag=1, kind='DMG' /",
         f"&MAT_DAMAGE rho=2670.d0, cp=6000.d0, cs=3464.d0, beta={beta}d0, R=1.d0, Cd=2.5d5, phi=30.9638d0, alpha={alpha0}d0,",
         f"   e0={e0[0]}d0,{e0[1]}d0,{e0[2]}d0, ep={ep0[0]}d0,{ep0[1]}d0,{ep0[2]}d0 /",
         "&TIME NbSteps=10, courant=0.5d0, kind='leapfrog' /"]
    o = orc.Oracle("\n".join(L) + "\n", renumber=False)
    assert o.i("ndm") == nx * nz
    rho, cp, cs = 2670.0, 6000.0, 3464.0
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=0, rho=rho, cp=cp, cs=cs)
    e.set_dt(o.f("dt"))
    e.set_damage([[rho * (cp * cp - 2 * cs * cs), rho * cs * cs, 30.9638, alpha0, 2.5e5, beta, 1.0, *e0, *ep0]],
                 np.ones(nx * nz, np.int32))
    e.commit()
    rng = np.random.default_rng(ngll * 100 + nx)
    base = rng.standard_normal(e.npoin * 2)
    for amp in (1e-3, 1e-2, 3e-2):
        d = amp * base
        e.set_fields(d, d)
        o.set_fields(d, d)
        ref = o.compute_fint()
        got = e.compute_fint()
        assert rel_l2(got, ref) <= 1e-12, (amp, rel_l2(got, ref))
    st_ref = o.arr("dm_state").reshape(nx * nz, 4, ngll, ngll)
    st = e.damage_state()
    assert st_ref[:, 0].max() > alpha0 + 1e-4          # damage has grown
    assert np.abs(st - st_ref).max() <= 1e-12 * np.abs(st_ref).max()
    e.close()
    o.close()


@pytest.mark.parametrize("rheology", ["plastic", "visco", "damage"])
def test_stateful_rheologies_fused_steps_fp64_and_fp32(rheology):
    """40 fused leapfrog steps (absorbing sides, a Ricker force strong enough to drive the medium into the inelastic
    range) with each stateful rheology, FP64 at 1e-10 and FP32 fields at 1e-4 against the oracle: the state lives
    through the fused update, the deferred nodes and the double displacement buffer."""
    nx, nz, nsteps, h, ngll = 29, 13, 40, 100.0, 5
    mat = {"plastic": ("PLAST", "&MAT_PLASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0, phi=30d0, coh=1.0d4, Tv=0.01d0, e0=-4d-6,-3d-6,2.5d-6 /"),
           "visco": ("VISCO", "&MAT_VISCO rho=2670.d0, cp=6000.d0, cs=3464.d0, QP=40d0, QS=25d0, Nbody=3, fmin=0.5d0, fmax=50d0 /"),
           "damage": ("DMG", "&MAT_DAMAGE rho=2670.d0, cp=6000.d0, cs=3464.d0, beta=0d0, R=1d0, Cd=1d5, phi=30.9638d0, alpha=0d0, e0=-1.487381d-5,-1.708729d-6,3.5d-6 /")}[rheology]
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='rheology', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz} /",
         f"&MATERIAL tag=1, kind='{mat[0]}' /", mat[1]]
    for t in (1, 2, 3, 4):
        L += [f"&BC_DEF tag={t}, kind='ABSORB' /"]
    L += [f"&TIME NbSteps={nsteps}, courant=0.5d0, kind='leapfrog' /",
          f"&SRC_DEF stf='RICKER', coord={0.37*nx*h}d0,{0.61*nz*h}d0, mechanism='FORCE' /",
          "&STF_RICKER f0=4.d0, onset=0.02d0, ampli=2.d10 /", "&SRC_FORCE angle=30d0 /"]
    deck = "\n".join(L) + "\n"
    rho, cp, cs = 2670.0, 6000.0, 3464.0
    for precision, tol in ((8, 1e-10), (4, 1e-4)):
        o = orc.Oracle(deck, renumber=False)
        e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), seed=0, rho=rho, cp=cp, cs=cs, precision=precision)
        e.set_dt(o.f("dt"))
        one = np.ones(nx * nz, np.int32)
        if rheology == "plastic":
            e.set_plastic([[1.0e4, 30.0, 0.01, -4e-6, -3e-6, 2.5e-6]], one)
        elif rheology == "visco":
            e.set_visco([3], o.arr("mat.1.moduli"), [o.arr("mat.1.wbody")], [o.arr("mat.1.theta").reshape(3, 3).T], one)
        else:
            e.set_damage([[rho * (cp * cp - 2 * cs * cs), rho * cs * cs, 30.9638, 0.0, 1e5, 0.0, 1.0,
                           -1.487381e-5, -1.708729e-6, 3.5e-6, 0.0, 0.0, 0.0]], one)
        for side in (1, 2, 3, 4):
            e.add_abso_side(side)
        e.add_force_at(0.37 * nx * h, 0.61 * nz * h, [o.f("src.0.dir1"), o.f("src.0.dir2")])
        e.commit()
        tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
        e.step(25, tab[:25])
        e.step(nsteps - 25, tab[25:])
        o.step(nsteps)
        d, v, a = e.get_fields()
        assert np.abs(o.arr("d")).max() > 0
        for nm, got in (("d", d), ("v", v), ("acc", a)):
            assert rel_l2(got, o.arr(nm)) <= tol, (rheology, precision, nm, rel_l2(got, o.arr(nm)))
        if rheology == "plastic":
            assert np.abs(o.arr("pl_ep")).max() > 0
        if rheology == "damage":
            assert o.arr("dm_state").reshape(-1, 4, 25)[:, 0].max() > 0
        e.close()
        o.close()


def test_damage_beyond_the_critical_value_aborts_like_the_reference():
    """compute_stress (mat_damage.f90:478-489) stops the run when the damaged moduli lose convexity; with a damage
    evolution coefficient this large the oracle aborts with 'MAT_DMG: damage exceeded critical value' -- and so does
    the engine (device error flag raised by the strip kernel, reported by s2d_step)."""
    from sem2dpack_b200.capi import S2DError
    nx, nz, nsteps, h, ngll = 29, 13, 40, 100.0, 5
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='rheology', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz} /",
         "&MATERIAL tag=1, kind='DMG' /",
         "&MAT_DAMAGE rho=2670.d0, cp=6000.d0, cs=3464.d0, beta=0d0, R=1d0, Cd=1d9, phi=30.9638d0, alpha=0d0, e0=-1.487381d-5,-1.708729d-6,3.5d-6 /",
         f"&TIME NbSteps={nsteps}, courant=0.5d0, kind='leapfrog' /",
         f"&SRC_DEF stf='RICKER', coord={0.37*nx*h}d0,{0.61*nz*h}d0, mechanism='FORCE' /",
         "&STF_RICKER f0=4.d0, onset=0.02d0, ampli=2.d10 /", "&SRC_FORCE angle=30d0 /"]
    o = orc.Oracle("\n".join(L) + "\n", renumber=False)
    with pytest.raises(Exception, match="damage exceeded critical value"):
        o.step(nsteps)
    rho, cp, cs = 2670.0, 6000.0, 3464.0
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), seed=0, rho=rho, cp=cp, cs=cs)
    e.set_dt(o.f("dt"))
    e.set_damage([[rho * (cp * cp - 2 * cs * cs), rho * cs * cs, 30.9638, 0.0, 1e9, 0.0, 1.0, -1.487381e-5, -1.708729e-6, 3.5e-6,
                   0.0, 0.0, 0.0]], np.ones(nx * nz, np.int32))
    e.add_force_at(0.37 * nx * h, 0.61 * nz * h, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.commit()
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    with pytest.raises(S2DError, match="damage exceeded critical value"):
        e.step(nsteps, tab)
    e.close()
    o.close()


@pytest.mark.parametrize("ngll,nx,nz", [(6, 23, 12), (9, 14, 7), (3, 47, 9), (8, 9, 6)])
def test_fp32_fused_steps_other_orders(ngll, nx, nz):
    """FP32 fields through the tensor-map instantiations of the other GLL orders (box widths and strides in 4-byte
    elements): 60 fused leapfrog steps with absorbing sides and a force source, 1e-4 against the FP64 oracle."""
    nsteps = 60
    o = orc.Oracle(harness.cart_deck(nx, nz, ngll=ngll, nsteps=nsteps, abso=(1, 2, 3, 4)), synthetic_seed=SEED, renumber=False)
    e = CartEngine(ngll, 2, nx, nz, (0.0, nx * 100.0), (0.0, nz * 100.0), seed=SEED, precision=4)
    for side in (1, 2, 3, 4):
        e.add_abso_side(side)
    e.add_force_at(0.37 * nx * 100, 0.61 * nz * 100, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.commit()
    e.step(nsteps, np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)]))
    o.step(nsteps)
    d, v, a = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-4 and rel_l2(v, o.arr("v")) <= 1e-4 and rel_l2(a, o.arr("acc")) <= 1e-3
    e.close()
    o.close()
