"""x-strip decomposition on ONE GPU: several strip engines of one box, exchanging interface partial
sums through the engine's halo path (the one the multi-GPU run uses), against the whole-box engine
and the oracle."""
import numpy as np
import pytest

import harness
import orc
from harness import rel_l2
from sem2dpack_b200 import CartEngine, strips

pytestmark = pytest.mark.gpu
SEED = 20261017
H = 100.0


def _build(nx_glob, nz, lo, hi, ezflt, scheme_kind, nsteps, world_rank, world, src_dir, ngll=5, ndof=2, stacey=False,
           dt=None):
    e = CartEngine(ngll, ndof, hi - lo, nz, (lo * H, hi * H), (0.0, nz * H), ezflt=ezflt, seed=SEED,
                   scheme_kind=scheme_kind, courant=0.5, ix0=lo * (ngll - 1), halo_left=world_rank > 0,
                   halo_right=world_rank < world - 1)
    if dt is not None:
        assert e.dt >= dt * (1 - 1e-12)  # the global Courant step is the minimum over the strips
        e.set_dt(dt)
    fid = None
    if ezflt:
        fid = e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx_glob * H / 2, harness.nuc_radius(nx_glob, H),
                              nt_max=nsteps)
    sides = [1, 3] + ([4] if world_rank == 0 else []) + ([2] if world_rank == world - 1 else [])
    for s in sorted(sides):
        e.add_abso_side(s, stacey)
    xs, zs = 0.37 * nx_glob * H, 0.61 * nz * H
    nsrc = 0
    if lo * H <= xs < hi * H:
        e.add_force_at(xs, zs, src_dir)
        nsrc = 1
    e.commit()
    return e, fid, nsrc


@pytest.mark.parametrize("world,nx,nz,ezflt,seg,scheme,peer", [
    (2, 20, 12, 5, 4, "leapfrog", False), (3, 30, 9, 4, 32, "newmark", False), (4, 16, 8, 0, 3, "leapfrog", False),
    # the device-side peer-memory protocol (slots + flags) of the multi-GPU run, here between two strips
    # that share one GPU; few streams on purpose: a spinning wait kernel must not share a hardware queue
    # with the kernel it waits for
    (2, 20, 12, 5, 4, "leapfrog", True), (2, 26, 10, 4, 32, "newmark", True)])
def test_strips_match_whole_box(world, nx, nz, ezflt, seg, scheme, peer, monkeypatch):
    monkeypatch.setenv("S2D_SEG", str(seg))
    nsteps = 200
    kind = 0 if scheme == "leapfrog" else 1
    o = orc.Oracle(harness.cart_deck(nx, nz, ezflt=ezflt, scheme=scheme, nsteps=nsteps, nrec=0), synthetic_seed=SEED,
                   renumber=False)
    sdir = [o.f("src.0.dir1"), o.f("src.0.dir2")]
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    whole, _, _ = _build(nx, nz, 0, nx, ezflt, kind, nsteps, 0, 1, sdir)
    whole.step(nsteps, tab)
    dw, vw, aw = whole.get_fields()
    ibw = whole.get_tables(rmass=False)[0].reshape(nz, nx, 25)
    o.step(nsteps)
    assert rel_l2(dw, o.arr("d")) <= 1e-10

    parts = strips.partition(nx, world)
    built = [_build(nx, nz, lo, hi, ezflt, kind, nsteps, r, world, sdir, dt=whole.dt) for r, (lo, hi) in enumerate(parts)]
    engines = [b[0] for b in built]
    grp = strips.LocalStrips(engines, peer=peer)
    grp.run(lambda r, e: e.step(nsteps, tab if built[r][2] else None))
    npw = whole.npoin
    fields = []
    for r, (lo, hi) in enumerate(parts):
        e = engines[r]
        d, v, acc = e.get_fields()   # accelerations: formed on demand, one strip at a time (no exchange needed)
        fields.append((d, v))
        ib = e.get_tables(rmass=False)[0].reshape(nz, hi - lo, 25) - 1
        gl = ibw[:, lo:hi, :] - 1
        for c in range(2):
            assert rel_l2(d[ib + c * e.npoin], dw[gl + c * npw]) <= 1e-11
            assert rel_l2(v[ib + c * e.npoin], vw[gl + c * npw]) <= 1e-11
            assert rel_l2(acc[ib + c * e.npoin], aw[gl + c * npw]) <= 1e-11
    # the two copies of every interface node are bit-identical
    for r in range(world - 1):
        eL, eR = engines[r], engines[r + 1]
        ibL = eL.get_tables(rmass=False)[0].reshape(nz, -1, 5, 5)[:, -1, :, 4] - 1   # i = N-1 column of the last element
        ibR = eR.get_tables(rmass=False)[0].reshape(nz, -1, 5, 5)[:, 0, :, 0] - 1    # i = 0 column of the first element
        for c in range(2):
            for q in (0, 1):
                a = fields[r][q][ibL + c * eL.npoin]
                b = fields[r + 1][q][ibR + c * eR.npoin]
                assert np.array_equal(a, b)
    if ezflt:
        stw = whole.fault_state(0, nx * 4 + 1)
        for r, (lo, hi) in enumerate(parts):
            st = engines[r].fault_state(built[r][1], (hi - lo) * 4 + 1)
            # slip (zero to rounding where the fault is still locked) and shear traction
            n1 = (hi - lo) * 4 + 1
            scale = max(np.abs(stw["D"]).max(), 1e-3)
            assert np.abs(st["D"][:n1] - stw["D"][lo * 4:hi * 4 + 1]).max() <= 1e-10 * scale
            assert rel_l2(st["T"][:n1], stw["T"][lo * 4:hi * 4 + 1]) <= 1e-10
    for e in engines:
        e.close()
    whole.close()
    o.close()


@pytest.mark.parametrize("rheology", ["plastic", "damage"])
def test_strips_with_a_stateful_rheology_match_the_whole_box(rheology, monkeypatch):
    """the per-element state of the stateful rheologies partitions with the elements: three x-strips of a plastic /
    damage box (heterogeneous hash medium, absorbing sides, a force source) against the same box in one piece"""
    monkeypatch.setenv("S2D_SEG", "4")
    world, nx, nz, nsteps = 3, 27, 10, 60
    sdir = [0.5, np.sqrt(0.75)]
    par_pl = [[2.0e5, 30.0, 0.01, -4e-6, -3e-6, 2.5e-6]]
    lam_mu = (2670.0 * (6000.0 ** 2 - 2 * 3464.0 ** 2), 2670.0 * 3464.0 ** 2)
    par_dm = [[lam_mu[0], lam_mu[1], 30.9638, 0.0, 1e6, 0.0, 1.0, -1.487381e-5, -1.708729e-6, 3.5e-6, 0.0, 0.0, 0.0]]

    def build(lo, hi, rank, nworld, dt=None):
        e = CartEngine(5, 2, hi - lo, nz, (lo * H, hi * H), (0.0, nz * H), seed=0, rho=2670.0, cp=6000.0, cs=3464.0,
                       scheme_kind=0, courant=0.5, ix0=lo * 4, halo_left=rank > 0, halo_right=rank < nworld - 1)
        if dt is not None:
            e.set_dt(dt)
        one = np.ones((hi - lo) * nz, np.int32)
        if rheology == "plastic":
            e.set_plastic(par_pl, one)
        else:
            e.set_damage(par_dm, one)
        for s in sorted([1, 3] + ([4] if rank == 0 else []) + ([2] if rank == nworld - 1 else [])):
            e.add_abso_side(s, False)
        xs, zs = 0.37 * nx * H, 0.61 * nz * H
        has = lo * H <= xs < hi * H
        if has:
            e.add_force_at(xs, zs, sdir)
        e.commit()
        return e, has

    whole, _ = build(0, nx, 0, 1)
    t = (np.arange(nsteps) + 1) * whole.dt
    arg = (np.pi * 4.0 * (t - 0.05)) ** 2
    amp = 2e10 if rheology == "plastic" else 3e9                # strong enough to leave the elastic range, below critical damage
    tab = (amp * (1 - 2 * arg) * np.exp(-arg))[:, None]
    whole.step(nsteps, tab)
    dw, vw, aw = whole.get_fields()
    ibw = whole.get_tables(rmass=False)[0].reshape(nz, nx, 25)
    state_w = (whole.plastic_strain() if rheology == "plastic" else whole.damage_state()).reshape(nz, nx, -1)
    assert np.abs(state_w).max() > 0
    parts = strips.partition(nx, world)
    built = [build(lo, hi, r, world, dt=whole.dt) for r, (lo, hi) in enumerate(parts)]
    engines = [b[0] for b in built]
    grp = strips.LocalStrips(engines, peer=False)
    grp.run(lambda r, e: e.step(nsteps, tab if built[r][1] else None))
    for r, (lo, hi) in enumerate(parts):
        e = engines[r]
        d, v, a = e.get_fields()
        ib = e.get_tables(rmass=False)[0].reshape(nz, hi - lo, 25) - 1
        gl = ibw[:, lo:hi, :] - 1
        for c in range(2):
            assert rel_l2(d[ib + c * e.npoin], dw[gl + c * whole.npoin]) <= 1e-11
            assert rel_l2(v[ib + c * e.npoin], vw[gl + c * whole.npoin]) <= 1e-11
        st = (e.plastic_strain() if rheology == "plastic" else e.damage_state()).reshape(nz, hi - lo, -1)
        assert np.abs(st - state_w[:, lo:hi]).max() <= 1e-11 * np.abs(state_w).max()
    for e in engines:
        e.close()
    whole.close()
