"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-10 in FP64, <= 1e-5 in FP32, on FP64 state
(fields, fault state); float32 outputs (seismograms, fault records) are compared after the same
float32 cast on both sides, so their floor is one float32 ulp of the trace maximum.
"""
import numpy as np
import pytest

import harness
import orc
from harness import Rig, rel_l2
from sem2dpack_b200.capi import S2D_ASM_ATOMIC, S2D_ASM_COLOR, S2D_ASM_PATCH

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-5
VARIANTS = {"patch": S2D_ASM_PATCH, "color": S2D_ASM_COLOR, "atomic": S2D_ASM_ATOMIC}


def _rand_fields(o, seed=1):
    rng = np.random.default_rng(seed)
    n = o.i("npoin") * o.i("ndof")
    return rng.standard_normal(n), rng.standard_normal(n)


@pytest.fixture(params=["strip-routing", "any-mesh"])
def route(request, monkeypatch):
    """structured decks handed over through the generic API run on the strip kernel by default
    (s2d_kernel_route); S2D_ROUTE_STRIP=0 keeps them on the any-mesh kernels, which stay covered this way"""
    monkeypatch.setenv("S2D_ROUTE_STRIP", "1" if request.param == "strip-routing" else "0")
    return request.param


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", ["testsh", "lamb", "tpv3", "ratestate"])
def test_fint_matches_oracle(name, variant, route):
    """compute_Fint (solver.f90:273-320) on random fields: every kernel variant, NGLL 5/6/9, SH and
    P-SV, shared and KV-carrying elements."""
    o = orc.Oracle(harness.deck(name))
    r = Rig(o, variant=VARIANTS[variant])
    d, v = _rand_fields(o)
    o.set_fields(d, v)
    r.e.set_fields(d, v)
    ref = o.compute_fint()
    got = r.e.compute_fint()
    assert rel_l2(got, ref) <= 1e-13
    r.close()


@pytest.mark.parametrize("ngll", [3, 4, 5, 6, 7, 8, 9, 10])
@pytest.mark.parametrize("ndof", [1, 2])
def test_fint_all_ngll_hetero(ngll, ndof, route):
    """heterogeneous medium (one coefficient block per element -> patch-major plane layout)."""
    o = orc.Oracle(harness.cart_deck(12, 9, ngll=ngll, ndof=ndof, nrec=0, src=False), synthetic_seed=20261017)
    assert o.i("ncoefsets") == o.i("nelem")
    d, v = _rand_fields(o, seed=ngll)
    o.set_fields(d, v)
    ref = o.compute_fint()
    for variant in VARIANTS.values():
        r = Rig(orc.Oracle(harness.cart_deck(12, 9, ngll=ngll, ndof=ndof, nrec=0, src=False),
                           synthetic_seed=20261017), variant=variant)
        r.e.set_fields(d, v)
        assert rel_l2(r.e.compute_fint(), ref) <= 1e-13
        r.close()
    o.close()


def test_fint_kd1_vs_kd2_nonsquare():
    """ELAST_KD2_PSV uses a4*(dUx_deta+dUz_dxi) where KD1 uses a4*dUx_deta+a5*dUz_dxi
    (mat_elastic.f90:612 vs :489): they differ on non-square elements and the engine must follow
    whichever the reference would take."""
    txt = harness.cart_deck(10, 10, ngll=5, ndof=2, nrec=0, src=False).replace("zlim=0d0,1000.0d0", "zlim=0d0,700.0d0")
    for force_kd1 in (False, True):
        o = orc.Oracle(txt, kd_force_kd1=force_kd1)
        r = Rig(o, kd2=not force_kd1)
        d, v = _rand_fields(o)
        o.set_fields(d, v)
        r.e.set_fields(d, v)
        assert rel_l2(r.e.compute_fint(), o.compute_fint()) <= 1e-13
        r.close()


def test_coloring_is_greedy_first_fit():
    o = orc.Oracle(harness.cart_deck(9, 7, ngll=5, ndof=2, ezflt=3, nrec=0, src=False))
    r = Rig(o, variant=S2D_ASM_COLOR)
    nc, col = r.e.coloring()
    ib = o.arr("ibool").reshape(o.i("nelem"), -1)
    used = {}
    exp = np.zeros(len(ib), np.int32)
    for e, nodes in enumerate(ib):
        taken = set()
        for n in nodes:
            taken |= used.get(int(n), set())
        c = 0
        while c in taken:
            c += 1
        exp[e] = c
        for n in nodes:
            used.setdefault(int(n), set()).add(c)
    assert np.array_equal(col, exp)
    assert nc == exp.max() + 1
    # validity: no two elements of one colour share a node
    for c in range(nc):
        nodes = ib[col == c].ravel()
        assert len(nodes) == len(set(nodes.tolist()))
    r.close()


def _lockstep(o, r, nsteps, chunk, tol, check_fault=True, skip_tstick=False):
    done = 0
    while done < nsteps:
        n = min(chunk, nsteps - done)
        o.step(n)
        r.step(n)
        done += n
    d, v, a = r.e.get_fields()
    for name, got in (("d", d), ("v", v), ("acc", a)):
        ref = o.arr(name)
        assert rel_l2(got, ref) <= tol, (name, rel_l2(got, ref))
    if o.i("rec.present"):
        s_ref, s_got = o.seis(), r.e.seis()
        scale = np.abs(s_ref).max()
        assert np.abs(s_got - s_ref).max() <= max(tol, 2e-7) * scale
    if check_fault:
        for fid, ibc, np_, onx in r.faults:
            st = r.e.fault_state(fid, np_)
            p = f"bc.{ibc}."
            for k, key in (("D", "D"), ("V", "V"), ("T", "T"), ("MU", "MU"), ("sigma", "sigma")):
                assert rel_l2(st[k], o.arr(p + key)) <= tol, (k, rel_l2(st[k], o.arr(p + key)))
            rec, pot = r.e.fault(fid, onx)
            rec_ref = o.arr(p + "out").reshape(-1, 6, onx)
            assert rec.shape == rec_ref.shape
            ncol = 5 if skip_tstick else 6
            for c in range(ncol):
                scale = max(np.abs(rec_ref[:, c]).max(), 1e-30)
                assert np.abs(rec[:, c] - rec_ref[:, c]).max() <= max(tol, 2e-7) * scale, c
            pot_ref = o.arr(p + "potency").reshape(-1, pot.shape[1])
            assert pot.shape == pot_ref.shape
            assert np.abs(pot - pot_ref).max() <= max(tol, 1e-12) * max(np.abs(pot_ref).max(), 1e-300)


def test_testsh_full_run_and_known_answer(route):
    """EXAMPLES/TestSH end to end (1987 leapfrog steps, NGLL=6 SH, ABSORB, Ricker force, receivers 'D'),
    in lock step with the oracle AND against the analytic trace of analyze_test.m (< 2 %)."""
    o = orc.Oracle(harness.deck("testsh"))
    r = Rig(o)
    _lockstep(o, r, o.i("nt"), 500, TOL64)
    u = r.e.seis()[:, 4, 0].astype(np.float64)
    uref = harness.refdata()["testsh_uref"]
    assert np.abs(u - uref).max() / np.abs(uref).max() < 0.02
    r.close()


def test_lamb_full_run_and_known_answer(route):
    """EXAMPLES/LambsProblem (3000 steps, NGLL=9 P-SV, P1 absorbing, free surface): misfits of
    analyze_test.m must reproduce the values recorded in test.out:12."""
    o = orc.Oracle(harness.deck("lamb"))
    r = Rig(o)
    _lockstep(o, r, o.i("nt"), 1000, TOL64)
    s = r.e.seis().astype(np.float64)  # (nt, 2, 2)
    g = harness.refdata()
    uxa = np.vstack([np.zeros((1, 2)), g["lamb_ux"].reshape(2, -1).T])
    uza = np.vstack([np.zeros((1, 2)), g["lamb_uz"].reshape(2, -1).T])
    num = np.abs(np.hstack([s[:, :, 0] - uxa, s[:, :, 1] - uza])).max(axis=0)
    den = np.abs(np.hstack([uxa, uza])).max(axis=0)
    err = num / den
    assert err.max() < 0.005
    assert np.allclose(err, g["lamb_misfits"], rtol=2e-3)
    r.close()


def test_tpv3_slip_weakening_run():
    """EXAMPLES/TestFlt2D_SCEC_TPV3_inplane: Newmark, KV elements, one-sided SWF fault, ABSORB, DIRNEU,
    interpolated receivers; 1200 steps (rupture well under way)."""
    o = orc.Oracle(harness.deck("tpv3"))
    r = Rig(o)
    _lockstep(o, r, 1200, 400, TOL64)
    st = r.e.fault_state(r.faults[0][0], r.faults[0][2])
    assert np.abs(st["D"]).max() > 0.5  # the fault has slipped
    r.close()


def test_ratestate_run(route):
    """EXAMPLES/RateState: rate-and-state slip law (kind 3, Newton/bisection per node), 803 steps."""
    o = orc.Oracle(harness.deck("ratestate"))
    r = Rig(o)
    _lockstep(o, r, o.i("nt"), 300, 1e-9, skip_tstick=True)
    r.close()


@pytest.mark.parametrize("variant", ["color", "atomic"])
def test_alternative_assembly_variants_run(variant):
    o = orc.Oracle(harness.deck("tpv3"))
    r = Rig(o, variant=VARIANTS[variant])
    _lockstep(o, r, 300, 300, TOL64)
    r.close()


@pytest.mark.parametrize("scheme", ["leapfrog", "newmark"])
@pytest.mark.parametrize("stacey", [False, True])
def test_synthetic_two_sided_fault(scheme, stacey):
    """the benchmark family at test size: heterogeneous P-SV, two-sided SWF fault on tags 5,6,
    absorbing boundaries on all sides (P1 or Stacey), force source, nearest-node receivers."""
    txt = harness.cart_deck(24, 16, ezflt=8, scheme=scheme, stacey=stacey, nsteps=400)
    o = orc.Oracle(txt, synthetic_seed=20261017)
    r = Rig(o)
    _lockstep(o, r, 400, 200, TOL64)
    r.close()


def test_fp32_mode():
    o = orc.Oracle(harness.deck("testsh"))
    r = Rig(o, precision=4)
    n = 600
    o.step(n)
    r.step(n)
    d, v, _ = r.e.get_fields()
    assert rel_l2(d, o.arr("d")) <= TOL32
    assert rel_l2(v, o.arr("v")) <= TOL32
    r.close()


def test_progress_and_energy():
    o = orc.Oracle(harness.deck("lamb"))
    r = Rig(o)
    o.step(300)
    r.step(300)
    vmax, dmax = r.e.progress()
    assert abs(vmax - np.abs(o.arr("v")).max()) <= 1e-12 * vmax
    assert abs(dmax - np.abs(o.arr("d")).max()) <= 1e-12 * dmax
    r.close()


def _engine_from(o, a, nelast, kd2, variant=0, abso_flat=True):
    """an Engine fed like harness.Rig, but with the caller's coefficient planes / absorbing-boundary form"""
    from sem2dpack_b200 import Engine
    e = Engine(o.i("ngll"), o.i("ndof"), o.arr("ibool"), o.arr("H"), o.arr("rmass"), o.i("scheme"), o.f("dt"),
               o.f("beta"), o.f("gamma"), o.f("alpha"))
    e.set_elastic(nelast, a, o.arr("elem2set"), kd2)
    for i in range(o.i("nbc")):
        p = f"bc.{i}."
        if o.i(p + "kind") == harness.IS_ABSORB:
            e.add_abso(o.arr(p + "node"), o.arr(p + "C"), is_flat=abso_flat, n=o.arr(p + "n"))
    for s in range(o.i("nsrc")):
        e.add_force(o.i(f"src.{s}.iglob"), [o.f(f"src.{s}.dir1"), o.f(f"src.{s}.dir2")])
    e.commit(variant)
    return e


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("ndof", [1, 2])
def test_general_planes_reduce_to_the_flat_ones(ndof, variant):
    """nelast = 3 / 10 (curved meshes, mat_elastic.f90:344-358,497-515,663-676): with the planes of a flat grid
    embedded (the cross terms DxiDz, DetaDx vanish) the general kernels must give the flat KD1 result"""
    deck = harness.cart_deck(11, 9, ngll=6, ndof=ndof, nrec=0, src=False).replace("zlim=0d0,900.0d0", "zlim=0d0,700.0d0")
    o = orc.Oracle(deck, synthetic_seed=20261017)
    n2, nset = 36, o.i("ncoefsets")
    af = o.arr("a").reshape(nset, -1, n2)
    ag = np.zeros((nset, 3 if ndof == 1 else 10, n2))
    ag[:, :af.shape[1]] = af                       # a1..a2 (SH) / a1..a6 (P-SV); a3 / a7..a10 = 0
    d, v = _rand_fields(o, seed=7)
    o.set_fields(d, v)
    ref = o.compute_fint()
    e = _engine_from(o, ag, ag.shape[1], False, VARIANTS[variant])
    e.set_fields(d, v)
    assert rel_l2(e.compute_fint(), ref) <= 1e-13
    e.close()
    o.close()


@pytest.mark.parametrize("ndof", [1, 2])
def test_general_planes_give_a_symmetric_operator(ndof):
    """with arbitrary cross planes (a3 / a7..a10 != 0) no oracle exists (the oracle meshes are flat), but the
    operator of MAT_ELAST_f stays symmetric: <w, K u> = <u, K w>"""
    o = orc.Oracle(harness.cart_deck(9, 7, ngll=5, ndof=ndof, nrec=0, src=False), synthetic_seed=20261017)
    n2, nset = 25, o.i("ncoefsets")
    rng = np.random.default_rng(3)
    ag = rng.standard_normal((nset, 3 if ndof == 1 else 10, n2))
    e = _engine_from(o, ag, ag.shape[1], False)
    n = o.i("npoin") * ndof
    u, w = rng.standard_normal(n), rng.standard_normal(n)
    e.set_fields(u, np.zeros(n))
    Ku = e.compute_fint()
    e.set_fields(w, np.zeros(n))
    Kw = e.compute_fint()
    a, b = float(w @ Ku), float(u @ Kw)
    assert abs(a - b) <= 1e-12 * max(abs(a), abs(b)), (a, b)
    e.close()
    o.close()


def test_nonflat_absorbing_form_on_axis_aligned_sides():
    """BC_ABSO_apply's normal / tangential form (bc_abso.f90:311-317) on boundaries whose normals are the axes
    must reproduce the flat form: Lamb's deck stepped with is_flat = 0"""
    o = orc.Oracle(harness.deck("lamb"))
    e = _engine_from(o, o.arr("a"), o.i("nelast"), False, abso_flat=False)
    nsteps = 400
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    e.step(nsteps, tab)
    o.step(nsteps)
    d, v, _ = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10 and rel_l2(v, o.arr("v")) <= 1e-10
    assert np.abs(d).max() > 0
    e.close()
    o.close()


def _fault_run(deck, nsteps, tol=1e-10, skip_tstick=False):
    o = orc.Oracle(deck)
    r = Rig(o)
    _lockstep(o, r, nsteps, 97, tol, skip_tstick=skip_tstick)
    slip = np.abs(o.arr("bc.0.D")).max()
    r.close()
    return slip


@pytest.mark.parametrize("edit", [
    ("Dc=0.4d0,", "kind=2, Dc=0.4d0,"),                                  # exponential weakening (bc_dynflt_swf.f90:160-163)
    ("Dc=0.4d0,", "kind=3, p=2.5d0, Dc=0.4d0,"),                         # power law (:164-167)
    ("Dc=0.4d0,", "healing=T, Dc=0.4d0,"),                               # state heals with the slip rate (:178-190)
    ("Dc=0.4d0,", "alpha=0.02d0, Dc=0.4d0,"),                            # + alpha*theta
])
def test_slip_weakening_variants(edit):
    """the TPV3 deck (one-sided fault, Kelvin-Voigt layer, Newmark) with every slip-weakening option"""
    deck = harness.deck("tpv3").replace(*edit)
    assert deck != harness.deck("tpv3")
    assert _fault_run(deck, 700) > 1e-2


@pytest.mark.parametrize("nor", ["kind=0", "kind=2, T=0.05d0", "kind=3, L=0.2d0, V=1d0"])
def test_normal_stress_laws(nor):
    """normal_update / normal_getSigma kinds 0, 2, 3 (bc_dynflt_normal.f90:118-147); kind 1 is the default of
    every other fault test.  A P-SV two-sided fault in a heterogeneous medium so that the normal stress varies."""
    deck = harness.cart_deck(24, 16, ezflt=8, scheme="newmark", nsteps=400)
    deck = deck.replace("&BC_DYNFLT_SWF Dc=0.4d0, MuS=0.677d0, MuD=0.525d0 /",
                        f"&BC_DYNFLT_SWF Dc=0.4d0, MuS=0.677d0, MuD=0.525d0 /\n&BC_DYNFLT_NOR {nor} /")
    o = orc.Oracle(deck, synthetic_seed=20261017)
    r = Rig(o)
    _lockstep(o, r, 400, 97, 1e-10)
    assert np.abs(o.arr("bc.0.D")).max() > 1e-3
    r.close()


@pytest.mark.parametrize("kind", [1, 2, 4])
def test_rate_and_state_kinds(kind):
    """rsf kinds 1 (strong velocity weakening, closed form), 2 (aging law), 4 (V-shape) on the RateState deck
    (kind 3, the slip law, is the deck's own: test_ratestate_run); T_stick is undefined for RSF"""
    deck = harness.deck("ratestate").replace("kind=3,", f"kind={kind},")
    assert f"kind={kind}," in deck
    _fault_run(deck, 400, tol=1e-9, skip_tstick=True)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_time_weakening_nucleation(kind):
    """friction = 'SWF','TWF': mu = min(mu_swf, mu_twf(t)) with the prescribed front of bc_dynflt_twf.f90:117-184"""
    deck = harness.deck("tpv3").replace("friction='SWF',", "friction='SWF','TWF',")
    deck = deck.replace("&BC_DYNFLT_SWF ", f"&BC_DYNFLT_TWF kind={kind}, MuS=0.677d0, MuD=0.525d0, Mu0=0.6d0, X=0d0, Z=0d0, "
                                           "V=2d3, L=1d3, T=2d0 /\n&BC_DYNFLT_SWF ")
    assert "BC_DYNFLT_TWF" in deck
    assert _fault_run(deck, 700) > 1e-2


def test_time_dependent_neumann_traction_equals_point_forces():
    """bc_DIRNEU_apply with a source time function (bc_dirneu.f90:160-167: f += stf(t)*B on the boundary
    nodes) has no oracle counterpart; it must equal the same load applied as collocated forces"""
    o = orc.Oracle(harness.deck("lamb"))
    top = [i for i in range(o.i("nbnd")) if o.i(f"bnd.{i}.tag") == 3][0]
    nodes = o.arr(f"bnd.{top}.node")
    rng = np.random.default_rng(2)
    Bv = rng.uniform(0.5, 1.5, nodes.size)
    Bh = rng.uniform(-0.3, 0.3, nodes.size)
    nsteps = 300
    amp = np.array([o.stf(0, (k + 1) * o.f("dt")) for k in range(nsteps)])
    out = []
    for mode in ("neumann", "forces"):
        from sem2dpack_b200 import Engine
        e = Engine(o.i("ngll"), 2, o.arr("ibool"), o.arr("H"), o.arr("rmass"), 0, o.f("dt"))
        e.set_elastic(o.i("nelast"), o.arr("a"), o.arr("elem2set"), False)
        if mode == "neumann":
            e.add_dirneu(nodes, 1, 1, B_h=Bh, B_v=Bv)
            e.commit()
            e.step(nsteps, None, np.stack([0.5 * amp, amp], axis=1))       # (nsteps, 2): h and v amplitudes
        else:
            for k, nd in enumerate(nodes):
                e.add_force(int(nd), [0.5 * Bh[k], Bv[k]])
            e.commit()
            e.step(nsteps, np.repeat(amp[:, None], nodes.size, axis=1))
        out.append(e.get_fields())
        e.close()
    for x, y in zip(*out):
        assert np.abs(x).max() > 0
        assert rel_l2(x, y) <= 1e-13
    o.close()
