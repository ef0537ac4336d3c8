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


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", ["testsh", "lamb", "tpv3", "ratestate"])
def test_fint_matches_oracle(name, variant):
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
def test_fint_all_ngll_hetero(ngll, ndof):
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


def test_testsh_full_run_and_known_answer():
    """EXAMPLES/TestSH end to end (1987 leapfrog steps, NGLL=6 SH, ABSORB, Ricker force, receivers 'D'),
    in lock step with the oracle AND against the analytic trace of analyze_test.m (< 2 %)."""
    o = orc.Oracle(harness.deck("testsh"))
    r = Rig(o)
    _lockstep(o, r, o.i("nt"), 500, TOL64)
    u = r.e.seis()[:, 4, 0].astype(np.float64)
    uref = harness.refdata()["testsh_uref"]
    assert np.abs(u - uref).max() / np.abs(uref).max() < 0.02
    r.close()


def test_lamb_full_run_and_known_answer():
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


def test_ratestate_run():
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
