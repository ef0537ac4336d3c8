"""CPU tests (no GPU): the oracle -- the CPU restatement of the reference's time-stepping path --
against every known-answer artefact the reference ships for this path, plus the
structural identities its mesh numbering must satisfy.  This is what pins the oracle; the GPU
parity tests then pin the CUDA path to the oracle.

  TestSH        EXAMPLES/TestSH/uyref.mat + analyze_test.m          (analytic, 2 % max-norm)
  LambsProblem  EXAMPLES/LambsProblem/U{x,z}_file_ascii + test.out   (0.5 % max-norm, 4 recorded misfits)
  RateState     EXAMPLES/RateState/{Tau,Ux,Vx}_{0,3,6,9}km_ascii     (no script / tolerance shipped: loose pin)
"""
import numpy as np
import pytest

import harness
import orc


@pytest.fixture(scope="module")
def golden():
    return harness.refdata()


def test_testsh_analytic_trace(golden):
    """analyze_test.m: max|uy - uyref| / max|uyref| < 2 % at station 5 (the one the script checks)."""
    o = orc.Oracle(harness.deck("testsh"))
    assert (o.i("ngll"), o.i("ndof"), o.i("nelem"), o.i("npoin")) == (6, 1, 3600, 301 * 301)
    assert o.i("nt") == 1987 and abs(o.f("dt") - 1.76209e-2) < 1e-6
    o.step(o.i("nt"))
    s = o.seis()  # (nt_rec, nx, ndof) float32
    assert s.shape == (1988, 7, 1)
    u = s[:, 4, 0].astype(np.float64)
    uref = golden["testsh_uref"]
    err = np.abs(u - uref).max() / np.abs(uref).max()
    assert err < 0.02, err
    o.close()


def test_lamb_known_answer_and_recorded_misfits(golden):
    """analyze_test.m of LambsProblem: every misfit < 0.5 %, and equal to the 4 numbers the
    reference recorded in test.out:12 to their printed precision."""
    o = orc.Oracle(harness.deck("lamb"))
    assert (o.i("ngll"), o.i("ndof"), o.i("nelem"), o.i("npoin")) == (9, 2, 800, 321 * 161)
    o.step(o.i("nt"))
    s = o.seis().astype(np.float64)  # (nt+1, 2 stations, 2 comps)
    uxa = np.vstack([np.zeros((1, 2)), golden["lamb_ux"].reshape(2, -1).T])
    uza = np.vstack([np.zeros((1, 2)), golden["lamb_uz"].reshape(2, -1).T])
    num = np.abs(np.hstack([s[:, :, 0] - uxa, s[:, :, 1] - uza])).max(axis=0)
    den = np.abs(np.hstack([uxa, uza])).max(axis=0)
    err = num / den
    assert err.max() < 0.005
    assert np.allclose(err, golden["lamb_misfits"], rtol=2e-3), (err, golden["lamb_misfits"])
    o.close()


def test_ratestate_series_loose_pin(golden):
    """EXAMPLES/RateState ships 803-sample slip / slip-rate / shear-stress series at x = 0,3,6,9 km
    without a comparison script or tolerance ('loosely pinned').  They were written by
    an earlier revision of the reference, so only loose bounds hold: slip within 1 %, slip rate
    within 6 %, stress within 15 % of the series maximum (max-norm; dominated by a one-sample shift
    of the rupture front), and the rupture front arrives within 2 samples."""
    o = orc.Oracle(harness.deck("ratestate"))
    p = "bc.0."
    np_, onx = o.i(p + "np"), o.i(p + "onx")
    assert (o.i("nelem"), o.i("npoin"), np_) == (24300, 1081 * 361, 1081)
    assert o.i("nt") == 803
    x = o.arr(p + "coord").reshape(-1, 2)[:, 0]
    o.step(o.i("nt"))
    out = o.arr(p + "out").reshape(-1, 6, onx)[1:]  # drop the it=0 record
    assert out.shape[0] == 803
    bounds = {"Ux": (0, 0.01), "Vx": (1, 0.06), "Tau": (2, 0.15)}
    for km in (0, 3, 6, 9):
        k = int(np.argmin(np.abs(x - km * 1e3)))
        for q, (col, tol) in bounds.items():
            ref = golden[f"ratestate_{q}_{km}km"]
            err = np.abs(out[:, col, k] - ref).max() / np.abs(ref).max()
            assert err < tol, (km, q, err)
        vref = golden[f"ratestate_Vx_{km}km"]
        t_ref = int(np.argmax(vref > 0.1 * vref.max()))
        t_got = int(np.argmax(out[:, 1, k] > 0.1 * vref.max()))
        assert abs(t_ref - t_got) <= 2, (km, t_ref, t_got)
    o.close()


def test_tpv3_rupture_physics():
    """TPV3 in-plane has no shipped series (parity unpinned by the reference): check
    the restatement against the SCEC TPV3 physics it must obey -- the rupture nucleates in the
    overstressed patch, never exceeds the P-wave speed, and slip stays one-signed."""
    o = orc.Oracle(harness.deck("tpv3"))
    p = "bc.0."
    onx, oitd = o.i(p + "onx"), o.i(p + "oitd")
    assert (o.i("ngll"), o.i("ndof"), o.i("nelem")) == (6, 2, 2800)
    assert abs(o.f("dt") - 5.38415e-3) < 1e-7 and o.i("nt") == 2972
    nsteps = 1200
    o.step(nsteps)
    out = o.arr(p + "out").reshape(-1, 6, onx)
    x = o.arr(p + "coord").reshape(-1, 2)[o.i(p + "oix1") - 1::o.i(p + "oixd"), 0][:onx]
    V = out[:, 1, :]
    assert out[-1, 0].max() > 0.5 and out[:, 0].min() > -1e-6
    t_arr = np.array([np.argmax(V[:, k] > 1e-3) if (V[:, k] > 1e-3).any() else -1 for k in range(onx)])
    ruptured = t_arr >= 0
    assert ruptured.sum() > 10
    k0 = int(np.argmin(np.where(ruptured, t_arr, 10 ** 9)))
    assert abs(x[k0]) <= 1.5e3 + 1.0  # inside the nucleation patch (half-width 1.5 km, symmetric model)
    dt = o.f("dt") * oitd
    for k in np.nonzero(ruptured)[0]:
        if abs(x[k] - x[k0]) > 3e3:
            speed = abs(x[k] - x[k0]) / max((t_arr[k] - t_arr[k0]) * dt, 1e-9)
            assert speed < 6000.0 * 1.05, (x[k], speed)
    o.close()


@pytest.mark.parametrize("nx,nz,ngll,ezflt", [(40, 40, 5, 0), (9, 7, 5, 3), (11, 6, 6, 2), (5, 4, 9, 0), (30, 30, 3, 15)])
def test_numbering_structure(nx, nz, ngll, ezflt):
    """SE_init_numbering identities: npoin = (nx(N-1)+1)(nz(N-1)+1) (+ one lattice row per fault);
    EXAMPLES/InaBox/info:191-192 (1600 elements, NGLL=5 -> 25921 points); every id used; valence of
    a node = 1 (interior), 2 (edge), 4 (vertex) except on the box sides and the fault."""
    o = orc.Oracle(harness.cart_deck(nx, nz, ngll=ngll, ezflt=ezflt, nrec=0, src=False, abso=(), fault=None))
    lx, lz = nx * (ngll - 1) + 1, nz * (ngll - 1) + 1
    npoin = lx * (lz + (1 if ezflt else 0))
    assert o.i("npoin") == npoin
    if (nx, nz, ngll, ezflt) == (40, 40, 5, 0):
        assert npoin == 25921
    ib = o.arr("ibool").reshape(nx * nz, ngll * ngll)
    assert ib.min() == 1 and ib.max() == npoin
    val = np.bincount(ib.ravel(), minlength=npoin + 1)[1:]
    assert val.min() == 1 and val.max() <= 4
    n4 = (nx - 1) * (nz - 1 - (1 if ezflt else 0))
    assert (val == 4).sum() == n4
    # element order is a permutation (RCM) and each element's nodes are distinct
    assert all(len(set(r.tolist())) == ngll * ngll for r in ib)
    o.close()


def test_rcm_is_a_permutation_with_small_bandwidth():
    """genrcm (SRC/rcm.f90): perm is a permutation of 1..E and the element-adjacency bandwidth after
    renumbering is of the order of the short side of the box (level sets are diagonals), far below
    the nx*nz of an arbitrary order."""
    nx, nz = 12, 7
    perm = orc.rcm(nx, nz)
    assert sorted(perm.tolist()) == list(range(1, nx * nz + 1))
    inv = np.empty(nx * nz, np.int64)
    inv[perm - 1] = np.arange(nx * nz)

    def bw(pos):
        b = 0
        for j in range(nz):
            for i in range(nx):
                for dj in (-1, 0, 1):
                    for di in (-1, 0, 1):
                        ii, jj = i + di, j + dj
                        if 0 <= ii < nx and 0 <= jj < nz:
                            b = max(b, abs(int(pos[i + nx * j]) - int(pos[ii + nx * jj])))
        return b
    assert bw(inv) <= 2 * min(nx, nz) + 2


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8, 9, 10])
def test_gll_tables(n):
    """get_GLL_info (SRC/gll.f90:19-36): end points +-1, symmetric, weights sum to 2 and integrate
    x^(2n-3) exactly; hprime differentiates polynomials of degree < n exactly; rows sum to 0."""
    x, w, H = orc.gll(n)  # H[i, j] = h'_i(x_j)
    assert x[0] == -1.0 and x[-1] == 1.0
    assert np.allclose(x, -x[::-1], atol=1e-15) and np.allclose(w, w[::-1], atol=1e-15)
    if n % 2:
        assert x[n // 2] == 0.0
    assert abs(w.sum() - 2.0) < 1e-14
    for p in range(0, 2 * n - 2):
        exact = 0.0 if p % 2 else 2.0 / (p + 1)
        assert abs((w * x ** p).sum() - exact) < 1e-13
    for p in range(n):
        f = x ** p
        df = p * x ** (p - 1) if p else np.zeros(n)
        assert np.allclose(H.T @ f, df, atol=1e-12)
    assert np.allclose(H.sum(axis=0), 0.0, atol=1e-13)


def test_fint_is_minus_K_d_symmetric_and_kills_rigid_motion():
    """compute_Fint (solver.f90:273-320) on the heterogeneous synthetic medium: K is symmetric
    (u.Kv == v.Ku) and a rigid translation produces no force."""
    o = orc.Oracle(harness.cart_deck(10, 8, nrec=0, src=False, abso=(), fault=None), synthetic_seed=20261017)
    n = o.i("npoin") * o.i("ndof")
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    z = np.zeros(n)
    o.set_fields(u, z)
    ku = o.compute_fint().copy()
    o.set_fields(v, z)
    kv = o.compute_fint().copy()
    assert abs(v @ ku - u @ kv) <= 1e-12 * abs(v @ ku)
    o.set_fields(np.ones(n), z)
    assert np.abs(o.compute_fint()).max() <= 1e-6 * np.abs(ku).max()
    o.close()


def test_synthetic_material_hash_is_partition_independent():
    """the benchmark medium depends only on the global GLL lattice coordinates."""
    L = orc.lib()
    a = L.orc_hash_u(20261017, 12345, 678, 1)
    assert -1.0 <= a <= 1.0
    assert a == L.orc_hash_u(20261017, 12345, 678, 1)
    assert a != L.orc_hash_u(20261017, 12346, 678, 1)
    vals = np.array([L.orc_hash_u(20261017, i, 3 * i + 1, 2) for i in range(4000)])
    assert abs(vals.mean()) < 0.05 and abs(vals.std() - 1 / np.sqrt(3)) < 0.03


def test_inabox_time_solver_lines():
    """EXAMPLES/InaBox/info:191-192,224-227 -- the reference's own log of this deck (NGLL=5 P-SV Newmark box, the
    KD2 path of the benchmark): 25921 GLL points, `Time step (secs) = 681.605E-06`, `Number of time steps = 2935`,
    `Total duration = 2.001E+00`.  Pins CHECK_grid's max(c/dx) and TIME_init (init.f90:187-225, time.f90:323-341)."""
    o = orc.Oracle(harness.deck("inabox"))
    assert o.i("npoin") == 25921 and o.i("nelem") == 1600
    assert f"{o.f('dt') * 1e6:.3f}" == "681.605"
    assert o.i("nt") == 2935
    assert f"{o.i('nt') * o.f('dt'):.3f}" == "2.001"
    assert abs(o.f("courant") - 0.3) < 1e-15
    o.close()


def test_damage_deck_prestrain_reproduces_the_fault_stresses():
    """EXAMPLES/Damage gives the medium an initial strain e0 = (-1.487381, -0.1708729, 0.35) and the fault the initial
    tractions Szz = -2, Sxz = 0.7: they are the same stress state only if xi_zero_2d, gamma_r_2d and compute_stress
    (mat_damage.f90:295-317,453-491) are restated right -- a consistency pin the reference deck itself holds."""
    deck = harness.deck("damage").replace("kind='DMG','KV'", "kind='DMG'").replace("nelem=240,100", "nelem=12,4")
    o = orc.Oracle(deck, renumber=False)
    par = o.arr("dm_par")[:16]
    lam, mu, xi0, gr = par[:4]
    assert abs(lam - 1.0) < 1e-6 and abs(mu - 1.0) < 1e-12      # cp = sqrt(3) cs
    assert -np.sqrt(2) < xi0 < 0 and gr > 0
    s0 = par[10:13]
    assert abs(s0[1] + 2.0) <= 1e-6 and abs(s0[2] - 0.7) <= 1e-6, s0
    o.close()


def test_visco_attenuation_fit_is_constant_q_within_the_documented_five_percent():
    """mat_visco.f90:58-61 documents get_attenuation's accuracy: 'For Nbody=3, constant Q with less than 5% error can be
    achieved over a maximum bandwidth fmax/fmin ~ 100'.  The oracle's restatement (relaxation frequencies, least-squares
    anelastic coefficients, theta) must reproduce that: 1/Q(w) of the generalized Maxwell body it builds, evaluated at
    200 frequencies across the band, stays within 5 % of 1/QP and 1/QS."""
    QP, QS, nb, fmin, fmax = 30.0, 20.0, 3, 1.8, 180.0
    deck = harness.deck("attenuation").replace("Nbody=5", f"Nbody={nb}").replace("nelem=44,44", "nelem=4,4")
    o = orc.Oracle(deck, renumber=False)
    theta = o.arr("mat.1.theta").reshape(3, nb)
    wb = o.arr("mat.1.wbody")
    lam, mu = o.arr("mat.1.moduli")
    assert np.allclose(wb / (2 * np.pi), [fmin, np.sqrt(fmin * fmax), fmax], rtol=1e-12)
    Ya, Yb = theta[0] / (lam + 2 * mu), theta[2] / (2 * mu)
    w = 2 * np.pi * np.exp(np.linspace(np.log(fmin), np.log(fmax), 200))
    for Y, Q in ((Ya, QP), (Yb, QS)):
        num = (Y[None, :] * wb[None, :] * w[:, None] / (wb[None, :] ** 2 + w[:, None] ** 2)).sum(axis=1)
        den = 1.0 - (Y[None, :] * wb[None, :] ** 2 / (wb[None, :] ** 2 + w[:, None] ** 2)).sum(axis=1)
        qinv = num / den
        assert np.abs(qinv * Q - 1.0).max() <= 0.05, np.abs(qinv * Q - 1.0).max()
    # unrelaxed moduli exceed the relaxed (reference-frequency) ones
    assert mu > 2000.0 * 2000.0 ** 2 and lam + 2 * mu > 2000.0 * 3000.0 ** 2
    o.close()


def test_plastic_return_puts_the_stress_on_the_coulomb_yield_surface():
    """MAT_PLAST_stress with Tv = 0 (vp_factor = 1, mat_plastic.f90:181-185,320-341) is the classical return mapping:
    wherever the trial stress exceeds the Coulomb yield stress Y = coh cos(phi) - sin(phi) sigma_m, the updated stress
    has max shear stress EXACTLY Y, the mean stress is unchanged and the plastic strain increment is deviatoric.  Checked
    on the oracle after one force evaluation of a large random displacement (absolute stress = snapshot + s0)."""
    nx, nz, h = 6, 5, 100.0
    coh, phi, e0 = 2.0e6, 30.0, (-4.0e-4, -3.0e-4, 2.5e-4)
    L = ["&GENERAL iexec=1, ngll=5, fmax=3.d0, ndof=2, title='plastic', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /", f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz} /",
         "&MATERIAL tag=1, kind='PLAST' /",
         f"&MAT_PLASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0, phi={phi}d0, coh={coh}d0, Tv=0d0, e0={e0[0]}d0,{e0[1]}d0,{e0[2]}d0 /",
         "&TIME NbSteps=10, courant=0.5d0, kind='leapfrog' /"]
    o = orc.Oracle("\n".join(L) + "\n", renumber=False)
    rng = np.random.default_rng(3)
    d = 3e-2 * rng.standard_normal(o.i("npoin") * 2)
    o.set_fields(d, d)
    o.compute_fint()                                  # advances ep
    ep = o.arr("pl_ep").reshape(nx * nz, 3, 25)
    assert np.abs(ep[:, 0] + ep[:, 1]).max() <= 1e-12 * np.abs(ep).max()       # deviatoric
    par = o.arr("pl_par")[:10]
    lam, mu = par[0], par[1]
    s0 = np.array([(lam + 2 * mu) * e0[0] + lam * e0[1], lam * e0[0] + (lam + 2 * mu) * e0[1], 2 * mu * e0[2]])
    s = o.snapshot("S").astype(np.float64).reshape(3, nx * nz, 25) + s0[:, None, None]   # C:(e - ep) + s0
    tau = np.sqrt(0.25 * (s[0] - s[1]) ** 2 + s[2] ** 2)
    Y = coh * np.cos(np.radians(phi)) - np.sin(np.radians(phi)) * 0.5 * (s[0] + s[1])
    yielded = np.abs(ep).max(axis=1) > 0
    assert yielded.mean() > 0.5
    # |Y|: in strong tension Y < 0 and the reference's factor Y/tau reverses the deviator instead of cutting it off
    # (no tension cut-off in mat_plastic.f90:331-336) -- restated as it is
    assert np.abs(tau[yielded] / np.abs(Y[yielded]) - 1.0).max() <= 1e-4   # float32 snapshot of stresses up to 5e8 Pa
    assert (tau[~yielded] <= Y[~yielded] * (1 + 1e-6)).all()
    o.close()
