"""sem2dsolve_b200 -- the C++ host of the path (host/: read_main, init_main, solve, REC_write over the
C-ABI) -- run as the reference's executable is run: a Par.inp in the working directory, output files
in the reference's formats.  Checked DIRECTLY against the reference's own known-answer artefacts
(no oracle in between) and, for the fault files, against the oracle.

  TestSH        EXAMPLES/TestSH/uyref.mat + analyze_test.m          analytic, 2 % max-norm
  LambsProblem  EXAMPLES/LambsProblem/U{x,z}_file_ascii + test.out   0.5 % max-norm, 4 recorded misfits
"""
import os
import subprocess

import numpy as np
import pytest

import harness
import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sem2dpack_b200", "lib", "sem2dsolve_b200")


def run(tmp_path, deck_text, *args):
    assert os.path.exists(EXE), "host program missing: run __graft_entry__.build()"
    (tmp_path / "Par.inp").write_text(deck_text)
    p = subprocess.run([EXE, *args], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    return p


def read_sep(tmp_path, name):
    """POST/sem2d_read_seis.m: header 'DT NSAMP NSTA', then one float32 record of NSAMP per station"""
    hdr = (tmp_path / "SeisHeader_sem2d.hdr").read_text().split("\n")
    dt, nsamp, nsta = hdr[1].split()
    coord = np.array([[float(v) for v in ln.split()] for ln in hdr[3:3 + int(nsta)]])
    u = np.fromfile(tmp_path / name, dtype=np.float32).reshape(int(nsta), int(nsamp)).T
    return float(dt), coord, u


def test_testsh_against_the_analytic_trace(tmp_path):
    p = run(tmp_path, harness.deck("testsh"))
    assert p.returncode == 0, p.stdout + p.stderr
    dt, coord, u = read_sep(tmp_path, "Uy_sem2d.dat")
    assert u.shape == (1988, 7) and abs(dt - 1.76209e-2) < 1e-6
    assert np.allclose(coord[:, 0], np.linspace(0, 30, 7)) and np.allclose(coord[:, 1], 0)
    uref = harness.refdata()["testsh_uref"]
    err = np.abs(u[:, 4].astype(np.float64) - uref).max() / np.abs(uref).max()
    assert err < 0.02, err   # analyze_test.m


def test_lamb_against_the_recorded_misfits(tmp_path):
    p = run(tmp_path, harness.deck("lamb"))
    assert p.returncode == 0, p.stdout + p.stderr
    g = harness.refdata()
    _, _, ux = read_sep(tmp_path, "Ux_sem2d.dat")
    _, _, uz = read_sep(tmp_path, "Uz_sem2d.dat")
    uxa = np.vstack([np.zeros((1, 2)), g["lamb_ux"].reshape(2, -1).T])
    uza = np.vstack([np.zeros((1, 2)), g["lamb_uz"].reshape(2, -1).T])
    num = np.abs(np.hstack([ux - uxa, uz - uza])).max(axis=0)
    den = np.abs(np.hstack([uxa, uza])).max(axis=0)
    err = num / den
    assert err.max() < 0.005
    assert np.allclose(err, g["lamb_misfits"], rtol=2e-3), (err, g["lamb_misfits"])


def test_fault_files_against_the_oracle(tmp_path):
    """a MESH_CART ezflt deck with the split-node slip-weakening fault: FltXX_sem2d.dat (sequential
    unformatted records), the potency table and the seismograms equal the oracle's after the same
    float32 cast"""
    nx, nz, nsteps = 24, 16, 200
    deck = harness.cart_deck(nx, nz, ezflt=8, nsteps=nsteps)
    p = run(tmp_path, deck, "--hash-seed", "20261017")    # the heterogeneous medium of the benchmark family
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, synthetic_seed=20261017, renumber=False)
    o.step(nsteps)
    _, _, ux = read_sep(tmp_path, "Ux_sem2d.dat")
    ref = o.seis()
    assert np.abs(ux - ref[:, :, 0]).max() <= 2e-7 * np.abs(ref).max()
    hdr = (tmp_path / "Flt05_sem2d.hdr").read_text().split("\n")
    npts, ndat, nsamp = (int(v) for v in hdr[1].split()[:3])
    assert (npts, ndat, nsamp) == (o.i("bc.0.onx"), 6, nsteps + 1)
    raw = np.fromfile(tmp_path / "Flt05_sem2d.dat", dtype=np.int32).reshape(nsamp, ndat, npts + 2)
    assert (raw[:, :, 0] == 4 * npts).all() and (raw[:, :, -1] == 4 * npts).all()   # record markers
    rec = raw[:, :, 1:-1].copy().view(np.float32)
    want = o.arr("bc.0.out").reshape(-1, 6, npts)
    for c in range(6):
        assert np.abs(rec[:, c] - want[:, c]).max() <= 2e-7 * max(np.abs(want[:, c]).max(), 1e-30), c
    assert rec[-1, 0].max() > 1e-3   # the fault did slip
    pot = np.loadtxt(tmp_path / "Flt05_potency_sem2d.tab")
    assert pot.shape == (nsteps + 1, 6)
    o.close()


def read_fault(tmp_path, tag):
    """POST/python/sem2d_read_fault.py: header, then NSAMP x NDAT records framed by 4-byte lengths"""
    hdr = (tmp_path / f"Flt{tag:02d}_sem2d.hdr").read_text().split("\n")
    npts, ndat, nsamp = (int(v) for v in hdr[1].split()[:3])
    x = np.array([[float(v) for v in ln.split()] for ln in hdr[4:4 + npts]])
    raw = np.fromfile(tmp_path / f"Flt{tag:02d}_sem2d.dat", dtype=np.int32).reshape(nsamp, ndat, npts + 2)
    assert (raw[:, :, 0] == 4 * npts).all() and (raw[:, :, -1] == 4 * npts).all()
    return x, raw[:, :, 1:-1].copy().view(np.float32)


def test_ratestate_deck_against_the_shipped_series_and_the_oracle(tmp_path):
    """EXAMPLES/RateState through the host program: SH, one-sided rate-and-state fault on the bottom side
    (tag 1, slip law, ORDER0 distributions of the initial shear stress and state), DIRNEU sides, absorbing
    top, Newmark, fztag.  Against the 12 series the reference ships (the loose bounds of
    tests/test_oracle_golden.py: they predate the current RSF solver) and against the oracle (float32)."""
    deck = harness.deck("ratestate")
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    x, rec = read_fault(tmp_path, 1)
    assert rec.shape == (804, 6, 1081)
    out = rec[1:]
    g = harness.refdata()
    bounds = {"Ux": (0, 0.01), "Vx": (1, 0.06), "Tau": (2, 0.15)}
    for km in (0, 3, 6, 9):
        k = int(np.argmin(np.abs(x[:, 0] - km * 1e3)))
        for q, (col, tol) in bounds.items():
            ref = g[f"ratestate_{q}_{km}km"]
            err = np.abs(out[:, col, k] - ref).max() / np.abs(ref).max()
            assert err < tol, (km, q, err)
    o = orc.Oracle(deck, renumber=False)
    o.step(o.i("nt"))
    want = o.arr("bc.0.out").reshape(-1, 6, 1081)
    for c in range(5):   # column 6 (T_stick) is not defined for rate-and-state faults (parity trap 3)
        assert np.abs(rec[:, c] - want[:, c]).max() <= 2e-6 * max(np.abs(want[:, c]).max(), 1e-30), c
    _, _, uy = read_sep(tmp_path, "Uy_sem2d.dat")
    ref = o.seis()[:, :, 0]
    assert np.abs(uy - ref).max() <= 2e-6 * np.abs(ref).max()
    o.close()


def test_tpv3_deck_against_the_oracle(tmp_path):
    """EXAMPLES/TestFlt2D_SCEC_TPV3_inplane through the host program: NGLL=6 P-SV, ELAST + Kelvin-Voigt with
    a GAUSSIAN eta, one-sided slip-weakening fault on the bottom side with PWCONR Tt and MuS, absorbing
    sides 2,3, DIRNEU side 4 (v='D'), Newmark, 10 interpolated stations (AtNode=F, isamp=20), fault
    output nodes 1:65:4.  No artefact of the reference pins this deck: checked against the
    oracle (float32 files) and the SCEC TPV3 physics the oracle test uses."""
    deck = harness.deck("tpv3").replace("TotalTime=16.d0", "TotalTime=7.d0")
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    nt = o.i("nt")
    o.step(nt)
    x, rec = read_fault(tmp_path, 1)
    onx = o.i("bc.0.onx")
    want = o.arr("bc.0.out").reshape(-1, 6, onx)
    assert rec.shape == want.shape and onx == 17
    for c in range(6):
        assert np.abs(rec[:, c] - want[:, c]).max() <= 5e-6 * max(np.abs(want[:, c]).max(), 1e-30), c
    assert rec[-1, 0].max() > 0.5 and rec[:, 0].min() > -1e-6      # it ruptured, slip is one-signed
    _, coord, ux = read_sep(tmp_path, "Ux_sem2d.dat")
    _, _, uz = read_sep(tmp_path, "Uz_sem2d.dat")
    ref = o.seis()
    assert ux.shape == ref[:, :, 0].shape == (nt // 20 + 1, 10)
    assert np.allclose(coord[:, 0], np.linspace(0, 15e3, 10)) and np.allclose(coord[:, 1], 500.0)
    for got, c in ((ux, 0), (uz, 1)):
        assert np.abs(got - ref[:, :, c]).max() <= 5e-6 * np.abs(ref[:, :, c]).max(), c
    o.close()


@pytest.mark.parametrize("renumber", [True, False])
def test_binary_snapshots_and_grid_files(tmp_path, renumber):
    """&SNAP_DEF bin=T: PLOT_FIELD's node-wise float32 files and the grid files POST/ reads them with.  By default in
    the element order and node numbering of a stock reference build (OPT_RENUMBER = .true., constants.f90:10-15):
    ibool_sem2d.dat is bit for bit the oracle's RCM-ordered table; --natural-order gives the row-by-row order."""
    deck = harness.deck("lamb").replace("TotalTime=1.5d0, Dt=0.5d-3", "NbSteps=250, Dt=0.5d-3")
    deck = deck.replace("&SNAP_DEF itd=5000, fields='V'/", "&SNAP_DEF itd=100, fields='DV', ps=F /")
    p = run(tmp_path, deck, "--quiet", *([] if renumber else ["--natural-order"]))
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=renumber)
    npoin, nelem = o.i("npoin"), o.i("nelem")
    hdr = (tmp_path / "grid_sem2d.hdr").read_text().split("\n")[1].split()
    assert [int(v) for v in hdr] == [nelem, 41 * 21, 4, npoin, 9]
    ib = np.fromfile(tmp_path / "ibool_sem2d.dat", dtype=np.int32)
    assert np.array_equal(ib, o.arr("ibool"))
    co = np.fromfile(tmp_path / "coord_sem2d.dat", dtype=np.float32).reshape(npoin, 2)
    assert np.abs(co - o.arr("coord").reshape(npoin, 2)).max() <= 1e-3
    for snap, it in ((0, 0), (1, 100), (2, 200)):
        if it:
            o.step(100)
        for name, key, c in (("dx", "d", 0), ("dz", "d", 1), ("vx", "v", 0), ("vz", "v", 1)):
            got = np.fromfile(tmp_path / f"{name}_{snap:03d}_sem2d.dat", dtype=np.float32)
            ref = o.arr(key)[c * npoin:(c + 1) * npoin]
            assert got.shape == (npoin,)
            assert np.abs(got - ref).max() <= 2e-7 * max(np.abs(ref).max(), 1e-30), (name, snap)
    assert not (tmp_path / "vx_003_sem2d.dat").exists()
    o.close()


def test_unsupported_input_aborts_like_io_abort(tmp_path):
    """what the host does not provide is refused the way the reference refuses bad input: message +
    non-zero exit (IO_abort, stdio.f90:205-214), never ignored"""
    p = run(tmp_path, harness.deck("tpv3").replace("'ELAST' ,'KV'", "'ELAST' ,'DMG'"))
    assert p.returncode == 1 and "FATAL ERROR" in p.stdout and "MAT_read" in p.stdout
    p = run(tmp_path, harness.deck("testsh").replace("courant = 0.3d0", "courant = 0.9d0"))
    assert p.returncode == 1 and "Courant out of range" in p.stdout


def test_velocity_weakening_deck_two_materials_and_a_kelvin_voigt_layer(tmp_path):
    """EXAMPLES/Velocity_weakening unchanged but for the run length: NGLL=6 P-SV, fztag=2 -> tag 1 ELAST, tag 2
    ELAST+KV (one element row next to the fault), one-sided rate-and-state fault, ABSORB, DIRNEU, leapfrog.  The host
    evaluates both materials at the GLL points (s2d_cart_set_material) and hands the KV row over element by element
    (s2d_cart_set_kv_elems); fault records and potency against the oracle after the same float32 cast."""
    deck = harness.deck("velweak").replace("TotalTime=0.10d0", "NbSteps=400")
    assert "NbSteps=400" in deck
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    assert o.i("nkv") == 100 and o.i("ncoefsets") == 2
    o.step(400)
    x, rec = read_fault(tmp_path, 1)
    want = o.arr("bc.0.out").reshape(-1, 6, rec.shape[2])
    assert rec.shape == want.shape
    for c in range(5):     # T_stick is undefined for rate-and-state faults
        assert np.abs(rec[:, c] - want[:, c]).max() <= 1e-6 * max(np.abs(want[:, c]).max(), 1e-30), c
    assert np.abs(rec[-1, 1]).max() > 1e-3            # the fault slips (nucleation patch)
    pot = np.loadtxt(tmp_path / "Flt01_potency_sem2d.tab")
    ref = o.arr("bc.0.potency").reshape(-1, pot.shape[1])
    assert np.abs(pot - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300)
    o.close()


def test_material_distributions_per_tag(tmp_path):
    """MAT_read / MAT_init_prop (mat_gen.f90:101-303): two MESH_CART_DOMAIN tags, the upper one with an ORDER0
    density and a GAUSSIAN shear velocity -- seismograms against the oracle"""
    deck = """&GENERAL iexec=1, ngll=5, fmax=3.d0, ndof=2, title='two materials', verbose='0000', ItInfo=1000 /
&MESH_DEF method='CARTESIAN' /
&MESH_CART xlim=0d0,2400d0, zlim=0d0,1600d0, nelem=24,16 /
&MESH_CART_DOMAIN tag=1, ex=1,24, ez=1,7 /
&MESH_CART_DOMAIN tag=2, ex=1,24, ez=8,16 /
&MATERIAL tag=1, kind='ELAST' /
&MAT_ELASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0 /
&MATERIAL tag=2, kind='ELAST' /
&MAT_ELASTIC rhoH='ORDER0', cp=5200.d0, csH='GAUSSIAN' /
&DIST_ORDER0 xn=2, zn=1 /
1250d0
2500d0 2300d0
&DIST_GAUSSIAN centered_at=1200d0,1200d0, length=700d0,500d0, offset=2800d0, ampli=-400d0, order=1 /
&BC_DEF tag=1, kind='ABSORB' /
&BC_DEF tag=2, kind='ABSORB' /
&BC_DEF tag=4, kind='ABSORB' /
&TIME NbSteps=300, courant=0.5d0, kind='newmark' /
&SRC_DEF stf='RICKER', coord=900d0,1000d0, mechanism='FORCE' /
&STF_RICKER f0=4.d0, onset=0.3d0, ampli=1.d9 /
&SRC_FORCE angle=30d0 /
&REC_LINE number=9, field='V', first=200d0,300d0, last=2200d0,1400d0, isamp=1 /
"""
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    assert o.i("ncoefsets") > 2
    o.step(300)
    dt, _, ux = read_sep(tmp_path, "Ux_sem2d.dat")
    _, _, uz = read_sep(tmp_path, "Uz_sem2d.dat")
    assert abs(dt - o.f("dt")) <= 1e-6 * dt
    ref = o.seis()
    assert np.abs(ref).max() > 0
    assert np.abs(ux - ref[:, :, 0]).max() <= 2e-6 * np.abs(ref).max()
    assert np.abs(uz - ref[:, :, 1]).max() <= 2e-6 * np.abs(ref).max()
    o.close()


def test_2p5d_inplane_deck_through_the_host_program(tmp_path):
    """EXAMPLES/2.5D_inplane (W = 10 km) shortened: the host hands W to the builder (s2d_cart_set_w25d), which forms
    beta at every GLL point; fault records against the oracle"""
    deck = harness.deck("inplane25d").replace("TotalTime=30", "NbSteps=300")
    assert "NbSteps=300" in deck
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    o.step(300)
    x, rec = read_fault(tmp_path, 1)
    want = o.arr("bc.0.out").reshape(-1, 6, rec.shape[2])
    assert rec.shape == want.shape
    for c in range(6):
        assert np.abs(rec[:, c] - want[:, c]).max() <= 1e-6 * max(np.abs(want[:, c]).max(), 1e-30), c
    assert np.abs(rec[-1, 0]).max() > 0
    o.close()


@pytest.mark.parametrize("kind", ["USER", "TAB"])
def test_user_and_tabulated_source_time_functions(tmp_path, kind):
    """STF_USER_fun (stf_user.f90:66-79) and STF_TAB_fun (stf_tabulated.f90:76-93: cubic spline with zero end slopes,
    clamped to the table) through the host program, against the same deck driven through the generic C-ABI with the
    amplitude table evaluated here (scipy's clamped CubicSpline for TAB)"""
    from scipy.interpolate import CubicSpline
    nsteps = 300
    base = harness.deck("testsh").replace("TotalTime=35.d0", f"NbSteps={nsteps}")
    o = orc.Oracle(base)
    dt = o.f("dt")
    t = (np.arange(nsteps) + 1) * dt
    if kind == "USER":
        deck = base.replace("'RICKER'", "'USER'") + "&STF_USER ampli=2.0, onset=0.25, par1=0.05 /\n"
        arg = t - np.float64(np.float32(0.25))
        amp = np.float64(np.float32(2.0)) * np.sin(arg) + np.float64(np.float32(0.05)) * arg ** 2
    else:
        tt = np.linspace(0.5, 4.0, 36)
        vv = np.exp(-((tt - 2.0) / 0.6) ** 2) * np.cos(3.0 * tt)
        (tmp_path / "stf.tab").write_text("".join(f"{a:.16e} {b:.16e}\n" for a, b in zip(tt, vv)))
        deck = base.replace("'RICKER'", "'TAB'") + "&STF_TAB file='stf.tab' /\n"
        amp = CubicSpline(tt, vv, bc_type=((1, 0.0), (1, 0.0)))(np.clip(t, tt[0], tt[-1]))
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    _, _, u = read_sep(tmp_path, "Uy_sem2d.dat")
    r = harness.Rig(o)
    r.e.step(nsteps, amp.reshape(-1, 1))
    ref = r.e.seis()[:, :, 0]
    assert np.abs(ref).max() > 0
    assert np.abs(u - ref).max() <= 1e-6 * np.abs(ref).max()
    r.close()


def test_strain_stress_divcurl_snapshots(tmp_path):
    """&SNAP_DEF fields='ESdc' (plot_gen.f90:222-305): e11/e22/e12, s11/s22/s12, div, curl -- one record of
    ngll*ngll reals per element, RCM element order -- computed on the device (s2d_cart_snapshot_elem) against the
    oracle's restatement of FIELD_strain_elem / MAT_stress_dv / FIELD_divcurl_elem; the TPV3 deck so that the stress
    carries the Kelvin-Voigt d + eta*v"""
    deck = harness.deck("tpv3").replace("TotalTime=16.d0", "NbSteps=300")
    deck = deck.replace("&SNAP_DEF itd=200, fields ='V',bin=F,ps=T /", "&SNAP_DEF itd=300, fields='ESdc', bin=T, ps=F /")
    assert "fields='ESdc'" in deck
    p = run(tmp_path, deck, "--quiet")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=True)
    o.step(300)
    ne, n2 = o.i("nelem"), o.i("ngll") ** 2
    for what, names in (("E", ("e11", "e22", "e12")), ("S", ("s11", "s22", "s12")), ("d", ("div",)), ("c", ("curl",))):
        ref = o.snapshot(what).reshape(len(names), ne * n2)
        for k, nm in enumerate(names):
            got = np.fromfile(tmp_path / f"{nm}_001_sem2d.dat", dtype=np.float32)
            assert got.shape == (ne * n2,)
            assert np.abs(ref[k]).max() > 0
            assert np.abs(got - ref[k]).max() <= 2e-6 * np.abs(ref[k]).max(), nm
    o.close()


def test_energy_table(tmp_path):
    """--energies = a reference build with COMPUTE_ENERGIES (constants.f90:22-27): energy_sem2d.tab, one line
    `time, E_ep, E_k, E_el, E_W` per step (main.f90:90-93, energy.f90:49-116); 2.5D deck so that E_W is non-zero"""
    nsteps = 60
    deck = harness.deck("inplane25d").replace("TotalTime=30", f"NbSteps={nsteps}")
    p = run(tmp_path, deck, "--quiet", "--energies")
    assert p.returncode == 0, p.stdout + p.stderr
    tab = np.loadtxt(tmp_path / "energy_sem2d.tab")
    assert tab.shape == (nsteps, 5)
    o = orc.Oracle(deck)
    for k in range(nsteps):
        o.step(1)
        ek, ew = o.L.orc_energy_Ek(o.h), o.L.orc_energy_EW(o.h)
        assert abs(tab[k, 0] - (k + 1) * o.f("dt")) <= 1e-12 * tab[k, 0]
        assert tab[k, 1] == 0.0 and tab[k, 3] == 0.0
        assert abs(tab[k, 2] - ek) <= 1e-9 * max(ek, 1e-300), (k, tab[k, 2], ek)
        assert abs(tab[k, 4] - ew) <= 1e-9 * max(ew, 1e-300), (k, tab[k, 4], ew)
    assert tab[-1, 2] > 0 and tab[-1, 4] > 0
    o.close()


def read_snapshot(tmp_path, name, it, nelem, ngll):
    f = tmp_path / f"{name}_{it:03d}_sem2d.dat"
    return np.fromfile(f, dtype=np.float32)


@pytest.mark.parametrize("nsteps,coh", [(400, "9.669501d6"), (250, "4.0d6")])
def test_plastic_deck_coulomb_yielding_off_the_fault(tmp_path, nsteps, coh):
    """EXAMPLES/2.5D_plastic (kind='PLAST', W = 10 km, SWF + TWF fault, ABSORB x3 + DIRNEU, leapfrog), coarsened
    to 48 x 48 elements: MAT_Fint's strain -> MAT_PLAST_stress(update) -> MAT_forces branch (mat_gen.f90:445-449,
    mat_plastic.f90:281-387) runs inside the strip kernel with the plastic strain of every element GLL point resident
    in HBM.  Fields (snapshots D, V), the stress snapshot (relative stress from e - ep) and the fault records against
    the oracle; the second case lowers the cohesion so that the medium yields from the first steps.
    No reference artefact pins this deck (none ships): oracle parity."""
    deck = harness.deck("plastic25d").replace("nelem=160,160", "nelem=48,48").replace("TotalTime=80", f"NbSteps={nsteps}")
    deck = deck.replace("coh = 9.669501d6", f"coh = {coh}").replace("itd=1000", f"itd={nsteps}").replace("fields ='DVSE'", "fields ='DVS'")
    assert f"NbSteps={nsteps}" in deck and f"coh = {coh}" in deck and f"itd={nsteps}" in deck
    p = run(tmp_path, deck, "--quiet", "--natural-order")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    assert o.i("npl") == 48 * 48
    o.step(nsteps)
    ep = o.arr("pl_ep")
    assert np.abs(ep).max() > 1e-6          # the medium has yielded
    n = o.i("npoin")
    for nm, key in (("d", "d"), ("v", "v")):
        ref = o.arr(key).reshape(2, n)
        for c, ax in enumerate("xz"):
            got = np.fromfile(tmp_path / f"{nm}{ax}_001_sem2d.dat", dtype=np.float32)
            assert got.size == n
            assert np.abs(got - ref[c].astype(np.float32)).max() <= 2e-6 * np.abs(ref[c]).max(), (nm, ax)
    sref = o.snapshot("S")                   # (3, nelem, ngll, ngll) float32, relative stress C:(e - ep)
    for c, nm in enumerate(("s11", "s22", "s12")):
        got = np.fromfile(tmp_path / f"{nm}_001_sem2d.dat", dtype=np.float32)
        assert got.size == sref[c].size
        assert np.abs(got - sref[c].ravel()).max() <= 5e-6 * np.abs(sref).max(), nm
    x, rec = read_fault(tmp_path, 5)
    want = o.arr("bc.0.out").reshape(-1, 6, rec.shape[2])
    assert rec.shape == want.shape
    for c in range(6):
        assert np.abs(rec[:, c] - want[:, c]).max() <= 1e-6 * max(np.abs(want[:, c]).max(), 1e-30), c
    o.close()


def test_attenuation_deck_visco_elastic_medium(tmp_path):
    """EXAMPLES/Attenuation (kind='VISCO': Nbody = 5 mechanisms fitted to QP = 30, QS = 20 over 1.8-180 Hz, NGLL 6,
    Ricker force, four absorbing sides), first 500 of 1004 steps: the host evaluates get_attenuation (its own
    least-squares solver) and hands theta / wbody / unrelaxed moduli to s2d_cart_set_visco; the seismogram and the
    velocity snapshot against the oracle, and against an elastic run of the same deck (the pulse must be damped).
    No reference artefact pins this deck (benchmark_attenuation.m needs an external analytical code): oracle parity."""
    nsteps = 500
    deck = harness.deck("attenuation").replace("TotalTime=0.75d0", f"NbSteps={nsteps}").replace("itd=10000", f"itd={nsteps}")
    assert f"NbSteps={nsteps}" in deck and f"itd={nsteps}" in deck
    p = run(tmp_path, deck, "--quiet", "--natural-order")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    assert o.i("nvs") == 44 * 44
    o.step(nsteps)
    dt, coord, ux = read_sep(tmp_path, "Ux_sem2d.dat")
    ref = o.seis()[:, 0, 0]
    assert ux.shape[0] == ref.shape[0]
    assert np.abs(ux[:, 0] - ref).max() <= 2e-6 * np.abs(ref).max()
    n = o.i("npoin")
    vref = o.arr("v").reshape(2, n)
    got = np.fromfile(tmp_path / "vx_001_sem2d.dat", dtype=np.float32)
    assert np.abs(got - vref[0].astype(np.float32)).max() <= 2e-6 * np.abs(vref[0]).max()
    # the same deck without attenuation: larger peak at the receiver
    el = deck.replace("kind='VISCO'", "kind='ELAST'").replace(
        "&MAT_VISCO rho=2000d0, cp=3000d0, cs=2000d0, QP=30d0, QS=20d0, Nbody=5,fmin=1.8d0,fmax=180d0 /",
        "&MAT_ELASTIC rho=2000d0, cp=3000d0, cs=2000d0 /")
    assert "MAT_ELASTIC" in el
    oe = orc.Oracle(el, renumber=False)
    oe.step(nsteps)
    assert np.abs(oe.seis()[:, 0, 0]).max() > 1.1 * np.abs(ref).max()
    oe.close()
    o.close()


def test_damage_deck_off_fault_damage(tmp_path):
    """EXAMPLES/Damage unchanged but for size and length (60 x 24 elements, 500 steps): kind='DMG' on tag 1,
    kind='DMG','KV' on the element rows next to the fault (fztag = 2: the damage rheology under a Kelvin-Voigt layer,
    d + eta*v before the constitutive law, mat_gen.f90:435), background stress Szz / Sxz on the SWF + TWF fault, four
    absorbing sides, leapfrog, nondimensional units.  Rupture nucleates, the off-fault medium accumulates damage and
    damage-related plastic strain.  Velocity snapshot and fault records against the oracle.  No reference artefact
    pins the time series (the deck's prestrain / fault-stress consistency is pinned in tests/test_oracle_golden.py)."""
    nsteps = 500
    deck = harness.deck("damage").replace("nelem=240,100", "nelem=60,24")
    deck = deck.replace("TotalTime=30d0", f"NbSteps={nsteps}").replace("itd=200", f"itd={nsteps}").replace("iexec=0", "iexec=1")
    assert f"NbSteps={nsteps}" in deck and "nelem=60,24" in deck and "'DMG','KV'" in deck
    p = run(tmp_path, deck, "--quiet", "--natural-order")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    assert o.i("ndm") == 60 * 24 and o.i("nkv") == 2 * 60
    o.step(nsteps)
    st = o.arr("dm_state").reshape(-1, 4, 25)
    assert st[:, 0].max() > 0.05 and np.abs(st[:, 1:]).max() > 1e-3      # damage and plastic strain have grown
    n = o.i("npoin")
    vref = o.arr("v").reshape(2, n)
    got = np.fromfile(tmp_path / "vx_001_sem2d.dat", dtype=np.float32)
    assert np.abs(got - vref[0].astype(np.float32)).max() <= 5e-6 * np.abs(vref[0]).max()
    x, rec = read_fault(tmp_path, 5)
    want = o.arr("bc.0.out").reshape(-1, 6, rec.shape[2])
    assert rec.shape == want.shape
    for c in range(6):
        assert np.abs(rec[:, c] - want[:, c]).max() <= 5e-6 * max(np.abs(want[:, c]).max(), 1e-30), c
    o.close()


@pytest.mark.parametrize("rheology", ["plastic", "visco"])
def test_kelvin_voigt_layer_on_top_of_plasticity_and_visco_elasticity(tmp_path, rheology):
    """KV is the one non-exclusive material (mat_gen.f90:350-354,435): kind='PLAST','KV' and kind='VISCO','KV' on the
    element rows next to the fault (fztag = 2), the plain rheology elsewhere; SWF fault, absorbing sides, leapfrog.
    Velocity snapshot and fault records against the oracle."""
    nsteps = 300
    if rheology == "plastic":
        kind, block = "PLAST", "&MAT_PLASTIC cp=5770.d0, cs=3330.d0, rho=2705.d0, phi = 30.d0, coh = 4.0d6, Tv = 0.0561d0,\n   e0 = -4.162378e-04, -4.162378e-04, 3.504802e-04 /"
    else:
        kind, block = "VISCO", "&MAT_VISCO rho=2705.d0, cp=5770.d0, cs=3330.d0, QP=60d0, QS=30d0, Nbody=3, fmin=0.05d0, fmax=5d0 /"
    deck = f"""&GENERAL iexec=1, ngll=5, fmax=3.d0 , ndof=2 , title = 'rheology under a KV layer', verbose='0000' , ItInfo = 400/
&MESH_DEF  method = 'CARTESIAN'/
&MESH_CART ezflt=-1, fztag=2, xlim=0d3,30d3, zlim=-12d3,12d3, nelem=40,32/
&MATERIAL tag=1, kind='{kind}'  /
&MATERIAL tag=2, kind='{kind}','KV'  /
{block}
&MAT_KV eta=0.2d0 /
&BC_DEF  tags = 5,6 , kind = 'DYNFLT' /
&BC_DYNFLT friction='SWF','TWF', Tn=-50d6, Tt=2.102564d7 /
&BC_DYNFLT_SWF Dc=2d0, MuS=0.6d0, MuD=0.1d0 /
&BC_DYNFLT_TWF kind=1, MuS=0.6d0, MuD=0.1d0, Mu0=0.6d0, X=15.d3, Z=0.d0, V=0.333d3, L=0.1665d3, T=60d0 /
&BC_DEF  tag = 1 , kind = 'ABSORB' /
&BC_DEF  tag = 2 , kind = 'ABSORB' /
&BC_DEF  tag = 3 , kind = 'ABSORB' /
&BC_DEF  tag = 4 , kind = 'ABSORB' /
&TIME  kind='leapfrog', NbSteps={nsteps}, courant=0.4d0 /
&SNAP_DEF itd={nsteps}, fields ='V',bin=T,ps=F /
"""
    p = run(tmp_path, deck, "--quiet", "--natural-order")
    assert p.returncode == 0, p.stdout + p.stderr
    o = orc.Oracle(deck, renumber=False)
    assert o.i("nkv") == 2 * 40 and o.i("npl" if rheology == "plastic" else "nvs") == 40 * 32
    o.step(nsteps)
    n = o.i("npoin")
    vref = o.arr("v").reshape(2, n)
    assert np.abs(vref).max() > 0
    for c, ax in enumerate("xz"):
        got = np.fromfile(tmp_path / f"v{ax}_001_sem2d.dat", dtype=np.float32)
        assert np.abs(got - vref[c].astype(np.float32)).max() <= 5e-6 * np.abs(vref).max(), ax
    x, rec = read_fault(tmp_path, 5)
    want = o.arr("bc.0.out").reshape(-1, 6, rec.shape[2])
    assert rec.shape == want.shape
    floor = 1e-6 * np.abs(want).max()   # a column at rounding level (normal-stress change of a symmetric problem: 1e-7 Pa)
    for c in range(6):
        assert np.abs(rec[:, c] - want[:, c]).max() <= 5e-6 * max(np.abs(want[:, c]).max(), floor), c
    o.close()
