"""HHT-alpha and symplectic time schemes (SRC/solver.f90:89-128,169-199) and
moment-tensor sources (SRC/src_moment.f90), GPU against the oracle through the C-ABI.  No reference
artefact pins these (parity unpinned by the reference, as for SWF): the oracle restates the few
lines of Fortran involved and the bar is the usual relative L2 <= 1e-10 in FP64."""
import os
import subprocess

import numpy as np
import pytest

import harness
import orc
from harness import rel_l2
from sem2dpack_b200 import CartEngine

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sem2dpack_b200", "lib", "sem2dsolve_b200")


def _run(deck, nsteps, **okw):
    o = orc.Oracle(deck, **okw)
    r = harness.Rig(o)
    o.step(nsteps)
    r.step(nsteps // 2)
    r.step(nsteps - nsteps // 2)
    d, v, a = r.e.get_fields()
    out = (rel_l2(d, o.arr("d")), rel_l2(v, o.arr("v")), rel_l2(a, o.arr("acc")))
    seis = (r.e.seis(), o.seis()) if o.i("rec.present") else None
    return out, seis, r


@pytest.mark.parametrize("deck,edit", [
    ("testsh", ("courant = 0.3d0", "courant = 0.3d0, kind='HHT-alpha'")),                       # SH, ABSORB
    ("lamb", ("Dt=0.5d-3 /", "Dt=0.5d-3, kind='HHT-alpha' /\n&TIME_HHTA alpha=0.7d0, rho=0.8d0 /")),  # P-SV, ngll=9
])
def test_hht_alpha_decks(deck, edit):
    text = harness.deck(deck).replace(*edit)
    assert "HHT-alpha" in text
    (ed, ev, ea), (s_got, s_ref), r = _run(text, 300)
    assert r.o.i("scheme") == 2 and r.o.f("gamma") == pytest.approx(1.5 - r.o.f("alpha"))
    assert ed <= 1e-10 and ev <= 1e-10 and ea <= 1e-10, (ed, ev, ea)
    assert np.abs(s_got - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    assert np.abs(s_ref).max() > 0
    r.close()


def test_hht_alpha_with_a_slip_weakening_fault():
    """BC_apply under HHT-alpha sees fields%displ / veloc and t_alpha (solver.f90:116-121)"""
    deck = harness.cart_deck(24, 16, ezflt=8, scheme="HHT-alpha", nsteps=300)
    (ed, ev, ea), _, r = _run(deck, 300, synthetic_seed=20261017)
    assert ed <= 1e-10 and ev <= 1e-10 and ea <= 1e-10, (ed, ev, ea)
    fid, i, np_f, _ = r.faults[0]
    st = r.e.fault_state(fid, np_f)
    for k in ("D", "V", "T"):
        assert rel_l2(st[k], r.o.arr(f"bc.{i}.{k}")) <= 1e-10, k
    assert np.abs(st["D"]).max() > 1e-3
    r.close()


@pytest.mark.parametrize("kind,nst", [("symp_PV", 1), ("symp_PFR", 3), ("symp_PEFRL", 4)])
def test_symplectic_schemes(kind, nst):
    text = harness.deck("testsh").replace("courant = 0.3d0", f"courant = 0.3d0, kind='{kind}'")
    (ed, ev, ea), (s_got, s_ref), r = _run(text, 200)
    assert r.o.i("scheme") == 3 and r.o.i("nstages") == nst
    assert ed <= 1e-10 and ev <= 1e-10 and ea <= 1e-10, (ed, ev, ea)
    assert np.abs(s_got - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    r.close()


@pytest.mark.parametrize("deck,mech,block", [
    ("lamb", "EXPLOSION", ""),
    ("lamb", "DOUBLE_COUPLE", "&SRC_DOUBLE_COUPLE dip=60d0 /"),
    ("lamb", "MOMENT", "&SRC_MOMENT Mxx=1d0, Mxz=0.3d0, Mzx=-0.2d0, Mzz=0.5d0 /"),
    ("testsh", "MOMENT", "&SRC_MOMENT Myx=1d0, Myz=0.4d0 /"),
])
def test_moment_sources(deck, mech, block):
    text = harness.deck(deck).replace("mechanism= 'FORCE'", f"mechanism= '{mech}'")
    text = text.replace("&SRC_FORCE angle = 0d0/", block)
    if deck == "testsh":   # a source inside the box: four elements share the node
        text = text.replace("coord= 0.d0,0.d0", "coord= 15.d0,15.d0")
    (ed, ev, ea), (s_got, s_ref), r = _run(text, 300)
    assert r.o.i("src.0.moment") == 1 and r.o.i("src.0.nterms") in (2 * r.ngll, 4 * r.ngll, 8 * r.ngll)
    assert ed <= 1e-10 and ev <= 1e-10 and ea <= 1e-10, (ed, ev, ea)
    assert np.abs(r.o.arr("d")).max() > 0
    r.close()


@pytest.mark.parametrize("x,z", [(1500.0, -1000.0), (1537.0, -963.0), (2000.0, 0.0)])
def test_structured_builder_moment_and_fused_step(x, z):
    """s2d_cart_add_moment builds the terms of SRC_MOMENT_init on the box (a vertex shared by four
    elements, an interior node, a free-surface node); the fused leapfrog step defers the cross of rows
    and columns the source touches"""
    nsteps = 250
    text = harness.deck("lamb").replace("mechanism= 'FORCE'", "mechanism= 'DOUBLE_COUPLE'")
    text = text.replace("&SRC_FORCE angle = 0d0/", "&SRC_DOUBLE_COUPLE dip=35d0 /").replace("1500.d0,-50.d0", f"{x}d0,{z}d0")
    o = orc.Oracle(text, renumber=False)
    e = CartEngine(9, 2, 40, 20, (0.0, 4000.0), (-2000.0, 0.0), rho=2000.0, cp=3200.0, cs=1847.5, dt=0.5e-3)
    for side in (1, 2, 4):
        e.add_abso_side(side, False)
    dip = np.deg2rad(35.0)
    n1, n2, r1, r2 = np.sin(dip), np.cos(dip), -np.cos(dip), np.sin(dip)
    e.add_moment_at(x, z, [2 * r1 * n1, r1 * n2 + r2 * n1, r1 * n2 + r2 * n1, 2 * r2 * n2])
    e.add_receiver_line(2, (2200.0, 0.0), (2700.0, 0.0), "D", 1, nsteps + 1)
    e.commit()
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    e.step(nsteps, tab)
    o.step(nsteps)
    d, v, a = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10 and rel_l2(v, o.arr("v")) <= 1e-10 and rel_l2(a, o.arr("acc")) <= 1e-10
    s_ref = o.seis()[:nsteps + 1]
    assert np.abs(e.seis() - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
    e.close()
    o.close()


def test_periodic_boundary_generic_engine():
    """bc_periodic.f90: Lamb's deck with the left and right sides tied together instead of absorbing;
    the bottom absorbing boundary meets the periodic one at both ends (C merged, bc_abso.f90:226-230)"""
    text = harness.deck("lamb")
    text = text.replace("&BC_DEF  tag = 2 , kind = 'ABSORB' /\n&BC_ABSORB  stacey=F/\n", "&BC_DEF  tags = 4,2 , kind = 'PERIOD' /\n")
    text = text.replace("&BC_DEF  tag = 4 , kind = 'ABSORB' /\n&BC_ABSORB  stacey=F/\n", "")
    assert "PERIOD" in text and text.count("ABSORB") == 2
    (ed, ev, ea), (s_got, s_ref), r = _run(text, 600)
    assert r.o.i("bc.1.kind") == harness.IS_PERIOD or r.o.i("bc.0.kind") == harness.IS_PERIOD
    assert ed <= 1e-10 and ev <= 1e-10 and ea <= 1e-10, (ed, ev, ea)
    d = r.o.arr("d")
    m = r.o.arr("bc.1.master") if r.o.i("bc.1.kind") == harness.IS_PERIOD else r.o.arr("bc.0.master")
    s = r.o.arr("bc.1.slave") if r.o.i("bc.1.kind") == harness.IS_PERIOD else r.o.arr("bc.0.slave")
    assert np.array_equal(d[m - 1], d[s - 1]) and np.abs(d[m - 1]).max() > 0   # the two sides move together
    r.close()


@pytest.mark.parametrize("scheme", ["leapfrog", "newmark"])
def test_periodic_boundary_structured_builder(scheme):
    """s2d_cart_add_periodic + fused step: periodic left/right, absorbing bottom/top, the two-sided fault
    running into the periodic sides (its end weights merged, spec_grid.f90:1000-1005)"""
    nx, nz, ezflt, nsteps, h = 24, 16, 8, 300, 100.0
    o = orc.Oracle(harness.cart_deck(nx, nz, ezflt=ezflt, scheme=scheme, nsteps=nsteps, abso=(1, 3), periodic=(4, 2)),
                   synthetic_seed=20261017, renumber=False)
    e = CartEngine(5, 2, nx, nz, (0.0, nx * h), (0.0, nz * h), ezflt=ezflt, seed=20261017,
                   scheme_kind=0 if scheme == "leapfrog" else 1, courant=0.5)
    e.add_periodic_sides(4, 2)
    fid = e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nx * h / 2, harness.nuc_radius(nx, h), nt_max=nsteps)
    for side in (1, 3):
        e.add_abso_side(side, False)
    e.add_force_at(0.37 * nx * h, 0.61 * nz * h, [o.f("src.0.dir1"), o.f("src.0.dir2")])
    e.commit()
    tab = np.array([[o.stf(0, (k + 1) * o.f("dt"))] for k in range(nsteps)])
    e.step(nsteps, tab)
    o.step(nsteps)
    d, v, a = e.get_fields()
    assert rel_l2(d, o.arr("d")) <= 1e-10 and rel_l2(v, o.arr("v")) <= 1e-10 and rel_l2(a, o.arr("acc")) <= 1e-10
    ibf = [i for i in range(o.i("nbc")) if o.i(f"bc.{i}.kind") == harness.IS_DYNFLT][0]
    st = e.fault_state(fid, o.i(f"bc.{ibf}.np"))
    for k in ("D", "V", "T"):
        assert rel_l2(st[k], o.arr(f"bc.{ibf}.{k}")) <= 1e-10, k
    assert np.abs(st["D"]).max() > 1e-3
    e.close()
    o.close()


def test_host_program_takes_the_new_blocks(tmp_path):
    """sem2dsolve_b200 on a TestSH deck with a moment source and the PFR symplectic scheme"""
    text = harness.deck("testsh").replace("mechanism= 'FORCE'", "mechanism= 'MOMENT'")
    text = text.replace("&SRC_FORCE angle = 0d0/", "&SRC_MOMENT Myx=1d0, Myz=0.4d0 /").replace("coord= 0.d0,0.d0", "coord= 15.d0,15.d0")
    text = text.replace("TotalTime=35.d0, courant = 0.3d0", "NbSteps=300, courant = 0.3d0, kind='symp_PFR'")
    (tmp_path / "Par.inp").write_text(text)
    p = subprocess.run([EXE, "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    u = np.fromfile(tmp_path / "Uy_sem2d.dat", dtype=np.float32).reshape(7, 301).T
    o = orc.Oracle(text, renumber=False)
    o.step(300)
    ref = o.seis()[:, :, 0]
    assert np.abs(ref).max() > 0 and np.abs(u - ref).max() <= 2e-7 * np.abs(ref).max()
    o.close()
