"""CPU checks of the C++ host program (host/): Par.inp reading and the reference's error behaviour.
Everything that reaches the device is covered by tests/test_gpu_host_program.py."""
import os
import subprocess

import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sem2dpack_b200", "lib", "sem2dsolve_b200")


def run(tmp_path, text):
    (tmp_path / "Par.inp").write_text(text)
    return subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True, timeout=120)


def test_refusals_happen_while_reading_the_deck(tmp_path):
    assert os.path.exists(EXE), "host program missing: run __graft_entry__.build()"
    cases = [
        (harness.deck("tpv3").replace("'ELAST' ,'KV'", "'ELAST' ,'PLAST'"), "MAT_read"),
        (harness.deck("tpv3").replace("etaH='GAUSSIAN'", "etaH='LINEAR'"), "DIST_read"),
        (harness.deck("ratestate").replace("friction='RSF'", "friction='XYZ'"), "invalid friction"),
        (harness.deck("ratestate").replace("TtH='ORDER0'", "TtH='SPLINE'"), "DIST_read"),
        (harness.deck("testsh").replace("courant = 0.3d0", "courant = 0.9d0"), "Courant out of range [0,0.6]"),
        (harness.deck("testsh").replace("'RICKER'", "'BUTTERWORTH'"), "BUTTER_read: not implemented"),   # as the reference
        (harness.deck("testsh").replace("'RICKER'", "'SPIKE'"), "STF_read"),
        (harness.deck("testsh").replace("'RICKER'", "'TAB'"), "STF_TAB_read"),                              # no stf.tab here
        (harness.deck("testsh").replace("kind = 'ABSORB'", "kind = 'PERIOD'", 1), "BC_read"),
        (harness.deck("lamb").replace("TotalTime=1.5d0, Dt=0.5d-3", "TotalTime=1.5d0, NbSteps=10"), "bad combination"),
        ("&GENERAL iexec=1 /\n", "MESH_DEF input block not found"),
    ]
    for text, needle in cases:
        p = run(tmp_path, text)
        assert p.returncode == 1 and "FATAL ERROR" in p.stdout and needle in p.stdout, (needle, p.stdout)


def test_a_good_deck_gets_as_far_as_the_device(tmp_path):
    """without a GPU the only possible failure of a supported deck is the missing device: there is no
    CPU fallback behind the host program"""
    for name in ("lamb", "testsh", "ratestate", "tpv3"):
        p = run(tmp_path, harness.deck(name))
        if p.returncode != 0:
            assert "no CUDA device" in p.stdout, (name, p.stdout)
