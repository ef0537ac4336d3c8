"""CPU checks of the windowed-parity helper (tests/window.py): its restatement of the hash, and that an oracle
window reproduces the oracle of the whole mesh on the nodes whose domain of dependence stays inside."""
import numpy as np

import harness
import orc
import window


def test_hash_restatement_matches_the_oracle():
    L = orc.lib()
    rng = np.random.default_rng(0)
    ix = rng.integers(0, 40000, 64)
    iz = rng.integers(0, 40000, 64)
    for k in (1, 2, 3, 16, 17, 32, 33):
        got = window.hash_u(20261017, ix, iz, k)
        ref = np.array([L.orc_hash_u(20261017, int(a), int(b), k) for a, b in zip(ix, iz)])
        assert np.array_equal(got, ref)


def test_oracle_window_equals_oracle_whole_mesh_inside_the_domain_of_dependence():
    NX, NZ, ez, k, seed = 30, 24, 12, 3, 20261017
    fill = (5, 1e-3, 1.0)
    whole = window.Window(0, 0, NX, NZ, NX, NZ, ez, 1.0e-3, k, seed, fill, src=(NX * 50.0 + 130.0, ez * 100.0 + 210.0))
    win = window.Window(6, 4, 18, 16, NX, NZ, ez, 1.0e-3, k, seed, fill, src=(NX * 50.0 + 130.0, ez * 100.0 + 210.0))
    whole.o.step(k)
    win.o.step(k)
    m = k * 4
    npw, npg = win.o.i("npoin"), whole.o.i("npoin")
    sel_w = win.lat[m:win.LZw - m, m:win.LXw - m]
    sel_g = whole.lat[win.gz0 + m:win.gz0 + win.LZw - m, win.gx0 + m:win.gx0 + win.LXw - m]
    for name in ("d", "v"):
        a, b = win.o.arr(name), whole.o.arr(name)
        for c in range(2):
            assert harness.rel_l2(a[c * npw + sel_w], b[c * npg + sel_g]) <= 1e-13, (name, c)
    # ... and differs outside it (the window's artificial free edges have been felt)
    a, b = win.o.arr("v"), whole.o.arr("v")
    assert harness.rel_l2(a[win.lat[:, 0]], b[whole.lat[win.gz0:win.gz0 + win.LZw, win.gx0]]) > 1e-8
    whole.close()
    win.close()
