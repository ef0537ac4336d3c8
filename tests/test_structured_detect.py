"""Host logic of the kernel routing (include/sem2d_b200.h `s2d_detect_structured`): the box is recognised from the
topology of ibool alone, in the RCM element order the reference uses by default (OPT_RENUMBER,
SRC/constants.f90:11; mesh_structured.f90:204-269) as well as in natural order, with and without the split-node
row of `ezflt`.  No device needed."""
import numpy as np
import pytest

import harness
import orc
from sem2dpack_b200.engine import detect_structured


def _oracle(nx, nz, ngll, ezflt, renumber):
    return orc.Oracle(harness.cart_deck(nx, nz, ngll=ngll, ezflt=ezflt, nrec=0, src=False, abso=(), fault=None),
                      renumber=renumber)


@pytest.mark.parametrize("renumber", [True, False])
@pytest.mark.parametrize("ngll,nx,nz,ezflt", [(5, 19, 13, 0), (5, 17, 16, 5), (6, 11, 9, 4), (3, 30, 7, 3), (9, 4, 5, 0),
                                             (5, 1, 6, 2), (4, 7, 1, 0)])
def test_box_recognised_in_any_element_order(ngll, nx, nz, ezflt, renumber):
    o = _oracle(nx, nz, ngll, ezflt, renumber)
    ib, npoin = o.arr("ibool"), o.i("npoin")
    hint = 0
    if ezflt:   # a node of the lower side of the split row: the element just below the fault, in the caller's order
        perm = o.arr("perm")            # perm(new) = old (1-based, fem_grid.f90:505-519), index 0 unused
        old = (ezflt - 1) * nx + 1
        new = int(np.nonzero(perm[1:] == old)[0][0])
        hint = int(ib.reshape(-1, ngll, ngll)[new, ngll - 1, 0])
    box = detect_structured(ngll, ib, npoin, hint)
    assert box is not None
    assert (box["nx"], box["nz"], box["ezflt"]) == (nx, nz, ezflt)
    # element positions: the natural (pre-RCM) index of element e is perm(e)
    perm = o.arr("perm")[1:] - 1
    assert np.array_equal(box["ex"], perm % nx) and np.array_equal(box["ez"], perm // nx)
    # node positions against the coordinates
    co = o.arr("coord").reshape(-1, 2)
    h = 100.0
    xg, _, _ = orc.gll(ngll)
    ex, i = np.divmod(box["gx"], ngll - 1)
    on_edge = ex == nx
    ex, i = np.where(on_edge, nx - 1, ex), np.where(on_edge, ngll - 1, i)
    assert np.abs(co[:, 0] - h * (ex + 0.5 * (xg[i] + 1.0))).max() < 1e-9
    gzg = box["gz"] - ((box["gz"] >= ezflt * (ngll - 1) + 1) if ezflt else 0)
    ez, j = np.divmod(gzg, ngll - 1)
    top = ez == nz
    ez, j = np.where(top, nz - 1, ez), np.where(top, ngll - 1, j)
    assert np.abs(co[:, 1] - h * (ez + 0.5 * (xg[j] + 1.0))).max() < 1e-9
    # every lattice position is taken exactly once
    LX = nx * (ngll - 1) + 1
    assert len(np.unique(box["gz"].astype(np.int64) * LX + box["gx"])) == npoin
    o.close()


def test_not_a_box():
    o = _oracle(6, 5, 5, 0, True)
    ib, npoin = o.arr("ibool").reshape(-1, 5, 5).copy(), o.i("npoin")
    # one element turned by 90 degrees: same nodes, different orientation
    turned = ib.copy()
    turned[7] = np.rot90(ib[7])
    assert detect_structured(5, turned, npoin) is None
    # a hole: drop one interior element (its private nodes disappear from the table)
    natural = _oracle(6, 5, 5, 0, False)
    ibn = natural.arr("ibool").reshape(-1, 5, 5)
    holed = np.delete(ibn, 14, axis=0)
    _, inv = np.unique(holed, return_inverse=True)
    assert detect_structured(5, (inv + 1).astype(np.int32), int(inv.max()) + 1) is None
    o.close()
    natural.close()


@pytest.mark.parametrize("nx,nz", [(1, 1), (1, 7), (9, 1), (2, 2), (19, 13), (40, 24), (270, 90), (7, 64), (60, 60), (100, 3)])
def test_rcm_order_equals_the_oracles_restatement_of_genrcm(nx, nz):
    """s2d_rcm_box (csrc/rcm_box.hpp) against oracle/rcm.hpp, two independent restatements of SRC/rcm.f90 +
    mesh_structured.f90:204-269; integer data, bit-exact"""
    import ctypes as C
    from sem2dpack_b200 import capi
    perm = np.empty(nx * nz, np.int32)
    assert capi.lib().s2d_rcm_box(nx, nz, perm.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(perm, orc.rcm(nx, nz))
    assert np.array_equal(np.sort(perm), np.arange(1, nx * nz + 1))
