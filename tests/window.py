"""Windowed parity check of a large synthetic mesh against the CPU oracle (test infrastructure).

The oracle cannot hold an 8192 x 8192 mesh, but it does not have to: after k steps the value at a node
depends only on the state within k elements of it.  So the oracle is run on a WINDOW of the global mesh
-- same elements (the hash medium is a function of global lattice coordinates), same initial state
(`s2d_cart_fill_fields`, restated here), same fault / absorbing sides / source where the window contains
them -- and compared with the engine on the nodes whose domain of dependence stays inside the window.
"""
import numpy as np

import harness
import orc

H = 100.0
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _sm64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
    return x ^ (x >> np.uint64(31))


def hash_u(seed, ix, iz, k):
    """the counter-based hash of the synthetic medium (oracle hash_u / cart.cu hash_u), vectorised"""
    with np.errstate(over="ignore"):
        ix = np.asarray(ix, np.uint64)
        iz = np.asarray(iz, np.uint64)
        a = _sm64(ix * np.uint64(0x9E3779B97F4A7C15) + np.uint64(k))
        b = _sm64(iz * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(0x165667B19E3779F9) * np.uint64(k + 1))
        h = _sm64(np.uint64(seed) ^ a ^ b)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0


def fmt(x):
    return f"{x:.17e}".replace("e", "d")


def window_deck(X0, Z0, nxw, nzw, NX, NZ, ezflt_g, dt, nsteps, ngll=5, scheme="leapfrog", src=None, half_nuc=1500.0):
    """Par.inp of the window [X0, X0+nxw) x [Z0, Z0+nzw) (elements) of the NX x NZ synthetic benchmark mesh"""
    ez = ezflt_g - Z0 if (ezflt_g and Z0 < ezflt_g < Z0 + nzw) else 0
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof=2, title='window', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim={fmt(X0 * H)},{fmt((X0 + nxw) * H)}, zlim={fmt(Z0 * H)},{fmt((Z0 + nzw) * H)}, nelem={nxw},{nzw}"
         + (f", ezflt={ez}" if ez else "") + " /",
         "&MATERIAL tag=1, kind='ELAST' /",
         "&MAT_ELASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0 /"]
    if ez:
        L += ["&BC_DEF tags=5,6, kind='DYNFLT' /",
              "&BC_DYNFLT friction='SWF', Tn=-120.d6, TtH='PWCONR' /",
              "&DIST_PWCONR num=2, ref=%s,%s /" % (fmt(NX * H / 2), fmt(ezflt_g * H)),
              "     %s" % fmt(half_nuc),
              "81.6d6 70.d6",
              "&BC_DYNFLT_SWF Dc=0.4d0, MuS=0.677d0, MuD=0.525d0 /"]
    sides = []
    if Z0 == 0:
        sides.append(1)
    if X0 + nxw == NX:
        sides.append(2)
    if Z0 + nzw == NZ:
        sides.append(3)
    if X0 == 0:
        sides.append(4)
    for t in sides:
        L += [f"&BC_DEF tag={t}, kind='ABSORB' /", "&BC_ABSORB stacey=F /"]
    L += [f"&TIME NbSteps={nsteps}, dt={fmt(dt)}, kind='{scheme}' /"]
    if src is not None:
        L += [f"&SRC_DEF stf='RICKER', coord={fmt(src[0])},{fmt(src[1])}, mechanism='FORCE' /",
              "&STF_RICKER f0=2.d0, onset=0.05d0, ampli=1.d9 /",
              "&SRC_FORCE angle=30d0 /"]
    return "\n".join(L) + "\n", ez, sides


class Window:
    """the oracle on one window, its node -> lattice map, and the comparison with an engine"""

    def __init__(self, X0, Z0, nxw, nzw, NX, NZ, ezflt_g, dt, nsteps, seed, fill, ngll=5, scheme="leapfrog", src=None,
                 ix0=0, half_nuc=1500.0):
        """fill = (seed, amp_d, amp_v) of s2d_cart_fill_fields; ix0 = lattice column of the engine's first column
        inside the global mesh (x-strips)"""
        self.X0, self.Z0, self.nxw, self.nzw, self.N = X0, Z0, nxw, nzw, ngll
        deck, ez, sides = window_deck(X0, Z0, nxw, nzw, NX, NZ, ezflt_g, dt, nsteps, ngll, scheme, src, half_nuc)
        self.ez, self.sides, self.nsteps = ez, sides, nsteps
        N1 = ngll - 1
        self.o = o = orc.Oracle(deck, synthetic_seed=seed, renumber=False, lattice_origin=(X0 * N1, Z0 * N1))
        assert abs(o.f("dt") - dt) <= 1e-15 * dt
        npoin = o.i("npoin")
        ib = o.arr("ibool").reshape(nxw * nzw, ngll, ngll)      # [e][j][i], natural element order
        self.LXw = nxw * N1 + 1
        self.LZw = nzw * N1 + 1 + (1 if ez else 0)
        lat = np.zeros((self.LZw, self.LXw), np.int64)          # oracle node id (1-based) at every lattice point
        for iz in range(nzw):
            r0 = iz * N1 + (1 if (ez and iz >= ez) else 0)
            for ix in range(nxw):
                lat[r0:r0 + ngll, ix * N1:ix * N1 + ngll] = ib[iz * nxw + ix]
        assert lat.min() >= 1 and len(np.unique(lat)) == npoin == lat.size
        self.lat = lat - 1
        # the engine's lattice rows of this window (the global mesh duplicates its fault row too)
        self.gx0 = X0 * N1 - ix0
        self.gz0 = Z0 * N1 + (1 if (ezflt_g and Z0 >= ezflt_g) else 0)
        # initial state: the restatement of k_cart_fill (cart.cu)
        fseed, amp_d, amp_v = fill
        gz = np.arange(self.LZw)
        gzg = gz - ((gz >= ez * N1 + 1) if ez else 0)
        Xg = (X0 * N1 + np.arange(self.LXw))[None, :].repeat(self.LZw, 0)
        Zg = (Z0 * N1 + gzg)[:, None].repeat(self.LXw, 1)
        d0 = np.zeros(npoin * 2)
        v0 = np.zeros(npoin * 2)
        for c in range(2):
            d0[c * npoin + self.lat] = amp_d * hash_u(fseed, Xg, Zg, 16 + c)
            v0[c * npoin + self.lat] = amp_v * hash_u(fseed, Xg, Zg, 32 + c)
        o.set_fields(d0, v0)

    def stf_table(self, nsteps):
        dt = self.o.f("dt")
        return np.array([[self.o.stf(0, (k + 1) * dt)] for k in range(nsteps)])

    def compare(self, e, tol=1e-10):
        """steps the oracle, fetches the engine's window, compares the nodes that cannot have seen the window's
        artificial edges; returns the relative L2 errors (d, v) and the number of nodes compared"""
        o, N1, k = self.o, self.N - 1, self.nsteps
        o.step(k)
        d, v = e.get_window(self.gx0, self.gz0, self.LXw, self.LZw)
        npoin = o.i("npoin")
        m = k * N1
        x_lo = 0 if 4 in self.sides else m
        x_hi = self.LXw if 2 in self.sides else self.LXw - m
        z_lo = 0 if 1 in self.sides else m
        z_hi = self.LZw if 3 in self.sides else self.LZw - m
        assert x_hi - x_lo > 2 * N1 and z_hi - z_lo > 2 * N1, "window too small for this many steps"
        sel = self.lat[z_lo:z_hi, x_lo:x_hi]
        od, ov = o.arr("d"), o.arr("v")
        errs = []
        for got, ref in ((d, od), (v, ov)):
            g = np.stack([got[c, z_lo:z_hi, x_lo:x_hi] for c in range(2)])
            r = np.stack([ref[c * npoin + sel] for c in range(2)])
            errs.append(harness.rel_l2(g, r))
        return errs[0], errs[1], sel.size

    def close(self):
        self.o.close()
