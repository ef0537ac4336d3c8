"""Test harness: plays the role of the reference's Fortran init (SRC/init.f90:16-131).  The oracle
builds the init-time arrays (ibool, hprime, rmass, coefficient planes, boundary tables ...) exactly
as the Fortran modules would, and this module hands them to the product's C-ABI in the order the
ISO_C_BINDING shim would.  The oracle is only the data source / checker here."""
import os

import numpy as np

import orc
from sem2dpack_b200 import Engine
from sem2dpack_b200.capi import S2D_ASM_PATCH

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IS_ABSORB, IS_DIRNEU, IS_DYNFLT, IS_PERIOD = 3, 1, 6, 4


def deck(name):
    with open(os.path.join(GOLDEN, name + ".par")) as f:
        return f.read()


def refdata():
    return np.load(os.path.join(GOLDEN, "refdata.npz"))


def nuc_radius(nx, h=100.0):
    """nucleation half-width: deliberately not a multiple of the GLL spacing, so that no node sits on
    the patch edge where a last-bit difference in its coordinate would flip the `<=` test"""
    return max(3 * h, 0.1 * nx * h) + 0.37 * h


def cart_deck(nx, nz, ngll=5, ndof=2, ezflt=0, scheme="leapfrog", courant=0.5, nsteps=50, h=100.0, fault="swf",
              abso=(1, 2, 3, 4), stacey=False, nrec=8, src=True, periodic=None):
    """A MESH_CART deck of the synthetic benchmark family at test size."""
    L = [f"&GENERAL iexec=1, ngll={ngll}, fmax=3.d0, ndof={ndof}, title='synthetic', verbose='0000', ItInfo=1000 /",
         "&MESH_DEF method='CARTESIAN' /",
         f"&MESH_CART xlim=0d0,{nx*h}d0, zlim=0d0,{nz*h}d0, nelem={nx},{nz}" + (f", ezflt={ezflt}" if ezflt else "") + " /",
         "&MATERIAL tag=1, kind='ELAST' /",
         "&MAT_ELASTIC rho=2670.d0, cp=6000.d0, cs=3464.d0 /"]
    if periodic:
        L += [f"&BC_DEF tags={periodic[0]},{periodic[1]}, kind='PERIOD' /"]
    if ezflt and fault:
        L += ["&BC_DEF tags=5,6, kind='DYNFLT' /"]
        if fault == "swf":
            L += ["&BC_DYNFLT friction='SWF', Tn=-120.d6, TtH='PWCONR' /",
                  "&DIST_PWCONR num=2, ref=%gd0,%gd0 /" % (nx * h / 2, ezflt * h),
                  "     %gd0" % nuc_radius(nx, h),
                  "81.6d6 70.d6",
                  "&BC_DYNFLT_SWF Dc=0.4d0, MuS=0.677d0, MuD=0.525d0 /"]
    for t in abso:
        L += [f"&BC_DEF tag={t}, kind='ABSORB' /", f"&BC_ABSORB stacey={'T' if stacey else 'F'} /"]
    L += [f"&TIME NbSteps={nsteps}, courant={courant}d0, kind='{scheme}' /"]
    if src:
        L += [f"&SRC_DEF stf='RICKER', coord={0.37*nx*h}d0,{0.61*nz*h}d0, mechanism='FORCE' /",
              "&STF_RICKER f0=2.d0, onset=0.05d0, ampli=1.d9 /",
              "&SRC_FORCE angle=30d0 /"]
    if nrec:
        L += [f"&REC_LINE number={nrec}, field='V', first={0.1*nx*h}d0,{0.3*nz*h}d0, last={0.9*nx*h}d0,{0.8*nz*h}d0, isamp=1 /"]
    return "\n".join(L) + "\n"


class Rig:
    """An Engine configured from an Oracle, plus what is needed to drive it in lock step."""

    def __init__(self, o, precision=8, variant=S2D_ASM_PATCH, kd2=None, device=-1):
        self.o = o
        ngll, ndof = o.i("ngll"), o.i("ndof")
        self.ngll, self.ndof = ngll, ndof
        self.dt = o.f("dt")
        self.nt = o.i("nt")
        kind = o.i("scheme")   # 0 leapfrog, 1 newmark, 2 HHT-alpha, 3 symplectic
        assert kind in (0, 1, 2, 3)
        self.kind = kind
        self.alpha = o.f("alpha")
        self.stages = (list(o.arr("time.a")), list(o.arr("time.b"))) if kind == 3 else None
        e = Engine(ngll, ndof, o.arr("ibool"), o.arr("H"), o.arr("rmass"), kind, self.dt, o.f("beta"), o.f("gamma"),
                   o.f("alpha"), precision=precision, device=device, stages=self.stages)
        self.e = e
        if kd2 is None:
            kd2 = (ngll == 5)  # OPT_NGLL (SRC/constants.f90:6, mat_elastic.f90:412)
        beta = o.arr("beta25d")   # finite seismogenic width W (2.5D): matwrk_elast_type%beta per coefficient block
        e.set_elastic(o.i("nelast"), o.arr("a"), o.arr("elem2set"), kd2, beta25d=beta if beta.size else None)
        if o.i("nkv") > 0:
            e.set_kv(o.arr("kv_elem"), o.arr("kv_eta"))
        self.faults = []
        for i in range(o.i("nbc")):
            k = o.i(f"bc.{i}.kind")
            p = f"bc.{i}."
            if k == IS_PERIOD:
                e.add_periodic(o.arr(p + "master"), o.arr(p + "slave"))
            elif k == IS_ABSORB:
                st = bool(o.i(p + "stacey"))
                e.add_abso(o.arr(p + "node"), o.arr(p + "C"), is_flat=bool(o.i(p + "is_flat")), n=o.arr(p + "n"),
                           stacey=st, bibool=o.arr(p + "bibool") if st else None, K=o.arr(p + "K") if st else None)
            elif k == IS_DIRNEU:
                e.add_dirneu(o.arr(p + "node"), o.i(p + "kind_h"), o.i(p + "kind_v"))
            elif k == IS_DYNFLT:
                two = bool(o.i(p + "two_sides"))
                kw = dict(np=o.i(p + "np"), node1=o.arr(p + "node1"), node2=o.arr(p + "node2") if two else None,
                          n1=o.arr(p + "n1"), B=o.arr(p + "B"), invM1=o.arr(p + "invM1"),
                          invM2=o.arr(p + "invM2") if two else None, Z=o.arr(p + "Z"), T0=o.arr(p + "T0"),
                          cohesion=o.arr(p + "cohesion"), coord=o.arr(p + "coord"), V0=o.arr(p + "V"),
                          CoefA2V=o.f(p + "CoefA2V"), CoefA2D=o.f(p + "CoefA2D"),
                          allow_opening=o.i(p + "allow_opening"),
                          normal_kind=o.i(p + "normal.kind"), normal_T=o.f(p + "normal.T"),
                          normal_L=o.f(p + "normal.L"), normal_V=o.f(p + "normal.V"),
                          oix1=o.i(p + "oix1"), oixn=o.i(p + "oixn"), oixd=o.i(p + "oixd"), oit=o.i(p + "oit0"),
                          oitd=o.i(p + "oitd"), nt_max=max(self.nt, 1))
                if o.i(p + "has_swf"):
                    kw.update(swf_kind=o.i(p + "swf.kind"), swf_healing=o.i(p + "swf.healing"))
                    for q in ("dc", "mus", "mud", "p", "alpha", "theta"):
                        kw["swf_" + q] = o.arr(p + "swf." + q)
                if o.i(p + "has_rsf"):
                    kw.update(rsf_kind=o.i(p + "rsf.kind"))
                    for q in ("dc", "mus", "a", "b", "Vstar", "theta", "Vc"):
                        kw["rsf_" + q] = o.arr(p + "rsf." + q)
                if o.i(p + "has_twf"):
                    kw.update(twf_kind=o.i(p + "twf.kind"))
                    for q in ("X", "Z", "mus", "mud", "mu0", "L", "V", "T", "Dc"):
                        kw["twf_" + q] = o.f(p + "twf." + q)
                fid = e.add_dynflt(**kw)
                self.faults.append((fid, i, o.i(p + "np"), o.i(p + "onx")))
        self.nsrc = o.i("nsrc")
        for s in range(self.nsrc):
            if o.i(f"src.{s}.moment"):   # what SRC_MOMENT_init built (src_moment.f90:129-180)
                e.add_moment(o.arr(f"src.{s}.mnode"), o.arr(f"src.{s}.mcoef"))
            else:
                e.add_force(o.i(f"src.{s}.iglob"), [o.f(f"src.{s}.dir1"), o.f(f"src.{s}.dir2")])
        if o.i("rec.present"):
            fld = chr(o.i("rec.field"))
            if o.i("rec.atnode"):
                e.add_receivers(fld, o.i("rec.isamp"), o.i("rec.nt"), iglob=o.arr("rec.iglob"))
            else:
                e.add_receivers(fld, o.i("rec.isamp"), o.i("rec.nt"), einterp=o.arr("rec.einterp"),
                                interp=o.arr("rec.interp"))
        e.commit(variant)

    def stf_table(self, it_first, nsteps):
        """STF_get(t-tdelay)*ampli evaluated by the host per step (src_gen.f90:300-303)."""
        if self.nsrc == 0:
            return None
        if self.kind == 3:   # one row per stage: t = (it-1)*dt + dt*sum(a(1:k)) (solver.f90:186-191)
            coa, cob = self.stages
            nst = len(cob)
            tab = np.empty((nsteps * nst, self.nsrc))
            for k in range(nsteps):
                t = (it_first + k) * self.dt - self.dt
                for q in range(nst):
                    t = t + self.dt * coa[q]
                    for s in range(self.nsrc):
                        tab[k * nst + q, s] = self.o.stf(s, t)
            return tab
        shift = (self.alpha - 1.0) * self.dt if self.kind == 2 else 0.0   # t_alpha (solver.f90:116)
        tab = np.empty((nsteps, self.nsrc))
        for k in range(nsteps):
            t = (it_first + k) * self.dt + shift
            for s in range(self.nsrc):
                tab[k, s] = self.o.stf(s, t)
        return tab

    def step(self, nsteps):
        it0 = self.e.it + 1
        self.e.step(nsteps, self.stf_table(it0, nsteps))

    def close(self):
        self.e.close()
        self.o.close()


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    if den == 0.0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / den)
