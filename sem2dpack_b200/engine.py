"""Object wrapper over the C-ABI handle.  Method names follow the reference's module entry points
they stand for (SRC/solver.f90 `solve`, SRC/bc_gen.f90 `BC_*`, SRC/receivers.f90 `REC_*`)."""
import ctypes as C

import numpy as np

from . import capi
from .capi import (S2D_ASM_PATCH, DynfltDesc, S2DError, Scheme, _f64, _i32, _pd, _pi, _ptr)


def symplectic_stages(kind):
    """time%a, time%b of the symplectic schemes (SRC/time.f90:248-300)."""
    if kind == "symp_PV":
        return [0.5, 0.5], [1.0]
    if kind == "symp_PFR":
        th = 1.0 / (2.0 - 2.0 ** (1.0 / 3.0))
        return [th / 2, (1 - th) / 2, (1 - th) / 2, th / 2], [th, 1 - 2 * th, th]
    if kind == "symp_PEFRL":
        xi, lam, chi = 0.1786178958448091, -0.2123418310626054, -0.06626458266981849
        return ([xi, chi, 1 - 2 * (chi + xi), chi, xi], [0.5 - lam, lam, lam, 0.5 - lam])
    raise ValueError(kind)


def make_scheme(kind, dt, beta, gamma, alpha, stages=None):
    s = Scheme(kind, dt, beta, gamma, alpha)
    if stages is not None:
        coa, cob = stages
        s.nstages = len(cob)
        for k, x in enumerate(coa):
            s.coa[k] = x
        for k, x in enumerate(cob):
            s.cob[k] = x
    return s


def detect_structured(ngll, ibool, npoin, lower_hint=0):
    """s2d_detect_structured: None if ibool is not a MESH_CART box, else dict(nx, nz, ezflt, ex, ez, gx, gz)"""
    L = capi.lib()
    ibool = _i32(ibool)
    nelem = ibool.size // (ngll * ngll)
    nx, nz, ez = C.c_int32(), C.c_int32(), C.c_int32()
    ex, ezz = np.empty(nelem, np.int32), np.empty(nelem, np.int32)
    gx, gz = np.empty(npoin, np.int32), np.empty(npoin, np.int32)
    rc = L.s2d_detect_structured(ngll, nelem, npoin, _ptr(ibool), lower_hint, C.byref(nx), C.byref(nz), C.byref(ez),
                                 _ptr(ex), _ptr(ezz), _ptr(gx), _ptr(gz))
    if rc != 0:
        raise S2DError(rc, "s2d_detect_structured: invalid argument")
    if nx.value == 0:
        return None
    return dict(nx=nx.value, nz=nz.value, ezflt=ez.value, ex=ex, ez=ezz, gx=gx, gz=gz)


class Engine:
    """One device-resident SEM2DPACK problem (problem_type, SRC/problem_class.f90:19-46)."""

    def __init__(self, ngll, ndof, ibool, hprime, rmass, scheme_kind, dt, beta=0.0, gamma=0.5, alpha=1.0,
                 precision=8, device=-1, _handle=None, stages=None):
        self.L = capi.lib()
        self._keep = []
        if _handle is not None:
            self.h = _handle
            return
        ibool = _i32(ibool)
        rmass = _f64(rmass)
        hprime = _f64(hprime)
        n2 = ngll * ngll
        nelem = ibool.size // n2
        npoin = rmass.size // ndof
        self.ngll, self.ndof, self.nelem, self.npoin = ngll, ndof, nelem, npoin
        self.dt = dt
        sch = make_scheme(scheme_kind, dt, beta, gamma, alpha, stages)
        h = C.c_void_p()
        rc = self.L.s2d_create(C.byref(h), ngll, ndof, nelem, npoin, _ptr(ibool), _ptr(hprime), _ptr(rmass),
                               precision, C.byref(sch), device)
        if rc != 0:
            raise S2DError(rc, self.L.s2d_last_error(None).decode())
        self.h = h

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise S2DError(rc, self.L.s2d_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.s2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration (what init_main hands over, SRC/init.f90:16-131) ---------------------
    def set_elastic(self, nelast, a, elem2set, kd2, beta25d=None):
        """beta25d: matwrk_elast_type%beta (ngll,ngll) per coefficient block of a finite-width 2.5D run, or None"""
        a = _f64(a)
        ncoefsets = a.size // (self.ngll * self.ngll * nelast)
        beta25d = _f64(beta25d)
        if beta25d is not None:
            assert beta25d.size == self.ngll * self.ngll * ncoefsets
        self._ck(self.L.s2d_set_elastic(self.h, nelast, ncoefsets, _ptr(a), _ptr(_i32(elem2set)), _ptr(beta25d),
                                        int(kd2)))

    def set_kv(self, elem_ids, eta):
        elem_ids = _i32(elem_ids)
        self._ck(self.L.s2d_set_kv(self.h, elem_ids.size, _ptr(elem_ids), _ptr(_f64(eta))))

    def set_mass(self, mass):
        self._ck(self.L.s2d_set_mass(self.h, _ptr(_f64(mass))))

    def add_abso(self, node, Cmat, is_flat=True, n=None, stacey=False, bibool=None, K=None):
        node = _i32(node)
        bibool = _i32(bibool)
        nbe = 0 if bibool is None else bibool.size // self.ngll
        self._ck(self.L.s2d_add_abso(self.h, node.size, _ptr(node), _ptr(_f64(Cmat)), int(is_flat), _ptr(_f64(n)),
                                     int(stacey), nbe, _ptr(bibool), _ptr(_f64(K))))

    def add_dirneu(self, node, kind_h, kind_v, B_h=None, B_v=None):
        node = _i32(node)
        self._ck(self.L.s2d_add_dirneu(self.h, node.size, _ptr(node), kind_h, kind_v, _ptr(_f64(B_h)),
                                       _ptr(_f64(B_v))))

    def add_dynflt(self, **kw):
        """Keyword names are the fields of s2d_dynflt_desc."""
        d = DynfltDesc()
        keep = []
        for name, ctype in DynfltDesc._fields_:
            val = kw.get(name)
            if ctype is capi._PD:
                arr = _f64(val)
                keep.append(arr)
                setattr(d, name, _pd(arr))
            elif ctype is capi._PI:
                arr = _i32(val)
                keep.append(arr)
                setattr(d, name, _pi(arr))
            elif val is not None:
                setattr(d, name, val)
        fid = C.c_int32(-1)
        self._ck(self.L.s2d_add_dynflt(self.h, C.byref(d), C.byref(fid)))
        return fid.value

    def add_force(self, iglob, direction):
        sid = C.c_int32(-1)
        self._ck(self.L.s2d_add_force(self.h, int(iglob), _ptr(_f64(direction)), C.byref(sid)))
        return sid.value

    def add_periodic(self, master, slave):
        """bc_periodic_type (SRC/bc_periodic.f90): f(master) += f(slave); f(slave) = f(master)"""
        master, slave = _i32(master), _i32(slave)
        self._ck(self.L.s2d_add_periodic(self.h, master.size, _ptr(master), _ptr(slave)))

    def add_moment(self, node, coef):
        """so_moment_type after SRC_MOMENT_init (SRC/src_moment.f90:129-180): node (nterms) and
        coef (nterms, ndof) column-major, in the order SRC_MOMENT_add applies them."""
        node = _i32(node)
        sid = C.c_int32(-1)
        self._ck(self.L.s2d_add_moment(self.h, node.size, _ptr(node), _ptr(_f64(coef)), C.byref(sid)))
        return sid.value

    def add_receivers(self, field, isamp, nt_rec, iglob=None, einterp=None, interp=None):
        at_node = iglob is not None
        nx = (_i32(iglob) if at_node else _i32(einterp)).size
        self.rec_shape = (self.ndof, nx, nt_rec)
        self._ck(self.L.s2d_add_receivers(self.h, nx, field.encode()[:1], isamp, nt_rec, int(at_node),
                                          _ptr(_i32(iglob)), _ptr(_i32(einterp)), _ptr(_f64(interp))))

    def commit(self, variant=S2D_ASM_PATCH):
        self._ck(self.L.s2d_commit(self.h, variant))

    # -- time loop ----------------------------------------------------------------------------
    def set_fields(self, d=None, v=None, a=None):
        self._ck(self.L.s2d_set_fields(self.h, _ptr(_f64(d)), _ptr(_f64(v)), _ptr(_f64(a))))

    def get_fields(self):
        n = self.npoin * self.ndof
        d, v, a = np.empty(n), np.empty(n), np.empty(n)
        self._ck(self.L.s2d_get_fields(self.h, _ptr(d), _ptr(v), _ptr(a)))
        return d, v, a

    def step(self, nsteps=1, src_ampli=None, bc_ampli=None):
        """nsteps passes of the loop body of SRC/main.f90:51-99 (solve + REC_store + BC_write)."""
        self._ck(self.L.s2d_step(self.h, nsteps, _ptr(_f64(src_ampli)), _ptr(_f64(bc_ampli))))

    def compute_fint(self):
        f = np.empty(self.npoin * self.ndof)
        self._ck(self.L.s2d_compute_fint(self.h, _ptr(f)))
        return f

    @property
    def it(self):
        v = C.c_int32()
        self._ck(self.L.s2d_get_it(self.h, C.byref(v)))
        return v.value

    def seis(self):
        """(nt, nx, ndof) float32 like rec%sis (SRC/receivers.f90:12)."""
        nd, nx, nt = self.rec_shape
        s = np.empty(nd * nx * nt, np.float32)
        self._ck(self.L.s2d_get_seis(self.h, _ptr(s)))
        return s.reshape(nd, nx, nt).transpose(2, 1, 0)

    def seis_row(self, it):
        nd, nx, _ = self.rec_shape
        row = np.empty(nd * nx, np.float32)
        self._ck(self.L.s2d_get_seis_row(self.h, it, _ptr(row)))
        return row.reshape(nd, nx)

    def fault(self, fid, onx):
        nout, ncalls = C.c_int32(), C.c_int32()
        self._ck(self.L.s2d_get_fault(self.h, fid, None, C.byref(nout), None, C.byref(ncalls)))
        rec = np.empty((max(nout.value, 0), 6, onx), np.float32)
        pot = np.empty((max(ncalls.value, 0), 2 * (self.ndof + 1)))
        self._ck(self.L.s2d_get_fault(self.h, fid, _ptr(rec), C.byref(nout), _ptr(pot), C.byref(ncalls)))
        return rec, pot

    def fault_state(self, fid, np_):
        nd = self.ndof
        out = dict(D=np.empty(np_ * nd), V=np.empty(np_ * nd), T=np.empty(np_ * 2), Tstick=np.empty(np_ * 2),
                   MU=np.empty(np_), theta=np.empty(np_), sigma=np.empty(np_))
        self._ck(self.L.s2d_get_fault_state(self.h, fid, *[_ptr(out[k]) for k in
                                                           ("D", "V", "T", "Tstick", "MU", "theta", "sigma")]))
        return out

    def progress(self):
        vm, dm = C.c_double(), C.c_double()
        self._ck(self.L.s2d_progress(self.h, C.byref(vm), C.byref(dm)))
        return vm.value, dm.value

    def energy(self):
        e = C.c_double()
        self._ck(self.L.s2d_energy(self.h, C.byref(e)))
        return e.value

    def energy_w25d(self):
        e = C.c_double()
        self._ck(self.L.s2d_energy_w25d(self.h, C.byref(e)))
        return e.value

    def coloring(self):
        nc = C.c_int32()
        col = np.empty(self.nelem, np.int32)
        self._ck(self.L.s2d_get_coloring(self.h, C.byref(nc), _ptr(col)))
        return nc.value, col

    def time_fint(self, reps):
        ms = C.c_float()
        self._ck(self.L.s2d_time_fint(self.h, reps, C.byref(ms)))
        return ms.value

    def time_steps(self, nsteps):
        ms = C.c_float()
        self._ck(self.L.s2d_time_steps(self.h, nsteps, C.byref(ms)))
        return ms.value

    def route(self):
        """0 = any-mesh kernel, 1 = z-marching strip kernel (s2d_kernel_route)"""
        r = C.c_int32()
        self._ck(self.L.s2d_kernel_route(self.h, C.byref(r)))
        return r.value

    PHASES = ("predictor", "element_force", "halo_fold_exchange", "sources", "boundary_conditions", "node_update",
              "outputs")

    def time_phases(self, nsteps):
        """average ms per step of each phase of a step (s2d_time_phases)"""
        ms = (C.c_float * len(self.PHASES))()
        self._ck(self.L.s2d_time_phases(self.h, nsteps, ms))
        return dict(zip(self.PHASES, [float(x) for x in ms]))

    def kernel_ms(self):
        ms = C.c_float()
        self._ck(self.L.s2d_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.s2d_launch_count(self.h, C.byref(n)))
        return n.value

    def stream(self):
        p = C.c_void_p()
        self._ck(self.L.s2d_stream(self.h, C.byref(p)))
        return p.value


class CartEngine(Engine):
    """A MESH_CART problem (SRC/mesh_cartesian.f90) generated on the device by the structured
    builder; afterwards it is driven exactly like an `Engine`."""

    def __init__(self, ngll, ndof, nx, nz, xlim, zlim, ezflt=0, seed=0, rho=0.0, cp=0.0, cs=0.0, scheme_kind=0,
                 dt=0.0, courant=0.5, beta=0.0, gamma=0.5, alpha=1.0, precision=8, device=-1, ix0=0, iz0=0,
                 halo_left=False, halo_right=False, coef_mode=0, stages=None, renumber=False):
        L = capi.lib()
        d = capi.CartDesc()
        d.ngll, d.ndof, d.nx, d.nz, d.ezflt = ngll, ndof, nx, nz, ezflt
        d.x0, d.x1 = xlim
        d.z0, d.z1 = zlim
        d.seed, d.ix0, d.iz0 = seed, ix0, iz0
        d.rho, d.cp, d.cs = rho, cp, cs
        d.precision = precision
        d.scheme = make_scheme(scheme_kind, dt, beta, gamma, alpha, stages)
        d.courant = courant
        d.device = device
        d.halo_left, d.halo_right = int(halo_left), int(halo_right)
        d.coef_mode = int(coef_mode)  # 0: (lambda, mu) only where the rheology allows, 1: all planes in HBM
        d.renumber = int(renumber)    # OPT_RENUMBER: RCM element order and the node numbering that follows from it
        h = C.c_void_p()
        rc = L.s2d_cart_create(C.byref(h), C.byref(d))
        if rc != 0:
            raise S2DError(rc, "s2d_cart_create failed (see stderr)")
        super().__init__(ngll, ndof, None, None, None, scheme_kind, dt, _handle=h)
        self.ngll, self.ndof = ngll, ndof
        npoin, nelem, dtv = C.c_int64(), C.c_int64(), C.c_double()
        self._ck(self.L.s2d_cart_info(self.h, C.byref(npoin), C.byref(nelem), C.byref(dtv)))
        self.npoin, self.nelem, self.dt = npoin.value, nelem.value, dtv.value
        self.nx, self.nz = nx, nz

    def set_dt(self, dt):
        """time%dt shared by every x-strip of one global mesh (min over strips of the Courant step)."""
        self._ck(self.L.s2d_cart_set_dt(self.h, float(dt)))
        self.dt = float(dt)

    def add_abso_side(self, side, stacey=False):
        self._ck(self.L.s2d_cart_add_abso(self.h, side, int(stacey)))

    def add_fault_swf(self, Dc, MuS, MuD, Tn, Tt, Tt_nuc, x_nuc, half_nuc, oixd=1, oitd=1, nt_max=0):
        fid = C.c_int32(-1)
        self._ck(self.L.s2d_cart_add_fault_swf(self.h, Dc, MuS, MuD, Tn, Tt, Tt_nuc, x_nuc, half_nuc, oixd, oitd,
                                               nt_max, C.byref(fid)))
        return fid.value

    def add_force_at(self, x, z, direction):
        sid = C.c_int32(-1)
        self._ck(self.L.s2d_cart_add_force(self.h, x, z, _ptr(_f64(direction)), C.byref(sid)))
        return sid.value

    def add_periodic_sides(self, master_tag, slave_tag):
        self._ck(self.L.s2d_cart_add_periodic(self.h, master_tag, slave_tag))

    def add_moment_at(self, x, z, M):
        """M(2,ndof) column-major as so%M (SRC/src_moment.f90:44-100)"""
        sid = C.c_int32(-1)
        self._ck(self.L.s2d_cart_add_moment(self.h, x, z, _ptr(_f64(M)), C.byref(sid)))
        return sid.value

    def add_receiver_line(self, nx, first, last, field, isamp, nt_rec):
        self.rec_shape = (self.ndof, nx, nt_rec)  # stations must snap to distinct nodes
        self._ck(self.L.s2d_cart_add_receivers(self.h, nx, first[0], first[1], last[0], last[1],
                                               field.encode()[:1], isamp, nt_rec))

    def set_material(self, rho, cp, cs):
        """rho, cp, cs (nelem, ngll, ngll) [e][j][i] at the GLL points, natural element order (s2d_cart_set_material)"""
        self._ck(self.L.s2d_cart_set_material(self.h, _ptr(_f64(rho)), _ptr(_f64(cp)), _ptr(_f64(cs))))
        npoin, nelem, dtv = C.c_int64(), C.c_int64(), C.c_double()
        self._ck(self.L.s2d_cart_info(self.h, C.byref(npoin), C.byref(nelem), C.byref(dtv)))
        self.dt = dtv.value

    def set_kv_elems(self, elem_ids, eta):
        """eta (nkv, ngll, ngll) of the Kelvin-Voigt elements, ids 1-based in natural order (s2d_cart_set_kv_elems)"""
        elem_ids = _i32(elem_ids)
        self._ck(self.L.s2d_cart_set_kv_elems(self.h, elem_ids.size, _ptr(elem_ids), _ptr(_f64(eta))))

    def set_plastic(self, par, elem_set):
        """par (nsets, 6) = coh, phi [deg], Tv, e0(3) per plastic material; elem_set (nelem) natural order, 0 = elastic
        (s2d_cart_set_plastic)"""
        par = _f64(par).reshape(-1, 6)
        elem_set = _i32(elem_set)
        self._ck(self.L.s2d_cart_set_plastic(self.h, par.shape[0], _ptr(par), _ptr(elem_set)))

    def set_visco(self, nbody, moduli, wbody, theta, elem_set):
        """per visco-elastic material: Nbody, (lambda_inf, mu_inf), wbody (<= 8), theta (Nbody, 3) as get_attenuation
        returns them; elem_set (nelem) natural order, 0 = elastic (s2d_cart_set_visco)"""
        nbody = _i32(nbody)
        ns = nbody.size
        wb = np.zeros((ns, 8))
        th = np.zeros((ns, 3, 8))
        for k in range(ns):
            nb = int(nbody[k])
            wb[k, :nb] = np.asarray(wbody[k])[:nb]
            th[k, :, :nb] = np.asarray(theta[k]).reshape(nb, 3).T
        mod = _f64(moduli).reshape(ns, 2)
        elem_set = _i32(elem_set)
        self._ck(self.L.s2d_cart_set_visco(self.h, ns, _ptr(nbody), _ptr(mod), _ptr(wb), _ptr(th), _ptr(elem_set)))

    def set_damage(self, par, elem_set):
        """par (nsets, 13) = lambda, mu, phi [deg], alpha0, Cd, beta, R, e0(3), ep0(3) per damage material; elem_set
        (nelem) natural order, 0 = elastic (s2d_cart_set_damage)"""
        par = _f64(par).reshape(-1, 13)
        elem_set = _i32(elem_set)
        self._ck(self.L.s2d_cart_set_damage(self.h, par.shape[0], _ptr(par), _ptr(elem_set)))

    def damage_state(self):
        """(nelem, 4, ngll, ngll): alpha, ep11, ep22, ep12, natural element order (s2d_cart_get_damage_state)"""
        out = np.empty((self.nelem, 4, self.ngll, self.ngll))
        self._ck(self.L.s2d_cart_get_damage_state(self.h, _ptr(out)))
        return out

    def plastic_strain(self):
        """ep (nelem, 3, ngll, ngll), natural element order (s2d_cart_get_plastic_strain)"""
        out = np.empty((self.nelem, 3, self.ngll, self.ngll))
        self._ck(self.L.s2d_cart_get_plastic_strain(self.h, _ptr(out)))
        return out

    def set_w25d(self, W):
        """&GENERAL W: finite seismogenic width (s2d_cart_set_w25d)"""
        self._ck(self.L.s2d_cart_set_w25d(self.h, float(W)))

    def fill_fields(self, seed, amp_d, amp_v):
        """seeded non-trivial state (hash noise keyed by global lattice coordinates); accel = 0"""
        self._ck(self.L.s2d_cart_fill_fields(self.h, int(seed), float(amp_d), float(amp_v)))

    def get_window(self, gx0, gz0, nwx, nwz, a=False):
        """(d, v[, a]) on a window of the GLL lattice, each (ndof, nwz, nwx); a split fault row is two rows"""
        shp = (self.ndof, nwz, nwx)
        d, v = np.empty(shp), np.empty(shp)
        aa = np.empty(shp) if a else None
        self._ck(self.L.s2d_cart_get_window(self.h, gx0, gz0, nwx, nwz, _ptr(d), _ptr(v), _ptr(aa)))
        return (d, v, aa) if a else (d, v)

    def snapshot_elem(self, what):
        """PLOT_FIELD's element-wise field 'E' | 'S' | 'd' | 'c': float32 (ncomp, nelem, ngll, ngll)"""
        ncomp = self.ndof + 1 if what in "ES" else 1
        out = np.empty((ncomp, self.nelem, self.ngll, self.ngll), np.float32)
        self._ck(self.L.s2d_cart_snapshot_elem(self.h, what.encode()[:1], _ptr(out)))
        return out

    def get_gll(self):
        """(xgll, wgll, hprime[column-major flat]) of the builder (s2d_cart_get_gll)"""
        n = self.ngll
        x, w, H = np.empty(n), np.empty(n), np.empty(n * n)
        self._ck(self.L.s2d_cart_get_gll(self.h, _ptr(x), _ptr(w), _ptr(H)))
        return x, w, H

    def get_tables(self, ibool=True, a=False, rmass=True, coord=False, nelast=None):
        n2 = self.ngll * self.ngll
        ib = np.empty(self.nelem * n2, np.int32) if ibool else None
        nel = nelast or (2 if self.ndof == 1 else 6)
        aa = np.empty(self.nelem * n2 * nel) if a else None
        rm = np.empty(self.npoin * self.ndof) if rmass else None
        co = np.empty(self.npoin * 2) if coord else None
        self._ck(self.L.s2d_cart_get(self.h, _ptr(ib), _ptr(aa), _ptr(rm), _ptr(co)))
        return ib, aa, rm, co
