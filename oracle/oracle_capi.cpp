// ORACLE -- TEST INFRASTRUCTURE ONLY (see sem2d_oracle.hpp).  C entry points for ctypes so that
// tests/, smoke() and bench.py's cpu_baseline leg can (a) run the CPU restatement and (b) fetch the
// init-time arrays (ibool, H, rmass, coefficient planes, boundary tables ...) that the reference's
// Fortran init would hand to the product's C-ABI.
#include <cstring>
#include <string>
#include <chrono>

#include "sem2d_oracle.hpp"

using namespace orc;

struct OrcHandle {
  Problem pb;
  CartSpec cart;
  std::vector<double> scratch_d;
  std::vector<int> scratch_i;
  std::vector<float> scratch_f;
};

static void set_err(char* err, int errlen, const std::string& m) {
  if (err && errlen > 0) {
    std::strncpy(err, m.c_str(), errlen - 1);
    err[errlen - 1] = 0;
  }
}

extern "C" {

// Build a problem from the text of a Par.inp.  synthetic_seed != 0 replaces every material by the
// hash-defined heterogeneous model of the synthetic benchmark config.
void* orc_create_at(const char* parinp_text, unsigned long long synthetic_seed, int renumber, int kd_force_kd1,
                    long long ix0, long long iz0, char* err, int errlen) {
  try {
    OrcHandle* h = new OrcHandle();
    ParInp in = ParInp::from_string(parinp_text);
    read_main(h->pb, h->cart, in);
    h->cart.renumber = renumber != 0;
    h->pb.kd_force_kd1 = kd_force_kd1 != 0;
    if (synthetic_seed != 0) {
      for (auto& mi : h->pb.mat.inputs) {
        mi.synthetic = true;
        mi.seed = synthetic_seed;
        mi.homogeneous = false;
        mi.has_lambda = false;
        mi.ix0 = ix0;
        mi.iz0 = iz0;
      }
    }
    init_main(h->pb, h->cart);
    return h;
  } catch (std::exception& e) {
    set_err(err, errlen, e.what());
    return nullptr;
  }
}

void* orc_create(const char* parinp_text, unsigned long long synthetic_seed, int renumber, int kd_force_kd1,
                 char* err, int errlen) {
  return orc_create_at(parinp_text, synthetic_seed, renumber, kd_force_kd1, 0, 0, err, errlen);
}

void orc_destroy(void* hv) { delete (OrcHandle*)hv; }

int orc_step(void* hv, int nsteps, char* err, int errlen) {
  OrcHandle* h = (OrcHandle*)hv;
  try {
    for (int k = 0; k < nsteps; ++k) step(h->pb);
    return 0;
  } catch (std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
}

// time nsteps of solve() only (no receivers / fault writes), seconds of wall clock
double orc_time_solve(void* hv, int nsteps) {
  OrcHandle* h = (OrcHandle*)hv;
  auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < nsteps; ++k) {
    h->pb.it++;
    h->pb.time.time = h->pb.it * h->pb.time.dt;
    solve(h->pb);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// one compute_Fint on the current d,v into a scratch array (exposed as "fint")
int orc_compute_fint(void* hv) {
  OrcHandle* h = (OrcHandle*)hv;
  h->scratch_d.assign(h->pb.d.size(), 0.0);
  compute_Fint(h->pb, h->scratch_d, h->pb.d, h->pb.v);
  return 0;
}

int orc_set_fields(void* hv, const double* d, const double* v, const double* a) {
  OrcHandle* h = (OrcHandle*)hv;
  size_t n = h->pb.d.size();
  if (d) std::memcpy(h->pb.d.data(), d, n * 8);
  if (v) std::memcpy(h->pb.v.data(), v, n * 8);
  if (a) std::memcpy(h->pb.a_.data(), a, n * 8);
  return 0;
}

double orc_energy_Ek(void* hv) { return energy_Ek(((OrcHandle*)hv)->pb); }
double orc_energy_EW(void* hv) { return energy_EW(((OrcHandle*)hv)->pb); }

// element-wise snapshot fields of PLOT_FIELD: fills the scratch array exposed as "snap"
int orc_snapshot(void* hv, int what) {
  OrcHandle* h = (OrcHandle*)hv;
  snapshot_elem(h->pb, (char)what, h->scratch_f);
  return 0;
}

long long orc_get_int(void* hv, const char* name_) {
  OrcHandle* h = (OrcHandle*)hv;
  Problem& pb = h->pb;
  std::string n(name_);
  if (n == "ngll") return pb.grid.ngll;
  if (n == "ndof") return pb.ndof;
  if (n == "nelem") return pb.grid.nelem;
  if (n == "npoin") return pb.grid.npoin;
  if (n == "npoin_fem") return pb.grid.npoin_fem;
  if (n == "nx") return pb.grid.nx;
  if (n == "nz") return pb.grid.nz;
  if (n == "ezflt") return pb.grid.ezflt;
  if (n == "nelast") return pb.nelast;
  if (n == "ncoefsets") return pb.ncoefsets;
  if (n == "nkv") return (long long)pb.kv_elem.size();
  if (n == "npl") return (long long)pb.pl_elem.size();
  if (n == "nvs") return (long long)pb.vs_elem.size();
  if (n == "ndm") return (long long)pb.dm_elem.size();
  if (n == "nt") return pb.time.nt;
  if (n == "it") return pb.it;
  if (n == "nbc") return (long long)pb.bc.size();
  if (n == "nbnd") return (long long)pb.grid.bnds.size();
  if (n == "nsrc") return (long long)pb.src.size();
  if (n == "nstages") return pb.time.nstages;
  if (n == "scheme")
    return pb.time.kind == "leapfrog" ? 0 : (pb.time.kind == "newmark" ? 1 : (pb.time.kind == "HHT-alpha" ? 2 : (pb.time.nstages > 0 ? 3 : -1)));
  if (n == "rec.present") return pb.rec.present;
  if (n == "rec.nx") return pb.rec.nx;
  if (n == "rec.nt") return pb.rec.nt;
  if (n == "rec.isamp") return pb.rec.isamp;
  if (n == "rec.atnode") return pb.rec.AtNode;
  if (n == "rec.field") return pb.rec.field;
  if (n.rfind("bc.", 0) == 0) {
    size_t p2 = n.find('.', 3);
    int i = std::atoi(n.substr(3, p2 - 3).c_str());
    std::string f = n.substr(p2 + 1);
    if (i < 0 || i >= (int)pb.bc.size()) return -999;
    Bc& b = pb.bc[i];
    if (f == "kind") return b.kind;
    if (f == "tag1") return b.tag[0];
    if (f == "tag2") return b.tag[1];
    if (b.kind == IS_ABSORB) {
      if (f == "np") return b.abso->topo->npoin;
      if (f == "nbe") return b.abso->topo->nelem;
      if (f == "stacey") return b.abso->stacey;
      if (f == "is_flat") return b.abso->is_flat;
      if (f == "periodic") return b.abso->periodic;
    }
    if (b.kind == IS_DIRNEU) {
      if (f == "np") return b.dirneu->topo->npoin;
      if (f == "kind_h") return b.dirneu->kind[0];
      if (f == "kind_v") return b.dirneu->kind[1];
    }
    if (b.kind == IS_DYNFLT) {
      BcDynflt& d = *b.dynflt;
      if (f == "np") return d.npoin;
      if (f == "two_sides") return d.two_sides;
      if (f == "allow_opening") return d.allow_opening;
      if (f == "has_swf") return d.swf != nullptr;
      if (f == "has_rsf") return d.rsf != nullptr;
      if (f == "has_twf") return d.twf != nullptr;
      if (f == "swf.kind") return d.swf ? d.swf->kind : 0;
      if (f == "swf.healing") return d.swf ? d.swf->healing : 0;
      if (f == "rsf.kind") return d.rsf ? d.rsf->kind : 0;
      if (f == "twf.kind") return d.twf ? d.twf->kind : 0;
      if (f == "normal.kind") return d.normal.kind;
      if (f == "oix1") return d.oix1;
      if (f == "oixn") return d.oixn;
      if (f == "oixd") return d.oixd;
      if (f == "oit") return d.oit;
      if (f == "oitd") return d.oitd;
      if (f == "oit0") return (long long)std::lround(d.ot1 / pb.time.dt);
      if (f == "onx") return d.onx();
      if (f == "nout") return d.nout;
    }
  }
  if (n.rfind("bnd.", 0) == 0) {
    size_t p2 = n.find('.', 4);
    int i = std::atoi(n.substr(4, p2 - 4).c_str());
    std::string f = n.substr(p2 + 1);
    if (i < 0 || i >= (int)pb.grid.bnds.size()) return -999;
    Boundary& b = pb.grid.bnds[i];
    if (f == "tag") return b.tag;
    if (f == "nelem") return b.nelem;
    if (f == "npoin") return b.npoin;
  }
  if (n.rfind("src.", 0) == 0) {
    size_t p2 = n.find('.', 4);
    int i = std::atoi(n.substr(4, p2 - 4).c_str());
    std::string f = n.substr(p2 + 1);
    if (f == "iglob") return pb.src[i].iglob;
    if (f == "moment") return pb.src[i].moment ? 1 : 0;
    if (f == "nterms") return (long long)pb.src[i].mnode.size();
  }
  return -999;
}

double orc_get_double(void* hv, const char* name_) {
  OrcHandle* h = (OrcHandle*)hv;
  Problem& pb = h->pb;
  std::string n(name_);
  if (n == "dt") return pb.time.dt;
  if (n == "beta") return pb.time.beta;
  if (n == "gamma") return pb.time.gamma;
  if (n == "alpha") return pb.time.alpha;
  if (n == "courant") return pb.time.courant;
  if (n == "total") return pb.time.total;
  if (n == "time") return pb.time.time;
  if (n == "grid_cfl") return pb.grid_cfl;
  if (n == "CoefA2V") return pb.time.CoefA2V();
  if (n == "CoefA2D") return pb.time.CoefA2D();
  if (n == "CoefA2Vrhs") return pb.time.CoefA2Vrhs();
  if (n == "rec.tsamp") return pb.rec.tsamp;
  if (n.rfind("bc.", 0) == 0) {
    size_t p2 = n.find('.', 3);
    int i = std::atoi(n.substr(3, p2 - 3).c_str());
    std::string f = n.substr(p2 + 1);
    Bc& b = pb.bc[i];
    if (b.kind == IS_DYNFLT) {
      BcDynflt& d = *b.dynflt;
      if (f == "CoefA2V") return d.CoefA2V;
      if (f == "CoefA2D") return d.CoefA2D;
      if (f == "normal.T") return d.normal.T;
      if (f == "normal.L") return d.normal.L;
      if (f == "normal.V") return d.normal.V;
      if (f == "normal.coef") return d.normal.coef;
      if (f == "swf.dt") return d.swf ? d.swf->dt : 0;
      if (f == "rsf.dt") return d.rsf ? d.rsf->dt : 0;
      if (d.twf) {
        Twf& t = *d.twf;
        if (f == "twf.X") return t.X;
        if (f == "twf.Z") return t.Z;
        if (f == "twf.mus") return t.mus;
        if (f == "twf.mud") return t.mud;
        if (f == "twf.mu0") return t.mu0;
        if (f == "twf.L") return t.L;
        if (f == "twf.V") return t.V;
        if (f == "twf.T") return t.T;
        if (f == "twf.Dc") return t.Dc;
      }
    }
  }
  if (n.rfind("src.", 0) == 0) {
    size_t p2 = n.find('.', 4);
    int i = std::atoi(n.substr(4, p2 - 4).c_str());
    std::string f = n.substr(p2 + 1);
    Source& s = pb.src[i];
    if (f == "dir1") return s.dir[0];
    if (f == "dir2") return s.dir[1];
    if (f == "f0") return s.stf.f0;
    if (f == "t0") return s.stf.t0;
    if (f == "ampli") return s.stf.ampli;
    if (f == "tdelay") return s.tdelay;
    if (f == "src_ampli") return s.ampli;
  }
  return std::nan("");
}

// source time function value STF_get(t - tdelay) * ampli for source i at time t (src_gen.f90:300-303)
double orc_stf(void* hv, int i, double t) {
  OrcHandle* h = (OrcHandle*)hv;
  Source& s = h->pb.src[i];
  double a = s.stf.eval(t - s.tdelay);
  return a * s.ampli;
}

// Pointer access to internal arrays.  dtype: 'i' int32, 'd' float64, 'f' float32.  Returns element
// count or -1 if the name is unknown.
long long orc_array(void* hv, const char* name_, const void** ptr, char* dtype) {
  OrcHandle* h = (OrcHandle*)hv;
  Problem& pb = h->pb;
  std::string n(name_);
#define RET_D(v) { *ptr = (v).data(); *dtype = 'd'; return (long long)(v).size(); }
#define RET_I(v) { *ptr = (v).data(); *dtype = 'i'; return (long long)(v).size(); }
#define RET_F(v) { *ptr = (v).data(); *dtype = 'f'; return (long long)(v).size(); }
  if (n == "ibool") RET_I(pb.grid.ibool);
  if (n == "time.a") RET_D(pb.time.a);
  if (n == "time.b") RET_D(pb.time.b);
  if (n.rfind("src.", 0) == 0) {
    size_t p2 = n.find('.', 4);
    int i = std::atoi(n.substr(4, p2 - 4).c_str());
    std::string f = n.substr(p2 + 1);
    if (i >= 0 && i < (int)pb.src.size()) {
      if (f == "mnode") RET_I(pb.src[i].mnode);
      if (f == "mcoef") RET_D(pb.src[i].mcoef);
    }
  }
  if (n == "coord") RET_D(pb.grid.coord);
  if (n == "coord_fem") RET_D(pb.grid.coord_fem);
  if (n == "knods") RET_I(pb.grid.knods);
  if (n == "tag") RET_I(pb.grid.tag);
  if (n == "perm") RET_I(pb.grid.perm);
  if (n == "xgll") RET_D(pb.grid.xgll);
  if (n == "wgll") RET_D(pb.grid.wgll);
  if (n == "H") RET_D(pb.grid.H);
  if (n == "Ht") RET_D(pb.grid.Ht);
  if (n == "a") RET_D(pb.a);
  if (n == "beta25d") RET_D(pb.beta25d);
  if (n == "elem2set") RET_I(pb.elem2set);
  if (n == "kv_elem") RET_I(pb.kv_elem);
  if (n == "kv_eta") RET_D(pb.kv_eta);
  if (n == "pl_elem") RET_I(pb.pl_elem);
  if (n == "pl_ep") RET_D(pb.pl_ep);
  if (n == "pl_par") RET_D(pb.pl_par);
  if (n == "vs_el") RET_D(pb.vs_el);
  if (n.rfind("mat.", 0) == 0) {  // mat.<tag>.theta | wbody | moduli (lambda_inf, mu_inf) of a VISCO material
    const size_t dot = n.find('.', 4);
    const int tag = std::atoi(n.substr(4, dot - 4).c_str());
    if (tag >= 1 && tag <= (int)pb.mat.inputs.size()) {
      const auto& mi = pb.mat.inputs[tag - 1];
      const std::string what = n.substr(dot + 1);
      if (what == "theta") RET_D(mi.theta);
      if (what == "wbody") RET_D(mi.wbody);
      if (what == "moduli") {
        static thread_local std::vector<double> mm;
        mm = {mi.lambda, mi.mu};
        RET_D(mm);
      }
    }
  }
  if (n == "vs_etot") RET_D(pb.vs_etot);
  if (n == "dm_state") RET_D(pb.dm_state);
  if (n == "dm_par") RET_D(pb.dm_par);
  if (n == "rmass") RET_D(pb.rmass);
  if (n == "mass") RET_D(pb.mass);
  if (n == "d") RET_D(pb.d);
  if (n == "v") RET_D(pb.v);
  if (n == "acc") RET_D(pb.a_);
  if (n == "fint") RET_D(h->scratch_d);
  if (n == "snap") RET_F(h->scratch_f);
  if (n == "rec.coord") RET_D(pb.rec.coord);
  if (n == "rec.iglob") RET_I(pb.rec.iglob);
  if (n == "rec.interp") RET_D(pb.rec.interp);
  if (n == "rec.einterp") RET_I(pb.rec.einterp);
  if (n == "rec.sis") RET_F(pb.rec.sis);
  if (n.rfind("bnd.", 0) == 0) {
    size_t p2 = n.find('.', 4);
    int i = std::atoi(n.substr(4, p2 - 4).c_str());
    std::string f = n.substr(p2 + 1);
    Boundary& b = pb.grid.bnds[i];
    if (f == "elem") RET_I(b.elem);
    if (f == "edge") RET_I(b.edge);
    if (f == "node") RET_I(b.node);
    if (f == "ibool") RET_I(b.ibool);
  }
  if (n.rfind("bc.", 0) == 0) {
    size_t p2 = n.find('.', 3);
    int i = std::atoi(n.substr(3, p2 - 3).c_str());
    std::string f = n.substr(p2 + 1);
    Bc& b = pb.bc[i];
    if (b.kind == IS_PERIOD) {
      if (f == "master") RET_I(b.perio->master->node);
      if (f == "slave") RET_I(b.perio->slave->node);
    }
    if (b.kind == IS_ABSORB) {
      BcAbso& a = *b.abso;
      if (f == "node") RET_I(a.topo->node);
      if (f == "bibool") RET_I(a.topo->ibool);
      if (f == "C") RET_D(a.C);
      if (f == "K") RET_D(a.K);
      if (f == "n") RET_D(a.n);
    }
    if (b.kind == IS_DIRNEU) {
      if (f == "node") RET_I(b.dirneu->topo->node);
    }
    if (b.kind == IS_DYNFLT) {
      BcDynflt& d = *b.dynflt;
      if (f == "node1") RET_I(d.node1);
      if (f == "node2") RET_I(d.node2);
      if (f == "n1") RET_D(d.n1);
      if (f == "B") RET_D(d.B);
      if (f == "invM1") RET_D(d.invM1);
      if (f == "invM2") RET_D(d.invM2);
      if (f == "Z") RET_D(d.Z);
      if (f == "T0") RET_D(d.T0);
      if (f == "T") RET_D(d.T);
      if (f == "Tstick") RET_D(d.Tstick);
      if (f == "V") RET_D(d.V);
      if (f == "D") RET_D(d.D);
      if (f == "MU") RET_D(d.MU);
      if (f == "cohesion") RET_D(d.cohesion);
      if (f == "coord") RET_D(d.coord);
      if (f == "sigma") RET_D(d.normal.sigma);
      if (f == "out") RET_F(d.out);
      if (f == "potency") RET_D(d.potency);
      if (d.swf) {
        if (f == "swf.dc") RET_D(d.swf->dc);
        if (f == "swf.mus") RET_D(d.swf->mus);
        if (f == "swf.mud") RET_D(d.swf->mud);
        if (f == "swf.theta") RET_D(d.swf->theta);
        if (f == "swf.p") RET_D(d.swf->p);
        if (f == "swf.alpha") RET_D(d.swf->alpha);
      }
      if (d.rsf) {
        if (f == "rsf.dc") RET_D(d.rsf->dc);
        if (f == "rsf.mus") RET_D(d.rsf->mus);
        if (f == "rsf.a") RET_D(d.rsf->a);
        if (f == "rsf.b") RET_D(d.rsf->b);
        if (f == "rsf.Vstar") RET_D(d.rsf->Vstar);
        if (f == "rsf.theta") RET_D(d.rsf->theta);
        if (f == "rsf.Vc") RET_D(d.rsf->Vc);
        if (f == "rsf.Tc") RET_D(d.rsf->Tc);
        if (f == "rsf.coeft") RET_D(d.rsf->coeft);
      }
    }
  }
  return -1;
}

// stand-alone pieces for unit tests ---------------------------------------------------------
int orc_gll(int n, double* x, double* w, double* H) {
  std::vector<double> xv, wv, Hv;
  gll::get_GLL_info(n, xv, wv, Hv);
  std::memcpy(x, xv.data(), n * 8);
  std::memcpy(w, wv.data(), n * 8);
  std::memcpy(H, Hv.data(), (size_t)n * n * 8);
  return 0;
}
int orc_rcm(int nx, int nz, int* perm /*nelem, 1-based values*/) {
  std::vector<int> p, pi;
  rcmlib::structured_rcm(nx, nz, p, pi);
  std::memcpy(perm, p.data() + 1, (size_t)nx * nz * 4);
  return 0;
}
double orc_hash_u(unsigned long long seed, unsigned long long ix, unsigned long long iz, unsigned long long k) {
  return hash_u(seed, ix, iz, k);
}

}  // extern "C"
