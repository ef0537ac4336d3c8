// C++ host side of the B200 time-stepping path: the reference's module API for this path
// (problem_type, read_main, init_main, solve, REC_store/REC_write, BC_write, IO_abort), with the same
// names, argument meaning and error behaviour, over the C-ABI of include/sem2d_b200.h.
//
// The reference is a Fortran program (no Fortran compiler exists in the build image, SURVEY.md
// header), so this layer is what a maintainer's ISO_C_BINDING shim looks like when
// written in C++.  Scope = what the device-side structured builder provides: &MESH_CART boxes with
// one ELAST material, ABSORB, PERIOD and DIRNEU sides, DYNFLT (the split-node row of `ezflt`, or the bottom / top
// side as a one-sided fault) with slip-weakening or rate-and-state friction and ORDER0 / PWCONR / GAUSSIAN
// distributions, Kelvin-Voigt damping (ELAST + KV), FORCE
// and moment-tensor sources, REC_LINE stations at nodes, the leapfrog, Newmark, HHT-alpha and symplectic schemes.  Anything else in a Par.inp is
// refused with IO_abort, never silently ignored -- except plotting (&SNAP_*), which is not on the path.
//
//   reference                                           here
//   read_main   SRC/input.f90:12-63                     read_main(pb, "Par.inp")
//   init_main   SRC/init.f90:16-131                     init_main(pb)        (builds the problem in HBM)
//   solve       SRC/solver.f90:20-35                    solve(pb, nsteps)    (nsteps passes of main.f90:51-99)
//   REC_store   SRC/receivers.f90:309-344               inside solve (device), REC_fetch copies rec%sis back
//   REC_write   SRC/receivers.f90:351-392               REC_write(rec)       Ux/Uy/Uz_sem2d.dat (SEP), header
//   BC_write    SRC/bc_gen.f90:313-337                  inside solve (device), BC_DYNFLT_flush writes the files
//   IO_abort    SRC/stdio.f90:205-214                   throws io_abort; main prints and stops
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/sem2d_b200.h"
#include "namelist.hpp"
#include "stf.hpp"

namespace sem2d {

struct io_abort : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] inline void IO_abort(const std::string& msg) { throw io_abort(msg); }

// timescheme_type (SRC/time.f90:5-11)
struct timescheme_type {
  std::string kind = "leapfrog";
  double dt = 0.0, courant = 0.5, total = 0.0, time = 0.0;
  double alpha = 1.0, beta = 0.0, gamma = 0.5;
  int nt = 0;
  int nstages = 0;               // symplectic schemes: time%a(1:nstages+1), time%b(1:nstages)
  std::vector<double> a, b;
};

// source_type with a so_force_type mechanism (SRC/src_gen.f90:17-27, SRC/src_force.f90:9-12)
struct source_type {
  double coord[2] = {0, 0};
  double tdelay = 0.0;
  stf_type stf;
  double dir[2] = {0, 1};
  bool moment = false;           // so_moment_type (SRC/src_moment.f90:9-14): M(2,ndof) column-major
  double M[4] = {0, 0, 0, 0};
  int32_t id = -1;
};

// rec_type (SRC/receivers.f90:9-20)
struct rec_type {
  int number = 0, isamp = 1, nx = 0, nt = 0;
  char SeisField = 'V', irepr = 'D';
  bool AtNode = true;
  double first[2] = {0, 0}, last[2] = {0, 0};
  double tsamp = 0.0;
  std::vector<double> coord;  // (2,nx)
  std::vector<float> sis;     // (nt,nx,ndof)
};

// cd_type (SRC/distribution_cd.f90): a constant or a spatial distribution evaluated at node coordinates.
// Distributions provided: ORDER0 (SRC/distribution_order0.f90: blocks of constant value on an x-z grid of
// zones) and PWCONR (SRC/distribution_pwconr.f90: constant in concentric rings around a point).
// get_attenuation (SRC/mat_visco.f90:251-340): relaxation frequencies log-spaced over [fmin, fmax], the anelastic
// coefficients Y_alpha, Y_beta that make 1/Q constant at 2 Nbody - 1 frequencies in the least-squares sense, the
// unrelaxed moduli and theta(Nbody,3).  The reference solves the two small least-squares problems with Numerical
// Recipes' SVD; the minimiser is unique, here it comes from Householder QR.
struct attenuation_type {
  std::vector<double> theta, wbody;  // theta[c * Nbody + b]
  double mu_inf = 0, lambda_inf = 0;
};
inline std::vector<double> lsq_qr(std::vector<double> A, int m, int n, std::vector<double> b) {  // A(m,n) column-major
  for (int k = 0; k < n; ++k) {
    double nrm = 0;
    for (int i = k; i < m; ++i) nrm += A[i + (size_t)m * k] * A[i + (size_t)m * k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = A[k + (size_t)m * k] > 0 ? -nrm : nrm;
    std::vector<double> v((size_t)m, 0.0);
    for (int i = k; i < m; ++i) v[(size_t)i] = A[i + (size_t)m * k];
    v[(size_t)k] -= alpha;
    double vv = 0;
    for (int i = k; i < m; ++i) vv += v[(size_t)i] * v[(size_t)i];
    if (vv == 0.0) continue;
    for (int j = k; j < n; ++j) {
      double s = 0;
      for (int i = k; i < m; ++i) s += v[(size_t)i] * A[i + (size_t)m * j];
      s = 2.0 * s / vv;
      for (int i = k; i < m; ++i) A[i + (size_t)m * j] -= s * v[(size_t)i];
    }
    double s = 0;
    for (int i = k; i < m; ++i) s += v[(size_t)i] * b[(size_t)i];
    s = 2.0 * s / vv;
    for (int i = k; i < m; ++i) b[(size_t)i] -= s * v[(size_t)i];
  }
  std::vector<double> x((size_t)n, 0.0);
  for (int k = n - 1; k >= 0; --k) {
    double s = b[(size_t)k];
    for (int j = k + 1; j < n; ++j) s -= A[k + (size_t)m * j] * x[(size_t)j];
    x[(size_t)k] = s / A[k + (size_t)m * k];
  }
  return x;
}
inline attenuation_type get_attenuation(double cp, double cs, double rho, double QP, double QS, int Nbody, double fmin, double fmax) {
  const double PI_ = 3.141592653589793;
  const int Nf = 2 * Nbody - 1;
  const double w0 = 2.0 * PI_ * std::pow(fmin * fmax, 0.5), wmin = 2.0 * PI_ * fmin, wmax = 2.0 * PI_ * fmax;
  std::vector<double> w((size_t)Nf, w0);
  if (Nbody > 1)
    for (int i = 1; i <= Nf; ++i) w[(size_t)i - 1] = std::exp(std::log(wmin) + (i - 1) * (std::log(wmax) - std::log(wmin)) / (Nf - 1));
  attenuation_type at;
  at.wbody.resize((size_t)Nbody);
  for (int j = 1; j <= Nbody; ++j) at.wbody[(size_t)j - 1] = w[(size_t)(2 * j - 2)];
  std::vector<double> AP((size_t)Nf * Nbody), AS((size_t)Nf * Nbody);
  for (int i = 0; i < Nf; ++i)
    for (int j = 0; j < Nbody; ++j) {
      const double wb = at.wbody[(size_t)j], wi = w[(size_t)i];
      AP[i + (size_t)Nf * j] = (wb * wi + wb * wb / QP) / (wb * wb + wi * wi);
      AS[i + (size_t)Nf * j] = (wb * wi + wb * wb / QS) / (wb * wb + wi * wi);
    }
  const std::vector<double> Ya = lsq_qr(AP, Nf, Nbody, std::vector<double>((size_t)Nf, 1.0 / QP));
  const std::vector<double> Yb = lsq_qr(AS, Nf, Nbody, std::vector<double>((size_t)Nf, 1.0 / QS));
  double RP1 = 1, RP2 = 0, RS1 = 1, RS2 = 0;
  for (int j = 0; j < Nbody; ++j) {
    const double r = w0 / at.wbody[(size_t)j], den = 1.0 + r * r;
    RP1 -= Ya[(size_t)j] / den;
    RP2 += Ya[(size_t)j] * r / den;
    RS1 -= Yb[(size_t)j] / den;
    RS2 += Yb[(size_t)j] * r / den;
  }
  const double RP = std::sqrt(RP1 * RP1 + RP2 * RP2), RS = std::sqrt(RS1 * RS1 + RS2 * RS2);
  const double mu = rho * cs * cs, lambda = rho * (cp * cp - 2.0 * cs * cs);
  at.mu_inf = mu * (RS + RS1) / (2 * RS * RS);
  at.lambda_inf = (lambda + 2.0 * mu) * (RP + RP1) / (2 * RP * RP) - 2.0 * at.mu_inf;
  at.theta.assign((size_t)Nbody * 3, 0.0);
  for (int j = 0; j < Nbody; ++j) {
    at.theta[(size_t)j] = (at.lambda_inf + 2.0 * at.mu_inf) * Ya[(size_t)j];
    at.theta[(size_t)(j + Nbody)] = (at.lambda_inf + 2.0 * at.mu_inf) * Ya[(size_t)j] - 2.0 * at.mu_inf * Yb[(size_t)j];
    at.theta[(size_t)(j + 2 * Nbody)] = 2.0 * at.mu_inf * Yb[(size_t)j];
  }
  return at;
}

struct cd_type {
  double c = 0.0;
  int dist = 0;  // 0 constant, 1 ORDER0, 2 PWCONR, 3 GAUSSIAN
  double gx0 = 0, gz0 = 0, glx = 1, glz = 1, goff = 0, gamp = 1;  // GAUSSIAN (SRC/distribution_gaussian.f90:25-50)
  int gorder = 1;
  int xn = 1, zn = 1;
  std::vector<double> xb, zb, val;  // ORDER0: zone boundaries, val(xn,zn) column-major
  double ref[2] = {0, 0};           // PWCONR: reference point, radii, values
  std::vector<double> rad;
  double eval(double x, double z) const {
    if (dist == 0) return c;
    if (dist == 3) {  // generate_gaussian_dist (distribution_gaussian.f90:72-73)
      const double a = (x - gx0) / glx, b = (z - gz0) / glz;
      return goff + gamp * std::exp(-std::pow(std::pow(a, 2.0) + std::pow(b, 2.0), (double)gorder));
    }
    if (dist == 1) {  // generate_order0_dist / zone (distribution_order0.f90:73-105)
      auto zone = [](double q, int nz, const std::vector<double>& b) {
        if (nz == 1) return 1;
        int k = 1;
        for (; k <= nz - 1; ++k)
          if (q < b[k - 1]) break;
        return k;
      };
      return val[(zone(x, xn, xb) - 1) + (size_t)xn * (zone(z, zn, zb) - 1)];
    }
    const double r = std::sqrt((x - ref[0]) * (x - ref[0]) + (z - ref[1]) * (z - ref[1]));  // distribution_pwconr.f90:69-85
    size_t iz = 0;
    for (; iz + 1 < val.size(); ++iz)
      if (r <= rad[iz]) break;
    return val[iz];
  }
};

// bc_type (SRC/bc_gen.f90:29-41), the kinds this host hands to the device
struct bc_type {
  int tag[2] = {0, 0};
  std::string kind;
  bool stacey = false;  // &BC_ABSORB
  int oxi[3] = {1, 2147483647, 1};
  double ot1 = 0.0, otd = 0.0;
  int32_t fault_id = -1;
  int np = 0, oitd = 1, oit = 0;
  std::vector<double> MU0;   // initial friction coefficient of the output nodes (FltXX_init_sem2d.tab)
  // general &BC_DYNFLT (SRC/bc_dynflt.f90:104-229): constants or distributions, one or two friction laws
  cd_type cd_Tn, cd_Tt, cd_cohesion, cd_V;
  cd_type cd_S[5];   // background stress Sxx, Sxy, Sxz, Syz, Szz (SRC/bc_dynflt.f90:165-169,376-400)
  bool opening = true, has_swf = false, has_rsf = false;
  int swf_kind = 1, rsf_kind = 1, nor_kind = 1;
  bool has_twf = false;                                                  // SRC/bc_dynflt_twf.f90:12-16
  int twf_kind = 1;
  double twf[9] = {0, 0, 0.6, 0.5, 0.6, 1.0, 1e3, 1.7976931348623157e308, 1.7976931348623157e308};  // X,Z,MuS,MuD,Mu0,L,V,T,Dc
  bool swf_healing = false;
  cd_type swf_Dc, swf_MuS, swf_MuD, swf_alpha, swf_p;                    // SRC/bc_dynflt_swf.f90:30-100
  cd_type rsf_Dc, rsf_MuS, rsf_a, rsf_b, rsf_Vstar, rsf_theta, rsf_Vc;   // SRC/bc_dynflt_rsf.f90:30-90
  double nor_L = 1, nor_V = 1, nor_T = 1;                                // SRC/bc_dynflt_normal.f90:30-60
  int kind_h = 1, kind_v = 1;                                            // &BC_DIRNEU: 1 Neumann, 2 Dirichlet
};

// problem_type (SRC/problem_class.f90:19-46): what the host keeps; fields, operator data and boundary
// tables live in HBM behind pb.gpu
struct problem_type {
  s2d_handle gpu = nullptr;
  int iexec = 0, ngll = 9, ndof = 2, ItInfo = 100;
  double W = 0.0;   // &GENERAL W when given (finite seismogenic width, 2.5D), else 0
  std::string title;
  // &MESH_CART
  double xlim[2] = {0, 0}, zlim[2] = {0, 0};
  int nelem[2] = {0, 0}, ezflt = 0;
  int fztag = 0, fznz = 1;                       // mesh_cartesian.f90:281-290
  struct domain_type {                           // &MESH_CART_DOMAIN (mesh_cartesian.f90:150-176)
    int tag = 1, ex[2] = {0, 0}, ez[2] = {0, 0};
  };
  std::vector<domain_type> domains;
  // matpro_input_type per tag (SRC/prop_mat.f90:21-25): &MAT_ELASTIC (SRC/mat_elastic.f90:104-129) with constants
  // or DIST_* fields, &MAT_KV (SRC/mat_kelvin_voigt.f90:35-66)
  struct material_type {
    bool set = false, kv = false, ETAxDT = true;
    bool plastic = false;                        // kind='PLAST' (&MAT_PLASTIC, SRC/mat_plastic.f90:46-118)
    double phi = 0, coh = 0, Tv = 0, e0[3] = {0, 0, 0};
    bool damage = false;                         // kind='DMG' (&MAT_DAMAGE, SRC/mat_damage.f90:80-173): phi, e0 above
    double Cd = 0, Rdmg = 0, beta = 0, alpha0 = 0, ep0[3] = {0, 0, 0};
    bool visco = false;                          // kind='VISCO' (&MAT_VISCO, SRC/mat_visco.f90:42-113)
    double QP = 0, QS = 0, fmin = 0, fmax = 0;
    int Nbody = 0;
    cd_type rho, cp, cs, eta;
    bool homogeneous() const { return rho.dist == 0 && cp.dist == 0 && cs.dist == 0; }
  };
  std::vector<material_type> mat;                // index tag-1
  double rho = 0, cp = 0, cs = 0;                // material of tag 1 when it is homogeneous (builder default)
  bool has_kv = false;
  bool has_plastic = false;
  bool has_visco = false;
  bool has_damage = false;
  timescheme_type time;
  std::vector<bc_type> bc;
  std::vector<source_type> src;
  std::unique_ptr<rec_type> rec;
  // &SNAP_DEF (SRC/plot_gen.f90:58-110): binary snapshots; field_names = 'DVAESdc' (plot_gen.f90:13)
  bool snap_bin = false, snap_fields[7] = {false, false, false, false, false, false, false};
  int snap_itd = 100, snap_it1 = 0;
  int64_t npoin = 0, nelem_total = 0;
  int it = 0;
  int precision = 8, device = -1;
  bool compute_energies = false;   // COMPUTE_ENERGIES (SRC/constants.f90:22): energy_sem2d.tab, one line per step
  bool renumber = true;   // RCM element order and the node numbering of a stock reference build (--natural-order: off)
  unsigned long long hash_seed = 0;  // != 0: the heterogeneous hash medium of the synthetic benchmark family
  ~problem_type() {
    if (gpu) s2d_destroy(gpu);
  }
};

inline void s2d_check(const problem_type& pb, int rc, const char* where) {
  if (rc == S2D_OK) return;
  const char* m = pb.gpu ? s2d_last_error(pb.gpu) : "";
  IO_abort(std::string(where) + ": " + (m && *m ? m : "device call failed") + " (code " + std::to_string(rc) + ")");
}

// DIST_CD_Read (SRC/distribution_cd.f90:26-61): the constant `key`, or -- when `key`H names a distribution --
// the next &DIST_<name> block after position `cur` (the reference reads the file forward) and its records
inline cd_type DIST_CD_Read(const namelist_file& in, const nml_group& g, const std::string& key, double dflt, size_t& cur) {
  cd_type cd;
  cd.c = g.real8(key, dflt);
  const std::string name = g.text(key + "h", "");
  if (name.empty()) return cd;
  const long k = in.find("DIST_" + name, cur);
  if (k < 0) IO_abort("DIST_read: DIST_" + name + " input block not found");
  cur = (size_t)k + 1;
  const nml_group& d = in.at((size_t)k);
  if (name == "ORDER0") {  // read_order0_dist (distribution_order0.f90:40-69)
    cd.dist = 1;
    cd.xn = d.integer("xn", 1);
    cd.zn = d.integer("zn", 1);
    const size_t nxb = cd.xn > 1 ? cd.xn - 1 : 0, nzb = cd.zn > 1 ? cd.zn - 1 : 0, nv = (size_t)cd.xn * cd.zn;
    const std::vector<double> r = in.records_after((size_t)k, nxb + nzb + nv);
    if (r.size() < nxb + nzb + nv) IO_abort("read_order0_dist: missing records after DIST_ORDER0");
    cd.xb.assign(r.begin(), r.begin() + nxb);
    cd.zb.assign(r.begin() + nxb, r.begin() + nxb + nzb);
    cd.val.assign(r.begin() + nxb + nzb, r.end());
  } else if (name == "PWCONR") {  // read_pwconr_dist (distribution_pwconr.f90:25-41)
    cd.dist = 2;
    const int num = d.integer("num", 0);
    if (num < 2) IO_abort("read_pwconr_dist: needs more than 2 zones (num)");
    cd.ref[0] = d.real8("ref", 0.0, 0);
    cd.ref[1] = d.real8("ref", 0.0, 1);
    const std::vector<double> r = in.records_after((size_t)k, (size_t)(2 * num - 1));
    if ((int)r.size() < 2 * num - 1) IO_abort("read_pwconr_dist: missing records after DIST_PWCONR");
    cd.rad.assign(r.begin(), r.begin() + (num - 1));
    cd.val.assign(r.begin() + (num - 1), r.end());
  } else if (name == "GAUSSIAN") {  // read_gaussian_dist (distribution_gaussian.f90:25-50)
    cd.dist = 3;
    cd.gx0 = d.real8("centered_at", 0.0, 0);
    cd.gz0 = d.real8("centered_at", 0.0, 1);
    cd.glx = d.real8("length", 1.0, 0);
    cd.glz = d.real8("length", 1.0, 1);
    cd.goff = d.real8("offset", 0.0);
    cd.gamp = d.real8("ampli", 1.0);
    cd.gorder = d.integer("order", 1);
  } else {
    IO_abort("DIST_read: distribution '" + name + "' is not provided here (ORDER0, PWCONR, GAUSSIAN are)");
  }
  return cd;
}

// ---------------------------------------------------------------------------------------------
// read_main (SRC/input.f90:12-63): &GENERAL, MESH_read, MAT_read, BC_read, TIME_read, SO_read, REC_read
inline void read_main(problem_type& pb, const std::string& file) {
  namelist_file in(file);
  long k = in.find("GENERAL");
  if (k < 0) IO_abort("GENERAL parameters not found");
  {
    const nml_group& g = in.at((size_t)k);  // SRC/input.f90:64-100
    pb.iexec = g.integer("iexec", 0);
    pb.ndof = g.integer("ndof", 2);
    pb.ngll = g.integer("ngll", 9);
    pb.ItInfo = g.integer("itinfo", 100);
    pb.title = g.text("title", "");
    if (pb.ndof > 2 || pb.ndof < 1) IO_abort("GENERAL input block: ndof must be 1 or 2 (SH or P-SV)");
    if (pb.ngll <= 0) IO_abort("GENERAL input block: ngll must be positive");
    if (pb.ItInfo <= 0) IO_abort("GENERAL input block: itInfo must be positive");
    if (g.has("w")) {  // seismogenic width of a 2.5D run (SRC/input.f90:82,119,130)
      pb.W = g.real8("w", 0.0);
      if (pb.W <= 0.0) IO_abort("GENERAL input block: W must be positive");
    }
  }
  // MESH_read (SRC/mesh_gen.f90:61-100), CART_read (SRC/mesh_cartesian.f90:82-170)
  k = in.find("MESH_DEF");
  if (k < 0) IO_abort("MESH_read: MESH_DEF input block not found");
  if (in.at((size_t)k).text("method", "") != "CARTESIAN")
    IO_abort("MESH_read: only method='CARTESIAN' is provided by the B200 structured builder");
  k = in.find("MESH_CART", (size_t)k);
  if (k < 0) IO_abort("CART_read: MESH_CART input block not found");
  {
    const nml_group& g = in.at((size_t)k);
    if (g.count("xlim") < 2 || g.count("zlim") < 2 || g.count("nelem") < 2) IO_abort("CART_read: xlim, zlim and nelem are required");
    for (int q = 0; q < 2; ++q) {
      pb.xlim[q] = g.real8("xlim", 0, q);
      pb.zlim[q] = g.real8("zlim", 0, q);
      pb.nelem[q] = g.integer("nelem", 0, q);
    }
    pb.ezflt = g.integer("ezflt", 0);
    if (g.logical("faultx", false)) pb.ezflt = -1;
    if (g.has("splitd") || g.logical("split", false)) IO_abort("CART_read: split / splitD is not provided by the B200 path");
    if (pb.ezflt == -1) pb.ezflt = pb.nelem[1] / 2;   // mesh_cartesian.f90:124
    if (pb.ezflt < -1) IO_abort("CART_read: ezflt must be >= -1");
    if (pb.ezflt >= pb.nelem[1]) IO_abort("CART_read: ezflt must be < nelem(2)");
    pb.fztag = g.integer("fztag", 0);
    pb.fznz = g.integer("fznz", 1);
    if (pb.fztag < 0) IO_abort("MESH_LAYERS_read: fztag must be positive");
    if (pb.fznz < 1) IO_abort("MESH_LAYERS_read: fznz must be strictly positive");
  }
  for (long q = in.find("MESH_CART_DOMAIN"); q >= 0; q = in.find("MESH_CART_DOMAIN", (size_t)q + 1)) {
    const nml_group& g = in.at((size_t)q);
    problem_type::domain_type dm;
    dm.tag = g.integer("tag", 0);
    for (int c = 0; c < 2; ++c) {
      dm.ex[c] = g.integer("ex", 0, c);
      dm.ez[c] = g.integer("ez", 0, c);
    }
    if (dm.tag < 1) IO_abort("CART_read: tag null, negative or missing");
    if (dm.ex[0] < 1 || dm.ex[0] > pb.nelem[0] || dm.ex[1] < 1 || dm.ex[1] > pb.nelem[0]) IO_abort("CART_read: ex out of bounds or missing");
    if (dm.ez[0] < 1 || dm.ez[0] > pb.nelem[1] || dm.ez[1] < 1 || dm.ez[1] > pb.nelem[1]) IO_abort("CART_read: ez out of bounds or missing");
    pb.domains.push_back(dm);
  }
  // MAT_read (SRC/mat_gen.f90:101-200): one MATERIAL block per tag; each kind reads ITS input block forward from
  // the position right after the MATERIAL block (so several tags can share one &MAT_ELASTIC that follows them)
  k = in.find("MATERIAL");
  if (k < 0) IO_abort("MAT_read: MATERIAL block not found");
  {
    int numat = 0, ntags = 0;
    for (long m0 = k; m0 >= 0; m0 = in.find("MATERIAL", (size_t)m0 + 1)) {
      const int tag = in.at((size_t)m0).integer("tag", 0);
      if (tag <= 0) IO_abort("MAT_read: tag must be positive");
      ntags = std::max(ntags, tag);
      ++numat;
    }
    if (numat != ntags) IO_abort("MAT_read: inconsistent or missing tags");
    pb.mat.assign((size_t)numat, problem_type::material_type());
  }
  for (long m0 = k; m0 >= 0; m0 = in.find("MATERIAL", (size_t)m0 + 1)) {
    const nml_group& g = in.at((size_t)m0);
    problem_type::material_type& M = pb.mat[(size_t)g.integer("tag", 0) - 1];
    const std::string k1 = g.text("kind", "ELAST", 0), k2 = g.count("kind") >= 2 ? g.text("kind", "", 1) : std::string();
    if (k1 == "PLAST") {  // MAT_PLAST_read (SRC/mat_plastic.f90:66-118): constants only
      if (!(k2.empty() || k2 == "KV")) IO_abort("MAT_read: kind='PLAST' combined with '" + k2 + "' is not on the B200 path");
      if (k2 == "KV") {  // MAT_KV_read (SRC/mat_kelvin_voigt.f90:35-66): the one non-exclusive material (mat_gen.f90:350-354)
        const long q = in.find("MAT_KV", (size_t)m0);
        if (q < 0) IO_abort("MAT_KV_read: MAT_KV input block not found");
        size_t c2 = (size_t)q + 1;
        M.eta = DIST_CD_Read(in, in.at((size_t)q), "eta", 0.0, c2);
        M.ETAxDT = in.at((size_t)q).logical("etaxdt", true);
        M.kv = true;
        pb.has_kv = true;
      }
      const long m = in.find("MAT_PLASTIC", (size_t)m0);
      if (m < 0) IO_abort("MAT_PLAST_read: MAT_PLASTIC input block not found");
      const nml_group& e = in.at((size_t)m);
      M.rho.c = e.real8("rho", 0.0);
      M.cp.c = e.real8("cp", 0.0);
      M.cs.c = e.real8("cs", 0.0);
      if (!(M.rho.c > 0) || !(M.cp.c > 0) || !(M.cs.c > 0)) IO_abort("MAT_PLAST_read: incomplete input (rho, cp, cs)");
      M.phi = e.real8("phi", 0.0);
      M.coh = e.real8("coh", 0.0);
      M.Tv = e.real8("tv", 0.0);
      for (int q = 0; q < 3; ++q) M.e0[q] = e.real8("e0", 0.0, (size_t)q);
      M.plastic = true;
      M.set = true;
      pb.has_plastic = true;
      continue;
    }
    if (k1 == "DMG") {  // MAT_DMG_read (SRC/mat_damage.f90:108-173): constants only
      if (!(k2.empty() || k2 == "KV")) IO_abort("MAT_read: kind='DMG' combined with '" + k2 + "' is not on the B200 path");
      if (k2 == "KV") {  // MAT_KV_read (SRC/mat_kelvin_voigt.f90:35-66): the one non-exclusive material (mat_gen.f90:350-354)
        const long q = in.find("MAT_KV", (size_t)m0);
        if (q < 0) IO_abort("MAT_KV_read: MAT_KV input block not found");
        size_t c2 = (size_t)q + 1;
        M.eta = DIST_CD_Read(in, in.at((size_t)q), "eta", 0.0, c2);
        M.ETAxDT = in.at((size_t)q).logical("etaxdt", true);
        M.kv = true;
        pb.has_kv = true;
      }
      const long m = in.find("MAT_DAMAGE", (size_t)m0);
      if (m < 0) IO_abort("MAT_DMG_read: MAT_DAMAGE input block not found");
      const nml_group& e = in.at((size_t)m);
      M.rho.c = e.real8("rho", 0.0);
      M.cp.c = e.real8("cp", 0.0);
      M.cs.c = e.real8("cs", 0.0);
      if (!(M.rho.c > 0) || !(M.cp.c > 0) || !(M.cs.c > 0)) IO_abort("MAT_DMG_read: incomplete input (rho, cp, cs)");
      M.phi = e.real8("phi", 0.0);
      M.alpha0 = e.real8("alpha", 0.0);
      M.Cd = e.real8("cd", 0.0);
      M.beta = e.real8("beta", 0.0);
      M.Rdmg = e.real8("r", 0.0);
      for (int q = 0; q < 3; ++q) M.e0[q] = e.real8("e0", 0.0, (size_t)q);
      for (int q = 0; q < 3; ++q) M.ep0[q] = e.real8("ep", 0.0, (size_t)q);
      M.damage = true;
      M.set = true;
      pb.has_damage = true;
      continue;
    }
    if (k1 == "VISCO") {  // MAT_VISCO_read (SRC/mat_visco.f90:65-113): constants only
      if (!(k2.empty() || k2 == "KV")) IO_abort("MAT_read: kind='VISCO' combined with '" + k2 + "' is not on the B200 path");
      if (k2 == "KV") {  // MAT_KV_read (SRC/mat_kelvin_voigt.f90:35-66): the one non-exclusive material (mat_gen.f90:350-354)
        const long q = in.find("MAT_KV", (size_t)m0);
        if (q < 0) IO_abort("MAT_KV_read: MAT_KV input block not found");
        size_t c2 = (size_t)q + 1;
        M.eta = DIST_CD_Read(in, in.at((size_t)q), "eta", 0.0, c2);
        M.ETAxDT = in.at((size_t)q).logical("etaxdt", true);
        M.kv = true;
        pb.has_kv = true;
      }
      const long m = in.find("MAT_VISCO", (size_t)m0);
      if (m < 0) IO_abort("MAT_VISCO_read: MAT_VISCO input block not found");
      const nml_group& e = in.at((size_t)m);
      M.rho.c = e.real8("rho", 0.0);
      M.cp.c = e.real8("cp", 0.0);
      M.cs.c = e.real8("cs", 0.0);
      M.QP = e.real8("qp", 0.0);
      M.QS = e.real8("qs", 0.0);
      M.Nbody = (int)e.real8("nbody", 0.0);
      M.fmin = e.real8("fmin", 0.0);
      M.fmax = e.real8("fmax", 0.0);
      if (!(M.rho.c > 0) || !(M.cp.c > 0) || !(M.cs.c > 0) || !(M.QP > 0) || !(M.QS > 0) || !(M.fmin > 0) || !(M.fmax > M.fmin))
        IO_abort("MAT_VISCO_read: incomplete input (rho, cp, cs, QP, QS, fmin < fmax)");
      if (M.Nbody < 1 || M.Nbody > 8) IO_abort("MAT_VISCO_read: Nbody must be in 1..8 on the B200 path");
      M.visco = true;
      M.set = true;
      pb.has_visco = true;
      continue;
    }
    if (k1 != "ELAST" || !(k2.empty() || k2 == "KV"))
      IO_abort("MAT_read: only kind='ELAST', kind='ELAST','KV', kind='PLAST', kind='VISCO' and kind='DMG' are on the B200 path");
    const long m = in.find("MAT_ELASTIC", (size_t)m0);
    if (m < 0) IO_abort("MAT_ELAST_read: MAT_ELASTIC input block not found");
    const nml_group& e = in.at((size_t)m);  // SRC/mat_elastic.f90:104-129
    if (e.has("c11") || e.has("c55") || e.has("c11h") || e.has("c55h"))
      IO_abort("MAT_ELAST_read: anisotropic materials go through the generic C-ABI (s2d_set_elastic), not this host");
    size_t cur = (size_t)m + 1;  // MAT_setProp reads each property's DIST_* block forward, in the order rho, cp, cs
    M.rho = DIST_CD_Read(in, e, "rho", 0.0, cur);
    M.cp = DIST_CD_Read(in, e, "cp", 0.0, cur);
    M.cs = DIST_CD_Read(in, e, "cs", 0.0, cur);
    if (M.rho.dist == 0 && !(M.rho.c > 0)) IO_abort("MAT_ELAST_read: undefined density (rho)");
    if ((M.cp.dist == 0 && !(M.cp.c > 0)) || (M.cs.dist == 0 && !(M.cs.c > 0))) IO_abort("MAT_ELAST_read: incomplete input");
    if (k2 == "KV") {  // MAT_KV_read (SRC/mat_kelvin_voigt.f90:35-66)
      const long q = in.find("MAT_KV", (size_t)m0);
      if (q < 0) IO_abort("MAT_KV_read: MAT_KV input block not found");
      size_t c2 = (size_t)q + 1;
      M.eta = DIST_CD_Read(in, in.at((size_t)q), "eta", 0.0, c2);
      M.ETAxDT = in.at((size_t)q).logical("etaxdt", true);
      M.kv = true;
      pb.has_kv = true;
    }
    M.set = true;
  }
  pb.rho = pb.mat[0].rho.c;
  pb.cp = pb.mat[0].cp.c;
  pb.cs = pb.mat[0].cs.c;
  // BC_read (SRC/bc_gen.f90:98-188): in input order
  for (long b = in.find("BC_DEF"); b >= 0; b = in.find("BC_DEF", (size_t)b + 1)) {
    const nml_group& g = in.at((size_t)b);
    bc_type bc;
    bc.kind = g.text("kind", "");
    if (g.has("tags")) {
      bc.tag[0] = g.integer("tags", 0, 0);
      bc.tag[1] = g.integer("tags", 0, 1);
    } else {
      bc.tag[0] = g.integer("tag", 0);
    }
    const long nxt = in.find("BC_DEF", (size_t)b + 1);
    auto sub = [&](const char* name) -> const nml_group* {
      const long s = in.find(name, (size_t)b);
      return (s >= 0 && (nxt < 0 || s < nxt)) ? &in.at((size_t)s) : nullptr;
    };
    if (bc.kind == "ABSORB") {  // SRC/bc_abso.f90:62-110
      if (const nml_group* a = sub("BC_ABSORB")) {
        bc.stacey = a->logical("stacey", false);
        if (a->logical("let_wave", true) == false) { /* only matters with an incident wave source */ }
      }
      if (bc.tag[0] < 1 || bc.tag[0] > 4) IO_abort("BC_read: ABSORB tag must be a side of the box (1..4)");
    } else if (bc.kind == "PERIOD") {  // SRC/bc_periodic.f90:29-40: no parameters
      if (bc.tag[1] == 0) IO_abort("BC_read: PERIOD needs tags = master, slave");
    } else if (bc.kind == "DIRNEU") {  // SRC/bc_dirneu.f90:50-112
      const nml_group* dn = sub("BC_DIRNEU");
      if (!dn) IO_abort("bc_DIRNEU_read: no BC_DIRNEU block found");
      if (dn->text("hstf", "none") != "none" || dn->text("vstf", "none") != "none")
        IO_abort("bc_DIRNEU_read: time-dependent Neumann conditions are provided by the generic C-ABI, not by this host");
      bc.kind_h = dn->text("h", "N") == "D" ? 2 : 1;
      bc.kind_v = dn->text("v", "N") == "D" ? 2 : 1;
      if (bc.tag[0] < 1 || bc.tag[0] > 4) IO_abort("BC_read: DIRNEU tag must be a side of the box (1..4)");
    } else if (bc.kind == "DYNFLT") {  // BC_DYNFLT_read (SRC/bc_dynflt.f90:104-229)
      const bool two = bc.tag[0] == 5 && bc.tag[1] == 6, one = (bc.tag[0] == 1 || bc.tag[0] == 3) && bc.tag[1] == 0;
      if (!two && !one)
        IO_abort("BC_read: DYNFLT is provided on tags=5,6 (split-node row of MESH_CART ezflt) or on the bottom / top side (tag=1 or 3)");
      const nml_group* f = sub("BC_DYNFLT");
      if (!f) IO_abort("BC_DYNFLT_read: BC_DYNFLT input block not found");
      for (int q = 0; q < 3; ++q) bc.oxi[q] = f->integer("oxi", bc.oxi[q], q);
      bc.ot1 = f->real8("ot1", 0.0);
      bc.otd = f->real8("otd", 0.0);
      bc.opening = f->logical("opening", true);
      if (f->logical("osides", false)) IO_abort("BC_DYNFLT_read: osides=T is not provided here");
      size_t cur = (size_t)(f - &in.at(0)) + 1;  // distributions are read forward from the block (bc_dynflt.f90:164-172)
      bc.cd_cohesion = DIST_CD_Read(in, *f, "cohesion", 0.0, cur);
      bc.cd_Tn = DIST_CD_Read(in, *f, "tn", 0.0, cur);
      bc.cd_Tt = DIST_CD_Read(in, *f, "tt", 0.0, cur);
      {
        static const char* skey[5] = {"sxx", "sxy", "sxz", "syz", "szz"};
        for (int q = 0; q < 5; ++q) bc.cd_S[q] = DIST_CD_Read(in, *f, skey[q], 0.0, cur);
      }
      bc.cd_V = DIST_CD_Read(in, *f, "v", 1e-12, cur);
      for (size_t q = 0; q < f->count("friction") || q < 1; ++q) {
        const std::string law = f->text("friction", "SWF", q);
        if (law == "SWF") {  // swf_read (SRC/bc_dynflt_swf.f90:30-100)
          const nml_group* w = sub("BC_DYNFLT_SWF");
          static const nml_group none;
          if (!w) w = &none;
          bc.has_swf = true;
          bc.swf_kind = w->integer("kind", 1);
          bc.swf_healing = w->logical("healing", false);
          if (bc.swf_kind < 1 || bc.swf_kind > 3) IO_abort("BC_DYNFLT_SWF: invalid kind");
          size_t c2 = w == &none ? cur : (size_t)(w - &in.at(0)) + 1;
          bc.swf_Dc = DIST_CD_Read(in, *w, "dc", 0.5, c2);
          bc.swf_MuS = DIST_CD_Read(in, *w, "mus", 0.6, c2);
          bc.swf_MuD = DIST_CD_Read(in, *w, "mud", 0.5, c2);
          bc.swf_alpha = DIST_CD_Read(in, *w, "alpha", 0.0, c2);
          bc.swf_p = DIST_CD_Read(in, *w, "p", 3.0, c2);
        } else if (law == "RSF") {  // rsf_read (SRC/bc_dynflt_rsf.f90:30-90)
          const nml_group* w = sub("BC_DYNFLT_RSF");
          static const nml_group none;
          if (!w) w = &none;
          bc.has_rsf = true;
          bc.rsf_kind = w->integer("kind", 1);
          if (bc.rsf_kind < 1 || bc.rsf_kind > 4) IO_abort("BC_DYNFLT_RSF: invalid kind");
          size_t c2 = w == &none ? cur : (size_t)(w - &in.at(0)) + 1;
          bc.rsf_Dc = DIST_CD_Read(in, *w, "dc", 0.5, c2);
          bc.rsf_MuS = DIST_CD_Read(in, *w, "mus", 0.6, c2);
          bc.rsf_a = DIST_CD_Read(in, *w, "a", 0.01, c2);
          bc.rsf_b = DIST_CD_Read(in, *w, "b", 0.02, c2);
          bc.rsf_Vstar = DIST_CD_Read(in, *w, "vstar", 1.0, c2);
          bc.rsf_theta = DIST_CD_Read(in, *w, "theta", 0.0, c2);
          bc.rsf_Vc = DIST_CD_Read(in, *w, "vc", 1e-6, c2);
        } else if (law == "TWF") {  // twf_read (SRC/bc_dynflt_twf.f90:56-100)
          const nml_group* w = sub("BC_DYNFLT_TWF");
          static const nml_group none;
          if (!w) w = &none;
          bc.has_twf = true;
          bc.twf_kind = w->integer("kind", 1);
          bc.twf[0] = w->real8("x", 0.0);
          bc.twf[1] = w->real8("z", 0.0);
          bc.twf[2] = w->real8("mus", 0.6);
          bc.twf[3] = w->real8("mud", 0.5);
          bc.twf[4] = w->real8("mu0", 0.6);
          bc.twf[5] = w->real8("l", 1.0);
          bc.twf[6] = w->real8("v", 1e3);
          bc.twf[7] = w->real8("t", 1.7976931348623157e308);
          bc.twf[8] = w->real8("dc", 1.7976931348623157e308);
          if (bc.twf_kind < 1) IO_abort("BC_SWFF_init: kind must be > 0");
          if (bc.twf_kind > 3) IO_abort("BC_SWFF_init: kind must be < 4");
          if (bc.twf[2] < 0 || bc.twf[3] < 0 || bc.twf[4] < 0) IO_abort("BC_SWFF_init: MuS, MuD, Mu0 must be positive in BC_DYNFLT_TWF input block");
          if (bc.twf[5] <= 0 || bc.twf[6] <= 0 || bc.twf[7] <= 0 || bc.twf[8] <= 0) IO_abort("BC_SWFF_init: L, V, T, Dc must be positive in BC_DYNFLT_TWF input block");
        } else if (!law.empty()) {
          IO_abort("BC_DYNFLT: invalid friction");
        }
      }
      if (bc.has_swf && bc.has_rsf) IO_abort("BC_DYNFLT_read: SWF and RSF together are not provided here");
      if (const nml_group* nr = sub("BC_DYNFLT_NOR")) {  // normal_read (SRC/bc_dynflt_normal.f90:30-60)
        bc.nor_kind = nr->integer("kind", 1);
        bc.nor_L = nr->real8("l", 1.0);
        bc.nor_V = nr->real8("v", 1.0);
        bc.nor_T = nr->real8("t", 1.0);
        if (bc.nor_kind > 3 || bc.nor_kind < 0) IO_abort("BC_SWFF_init: invalid kind in BC_DYNFLT_NOR input block");
      }
    } else {
      IO_abort("BC_read: boundary kind '" + bc.kind + "' is not provided by the B200 structured builder (ABSORB, PERIOD, DIRNEU, DYNFLT are)");
    }
    pb.bc.push_back(bc);
  }
  // TIME_read (SRC/time.f90:122-230)
  k = in.find("TIME");
  if (k < 0) IO_abort("TIME parameters not found");
  {
    const nml_group& g = in.at((size_t)k);
    timescheme_type& t = pb.time;
    t.kind = g.text("kind", "leapfrog");
    int NbSteps = g.integer("nbsteps", 0);
    t.dt = g.real8("dt", 0.0);
    t.courant = g.real8("courant", 0.5);
    double TotalTime = g.real8("totaltime", 0.0);
    if (NbSteps < 0) IO_abort("TIME: NbSteps must be positive");
    if (t.dt < 0.0) IO_abort("TIME: Dt must be positive");
    if (t.courant < 0.0 || t.courant > 0.6) IO_abort("TIME: Courant out of range [0,0.6]");
    if (TotalTime < 0.0) IO_abort("TIME: TotalTime must be positive");
    if (NbSteps * TotalTime != 0.0) IO_abort("TIME: bad combination of settings, NbSteps or TotalTime");
    if (t.dt > 0.0) {
      if (TotalTime > 0.0) NbSteps = (int)std::ceil(TotalTime / t.dt);
      TotalTime = t.dt * NbSteps;
    }
    t.nt = NbSteps;
    t.total = TotalTime;
    if (t.kind == "newmark") {
      const long m = in.find("TIME_NEWMARK", (size_t)k);
      if (m >= 0) {
        t.beta = in.at((size_t)m).real8("beta", 0.0);
        t.gamma = in.at((size_t)m).real8("gamma", 0.5);
      }
      if (t.beta != 0.0) IO_abort("TIME: only the explicit Newmark scheme (beta=0) is on the B200 path");
    } else if (t.kind == "HHT-alpha") {  // SRC/time.f90:232-246
      double alpha = 0.5, rho = 0.5;
      const long m = in.find("TIME_HHTA", (size_t)k);
      if (m >= 0) {
        alpha = in.at((size_t)m).real8("alpha", 0.5);
        rho = in.at((size_t)m).real8("rho", 0.5);
      }
      if (alpha < 0.0 || alpha > 1.0) IO_abort("TIME_HHTA: alpha is out of range [0,1]");
      if (rho < 0.5 || rho > 1.0) IO_abort("TIME_HHTA: rho is out of range [0.5,1]");
      t.alpha = alpha;
      t.gamma = 1.5 - alpha;
      t.beta = alpha != 1.0 ? 1.0 - alpha - rho * rho * (rho - 1.0) / ((1.0 - alpha) * ((1.0 + rho) * (1.0 + rho) * (1.0 + rho))) : 0.0;
    } else if (t.kind == "symp_PV") {  // SRC/time.f90:248-255
      t.nstages = 1;
      t.a = {0.5, 0.5};
      t.b = {1.0};
    } else if (t.kind == "symp_PFR") {  // :257-269
      const double theta = 1.0 / (2.0 - std::pow(2.0, 1.0 / 3.0));
      t.nstages = 3;
      t.a = {theta / 2.0, (1.0 - theta) / 2.0, (1.0 - theta) / 2.0, theta / 2.0};
      t.b = {theta, 1.0 - 2.0 * theta, theta};
    } else if (t.kind == "symp_PEFRL") {  // :271-287
      const double xi = 0.1786178958448091, lambda = -0.2123418310626054, chi = -0.06626458266981849;
      t.nstages = 4;
      t.a = {xi, chi, 1.0 - 2.0 * (chi + xi), chi, xi};
      t.b = {0.5 - lambda, lambda, lambda, 0.5 - lambda};
    } else if (t.kind != "leapfrog") {
      IO_abort("TIME: unknown kind");
    }
  }
  // SO_read (SRC/src_gen.f90:126-214)
  for (long s = in.find("SRC_DEF"); s >= 0; s = in.find("SRC_DEF", (size_t)s + 1)) {
    const nml_group& g = in.at((size_t)s);
    source_type so;
    if (g.has("file")) IO_abort("SO_read: source positions from a file are not provided here");
    if (g.count("coord") < 2) IO_abort("SO_read: coord is required");
    so.coord[0] = g.real8("coord", 0, 0);
    so.coord[1] = g.real8("coord", 0, 1);
    so.tdelay = g.real8("delay", 0.0);
    const std::string mech = g.text("mechanism", "");
    so.stf = STF_read(g.text("stf", ""), in, (size_t)s);
    const double PI = 3.141592653589793238462643383279502884197;
    if (mech == "FORCE") {
      const long m = in.find("SRC_FORCE", (size_t)s);  // SRC/src_force.f90:40-58
      const double angle = (m >= 0 ? in.at((size_t)m).real8("angle", 0.0) : 0.0) * PI / 180.0;
      so.dir[0] = -std::sin(angle);
      so.dir[1] = std::cos(angle);
    } else if (mech == "EXPLOSION" || mech == "DOUBLE_COUPLE" || mech == "MOMENT") {  // SRC/src_moment.f90:26-104
      so.moment = true;
      if (mech == "EXPLOSION") {
        if (pb.ndof != 2) IO_abort("SRC_MOMENT_read: explosion only allowed in PSV (ndof=2)");
        so.M[0] = so.M[3] = 1.0;
      } else if (mech == "DOUBLE_COUPLE") {
        const long m = in.find("SRC_DOUBLE_COUPLE", (size_t)s);
        if (m < 0) IO_abort("SRC_MOMENT_read: SRC_DOUBLE_COUPLE input block not found");
        const double dip = in.at((size_t)m).real8("dip", 90.0) * PI / 180.0;
        const double n1 = std::sin(dip), n2 = std::cos(dip);
        if (pb.ndof == 2) {
          const double r1 = -std::cos(dip), r2 = std::sin(dip);
          so.M[0] = 2.0 * r1 * n1;
          so.M[2] = r1 * n2 + r2 * n1;
          so.M[1] = so.M[2];
          so.M[3] = 2.0 * r2 * n2;
        } else {
          so.M[0] = n1;
          so.M[1] = n2;
        }
      } else {
        const long m = in.find("SRC_MOMENT", (size_t)s);
        if (m < 0) IO_abort("SRC_MOMENT_read: SRC_MOMENT input block not found");
        const nml_group& q = in.at((size_t)m);
        if (pb.ndof == 2) {
          so.M[0] = q.real8("mxx", 0.0);
          so.M[2] = q.real8("mxz", 0.0);
          so.M[1] = q.real8("mzx", 0.0);
          so.M[3] = q.real8("mzz", 0.0);
        } else {
          so.M[0] = q.real8("myx", 0.0);
          so.M[1] = q.real8("myz", 0.0);
        }
      }
    } else {
      IO_abort("SO_read: mechanism '" + mech + "' is not provided here (FORCE, EXPLOSION, DOUBLE_COUPLE, MOMENT are)");
    }
    pb.src.push_back(so);
  }
  // read_plot_gen (SRC/plot_gen.f90:58-110).  PostScript / AVS / Visual3 / GMT plotting is not on the
  // path and is skipped; binary snapshots (bin=T, the default) of D, V, A are written as PLOT_FIELD does.
  k = in.find("SNAP_DEF");
  {
    std::string fields = "V";
    pb.snap_bin = true;
    if (k >= 0) {
      const nml_group& g = in.at((size_t)k);
      pb.snap_bin = g.logical("bin", true);
      fields = g.text("fields", "V");
      pb.snap_itd = g.integer("itd", 100);
      pb.snap_it1 = g.integer("it1", 0);
    }
    static const char field_names[] = "DVAESdc";
    for (int i = 0; i < 7; ++i) pb.snap_fields[i] = fields.find(field_names[i]) != std::string::npos;  // scan(fields, ...) > 0
    if (pb.ndof == 1) pb.snap_fields[5] = pb.snap_fields[6] = false;  // no div / curl for SH (plot_gen.f90:100)
    if (pb.snap_itd <= 0) IO_abort("SNAP_DEF: itd must be positive");
  }
  // REC_read (SRC/receivers.f90:62-140)
  k = in.find("REC_LINE");
  if (k >= 0) {
    const nml_group& g = in.at((size_t)k);
    pb.rec.reset(new rec_type());
    rec_type& r = *pb.rec;
    r.number = g.integer("number", 0);
    r.isamp = g.integer("isamp", 1);
    r.SeisField = g.text("field", "V")[0];
    r.AtNode = g.logical("atnode", true);
    r.irepr = g.text("irepr", "D")[0];
    if (r.number < 0) IO_abort("REC_read: \"number\" must be positive");
    if (r.SeisField != 'D' && r.SeisField != 'V' && r.SeisField != 'A') IO_abort("REC_read: parameter field has wrong value [D,V,A]");
    if (g.text("file", "none") != "none") IO_abort("REC_read: station files are not provided here");
    if (g.count("first") < 2 || g.count("last") < 2) IO_abort("REC_read: first and last are required");
    for (int q = 0; q < 2; ++q) {
      r.first[q] = g.real8("first", 0, q);
      r.last[q] = g.real8("last", 0, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// init_main (SRC/init.f90:16-131): mesh, numbering, operator data, mass, dt, boundaries, sources and
// receivers -- built on the device by the structured builder, in the order init_main uses
inline void init_main(problem_type& pb) {
  s2d_cart_desc d;
  std::memset(&d, 0, sizeof(d));
  d.ngll = pb.ngll;
  d.ndof = pb.ndof;
  d.nx = pb.nelem[0];
  d.nz = pb.nelem[1];
  d.ezflt = pb.ezflt;
  d.x0 = pb.xlim[0];
  d.x1 = pb.xlim[1];
  d.z0 = pb.zlim[0];
  d.z1 = pb.zlim[1];
  d.seed = pb.hash_seed;
  d.rho = pb.rho;
  d.cp = pb.cp;
  d.cs = pb.cs;
  d.precision = pb.precision;
  d.scheme.kind = pb.time.kind == "newmark" ? 1 : (pb.time.kind == "HHT-alpha" ? 2 : (pb.time.nstages > 0 ? 3 : 0));
  d.scheme.nstages = pb.time.nstages;
  for (size_t q = 0; q < pb.time.a.size(); ++q) d.scheme.coa[q] = pb.time.a[q];
  for (size_t q = 0; q < pb.time.b.size(); ++q) d.scheme.cob[q] = pb.time.b[q];
  d.scheme.dt = pb.time.dt;
  d.scheme.beta = pb.time.beta;
  d.scheme.gamma = pb.time.gamma;
  d.scheme.alpha = pb.time.alpha;
  d.courant = pb.time.courant;
  d.device = pb.device;
  d.renumber = pb.renumber ? 1 : 0;  // OPT_RENUMBER (SRC/constants.f90:10-15): the reference's default
  const int rc = s2d_cart_create(&pb.gpu, &d);
  if (rc == S2D_ENODEV) IO_abort("init_main: no CUDA device (the B200 path has no CPU fallback)");
  if (rc != S2D_OK) IO_abort("init_main: s2d_cart_create failed (code " + std::to_string(rc) + ")");
  double dt = 0;
  s2d_check(pb, s2d_cart_info(pb.gpu, &pb.npoin, &pb.nelem_total, &dt), "init_main");
  // MAT_init_prop (SRC/mat_gen.f90:204-303): the element tags of CART_build (mesh_cartesian.f90:270-290), then every
  // tag's material evaluated at the GLL points of its elements.  One homogeneous material for the whole box is what
  // s2d_cart_create already built; anything else is handed over with s2d_cart_set_material.
  const int nx = pb.nelem[0], nz = pb.nelem[1], N = pb.ngll, n2 = N * N;
  std::vector<int> tag((size_t)nx * nz, pb.domains.empty() ? 1 : 0);
  for (const auto& dm : pb.domains)
    for (int i = dm.ex[0]; i <= dm.ex[1]; ++i)
      for (int j = dm.ez[0]; j <= dm.ez[1]; ++j) tag[(size_t)(i - 1) + (size_t)nx * (j - 1)] = dm.tag;
  if (pb.fztag > 0) {  // the element rows next to the fault (at the bottom when ezflt = 0)
    const int j1 = std::max(pb.ezflt + 1 - pb.fznz, 1), j2 = std::min(pb.ezflt + pb.fznz, nz);
    for (int j = j1; j <= j2; ++j)
      for (int i = 1; i <= nx; ++i) tag[(size_t)(i - 1) + (size_t)nx * (j - 1)] = pb.fztag;
  }
  bool one_material = pb.mat[0].homogeneous();
  for (int tg : tag) {
    if (tg == 0) IO_abort("CART_build: Domain tags not entirely set");
    if (tg < 1 || tg > (int)pb.mat.size() || !pb.mat[(size_t)tg - 1].set)
      IO_abort("ELAST_init: element tag does not correspond to a material number");
    const auto& M = pb.mat[(size_t)tg - 1];
    one_material = one_material && M.homogeneous() && M.rho.c == pb.rho && M.cp.c == pb.cp && M.cs.c == pb.cs;
  }
  // coordinates of the GLL points of element e = ix + nx*iz (natural order: the order s2d_cart_set_material and
  // s2d_cart_set_kv_elems expect whatever the numbering of the outputs), as the builder computes them
  std::vector<double> xgll((size_t)N);
  s2d_check(pb, s2d_cart_get_gll(pb.gpu, xgll.data(), nullptr, nullptr), "MAT_init_prop");
  const double hx = (pb.xlim[1] - pb.xlim[0]) / nx, hz = (pb.zlim[1] - pb.zlim[0]) / nz;
  auto gll_xz = [&](size_t e, int q, double& x, double& z) {
    const int ix = (int)(e % (size_t)nx), iz = (int)(e / (size_t)nx), i = q % N, j = q / N;
    x = pb.xlim[0] + hx * (ix + 0.5 * (xgll[(size_t)i] + 1.0));
    z = pb.zlim[0] + hz * (iz + 0.5 * (xgll[(size_t)j] + 1.0));
  };
  if (!one_material) {
    if (pb.hash_seed) IO_abort("init_main: the synthetic hash medium replaces the deck's materials; give one homogeneous material");
    const size_t nn = (size_t)pb.nelem_total * n2;
    std::vector<double> rho(nn), cp(nn), cs(nn);
    for (size_t e = 0; e < (size_t)pb.nelem_total; ++e) {
      const auto& M = pb.mat[(size_t)tag[e] - 1];
      for (int q = 0; q < n2; ++q) {
        double x, z;
        gll_xz(e, q, x, z);
        rho[e * n2 + q] = M.rho.eval(x, z);
        cp[e * n2 + q] = M.cp.eval(x, z);
        cs[e * n2 + q] = M.cs.eval(x, z);
      }
    }
    s2d_check(pb, s2d_cart_set_material(pb.gpu, rho.data(), cp.data(), cs.data()), "MAT_init_prop");
    s2d_check(pb, s2d_cart_info(pb.gpu, &pb.npoin, &pb.nelem_total, &dt), "init_main");
  }
  if (pb.W > 0.0) s2d_check(pb, s2d_cart_set_w25d(pb.gpu, pb.W), "MAT_ELAST_init_25D");
  if (pb.has_damage) {  // MAT_DMG_init_elem_prop / _work (SRC/mat_damage.f90:176-279): one set per DMG tag
    if (pb.ndof != 2) IO_abort("MAT_init_work: the damage rheology requires ndof=2 (P-SV) ");
    if (pb.has_plastic || pb.has_visco) IO_abort("MAT_read: DMG together with PLAST or VISCO materials is not on the B200 path");
    std::vector<int> set_of_tag(pb.mat.size(), 0);
    std::vector<double> par;
    int nsets = 0;
    for (size_t tg = 0; tg < pb.mat.size(); ++tg) {
      const auto& M = pb.mat[tg];
      if (!M.damage) continue;
      set_of_tag[tg] = ++nsets;
      const double lam = M.rho.c * (M.cp.c * M.cp.c - 2.0 * M.cs.c * M.cs.c), mu = M.rho.c * M.cs.c * M.cs.c;
      const double row[13] = {lam, mu, M.phi, M.alpha0, M.Cd, M.beta, M.Rdmg, M.e0[0], M.e0[1], M.e0[2], M.ep0[0], M.ep0[1], M.ep0[2]};
      par.insert(par.end(), row, row + 13);
    }
    std::vector<int32_t> eset((size_t)pb.nelem_total);
    for (size_t e = 0; e < eset.size(); ++e) eset[e] = set_of_tag[(size_t)tag[e] - 1];
    s2d_check(pb, s2d_cart_set_damage(pb.gpu, nsets, par.data(), eset.data()), "MAT_DMG_init_elem_work");
  }
  if (pb.has_visco) {  // MAT_VISCO_init_elem_prop (SRC/mat_visco.f90:116-161): get_attenuation per VISCO tag
    if (pb.ndof != 2) IO_abort("MAT_init_work: visco-elasticity requires ndof=2 (P-SV) ");
    if (pb.has_plastic || pb.has_damage) IO_abort("MAT_read: VISCO together with PLAST or DMG materials is not on the B200 path");
    std::vector<int> set_of_tag(pb.mat.size(), 0);
    std::vector<int32_t> nbody;
    std::vector<double> moduli, wbody, theta;
    int nsets = 0;
    for (size_t tg = 0; tg < pb.mat.size(); ++tg) {
      const auto& M = pb.mat[tg];
      if (!M.visco) continue;
      set_of_tag[tg] = ++nsets;
      attenuation_type at = get_attenuation(M.cp.c, M.cs.c, M.rho.c, M.QP, M.QS, M.Nbody, M.fmin, M.fmax);
      nbody.push_back(M.Nbody);
      moduli.push_back(at.lambda_inf);
      moduli.push_back(at.mu_inf);
      for (int b = 0; b < 8; ++b) wbody.push_back(b < M.Nbody ? at.wbody[(size_t)b] : 0.0);
      for (int c = 0; c < 3; ++c)
        for (int b = 0; b < 8; ++b) theta.push_back(b < M.Nbody ? at.theta[(size_t)c * M.Nbody + b] : 0.0);
    }
    std::vector<int32_t> eset((size_t)pb.nelem_total);
    for (size_t e = 0; e < eset.size(); ++e) eset[e] = set_of_tag[(size_t)tag[e] - 1];
    s2d_check(pb, s2d_cart_set_visco(pb.gpu, nsets, nbody.data(), moduli.data(), wbody.data(), theta.data(), eset.data()),
              "MAT_VISCO_init_elem_work");
  }
  if (pb.has_plastic) {  // MAT_init_work (SRC/mat_gen.f90:367-372): one plastic material set per PLAST tag
    if (pb.ndof != 2) IO_abort("MAT_init_work: plasticity requires ndof=2 (P-SV) ");
    std::vector<int> set_of_tag(pb.mat.size(), 0);
    std::vector<double> par;
    int nsets = 0;
    for (size_t tg = 0; tg < pb.mat.size(); ++tg) {
      const auto& M = pb.mat[tg];
      if (!M.plastic) continue;
      set_of_tag[tg] = ++nsets;
      const double six[6] = {M.coh, M.phi, M.Tv, M.e0[0], M.e0[1], M.e0[2]};
      par.insert(par.end(), six, six + 6);
    }
    std::vector<int32_t> eset((size_t)pb.nelem_total);
    for (size_t e = 0; e < eset.size(); ++e) eset[e] = set_of_tag[(size_t)tag[e] - 1];
    s2d_check(pb, s2d_cart_set_plastic(pb.gpu, nsets, par.data(), eset.data()), "MAT_PLAST_init_elem_work");
  }
  // TIME_init (SRC/time.f90:323-341)
  timescheme_type& t = pb.time;
  if (!(t.dt > 0.0)) {
    t.dt = dt;
    if (t.total > 0.0) t.nt = (int)std::ceil(t.total / t.dt);
    t.total = t.nt * t.dt;
  }
  if (pb.has_kv) {  // MAT_KV_init_elem_work (SRC/mat_kelvin_voigt.f90:117-133): eta(ngll,ngll) of every KV element, times dt
    std::vector<int32_t> ids;
    std::vector<double> eta;
    for (size_t e = 0; e < (size_t)pb.nelem_total; ++e) {
      const auto& M = pb.mat[(size_t)tag[e] - 1];
      if (!M.kv) continue;
      ids.push_back((int32_t)e + 1);
      for (int q = 0; q < n2; ++q) {
        double x, z;
        gll_xz(e, q, x, z);
        double v = M.eta.eval(x, z);
        if (M.ETAxDT) v = t.dt * v;
        eta.push_back(v);
      }
    }
    if (!ids.empty()) s2d_check(pb, s2d_cart_set_kv_elems(pb.gpu, (int32_t)ids.size(), ids.data(), eta.data()), "MAT_KV_init");
  }
  // BC_init (SRC/bc_gen.f90:190-251): periodic boundaries first, then input order
  for (bc_type& bc : pb.bc)
    if (bc.kind == "PERIOD") s2d_check(pb, s2d_cart_add_periodic(pb.gpu, bc.tag[0], bc.tag[1]), "BC_PERIO_init");
  for (bc_type& bc : pb.bc) {
    if (bc.kind == "PERIOD") continue;
    if (bc.kind == "ABSORB") {
      s2d_check(pb, s2d_cart_add_abso(pb.gpu, bc.tag[0], bc.stacey ? 1 : 0), "BC_ABSO_init");
    } else if (bc.kind == "DIRNEU") {
      s2d_check(pb, s2d_cart_add_dirneu(pb.gpu, bc.tag[0], bc.kind_h, bc.kind_v), "bc_DIRNEU_init");
    } else {  // BC_DYNFLT_init (SRC/bc_dynflt.f90:231-520): the host evaluates constants / distributions at the
              // fault nodes, the builder adds the topology (nodes, normals, weights, impedances)
      bc.oitd = std::max(1, (int)std::lround(bc.otd / t.dt));  // :452-456
      const int oit = (int)std::lround(bc.ot1 / t.dt);
      int32_t np = 0;
      s2d_check(pb, s2d_cart_fault_nodes(pb.gpu, bc.tag[0], bc.tag[1], &np, nullptr), "BC_DYNFLT_init");
      std::vector<double> xz(2 * (size_t)np);
      s2d_check(pb, s2d_cart_fault_nodes(pb.gpu, bc.tag[0], bc.tag[1], &np, xz.data()), "BC_DYNFLT_init");
      auto gen = [&](const cd_type& cd) {  // DIST_CD_Init (SRC/distribution_cd.f90:65-90)
        std::vector<double> out(np);
        for (int k = 0; k < np; ++k) out[k] = cd.eval(xz[2 * k], xz[2 * k + 1]);
        return out;
      };
      const std::vector<double> Tt = gen(bc.cd_Tt), Tn = gen(bc.cd_Tn), coh = gen(bc.cd_cohesion), V = gen(bc.cd_V);
      for (double c : coh)
        if (c < 0.0) IO_abort("bc_dynflt_init: cohesion must be positive");
      std::vector<double> T0(2 * (size_t)np), V0((size_t)np * pb.ndof, 0.0);
      // background stress resolved on the fault (:376-400); the faults of the builder are horizontal: the normal of
      // side 1 is +z on the split-node row and on the top side, -z on the bottom side
      const double fnx = 0.0, fnz = (bc.tag[1] == 0 && bc.tag[0] == 1) ? -1.0 : 1.0;
      const std::vector<double> Sxx = gen(bc.cd_S[0]), Sxy = gen(bc.cd_S[1]), Sxz = gen(bc.cd_S[2]), Syz = gen(bc.cd_S[3]),
                                Szz = gen(bc.cd_S[4]);
      for (int k = 0; k < np; ++k) {
        const double Tx = Sxx[k] * fnx + Sxz[k] * fnz, Ty = Sxy[k] * fnx + Syz[k] * fnz, Tz = Sxz[k] * fnx + Szz[k] * fnz;
        T0[k] = pb.ndof == 1 ? Tt[k] + Ty : Tt[k] + Tx * fnz - Tz * fnx;
        T0[k + np] = Tn[k] + Tx * fnx + Tz * fnz;
        if (bc.has_rsf) V0[k] = V[k];  // bc%V(:,1) = V for rate-and-state faults (:371-376)
      }
      s2d_dynflt_desc d;
      std::memset(&d, 0, sizeof(d));
      d.np = np;
      d.T0 = T0.data();
      d.cohesion = coh.data();
      d.V0 = V0.data();
      d.allow_opening = bc.opening ? 1 : 0;
      std::vector<double> p1, p2, p3, p4, p5, p6, p7, th;
      if (bc.has_swf) {
        p1 = gen(bc.swf_Dc); p2 = gen(bc.swf_MuS); p3 = gen(bc.swf_MuD); p4 = gen(bc.swf_p); p5 = gen(bc.swf_alpha);
        th.assign(np, 0.0);
        d.swf_kind = bc.swf_kind;
        d.swf_healing = bc.swf_healing ? 1 : 0;
        d.swf_dc = p1.data(); d.swf_mus = p2.data(); d.swf_mud = p3.data(); d.swf_p = p4.data(); d.swf_alpha = p5.data();
        d.swf_theta = th.data();
      }
      if (bc.has_rsf) {
        p1 = gen(bc.rsf_Dc); p2 = gen(bc.rsf_MuS); p3 = gen(bc.rsf_a); p4 = gen(bc.rsf_b); p5 = gen(bc.rsf_Vstar);
        p6 = gen(bc.rsf_theta); p7 = gen(bc.rsf_Vc);
        d.rsf_kind = bc.rsf_kind;
        d.rsf_dc = p1.data(); d.rsf_mus = p2.data(); d.rsf_a = p3.data(); d.rsf_b = p4.data(); d.rsf_Vstar = p5.data();
        d.rsf_theta = p6.data(); d.rsf_Vc = p7.data();
      }
      if (bc.has_twf) {
        d.twf_kind = bc.twf_kind;
        d.twf_X = bc.twf[0]; d.twf_Z = bc.twf[1]; d.twf_mus = bc.twf[2]; d.twf_mud = bc.twf[3]; d.twf_mu0 = bc.twf[4];
        d.twf_L = bc.twf[5]; d.twf_V = bc.twf[6]; d.twf_T = bc.twf[7]; d.twf_Dc = bc.twf[8];
      }
      d.normal_kind = bc.nor_kind;
      d.normal_T = bc.nor_T;
      d.normal_L = bc.nor_L;
      d.normal_V = bc.nor_V;
      d.oix1 = bc.oxi[0];
      d.oixn = bc.oxi[1];
      d.oixd = bc.oxi[2];
      d.oit = oit;
      d.oitd = bc.oitd;
      d.nt_max = t.nt;
      s2d_check(pb, s2d_cart_add_dynflt(pb.gpu, bc.tag[0], bc.tag[1], &d, &bc.fault_id), "BC_DYNFLT_init");
      bc.np = np;
      bc.oxi[0] = std::max(bc.oxi[0], 1);
      bc.oxi[1] = std::min(bc.oxi[1], (int)np);
      bc.oit = oit;
    }
  }
  // SO_init (SRC/src_gen.f90:216-262): nearest node
  for (source_type& so : pb.src) {
    if (so.moment) s2d_check(pb, s2d_cart_add_moment(pb.gpu, so.coord[0], so.coord[1], so.M, &so.id), "SRC_MOMENT_init");
    else s2d_check(pb, s2d_cart_add_force(pb.gpu, so.coord[0], so.coord[1], so.dir, &so.id), "SO_init");
  }
  // REC_init (SRC/receivers.f90:143-226)
  if (pb.rec) {
    rec_type& r = *pb.rec;
    r.nt = t.nt / r.isamp + 1;  // receivers.f90:172
    r.tsamp = t.dt * r.isamp;
    if (r.AtNode)
      s2d_check(pb, s2d_cart_add_receivers(pb.gpu, r.number, r.first[0], r.first[1], r.last[0], r.last[1], r.SeisField,
                                           r.isamp, r.nt),
                "REC_init");
    else
      s2d_check(pb, s2d_cart_add_receivers_interp(pb.gpu, r.number, r.first[0], r.first[1], r.last[0], r.last[1],
                                                  r.SeisField, r.isamp, r.nt),
                "REC_init");
    int32_t nx = 0;
    s2d_check(pb, s2d_cart_receiver_info(pb.gpu, &nx, nullptr), "REC_init");
    r.nx = nx;
    r.coord.resize(2 * (size_t)nx);
    s2d_check(pb, s2d_cart_receiver_info(pb.gpu, &nx, r.coord.data()), "REC_init");
  }
  s2d_check(pb, s2d_commit(pb.gpu, S2D_ASM_PATCH), "init_main");
  for (bc_type& bc : pb.bc)
    if (bc.kind == "DYNFLT") {  // bc%MU as BC_DYNFLT_init leaves it (SRC/bc_dynflt.f90:402-420)
      bc.MU0.resize(bc.np);
      s2d_check(pb, s2d_get_fault_state(pb.gpu, bc.fault_id, nullptr, nullptr, nullptr, nullptr, bc.MU0.data(), nullptr, nullptr),
                "BC_DYNFLT_init");
    }
  pb.it = 0;
}

// ---------------------------------------------------------------------------------------------
// solve (SRC/solver.f90:20-35) + the per-step outputs of the main loop (SRC/main.f90:51-99):
// nsteps passes on the device.  The source amplitudes are evaluated here as SO_add does
// (SRC/src_gen.f90:300-303): stf(time - tdelay), time = it*dt.
inline void solve(problem_type& pb, int nsteps = 1) {
  std::vector<double> ampli;
  const size_t ns = pb.src.size();
  if (ns) {
    const timescheme_type& t = pb.time;
    const int nst = t.nstages > 0 ? t.nstages : 1;
    ampli.resize(ns * (size_t)nsteps * nst);
    for (int k = 0; k < nsteps; ++k) {
      const double time = (pb.it + k + 1) * t.dt;
      if (t.nstages > 0) {  // solve_symplectic (SRC/solver.f90:186-191): one evaluation per stage
        double ts = time - t.dt;
        for (int q = 0; q < nst; ++q) {
          ts = ts + t.dt * t.a[q];
          for (size_t s = 0; s < ns; ++s) ampli[s + ns * ((size_t)k * nst + q)] = STF_get(pb.src[s].stf, ts - pb.src[s].tdelay);
        }
      } else {  // HHT-alpha evaluates at t_alpha = time + (alpha-1)*dt (SRC/solver.f90:116-117)
        const double te = t.kind == "HHT-alpha" ? time + (t.alpha - 1.0) * t.dt : time;
        for (size_t s = 0; s < ns; ++s) ampli[s + ns * (size_t)k] = STF_get(pb.src[s].stf, te - pb.src[s].tdelay);
      }
    }
  }
  s2d_check(pb, s2d_step(pb.gpu, nsteps, ns ? ampli.data() : nullptr, nullptr), "solve");
  pb.it += nsteps;
  pb.time.time = pb.it * pb.time.dt;
}

// Grid files every reader of the snapshots needs (SE_init, SRC/spec_grid.f90:136-141,296-306,365-371):
// grid_sem2d.hdr, ibool_sem2d.dat (int32 (ngll,ngll) per element), coord_sem2d.dat (float32 (2) per node).
// Element order and node numbering are those of the reference's default OPT_RENUMBER = .true. (constants.f90:10-15)
// unless the program was started with --natural-order.
inline void SE_write_grid(problem_type& pb, const std::string& dir = ".") {
  const size_t n2 = (size_t)pb.ngll * pb.ngll;
  std::vector<int32_t> ibool(n2 * (size_t)pb.nelem_total);
  std::vector<double> coord(2 * (size_t)pb.npoin);
  s2d_check(pb, s2d_cart_get(pb.gpu, ibool.data(), nullptr, nullptr, coord.data()), "SE_init");
  if (FILE* f = std::fopen((dir + "/grid_sem2d.hdr").c_str(), "w")) {
    std::fprintf(f, " NELEM  NPGEO  NGNOD  NPOIN  NGLL\n %lld %lld 4 %lld %d\n", (long long)pb.nelem_total,
                 (long long)(pb.nelem[0] + 1) * (pb.nelem[1] + 1) + (pb.ezflt > 0 ? pb.nelem[0] + 1 : 0), (long long)pb.npoin, pb.ngll);
    std::fclose(f);
  }
  if (FILE* f = std::fopen((dir + "/ibool_sem2d.dat").c_str(), "wb")) {
    std::fwrite(ibool.data(), sizeof(int32_t), ibool.size(), f);
    std::fclose(f);
  }
  if (FILE* f = std::fopen((dir + "/coord_sem2d.dat").c_str(), "wb")) {
    std::vector<float> c4(coord.begin(), coord.end());
    std::fwrite(c4.data(), sizeof(float), c4.size(), f);
    std::fclose(f);
  }
}

// PLOT_FIELD's binary branch (SRC/plot_gen.f90:168-215 -> IO_rw_field, SRC/stdio.f90:112-139): one float32
// per node, files <d|v|a><x|z|y>_NNN_sem2d.dat, NNN = (it-IT1)/ITD
inline bool snapshot_due(const problem_type& pb, int it) {
  bool any = false;
  for (bool f : pb.snap_fields) any = any || f;
  if (!pb.snap_bin || !any) return false;
  return it >= pb.snap_it1 && (it - pb.snap_it1) % pb.snap_itd == 0;
}
inline void PLOT_FIELD(problem_type& pb, int it, const std::string& dir = ".") {
  if (!snapshot_due(pb, it)) return;
  const size_t n = (size_t)pb.npoin;
  std::vector<double> d, v, a;
  if (pb.snap_fields[0]) d.resize(n * pb.ndof);
  if (pb.snap_fields[1]) v.resize(n * pb.ndof);
  if (pb.snap_fields[2]) a.resize(n * pb.ndof);
  s2d_check(pb, s2d_get_fields(pb.gpu, d.empty() ? nullptr : d.data(), v.empty() ? nullptr : v.data(), a.empty() ? nullptr : a.data()),
            "PLOT_FIELD");
  const std::vector<double>* fld[3] = {&d, &v, &a};
  const char fchar[3] = {'d', 'v', 'a'};
  std::vector<float> buf(n);
  for (int i = 0; i < 3; ++i) {
    if (!pb.snap_fields[i]) continue;
    for (int c = 0; c < pb.ndof; ++c) {
      char name[64];
      std::snprintf(name, sizeof(name), "%c%c_%03d_sem2d.dat", fchar[i], pb.ndof == 1 ? 'y' : (c == 0 ? 'x' : 'z'),
                    (it - pb.snap_it1) / pb.snap_itd);
      for (size_t q = 0; q < n; ++q) buf[q] = (float)(*fld[i])[q + n * c];
      FILE* f = std::fopen((dir + "/" + name).c_str(), "wb");
      if (!f) IO_abort(std::string("PLOT_FIELD: cannot open ") + name);
      std::fwrite(buf.data(), sizeof(float), n, f);
      std::fclose(f);
    }
  }
  // element-wise fields (plot_gen.f90:222-305): direct-access files, record e = real(field(:,:,k)) of element e
  const size_t ne = (size_t)pb.nelem_total, n2 = (size_t)pb.ngll * pb.ngll;
  const int tagn = (it - pb.snap_it1) / pb.snap_itd;
  auto dump = [&](char what, const std::vector<std::string>& names) {
    std::vector<float> out(names.size() * ne * n2);
    s2d_check(pb, s2d_cart_snapshot_elem(pb.gpu, what, out.data()), "PLOT_FIELD");
    for (size_t k = 0; k < names.size(); ++k) {
      char name[64];
      std::snprintf(name, sizeof(name), "%s_%03d_sem2d.dat", names[k].c_str(), tagn);
      FILE* f = std::fopen((dir + "/" + name).c_str(), "wb");
      if (!f) IO_abort(std::string("PLOT_FIELD: cannot open ") + name);
      std::fwrite(out.data() + k * ne * n2, sizeof(float), ne * n2, f);
      std::fclose(f);
    }
  };
  if (pb.snap_fields[3]) dump('E', pb.ndof == 1 ? std::vector<std::string>{"e13", "e23"} : std::vector<std::string>{"e11", "e22", "e12"});
  if (pb.snap_fields[4]) dump('S', pb.ndof == 1 ? std::vector<std::string>{"s13", "s23"} : std::vector<std::string>{"s11", "s22", "s12"});
  if (pb.snap_fields[5]) dump('d', {"div"});
  if (pb.snap_fields[6]) dump('c', {"curl"});
}

// rec%sis as REC_store has filled it up to now
inline void REC_fetch(problem_type& pb) {
  if (!pb.rec) return;
  rec_type& r = *pb.rec;
  r.sis.resize((size_t)r.nt * r.nx * pb.ndof);
  s2d_check(pb, s2d_get_seis(pb.gpu, r.sis.data()), "REC_store");
}

// REC_init's header (SRC/receivers.f90:214-221) and REC_write's SEP files (:351-392): one direct-access
// record of nt float32 per station, Uy for SH, Ux and Uz for P-SV
inline void REC_write(const rec_type& r, int ndof, const std::string& dir = ".") {
  auto path = [&](const char* n) { return dir + "/" + n; };
  if (FILE* f = std::fopen(path("SeisHeader_sem2d.hdr").c_str(), "w")) {
    std::fprintf(f, " DT NSAMP NSTA\n %.7E %d %d\n XSTA ZSTA\n", (double)(float)r.tsamp, r.nt, r.nx);
    for (int k = 0; k < r.nx; ++k) std::fprintf(f, " %.16E %.16E\n", r.coord[2 * k], r.coord[2 * k + 1]);
    std::fclose(f);
  } else {
    IO_abort("REC_write: cannot open SeisHeader_sem2d.hdr");
  }
  const char* names1[] = {"Uy_sem2d.dat"};
  const char* names2[] = {"Ux_sem2d.dat", "Uz_sem2d.dat"};
  for (int c = 0; c < ndof; ++c) {
    const char* nm = ndof == 1 ? names1[0] : names2[c];
    FILE* f = std::fopen(path(nm).c_str(), "wb");
    if (!f) IO_abort(std::string("REC_write: cannot open ") + nm);
    std::fwrite(r.sis.data() + (size_t)c * r.nt * r.nx, sizeof(float), (size_t)r.nt * r.nx, f);
    std::fclose(f);
  }
}

// The files BC_DYNFLT_init / BC_DYNFLT_write produce (SRC/bc_dynflt.f90:429-520,751-778):
// FltXX_sem2d.hdr, FltXX_init_sem2d.tab, FltXX_sem2d.dat (sequential unformatted: every record framed
// by its 4-byte length, as gfortran/ifort write it) and FltXX_potency_sem2d.tab (6D24.16 per call).
inline void BC_DYNFLT_flush(problem_type& pb, const bc_type& bc, const std::string& dir = ".") {
  int32_t np = 0;
  s2d_check(pb, s2d_cart_fault_info(pb.gpu, bc.fault_id, &np, nullptr, nullptr, nullptr), "BC_write");
  std::vector<double> coord(2 * (size_t)np), T0(2 * (size_t)np), B(np);
  s2d_check(pb, s2d_cart_fault_info(pb.gpu, bc.fault_id, &np, coord.data(), T0.data(), B.data()), "BC_write");
  const int oix1 = bc.oxi[0], oixn = bc.oxi[1], oixd = bc.oxi[2], onx = (oixn - oix1) / oixd + 1;
  int32_t nout = 0, ncalls = 0;
  s2d_check(pb, s2d_get_fault(pb.gpu, bc.fault_id, nullptr, &nout, nullptr, &ncalls), "BC_write");
  std::vector<float> rec((size_t)nout * 6 * onx);
  std::vector<double> pot((size_t)ncalls * 2 * (pb.ndof + 1));
  s2d_check(pb, s2d_get_fault(pb.gpu, bc.fault_id, rec.data(), &nout, pot.data(), &ncalls), "BC_write");
  char base[64];
  std::snprintf(base, sizeof(base), "%s/Flt%02d", dir.c_str(), bc.tag[0]);
  const std::string b(base);
  if (FILE* f = std::fopen((b + "_sem2d.hdr").c_str(), "w")) {
    std::fprintf(f, " NPTS NDAT NSAMP DELT\n %d %d %d %.16E\n", onx, 6, (pb.time.nt - bc.oit) / bc.oitd + 1, pb.time.dt * bc.oitd);
    std::fprintf(f, " Slip:Slip_Rate:Shear_Stress:Normal_Stress:Friction:T_stick\n XPTS ZPTS\n");
    for (int i = oix1 - 1; i < oixn; i += oixd) std::fprintf(f, " %.16E %.16E\n", coord[2 * i], coord[2 * i + 1]);
    std::fclose(f);
  }
  if (FILE* f = std::fopen((b + "_init_sem2d.tab").c_str(), "w")) {
    for (int i = oix1 - 1; i < oixn; i += oixd) std::fprintf(f, " %.16E %.16E %.16E %.16E\n", T0[i], T0[i + np], bc.MU0[i], B[i]);
    std::fclose(f);
  }
  if (FILE* f = std::fopen((b + "_sem2d.dat").c_str(), "wb")) {
    const int32_t len = (int32_t)(onx * sizeof(float));
    for (int r = 0; r < nout * 6; ++r) {
      std::fwrite(&len, 4, 1, f);
      std::fwrite(rec.data() + (size_t)r * onx, sizeof(float), onx, f);
      std::fwrite(&len, 4, 1, f);
    }
    std::fclose(f);
  }
  if (FILE* f = std::fopen((b + "_potency_sem2d.tab").c_str(), "w")) {
    const int w = 2 * (pb.ndof + 1);
    for (int c = 0; c < ncalls; ++c) {
      for (int q = 0; q < w; ++q) std::fprintf(f, "%24.16E", pot[(size_t)c * w + q]);
      std::fprintf(f, "\n");
    }
    std::fclose(f);
  }
}

}  // namespace sem2d
