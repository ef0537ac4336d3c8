// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed
// by the product (sem2dpack_b200/, include/); only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker / reported baseline.
//
// CPU restatement (C++17, FP64, one thread) of SEM2DPACK's explicit time-stepping path and of
// the init code that feeds it.  The reference is Fortran 90 and no Fortran compiler exists in
// this image, so the reference itself cannot be built here; this restatement is
// pinned against the reference's own known-answer artefacts in tests/ (TestSH uyref.mat,
// Lamb's problem EX2DDIR traces, RateState series).
//
// Every function cites the reference file:line it follows (paths under /root/reference/SRC).
// Index conventions: node / element / boundary-node ids stored in tables are 1-based exactly as
// in the reference (so ibool can be compared bit for bit); C arrays are 0-based containers.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "gll.hpp"
#include "parinp.hpp"
#include "rcm.hpp"

namespace orc {

static const double PI = 3.141592653589793;  // constants.f90:43
static const double TINY_XABS = 1e-3;         // constants.f90:35
static const int OPT_NGLL = 5;                // constants.f90:6
static const double HUGE_D = std::numeric_limits<double>::max();

[[noreturn]] inline void IO_abort(const std::string& msg) {  // stdio.f90:205-214
  throw std::runtime_error("IO_abort: " + msg);
}

enum { edge_D = 1, edge_R = 2, edge_U = 3, edge_L = 4 };  // fem_grid.f90:73-76

// ------------------------------------------------------------------------------------------
// counter-based hash for the synthetic heterogeneous material (NOT part of the
// reference: it stands in for a user-supplied heterogeneous model).  The product's device-side
// generator implements the same function so both sides see bit-identical cp, cs, rho.
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double hash_u(uint64_t seed, uint64_t ix, uint64_t iz, uint64_t k) {  // U(-1,1)
  uint64_t h = splitmix64(seed ^ splitmix64(ix * 0x9E3779B97F4A7C15ull + k) ^
                          splitmix64(iz * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull * (k + 1)));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

// ------------------------------------------------------------------------------------------
// distributions (distribution_cd.f90, _gaussian.f90:72-84, _pwconr.f90:69-85, _order0.f90:73-105)
struct Dist {
  enum Kind { CONST = 0, ORDER0, GAUSSIAN, PWCONR } kind = CONST;
  double c = 0.0;
  // gaussian
  double x0 = 0, z0 = 0, lx = 1, lz = 1, level0 = 0, ampli = 1;
  int order = 1;
  // pwconr
  int numzon = 0;
  double ref[2] = {0, 0};
  std::vector<double> radzon, valzon;
  // order0
  int xn = 0, zn = 0;
  std::vector<double> xb, zb, val;  // val(xn,zn) column-major

  static int zone(double coord, int nzones, const std::vector<double>& bound) {  // order0 :93-105
    int z = nzones;
    if (nzones == 1) return z;
    int k;
    for (k = 1; k <= nzones - 1; ++k)
      if (coord < bound[k - 1]) break;
    return k;
  }
  double eval(double x, double z) const {
    switch (kind) {
      case CONST:
        return c;
      case GAUSSIAN: {
        double ax = (x - x0) / lx, az = (z - z0) / lz;
        double r = ax * ax + az * az;  // (..)**2d0 + (..)**2d0
        double rp = r;
        for (int k = 1; k < order; ++k) rp *= r;  // **order (integer power)
        return level0 + ampli * std::exp(-rp);
      }
      case PWCONR: {
        double rad = std::sqrt((x - ref[0]) * (x - ref[0]) + (z - ref[1]) * (z - ref[1]));
        int izone;
        for (izone = 1; izone <= numzon - 1; ++izone)
          if (rad <= radzon[izone - 1]) break;
        return valzon[izone - 1];
      }
      case ORDER0: {
        int ix = zone(x, xn, xb), iz = zone(z, zn, zb);
        return val[(ix - 1) + (size_t)xn * (iz - 1)];
      }
    }
    return 0;
  }
  bool is_dist() const { return kind != CONST; }
};

// DIST_CD_Read + DIST_read (distribution_cd.f90:29-66, distribution_general.f90:52-95)
inline Dist read_cd(ParInp& in, double C, const std::string& Dname_) {
  Dist d;
  std::string Dname = upper(Dname_);
  // trim
  while (!Dname.empty() && Dname.back() == ' ') Dname.pop_back();
  if (Dname.empty()) {
    d.kind = Dist::CONST;
    d.c = C;
    return d;
  }
  if (Dname == "GAUSSIAN") {
    const NmlGroup* g = in.next("DIST_GAUSSIAN");
    if (!g) IO_abort("read_gaussian_dist: DIST_GAUSSIAN parameters missing");
    d.kind = Dist::GAUSSIAN;
    d.x0 = g->dbl("centered_at", 0.0, 0);
    d.z0 = g->dbl("centered_at", 0.0, 1);
    d.lx = g->dbl("length", 1.0, 0);
    d.lz = g->dbl("length", 1.0, 1);
    d.level0 = g->dbl("offset", 0.0);
    d.ampli = g->dbl("ampli", 1.0);
    d.order = g->integer("order", 1);
  } else if (Dname == "PWCONR") {
    const NmlGroup* g = in.next("DIST_PWCONR");
    if (!g) IO_abort("read_pwconr_dist: DIST_PWCONR missing");
    d.kind = Dist::PWCONR;
    d.numzon = g->integer("num", 0);
    if (d.numzon < 2) IO_abort("read_pwconr_dist: needs more than 2 zones (num)");
    d.ref[0] = g->dbl("ref", 0.0, 0);
    d.ref[1] = g->dbl("ref", 0.0, 1);
    d.radzon = in.read_list(d.numzon - 1);
    d.valzon = in.read_list(d.numzon);
  } else if (Dname == "ORDER0") {
    const NmlGroup* g = in.next("DIST_ORDER0");
    if (!g) IO_abort("read_order0_dist: DIST_ORDER0 missing");
    d.kind = Dist::ORDER0;
    d.xn = g->integer("xn", 0);
    d.zn = g->integer("zn", 0);
    if (d.xn > 1) d.xb = in.read_list(d.xn - 1);
    if (d.zn > 1) d.zb = in.read_list(d.zn - 1);
    d.val.resize((size_t)d.xn * d.zn);
    for (int i = 0; i < d.zn; ++i) {
      std::vector<double> row = in.read_list(d.xn);
      for (int k = 0; k < d.xn; ++k) d.val[k + (size_t)d.xn * i] = row[k];
    }
  } else {
    IO_abort("DIST_read: unknown distribution name " + Dname);
  }
  return d;
}

// ------------------------------------------------------------------------------------------
struct Boundary {  // bnd_grid.f90:22-28
  int tag = 0, nelem = 0, npoin = 0, ngnod = 0;
  std::vector<int> elem, edge;  // (nelem) 1-based bulk element / edge id
  std::vector<int> node;        // (npoin) 1-based bulk node ids, sorted
  std::vector<int> ibool;       // (ngnod,nelem) 1-based boundary node index
  bool exists() const { return tag != 0; }
};

struct Grid {  // fem_grid_type (fem_grid.f90:60-71) + sem_grid_type (spec_grid.f90:49-64)
  // macro mesh
  int nx = 0, nz = 0, ezflt = 0;
  int npoin_fem = 0, nelem = 0;
  std::vector<double> coord_fem;  // (2,npoin_fem)
  std::vector<int> knods;         // (4,nelem) 1-based
  std::vector<int> tag;           // (nelem)
  std::vector<int> perm;          // new -> old element id (1-based, [0] unused); identity if no RCM
  bool flat = true;
  std::vector<Boundary> bnds;  // slot k holds tag k+1 (mesh_structured.f90:86-196)
  // spectral grid
  int ngll = 0, npoin = 0;
  double fmax = 1.0, W = HUGE_D;
  std::vector<double> xgll, wgll, H, Ht, wgll2;  // H(ip,ix) column-major
  std::vector<double> shape;                     // (4,ngll,ngll)
  std::vector<double> dshape;                    // (4,2,ngll,ngll)
  std::vector<int> ibool;                        // (ngll,ngll,nelem) 1-based
  std::vector<double> coord;                     // (2,npoin)

  inline int ib(int i, int j, int e) const {  // 1-based i,j,e
    return ibool[(size_t)(i - 1) + (size_t)ngll * ((j - 1) + (size_t)ngll * (e - 1))];
  }
  inline int& ib(int i, int j, int e) {
    return ibool[(size_t)(i - 1) + (size_t)ngll * ((j - 1) + (size_t)ngll * (e - 1))];
  }
  const Boundary* bc_inquire(int tagv) const {  // spec_grid.f90:921-937
    for (const auto& b : bnds)
      if (b.tag == tagv) return &b;
    return nullptr;
  }
  Boundary* bc_inquire(int tagv) {
    for (auto& b : bnds)
      if (b.tag == tagv) return &b;
    return nullptr;
  }
};

// Q4 shape functions (elem_q4.f90:40-79)
inline void Q4_getshape(double s, double t, double* sh) {
  double sp = s + 1.0, sm = s - 1.0, tp = t + 1.0, tm = t - 1.0;
  sh[0] = 0.25 * sm * tm;
  sh[1] = -0.25 * sp * tm;
  sh[2] = 0.25 * sp * tp;
  sh[3] = -0.25 * sm * tp;
}
inline void Q4_getdershape(double s, double t, double* d /*(4,2) col-major*/) {
  double sp = s + 1.0, sm = s - 1.0, tp = t + 1.0, tm = t - 1.0;
  d[0] = 0.25 * tm;
  d[1] = -0.25 * tm;
  d[2] = 0.25 * tp;
  d[3] = -0.25 * tp;
  d[4] = 0.25 * sm;
  d[5] = -0.25 * sp;
  d[6] = 0.25 * sp;
  d[7] = -0.25 * sm;
}

// utils.f90:92-111
inline void invert2(const double A[4] /*col-major 2x2*/, double B[4]) {
  double det = A[0] * A[3] - A[2] * A[1];
  if (det <= 0.0) IO_abort("SE_InverseJacobian: undefined Jacobian");
  B[0] = A[3];
  B[1] = -A[1];
  B[2] = -A[2];
  B[3] = A[0];
  for (int k = 0; k < 4; ++k) B[k] = B[k] / det;
}

// ------------------------------------------------------------------------------------------
// mesh_cartesian.f90:219-314 (CART_build) + mesh_structured.f90:12-196
struct CartSpec {
  double xmin = 0, xmax = 0, zmin = 0, zmax = 0;
  int nx = 0, nz = 0, ezflt = 0, fztag = 0, fznz = 1;
  bool split = false;
  double splitD = HUGE_D;
  struct Dom {
    int tag, ex[2], ez[2];
  };
  std::vector<Dom> domains;
  bool renumber = true;  // constants.f90:11 OPT_RENUMBER
};

inline int sub2ind(int i, int j, int n) { return (j - 1) * n + i; }  // utils.f90:120-123

inline void CART_build(const CartSpec& m, Grid& g) {
  int nxp = m.nx + 1;
  int nzp = (m.ezflt > 0) ? m.nz + 2 : m.nz + 1;
  g.nx = m.nx;
  g.nz = m.nz;
  g.ezflt = m.ezflt;
  g.npoin_fem = nxp * nzp;
  g.nelem = m.nx * m.nz;
  g.flat = true;
  g.coord_fem.assign((size_t)2 * g.npoin_fem, 0.0);
  g.knods.assign((size_t)4 * g.nelem, 0);
  g.tag.assign(g.nelem, 0);
  std::vector<double> x(nxp), z(nzp);
  for (int i = 0; i < nxp; ++i) x[i] = m.xmin + (m.xmax - m.xmin) / (double)m.nx * (double)i;
  if (m.ezflt > 0) {
    int k = 0;
    for (int j = 0; j <= m.ezflt; ++j) z[k++] = m.zmin + (m.zmax - m.zmin) / (double)m.nz * (double)j;
    for (int j = m.ezflt; j <= m.nz; ++j) z[k++] = m.zmin + (m.zmax - m.zmin) / (double)m.nz * (double)j;
  } else {
    for (int j = 0; j <= m.nz; ++j) z[j] = m.zmin + (m.zmax - m.zmin) / (double)m.nz * (double)j;
  }
  {
    size_t ilast = 0;
    for (int j = 0; j < nzp; ++j) {
      for (int i = 0; i < nxp; ++i) {
        g.coord_fem[2 * (ilast + i) + 0] = x[i];
        g.coord_fem[2 * (ilast + i) + 1] = z[j];
      }
      ilast += nxp;
    }
  }
  // domain tags
  for (const auto& d : m.domains)
    for (int i = d.ex[0]; i <= d.ex[1]; ++i)
      for (int j = d.ez[0]; j <= d.ez[1]; ++j) g.tag[sub2ind(i, j, m.nx) - 1] = d.tag;
  if (m.fztag > 0) {
    int j1 = std::max(m.ezflt + 1 - m.fznz, 1);
    int j2 = std::min(m.ezflt + m.fznz, m.nz);
    for (int j = j1; j <= j2; ++j)
      for (int i = 1; i <= m.nx; ++i) g.tag[sub2ind(i, j, m.nx) - 1] = m.fztag;
  }
  for (int e = 0; e < g.nelem; ++e)
    if (g.tag[e] == 0) IO_abort("CART_build: Domain tags not entirely set");
  // connectivity (mesh_structured.f90:24-35,79-81)
  {
    int k = 0;
    for (int j = 1; j <= m.nz; ++j)
      for (int i = 1; i <= m.nx; ++i) {
        g.knods[4 * k + 0] = sub2ind(i, j, nxp);
        g.knods[4 * k + 1] = sub2ind(i + 1, j, nxp);
        g.knods[4 * k + 2] = sub2ind(i + 1, j + 1, nxp);
        g.knods[4 * k + 3] = sub2ind(i, j + 1, nxp);
        ++k;
      }
    if (m.ezflt > 0)
      for (size_t q = (size_t)4 * m.nx * m.ezflt; q < g.knods.size(); ++q) g.knods[q] += nxp;
  }
  // boundaries (mesh_structured.f90:86-196)
  int splitN = 0;
  int nb = 4;
  if (m.ezflt > 0) {
    nb = 6;
  } else {
    if (m.split) splitN = (int)std::floor((m.splitD - m.xmin) / (m.xmax - m.xmin) * m.nx);
    if (splitN > 0) nb = 5;
  }
  g.bnds.assign(nb, Boundary());
  auto setb = [&](int slot, int tagv, int n, std::function<int(int)> el, int edge) {
    Boundary& b = g.bnds[slot - 1];
    b.tag = tagv;
    b.nelem = n;
    b.elem.resize(n);
    b.edge.assign(n, edge);
    for (int i = 1; i <= n; ++i) b.elem[i - 1] = el(i);
  };
  if (splitN > 0) {
    setb(5, 5, splitN, [&](int i) { return sub2ind(i, 1, m.nx); }, edge_D);
    setb(1, 1, m.nx - splitN, [&](int i) { return sub2ind(splitN + i, 1, m.nx); }, edge_D);
  } else {
    setb(1, 1, m.nx, [&](int i) { return sub2ind(i, 1, m.nx); }, edge_D);
  }
  setb(2, 2, m.nz, [&](int j) { return sub2ind(m.nx, j, m.nx); }, edge_R);
  setb(3, 3, m.nx, [&](int i) { return sub2ind(i, m.nz, m.nx); }, edge_U);
  setb(4, 4, m.nz, [&](int j) { return sub2ind(1, j, m.nx); }, edge_L);
  if (m.ezflt > 0) {
    setb(6, 6, m.nx, [&](int i) { return sub2ind(i, m.ezflt + 1, m.nx); }, edge_D);
    setb(5, 5, m.nx, [&](int i) { return sub2ind(i, m.ezflt, m.nx); }, edge_U);
  }
  // renumber (mesh_structured.f90:204-269, fem_grid.f90:505-519)
  g.perm.assign(g.nelem + 1, 0);
  for (int e = 1; e <= g.nelem; ++e) g.perm[e] = e;
  if (m.renumber) {
    std::vector<int> perm, perm_inv;
    rcmlib::structured_rcm(m.nx, m.nz, perm, perm_inv);
    std::vector<int> kn(g.knods.size()), tg(g.tag.size());
    for (int e = 1; e <= g.nelem; ++e) {
      for (int n = 0; n < 4; ++n) kn[4 * (e - 1) + n] = g.knods[4 * (size_t)(perm[e] - 1) + n];
      tg[e - 1] = g.tag[perm[e] - 1];
    }
    g.knods.swap(kn);
    g.tag.swap(tg);
    for (auto& b : g.bnds)
      for (auto& el : b.elem) el = perm_inv[el];
    g.perm = perm;
  }
}

// ------------------------------------------------------------------------------------------
// fem_grid.f90:103-204 FE_SetConnectivity
struct Connectivity {
  std::vector<int> vstart;       // CSR over control nodes (1-based node k -> [vstart[k],vstart[k+1]) )
  std::vector<int> velem, vnode;  // in the order LI_Remove_Head yields them (descending element)
  std::vector<int> edge_elem, edge_edge;  // (4,nelem)
};

inline void FE_SetConnectivity(const Grid& g, Connectivity& c) {
  int np = g.npoin_fem, ne = g.nelem;
  {
    std::vector<int> cnt(np + 2, 0);
    for (int e = 0; e < ne; ++e)
      for (int n = 0; n < 4; ++n) cnt[g.knods[4 * (size_t)e + n]]++;
    c.vstart.assign(np + 2, 0);
    for (int k = 1; k <= np; ++k) c.vstart[k + 1] = c.vstart[k] + cnt[k];
  }
  c.velem.assign(c.vstart[np + 1], 0);
  c.vnode.assign(c.vstart[np + 1], 0);
  {
    // head insertion while looping e=1..ne, n=1..4, then popping from the head:
    // resulting order = reverse insertion order
    std::vector<int> fill(np + 2, 0);
    for (int e = ne; e >= 1; --e)
      for (int n = 4; n >= 1; --n) {
        int k = g.knods[4 * (size_t)(e - 1) + (n - 1)];
        int p = c.vstart[k] + fill[k]++;
        c.velem[p] = e;
        c.vnode[p] = n;
      }
  }
  c.edge_elem.assign((size_t)4 * ne, 0);
  c.edge_edge.assign((size_t)4 * ne, 0);
  static const int EdgeKnod1[4] = {1, 2, 3, 4}, EdgeKnod2[4] = {2, 3, 4, 1};
  for (int e = 1; e <= ne; ++e)
    for (int n = 1; n <= 4; ++n) {
      if (c.edge_elem[4 * (size_t)(e - 1) + (n - 1)] > 0) continue;
      int k1 = g.knods[4 * (size_t)(e - 1) + EdgeKnod1[n - 1] - 1];
      int k2 = g.knods[4 * (size_t)(e - 1) + EdgeKnod2[n - 1] - 1];
      int nn = 0;
      for (int n1 = c.vstart[k1]; n1 < c.vstart[k1 + 1]; ++n1) {
        int ee = c.velem[n1];
        if (ee == e) continue;
        for (int n2 = c.vstart[k2]; n2 < c.vstart[k2 + 1]; ++n2) {
          if (c.velem[n2] == ee) {
            nn = c.vnode[n2];
            c.edge_elem[4 * (size_t)(e - 1) + (n - 1)] = ee;
            c.edge_edge[4 * (size_t)(e - 1) + (n - 1)] = nn;
            c.edge_elem[4 * (size_t)(ee - 1) + (nn - 1)] = e;
            c.edge_edge[4 * (size_t)(ee - 1) + (nn - 1)] = n;
            break;
          }
        }
        if (nn > 0) break;
      }
    }
}

// edge GLL index tables, counterclockwise (spec_grid.f90:876-883 SE_inquire)
inline void edge_tabs(int ngll, int edge, std::vector<int>& itab, std::vector<int>& jtab) {
  itab.resize(ngll);
  jtab.resize(ngll);
  for (int k = 1; k <= ngll; ++k) {
    switch (edge) {
      case edge_D: itab[k - 1] = k; jtab[k - 1] = 1; break;
      case edge_R: itab[k - 1] = ngll; jtab[k - 1] = k; break;
      case edge_U: itab[k - 1] = ngll + 1 - k; jtab[k - 1] = ngll; break;
      case edge_L: itab[k - 1] = 1; jtab[k - 1] = ngll + 1 - k; break;
    }
  }
}

// spec_grid.f90:149-190 SE_init_gll
inline void SE_init_gll(Grid& g) {
  int n = g.ngll;
  gll::get_GLL_info(n, g.xgll, g.wgll, g.H);
  g.Ht.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) g.Ht[i + (size_t)n * j] = g.H[j + (size_t)n * i];
  g.wgll2.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) g.wgll2[i + (size_t)n * j] = g.wgll[i] * g.wgll[j];
  g.shape.assign((size_t)4 * n * n, 0.0);
  g.dshape.assign((size_t)8 * n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      Q4_getshape(g.xgll[i], g.xgll[j], &g.shape[4 * (i + (size_t)n * j)]);
      Q4_getdershape(g.xgll[i], g.xgll[j], &g.dshape[8 * (i + (size_t)n * j)]);
    }
}

// spec_grid.f90:198-314 SE_init_numbering
inline void SE_init_numbering(Grid& g) {
  int ngll = g.ngll, ne = g.nelem;
  Connectivity c;
  FE_SetConnectivity(g, c);  // lazily built on first FE_GetEdgeConn, after RCM
  g.ibool.assign((size_t)ngll * ngll * ne, 0);
  std::vector<int> iedg[4], jedg[4];
  for (int n = 1; n <= 4; ++n) edge_tabs(ngll, n, iedg[n - 1], jedg[n - 1]);
  const int ivtx[4] = {1, ngll, ngll, 1}, jvtx[4] = {1, 1, ngll, ngll};
  int npoin = 0;
  for (int e = 1; e <= ne; ++e) {
    for (int j = 2; j <= ngll - 1; ++j)
      for (int i = 2; i <= ngll - 1; ++i) g.ib(i, j, e) = ++npoin;
    for (int n = 1; n <= 4; ++n) {
      if (g.ib(iedg[n - 1][1], jedg[n - 1][1], e) > 0) continue;
      int ee = c.edge_elem[4 * (size_t)(e - 1) + (n - 1)];
      int nn = c.edge_edge[4 * (size_t)(e - 1) + (n - 1)];
      for (int k = 2; k <= ngll - 1; ++k) {
        ++npoin;
        g.ib(iedg[n - 1][k - 1], jedg[n - 1][k - 1], e) = npoin;
        if (ee > 0) {
          // iedgR(k,nn) = iedg(ngll+1-k,nn)
          g.ib(iedg[nn - 1][ngll - k], jedg[nn - 1][ngll - k], ee) = npoin;
        }
      }
    }
    for (int n = 1; n <= 4; ++n) {
      int i = ivtx[n - 1], j = jvtx[n - 1];
      if (g.ib(i, j, e) > 0) continue;
      ++npoin;
      int k = g.knods[4 * (size_t)(e - 1) + (n - 1)];
      for (int q = c.vstart[k]; q < c.vstart[k + 1]; ++q) {
        int nn = c.vnode[q];
        g.ib(ivtx[nn - 1], jvtx[nn - 1], c.velem[q]) = npoin;
      }
    }
  }
  g.npoin = npoin;
}

// spec_grid.f90:321-345 SE_init_coord: coord(:,ibool) = matmul(coorg, shape); last writer wins
inline void SE_init_coord(Grid& g) {
  int n = g.ngll;
  g.coord.assign((size_t)2 * g.npoin, 0.0);
  for (int e = 1; e <= g.nelem; ++e) {
    double cg[8];
    for (int k = 0; k < 4; ++k) {
      int kn = g.knods[4 * (size_t)(e - 1) + k];
      cg[2 * k] = g.coord_fem[2 * (size_t)(kn - 1)];
      cg[2 * k + 1] = g.coord_fem[2 * (size_t)(kn - 1) + 1];
    }
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i) {
        const double* sh = &g.shape[4 * ((i - 1) + (size_t)n * (j - 1))];
        double x = 0.0, z = 0.0;
        for (int k = 0; k < 4; ++k) {
          x += cg[2 * k] * sh[k];
          z += cg[2 * k + 1] * sh[k];
        }
        int ip = g.ib(i, j, e);
        g.coord[2 * (size_t)(ip - 1)] = x;
        g.coord[2 * (size_t)(ip - 1) + 1] = z;
      }
  }
}

// spec_grid.f90:551-586 SE_Jacobian_eij: jac = matmul(coorg(2,4), dshape(4,2)) (col-major 2x2)
inline void SE_Jacobian(const Grid& g, int e, int i, int j, double jac[4]) {
  int n = g.ngll;
  const double* ds = &g.dshape[8 * ((i - 1) + (size_t)n * (j - 1))];
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < 2; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) {
        int kn = g.knods[4 * (size_t)(e - 1) + k];
        s += g.coord_fem[2 * (size_t)(kn - 1) + r] * ds[k + 4 * c];
      }
      jac[r + 2 * c] = s;
    }
}
inline double SE_VolumeWeight(const Grid& g, int e, int i, int j) {  // spec_grid.f90:608-627
  double jac[4];
  SE_Jacobian(g, e, i, j, jac);
  return (jac[0] * jac[3] - jac[2] * jac[1]) * g.wgll2[(i - 1) + (size_t)g.ngll * (j - 1)];
}

// spec_grid.f90:530-543 SE_get_edge_nodes
inline void SE_get_edge_nodes(const Grid& g, int element, int edge, std::vector<int>& nodes) {
  std::vector<int> it, jt;
  edge_tabs(g.ngll, edge, it, jt);
  nodes.resize(g.ngll);
  for (int k = 0; k < g.ngll; ++k) nodes[k] = g.ib(it[k], jt[k], element);
}

// spec_grid.f90:676-822 BC_set_bulk_node (+ utils.f90 drank = stable merge ranking)
inline void BC_set_bulk_node(Boundary& bc, const Grid& g) {
  if (!bc.exists()) return;
  int ngll = g.ngll;
  bc.ngnod = ngll;
  bc.ibool.assign((size_t)ngll * bc.nelem, 0);
  std::vector<int> nodes_list;                        // bc node id -> bulk node
  std::vector<std::pair<int, int>> corners;           // (bulk, bc) newest first is irrelevant
  std::vector<int> ev;
  for (int n = 1; n <= bc.nelem; ++n) {
    SE_get_edge_nodes(g, bc.elem[n - 1], bc.edge[n - 1], ev);
    for (int kloc = 1; kloc <= ngll; ++kloc) {
      int bulk = ev[kloc - 1];
      bool new_node = true;
      int bc_inode = 0;
      bool at_corner = (kloc == 1 || kloc == ngll);
      if (at_corner) {
        for (auto& cn : corners)
          if (cn.first == bulk) {
            new_node = false;
            bc_inode = cn.second;
            break;
          }
      }
      if (new_node) {
        nodes_list.push_back(bulk);
        bc_inode = (int)nodes_list.size();
        if (at_corner) {
          corners.push_back({bulk, bc_inode});
          // only the two most recent corners can match on a simple chain, but keep all
        }
      }
      bc.ibool[(kloc - 1) + (size_t)ngll * (n - 1)] = bc_inode;
    }
    // keep the corner list short (chain boundaries): matches can only involve recent corners,
    // but a closed loop could match the very first -> keep first + last few
    if (corners.size() > 8) corners.erase(corners.begin() + 1, corners.end() - 4);
  }
  bc.npoin = (int)nodes_list.size();
  // sort by the coordinate with the largest range
  double xmn = HUGE_D, xmx = -HUGE_D, zmn = HUGE_D, zmx = -HUGE_D;
  for (int k = 0; k < bc.npoin; ++k) {
    double x = g.coord[2 * (size_t)(nodes_list[k] - 1)], z = g.coord[2 * (size_t)(nodes_list[k] - 1) + 1];
    xmn = std::min(xmn, x);
    xmx = std::max(xmx, x);
    zmn = std::min(zmn, z);
    zmx = std::max(zmx, z);
  }
  double Lx = xmx - xmn, Lz = zmx - zmn;
  int dim = (Lx > Lz) ? 0 : 1;
  std::vector<int> isort(bc.npoin);
  std::iota(isort.begin(), isort.end(), 0);
  std::stable_sort(isort.begin(), isort.end(), [&](int a, int b) {
    return g.coord[2 * (size_t)(nodes_list[a] - 1) + dim] < g.coord[2 * (size_t)(nodes_list[b] - 1) + dim];
  });
  bc.node.resize(bc.npoin);
  std::vector<int> iback(bc.npoin);
  for (int k = 0; k < bc.npoin; ++k) {
    bc.node[k] = nodes_list[isort[k]];
    iback[isort[k]] = k + 1;
  }
  for (auto& v : bc.ibool) v = iback[v - 1];
}

// spec_grid.f90:961-1011 BC_get_normal_and_weights
inline void BC_get_normal_and_weights(const Boundary& bc, const Grid& g, std::vector<double>& NORM /*(npoin,2)*/,
                                      std::vector<double>& W, bool periodic) {
  int np = bc.npoin, ngll = g.ngll;
  NORM.assign((size_t)2 * np, 0.0);
  W.assign(np, 0.0);
  std::vector<int> it, jt;
  for (int be = 1; be <= bc.nelem; ++be) {
    int edge = bc.edge[be - 1];
    edge_tabs(ngll, edge, it, jt);
    int dim_t = (edge == edge_D || edge == edge_U) ? 1 : 2;
    double SignTang = (edge == edge_U || edge == edge_L) ? -1.0 : 1.0;
    for (int k = 1; k <= ngll; ++k) {
      double jac[4];
      SE_Jacobian(g, bc.elem[be - 1], it[k - 1], jt[k - 1], jac);
      double dx = jac[0 + 2 * (dim_t - 1)], dz = jac[1 + 2 * (dim_t - 1)];
      double Jac1D = std::sqrt(dx * dx + dz * dz);
      double t1 = SignTang * dx / Jac1D, t2 = SignTang * dz / Jac1D;
      int bn = bc.ibool[(k - 1) + (size_t)ngll * (be - 1)];
      NORM[(bn - 1)] = NORM[(bn - 1)] + t2;
      NORM[(bn - 1) + np] = NORM[(bn - 1) + np] + (-t1);
      W[bn - 1] = W[bn - 1] + g.wgll[k - 1] * Jac1D;
    }
  }
  if (periodic) {
    for (int c = 0; c < 2; ++c) {
      NORM[0 + np * c] = NORM[0 + np * c] + NORM[(np - 1) + np * c];
      NORM[(np - 1) + np * c] = NORM[0 + np * c];
    }
    W[0] = W[0] + W[np - 1];
    W[np - 1] = W[0];
  }
  for (int be = 1; be <= bc.nelem; ++be) {
    for (int kk = 0; kk < 2; ++kk) {
      int bn = bc.ibool[((kk == 0) ? 0 : ngll - 1) + (size_t)ngll * (be - 1)];
      double a = NORM[bn - 1], b = NORM[bn - 1 + np];
      double nrm = std::sqrt(a * a + b * b);
      NORM[bn - 1] = a / nrm;
      NORM[bn - 1 + np] = b / nrm;
    }
  }
}

// spec_grid.f90:411-434 SE_find_nearest_node (ties -> highest node id)
inline int SE_find_nearest_node(const Grid& g, double x, double z, double* dist = nullptr) {
  int iglob = 0;
  double d2min = HUGE_D;
  for (int ip = 1; ip <= g.npoin; ++ip) {
    double dx = x - g.coord[2 * (size_t)(ip - 1)], dz = z - g.coord[2 * (size_t)(ip - 1) + 1];
    double d2 = dx * dx + dz * dz;
    if (d2 <= d2min) {
      d2min = d2;
      iglob = ip;
    }
  }
  if (dist) *dist = std::sqrt(d2min);
  return iglob;
}

// ------------------------------------------------------------------------------------------
// materials
struct MatInput {  // matpro_input_type (prop_mat.f90:21-25) for ELAST (+KV)
  bool elastic = false, isotropic = false, homogeneous = false, kv = false;
  bool plastic = false;                    // kind='PLAST' (mat_plastic.f90)
  double phi = 0, coh = 0, Tv = 0, e0[3] = {0, 0, 0};
  bool damage = false;                     // kind='DMG' (mat_damage.f90): phi, e0 as above; Cd, R, beta, alpha, ep
  double Cd = 0, Rdmg = 0, beta_dmg = 0, alpha0 = 0, ep0[3] = {0, 0, 0};
  bool visco = false;                      // kind='VISCO' (mat_visco.f90): generalized Maxwell body, Nbody mechanisms
  double QP = 0, QS = 0, fmin = 0, fmax = 0;
  int Nbody = 0;
  std::vector<double> theta, wbody;        // theta(Nbody,3) column-major, wbody(Nbody): get_attenuation (mat_visco.f90:251-340)
  Dist rho, cp, cs, eta;
  double lambda = 0, mu = 0;  // set if homogeneous (mat_elastic.f90:118-125)
  bool has_lambda = false;
  bool etaxdt = true;
  bool synthetic = false;  // hash-based heterogeneous model (not in the reference)
  uint64_t seed = 0;
  int64_t ix0 = 0, iz0 = 0;  // lattice origin of this mesh inside a larger synthetic mesh (window tests)
};

struct ElemProp {  // prop_elem_type (prop_elem.f90:10-14): homogeneous scalar or ngll x ngll values
  double homo = 0.0;
  int64_t hete = -1;  // offset into Materials::pool
};

struct Materials {
  std::vector<MatInput> inputs;  // by tag (1-based -> index tag-1)
  std::vector<double> pool;
  std::vector<ElemProp> rho, cp, cs, lambda, mu, eta;  // per element
  int ngll = 0;
  double get(const std::vector<ElemProp>& p, int e, int i, int j) const {  // PROP_get_ij
    const ElemProp& q = p[e - 1];
    if (q.hete >= 0) return pool[q.hete + (i - 1) + (size_t)ngll * (j - 1)];
    return q.homo;
  }
  void get(const std::vector<ElemProp>& p, int e, double* out) const {  // PROP_get (ngll,ngll)
    const ElemProp& q = p[e - 1];
    int n2 = ngll * ngll;
    if (q.hete >= 0)
      for (int k = 0; k < n2; ++k) out[k] = pool[q.hete + k];
    else
      for (int k = 0; k < n2; ++k) out[k] = q.homo;
  }
  ElemProp set_vals(const double* v) {
    ElemProp q;
    q.homo = 0.0;
    q.hete = (int64_t)pool.size();
    pool.insert(pool.end(), v, v + (size_t)ngll * ngll);
    return q;
  }
};

// ------------------------------------------------------------------------------------------
struct TimeScheme {  // timescheme_type (time.f90:5-11)
  std::string kind = "leapfrog";
  double dt = 0, courant = 0.5, time = 0, total = 0, alpha = 1.0, beta = 0.0, gamma = 0.5, Omega_max = 2.0;
  int nt = 0;
  int nstages = 0;              // symplectic schemes (time.f90:248-300)
  std::vector<double> a, b;     // time%a(1:nstages+1), time%b(1:nstages)
  double CoefA2D() const {  // time.f90:426-440
    if (kind == "newmark" || kind == "HHT-alpha") return beta * dt * dt;
    return 0.0;
  }
  double CoefA2V() const {  // time.f90:443-456
    if (kind == "newmark" || kind == "HHT-alpha") return gamma * dt;
    return dt;
  }
  double CoefA2Vrhs() const {  // time.f90:465-486
    if (kind == "newmark" || kind == "HHT-alpha") return alpha * CoefA2V();
    return 0.5 * CoefA2V();
  }
};

// ------------------------------------------------------------------------------------------
// boundary conditions
struct BcAbso {  // bc_abso_type (bc_abso.f90:38-46)
  const Boundary* topo = nullptr;
  std::vector<double> C;  // (npoin,ndof)
  std::vector<double> K;  // (ngll,ndof,nelem)
  std::vector<double> n;  // (npoin,2)
  bool stacey = false, periodic = false, is_flat = true, let_wave = true;
};
struct BcDirneu {  // bc_dirneu_type (bc_dirneu.f90:17-23)
  const Boundary* topo = nullptr;
  int kind[2] = {1, 1};  // 1 Neumann, 2 Dirichlet
};
struct Swf {  // swf_type (bc_dynflt_swf.f90:12-20)
  int kind = 1;
  double dt = 0;
  bool healing = false;
  Dist in_dc, in_mus, in_mud, in_alpha, in_p;
  std::vector<double> dc, mus, mud, theta, p, alpha;
};
struct Rsf {  // rsf_type (bc_dynflt_rsf.f90:14-22)
  int kind = 1;
  double dt = 0;
  Dist in_dc, in_mus, in_a, in_b, in_Vstar, in_theta, in_Vc;
  std::vector<double> dc, mus, a, b, Vstar, theta, Vc, Tc, coeft;
};
struct Twf {  // twf_type (bc_dynflt_twf.f90:12-16)
  int kind = 1;
  double X = 0, Z = 0, mus = 0.6, mud = 0.5, mu0 = 0.6, L = 1, V = 1e3, T = HUGE_D, Dc = HUGE_D;
};
struct NormalLaw {  // normal_type (bc_dynflt_normal.f90:8-13)
  int kind = 1;
  std::vector<double> sigma;
  double T = 1, L = 1, V = 1, coef = 0;
};
struct BcDynflt {  // bc_dynflt_type (bc_dynflt.f90:18-38)
  int tags[2] = {0, 0};
  int npoin = 0;
  std::vector<int> node1, node2;
  bool two_sides = false;
  double CoefA2V = 0, CoefA2D = 0;
  std::vector<double> n1, B, invM1, invM2, Z, T0, Tstick, T, V, D, coord;  // (npoin,*) col-major
  std::vector<double> MU, cohesion;
  std::unique_ptr<Swf> swf;
  std::unique_ptr<Rsf> rsf;
  std::unique_ptr<Twf> twf;
  bool allow_opening = true;
  NormalLaw normal;
  const Boundary *bc1 = nullptr, *bc2 = nullptr;
  Dist in_T, in_N, in_Sxx, in_Sxy, in_Sxz, in_Syz, in_Szz, in_cohesion, in_V;
  double ot1 = 0, odt = 0;
  int oit = 0, oitd = 1, oix1 = 1, oixn = std::numeric_limits<int>::max(), oixd = 1;
  bool osides = false;
  // recorded outputs (what BC_DYNFLT_write would put in FltXX_sem2d.dat / _potency_sem2d.tab)
  std::vector<float> out;       // records: per output time 6 x onx floats
  std::vector<double> potency;  // per call: 2*(ndof+1) doubles
  int nout = 0;
  int onx() const { return (oixn - oix1) / oixd + 1; }
};

enum BcKind { IS_EMPTY = 0, IS_DIRNEU = 1, IS_KINFLT = 2, IS_ABSORB = 3, IS_PERIOD = 4, IS_LISFLT = 5, IS_DYNFLT = 6 };
struct BcPerio {  // bc_periodic_type (bc_periodic.f90:11-14)
  const Boundary* master = nullptr;
  const Boundary* slave = nullptr;
};
// bc_periodic.f90:107-121 BC_PERIO_intersects
inline bool BC_PERIO_intersects(const Boundary& bnd, const BcPerio* perio) {
  if (!perio) return false;
  auto on = [&](int node) {
    for (int v : perio->master->node)
      if (v == node) return true;
    for (int v : perio->slave->node)
      if (v == node) return true;
    return false;
  };
  return on(bnd.node[0]) && on(bnd.node[bnd.npoin - 1]);
}
struct Bc {  // bc_type (bc_gen.f90:29-41)
  int tag[2] = {0, 0};
  int kind = IS_EMPTY;
  std::unique_ptr<BcPerio> perio;
  std::unique_ptr<BcAbso> abso;
  std::unique_ptr<BcDirneu> dirneu;
  std::unique_ptr<BcDynflt> dynflt;
};

// ------------------------------------------------------------------------------------------
struct Ricker {  // stf_ricker.f90:12-15
  double f0 = 0, t0 = 0, ampli = 1;
  double eval(double t) const {  // :89-101
    double arg = PI * f0 * (t - t0);
    arg = arg * arg;
    return -ampli * (1.0 - 2.0 * arg) * std::exp(-arg);
  }
};
struct Source {  // source_type (src_gen.f90:20-26) with FORCE mechanism (src_force.f90)
  double coord[2] = {0, 0};
  double tdelay = 0, ampli = 1;
  Ricker stf;
  double dir[2] = {0, 1};
  int iglob = 0;
  // so_moment_type (src_moment.f90:9-14); moment == false: collocated force
  bool moment = false;
  double M[4] = {0, 0, 0, 0};              // M(2,ndof) col-major
  std::vector<int> mnode;                  // terms of SRC_MOMENT_add in application order
  std::vector<double> mcoef;               // (nterms, ndof) col-major
};
struct Receivers {  // rec_type (receivers.f90:9-20)
  bool present = false;
  int nx = 0, nt = 0, isamp = 1;
  bool AtNode = true;
  char field = 'V';
  std::vector<double> coord;    // (2,nx)
  std::vector<int> iglob;       // (nx)
  std::vector<double> interp;   // (ngll*ngll,nx)
  std::vector<int> einterp;     // (nx)
  std::vector<float> sis;       // (nt,nx,ndof)
  double tsamp = 0;
};

// ------------------------------------------------------------------------------------------
struct Problem {  // problem_type (problem_class.f90:19-46)
  Grid grid;
  Materials mat;
  TimeScheme time;
  int ndof = 2;
  // work arrays: coefficient sets (mat_gen.f90:357-365 shares one set per homogeneous tag)
  int nelast = 0;
  std::vector<double> a;          // (ngll,ngll,nelast,ncoefsets)
  std::vector<double> beta25d;    // (ngll,ngll,ncoefsets) matwrk_elast_type%beta when W is finite (mat_elastic.f90:280-284), else empty
  std::vector<int> elem2set;      // (nelem) 1-based set id
  int ncoefsets = 0;
  std::vector<int> kv_elem;       // 1-based element ids with KV
  std::vector<int> elem2kv;       // (nelem) 0 or 1-based index into kv list
  std::vector<double> kv_eta;     // (ngll,ngll,nkv) already multiplied by dt if ETAxDT
  // Coulomb plasticity (matwrk_plast_type, mat_plastic.f90:10-19) + derint (mat_gen.f90:46-49), per plastic element
  std::vector<int> elem2pl;       // (nelem) 0 or 1-based index into the plastic element list
  std::vector<int> pl_elem;       // 1-based element ids
  std::vector<double> pl_par;     // (10,npl): lambda, mu, yield_co, yield_mu, vp_factor, e0(3), (unused 2)
  std::vector<double> pl_ep;      // (ngll,ngll,3,npl) plastic strain
  std::vector<double> pl_derint;  // (ngll,ngll,5,npl): dxi_dx, dxi_dy, deta_dx, deta_dy, weights
  std::vector<double> pl_beta;    // (ngll,ngll,npl) when W is finite
  // damage rheology (matwrk_dmg_type, mat_damage.f90:44-52), per damage element
  std::vector<int> elem2dm;       // (nelem) 0 or 1-based index into the damage element list
  std::vector<int> dm_elem;
  std::vector<double> dm_derint;  // (ngll,ngll,5,ndm)
  std::vector<double> dm_par;     // (16,ndm): lambda, mu, xi_0, gamma_r, beta, Cd, Cv, e0(3), s0(3)
  std::vector<double> dm_state;   // (ngll,ngll,4,ndm): alpha, ep(3)
  // visco-elasticity (matwrk_visco_type, mat_visco.f90:11-19), per visco element; derint shares pl_derint's layout
  std::vector<int> elem2vs;       // (nelem) 0 or 1-based index into the visco element list
  std::vector<int> vs_elem;       // 1-based element ids
  std::vector<double> vs_derint;  // (ngll,ngll,5,nvs)
  std::vector<double> vs_el;      // (ngll,ngll,Nbody,3) per element, concatenated (offsets vs_off)
  std::vector<size_t> vs_off;
  std::vector<double> vs_etot;    // (ngll,ngll,3,nvs) strain of the previous evaluation
  std::vector<double> rmass;      // (npoin,ndof) -- mass until init end, then inverse
  std::vector<double> mass;       // (npoin) assembled mass as MAT_MASS_init leaves it (mat_mass.f90:50-57), before BC_init
  std::vector<double> d, v, a_;   // fields (npoin,ndof) col-major
  std::vector<Bc> bc;
  const BcPerio* perio = nullptr;  // the periodic boundary the other BC_*_init routines receive (bc_gen.f90:221-246)
  std::vector<Source> src;
  Receivers rec;
  int it = 0;
  double grid_cfl = 0;
  // energy (energy.f90) -- optional
  double E_k = 0;
  bool kd_force_kd1 = false;  // testing hook: use the KD1 form even if ngll==5

  size_t idx(int ip, int c) const { return (size_t)(ip - 1) + (size_t)grid.npoin * c; }
};

// mat_damage.f90:453-491 compute_stress: sigma = (lambda i1 - gamma sqrt(i2)) delta + (2 mu - gamma i1 / sqrt(i2)) e,
// with the loss-of-convexity checks of the reference (they abort the run)
inline void DMG_compute_stress(double s[3], const double e[3], double rl, double rm, double rg, double& i1, double& i2, double& xi) {
  i1 = e[0] + e[1];
  i2 = e[0] * e[0] + e[1] * e[1] + 2.0 * e[2] * e[2];
  const double si2 = std::sqrt(i2);
  xi = si2 < 1e-10 ? 0.0 : i1 / si2;
  const double two_mue = 2.0 * rm - rg * xi;
  s[0] = rl * i1 - rg * si2 + two_mue * e[0];
  s[1] = rl * i1 - rg * si2 + two_mue * e[1];
  s[2] = two_mue * e[2];
  const double p = -(4.0 * rm + 2.0 * rl - 3.0 * rg * xi);
  const double q = two_mue * two_mue + two_mue * (2.0 * rl - rg * xi) + rg * (rl * xi - rg) * (2.0 - xi * xi);
  const double d = p * p / 4.0 - q;
  if (d <= 0.0) IO_abort("mat_damage:elastic: discriminant < 0");
  if (p / 2.0 + std::sqrt(d) >= 0.0) IO_abort("MAT_DMG: damage exceeded critical value (1st type)");
  if (two_mue <= 0.0) IO_abort("MAT_DMG: damage exceeded critical value (2nd type)");
}

// ------------------------------------------------------------------------------------------
// Least squares x = argmin |A x - b|, A (m,n) column-major, m >= n, full column rank.  The reference solves it with
// Numerical Recipes' svdcmp / svbksb (mat_visco.f90:343-607), i.e. x = V diag(1/w) U^T b; the solution is unique, and
// here it comes from a one-sided Jacobi SVD (Hestenes) -- same x to rounding (cond(A) ~ 1e2), not the same code.
inline void lsq_svd(std::vector<double> A, int m, int n, const std::vector<double>& b, std::vector<double>& x) {
  std::vector<double> V((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[i + (size_t)n * i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double app = 0, aqq = 0, apq = 0;
        for (int k = 0; k < m; ++k) {
          app += A[k + (size_t)m * p] * A[k + (size_t)m * p];
          aqq += A[k + (size_t)m * q] * A[k + (size_t)m * q];
          apq += A[k + (size_t)m * p] * A[k + (size_t)m * q];
        }
        if (std::abs(apq) <= 1e-300 || std::abs(apq) <= 1e-17 * std::sqrt(app * aqq)) continue;
        off = std::max(off, std::abs(apq) / std::sqrt(app * aqq));
        const double zeta = (aqq - app) / (2.0 * apq);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (int k = 0; k < m; ++k) {
          const double ap = A[k + (size_t)m * p], aq = A[k + (size_t)m * q];
          A[k + (size_t)m * p] = c * ap - sn * aq;
          A[k + (size_t)m * q] = sn * ap + c * aq;
        }
        for (int k = 0; k < n; ++k) {
          const double vp = V[k + (size_t)n * p], vq = V[k + (size_t)n * q];
          V[k + (size_t)n * p] = c * vp - sn * vq;
          V[k + (size_t)n * q] = sn * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  // columns of A are now U diag(w): x = V diag(1/w) U^T b = sum_j V(:,j) (A(:,j)^T b) / w_j^2
  x.assign(n, 0.0);
  for (int j = 0; j < n; ++j) {
    double w2 = 0, ub = 0;
    for (int k = 0; k < m; ++k) {
      w2 += A[k + (size_t)m * j] * A[k + (size_t)m * j];
      ub += A[k + (size_t)m * j] * b[k];
    }
    if (w2 == 0.0) continue;  // svbksb skips zero singular values
    for (int i = 0; i < n; ++i) x[i] += V[i + (size_t)n * j] * (ub / w2);
  }
}

// mat_visco.f90:251-340 get_attenuation: relaxation frequencies log-spaced in [fmin, fmax], anelastic coefficients
// Y_alpha, Y_beta fitted to constant 1/QP, 1/QS at 2 Nbody - 1 frequencies, unrelaxed moduli, theta(Nbody,3)
inline void get_attenuation(std::vector<double>& theta, std::vector<double>& wbody, double& mu_inf, double& lambda_inf,
                            double cp, double cs, double rho, double QP, double QS, int Nbody, double fmin, double fmax) {
  const int Nf = 2 * Nbody - 1;
  const double w0 = 2.0 * PI * std::pow(fmin * fmax, 0.5), wmin = 2.0 * PI * fmin, wmax = 2.0 * PI * fmax;
  std::vector<double> w(Nf);
  if (Nbody > 1)
    for (int i = 1; i <= Nf; ++i) w[i - 1] = std::exp(std::log(wmin) + (i - 1) * (std::log(wmax) - std::log(wmin)) / (Nf - 1));
  else
    for (int i = 0; i < Nf; ++i) w[i] = w0;
  wbody.assign(Nbody, 0.0);
  for (int j = 1; j <= Nbody; ++j) wbody[j - 1] = w[2 * j - 2];
  std::vector<double> AP((size_t)Nf * Nbody), AS((size_t)Nf * Nbody), qpi(Nf, 1.0 / QP), qsi(Nf, 1.0 / QS), Ya, Yb;
  for (int i = 0; i < Nf; ++i)
    for (int j = 0; j < Nbody; ++j) {
      AP[i + (size_t)Nf * j] = (wbody[j] * w[i] + wbody[j] * wbody[j] / QP) / (wbody[j] * wbody[j] + w[i] * w[i]);
      AS[i + (size_t)Nf * j] = (wbody[j] * w[i] + wbody[j] * wbody[j] / QS) / (wbody[j] * wbody[j] + w[i] * w[i]);
    }
  lsq_svd(AP, Nf, Nbody, qpi, Ya);
  lsq_svd(AS, Nf, Nbody, qsi, Yb);
  double RP1 = 1, RP2 = 0, RS1 = 1, RS2 = 0;
  for (int j = 0; j < Nbody; ++j) {
    const double r = w0 / wbody[j], den = 1.0 + r * r;
    RP1 = RP1 - Ya[j] / den;
    RP2 = RP2 + Ya[j] * r / den;
    RS1 = RS1 - Yb[j] / den;
    RS2 = RS2 + Yb[j] * r / den;
  }
  const double RP = std::sqrt(RP1 * RP1 + RP2 * RP2), RS = std::sqrt(RS1 * RS1 + RS2 * RS2);
  const double mu = rho * cs * cs, lambda = rho * (cp * cp - 2.0 * cs * cs);
  mu_inf = mu * (RS + RS1) / (2 * RS * RS);
  lambda_inf = (lambda + 2.0 * mu) * (RP + RP1) / (2 * RP * RP) - 2.0 * mu_inf;
  theta.assign((size_t)Nbody * 3, 0.0);
  for (int j = 0; j < Nbody; ++j) {
    theta[j] = (lambda_inf + 2.0 * mu_inf) * Ya[j];
    theta[j + Nbody] = (lambda_inf + 2.0 * mu_inf) * Ya[j] - 2.0 * mu_inf * Yb[j];
    theta[j + 2 * Nbody] = 2.0 * mu_inf * Yb[j];
  }
}

// ------------------------------------------------------------------------------------------
// material init
// mat_gen.f90:204-250 MAT_init_prop + mat_mass.f90:19-26 + mat_elastic.f90:189-236 + mat_kelvin_voigt.f90:117-124
inline void MAT_init_prop(Problem& pb, int N_for_lattice /*ngll*/) {
  Grid& g = pb.grid;
  Materials& m = pb.mat;
  int n = g.ngll, ne = g.nelem, n2 = n * n;
  m.ngll = n;
  m.rho.assign(ne, ElemProp());
  m.cp.assign(ne, ElemProp());
  m.cs.assign(ne, ElemProp());
  m.lambda.assign(ne, ElemProp());
  m.mu.assign(ne, ElemProp());
  m.eta.assign(ne, ElemProp());
  std::vector<double> ex(n2), ez(n2), buf(n2), rho(n2), cp(n2), cs(n2), tmp(n2);
  for (int e = 1; e <= ne; ++e) {
    int tag = g.tag[e - 1];
    if (tag > (int)m.inputs.size() || tag < 1)
      IO_abort("ELAST_init: element tag does not correspond to a material number");
    const MatInput& in = m.inputs[tag - 1];
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i) {
        int ip = g.ib(i, j, e);
        ex[(i - 1) + n * (j - 1)] = g.coord[2 * (size_t)(ip - 1)];
        ez[(i - 1) + n * (j - 1)] = g.coord[2 * (size_t)(ip - 1) + 1];
      }
    auto set_from_input = [&](const Dist& dd) -> ElemProp {  // PROP_set_cd1 (prop_elem.f90:33-52)
      ElemProp q;
      if (dd.is_dist()) {
        for (int k = 0; k < n2; ++k) buf[k] = dd.eval(ex[k], ez[k]);
        q = m.set_vals(buf.data());
      } else {
        q.homo = dd.c;
      }
      return q;
    };
    if (in.synthetic) {
      // lattice coordinates of the element's GLL points (pre-RCM element position)
      int eold = g.perm[e];
      int ie = (eold - 1) % g.nx, je = (eold - 1) / g.nx;
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
          uint64_t ix = (uint64_t)(in.ix0 + (int64_t)ie * (N_for_lattice - 1) + i), iz = (uint64_t)(in.iz0 + (int64_t)je * (N_for_lattice - 1) + j);
          double u1 = hash_u(in.seed, ix, iz, 1), u2 = hash_u(in.seed, ix, iz, 2), u3 = hash_u(in.seed, ix, iz, 3);
          double csv = 3464.0 * (1.0 + 0.10 * u1);
          double cpv = 1.7321 * csv * (1.0 + 0.02 * u2);
          double rhov = 2670.0 * (1.0 + 0.05 * u3);
          cs[i + n * j] = csv;
          cp[i + n * j] = cpv;
          rho[i + n * j] = rhov;
        }
      m.rho[e - 1] = m.set_vals(rho.data());
      m.cp[e - 1] = m.set_vals(cp.data());
      m.cs[e - 1] = m.set_vals(cs.data());
      for (int k = 0; k < n2; ++k) tmp[k] = rho[k] * (cp[k] * cp[k] - 2.0 * cs[k] * cs[k]);
      m.lambda[e - 1] = m.set_vals(tmp.data());
      for (int k = 0; k < n2; ++k) tmp[k] = rho[k] * cs[k] * cs[k];
      m.mu[e - 1] = m.set_vals(tmp.data());
      continue;
    }
    m.rho[e - 1] = set_from_input(in.rho);  // MAT_MASS_init_elem_prop
    if (in.elastic) {
      m.cp[e - 1] = set_from_input(in.cp);
      m.cs[e - 1] = set_from_input(in.cs);
      if (in.has_lambda) {
        m.lambda[e - 1].homo = in.lambda;
        m.mu[e - 1].homo = in.mu;
      } else {
        m.get(m.rho, e, rho.data());
        m.get(m.cp, e, cp.data());
        m.get(m.cs, e, cs.data());
        for (int k = 0; k < n2; ++k) tmp[k] = rho[k] * (cp[k] * cp[k] - 2.0 * cs[k] * cs[k]);
        m.lambda[e - 1] = m.set_vals(tmp.data());
        for (int k = 0; k < n2; ++k) tmp[k] = rho[k] * cs[k] * cs[k];
        m.mu[e - 1] = m.set_vals(tmp.data());
      }
    }
    if (in.visco) {  // MAT_VISCO_init_elem_prop (mat_visco.f90:116-161): unrelaxed moduli from get_attenuation
      m.cp[e - 1] = set_from_input(in.cp);
      m.cs[e - 1] = set_from_input(in.cs);
      MatInput& iw = m.inputs[tag - 1];
      if (iw.theta.empty()) {
        double mu_inf, la_inf;
        get_attenuation(iw.theta, iw.wbody, mu_inf, la_inf, in.cp.c, in.cs.c, in.rho.c, in.QP, in.QS, in.Nbody, in.fmin, in.fmax);
        iw.lambda = la_inf;
        iw.mu = mu_inf;
      }
      m.lambda[e - 1].homo = iw.lambda;
      m.mu[e - 1].homo = iw.mu;
    }
    if (in.plastic || in.damage) {  // MAT_PLAST_init_elem_prop (mat_plastic.f90:121-145) / MAT_DMG_init_elem_prop (mat_damage.f90:176-206): scalar properties
      m.cp[e - 1] = set_from_input(in.cp);
      m.cs[e - 1] = set_from_input(in.cs);
      const double rho1 = in.rho.c, cp1 = in.cp.c, cs1 = in.cs.c;
      m.lambda[e - 1].homo = rho1 * (cp1 * cp1 - 2.0 * cs1 * cs1);
      m.mu[e - 1].homo = rho1 * cs1 * cs1;
    }
    if (in.kv) m.eta[e - 1] = set_from_input(in.eta);
  }
}

// init.f90:145-289 CHECK_grid -> max_c_dx
inline double CHECK_grid(const Problem& pb) {
  const Grid& g = pb.grid;
  int n = g.ngll;
  double max_c_dx = 0.0;
  std::vector<double> celem((size_t)n * n);
  for (int e = 1; e <= g.nelem; ++e) {
    if (pb.ndof == 2)
      pb.mat.get(pb.mat.cp, e, celem.data());
    else
      pb.mat.get(pb.mat.cs, e, celem.data());
    double ratiomax = 0.0;
    for (int j = 1; j <= n - 1; ++j)
      for (int i = 1; i <= n - 1; ++i) {
        int p0 = g.ib(i, j, e), p1 = g.ib(i + 1, j, e), p2 = g.ib(i, j + 1, e);
        double x0 = g.coord[2 * (size_t)(p0 - 1)], z0 = g.coord[2 * (size_t)(p0 - 1) + 1];
        double x1 = g.coord[2 * (size_t)(p1 - 1)], z1 = g.coord[2 * (size_t)(p1 - 1) + 1];
        double x2 = g.coord[2 * (size_t)(p2 - 1)], z2 = g.coord[2 * (size_t)(p2 - 1) + 1];
        double rdist1 = std::sqrt((x1 - x0) * (x1 - x0) + (z1 - z0) * (z1 - z0));
        double rdist2 = std::sqrt((x2 - x0) * (x2 - x0) + (z2 - z0) * (z2 - z0));
        ratiomax = std::max(ratiomax, celem[(i - 1) + (size_t)n * (j - 1)] / std::min(rdist1, rdist2));
      }
    max_c_dx = std::max(max_c_dx, ratiomax);
  }
  return max_c_dx;
}

// time.f90:323-341 TIME_init
inline void TIME_init(TimeScheme& t, double grid_cfl) {
  if (t.dt > 0.0) {
    t.courant = grid_cfl * t.dt;
  } else {
    t.dt = t.courant / grid_cfl;
    if (t.total > 0.0) t.nt = (int)std::ceil(t.total / t.dt);
    t.total = t.nt * t.dt;
  }
}

// mat_elastic.f90:290-360 MAT_ELAST_init_a for one element -> a(ngll,ngll,nelast)
inline void MAT_ELAST_init_a(const Problem& pb, int e, int nelast, double* a) {
  const Grid& g = pb.grid;
  const Materials& m = pb.mat;
  int n = g.ngll, n2 = n * n;
  std::vector<double> la(n2), mu(n2);
  m.get(m.lambda, e, la.data());
  m.get(m.mu, e, mu.data());
  for (int j = 1; j <= n; ++j)
    for (int i = 1; i <= n; ++i) {
      int k = (i - 1) + n * (j - 1);
      double jac[4], ji[4];
      SE_Jacobian(g, e, i, j, jac);
      invert2(jac, ji);
      double DxiDx = ji[0], DetaDx = ji[1], DxiDz = ji[2], DetaDz = ji[3];
      double weights = (jac[0] * jac[3] - jac[2] * jac[1]) * g.wgll2[k];  // SE_VolumeWeights_e
      double mux = mu[k], muz = mu[k];
      double Kx = la[k] + 2.0 * mu[k], Kz = Kx;
      double av[10];
      switch (nelast) {
        case 2:
          av[0] = mux * DxiDx * DxiDx;
          av[1] = muz * DetaDz * DetaDz;
          break;
        case 3:
          av[0] = mux * DxiDx * DxiDx + muz * DxiDz * DxiDz;
          av[1] = mux * DetaDx * DetaDx + muz * DetaDz * DetaDz;
          av[2] = mux * DxiDx * DetaDx + muz * DxiDz * DetaDz;
          break;
        case 6:
          av[0] = Kx * DxiDx * DxiDx;
          av[1] = la[k] * DxiDx * DetaDz;
          av[2] = Kz * DetaDz * DetaDz;
          av[3] = mu[k] * DetaDz * DetaDz;
          av[4] = mu[k] * DxiDx * DetaDz;
          av[5] = mu[k] * DxiDx * DxiDx;
          break;
        case 10:
          av[0] = Kx * DxiDx * DxiDx + mu[k] * DxiDz * DxiDz;
          av[1] = la[k] * DxiDx * DetaDz + mu[k] * DxiDz * DetaDx;
          av[2] = Kz * DetaDz * DetaDz + mu[k] * DetaDx * DetaDx;
          av[3] = Kx * DetaDx * DetaDx + mu[k] * DetaDz * DetaDz;
          av[4] = la[k] * DxiDz * DetaDx + mu[k] * DxiDx * DetaDz;
          av[5] = Kz * DxiDz * DxiDz + mu[k] * DxiDx * DxiDx;
          av[6] = Kx * DxiDx * DetaDx + mu[k] * DxiDz * DetaDz;
          av[7] = (la[k] + mu[k]) * DxiDx * DxiDz;
          av[8] = (la[k] + mu[k]) * DetaDx * DetaDz;
          av[9] = Kz * DxiDz * DetaDz + mu[k] * DxiDx * DetaDx;
          break;
      }
      for (int q = 0; q < nelast; ++q) a[k + (size_t)n2 * q] = -weights * av[q];
    }
}

// mat_gen.f90:323-365 MAT_init_work (elastic + KV only)
inline void MAT_init_work(Problem& pb, bool force_general_nelast = false) {
  Grid& g = pb.grid;
  int n = g.ngll, n2 = n * n, ne = g.nelem;
  bool flat_grid = g.flat && !force_general_nelast;
  if (flat_grid)
    pb.nelast = (pb.ndof == 1) ? 2 : 6;
  else
    pb.nelast = (pb.ndof == 1) ? 3 : 10;
  // SE_firstElementTagged (spec_grid.f90:835-861): first element of each tag
  int maxtag = 0;
  for (int e = 0; e < ne; ++e) maxtag = std::max(maxtag, g.tag[e]);
  std::vector<int> first(maxtag + 1, 0);
  for (int e = ne; e >= 1; --e) first[g.tag[e - 1]] = e;
  pb.elem2set.assign(ne, 0);
  pb.elem2kv.assign(ne, 0);
  pb.a.clear();
  pb.beta25d.clear();
  const bool w25d = g.W < HUGE_D;
  std::vector<double> mu(n2), lambda(n2);
  pb.kv_elem.clear();
  pb.kv_eta.clear();
  pb.ncoefsets = 0;
  std::vector<double> abuf((size_t)n2 * pb.nelast), eta(n2);
  pb.elem2dm.assign(ne, 0);
  pb.dm_elem.clear();
  pb.dm_derint.clear();
  pb.dm_par.clear();
  pb.dm_state.clear();
  pb.elem2vs.assign(ne, 0);
  pb.vs_elem.clear();
  pb.vs_derint.clear();
  pb.vs_el.clear();
  pb.vs_off.clear();
  pb.vs_etot.clear();
  pb.elem2pl.assign(ne, 0);
  pb.pl_elem.clear();
  pb.pl_par.clear();
  pb.pl_ep.clear();
  pb.pl_derint.clear();
  pb.pl_beta.clear();
  for (int e = 1; e <= ne; ++e) {
    const MatInput& in = pb.mat.inputs[g.tag[e - 1] - 1];
    if (in.kv) {  // mat_kelvin_voigt.f90:117-135
      pb.mat.get(pb.mat.eta, e, eta.data());
      if (in.etaxdt)
        for (int k = 0; k < n2; ++k) eta[k] = pb.time.dt * eta[k];
      pb.kv_elem.push_back(e);
      pb.elem2kv[e - 1] = (int)pb.kv_elem.size();
      pb.kv_eta.insert(pb.kv_eta.end(), eta.begin(), eta.end());
    }
    if (in.damage) {  // mat_gen.f90:374-378: MAT_set_derint + MAT_DMG_init_elem_work (mat_damage.f90:209-279)
      if (pb.ndof != 2) IO_abort("oracle: damage rheology requires ndof=2 (P-SV)");
      pb.dm_elem.push_back(e);
      pb.elem2dm[e - 1] = (int)pb.dm_elem.size();
      const size_t o = pb.dm_derint.size();
      pb.dm_derint.resize(o + (size_t)5 * n2);
      for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
          const int k = (i - 1) + n * (j - 1);
          double jac[4], inv[4];
          SE_Jacobian(g, e, i, j, jac);
          invert2(jac, inv);
          pb.dm_derint[o + k] = inv[0];
          pb.dm_derint[o + n2 + k] = inv[2];
          pb.dm_derint[o + 2 * (size_t)n2 + k] = inv[1];
          pb.dm_derint[o + 3 * (size_t)n2 + k] = inv[3];
          pb.dm_derint[o + 4 * (size_t)n2 + k] = SE_VolumeWeight(g, e, i, j);
        }
      const double lam = pb.mat.lambda[e - 1].homo, mu1 = pb.mat.mu[e - 1].homo;
      double q = std::sin(in.phi * PI / 180.0);                                       // xi_zero_2d (:295-305)
      const double xi0 = -std::sqrt(2.0) / std::sqrt(q * q * ((lam / mu1 + 1.0) * (lam / mu1 + 1.0)) + 1.0);
      const double qq = 2.0 * (mu1 + lam) / (2.0 - xi0 * xi0);                        // gamma_r_2d (:308-317)
      const double pp = 0.5 * xi0 * (qq + lam);
      const double gr = pp + std::sqrt(pp * pp + 2.0 * mu1 * qq);
      double par[16] = {lam, mu1, xi0, gr, in.beta_dmg, in.Cd, in.Rdmg / mu1, in.e0[0], in.e0[1], in.e0[2], 0, 0, 0, 0, 0, 0};
      {  // initial stress (:262-265)
        const double mud = mu1 + xi0 * gr * in.alpha0;
        const double rg = gr * std::pow(in.alpha0, 1.0 + in.beta_dmg) / (1.0 + in.beta_dmg);
        const double ee[3] = {in.e0[0] - in.ep0[0], in.e0[1] - in.ep0[1], in.e0[2] - in.ep0[2]};
        double i1, i2, xi;
        DMG_compute_stress(&par[10], ee, lam, mud, rg, i1, i2, xi);
      }
      pb.dm_par.insert(pb.dm_par.end(), par, par + 16);
      const size_t so = pb.dm_state.size();
      pb.dm_state.resize(so + (size_t)4 * n2);
      for (int k = 0; k < n2; ++k) {
        pb.dm_state[so + k] = in.alpha0;
        for (int c = 0; c < 3; ++c) pb.dm_state[so + (size_t)(c + 1) * n2 + k] = in.ep0[c];
      }
      continue;
    }
    if (in.visco) {  // mat_gen.f90:380-385: MAT_set_derint + MAT_VISCO_init_elem_work (mat_visco.f90:164-200)
      if (pb.ndof != 2) IO_abort("MAT_init_work: visco-elasticity requires ndof=2 (P-SV) ");
      pb.vs_elem.push_back(e);
      pb.elem2vs[e - 1] = (int)pb.vs_elem.size();
      const size_t o = pb.vs_derint.size();
      pb.vs_derint.resize(o + (size_t)5 * n2);
      for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
          const int k = (i - 1) + n * (j - 1);
          double jac[4], inv[4];
          SE_Jacobian(g, e, i, j, jac);
          invert2(jac, inv);
          pb.vs_derint[o + k] = inv[0];
          pb.vs_derint[o + n2 + k] = inv[2];
          pb.vs_derint[o + 2 * (size_t)n2 + k] = inv[1];
          pb.vs_derint[o + 3 * (size_t)n2 + k] = inv[3];
          pb.vs_derint[o + 4 * (size_t)n2 + k] = SE_VolumeWeight(g, e, i, j);
        }
      pb.vs_off.push_back(pb.vs_el.size());
      pb.vs_el.resize(pb.vs_el.size() + (size_t)n2 * in.Nbody * 3, 0.0);
      pb.vs_etot.resize(pb.vs_etot.size() + (size_t)3 * n2, 0.0);
      continue;
    }
    if (in.plastic) {  // mat_gen.f90:367-372: MAT_set_derint (:645-681) + MAT_PLAST_init_elem_work (mat_plastic.f90:148-218)
      if (pb.ndof != 2) IO_abort("MAT_init_work: plasticity requires ndof=2 (P-SV) ");
      pb.pl_elem.push_back(e);
      pb.elem2pl[e - 1] = (int)pb.pl_elem.size();
      const size_t o = pb.pl_derint.size();
      pb.pl_derint.resize(o + (size_t)5 * n2);
      for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
          const int k = (i - 1) + n * (j - 1);
          double jac[4], inv[4];
          SE_Jacobian(g, e, i, j, jac);
          invert2(jac, inv);  // SE_InverseJacobian: xjaci(1,1)=dxi_dx, (1,2)=dxi_dy, (2,1)=deta_dx, (2,2)=deta_dy
          pb.pl_derint[o + k] = inv[0];
          pb.pl_derint[o + n2 + k] = inv[2];
          pb.pl_derint[o + 2 * (size_t)n2 + k] = inv[1];
          pb.pl_derint[o + 3 * (size_t)n2 + k] = inv[3];
          pb.pl_derint[o + 4 * (size_t)n2 + k] = SE_VolumeWeight(g, e, i, j);
        }
      const double lam = pb.mat.lambda[e - 1].homo, mu1 = pb.mat.mu[e - 1].homo;
      const double phi = PI / 180.0 * in.phi;
      double par[10] = {lam, mu1, in.coh * std::cos(phi), std::sin(phi), in.Tv > 0.0 ? 1.0 - std::exp(-pb.time.dt / in.Tv) : 1.0,
                        in.e0[0], in.e0[1], in.e0[2], 0.0, 0.0};
      pb.pl_par.insert(pb.pl_par.end(), par, par + 10);
      pb.pl_ep.resize(pb.pl_ep.size() + (size_t)3 * n2, 0.0);
      if (w25d) {  // MAT_PLAST_init_25D (mat_plastic.f90:221-241)
        for (int j = 1; j <= n; ++j)
          for (int i = 1; i <= n; ++i) {
            const double dvol = SE_VolumeWeight(g, e, i, j);
            const double nu = lam / (lam + mu1) / 2.0;
            const double t = 4.0 * std::atan(1.0) * (1 - nu) / g.W;
            pb.pl_beta.push_back(dvol * mu1 * (t * t));
          }
      }
      continue;
    }
    int e1 = first[g.tag[e - 1]];
    if (g.flat && !force_general_nelast && in.homogeneous && e > e1) {
      pb.elem2set[e - 1] = pb.elem2set[e1 - 1];
    } else {
      MAT_ELAST_init_a(pb, e, pb.nelast, abuf.data());
      pb.a.insert(pb.a.end(), abuf.begin(), abuf.end());
      if (w25d) {  // MAT_ELAST_init_25D (mat_elastic.f90:363-383)
        pb.mat.get(pb.mat.mu, e, mu.data());
        pb.mat.get(pb.mat.lambda, e, lambda.data());
        for (int j = 1; j <= n; ++j)
          for (int i = 1; i <= n; ++i) {
            const int k = (i - 1) + n * (j - 1);
            const double dvol = SE_VolumeWeight(g, e, i, j);
            const double nu = lambda[k] / (lambda[k] + mu[k]) / 2.0;
            const double t = (pb.ndof == 1) ? 4.0 * std::atan(1.0) / g.W : 4.0 * std::atan(1.0) * (1 - nu) / g.W;
            pb.beta25d.push_back(dvol * mu[k] * (t * t));
          }
      }
      pb.ncoefsets++;
      pb.elem2set[e - 1] = pb.ncoefsets;
    }
  }
}

// mat_mass.f90:29-61 MAT_MASS_init
inline void MAT_MASS_init(Problem& pb) {
  Grid& g = pb.grid;
  int n = g.ngll;
  pb.rmass.assign((size_t)g.npoin * pb.ndof, 0.0);
  std::vector<double> rho((size_t)n * n);
  for (int e = 1; e <= g.nelem; ++e) {
    pb.mat.get(pb.mat.rho, e, rho.data());
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i) {
        double ml = rho[(i - 1) + (size_t)n * (j - 1)] * SE_VolumeWeight(g, e, i, j);
        int k = g.ib(i, j, e);
        pb.rmass[k - 1] = pb.rmass[k - 1] + ml;
      }
  }
  if (pb.ndof == 2)
    for (int i = 0; i < g.npoin; ++i) pb.rmass[i + (size_t)g.npoin] = pb.rmass[i];
}

// ------------------------------------------------------------------------------------------
// bc_abso.f90:115-266 BC_ABSO_init
inline void BC_ABSO_init(BcAbso& bc, int tag, Problem& pb) {
  Grid& g = pb.grid;
  bc.topo = g.bc_inquire(tag);
  bc.periodic = BC_PERIO_intersects(*bc.topo, pb.perio);  // bc_abso.f90:139
  int ndof = pb.ndof, ngll = g.ngll;
  int bc_nelem = bc.topo->nelem, bc_npoin = bc.topo->npoin;
  std::vector<double> B;
  BC_get_normal_and_weights(*bc.topo, g, bc.n, B, bc.periodic);
  int GeoDimTan, GeoDimNor;
  bool all_x_small = true, all_z_small = true;
  for (int k = 0; k < bc_npoin; ++k) {
    if (!(std::fabs(bc.n[k]) < TINY_XABS)) all_x_small = false;
    if (!(std::fabs(bc.n[k + bc_npoin]) < TINY_XABS)) all_z_small = false;
  }
  if (all_x_small) {
    bc.is_flat = true;
    GeoDimTan = 1;
    GeoDimNor = 2;
  } else if (all_z_small) {
    bc.is_flat = true;
    GeoDimTan = 2;
    GeoDimNor = 1;
  } else {
    bc.is_flat = false;
    GeoDimTan = 2;
    GeoDimNor = 1;
  }
  bc.C.assign((size_t)bc_npoin * ndof, 0.0);
  bc.stacey = bc.stacey && (ndof == 2) && bc.is_flat;
  if (bc.stacey) bc.K.assign((size_t)ngll * ndof * bc_nelem, 0.0);
  std::vector<int> itab, jtab;
  for (int e = 1; e <= bc_nelem; ++e) {
    int ebulk = bc.topo->elem[e - 1];
    int edge = bc.topo->edge[e - 1];
    edge_tabs(ngll, edge, itab, jtab);
    int LocDimTan = (edge == edge_D || edge == edge_U) ? 1 : 2;
    for (int k = 1; k <= ngll; ++k) {
      int i = itab[k - 1], j = jtab[k - 1];
      double rho = pb.mat.get(pb.mat.rho, ebulk, i, j);
      double c[3];
      c[GeoDimNor] = pb.mat.get(pb.mat.cp, ebulk, i, j);
      c[GeoDimTan] = pb.mat.get(pb.mat.cs, ebulk, i, j);
      double xjac[4];
      SE_Jacobian(g, ebulk, i, j, xjac);
      double a0 = xjac[0 + 2 * (LocDimTan - 1)], a1 = xjac[1 + 2 * (LocDimTan - 1)];
      double CoefIntegr = g.wgll[k - 1] * std::sqrt(a0 * a0 + a1 * a1);
      int bck = bc.topo->ibool[(k - 1) + (size_t)ngll * (e - 1)];
      if (ndof == 1) {
        bc.C[bck - 1] = bc.C[bck - 1] + rho * c[GeoDimTan] * CoefIntegr;
      } else {
        bc.C[bck - 1] = bc.C[bck - 1] + rho * c[1] * CoefIntegr;
        bc.C[bck - 1 + bc_npoin] = bc.C[bck - 1 + bc_npoin] + rho * c[2] * CoefIntegr;
      }
      if (bc.stacey) {
        double xi[4];
        invert2(xjac, xi);  // DLocDGlob
        bc.K[(k - 1) + (size_t)ngll * (0 + 2 * (size_t)(e - 1))] =
            CoefIntegr * xi[(LocDimTan - 1) + 2 * (GeoDimTan - 1)] * rho * c[GeoDimTan] *
            (2.0 * c[GeoDimTan] - c[GeoDimNor]);
      }
    }
  }
  if (bc.stacey) {
    for (int e = 0; e < bc_nelem; ++e)
      for (int k = 0; k < ngll; ++k) bc.K[k + (size_t)ngll * (1 + 2 * (size_t)e)] = bc.K[k + (size_t)ngll * (0 + 2 * (size_t)e)];
    for (int e = 0; e < bc_nelem; ++e)
      for (int k = 0; k < ngll; ++k) {
        size_t q = k + (size_t)ngll * ((GeoDimTan - 1) + 2 * (size_t)e);
        bc.K[q] = -bc.K[q];
      }
  }
  if (bc.periodic && bc.is_flat) {  // bc_abso.f90:226-230
    if (bc.stacey) IO_abort("oracle: a Stacey boundary that meets a periodic one is not restated (bc_abso.f90:328-331)");
    for (int c = 0; c < ndof; ++c) {
      bc.C[0 + (size_t)bc_npoin * c] = bc.C[0 + (size_t)bc_npoin * c] + bc.C[(bc_npoin - 1) + (size_t)bc_npoin * c];
      bc.C[(bc_npoin - 1) + (size_t)bc_npoin * c] = bc.C[0 + (size_t)bc_npoin * c];
    }
  }
  if (bc.is_flat) {
    double coef = pb.time.CoefA2Vrhs();
    for (int c = 0; c < ndof; ++c)
      for (int k = 0; k < bc_npoin; ++k) {
        size_t q = pb.idx(bc.topo->node[k], c);
        pb.rmass[q] = pb.rmass[q] + coef * bc.C[k + (size_t)bc_npoin * c];
      }
  }
}

// bc_abso.f90:286-336 BC_ABSO_apply (no incident wave)
inline void BC_ABSO_apply(const BcAbso& bc, const Problem& pb, const std::vector<double>& D,
                          const std::vector<double>& V, std::vector<double>& MxA) {
  int ndof = pb.ndof, np = bc.topo->npoin, ngll = pb.grid.ngll;
  const std::vector<int>& nodes = bc.topo->node;
  if (bc.is_flat || ndof == 1) {
    for (int c = 0; c < ndof; ++c)
      for (int k = 0; k < np; ++k) {
        size_t q = pb.idx(nodes[k], c);
        MxA[q] = MxA[q] - bc.C[k + (size_t)np * c] * V[q];
      }
  } else {
    for (int k = 0; k < np; ++k) {
      size_t q1 = pb.idx(nodes[k], 0), q2 = pb.idx(nodes[k], 1);
      double vn = V[q1] * bc.n[k] + V[q2] * bc.n[k + np];
      double vn1 = vn * bc.n[k], vn2 = vn * bc.n[k + np];
      MxA[q1] = MxA[q1] - bc.C[k] * vn1 - bc.C[k + np] * (V[q1] - vn1);
      MxA[q2] = MxA[q2] - bc.C[k] * vn2 - bc.C[k + np] * (V[q2] - vn2);
    }
  }
  if (bc.stacey) {
    std::vector<double> KxD((size_t)np * 2, 0.0);
    const std::vector<double>& Ht = pb.grid.Ht;
    for (int e = 0; e < bc.topo->nelem; ++e) {
      for (int c = 0; c < 2; ++c)
        for (int i = 0; i < ngll; ++i) {
          double s = 0.0;  // matmul(Ht, D(nodes(k),:))
          for (int kk = 0; kk < ngll; ++kk) {
            int bn = bc.topo->ibool[kk + (size_t)ngll * e];
            s += Ht[i + (size_t)ngll * kk] * D[pb.idx(nodes[bn - 1], c)];
          }
          int bi = bc.topo->ibool[i + (size_t)ngll * e];
          KxD[(bi - 1) + (size_t)np * c] = KxD[(bi - 1) + (size_t)np * c] + bc.K[i + (size_t)ngll * (c + 2 * (size_t)e)] * s;
        }
    }
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < np; ++k) {
        size_t q = pb.idx(nodes[k], c);
        MxA[q] = MxA[q] - KxD[k + (size_t)np * c];
      }
  }
}

// bc_dirneu.f90:117-169
inline void BC_DIRNEU_init(BcDirneu& bc, int tag, Problem& pb) {
  bc.topo = pb.grid.bc_inquire(tag);
  std::vector<double> n, B;
  BC_get_normal_and_weights(*bc.topo, pb.grid, n, B, BC_PERIO_intersects(*bc.topo, pb.perio));
  int np = bc.topo->npoin;
  bool ax = true, az = true;
  for (int k = 0; k < np; ++k) {
    if (!(std::fabs(n[k]) < TINY_XABS)) ax = false;
    if (!(std::fabs(n[k + np]) < TINY_XABS)) az = false;
  }
  if (!(ax || az)) IO_abort("BC_DIRNEU_init: boundary is not vertical or horizontal");
}
inline void BC_DIRNEU_apply(const BcDirneu& bc, const Problem& pb, std::vector<double>& field) {
  if (bc.kind[0] == 2)
    for (int k : bc.topo->node) field[pb.idx(k, 0)] = 0.0;
  if (pb.ndof == 1) return;
  if (bc.kind[1] == 2)
    for (int k : bc.topo->node) field[pb.idx(k, 1)] = 0.0;
}

// ------------------------------------------------------------------------------------------
// friction laws
// bc_dynflt_swf.f90:142-160 swf_mu
inline double swf_mu_1(const Swf& f, int k) {
  double mu = 0;
  if (f.kind == 1)
    mu = f.mus[k] - (f.mus[k] - f.mud[k]) * std::min(f.theta[k] / f.dc[k], 1.0);
  else if (f.kind == 2)
    mu = f.mud[k] - (f.mud[k] - f.mus[k]) * std::exp(-f.theta[k] / f.dc[k]);
  else if (f.kind == 3)
    mu = f.mud[k] + (f.mus[k] - f.mud[k]) / std::pow(1.0 + f.theta[k] / f.dc[k], f.p[k]);
  mu = mu + f.alpha[k] * f.theta[k];
  return mu;
}
// bc_dynflt_twf.f90:117-184 twf_mu for one node
inline double twf_mu_1(const Twf& tw, double x, double z, double time, double d) {
  const double VERY_LARGE_VALUE = HUGE_D;
  double t, r = 0, mu = VERY_LARGE_VALUE;
  if (tw.kind == 1) {
    t = time + (tw.mus - tw.mu0) * tw.L / ((tw.mus - tw.mud) * tw.V);
    if (t > tw.T) t = 0.0;
    r = tw.V * t;
  } else if (tw.kind == 2) {
    t = time + 0.5 * tw.T * (1.0 - std::sqrt(1.0 - 4.0 * (tw.mus - tw.mu0) * tw.L / ((tw.mus - tw.mud) * tw.T * tw.V)));
    t = std::min(t, tw.T);
    r = tw.V * t * (1.0 - t / tw.T);
  }
  if (tw.kind == 1 || tw.kind == 2) {
    double rr = std::sqrt((x - tw.X) * (x - tw.X) + (z - tw.Z) * (z - tw.Z)) - r;
    if (rr < -tw.L)
      mu = tw.mud;
    else if (rr <= tw.L)
      mu = tw.mus + (tw.mus - tw.mud) / tw.L * rr;
    else
      mu = VERY_LARGE_VALUE;
  } else {
    t = time;
    double rr = std::sqrt((x - tw.X) * (x - tw.X) + (z - tw.Z) * (z - tw.Z));
    // NOTE: the reference leaves mu undefined when rr > V*T (bc_dynflt_twf.f90:168-181);
    // VERY_LARGE_VALUE (no weakening) is used for that case here.
    if (rr <= tw.V * tw.T && d <= tw.Dc) {
      if (rr < tw.V * t - tw.L)
        mu = tw.mud;
      else if (rr >= tw.V * t - tw.L && rr <= tw.V * t)
        mu = tw.mus + (tw.mus - tw.mud) / tw.L * (rr - tw.V * t);
      else
        mu = VERY_LARGE_VALUE;
    } else if (rr <= tw.V * tw.T && d > tw.Dc) {
      mu = VERY_LARGE_VALUE;
    }
  }
  return mu;
}
// bc_dynflt_rsf.f90:165-184 rsf_mu
inline double rsf_mu_1(const Rsf& f, int k, double v) {
  switch (f.kind) {
    case 1:
      return f.mus[k] + f.a[k] * std::fabs(v) / (std::fabs(v) + f.Vstar[k]) - f.b[k] * f.theta[k] / (f.theta[k] + f.dc[k]);
    case 2:
    case 3:
      return f.a[k] * std::asinh(std::fabs(v) / (2.0 * f.Vstar[k]) *
                                 std::exp((f.mus[k] + f.b[k] * std::log(f.Vstar[k] * f.theta[k] / f.dc[k])) / f.a[k]));
    case 4:
      return f.a[k] * std::asinh(std::fabs(v) / (2.0 * f.Vstar[k]) *
                                 std::exp((f.mus[k] + f.b[k] * std::log(f.Vc[k] * f.theta[k] / f.dc[k] + 1)) / f.a[k]));
  }
  return 0;
}
// bc_dynflt_rsf.f90:273-308 rsf_update_theta
inline double rsf_update_theta_1(const Rsf& f, int k, double theta, double v) {
  double theta_new = 0;
  switch (f.kind) {
    case 1:
      theta_new = theta * f.coeft[k] + f.Tc[k] * std::fabs(v) * (1.0 - f.coeft[k]);
      if (theta_new < 1.0e-12) theta_new = 0.0;
      break;
    case 2:
    case 4: {
      double x = std::fabs(v) / f.dc[k];
      double exp_x = std::exp(-f.dt * x);
      if (f.dt * x > 1e-8)
        theta_new = theta * exp_x + (1.0 - exp_x) / x;
      else
        theta_new = theta * exp_x + f.dt * (1.0 - 0.5 * f.dt * x);
      break;
    }
    case 3:
      theta_new = f.dc[k] / std::fabs(v);
      theta_new = theta_new * std::pow(theta / theta_new, std::exp(-f.dt / theta_new));
      break;
  }
  return theta_new;
}
// bc_dynflt_rsf.f90:533-568 nr_fric_func_tau
inline void nr_fric_func_tau(double tau, double& func_tau, double& dfunc_dtau, double& v, const Rsf& f,
                             double theta, int it, double tau_stick, double sigma, double Z) {
  double tmp;
  if (f.kind == 4)
    tmp = f.mus[it] + f.b[it] * std::log(f.Vc[it] * theta / f.dc[it] + 1.0);
  else
    tmp = f.mus[it] + f.b[it] * std::log(f.Vstar[it] * theta / f.dc[it]);
  tmp = 2.0 * f.Vstar[it] * std::exp(-tmp / f.a[it]);
  v = std::sinh(tau / (-sigma * f.a[it])) * tmp;
  func_tau = tau_stick - Z * v - tau;
  double dv_dtau = std::cosh(tau / (-sigma * f.a[it])) * tmp / (-sigma * f.a[it]);
  dfunc_dtau = -Z * dv_dtau - 1.0;
}
// bc_dynflt_rsf.f90:369-468 nr_solver (result is v, the velocity of the LAST function evaluation)
inline double nr_solver(double xL, double xR, double x_acc, const Rsf& f, int it, double theta,
                        double tau_stick, double sigma, double Z) {
  const int maxIteration = 200;
  double v = 0, dfunc_dx, f_low, f_high, func_x, x_est, dx, dx_old, x_high, x_low, temp;
  nr_fric_func_tau(xL, f_low, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
  nr_fric_func_tau(xR, f_high, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
  double xLeft = xL, xRight = xR;
  while (f_low * f_high > 0) {
    xLeft = xLeft / 2.0;
    xRight = xRight * 2.0;
    nr_fric_func_tau(xLeft, f_low, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
    nr_fric_func_tau(xRight, f_high, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
  }
  if (f_low == 0) {
    return v;
  } else if (f_high == 0) {
    return v;
  } else if (f_low < 0) {
    x_low = xLeft;
    x_high = xRight;
  } else {
    x_high = xLeft;
    x_low = xRight;
  }
  x_est = 0.5 * (xLeft + xRight);
  dx_old = std::fabs(xRight - xLeft);
  dx = dx_old;
  nr_fric_func_tau(x_est, func_x, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
  for (int is = 1; is <= maxIteration; ++is) {
    if (((x_est - x_high) * dfunc_dx - func_x) * ((x_est - x_low) * dfunc_dx - func_x) > 0 ||
        std::fabs(2 * func_x) > std::fabs(dx_old * dfunc_dx)) {
      dx_old = dx;
      dx = 0.5 * (x_high - x_low);
      x_est = x_low + dx;
      if (x_low == x_est) return v;
    } else {
      dx_old = dx;
      dx = func_x / dfunc_dx;
      temp = x_est;
      x_est = x_est - dx;
      if (temp == x_est) return v;
    }
    if (std::fabs(dx) < std::fabs(x_acc)) {
      nr_fric_func_tau(x_est, func_x, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
      return v;
    }
    nr_fric_func_tau(x_est, func_x, dfunc_dx, v, f, theta, it, tau_stick, sigma, Z);
    if (func_x < 0)
      x_low = x_est;
    else
      x_high = x_est;
  }
  IO_abort("NR_Solver has exceeded the maximum iterations (200)");
}
// bc_dynflt_rsf.f90:253-261 rsf_mu_no_direct
inline double rsf_mu_no_direct_1(const Rsf& f, int k) {
  if (f.kind == 1) return f.mus[k] - f.b[k] * f.theta[k] / (f.theta[k] + f.dc[k]);
  return f.mus[k] + f.b[k] * std::log(f.theta[k] * f.Vstar[k] / f.dc[k]);
}
// bc_dynflt_rsf.f90:325-358 rsf_update_V for one node
inline double rsf_update_V_1(const Rsf& f, int k, double tau_stick, double sigma, double theta, double Z) {
  if (f.kind == 1) {
    // NOTE: kind 1 uses f%theta (the stored state), not the theta argument (rsf.f90:338)
    double v = (tau_stick + sigma * rsf_mu_no_direct_1(f, k)) / Z;
    double tmp = v - f.Vstar[k] + sigma * f.a[k] / Z;
    v = 0.5 * (tmp + std::sqrt(tmp * tmp + 4.0 * v * f.Vstar[k]));
    v = std::max(0.0, v);
    if (v < 1.0e-12) v = 0.0;
    return v;
  }
  double tolerance = -0.001 * f.a[k] * sigma;
  double lo = std::min(0.0, tau_stick), hi = std::max(0.0, tau_stick);
  return nr_solver(lo, hi, tolerance, f, k, theta, tau_stick, sigma, Z);
}

// bc_dynflt_normal.f90:118-136 normal_update for one node
inline void normal_update_1(NormalLaw& n, int k, double Tn, double V) {
  switch (n.kind) {
    case 0: break;
    case 1: n.sigma[k] = Tn; break;
    case 2: n.sigma[k] = Tn + n.coef * (n.sigma[k] - Tn); break;
    case 3: n.sigma[k] = Tn + std::exp(-(std::fabs(V) + n.V) * n.coef) * (n.sigma[k] - Tn); break;
  }
}

// bc_dynflt.f90:832-855 BC_DYNFLT_potency
inline void BC_DYNFLT_potency(const BcDynflt& bc, const Problem& pb, const std::vector<double>& d, double* p) {
  int np = bc.npoin, ndof = pb.ndof;
  auto jump = [&](int k, int c) {
    if (bc.two_sides) return d[pb.idx(bc.node2[k], c)] - d[pb.idx(bc.node1[k], c)];
    return -2.0 * d[pb.idx(bc.node1[k], c)];
  };
  if (ndof == 2) {
    double s1 = 0, s2 = 0, s3 = 0;
    for (int k = 0; k < np; ++k) s1 += bc.n1[k] * jump(k, 0) * bc.B[k];
    for (int k = 0; k < np; ++k) s2 += bc.n1[k + np] * jump(k, 1) * bc.B[k];
    for (int k = 0; k < np; ++k) s3 += (bc.n1[k] * jump(k, 1) + bc.n1[k + np] * jump(k, 0)) * bc.B[k];
    p[0] = s1;
    p[1] = s2;
    p[2] = 0.5 * s3;
  } else {
    double s1 = 0, s2 = 0;
    for (int k = 0; k < np; ++k) s1 += bc.n1[k] * jump(k, 0) * bc.B[k];
    for (int k = 0; k < np; ++k) s2 += bc.n1[k + np] * jump(k, 0) * bc.B[k];
    p[0] = 0.5 * s1;
    p[1] = 0.5 * s2;
  }
}

// bc_dynflt.f90:751-778 BC_DYNFLT_write (records kept in memory)
inline void BC_DYNFLT_write(BcDynflt& bc, const Problem& pb, int itime) {
  int ndof = pb.ndof, np = bc.npoin;
  double p[6];
  BC_DYNFLT_potency(bc, pb, pb.d, p);
  BC_DYNFLT_potency(bc, pb, pb.v, p + (ndof + 1));
  bc.potency.insert(bc.potency.end(), p, p + 2 * (ndof + 1));
  if (itime < bc.oit) return;
  auto rec = [&](const double* arr) {
    for (int i = bc.oix1; i <= bc.oixn; i += bc.oixd) bc.out.push_back((float)arr[i - 1]);
  };
  rec(&bc.D[0]);
  rec(&bc.V[0]);
  rec(&bc.T[0]);
  rec(&bc.T[np]);
  rec(&bc.MU[0]);
  rec(&bc.Tstick[0]);
  bc.nout++;
  bc.oit = bc.oit + bc.oitd;
}

// bc_dynflt.f90:231-520 BC_DYNFLT_init
inline void BC_DYNFLT_init(BcDynflt& bc, const int tags[2], Problem& pb) {
  Grid& g = pb.grid;
  int ndof = pb.ndof;
  bc.tags[0] = tags[0];
  bc.tags[1] = tags[1];
  bc.two_sides = tags[1] > 0;
  bc.bc1 = g.bc_inquire(tags[0]);
  if (bc.two_sides) {
    bc.bc2 = g.bc_inquire(tags[1]);
    if (bc.bc1->nelem != bc.bc2->nelem) IO_abort("bc_dynflt_init: number of boundary elements do not match");
    if (bc.bc1->npoin != bc.bc2->npoin) IO_abort("bc_dynflt_init: number of nodes on boundaries do not match");
  }
  int np1 = bc.bc1->npoin;
  std::vector<char> keep(np1, 1);
  int npoin = np1;
  if (bc.two_sides)
    for (int k = 0; k < np1; ++k)
      if (bc.bc1->node[k] == bc.bc2->node[k]) {
        keep[k] = 0;
        npoin--;
      }
  bc.node1.clear();
  bc.node2.clear();
  for (int k = 0; k < np1; ++k)
    if (keep[k]) {
      bc.node1.push_back(bc.bc1->node[k]);
      if (bc.two_sides) bc.node2.push_back(bc.bc2->node[k]);
    }
  bc.npoin = npoin;
  bc.coord.resize((size_t)2 * npoin);
  for (int k = 0; k < npoin; ++k) {
    bc.coord[2 * k] = g.coord[2 * (size_t)(bc.node1[k] - 1)];
    bc.coord[2 * k + 1] = g.coord[2 * (size_t)(bc.node1[k] - 1) + 1];
  }
  if (bc.two_sides)
    for (int k = 0; k < npoin; ++k)
      if (std::fabs(bc.coord[2 * k] - g.coord[2 * (size_t)(bc.node2[k] - 1)]) > TINY_XABS ||
          std::fabs(bc.coord[2 * k + 1] - g.coord[2 * (size_t)(bc.node2[k] - 1) + 1]) > TINY_XABS)
        IO_abort("bc_dynflt_init: coordinates on boundaries do not match properly");
  {
    std::vector<double> tn, tB;
    BC_get_normal_and_weights(*bc.bc1, g, tn, tB, BC_PERIO_intersects(*bc.bc1, pb.perio));
    bc.n1.assign((size_t)2 * npoin, 0.0);
    bc.B.assign((size_t)npoin * ndof, 0.0);
    int j = 0;
    for (int k = 0; k < np1; ++k)
      if (keep[k]) {
        bc.B[j] = tB[k];
        bc.n1[j] = tn[k];
        bc.n1[j + npoin] = tn[k + np1];
        ++j;
      }
    if (ndof == 2)
      for (int k = 0; k < npoin; ++k) bc.B[k + npoin] = bc.B[k];
  }
  bc.CoefA2V = pb.time.CoefA2V();
  bc.CoefA2D = pb.time.CoefA2D();
  bc.invM1.assign((size_t)npoin * ndof, 0.0);
  for (int c = 0; c < ndof; ++c)
    for (int k = 0; k < npoin; ++k) bc.invM1[k + (size_t)npoin * c] = 1.0 / pb.rmass[pb.idx(bc.node1[k], c)];
  if (bc.two_sides) {
    bc.invM2.assign((size_t)npoin * ndof, 0.0);
    for (int c = 0; c < ndof; ++c)
      for (int k = 0; k < npoin; ++k) bc.invM2[k + (size_t)npoin * c] = 1.0 / pb.rmass[pb.idx(bc.node2[k], c)];
  }
  bc.Z.assign((size_t)npoin * ndof, 0.0);
  for (size_t q = 0; q < bc.Z.size(); ++q) {
    if (bc.two_sides)
      bc.Z[q] = 1.0 / (bc.CoefA2V * bc.B[q] * (bc.invM1[q] + bc.invM2[q]));
    else
      bc.Z[q] = 0.5 / (bc.CoefA2V * bc.B[q] * bc.invM1[q]);
  }
  double dt = pb.time.dt;
  auto gen = [&](const Dist& dd, std::vector<double>& out) {  // DIST_CD_Init_1
    out.resize(npoin);
    for (int k = 0; k < npoin; ++k) out[k] = dd.is_dist() ? dd.eval(bc.coord[2 * k], bc.coord[2 * k + 1]) : dd.c;
  };
  if (bc.swf) {  // swf_init (bc_dynflt_swf.f90:121-138)
    Swf& s = *bc.swf;
    gen(s.in_dc, s.dc);
    gen(s.in_mus, s.mus);
    gen(s.in_mud, s.mud);
    gen(s.in_p, s.p);
    gen(s.in_alpha, s.alpha);
    s.theta.assign(npoin, 0.0);
    s.dt = dt;
  } else if (bc.rsf) {  // rsf_init (bc_dynflt_rsf.f90:139-161)
    Rsf& r = *bc.rsf;
    gen(r.in_dc, r.dc);
    gen(r.in_mus, r.mus);
    gen(r.in_a, r.a);
    gen(r.in_b, r.b);
    gen(r.in_Vstar, r.Vstar);
    gen(r.in_theta, r.theta);
    gen(r.in_Vc, r.Vc);
    r.Tc.resize(npoin);
    r.coeft.resize(npoin);
    for (int k = 0; k < npoin; ++k) {
      r.Tc[k] = r.dc[k] / r.Vstar[k];
      r.coeft[k] = std::exp(-dt / r.Tc[k]);
    }
    r.dt = dt;
  }
  bc.T.assign((size_t)npoin * 2, 0.0);
  bc.Tstick.assign((size_t)npoin * 2, 0.0);
  bc.D.assign((size_t)npoin * ndof, 0.0);
  bc.V.assign((size_t)npoin * ndof, 0.0);
  if (bc.rsf) {
    std::vector<double> V;
    gen(bc.in_V, V);
    for (int k = 0; k < npoin; ++k) bc.V[k] = V[k];
  }
  {
    std::vector<double> Tt0, Tn0, Sxx, Sxy, Sxz, Syz, Szz;
    gen(bc.in_T, Tt0);
    gen(bc.in_N, Tn0);
    gen(bc.in_Sxx, Sxx);
    gen(bc.in_Sxy, Sxy);
    gen(bc.in_Sxz, Sxz);
    gen(bc.in_Syz, Syz);
    gen(bc.in_Szz, Szz);
    bc.T0.assign((size_t)npoin * 2, 0.0);
    for (int k = 0; k < npoin; ++k) {
      double nx = bc.n1[k], nz = bc.n1[k + npoin];
      double Tx = Sxx[k] * nx + Sxz[k] * nz;
      double Ty = Sxy[k] * nx + Syz[k] * nz;
      double Tz = Sxz[k] * nx + Szz[k] * nz;
      if (ndof == 1)
        bc.T0[k] = Tt0[k] + Ty;
      else
        bc.T0[k] = Tt0[k] + Tx * nz - Tz * nx;
      bc.T0[k + npoin] = Tn0[k] + Tx * nx + Tz * nz;
    }
  }
  bc.MU.assign(npoin, 0.0);
  for (int k = 0; k < npoin; ++k) {
    if (bc.swf) {
      bc.MU[k] = swf_mu_1(*bc.swf, k);
      if (bc.twf) bc.MU[k] = std::min(bc.MU[k], twf_mu_1(*bc.twf, bc.coord[2 * k], bc.coord[2 * k + 1], 0.0, bc.D[k]));
    } else if (bc.rsf) {
      bc.MU[k] = rsf_mu_1(*bc.rsf, k, bc.V[k]);
      if (bc.twf) bc.MU[k] = std::min(bc.MU[k], twf_mu_1(*bc.twf, bc.coord[2 * k], bc.coord[2 * k + 1], 0.0, bc.D[k]));
    } else if (bc.twf) {
      bc.MU[k] = twf_mu_1(*bc.twf, bc.coord[2 * k], bc.coord[2 * k + 1], 0.0, bc.D[k]);
    }
  }
  gen(bc.in_cohesion, bc.cohesion);
  for (int k = 0; k < npoin; ++k)
    if (bc.cohesion[k] < 0.0) IO_abort("bc_dynflt_init: cohesion must be positive");
  // normal_init (bc_dynflt_normal.f90:96-114)
  if (bc.normal.kind == 2) bc.normal.coef = std::exp(-dt / bc.normal.T);
  if (bc.normal.kind == 3) bc.normal.coef = dt / bc.normal.L;
  bc.normal.sigma.resize(npoin);
  for (int k = 0; k < npoin; ++k) bc.normal.sigma[k] = bc.T0[k + npoin];
  // outputs
  bc.oix1 = std::max(bc.oix1, 1);
  bc.oixn = std::min(bc.oixn, npoin);
  bc.oitd = std::max(1, (int)std::lround(bc.odt / dt));
  bc.odt = dt * bc.oitd;
  bc.oit = (int)std::lround(bc.ot1 / dt);
}

// bc_dynflt.f90:569-689 BC_DYNFLT_apply
inline void BC_DYNFLT_apply(BcDynflt& bc, const Problem& pb, std::vector<double>& MxA, const std::vector<double>& V,
                            const std::vector<double>& D, double time) {
  int ndof = pb.ndof, np = bc.npoin;
  std::vector<double> T((size_t)np * 2, 0.0), Tstick((size_t)np * 2, 0.0);
  std::vector<double> dD((size_t)np * ndof), dV((size_t)np * ndof), dA((size_t)np * ndof), strength(np);
  for (int c = 0; c < ndof; ++c)
    for (int k = 0; k < np; ++k) {
      size_t q = k + (size_t)np * c;
      size_t i1 = pb.idx(bc.node1[k], c);
      if (bc.two_sides) {
        size_t i2 = pb.idx(bc.node2[k], c);
        dD[q] = D[i2] - D[i1];
        dV[q] = V[i2] - V[i1];
        dA[q] = bc.invM2[q] * MxA[i2] - bc.invM1[q] * MxA[i1];
      } else {
        dD[q] = -2.0 * D[i1];
        dV[q] = -2.0 * V[i1];
        dA[q] = -2.0 * bc.invM1[q] * MxA[i1];
      }
    }
  for (int c = 0; c < ndof; ++c)
    for (int k = 0; k < np; ++k) {
      size_t q = k + (size_t)np * c;
      T[q] = bc.Z[q] * (dV[q] + bc.CoefA2V * dA[q]);
    }
  auto rotate = [&](std::vector<double>& v, int fb) {  // :722-739
    for (int k = 0; k < np; ++k) {
      double v1 = v[k], v2 = v[k + np], nx = bc.n1[k], nz = bc.n1[k + np];
      if (fb == 1) {
        v[k] = nz * v1 - nx * v2;
        v[k + np] = nx * v1 + nz * v2;
      } else {
        v[k] = nz * v1 + nx * v2;
        v[k + np] = -nx * v1 + nz * v2;
      }
    }
  };
  if (ndof == 2) {
    rotate(dD, 1);
    rotate(dV, 1);
    rotate(dA, 1);
    rotate(T, 1);
  }
  if (!bc.two_sides || ndof == 1)
    for (int k = 0; k < np; ++k) T[k + np] = 0.0;
  for (size_t q = 0; q < T.size(); ++q) T[q] = T[q] + bc.T0[q];
  if (bc.allow_opening)
    for (int k = 0; k < np; ++k) T[k + np] = std::min(T[k + np], 0.0);
  for (int k = 0; k < np; ++k) normal_update_1(bc.normal, k, T[k + np], dV[k]);
  if (bc.rsf) {
    Rsf& f = *bc.rsf;
    // rsf_solver (bc_dynflt_rsf.f90:229-249): two passes, array-wise
    std::vector<double> v_new(np), theta_new(np);
    for (int k = 0; k < np; ++k) theta_new[k] = rsf_update_theta_1(f, k, f.theta[k], bc.V[k]);
    for (int k = 0; k < np; ++k) v_new[k] = rsf_update_V_1(f, k, T[k], bc.normal.sigma[k], theta_new[k], bc.Z[k]);
    for (int k = 0; k < np; ++k) theta_new[k] = rsf_update_theta_1(f, k, f.theta[k], 0.5 * (bc.V[k] + v_new[k]));
    for (int k = 0; k < np; ++k) v_new[k] = rsf_update_V_1(f, k, T[k], bc.normal.sigma[k], theta_new[k], bc.Z[k]);
    for (int k = 0; k < np; ++k) {
      f.theta[k] = theta_new[k];
      bc.V[k] = v_new[k];
    }
    for (int k = 0; k < np; ++k) {
      bc.MU[k] = rsf_mu_1(f, k, bc.V[k]);
      if (bc.twf) bc.MU[k] = std::min(bc.MU[k], twf_mu_1(*bc.twf, bc.coord[2 * k], bc.coord[2 * k + 1], time, bc.D[k]));
      strength[k] = -bc.MU[k] * bc.normal.sigma[k];
      T[k] = std::copysign(std::fabs(strength[k]), T[k]);  // Fortran sign(strength,T)
    }
    // Tstick is left unassigned by the reference in this branch (garbage in output column 6)
  } else {
    for (int k = 0; k < np; ++k) {
      if (bc.swf) {
        Swf& s = *bc.swf;
        if (bc.CoefA2D == 0.0) {  // swf_update_state (:163-181)
          if (s.healing) {
            s.theta[k] = s.theta[k] + std::fabs(dV[k]) * s.dt;
            if (std::fabs(dV[k]) < 1e-14) s.theta[k] = 0.0;
          } else {
            s.theta[k] = std::fabs(dD[k]);
          }
        } else {
          s.theta[k] = std::fabs(bc.D[k]);  // swf_set_state
        }
        bc.MU[k] = swf_mu_1(s, k);
        if (bc.twf) bc.MU[k] = std::min(bc.MU[k], twf_mu_1(*bc.twf, bc.coord[2 * k], bc.coord[2 * k + 1], time, bc.D[k]));
      } else if (bc.twf) {
        bc.MU[k] = twf_mu_1(*bc.twf, bc.coord[2 * k], bc.coord[2 * k + 1], time, bc.D[k]);
      }
      strength[k] = bc.cohesion[k] - bc.MU[k] * bc.normal.sigma[k];
    }
    Tstick = T;
    for (int k = 0; k < np; ++k) {
      double m = std::min(std::fabs(T[k]), strength[k]);
      T[k] = std::copysign(std::fabs(m), T[k]);  // Fortran sign(a,b)
    }
  }
  for (size_t q = 0; q < T.size(); ++q) {
    T[q] = T[q] - bc.T0[q];
    Tstick[q] = Tstick[q] - bc.T0[q];
  }
  bc.T = T;
  bc.Tstick = Tstick;
  if (ndof == 2) rotate(T, -1);
  for (int c = 0; c < ndof; ++c)
    for (int k = 0; k < np; ++k) {
      size_t q = k + (size_t)np * c;
      size_t i1 = pb.idx(bc.node1[k], c);
      MxA[i1] = MxA[i1] + bc.B[q] * T[q];
    }
  if (bc.two_sides)
    for (int c = 0; c < ndof; ++c)
      for (int k = 0; k < np; ++k) {
        size_t q = k + (size_t)np * c;
        size_t i2 = pb.idx(bc.node2[k], c);
        MxA[i2] = MxA[i2] - bc.B[q] * T[q];
      }
  for (int c = 0; c < ndof; ++c)
    for (int k = 0; k < np; ++k) {
      size_t q = k + (size_t)np * c;
      dA[q] = dA[q] - bc.T[q] / (bc.Z[q] * bc.CoefA2V);
      bc.D[q] = dD[q] + bc.CoefA2D * dA[q];
      bc.V[q] = dV[q] + bc.CoefA2V * dA[q];
    }
}

// ------------------------------------------------------------------------------------------
// bc_periodic.f90:77-86 BC_PERIO_set_field: the vector-subscripted sum, then the copy back
inline void BC_PERIO_set_field(const BcPerio& bc, const Problem& pb, std::vector<double>& field) {
  int np = bc.master->npoin;
  for (int c = 0; c < pb.ndof; ++c) {
    for (int k = 0; k < np; ++k)
      field[pb.idx(bc.master->node[k], c)] = field[pb.idx(bc.master->node[k], c)] + field[pb.idx(bc.slave->node[k], c)];
    for (int k = 0; k < np; ++k) field[pb.idx(bc.slave->node[k], c)] = field[pb.idx(bc.master->node[k], c)];
  }
}
// bc_periodic.f90:44-74 BC_PERIO_init
inline void BC_PERIO_init(BcPerio& bc, const int tags[2], Problem& pb) {
  const Grid& g = pb.grid;
  bc.master = g.bc_inquire(tags[0]);
  bc.slave = g.bc_inquire(tags[1]);
  if (bc.master->nelem != bc.slave->nelem) IO_abort("bc_perio_init: number of boundary elements do not match");
  if (bc.master->npoin != bc.slave->npoin) IO_abort("bc_perio_init: number of nodes on boundaries do not match");
  const double TINY_XABS = 1e-3;  // constants.f90:36
  double s1 = g.coord[2 * (size_t)(bc.master->node[0] - 1)] - g.coord[2 * (size_t)(bc.slave->node[0] - 1)];
  double s2 = g.coord[2 * (size_t)(bc.master->node[0] - 1) + 1] - g.coord[2 * (size_t)(bc.slave->node[0] - 1) + 1];
  for (int k = 0; k < bc.master->npoin; ++k) {
    size_t m = (size_t)(bc.master->node[k] - 1), sl = (size_t)(bc.slave->node[k] - 1);
    if (std::fabs(g.coord[2 * m] - g.coord[2 * sl] - s1) > TINY_XABS || std::fabs(g.coord[2 * m + 1] - g.coord[2 * sl + 1] - s2) > TINY_XABS)
      IO_abort("bc_perio_init: coordinates on boundaries do not match properly");
  }
  BC_PERIO_set_field(bc, pb, pb.rmass);
}

// bc_gen.f90:190-250 bc_init and :256-308 bc_apply, :313-337 BC_write
inline void BC_write(Problem& pb, int itime) {
  for (auto& b : pb.bc)
    if (b.kind == IS_DYNFLT) BC_DYNFLT_write(*b.dynflt, pb, itime);
}
inline void BC_init(Problem& pb) {
  for (auto& b : pb.bc)
    for (int j = 0; j < 2; ++j)
      if (b.tag[j] != 0 && !pb.grid.bc_inquire(b.tag[j])) b.kind = IS_EMPTY;
  pb.perio = nullptr;  // first the periodic boundaries (bc_gen.f90:221-228)
  for (auto& b : pb.bc)
    if (b.kind == IS_PERIOD) {
      BC_PERIO_init(*b.perio, b.tag, pb);
      pb.perio = b.perio.get();
    }
  for (auto& b : pb.bc) {
    switch (b.kind) {
      case IS_DIRNEU: BC_DIRNEU_init(*b.dirneu, b.tag[0], pb); break;
      case IS_ABSORB: BC_ABSO_init(*b.abso, b.tag[0], pb); break;
      case IS_DYNFLT: BC_DYNFLT_init(*b.dynflt, b.tag, pb); break;
      default: break;
    }
  }
  BC_write(pb, 0);
}
inline void BC_apply(Problem& pb, std::vector<double>& field) {
  for (auto& b : pb.bc)  // first periodic, then absorbing, then the rest (bc_gen.f90:271-281)
    if (b.kind == IS_PERIOD) BC_PERIO_set_field(*b.perio, pb, field);
  for (auto& b : pb.bc)
    if (b.kind == IS_ABSORB) BC_ABSO_apply(*b.abso, pb, pb.d, pb.v, pb.a_);
  for (auto& b : pb.bc) {
    if (b.kind == IS_DIRNEU) BC_DIRNEU_apply(*b.dirneu, pb, field);
    if (b.kind == IS_DYNFLT) BC_DYNFLT_apply(*b.dynflt, pb, pb.a_, pb.v, pb.d, pb.time.time);
  }
}

// ------------------------------------------------------------------------------------------
// receivers.f90:231-303 REC_posit, :168-228 REC_init
inline void FE_find_point(const Grid& g, const double coord[2], int e, double& xi, double& eta, double newc[2]) {
  // fem_grid.f90:443-500.  NOTE: the reference tests the *initial* (xi,eta) for istatus
  // (the iterate lives in x(:) and is copied back after the test), so istatus is always 0.
  const int ntrial = 100;
  double cg[8];
  for (int k = 0; k < 4; ++k) {
    int kn = g.knods[4 * (size_t)(e - 1) + k];
    cg[2 * k] = g.coord_fem[2 * (size_t)(kn - 1)];
    cg[2 * k + 1] = g.coord_fem[2 * (size_t)(kn - 1) + 1];
  }
  double x1 = cg[0] - cg[4], x2 = cg[1] - cg[5], y1 = cg[2] - cg[6], y2 = cg[3] - cg[7];
  double area = 0.5 * std::fabs(x1 * y2 - x2 * y1);
  double esize = std::sqrt(area / PI);
  double x[2] = {xi, eta}, p[2] = {0, 0};
  double tolx = TINY_XABS, tolf = TINY_XABS * esize;
  int n;
  for (n = 1; n <= ntrial; ++n) {
    double sh[4], ds[8], fvec[2], fjac[4];
    Q4_getshape(x[0], x[1], sh);
    for (int r = 0; r < 2; ++r) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += cg[2 * k + r] * sh[k];
      fvec[r] = s - coord[r];
    }
    Q4_getdershape(x[0], x[1], ds);
    for (int c = 0; c < 2; ++c)
      for (int r = 0; r < 2; ++r) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += cg[2 * k + r] * ds[k + 4 * c];
        fjac[r + 2 * c] = s;
      }
    if (std::fabs(fvec[0]) + std::fabs(fvec[1]) <= tolf) break;
    p[0] = -fvec[0];
    p[1] = -fvec[1];
    double inv[4];
    invert2(fjac, inv);
    double q0 = inv[0] * p[0] + inv[2] * p[1], q1 = inv[1] * p[0] + inv[3] * p[1];
    p[0] = q0;
    p[1] = q1;
    x[0] += p[0];
    x[1] += p[1];
    if (std::fabs(p[0]) + std::fabs(p[1]) <= tolx) break;
  }
  if (n >= ntrial) IO_abort("FE_find_point: did not converge");
  xi = x[0];
  eta = x[1];
  double sh[4];
  Q4_getshape(xi, eta, sh);
  for (int r = 0; r < 2; ++r) {
    double s = 0;
    for (int k = 0; k < 4; ++k) s += cg[2 * k + r] * sh[k];
    newc[r] = s;
  }
}

inline void REC_init(Problem& pb) {
  Receivers& rec = pb.rec;
  if (!rec.present) return;
  Grid& g = pb.grid;
  int ngll = g.ngll;
  if (rec.AtNode) {
    std::vector<int> tmp;
    int irec = 0;
    for (int n = 0; n < rec.nx; ++n) {
      int ip = SE_find_nearest_node(g, rec.coord[2 * n], rec.coord[2 * n + 1]);
      bool dup = false;
      if (irec > 1)
        for (int q = 0; q < irec; ++q) dup = dup || (tmp[q] == ip);
      if (dup) continue;
      ++irec;
      tmp.push_back(ip);
    }
    rec.nx = irec;
    rec.iglob = tmp;
    rec.coord.resize((size_t)2 * rec.nx);
    for (int n = 0; n < rec.nx; ++n) {
      rec.coord[2 * n] = g.coord[2 * (size_t)(rec.iglob[n] - 1)];
      rec.coord[2 * n + 1] = g.coord[2 * (size_t)(rec.iglob[n] - 1) + 1];
    }
  } else {
    rec.interp.assign((size_t)ngll * ngll * rec.nx, 0.0);
    rec.einterp.assign(rec.nx, 0);
    for (int n = 0; n < rec.nx; ++n) {
      double c[2] = {rec.coord[2 * n], rec.coord[2 * n + 1]};
      int iglob = SE_find_nearest_node(g, c[0], c[1]);
      // SE_node_belongs_to_2 + SE_find_point (spec_grid.f90:457-524): first element wins (see note)
      int ef = 0, fi = 0, fj = 0;
      for (int e = 1; e <= g.nelem && !ef; ++e)
        for (int j = 1; j <= ngll && !ef; ++j)
          for (int i = 1; i <= ngll && !ef; ++i)
            if (g.ib(i, j, e) == iglob) {
              ef = e;
              fi = i;
              fj = j;
            }
      double xi = g.xgll[fi - 1], eta = g.xgll[fj - 1], newc[2];
      FE_find_point(g, c, ef, xi, eta, newc);
      rec.einterp[n] = ef;
      int k = 0;
      for (int j = 1; j <= ngll; ++j) {
        double fjv = gll::hgll(j - 1, eta, g.xgll.data(), ngll);
        for (int i = 1; i <= ngll; ++i) {
          double fiv = gll::hgll(i - 1, xi, g.xgll.data(), ngll);
          rec.interp[k + (size_t)ngll * ngll * n] = fiv * fjv;
          ++k;
        }
      }
      rec.coord[2 * n] = newc[0];
      rec.coord[2 * n + 1] = newc[1];
    }
  }
  rec.tsamp = pb.time.dt * rec.isamp;
  rec.nt = pb.time.nt / rec.isamp + 1;
  rec.sis.assign((size_t)rec.nt * rec.nx * pb.ndof, 0.0f);
}

// receivers.f90:309-344 REC_store
inline void REC_store(Problem& pb, int it) {
  Receivers& rec = pb.rec;
  if (!rec.present) return;
  if (it % rec.isamp != 0) return;
  int itsis = it / rec.isamp + 1;
  if (itsis > rec.nt) IO_abort("receivers.REC_store: storage is full");
  const std::vector<double>& field = (rec.field == 'D') ? pb.d : (rec.field == 'V') ? pb.v : pb.a_;
  int ngll = pb.grid.ngll;
  for (int c = 0; c < pb.ndof; ++c)
    for (int n = 0; n < rec.nx; ++n) {
      double val;
      if (rec.AtNode) {
        val = field[pb.idx(rec.iglob[n], c)];
      } else {
        int e = rec.einterp[n];
        double s = 0.0;
        int k = 0;
        for (int j = 1; j <= ngll; ++j)
          for (int i = 1; i <= ngll; ++i) {
            s += rec.interp[k + (size_t)ngll * ngll * n] * field[pb.idx(pb.grid.ib(i, j, e), c)];
            ++k;
          }
        val = s;
      }
      rec.sis[(size_t)(itsis - 1) + (size_t)rec.nt * (n + (size_t)rec.nx * c)] = (float)val;
    }
}

// ------------------------------------------------------------------------------------------
// element kernels (mat_elastic.f90:464-799) -- f,d local (ngll,ngll,ndof) col-major
// mxm / My_MATMUL: C(i,j) = sum_k A(i,k) B(k,j), k ascending (mxmlib.f90, mat_elastic.f90:779-799)
inline void mxm(const double* A, const double* B, double* C, int n) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = A[i] * B[(size_t)n * j];
      for (int k = 1; k < n; ++k) s = s + A[i + (size_t)n * k] * B[k + (size_t)n * j];
      C[i + (size_t)n * j] = s;
    }
}
inline void My_MATMUL(const double* A, const double* B, double* C, int n) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double Cij = 0.0;
      for (int k = 0; k < n; ++k) Cij = Cij + A[i + (size_t)n * k] * B[k + (size_t)n * j];
      C[i + (size_t)n * j] = Cij;
    }
}

struct ElemScratch {
  std::vector<double> t, g1, g2, g3, g4, r1, r2;
  explicit ElemScratch(int n) : t(n * n), g1(n * n), g2(n * n), g3(n * n), g4(n * n), r1(n * n), r2(n * n) {}
};

// MAT_ELAST_f (mat_elastic.f90:396-426): dispatch KD2 (ngll==OPT_NGLL) / KD1
inline void MAT_ELAST_f(double* f, const double* d, const double* a, int nelast, const double* H, const double* Ht,
                        int n, int ndof, ElemScratch& s, bool force_kd1) {
  int n2 = n * n;
  bool kd2 = (n == OPT_NGLL) && !force_kd1;
  auto MM = [&](const double* A, const double* B, double* C) {
    if (kd2)
      My_MATMUL(A, B, C, n);
    else
      mxm(A, B, C, n);
  };
  if (ndof == 1) {
    double *dU_dxi = s.g1.data(), *dU_deta = s.g2.data(), *tmp = s.t.data(), *r = s.r1.data();
    MM(Ht, d, dU_dxi);
    MM(d, H, dU_deta);
    if (nelast == 2) {
      for (int k = 0; k < n2; ++k) tmp[k] = a[k] * dU_dxi[k];
      MM(H, tmp, f);
      for (int k = 0; k < n2; ++k) tmp[k] = a[k + n2] * dU_deta[k];
      MM(tmp, Ht, r);
      for (int k = 0; k < n2; ++k) f[k] = f[k] + r[k];
    } else {
      for (int k = 0; k < n2; ++k) tmp[k] = a[k] * dU_dxi[k] + a[k + 2 * n2] * dU_deta[k];
      MM(H, tmp, f);
      for (int k = 0; k < n2; ++k) tmp[k] = a[k + 2 * n2] * dU_dxi[k] + a[k + n2] * dU_deta[k];
      MM(tmp, Ht, r);
      for (int k = 0; k < n2; ++k) f[k] = f[k] + r[k];
    }
    return;
  }
  double *dUx_dxi = s.g1.data(), *dUz_dxi = s.g2.data(), *dUx_deta = s.g3.data(), *dUz_deta = s.g4.data();
  double *tmp = s.t.data(), *r = s.r1.data();
  const double *dx = d, *dz = d + n2;
  double *fx = f, *fz = f + n2;
  MM(Ht, dx, dUx_dxi);
  MM(Ht, dz, dUz_dxi);
  MM(dx, H, dUx_deta);
  MM(dz, H, dUz_deta);
  auto A = [&](int q) { return a + (size_t)n2 * (q - 1); };
  if (nelast == 6) {
    for (int k = 0; k < n2; ++k) tmp[k] = A(1)[k] * dUx_dxi[k] + A(2)[k] * dUz_deta[k];
    MM(H, tmp, fx);
    if (kd2)
      for (int k = 0; k < n2; ++k) tmp[k] = A(4)[k] * (dUx_deta[k] + dUz_dxi[k]);  // :612
    else
      for (int k = 0; k < n2; ++k) tmp[k] = A(4)[k] * dUx_deta[k] + A(5)[k] * dUz_dxi[k];  // :489
    MM(tmp, Ht, r);
    for (int k = 0; k < n2; ++k) fx[k] = fx[k] + r[k];
    for (int k = 0; k < n2; ++k) tmp[k] = A(5)[k] * dUx_deta[k] + A(6)[k] * dUz_dxi[k];
    MM(H, tmp, fz);
    for (int k = 0; k < n2; ++k) tmp[k] = A(2)[k] * dUx_dxi[k] + A(3)[k] * dUz_deta[k];
    MM(tmp, Ht, r);
    for (int k = 0; k < n2; ++k) fz[k] = fz[k] + r[k];
  } else {
    for (int k = 0; k < n2; ++k)
      tmp[k] = A(1)[k] * dUx_dxi[k] + A(7)[k] * dUx_deta[k] + A(8)[k] * dUz_dxi[k] + A(2)[k] * dUz_deta[k];
    MM(H, tmp, fx);
    for (int k = 0; k < n2; ++k)
      tmp[k] = A(7)[k] * dUx_dxi[k] + A(4)[k] * dUx_deta[k] + A(5)[k] * dUz_dxi[k] + A(9)[k] * dUz_deta[k];
    MM(tmp, Ht, r);
    for (int k = 0; k < n2; ++k) fx[k] = fx[k] + r[k];
    for (int k = 0; k < n2; ++k)
      tmp[k] = A(8)[k] * dUx_dxi[k] + A(5)[k] * dUx_deta[k] + A(6)[k] * dUz_dxi[k] + A(10)[k] * dUz_deta[k];
    MM(H, tmp, fz);
    for (int k = 0; k < n2; ++k)
      tmp[k] = A(2)[k] * dUx_dxi[k] + A(9)[k] * dUx_deta[k] + A(10)[k] * dUz_dxi[k] + A(3)[k] * dUz_deta[k];
    MM(tmp, Ht, r);
    for (int k = 0; k < n2; ++k) fz[k] = fz[k] + r[k];
  }
}

// solver.f90:273-320 compute_Fint (elastic / Kelvin-Voigt branch of MAT_Fint, mat_gen.f90:418-445)
inline void compute_Fint(Problem& pb, std::vector<double>& f, const std::vector<double>& d, const std::vector<double>& v) {
  const Grid& g = pb.grid;
  int n = g.ngll, n2 = n * n, ndof = pb.ndof;
  std::fill(f.begin(), f.end(), 0.0);
  std::vector<double> dloc((size_t)n2 * ndof), vloc((size_t)n2 * ndof), floc((size_t)n2 * ndof);
  ElemScratch s(n);
  size_t np = g.npoin;
  for (int e = 1; e <= g.nelem; ++e) {
    const int* ib = &g.ibool[(size_t)n2 * (e - 1)];
    for (int c = 0; c < ndof; ++c)
      for (int k = 0; k < n2; ++k) {
        dloc[k + (size_t)n2 * c] = d[(size_t)(ib[k] - 1) + np * c];
        vloc[k + (size_t)n2 * c] = v[(size_t)(ib[k] - 1) + np * c];
      }
    // Kelvin-Voigt is the one non-exclusive material: d + eta*v before ANY constitutive law (mat_gen.f90:435)
    int ikv = pb.elem2kv[e - 1];
    if (ikv > 0) {  // MAT_KV_add_etav (mat_kelvin_voigt.f90:137-150)
      const double* eta = &pb.kv_eta[(size_t)n2 * (ikv - 1)];
      for (int c = 0; c < ndof; ++c)
        for (int k = 0; k < n2; ++k) dloc[k + (size_t)n2 * c] = dloc[k + (size_t)n2 * c] + eta[k] * vloc[k + (size_t)n2 * c];
    }
    if (!pb.elem2dm.empty() && pb.elem2dm[e - 1] > 0) {
      // mat_gen.f90:451-457: e = MAT_strain(d), MAT_DMG_stress(update = true, dt) (mat_damage.f90:337-445), f = MAT_forces(s)
      const int id = pb.elem2dm[e - 1] - 1;
      const double dt = pb.time.dt;
      const double* D = &pb.dm_derint[(size_t)5 * n2 * id];
      const double *dxi_dx = D, *dxi_dy = D + n2, *deta_dx = D + 2 * n2, *deta_dy = D + 3 * n2, *wts = D + 4 * n2;
      const double* par = &pb.dm_par[(size_t)16 * id];
      double* al = &pb.dm_state[(size_t)4 * n2 * id];
      double* ep = al + n2;
      std::vector<double> gx1(n2), gx2(n2), ge1(n2), ge2(n2), st((size_t)3 * n2), t1(n2), t2(n2), m1(n2), m2(n2);
      mxm(g.Ht.data(), dloc.data(), gx1.data(), n);
      mxm(g.Ht.data(), dloc.data() + n2, gx2.data(), n);
      mxm(dloc.data(), g.H.data(), ge1.data(), n);
      mxm(dloc.data() + n2, g.H.data(), ge2.data(), n);
      const double lam = par[0], mu0 = par[1], xi0 = par[2], gr = par[3], beta = par[4], Cd = par[5], Cv = par[6];
      for (int k = 0; k < n2; ++k) {
        const double et1 = gx1[k] * dxi_dx[k] + ge1[k] * deta_dx[k];
        const double et2 = gx2[k] * dxi_dy[k] + ge2[k] * deta_dy[k];
        const double et3 = 0.5 * (gx1[k] * dxi_dy[k] + ge1[k] * deta_dy[k] + gx2[k] * dxi_dx[k] + ge2[k] * deta_dx[k]);
        double ee[3] = {et1 + par[7], et2 + par[8], et3 + par[9]};
        for (int c = 0; c < 3; ++c) ee[c] = ee[c] - ep[(size_t)c * n2 + k];
        const double rm = mu0 + xi0 * gr * al[k];
        const double rg = gr * std::pow(al[k], 1.0 + beta) / (1.0 + beta);
        double sij[3], i1, i2, xi;
        DMG_compute_stress(sij, ee, lam, rm, rg, i1, i2, xi);
        double dalpha;
        if (beta == 0.0) dalpha = dt * Cd * i2 * std::max(xi - xi0, 0.0);
        else dalpha = dt * Cd * i2 * std::max(xi * std::pow(al[k], beta) - xi0, 0.0);
        al[k] = al[k] + dalpha;
        const double sm = 0.5 * (sij[0] + sij[1]);
        dalpha = Cv * std::max(dalpha, 0.0);
        ep[k] = ep[k] + (sij[0] - sm) * dalpha;
        ep[n2 + k] = ep[n2 + k] + (sij[1] - sm) * dalpha;
        ep[2 * n2 + k] = ep[2 * n2 + k] + sij[2] * dalpha;
        st[k] = sij[0] - par[10];
        st[n2 + k] = sij[1] - par[11];
        st[2 * n2 + k] = sij[2] - par[12];
      }
      for (int c = 0; c < 2; ++c) {  // MAT_forces (mat_gen.f90:834-866)
        const double* sa = c == 0 ? &st[0] : &st[2 * n2];
        const double* sb = c == 0 ? &st[2 * n2] : &st[n2];
        for (int k = 0; k < n2; ++k) {
          t1[k] = -wts[k] * (dxi_dx[k] * sa[k] + dxi_dy[k] * sb[k]);
          t2[k] = -wts[k] * (deta_dx[k] * sa[k] + deta_dy[k] * sb[k]);
        }
        mxm(g.H.data(), t1.data(), m1.data(), n);
        mxm(t2.data(), g.Ht.data(), m2.data(), n);
        for (int k = 0; k < n2; ++k) floc[k + (size_t)n2 * c] = m1[k] + m2[k];
      }
      for (int c = 0; c < ndof; ++c)
        for (int k = 0; k < n2; ++k) f[(size_t)(ib[k] - 1) + np * c] = f[(size_t)(ib[k] - 1) + np * c] + floc[k + (size_t)n2 * c];
      continue;
    }
    if (!pb.elem2vs.empty() && pb.elem2vs[e - 1] > 0) {
      // mat_gen.f90:451-457: e = MAT_strain(d), MAT_VISCO_stress (mat_visco.f90:206-248), f = MAT_forces(s); no 2.5D term
      const int iv = pb.elem2vs[e - 1] - 1;
      const MatInput& in = pb.mat.inputs[g.tag[e - 1] - 1];
      const int NB = in.Nbody;
      const double dt = pb.time.dt;
      const double* D = &pb.vs_derint[(size_t)5 * n2 * iv];
      const double *dxi_dx = D, *dxi_dy = D + n2, *deta_dx = D + 2 * n2, *deta_dy = D + 3 * n2, *wts = D + 4 * n2;
      double* el = &pb.vs_el[pb.vs_off[iv]];          // el(ngll,ngll,Nbody,3)
      double* eo = &pb.vs_etot[(size_t)3 * n2 * iv];  // etot_old(ngll,ngll,3)
      std::vector<double> gx1(n2), gx2(n2), ge1(n2), ge2(n2), st((size_t)3 * n2), t1(n2), t2(n2), m1(n2), m2(n2);
      mxm(g.Ht.data(), dloc.data(), gx1.data(), n);
      mxm(g.Ht.data(), dloc.data() + n2, gx2.data(), n);
      mxm(dloc.data(), g.H.data(), ge1.data(), n);
      mxm(dloc.data() + n2, g.H.data(), ge2.data(), n);
      const double lambda = in.lambda, two_mu = 2.0 * in.mu;
      for (int b = 0; b < NB; ++b) {  // memory variables advanced with the strain of the previous evaluation
        const double x = in.wbody[b] * dt;
        const double RK = x - (x * x) / 2.0 + (x * x * x) / 6.0 - (x * x * x * x) / 24.0;
        for (int c = 0; c < 3; ++c)
          for (int k = 0; k < n2; ++k) {
            double& q = el[k + (size_t)n2 * (b + (size_t)NB * c)];
            q = q + RK * (eo[k + (size_t)n2 * c] - q);
          }
      }
      for (int k = 0; k < n2; ++k) {
        const double et1 = gx1[k] * dxi_dx[k] + ge1[k] * deta_dx[k];
        const double et2 = gx2[k] * dxi_dy[k] + ge2[k] * deta_dy[k];
        const double et3 = 0.5 * (gx1[k] * dxi_dy[k] + ge1[k] * deta_dy[k] + gx2[k] * dxi_dx[k] + ge2[k] * deta_dx[k]);
        eo[k] = et1;
        eo[n2 + k] = et2;
        eo[2 * n2 + k] = et3;
        double sa1 = 0, sa2 = 0, sa3 = 0;
        for (int b = 0; b < NB; ++b) {
          const double e1 = el[k + (size_t)n2 * (b + (size_t)NB * 0)], e2 = el[k + (size_t)n2 * (b + (size_t)NB * 1)],
                       e3 = el[k + (size_t)n2 * (b + (size_t)NB * 2)];
          sa1 = sa1 + in.theta[b] * e1 + in.theta[b + NB] * e2;
          sa2 = sa2 + in.theta[b + NB] * e1 + in.theta[b] * e2;
          sa3 = sa3 + in.theta[b + 2 * NB] * e3;
        }
        st[k] = (lambda + two_mu) * et1 + lambda * et2 - sa1;
        st[n2 + k] = lambda * et1 + (lambda + two_mu) * et2 - sa2;
        st[2 * n2 + k] = two_mu * et3 - sa3;
      }
      for (int c = 0; c < 2; ++c) {  // MAT_forces (mat_gen.f90:834-866)
        const double* sa = c == 0 ? &st[0] : &st[2 * n2];
        const double* sb = c == 0 ? &st[2 * n2] : &st[n2];
        for (int k = 0; k < n2; ++k) {
          t1[k] = -wts[k] * (dxi_dx[k] * sa[k] + dxi_dy[k] * sb[k]);
          t2[k] = -wts[k] * (deta_dx[k] * sa[k] + deta_dy[k] * sb[k]);
        }
        mxm(g.H.data(), t1.data(), m1.data(), n);
        mxm(t2.data(), g.Ht.data(), m2.data(), n);
        for (int k = 0; k < n2; ++k) floc[k + (size_t)n2 * c] = m1[k] + m2[k];
      }
      for (int c = 0; c < ndof; ++c)
        for (int k = 0; k < n2; ++k) f[(size_t)(ib[k] - 1) + np * c] = f[(size_t)(ib[k] - 1) + np * c] + floc[k + (size_t)n2 * c];
      continue;
    }
    if (!pb.elem2pl.empty() && pb.elem2pl[e - 1] > 0) {
      // mat_gen.f90:445-449: e = MAT_strain(d), MAT_PLAST_stress(update = true), f = MAT_forces(s) (, 2.5D term)
      const int ip = pb.elem2pl[e - 1] - 1;
      const double* D = &pb.pl_derint[(size_t)5 * n2 * ip];
      const double *dxi_dx = D, *dxi_dy = D + n2, *deta_dx = D + 2 * n2, *deta_dy = D + 3 * n2, *wts = D + 4 * n2;
      const double* par = &pb.pl_par[(size_t)10 * ip];
      double* ep = &pb.pl_ep[(size_t)3 * n2 * ip];
      std::vector<double> gx1(n2), gx2(n2), ge1(n2), ge2(n2), st((size_t)3 * n2), t1(n2), t2(n2), m1(n2), m2(n2);
      // MAT_strain_PSV (mat_gen.f90:752-775)
      mxm(g.Ht.data(), dloc.data(), gx1.data(), n);
      mxm(g.Ht.data(), dloc.data() + n2, gx2.data(), n);
      mxm(dloc.data(), g.H.data(), ge1.data(), n);
      mxm(dloc.data() + n2, g.H.data(), ge2.data(), n);
      const double lambda = par[0], two_mu = 2.0 * par[1];
      const double s0[3] = {(lambda + two_mu) * par[5] + lambda * par[6], lambda * par[5] + (lambda + two_mu) * par[6],
                            two_mu * par[7]};  // mat_plastic.f90:199-204
      for (int k = 0; k < n2; ++k) {
        const double et1 = gx1[k] * dxi_dx[k] + ge1[k] * deta_dx[k];
        const double et2 = gx2[k] * dxi_dy[k] + ge2[k] * deta_dy[k];
        const double et3 = 0.5 * (gx1[k] * dxi_dy[k] + ge1[k] * deta_dy[k] + gx2[k] * dxi_dx[k] + ge2[k] * deta_dx[k]);
        // MAT_PLAST_stress (mat_plastic.f90:281-387), update = .true.
        double e1 = et1 - ep[k], e2 = et2 - ep[n2 + k], e3 = et3 - ep[2 * n2 + k];
        e1 = e1 + par[5];
        e2 = e2 + par[6];
        e3 = e3 + par[7];
        double s1 = (lambda + two_mu) * e1 + lambda * e2;
        double s2 = lambda * e1 + (lambda + two_mu) * e2;
        double s3 = two_mu * e3;
        const double tau = std::sqrt(0.25 * ((s1 - s2) * (s1 - s2)) + s3 * s3);
        const double sm = 0.5 * (s1 + s2);
        const double Y = par[2] - par[3] * sm;
        const double sdt1 = s1 - sm, sdt2 = s2 - sm, sdt3 = s3;
        const double factor = 1.0 - std::max(1.0 - Y / tau, 0.0) * par[4];
        const double sd1 = factor * sdt1, sd2 = factor * sdt2, sd3 = factor * sdt3;
        ep[k] = ep[k] + (sdt1 - sd1) / two_mu;
        ep[n2 + k] = ep[n2 + k] + (sdt2 - sd2) / two_mu;
        ep[2 * n2 + k] = ep[2 * n2 + k] + (sdt3 - sd3) / two_mu;
        s1 = sd1 + sm;
        s2 = sd2 + sm;
        s3 = sd3;
        st[k] = s1 - s0[0];
        st[n2 + k] = s2 - s0[1];
        st[2 * n2 + k] = s3 - s0[2];
      }
      // MAT_forces (mat_gen.f90:834-866)
      for (int c = 0; c < 2; ++c) {
        const double* sa = c == 0 ? &st[0] : &st[2 * n2];
        const double* sb = c == 0 ? &st[2 * n2] : &st[n2];
        for (int k = 0; k < n2; ++k) {
          t1[k] = -wts[k] * (dxi_dx[k] * sa[k] + dxi_dy[k] * sb[k]);
          t2[k] = -wts[k] * (deta_dx[k] * sa[k] + deta_dy[k] * sb[k]);
        }
        mxm(g.H.data(), t1.data(), m1.data(), n);
        mxm(t2.data(), g.Ht.data(), m2.data(), n);
        for (int k = 0; k < n2; ++k) floc[k + (size_t)n2 * c] = m1[k] + m2[k];
      }
      if (!pb.pl_beta.empty()) {  // MAT_PLAST_add_25D_f (mat_plastic.f90:262-274)
        const double* beta = &pb.pl_beta[(size_t)n2 * ip];
        for (int c = 0; c < ndof; ++c)
          for (int k = 0; k < n2; ++k) floc[k + (size_t)n2 * c] = floc[k + (size_t)n2 * c] - beta[k] * dloc[k + (size_t)n2 * c];
      }
      for (int c = 0; c < ndof; ++c)
        for (int k = 0; k < n2; ++k) f[(size_t)(ib[k] - 1) + np * c] = f[(size_t)(ib[k] - 1) + np * c] + floc[k + (size_t)n2 * c];
      continue;
    }
    const double* a = &pb.a[(size_t)n2 * pb.nelast * (pb.elem2set[e - 1] - 1)];
    MAT_ELAST_f(floc.data(), dloc.data(), a, pb.nelast, g.H.data(), g.Ht.data(), n, ndof, s, pb.kd_force_kd1);
    if (!pb.beta25d.empty()) {  // MAT_ELAST_add_25D_f (mat_gen.f90:440, mat_elastic.f90:447-459): the KV-modified d
      const double* beta = &pb.beta25d[(size_t)n2 * (pb.elem2set[e - 1] - 1)];
      for (int c = 0; c < ndof; ++c)
        for (int k = 0; k < n2; ++k) floc[k + (size_t)n2 * c] = floc[k + (size_t)n2 * c] - beta[k] * dloc[k + (size_t)n2 * c];
    }
    for (int c = 0; c < ndof; ++c)  // FIELD_add_elem (fields.f90:113-129)
      for (int k = 0; k < n2; ++k) f[(size_t)(ib[k] - 1) + np * c] = f[(size_t)(ib[k] - 1) + np * c] + floc[k + (size_t)n2 * c];
  }
}

// src_moment.f90:129-180 SRC_MOMENT_init, with SE_node_belongs_to (spec_grid.f90:385-409: elements in
// ascending order, each holding the node once)
inline void SRC_MOMENT_init(Source& so, Problem& pb) {
  const Grid& g = pb.grid;
  int n = g.ngll, ndof = pb.ndof;
  std::vector<int> etab, itab, jtab;
  for (int e = 1; e <= g.nelem; ++e)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i)
        if (g.ib(i, j, e) == so.iglob) {
          etab.push_back(e);
          itab.push_back(i);
          jtab.push_back(j);
        }
  int nel = (int)etab.size(), nterms = nel * 2 * n;
  so.mnode.assign(nterms, 0);
  so.mcoef.assign((size_t)nterms * ndof, 0.0);
  int t = 0;
  for (int k = 0; k < nel; ++k) {
    int e = etab[k], i = itab[k], j = jtab[k];
    double jac[4], ji[4], G[4] = {0, 0, 0, 0};
    SE_Jacobian(g, e, i, j, jac);
    invert2(jac, ji);  // jac_inv(r,c) at ji[r + 2*c]
    if (ndof == 2) {   // G = matmul(M, transpose(jac_inv)): G(a,b) = sum_k M(a,k)*jac_inv(b,k)
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) G[a + 2 * b] = so.M[a + 2 * 0] * ji[b + 2 * 0] + so.M[a + 2 * 1] * ji[b + 2 * 1];
    } else {           // G(:,1) = matmul(jac_inv, M(:,1))
      for (int a = 0; a < 2; ++a) G[a] = ji[a + 2 * 0] * so.M[0] + ji[a + 2 * 1] * so.M[1];
    }
    for (int q = 1; q <= n; ++q, ++t) {  // iglob_xi(:,k) = ibool(:,j,e); coef_xi(:,c,k) = G(c,1)*hprime(:,i)
      so.mnode[t] = g.ib(q, j, e);
      if (ndof == 2) {
        so.mcoef[t] = G[0 + 2 * 0] * g.H[(q - 1) + (size_t)n * (i - 1)];
        so.mcoef[t + (size_t)nterms] = G[1 + 2 * 0] * g.H[(q - 1) + (size_t)n * (i - 1)];
      } else {
        so.mcoef[t] = G[0] * g.H[(q - 1) + (size_t)n * (i - 1)];
      }
    }
    for (int q = 1; q <= n; ++q, ++t) {  // iglob_eta(:,k) = ibool(i,:,e); coef_eta(:,c,k) = G(c,2)*hprime(:,j)
      so.mnode[t] = g.ib(i, q, e);
      if (ndof == 2) {
        so.mcoef[t] = G[0 + 2 * 1] * g.H[(q - 1) + (size_t)n * (j - 1)];
        so.mcoef[t + (size_t)nterms] = G[1 + 2 * 1] * g.H[(q - 1) + (size_t)n * (j - 1)];
      } else {
        so.mcoef[t] = G[1] * g.H[(q - 1) + (size_t)n * (j - 1)];
      }
    }
  }
}

// src_gen.f90:290-317 SO_add with FORCE_add (src_force.f90:77-90) / SRC_MOMENT_add (src_moment.f90:183-197)
inline void SO_add(Problem& pb, double t, std::vector<double>& MxA) {
  for (auto& s : pb.src) {
    double ampli = s.stf.eval(t - s.tdelay);
    ampli = ampli * s.ampli;
    if (s.moment) {
      size_t nt = s.mnode.size();
      for (size_t q = 0; q < nt; ++q)
        for (int c = 0; c < pb.ndof; ++c)
          MxA[pb.idx(s.mnode[q], c)] = MxA[pb.idx(s.mnode[q], c)] + ampli * s.mcoef[q + nt * c];
      continue;
    }
    if (pb.ndof == 1) {
      MxA[pb.idx(s.iglob, 0)] = MxA[pb.idx(s.iglob, 0)] + ampli;
    } else {
      MxA[pb.idx(s.iglob, 0)] = MxA[pb.idx(s.iglob, 0)] + s.dir[0] * ampli;
      MxA[pb.idx(s.iglob, 1)] = MxA[pb.idx(s.iglob, 1)] + s.dir[1] * ampli;
    }
  }
}

// solver.f90:42-84 solve_Newmark, :89-128 solve_HHT_alpha, :140-160 solve_leapfrog, :169-199 solve_symplectic
inline void solve(Problem& pb) {
  size_t nn = pb.d.size();
  std::vector<double>&d = pb.d, &v = pb.v, &a = pb.a_, &f = pb.a_;
  double dt = pb.time.dt;
  if (pb.time.kind == "leapfrog") {
    for (size_t q = 0; q < nn; ++q) d[q] = d[q] + dt * v[q];
    compute_Fint(pb, f, d, v);
    SO_add(pb, pb.time.time, f);
    BC_apply(pb, f);
    for (size_t q = 0; q < nn; ++q) a[q] = pb.rmass[q] * f[q];
    for (size_t q = 0; q < nn; ++q) v[q] = v[q] + dt * a[q];
  } else if (pb.time.kind == "newmark") {
    double beta = pb.time.beta, gamma = pb.time.gamma;
    double c1 = (0.5 - beta) * dt * dt, c2 = (1.0 - gamma) * dt;
    for (size_t q = 0; q < nn; ++q) d[q] = d[q] + dt * v[q] + c1 * a[q];
    for (size_t q = 0; q < nn; ++q) v[q] = v[q] + c2 * a[q];
    compute_Fint(pb, f, d, v);
    SO_add(pb, pb.time.time, f);
    BC_apply(pb, f);
    for (size_t q = 0; q < nn; ++q) a[q] = f[q] * pb.rmass[q];
    double c3 = gamma * dt, c4 = beta * dt * dt;
    for (size_t q = 0; q < nn; ++q) v[q] = v[q] + c3 * a[q];
    for (size_t q = 0; q < nn; ++q) d[q] = d[q] + c4 * a[q];
  } else if (pb.time.kind == "HHT-alpha") {  // solver.f90:89-128
    double alpha = pb.time.alpha, beta = pb.time.beta, gamma = pb.time.gamma;
    std::vector<double> d_alpha = d, v_alpha = v;
    double c1 = (0.5 - beta) * dt * dt, c2 = (1.0 - gamma) * dt;
    for (size_t q = 0; q < nn; ++q) d[q] = d[q] + dt * v[q] + c1 * a[q];
    for (size_t q = 0; q < nn; ++q) v[q] = v[q] + c2 * a[q];
    for (size_t q = 0; q < nn; ++q) d_alpha[q] = alpha * d[q] + (1.0 - alpha) * d_alpha[q];
    for (size_t q = 0; q < nn; ++q) v_alpha[q] = alpha * v[q] + (1.0 - alpha) * v_alpha[q];
    compute_Fint(pb, f, d_alpha, v_alpha);
    double t_alpha = pb.time.time + (alpha - 1.0) * dt;
    SO_add(pb, t_alpha, f);
    double tmp = pb.time.time;
    pb.time.time = t_alpha;
    BC_apply(pb, f);
    pb.time.time = tmp;
    for (size_t q = 0; q < nn; ++q) a[q] = f[q] * pb.rmass[q];
    double c3 = gamma * dt, c4 = beta * dt * dt;
    for (size_t q = 0; q < nn; ++q) v[q] = v[q] + c3 * a[q];
    for (size_t q = 0; q < nn; ++q) d[q] = d[q] + c4 * a[q];
  } else if (pb.time.nstages > 0) {  // solve_symplectic (solver.f90:169-199): no boundary conditions
    double t = pb.time.time - dt;
    for (int k = 0; k < pb.time.nstages; ++k) {
      double ca = dt * pb.time.a[k];
      for (size_t q = 0; q < nn; ++q) d[q] = d[q] + ca * v[q];
      compute_Fint(pb, f, d, v);
      t = t + dt * pb.time.a[k];
      SO_add(pb, t, f);
      for (size_t q = 0; q < nn; ++q) a[q] = pb.rmass[q] * f[q];
      double cb = dt * pb.time.b[k];
      for (size_t q = 0; q < nn; ++q) v[q] = v[q] + cb * a[q];
    }
    double ca = dt * pb.time.a[pb.time.nstages];
    for (size_t q = 0; q < nn; ++q) d[q] = d[q] + ca * v[q];
  } else {
    IO_abort("solve: scheme not supported by the oracle: " + pb.time.kind);
  }
}

// energy.f90:49-106 kinetic energy
inline double energy_Ek(const Problem& pb) {
  const Grid& g = pb.grid;
  int n = g.ngll;
  double Ek = 0;
  std::vector<double> rho((size_t)n * n);
  for (int e = 1; e <= g.nelem; ++e) {
    pb.mat.get(pb.mat.rho, e, rho.data());
    double s = 0;
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i) {
        int ip = g.ib(i, j, e);
        double v2 = 0;
        for (int c = 0; c < pb.ndof; ++c) v2 += pb.v[pb.idx(ip, c)] * pb.v[pb.idx(ip, c)];
        s += SE_VolumeWeight(g, e, i, j) * rho[(i - 1) + (size_t)n * (j - 1)] * v2;
      }
    Ek += s;
  }
  return 0.5 * Ek;
}

// energy.f90:84-104: E_W = 1/2 sum over elements of sum(beta * d2), the additional elastic energy of a 2.5D run
inline double energy_EW(const Problem& pb) {
  if (pb.beta25d.empty()) return 0.0;
  const Grid& g = pb.grid;
  int n = g.ngll, n2 = n * n;
  double EW = 0;
  for (int e = 1; e <= g.nelem; ++e) {
    const double* beta = &pb.beta25d[(size_t)n2 * (pb.elem2set[e - 1] - 1)];
    double s = 0;
    for (int k = 0; k < n2; ++k) {
      int ip = g.ibool[(size_t)n2 * (e - 1) + k];
      double d2 = 0;
      for (int c = 0; c < pb.ndof; ++c) d2 += pb.d[pb.idx(ip, c)] * pb.d[pb.idx(ip, c)];
      s += beta[k] * d2;
    }
    EW += s;
  }
  return 0.5 * EW;
}

// PLOT_FIELD's element-wise snapshot fields (plot_gen.f90:239-300): 'E' strain (FIELD_strain_elem,
// fields.f90:192-237), 'S' stress (MAT_stress_dv, mat_gen.f90:626-641: Kelvin-Voigt d + eta*v, then MAT_ELAST_stress,
// mat_elastic.f90:804-839, isotropic), 'd' / 'c' divergence and curl of the velocity (FIELD_divcurl_elem,
// fields.f90:242-285).  out[(c*nelem + e-1)*n2 + k] as real (float32), c = 0..ndof for E/S, one component for d/c.
inline void snapshot_elem(const Problem& pb, char what, std::vector<float>& out) {
  const Grid& g = pb.grid;
  const int n = g.ngll, n2 = n * n, ndof = pb.ndof;
  const int ncomp = (what == 'E' || what == 'S') ? ndof + 1 : 1;
  out.assign((size_t)ncomp * g.nelem * n2, 0.f);
  std::vector<double> U((size_t)n2 * ndof), dxi((size_t)n2 * ndof), deta((size_t)n2 * ndof), la(n2), mu(n2);
  const std::vector<double>& fld = (what == 'd' || what == 'c') ? pb.v : pb.d;
  for (int e = 1; e <= g.nelem; ++e) {
    const int* ib = &g.ibool[(size_t)n2 * (e - 1)];
    for (int c = 0; c < ndof; ++c)
      for (int k = 0; k < n2; ++k) U[k + (size_t)n2 * c] = fld[pb.idx(ib[k], c)];
    if (what == 'S' && pb.elem2kv[e - 1] > 0) {
      const double* eta = &pb.kv_eta[(size_t)n2 * (pb.elem2kv[e - 1] - 1)];
      for (int c = 0; c < ndof; ++c)
        for (int k = 0; k < n2; ++k) U[k + (size_t)n2 * c] = U[k + (size_t)n2 * c] + eta[k] * pb.v[pb.idx(ib[k], c)];
    }
    for (int c = 0; c < ndof; ++c) {  // mxm(hTprime, U), mxm(U, hprime)
      mxm(g.Ht.data(), &U[(size_t)n2 * c], &dxi[(size_t)n2 * c], n);
      mxm(&U[(size_t)n2 * c], g.H.data(), &deta[(size_t)n2 * c], n);
    }
    if (what == 'S') {
      pb.mat.get(pb.mat.lambda, e, la.data());
      pb.mat.get(pb.mat.mu, e, mu.data());
    }
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i) {
        const int k = (i - 1) + n * (j - 1);
        double jac[4], xi[4];
        SE_Jacobian(g, e, i, j, jac);
        invert2(jac, xi);  // xjaci(1,1)=dxi_dx, (1,2)=dxi_dz, (2,1)=deta_dx, (2,2)=deta_dz, column-major
        const double dxi_dx = xi[0], deta_dx = xi[1], dxi_dz = xi[2], deta_dz = xi[3];
        double ev[3] = {0, 0, 0};
        if (what == 'E' || what == 'S') {
          if (ndof == 1) {
            ev[0] = 0.5 * (dxi[k] * dxi_dx + deta[k] * deta_dx);
            ev[1] = 0.5 * (dxi[k] * dxi_dz + deta[k] * deta_dz);
          } else {
            ev[0] = dxi[k] * dxi_dx + deta[k] * deta_dx;
            ev[1] = dxi[k + n2] * dxi_dz + deta[k + n2] * deta_dz;
            ev[2] = 0.5 * (dxi[k] * dxi_dz + deta[k] * deta_dz + dxi[k + n2] * dxi_dx + deta[k + n2] * deta_dx);
          }
          if (what == 'S' && !pb.elem2pl.empty() && pb.elem2pl[e - 1] > 0) {  // MAT_PLAST_stress, update = .false. (mat_plastic.f90:294,379-383)
            const double* ep = &pb.pl_ep[(size_t)3 * n2 * (pb.elem2pl[e - 1] - 1)];
            for (int c = 0; c < 3; ++c) ev[c] = ev[c] - ep[(size_t)c * n2 + k];
          }
          if (what == 'S') {
            if (ndof == 1) {
              ev[0] = 2.0 * mu[k] * ev[0];
              ev[1] = 2.0 * mu[k] * ev[1];
            } else {
              const double e1 = ev[0], e2 = ev[1];
              ev[0] = (la[k] + 2.0 * mu[k]) * e1 + la[k] * e2;
              ev[1] = la[k] * e1 + (la[k] + 2.0 * mu[k]) * e2;
              ev[2] = 2.0 * mu[k] * ev[2];
            }
          }
          for (int c = 0; c <= ndof; ++c) out[((size_t)c * g.nelem + (e - 1)) * n2 + k] = (float)ev[c];
        } else {
          const double dv = dxi[k] * dxi_dx + deta[k] * deta_dx + dxi[k + n2] * dxi_dz + deta[k + n2] * deta_dz;
          const double cu = dxi[k] * dxi_dz + deta[k] * deta_dz - dxi[k + n2] * dxi_dx - deta[k + n2] * deta_dx;
          out[(size_t)(e - 1) * n2 + k] = (float)(what == 'd' ? dv : cu);
        }
      }
  }
}

// ------------------------------------------------------------------------------------------
// main.f90:27-99: init_main (init.f90:16-131) then the time loop
inline void init_main(Problem& pb, const CartSpec& cart) {
  Grid& g = pb.grid;
  CART_build(cart, g);          // MESH_build
  SE_init_gll(g);               // SE_init
  SE_init_numbering(g);
  SE_init_coord(g);
  for (auto& b : g.bnds) BC_set_bulk_node(b, g);  // SE_BcTopoInit
  MAT_init_prop(pb, g.ngll);
  pb.grid_cfl = CHECK_grid(pb);
  TIME_init(pb.time, pb.grid_cfl);
  MAT_init_work(pb);
  size_t nn = (size_t)g.npoin * pb.ndof;
  pb.d.assign(nn, 0.0);
  pb.v.assign(nn, 0.0);
  pb.a_.assign(nn, 0.0);
  MAT_MASS_init(pb);
  pb.mass.assign(pb.rmass.begin(), pb.rmass.begin() + g.npoin);
  BC_init(pb);
  REC_init(pb);
  for (auto& s : pb.src) {  // SO_init (src_gen.f90:216-260)
    s.iglob = SE_find_nearest_node(g, s.coord[0], s.coord[1]);
    s.coord[0] = g.coord[2 * (size_t)(s.iglob - 1)];
    s.coord[1] = g.coord[2 * (size_t)(s.iglob - 1) + 1];
    if (s.moment) SRC_MOMENT_init(s, pb);
  }
  for (size_t q = 0; q < nn; ++q) pb.rmass[q] = 1.0 / pb.rmass[q];  // init.f90:112-116
  pb.time.time = 0.0;
  pb.it = 0;
  REC_store(pb, 0);  // main.f90:35
}

inline void step(Problem& pb) {  // body of the do-loop main.f90:51-99
  pb.it++;
  pb.time.time = pb.it * pb.time.dt;
  solve(pb);
  REC_store(pb, pb.it);
  BC_write(pb, pb.it);
}

// ------------------------------------------------------------------------------------------
// input.f90:12-63 read_main for the supported subset
inline void read_main(Problem& pb, CartSpec& cart, ParInp& in) {
  // GENERAL (input.f90:101-139)
  in.rewind();
  const NmlGroup* g = in.next("GENERAL");
  if (!g) IO_abort("Input: GENERAL parameters not found");
  pb.ndof = g->integer("ndof", 2);
  pb.grid.ngll = g->integer("ngll", 9);
  pb.grid.fmax = g->dbl("fmax", 1.0);
  pb.grid.W = g->dbl("W", HUGE_D);
  if (pb.grid.W <= 0.0) IO_abort("GENERAL input block: W must be positive");
  // MESH_DEF / MESH_CART (mesh_gen.f90, mesh_cartesian.f90:88-214)
  in.rewind();
  g = in.next("MESH_DEF");
  if (!g || upper(g->str("method", "")) != "CARTESIAN") IO_abort("oracle: only method='CARTESIAN' is supported");
  in.rewind();
  g = in.next("MESH_CART");
  if (!g) IO_abort("CART_read: input block not found");
  cart.xmin = g->dbl("xlim", 0.0, 0);
  cart.xmax = g->dbl("xlim", HUGE_D, 1);
  cart.zmin = g->dbl("zlim", 0.0, 0);
  cart.zmax = g->dbl("zlim", HUGE_D, 1);
  cart.nx = g->integer("nelem", 0, 0);
  cart.nz = g->integer("nelem", 0, 1);
  cart.ezflt = g->integer("ezflt", 0);
  if (g->logical("FaultX", false)) cart.ezflt = -1;
  if (cart.ezflt >= cart.nz) IO_abort("CART_read: ezflt must be < nelem(2)");
  if (cart.ezflt == -1) cart.ezflt = cart.nz / 2;
  cart.fztag = g->integer("fztag", 0);
  cart.fznz = g->integer("fznz", 1);
  cart.split = g->logical("split", false);
  cart.splitD = g->dbl("splitD", HUGE_D);
  in.rewind();
  int ndom = in.count("MESH_CART_DOMAIN");
  if (ndom == 0) {
    cart.domains.push_back({1, {1, cart.nx}, {1, cart.nz}});
  } else {
    in.rewind();
    for (int i = 0; i < ndom; ++i) {
      g = in.next("MESH_CART_DOMAIN");
      CartSpec::Dom dm;
      dm.tag = g->integer("tag", 0);
      dm.ex[0] = g->integer("ex", 0, 0);
      dm.ex[1] = g->integer("ex", 0, 1);
      dm.ez[0] = g->integer("ez", 0, 0);
      dm.ez[1] = g->integer("ez", 0, 1);
      cart.domains.push_back(dm);
    }
  }
  // TIME (time.f90:122-318)
  {
    TimeScheme& t = pb.time;
    in.rewind();
    g = in.next("TIME");
    if (!g) IO_abort("TIME parameters not found");
    t.kind = g->str("kind", "leapfrog");
    int NbSteps = g->integer("NbSteps", 0);
    double dt = g->dbl("dt", 0.0), courant = g->dbl("courant", 0.5), TotalTime = g->dbl("TotalTime", 0.0);
    if (courant < 0.0 || courant > (double)0.6f) IO_abort("TIME: Courant out of range [0,0.6]");
    if (dt > 0.0) {
      if (TotalTime > 0.0) NbSteps = (int)std::ceil(TotalTime / dt);
      TotalTime = dt * NbSteps;
    }
    t.nt = NbSteps;
    t.dt = dt;
    t.courant = courant;
    t.total = TotalTime;
    t.alpha = 1.0;
    t.beta = 0.0;
    t.gamma = 0.5;
    if (t.kind == "newmark") {
      const NmlGroup* gn = in.next("TIME_NEWMARK");
      if (gn) {
        t.beta = gn->dbl("beta", 0.0);
        t.gamma = gn->dbl("gamma", 0.5);
      }
    } else if (t.kind == "HHT-alpha") {  // time.f90:232-246
      double alpha = 0.5, rho = 0.5;
      const NmlGroup* gh = in.next("TIME_HHTA");
      if (gh) {
        alpha = gh->dbl("alpha", 0.5);
        rho = gh->dbl("rho", 0.5);
      }
      if (alpha < 0.0 || alpha > 1.0) IO_abort("TIME_HHTA: alpha is out of range [0,1]");
      if (rho < 0.5 || rho > 1.0) IO_abort("TIME_HHTA: rho is out of range [0.5,1]");
      t.alpha = alpha;
      t.gamma = 1.5 - alpha;
      if (alpha != 1.0) t.beta = 1.0 - alpha - rho * rho * (rho - 1.0) / ((1.0 - alpha) * ((1.0 + rho) * (1.0 + rho) * (1.0 + rho)));
      else t.beta = 0.0;
    } else if (t.kind == "symp_PV") {  // time.f90:248-255
      t.nstages = 1;
      t.a = {0.5, 0.5};
      t.b = {1.0};
    } else if (t.kind == "symp_PFR") {  // time.f90:257-269
      t.nstages = 3;
      double theta = 1.0 / (2.0 - std::pow(2.0, 1.0 / 3.0));
      t.a = {theta / 2.0, (1.0 - theta) / 2.0, (1.0 - theta) / 2.0, theta / 2.0};
      t.b = {theta, 1.0 - 2.0 * theta, theta};
    } else if (t.kind == "symp_PEFRL") {  // time.f90:271-287
      t.nstages = 4;
      double xi = 0.1786178958448091, lambda = -0.2123418310626054, chi = -0.06626458266981849;
      t.a = {xi, chi, 1.0 - 2.0 * (chi + xi), chi, xi};
      t.b = {0.5 - lambda, lambda, lambda, 0.5 - lambda};
    } else if (t.kind != "leapfrog") {
      IO_abort("oracle: time scheme not supported: " + t.kind);
    }
  }
  // MATERIAL (mat_gen.f90:101-189)
  {
    in.rewind();
    int numat = in.count("MATERIAL");
    if (numat == 0) IO_abort("MAT_read: MATERIAL block not found");
    pb.mat.inputs.assign(numat, MatInput());
    for (int i = 1; i <= numat; ++i) {
      in.rewind();
      const NmlGroup* gm = nullptr;
      for (int j = 1; j <= i; ++j) gm = in.next("MATERIAL");
      int tag = gm->integer("tag", 0);
      if (tag <= 0 || tag > numat) IO_abort("MAT_read: inconsistent or missing tags");
      std::string kinds[2] = {upper(gm->str("kind", "ELAST", 0)), upper(gm->str("kind", "", 1))};
      MatInput& mi = pb.mat.inputs[tag - 1];
      for (int k = 0; k < 2; ++k) {
        if (kinds[k] == "ELAST") {  // MAT_ELAST_read (mat_elastic.f90:71-184)
          const NmlGroup* ge = in.next("MAT_ELASTIC");
          if (!ge) IO_abort("MAT_ELAST_read: MAT_ELASTIC input block not found");
          mi.elastic = true;
          double rho = ge->dbl("rho", 0.0), cp = ge->dbl("cp", 0.0), cs = ge->dbl("cs", 0.0);
          std::string rhoH = ge->str("rhoH", ""), cpH = ge->str("cpH", ""), csH = ge->str("csH", "");
          if (rho <= 0.0 && rhoH == "") IO_abort("MAT_ELAST_read: undefined density (rho)");
          mi.rho = read_cd(in, rho, rhoH);
          if ((cp > 0.0 || cpH != "") && (cs > 0.0 || csH != "")) {
            mi.isotropic = true;
            mi.cp = read_cd(in, cp, cpH);
            mi.cs = read_cd(in, cs, csH);
            if (cp > 0.0 && cs > 0.0 && rho > 0.0) {
              mi.homogeneous = true;
              mi.mu = rho * cs * cs;
              mi.lambda = rho * (cp * cp - 2.0 * cs * cs);
              mi.has_lambda = true;
            }
          } else {
            IO_abort("oracle: anisotropic MAT_ELASTIC not supported");
          }
        } else if (kinds[k] == "KV") {  // MAT_KV_read (mat_kelvin_voigt.f90:68-107)
          const NmlGroup* gk = in.next("MAT_KV");
          if (!gk) IO_abort("MAT_KV_read: MAT_KV input block not found");
          mi.kv = true;
          mi.eta = read_cd(in, gk->dbl("eta", 0.0), gk->str("etaH", ""));
          mi.etaxdt = gk->logical("ETAxDT", true);
        } else if (kinds[k] == "PLAST") {  // MAT_PLAST_read (mat_plastic.f90:66-118)
          const NmlGroup* gp = in.next("MAT_PLASTIC");
          if (!gp) IO_abort("MAT_PLAST_read: MAT_PLASTIC input block not found");
          mi.plastic = true;
          mi.isotropic = true;
          mi.rho = read_cd(in, gp->dbl("rho", 0.0), "");
          mi.cp = read_cd(in, gp->dbl("cp", 0.0), "");
          mi.cs = read_cd(in, gp->dbl("cs", 0.0), "");
          mi.phi = gp->dbl("phi", 0.0);
          mi.coh = gp->dbl("coh", 0.0);
          mi.Tv = gp->dbl("Tv", 0.0);
          for (int q = 0; q < 3; ++q) mi.e0[q] = gp->dbl("e0", 0.0, q);
        } else if (kinds[k] == "DMG") {  // MAT_DMG_read (mat_damage.f90:108-173)
          const NmlGroup* gd = in.next("MAT_DAMAGE");
          if (!gd) IO_abort("MAT_DMG_read: MAT_DAMAGE input block not found");
          mi.damage = true;
          mi.isotropic = true;
          mi.rho = read_cd(in, gd->dbl("rho", 0.0), "");
          mi.cp = read_cd(in, gd->dbl("cp", 0.0), "");
          mi.cs = read_cd(in, gd->dbl("cs", 0.0), "");
          mi.phi = gd->dbl("phi", 0.0);
          mi.alpha0 = gd->dbl("alpha", 0.0);
          mi.Cd = gd->dbl("Cd", 0.0);
          mi.beta_dmg = gd->dbl("beta", 0.0);
          mi.Rdmg = gd->dbl("R", 0.0);
          for (int q = 0; q < 3; ++q) mi.e0[q] = gd->dbl("e0", 0.0, q);
          for (int q = 0; q < 3; ++q) mi.ep0[q] = gd->dbl("ep", 0.0, q);
        } else if (kinds[k] == "VISCO") {  // MAT_VISCO_read (mat_visco.f90:65-113)
          const NmlGroup* gv = in.next("MAT_VISCO");
          if (!gv) IO_abort("MAT_VISCO_read: MAT_VISCO input block not found");
          mi.visco = true;
          mi.isotropic = true;
          mi.rho = read_cd(in, gv->dbl("rho", 0.0), "");
          mi.cp = read_cd(in, gv->dbl("cp", 0.0), "");
          mi.cs = read_cd(in, gv->dbl("cs", 0.0), "");
          mi.QP = gv->dbl("QP", 0.0);
          mi.QS = gv->dbl("QS", 0.0);
          mi.Nbody = (int)gv->dbl("Nbody", 0.0);
          mi.fmin = gv->dbl("fmin", 0.0);
          mi.fmax = gv->dbl("fmax", 0.0);
          if (mi.Nbody < 1 || mi.Nbody > 8) IO_abort("oracle: MAT_VISCO Nbody must be in 1..8");
        } else if (kinds[k] == "") {
        } else {
          IO_abort("oracle: material kind not supported: " + kinds[k]);
        }
        // reposition right after the i-th MATERIAL block
        in.rewind();
        for (int j = 1; j <= i; ++j) in.next("MATERIAL");
      }
    }
  }
  // BC_DEF (bc_gen.f90:83-187)
  {
    in.rewind();
    int nbc = in.count("BC_DEF");
    pb.bc.clear();
    pb.bc.resize(nbc);
    for (int i = 1; i <= nbc; ++i) {
      in.rewind();
      const NmlGroup* gb = nullptr;
      for (int j = 1; j <= i; ++j) gb = in.next("BC_DEF");
      Bc& b = pb.bc[i - 1];
      int tag = gb->integer("tag", 0);
      if (tag > 0) {
        b.tag[0] = tag;
      } else if (gb->integer("tags", 0, 0) > 0) {
        b.tag[0] = gb->integer("tags", 0, 0);
        b.tag[1] = gb->integer("tags", 0, 1);
      } else {
        IO_abort("bc_read: tag(s) are null or not set");
      }
      std::string kind = upper(gb->str("kind", " "));
      if (kind == "ABSORB") {  // BC_ABSO_read (bc_abso.f90:70-108)
        b.kind = IS_ABSORB;
        b.abso.reset(new BcAbso());
        const NmlGroup* ga = in.next("BC_ABSORB");
        b.abso->stacey = ga ? ga->logical("stacey", false) : false;
        b.abso->let_wave = ga ? ga->logical("let_wave", true) : true;
      } else if (kind == "PERIOD") {  // BC_PERIO_read (bc_periodic.f90:29-40): no parameters
        b.kind = IS_PERIOD;
        b.perio.reset(new BcPerio());
        if (b.tag[1] == 0) IO_abort("bc_read: PERIOD needs tags = master, slave");
      } else if (kind == "DIRNEU") {  // bc_DIRNEU_read (bc_dirneu.f90:50-112)
        b.kind = IS_DIRNEU;
        b.dirneu.reset(new BcDirneu());
        const NmlGroup* gd = in.next("BC_DIRNEU");
        if (!gd) IO_abort("bc_DIRNEU_read: no BC_DIRNEU block found");
        std::string h = upper(gd->str("h", "N")), v = upper(gd->str("v", "N"));
        b.dirneu->kind[0] = (h == "D") ? 2 : 1;
        b.dirneu->kind[1] = (v == "D") ? 2 : 1;
        if (gd->str("hstf", "none") != "none" || gd->str("vstf", "none") != "none")
          IO_abort("oracle: time-dependent Neumann not supported");
      } else if (kind == "DYNFLT") {  // BC_DYNFLT_read (bc_dynflt.f90:104-229)
        b.kind = IS_DYNFLT;
        b.dynflt.reset(new BcDynflt());
        BcDynflt& f = *b.dynflt;
        const NmlGroup* gf = in.next("BC_DYNFLT");
        if (!gf) IO_abort("BC_DYNFLT_read: BC_DYNFLT input block not found");
        f.ot1 = gf->dbl("ot1", 0.0);
        f.odt = gf->dbl("otd", 0.0);
        f.oix1 = gf->integer("oxi", 1, 0);
        f.oixn = gf->integer("oxi", std::numeric_limits<int>::max(), 1);
        f.oixd = gf->integer("oxi", 1, 2);
        f.osides = gf->logical("osides", false);
        if (f.osides) IO_abort("oracle: osides not supported");
        f.in_cohesion = read_cd(in, gf->dbl("cohesion", 0.0), gf->str("cohesionH", ""));
        f.in_N = read_cd(in, gf->dbl("Tn", 0.0), gf->str("TnH", ""));
        f.in_T = read_cd(in, gf->dbl("Tt", 0.0), gf->str("TtH", ""));
        f.in_Sxx = read_cd(in, gf->dbl("Sxx", 0.0), gf->str("SxxH", ""));
        f.in_Sxy = read_cd(in, gf->dbl("Sxy", 0.0), gf->str("SxyH", ""));
        f.in_Sxz = read_cd(in, gf->dbl("Sxz", 0.0), gf->str("SxzH", ""));
        f.in_Syz = read_cd(in, gf->dbl("Syz", 0.0), gf->str("SyzH", ""));
        f.in_Szz = read_cd(in, gf->dbl("Szz", 0.0), gf->str("SzzH", ""));
        f.in_V = read_cd(in, gf->dbl("V", 1e-12), gf->str("VH", ""));
        f.allow_opening = gf->logical("opening", true);
        std::string fr[2] = {upper(gf->str("friction", "SWF", 0)), upper(gf->str("friction", "", 1))};
        for (int k = 0; k < 2; ++k) {
          if (fr[k] == "SWF") {  // swf_read (bc_dynflt_swf.f90:56-117)
            f.swf.reset(new Swf());
            const NmlGroup* gs = in.next("BC_DYNFLT_SWF");
            NmlGroup empty;
            if (!gs) gs = &empty;
            f.swf->kind = gs->integer("kind", 1);
            f.swf->healing = gs->logical("healing", false);
            f.swf->in_dc = read_cd(in, gs->dbl("Dc", 0.5), gs->str("DcH", ""));
            f.swf->in_mus = read_cd(in, gs->dbl("MuS", 0.6), gs->str("MuSH", ""));
            f.swf->in_mud = read_cd(in, gs->dbl("MuD", 0.5), gs->str("MuDH", ""));
            f.swf->in_alpha = read_cd(in, gs->dbl("alpha", 0.0), gs->str("alphaH", ""));
            f.swf->in_p = read_cd(in, gs->dbl("p", 3.0), gs->str("pH", ""));
          } else if (fr[k] == "RSF") {  // rsf_read (bc_dynflt_rsf.f90:63-135)
            f.rsf.reset(new Rsf());
            const NmlGroup* gs = in.next("BC_DYNFLT_RSF");
            NmlGroup empty;
            if (!gs) gs = &empty;
            f.rsf->kind = gs->integer("kind", 1);
            f.rsf->in_dc = read_cd(in, gs->dbl("Dc", 0.5), gs->str("DcH", ""));
            f.rsf->in_mus = read_cd(in, gs->dbl("MuS", 0.6), gs->str("MuSH", ""));
            f.rsf->in_a = read_cd(in, gs->dbl("a", 0.01), gs->str("aH", ""));
            f.rsf->in_b = read_cd(in, gs->dbl("b", 0.02), gs->str("bH", ""));
            f.rsf->in_Vstar = read_cd(in, gs->dbl("Vstar", 1.0), gs->str("VstarH", ""));
            f.rsf->in_theta = read_cd(in, gs->dbl("theta", 0.0), gs->str("thetaH", ""));
            f.rsf->in_Vc = read_cd(in, gs->dbl("Vc", 1e-6), gs->str("VcH", ""));
          } else if (fr[k] == "TWF") {  // twf_read (bc_dynflt_twf.f90:51-113)
            f.twf.reset(new Twf());
            const NmlGroup* gs = in.next("BC_DYNFLT_TWF");
            NmlGroup empty;
            if (!gs) gs = &empty;
            Twf& t = *f.twf;
            t.kind = gs->integer("kind", 1);
            t.mus = gs->dbl("mus", 0.6);
            t.mud = gs->dbl("mud", 0.5);
            t.mu0 = gs->dbl("mu0", 0.6);
            t.X = gs->dbl("X", 0.0);
            t.Z = gs->dbl("Z", 0.0);
            t.V = gs->dbl("V", 1e3);
            t.L = gs->dbl("L", 1.0);
            t.T = gs->dbl("T", HUGE_D);
            t.Dc = gs->dbl("Dc", HUGE_D);
          } else if (fr[k] == "") {
          } else {
            IO_abort("BC_DYNFLT: invalid friction");
          }
        }
        {  // normal_read (bc_dynflt_normal.f90:41-92)
          const NmlGroup* gn = in.next("BC_DYNFLT_NOR");
          NmlGroup empty;
          if (!gn) gn = &empty;
          f.normal.kind = gn->integer("kind", 1);
          f.normal.L = gn->dbl("L", 1.0);
          f.normal.V = gn->dbl("V", 1.0);
          f.normal.T = gn->dbl("T", 1.0);
        }
      } else {
        IO_abort("oracle: boundary condition kind not supported: " + kind);
      }
    }
  }
  // SRC_DEF (src_gen.f90:75-212), STF_RICKER (stf_ricker.f90:45-87), SRC_FORCE (src_force.f90:36-62)
  {
    in.rewind();
    const NmlGroup* gs = in.next("SRC_DEF");
    pb.src.clear();
    if (gs) {
      if (upper(gs->str("stf", " ")) != "RICKER") IO_abort("oracle: only stf='RICKER' supported");
      std::string mech = upper(gs->str("mechanism", " "));
      if (mech != "FORCE" && mech != "EXPLOSION" && mech != "DOUBLE_COUPLE" && mech != "MOMENT")
        IO_abort("oracle: mechanism not supported: " + mech);
      Source s;
      s.coord[0] = gs->dbl("coord", HUGE_D, 0);
      s.coord[1] = gs->dbl("coord", HUGE_D, 1);
      const NmlGroup* gr = in.next("STF_RICKER");
      if (!gr) IO_abort("RICKER_read: STF_RICKER input block not found");
      s.stf.f0 = gr->real_as_dbl("f0", 0.0);
      s.stf.t0 = gr->real_as_dbl("onset", 0.0);
      s.stf.ampli = gr->real_as_dbl("ampli", 1.0);
      in.rewind();
      if (mech == "FORCE") {
        const NmlGroup* gfo = in.next("SRC_FORCE");
        double angle = gfo ? gfo->dbl("angle", 0.0) : 0.0;
        angle = angle * PI / 180.0;
        s.dir[0] = -std::sin(angle);
        s.dir[1] = std::cos(angle);
      } else {  // SRC_MOMENT_read (src_moment.f90:26-104); M(2,ndof) col-major
        s.moment = true;
        int ndof = pb.ndof;
        if (mech == "EXPLOSION") {
          if (ndof != 2) IO_abort("SRC_MOMENT_read: explosion only allowed in PSV (ndof=2)");
          s.M[0] = 1.0; s.M[1] = 0.0; s.M[2] = 0.0; s.M[3] = 1.0;
        } else if (mech == "DOUBLE_COUPLE") {
          const NmlGroup* gd = in.next("SRC_DOUBLE_COUPLE");
          if (!gd) IO_abort("SRC_MOMENT_read: SRC_DOUBLE_COUPLE input block not found");
          double dip = gd->dbl("dip", 90.0) * PI / 180.0;
          double n1 = std::sin(dip), n2 = std::cos(dip);
          if (ndof == 2) {
            double r1 = -std::cos(dip), r2 = std::sin(dip);
            s.M[0] = 2.0 * r1 * n1;           // M(1,1)
            s.M[2] = r1 * n2 + r2 * n1;       // M(1,2)
            s.M[1] = s.M[2];                  // M(2,1)
            s.M[3] = 2.0 * r2 * n2;           // M(2,2)
          } else {
            s.M[0] = n1;
            s.M[1] = n2;
          }
        } else {
          const NmlGroup* gm = in.next("SRC_MOMENT");
          if (!gm) IO_abort("SRC_MOMENT_read: SRC_MOMENT input block not found");
          if (ndof == 2) {
            s.M[0] = gm->dbl("Mxx", 0.0);
            s.M[2] = gm->dbl("Mxz", 0.0);
            s.M[1] = gm->dbl("Mzx", 0.0);
            s.M[3] = gm->dbl("Mzz", 0.0);
          } else {
            s.M[0] = gm->dbl("Myx", 0.0);
            s.M[1] = gm->dbl("Myz", 0.0);
          }
        }
      }
      pb.src.push_back(s);
    }
  }
  // REC_LINE (receivers.f90:62-162)
  {
    in.rewind();
    const NmlGroup* gr = in.next("REC_LINE");
    Receivers& r = pb.rec;
    r.present = false;
    if (gr) {
      r.present = true;
      int number = gr->integer("number", 0);
      r.isamp = gr->integer("isamp", 1);
      r.field = upper(gr->str("field", "V"))[0];
      r.AtNode = gr->logical("AtNode", true);
      if (gr->str("file", "none") != "none") IO_abort("oracle: receivers from file not supported");
      double first[2] = {gr->dbl("first", HUGE_D, 0), gr->dbl("first", HUGE_D, 1)};
      double last[2] = {gr->dbl("last", HUGE_D, 0), gr->dbl("last", HUGE_D, 1)};
      r.nx = number;
      r.coord.resize((size_t)2 * number);
      if (number > 1) {
        for (int i = 1; i <= number; ++i)
          for (int c = 0; c < 2; ++c) r.coord[2 * (i - 1) + c] = first[c] + (i - 1) / (double)(number - 1) * (last[c] - first[c]);
      } else {
        r.coord[0] = first[0];
        r.coord[1] = first[1];
      }
    }
  }
}

}  // namespace orc
