// Device-side structured builder: MESH_CART problems are generated directly in HBM.
//
// Restates, for the flat Cartesian Q4 box, what the reference's init does on the host:
//   CART_build            SRC/mesh_cartesian.f90:219-314 (+ mesh_structured.f90:12-196), natural
//                         element order (the OPT_RENUMBER=.false. setting of constants.f90:11)
//   SE_init_numbering     SRC/spec_grid.f90:198-314   -> closed-form node ids (cart_node_id)
//   MAT_ELAST_init_a      SRC/mat_elastic.f90:290-360 -> flat-grid planes (nelast 2 | 6)
//   MAT_MASS_init         SRC/mat_mass.f90:29-61
//   CHECK_grid/TIME_init  SRC/init.f90:145-289, SRC/time.f90:323-341 -> dt from the Courant number
//   BC_ABSO_init          SRC/bc_abso.f90:115-266     (flat sides, P1 or Stacey)
//   BC_DYNFLT_init        SRC/bc_dynflt.f90:231-520   (two-sided fault on tags 5,6, linear SWF)
// and lays the result out for the CTA-patch kernel: rectangular tiles of elements, congruent tiles
// sharing one cache-resident shape table, halo slots addressed in closed form.
// The heterogeneous material is the counter-based hash model of the synthetic benchmark
// (SURVEY.md 8d): values depend only on the global GLL lattice coordinates, so x-strips of one
// global mesh built on different GPUs agree bit for bit.
#include <algorithm>
#include <array>
#include <map>

#include "engine.hpp"
#include "rcm_box.hpp"

namespace s2d {

// ---------------------------------------------------------------------------------------------
// hash material (the benchmark's stand-in for a user-supplied heterogeneous model)
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ inline double hash_u(uint64_t seed, uint64_t ix, uint64_t iz, uint64_t k) {
  uint64_t h = splitmix64(seed ^ splitmix64(ix * 0x9E3779B97F4A7C15ull + k) ^
                          splitmix64(iz * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull * (k + 1)));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

struct CartGeom {
  int N, ndof, nx, nz, ezflt;
  double x0, z0, hx, hz;
  uint64_t seed;
  long long ix0, iz0;
  double rho, cp, cs;
  int halo_left, halo_right;
  StripGeom S;  // lattice / strip decomposition of the z-marching kernel
  double xgll[10], wgll[10];
  // caller-supplied material (s2d_cart_set_material): rho, cp, cs at every GLL point of every element,
  // (ngll,ngll,nelem) in natural element order; host pointers in the host copy of this struct, device pointers
  // in the copy the kernels receive (CartState::dev_geom)
  const double* mat[3];
};

__host__ __device__ inline bool row_detached(const CartGeom& G, int iz) {  // no element below shares nodes
  return iz == 0 || (G.ezflt > 0 && iz == G.ezflt);
}
// new GLL nodes created by the first / any other element of a row (see the header comment)
__host__ __device__ inline long long c_first(const CartGeom& G, bool B) {
  const int m = G.N - 2;
  return (long long)m * m + (long long)m * ((B ? 1 : 0) + 3) + (B ? 2 : 0) + 2;
}
__host__ __device__ inline long long c_other(const CartGeom& G, bool B) {
  const int m = G.N - 2;
  return (long long)m * m + (long long)m * ((B ? 1 : 0) + 2) + (B ? 1 : 0) + 1;
}
__host__ __device__ inline long long row_count(const CartGeom& G, bool B) {
  return c_first(G, B) + (long long)(G.nx - 1) * c_other(G, B);
}
__host__ __device__ inline long long elem_base(const CartGeom& G, int ix, int iz) {
  long long nB = 0;
  if (iz > 0) nB = 1 + ((G.ezflt > 0 && iz > G.ezflt) ? 1 : 0);
  long long base = nB * row_count(G, true) + ((long long)iz - nB) * row_count(G, false);
  const bool B = row_detached(G, iz);
  if (ix > 0) base += c_first(G, B) + (long long)(ix - 1) * c_other(G, B);
  return base;
}
__host__ __device__ inline long long cart_npoin(const CartGeom& G) {
  const long long nB = 1 + (G.ezflt > 0 ? 1 : 0);
  return nB * row_count(G, true) + ((long long)G.nz - nB) * row_count(G, false);
}
// Global node id (1-based) of GLL point (i,j) (1-based) of element (ix,iz) (0-based) under the
// traversal of SE_init_numbering (spec_grid.f90:249-287) in natural element order: the point is
// first moved to the earliest element that contains it, where it is "new"; there the ids run
// interior (i fastest), then the new edges in the order D,R,U,L with their counter-clockwise
// interior points, then the new vertices SW,SE,NE,NW.
__host__ __device__ inline long long cart_node_id(const CartGeom& G, int ix, int iz, int i, int j) {
  const int N = G.N;
  if (i == 1 && ix > 0) {
    ix -= 1;
    i = N;
  }
  if (j == 1 && !row_detached(G, iz)) {
    iz -= 1;
    j = N;
  }
  const bool B = row_detached(G, iz);
  const bool Lf = (ix == 0);
  const int m = N - 2;
  long long o;
  const bool ii = (i > 1 && i < N), jj = (j > 1 && j < N);
  if (ii && jj) {
    o = (i - 2) + (long long)m * (j - 2);
  } else {
    o = (long long)m * m;
    const long long eD = o, eR = eD + (B ? m : 0), eU = eR + m, eL = eU + m, vs = eL + (Lf ? m : 0);
    if (jj) {                    // vertical edges
      if (i == N) o = eR + (j - 2);
      else o = eL + (N - 1 - j);           // L edge, counter-clockwise = j descending
    } else if (ii) {             // horizontal edges
      if (j == 1) o = eD + (i - 2);
      else o = eU + (N - 1 - i);           // U edge, counter-clockwise = i descending
    } else {                     // vertices SW,SE,NE,NW
      const long long vSW = vs, vSE = vSW + ((Lf && B) ? 1 : 0), vNE = vSE + (B ? 1 : 0), vNW = vNE + 1;
      if (i == 1 && j == 1) o = vSW;
      else if (i == N && j == 1) o = vSE;
      else if (i == N && j == N) o = vNE;
      else o = vNW;
    }
  }
  return elem_base(G, ix, iz) + o + 1;
}
// Internal (device) numbering: the GLL lattice, row-major, the split fault row stored twice.
// 1-based like every node id that crosses the engine's add_* calls.  (i,j) 0-based here.
__host__ __device__ inline long long cart_lat_id(const CartGeom& G, int ix, int iz, int i, int j) {
  return (long long)strip_lat_row(G.S, iz, j) * G.S.LXP + (long long)ix * (G.N - 1) + i + 1;
}

// material at GLL point (i,j) (0-based) of element (ix,iz)
__host__ __device__ inline void cart_material(const CartGeom& G, int ix, int iz, int i, int j, double& rho,
                                              double& cp, double& cs) {
  if (G.mat[0]) {  // what MAT_getProp returns once MAT_init_prop has run (mat_gen.f90:204-303)
    const size_t q = ((size_t)iz * G.nx + ix) * (G.N * G.N) + (size_t)j * G.N + i;
    rho = G.mat[0][q];
    cp = G.mat[1][q];
    cs = G.mat[2][q];
    return;
  }
  if (G.seed == 0) {
    rho = G.rho;
    cp = G.cp;
    cs = G.cs;
    return;
  }
  const uint64_t gx = (uint64_t)(G.ix0 + (long long)ix * (G.N - 1) + i);
  const uint64_t gz = (uint64_t)(G.iz0 + (long long)iz * (G.N - 1) + j);
  const double u1 = hash_u(G.seed, gx, gz, 1), u2 = hash_u(G.seed, gx, gz, 2), u3 = hash_u(G.seed, gx, gz, 3);
  cs = 3464.0 * (1.0 + 0.10 * u1);
  cp = 1.7321 * cs * (1.0 + 0.02 * u2);
  rho = 2670.0 * (1.0 + 0.05 * u3);
}

// flat-grid coefficient planes a(i,j,plane) (mat_elastic.f90:323-358)
__host__ __device__ inline void cart_planes(const CartGeom& G, double rho, double cp, double cs, int i, int j,
                                            int nelast, double* av) {
  const double DxiDx = 2.0 / G.hx, DetaDz = 2.0 / G.hz;
  const double det = (0.5 * G.hx) * (0.5 * G.hz);
  const double la = rho * (cp * cp - 2.0 * cs * cs);
  const double mu = rho * cs * cs;
  const double weights = det * (G.wgll[i] * G.wgll[j]);
  if (nelast == 2) {
    av[0] = mu * DxiDx * DxiDx;
    av[1] = mu * DetaDz * DetaDz;
  } else {
    const double Kx = la + 2.0 * mu;
    av[0] = Kx * DxiDx * DxiDx;
    av[1] = la * DxiDx * DetaDz;
    av[2] = Kx * DetaDz * DetaDz;
    av[3] = mu * DetaDz * DetaDz;
    av[4] = mu * DxiDx * DetaDz;
    av[5] = mu * DxiDx * DxiDx;
  }
  for (int pl = 0; pl < nelast; ++pl) av[pl] = -weights * av[pl];
}

// ---------------------------------------------------------------------------------------------
// kernels
// coefficient planes in the strip layout of strip_kernels.cuh; one thread per GLL point of an element
template <typename T>
__global__ void k_cart_coef(CartGeom G, T* __restrict__ coef, int nelast, int compact) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N;
  const long long total = (long long)G.nx * G.nz * N2;
  if (w >= total) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2);
  const int i = k % N, j = k / N;
  const int ix = (int)(e % G.nx), iz = (int)(e / G.nx);
  double rho, cp, cs, av[6];
  cart_material(G, ix, iz, i, j, rho, cp, cs);
  if (compact) {  // (lambda, mu) only: the strip kernel forms the planes (strip_kernels.cuh)
    coef[strip_coef_index(G.S, 2, ix, iz, i, j, 0)] = (T)(rho * (cp * cp - 2.0 * cs * cs));
    coef[strip_coef_index(G.S, 2, ix, iz, i, j, 1)] = (T)(rho * cs * cs);
    return;
  }
  cart_planes(G, rho, cp, cs, i, j, nelast, av);
  for (int pl = 0; pl < nelast; ++pl) coef[strip_coef_index(G.S, nelast, ix, iz, i, j, pl)] = (T)av[pl];
}

// 2.5D term (MAT_ELAST_init_25D, mat_elastic.f90:363-383): beta = dvol * mu * (pi (1 - nu) / W)^2 per GLL point of
// every element, in the strip layout; SH: without the (1 - nu) factor
template <typename T>
__global__ void k_cart_beta(CartGeom G, T* __restrict__ beta, double W) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N;
  if (w >= (long long)G.nx * G.nz * N2) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2), i = k % N, j = k / N;
  const int ix = (int)(e % G.nx), iz = (int)(e / G.nx);
  double rho, cp, cs;
  cart_material(G, ix, iz, i, j, rho, cp, cs);
  const double la = rho * (cp * cp - 2.0 * cs * cs), mu = rho * cs * cs;
  const double dvol = ((0.5 * G.hx) * (0.5 * G.hz)) * (G.wgll[i] * G.wgll[j]);
  const double nu = la / (la + mu) / 2.0;
  const double pi = 4.0 * atan(1.0);
  const double t = (G.ndof == 1) ? pi / W : pi * (1 - nu) / W;
  beta[strip_scalar_index(G.S, ix, iz, i, j)] = (T)(dvol * mu * (t * t));
}

// assembled mass (mat_mass.f90:50-57): each node is summed by its first element, over the elements
// that share it in ascending element order; virtual neighbour columns stand in for the elements of
// an adjacent x-strip so that interface nodes carry the full mass on both ranks.
template <typename T>
__global__ void k_cart_mass(CartGeom G, T* __restrict__ mass, size_t npoin) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N;
  const long long total = (long long)G.nx * G.nz * N2;
  if (w >= total) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2);
  const int i = k % N, j = k / N;
  const int ix = (int)(e % G.nx), iz = (int)(e / G.nx);
  // canonical owner?
  if (i == 0 && ix > 0) return;
  if (j == 0 && !row_detached(G, iz)) return;
  const double det = (0.5 * G.hx) * (0.5 * G.hz);
  double sum = 0.0;
  const bool up_ok = (j == N - 1) && (iz + 1 < G.nz) && !row_detached(G, iz + 1);
  for (int dz = 0; dz <= (up_ok ? 1 : 0); ++dz)
    for (int dx = -1; dx <= 1; ++dx) {
      // elements containing the node in this row: dx=0 always; dx=+1 when on the right edge;
      // dx=-1 only as the virtual column of the left neighbour strip
      int eix = ix + dx, li = i;
      if (dx == 1) {
        if (i != N - 1) continue;
        if (eix >= G.nx && !G.halo_right) continue;
        li = 0;
      } else if (dx == -1) {
        if (!(i == 0 && ix == 0 && G.halo_left)) continue;
        li = N - 1;
      }
      const int eiz = iz + dz;
      const int lj = dz ? 0 : j;
      double rho, cp, cs;
      cart_material(G, eix, eiz, li, lj, rho, cp, cs);
      sum = sum + rho * (det * (G.wgll[li] * G.wgll[lj]));
    }
  const size_t node = (size_t)(cart_lat_id(G, ix, iz, i, j) - 1);
  for (int c = 0; c < G.ndof; ++c) mass[node + npoin * c] = (T)sum;
}

// max over the grid of c / min(dx,dz) (CHECK_grid, init.f90:187-225)
__global__ void k_cart_cfl(CartGeom G, unsigned long long* out) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N;
  const long long total = (long long)G.nx * G.nz * N2;
  double r = 0.0;
  if (w < total) {
    const long long e = w / N2;
    const int k = (int)(w - e * N2);
    const int i = k % N, j = k / N;
    if (i < N - 1 && j < N - 1) {
      double rho, cp, cs;
      cart_material(G, (int)(e % G.nx), (int)(e / G.nx), i, j, rho, cp, cs);
      const double dx = 0.5 * G.hx * (G.xgll[i + 1] - G.xgll[i]);
      const double dz = 0.5 * G.hz * (G.xgll[j + 1] - G.xgll[j]);
      r = (G.ndof == 2 ? cp : cs) / fmin(dx, dz);
    }
  }
  for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  if ((threadIdx.x & 31) == 0 && r > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(r));
}

template <typename T>
__global__ void k_add_mass(T* mass, size_t npoin, int ndof, int np, const int* node, const double* C, double coef) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= np) return;
  for (int c = 0; c < ndof; ++c) {
    const size_t q = (size_t)(node[k] - 1) + npoin * c;
    mass[q] = (T)((double)mass[q] + coef * C[k + (size_t)np * c]);  // bc_abso.f90:243
  }
}
template <typename T>
__global__ void k_gather_nodes(const T* src, size_t npoin, int ndof, int np, const int* node, double* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= np) return;
  for (int c = 0; c < ndof; ++c) out[k + (size_t)np * c] = (double)src[(size_t)(node[k] - 1) + npoin * c];
}

template <typename T>
__global__ void k_fill(T* x, size_t n, T val) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) x[q] = val;
}
template <typename TS, typename TD>
__global__ void k_cast_copy(const TS* __restrict__ src, TD* __restrict__ dst, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) dst[q] = (TD)src[q];
}

// Seeded, non-trivial initial state: d = amp_d * u, v = amp_v * u' with u, u' in U(-1,1) from the counter-based
// hash of the node's GLOBAL geometric lattice coordinates (both sides of the split fault row get the same
// values: no initial slip; x-strips of one global mesh agree on their shared columns).
template <typename T>
__global__ void k_cart_fill(CartGeom G, T* __restrict__ d, T* __restrict__ v, size_t nlat, uint64_t seed,
                            double amp_d, double amp_v) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)G.S.LX * G.S.LZ;
  if (w >= total) return;
  const int gz = (int)(w / G.S.LX), gx = (int)(w - (long long)gz * G.S.LX);
  const int gzg = gz - ((G.ezflt > 0 && gz >= G.ezflt * (G.N - 1) + 1) ? 1 : 0);
  const uint64_t X = (uint64_t)(G.ix0 + gx), Z = (uint64_t)(G.iz0 + gzg);
  const size_t q = (size_t)gz * G.S.LXP + gx;
  for (int c = 0; c < G.ndof; ++c) {
    d[q + nlat * c] = (T)(amp_d * hash_u(seed, X, Z, 16 + c));
    v[q + nlat * c] = (T)(amp_v * hash_u(seed, X, Z, 32 + c));
  }
}
// a window of the lattice, out[c][gz - gz0][gx - gx0] in FP64
template <typename T>
__global__ void k_cart_window(const T* __restrict__ src, double* __restrict__ out, size_t nlat, int LXP, int ndof,
                              int gx0, int gz0, int nwx, int nwz) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nwx * nwz * ndof;
  if (w >= total) return;
  const int c = (int)(w / ((long long)nwx * nwz));
  const long long r = w - (long long)c * nwx * nwz;
  const int z = (int)(r / nwx), x = (int)(r - (long long)z * nwx);
  out[w] = (double)src[(size_t)(gz0 + z) * LXP + (gx0 + x) + nlat * c];
}

// PLOT_FIELD's element-wise snapshot fields on the flat box (plot_gen.f90:239-300): strain of d (FIELD_strain_elem,
// fields.f90:192-237), stress of the Kelvin-Voigt-modified d (MAT_stress_dv, mat_gen.f90:626-641; isotropic
// MAT_ELAST_stress, mat_elastic.f90:822-839), divergence / curl of v (FIELD_divcurl_elem, fields.f90:242-285).
// One thread per GLL point of an element; out[(c*nelem + e)*N2 + k] float32, e in the caller's element order.
struct SnapH {
  double H[100];
};
template <typename T>
__global__ void k_cart_snap(CartGeom G, SnapH Hm, const T* __restrict__ d, const T* __restrict__ v,
                            const T* __restrict__ eta_strip, const T* __restrict__ ep_strip, size_t nlat, int what,
                            const int* __restrict__ perm, float* __restrict__ out) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N, ndof = G.ndof;
  const long long nelem = (long long)G.nx * G.nz;
  if (w >= nelem * N2) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2), i = k % N, j = k / N;
  const long long old = perm ? perm[e] : e;
  const int ix = (int)(old % G.nx), iz = (int)(old / G.nx);
  const bool of_v = what == 'd' || what == 'c';
  const T* F = of_v ? v : d;
  const bool kv = what == 'S' && eta_strip != nullptr;
  auto val = [&](int c, int ii, int jj) {
    const size_t q = (size_t)(cart_lat_id(G, ix, iz, ii, jj) - 1) + nlat * c;
    double u = (double)F[q];
    if (kv) u = u + (double)eta_strip[strip_scalar_index(G.S, ix, iz, ii, jj)] * (double)v[q];
    return u;
  };
  double dxi[2] = {0, 0}, deta[2] = {0, 0};
  for (int c = 0; c < ndof; ++c)
    for (int m = 0; m < N; ++m) {
      dxi[c] += Hm.H[m + N * i] * val(c, m, j);   // (Ht U)(i,j)
      deta[c] += val(c, i, m) * Hm.H[m + N * j];  // (U H)(i,j)
    }
  const double dxi_dx = 2.0 / G.hx, deta_dz = 2.0 / G.hz;
  double ev[3] = {0, 0, 0};
  int ncomp = 1;
  if (what == 'E' || what == 'S') {
    ncomp = ndof + 1;
    if (ndof == 1) {
      ev[0] = 0.5 * (dxi[0] * dxi_dx);
      ev[1] = 0.5 * (deta[0] * deta_dz);
    } else {
      ev[0] = dxi[0] * dxi_dx;
      ev[1] = deta[1] * deta_dz;
      ev[2] = 0.5 * (deta[0] * deta_dz + dxi[1] * dxi_dx);
    }
    if (what == 'S' && ep_strip != nullptr) {  // MAT_PLAST_stress without update: the relative elastic strain e - ep
      for (int c = 0; c < 3; ++c) ev[c] = ev[c] - (double)ep_strip[strip_ep_index(G.S, ix, iz, i, j, c)];
    }
    if (what == 'S') {
      double rho, cp, cs;
      cart_material(G, ix, iz, i, j, rho, cp, cs);
      const double la = rho * (cp * cp - 2.0 * cs * cs), mu = rho * cs * cs;
      if (ndof == 1) {
        ev[0] = 2.0 * mu * ev[0];
        ev[1] = 2.0 * mu * ev[1];
      } else {
        const double e1 = ev[0], e2 = ev[1];
        ev[0] = (la + 2.0 * mu) * e1 + la * e2;
        ev[1] = la * e1 + (la + 2.0 * mu) * e2;
        ev[2] = 2.0 * mu * ev[2];
      }
    }
  } else if (what == 'd') {
    ev[0] = dxi[0] * dxi_dx + deta[1] * deta_dz;
  } else {
    ev[0] = deta[0] * deta_dz - dxi[1] * dxi_dx;
  }
  for (int c = 0; c < ncomp; ++c) out[((size_t)c * nelem + e) * N2 + k] = (float)ev[c];
}

// ibool in the reference layout (ngll,ngll,nelem), natural element order
__global__ void k_cart_ibool(CartGeom G, int* __restrict__ ibool) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N2 = G.N * G.N;
  const long long total = (long long)G.nx * G.nz * N2;
  if (w >= total) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2);
  ibool[w] = (int)cart_node_id(G, (int)(e % G.nx), (int)(e / G.nx), k % G.N + 1, k / G.N + 1);
}

// field transfer between the caller's numbering (SE_init_numbering ids) and the lattice:
// to_ref != 0: ref[id] = lat[lattice];  else lat[lattice] = ref[id].  Nodes shared by several
// elements are written several times with the same value.
template <typename TS, typename TD>
__global__ void k_cart_permute(CartGeom G, const TS* __restrict__ src, TD* __restrict__ dst, size_t np_ref,
                               size_t np_lat, int ncomp, int to_ref) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N;
  const long long total = (long long)G.nx * G.nz * N2;
  if (w >= total) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2);
  const int i = k % N, j = k / N;
  const int ix = (int)(e % G.nx), iz = (int)(e / G.nx);
  const size_t r = (size_t)(cart_node_id(G, ix, iz, i + 1, j + 1) - 1);
  const size_t l = (size_t)(cart_lat_id(G, ix, iz, i, j) - 1);
  for (int c = 0; c < ncomp; ++c) {
    if (to_ref) dst[r + np_ref * c] = (TD)src[l + np_lat * c];
    else dst[l + np_lat * c] = (TD)src[r + np_ref * c];
  }
}

// ---------------------------------------------------------------------------------------------
// GLL points, weights and derivative matrix (the quantities of SRC/gll.f90 get_GLL_info,
// computed here by Newton iteration on (1-x^2) P'_{n-1}(x))
static void legendre(int p, double x, double& P, double& dP) {
  double p0 = 1.0, p1 = x;
  if (p == 0) { P = 1.0; dP = 0.0; return; }
  for (int k = 2; k <= p; ++k) {
    const double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  P = p1;
  dP = (x * x == 1.0) ? 0.5 * p * (p + 1) * ((p % 2 == 0 || x > 0) ? 1.0 : -1.0) : p * (p0 - x * p1) / (1.0 - x * x);
}
static void gll_tables(int n, double* x, double* w, double* H /* H[i + n*j] = h'_i(x_j) */) {
  const int p = n - 1;
  const double PI = 3.141592653589793;
  x[0] = -1.0;
  x[p] = 1.0;
  for (int k = 1; k < p; ++k) {
    double z = -std::cos(PI * k / p);
    for (int itn = 0; itn < 100; ++itn) {
      double P, dP;
      legendre(p, z, P, dP);
      // q(z) = P'_p(z); q'(z) from the Legendre ODE: (1-z^2) P'' = 2 z P' - p(p+1) P
      const double ddP = (2.0 * z * dP - p * (p + 1.0) * P) / (1.0 - z * z);
      const double dz = dP / ddP;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[k] = z;
  }
  for (int k = 0; k <= p / 2; ++k) {  // enforce symmetry, middle point exactly 0 (gll.f90:27-28)
    const double s = 0.5 * (x[p - k] - x[k]);
    x[k] = -s;
    x[p - k] = s;
  }
  if (n % 2 == 1) x[p / 2] = 0.0;
  std::vector<double> Pv(n);
  for (int k = 0; k < n; ++k) {
    double P, dP;
    legendre(p, x[k], P, dP);
    Pv[k] = P;
    w[k] = 2.0 / (p * (p + 1.0) * P * P);
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double v;
      if (i != j) v = Pv[j] / (Pv[i] * (x[j] - x[i]));
      else if (i == 0) v = -0.25 * p * (p + 1.0);
      else if (i == p) v = 0.25 * p * (p + 1.0);
      else v = 0.0;
      H[i + n * j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
struct CartState {
  CartGeom G;
  s2d_scheme scheme;
  double courant, dt, grid_cfl;
  double H[100];
  bool mass_inverted = false;
  bool bc_added = false;
  int perio_lr = 0, perio_bt = 0;   // periodic pairs: left-right (tags 4,2), bottom-top (tags 1,3)
  struct FaultInfo {                // per fault id: bc%coord (2,np), bc%T0 (np,2), bc%B(:,1) (FltXX_sem2d.hdr, _init.tab)
    std::vector<double> coord, T0, B;
  };
  std::vector<FaultInfo> faults;
  std::vector<double> rec_coord;  // (2, nx) positions of the relocated stations (rec%coord, receivers.f90:231-303)
  int coef_mode = 0;  // 0 = compact (lambda, mu) where the rheology allows, 1 = all planes stored
  std::vector<double> h_mat[3];   // s2d_cart_set_material: host and device copies of rho, cp, cs
  DevBuf<double> d_mat[3];
  bool dt_given = false;
  double W25d = 0.0;              // &GENERAL W when finite (2.5D), else 0
  // OPT_RENUMBER (s2d_cart_desc.renumber): elements in reverse Cuthill-McKee order and the GLL numbering that
  // SE_init_numbering gives in that order (spec_grid.f90:198-314)
  bool renumber = false;
  std::vector<int32_t> perm;        // perm[new] = old element (0-based, natural order)
  std::vector<int32_t> id_of_lat;   // (LXP*LZ) 1-based node id of every lattice point (0 on the pad columns)
  CartGeom dev_geom() const {     // the geometry as the kernels see it
    CartGeom g = G;
    for (int q = 0; q < 3; ++q) g.mat[q] = d_mat[q].p;
    return g;
  }
  bool nm() const { return scheme.kind == 1 || scheme.kind == 2; }  // 'newmark', 'HHT-alpha'
  double CoefA2V() const { return nm() ? scheme.gamma * scheme.dt : scheme.dt; }        // time.f90:443-456
  double CoefA2D() const { return nm() ? scheme.beta * scheme.dt * scheme.dt : 0.0; }    // time.f90:426-440
  double CoefA2Vrhs() const { return nm() ? scheme.alpha * CoefA2V() : 0.5 * CoefA2V(); }  // :465-486
};

void cart_free(void* c) { delete (CartState*)c; }

int select_device(int device, std::string& err);
s2d_handle wrap_engine(std::unique_ptr<EngineBase> impl, void* cart);
void* engine_cart(s2d_handle h);
EngineBase* engine_impl(s2d_handle h);
void set_error(s2d_handle h, const std::string& m);

// max over the grid of c / min(dx,dz) (CHECK_grid, init.f90:187-225) with the current material
static double cart_grid_cfl(const CartState& S) {
  const CartGeom G = S.dev_geom();
  DevBuf<unsigned long long> mx;
  mx.alloc(1);
  mx.zero();
  const long long tot = (long long)G.nx * G.nz * G.N * G.N;
  k_cart_cfl<<<(unsigned)((tot + 255) / 256), 256>>>(G, mx.p);
  S2D_CUDA(cudaGetLastError());
  unsigned long long bits = 0;
  S2D_CUDA(cudaMemcpy(&bits, mx.p, 8, cudaMemcpyDeviceToHost));
  double v;
  std::memcpy(&v, &bits, 8);
  return v;
}

// operator data of the box from the current material: coefficient planes in the strip layout
// (MAT_ELAST_init_a, mat_elastic.f90:290-360), assembled mass (MAT_MASS_init, mat_mass.f90:29-61)
template <typename T>
static void cart_operator(Engine<T>& E, CartState& S) {
  const CartGeom G = S.dev_geom();
  const int N = G.N, N2 = N * N;
  const int nelast = (G.ndof == 1) ? 2 : 6;
  cudaStream_t st = E.stream;
  const long long tot = (long long)E.nelem * N2;
  const unsigned nblk = (unsigned)((tot + 255) / 256);
  E.ncoefsets = (G.seed != 0 || G.mat[0]) ? E.nelem : 1;
  E.p_coef.alloc((size_t)E.nelem * (E.cart_compact ? 2 : nelast) * N2);
  k_cart_coef<T><<<nblk, 256, 0, st>>>(G, E.p_coef.p, nelast, E.cart_compact);
  // mass (kept un-inverted in rmass until commit); the pad columns of the lattice rows get 1 so that 1/M and
  // every node-wise pass stay finite there (their fields are zero and nothing ever reads them)
  k_fill<T><<<(unsigned)std::min<size_t>((E.rmass.n + 255) / 256, 148 * 32), 256, 0, st>>>(E.rmass.p, E.rmass.n, (T)1);
  k_cart_mass<T><<<nblk, 256, 0, st>>>(G, E.rmass.p, E.npoin);
  S2D_CUDA(cudaGetLastError());
  // the assembled mass as MAT_MASS_init leaves it (mat_mass.f90:50-57), before BC_init augments it: what
  // energy_compute's sum(w*rho*v2) (energy.f90:49-106) amounts to node by node
  E.mass.alloc(E.npoin);
  k_cast_copy<T, double><<<(unsigned)((E.npoin + 255) / 256), 256, 0, st>>>(E.rmass.p, E.mass.p, E.npoin);
  S2D_CUDA(cudaGetLastError());
  if (S.W25d > 0.0) {
    E.strip_beta.alloc((size_t)E.nelem * N2);
    k_cart_beta<T><<<nblk, 256, 0, st>>>(G, E.strip_beta.p, S.W25d);
    S2D_CUDA(cudaGetLastError());
  }
  S2D_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static void cart_build(Engine<T>& E, CartState& S) {
  CartGeom& G = S.G;
  const int N = G.N, N2 = N * N;
  const int nelast = (G.ndof == 1) ? 2 : 6;
  cudaStream_t st = E.stream;
  const long long tot = (long long)E.nelem * N2;
  const unsigned nblk = (unsigned)((tot + 255) / 256);
  // coefficient planes, one block per element (mat_gen.f90:357-365 would share one block between
  // homogeneous elements; the strip kernel streams its planes, so they are materialised per element)
  E.nelast = nelast;
  E.kd2 = (N == 5) ? 1 : 0;  // OPT_NGLL (constants.f90:6, mat_elastic.f90:412)
  E.nkv = 0;
  E.p_hetero = true;
  E.cart_compact = (G.ndof == 2 && S.coef_mode == 0) ? 1 : 0;
  E.cart_cdx = 2.0 / G.hx;
  E.cart_cdz = 2.0 / G.hz;
  E.cart_cdet = (0.5 * G.hx) * (0.5 * G.hz);
  E.cart_wgll.assign(G.wgll, G.wgll + N);
  cart_operator<T>(E, S);
  E.init_strip_tables(G.S);  // halo arrays and decomposition flags of the strip kernel
  CartGeom Gc = G;
  Gc.mat[0] = Gc.mat[1] = Gc.mat[2] = nullptr;  // the permutations only use the numbering
  Engine<T>* Ep = &E;
  E.cart_to_ref = [Ep, Gc, nblk](const T* lat, double* ref) {
    k_cart_permute<T, double><<<nblk, 256, 0, Ep->stream>>>(Gc, lat, ref, Ep->npoin_ref, Ep->npoin, Gc.ndof, 1);
    S2D_CUDA(cudaGetLastError());
  };
  E.cart_from_ref = [Ep, Gc, nblk](const double* ref, T* lat) {
    k_cart_permute<double, T><<<nblk, 256, 0, Ep->stream>>>(Gc, ref, lat, Ep->npoin_ref, Ep->npoin, Gc.ndof, 0);
    S2D_CUDA(cudaGetLastError());
  };
  E.cart_from_ref1 = [Ep, Gc, nblk](const double* ref, T* lat) {
    k_cart_permute<double, T><<<nblk, 256, 0, Ep->stream>>>(Gc, ref, lat, Ep->npoin_ref, Ep->npoin, 1, 0);
    S2D_CUDA(cudaGetLastError());
  };
  E.cart_from_ref1d = [Ep, Gc, nblk](const double* ref, double* lat) {
    k_cart_permute<double, double><<<nblk, 256, 0, Ep->stream>>>(Gc, ref, lat, Ep->npoin_ref, Ep->npoin, 1, 0);
    S2D_CUDA(cudaGetLastError());
  };
  if (S.renumber) {  // RCM element order: the caller's node ids come from a table, like a routed generic handle
    std::vector<int32_t> lat_of(E.npoin_ref);
    for (size_t l = 0; l < S.id_of_lat.size(); ++l)
      if (S.id_of_lat[l] > 0) lat_of[(size_t)S.id_of_lat[l] - 1] = (int32_t)l;
    E.install_lat_table(lat_of);
  }
  S2D_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static Engine<T>* as_engine(EngineBase* b) {
  return static_cast<Engine<T>*>(b);
}

// coordinates of GLL point (i,j) (0-based) of element (ix,iz)
static inline double gx_of(const CartGeom& G, int ix, int i) { return G.x0 + G.hx * (ix + 0.5 * (G.xgll[i] + 1.0)); }
static inline double gz_of(const CartGeom& G, int iz, int j) { return G.z0 + G.hz * (iz + 0.5 * (G.xgll[j] + 1.0)); }

// node id (1-based, the caller's numbering) of GLL point (i,j) (0-based) of element (ix,iz)
static inline long long cart_ref_id(const CartState& S, int ix, int iz, int i, int j) {
  if (S.renumber) return S.id_of_lat[(size_t)(cart_lat_id(S.G, ix, iz, i, j) - 1)];
  return cart_node_id(S.G, ix, iz, i + 1, j + 1);
}

// SE_init_numbering (spec_grid.f90:249-287) over the elements in the order perm: interior points (i fastest),
// then the edges D, R, U, L whose points are still unnumbered (counter-clockwise interior points), then the
// vertices SW, SE, NE, NW that are still unnumbered.  A GLL point is identified by its lattice position, which is
// what the reference's neighbour / vertex lists establish (split fault rows are two lattice rows).
static void cart_number_nodes(CartState& S) {
  const CartGeom& G = S.G;
  const int N = G.N;
  S.id_of_lat.assign((size_t)G.S.LXP * G.S.LZ, 0);
  int32_t npoin = 0;
  auto at = [&](int ix, int iz, int i, int j) -> int32_t& { return S.id_of_lat[(size_t)(cart_lat_id(G, ix, iz, i, j) - 1)]; };
  for (size_t en = 0; en < S.perm.size(); ++en) {
    const int ix = S.perm[en] % G.nx, iz = S.perm[en] / G.nx;
    for (int j = 1; j < N - 1; ++j)
      for (int i = 1; i < N - 1; ++i) at(ix, iz, i, j) = ++npoin;
    for (int n = 0; n < 4; ++n) {  // edge tables (spec_grid.f90:876-883): D, R, U, L
      auto pt = [&](int k, int& i, int& j) {
        switch (n) {
          case 0: i = k; j = 0; break;
          case 1: i = N - 1; j = k; break;
          case 2: i = N - 1 - k; j = N - 1; break;
          default: i = 0; j = N - 1 - k; break;
        }
      };
      int i, j;
      pt(1, i, j);
      if (at(ix, iz, i, j) != 0) continue;
      for (int k = 1; k < N - 1; ++k) {
        pt(k, i, j);
        at(ix, iz, i, j) = ++npoin;
      }
    }
    const int vi[4] = {0, N - 1, N - 1, 0}, vj[4] = {0, 0, N - 1, N - 1};
    for (int n = 0; n < 4; ++n)
      if (at(ix, iz, vi[n], vj[n]) == 0) at(ix, iz, vi[n], vj[n]) = ++npoin;
  }
}

// nearest GLL lattice position along one axis: element index and local index
static void nearest_1d(const CartGeom& G, double x, double x0, double h, int n, int& e, int& i) {
  double best = 1e300;
  e = 0;
  i = 0;
  int ec = (int)std::floor((x - x0) / h);
  for (int ee = std::max(0, ec - 1); ee <= std::min(n - 1, ec + 1); ++ee)
    for (int k = 0; k < G.N; ++k) {
      const double xx = x0 + h * (ee + 0.5 * (G.xgll[k] + 1.0));
      const double dd = std::fabs(xx - x);
      if (dd <= best) {  // ties -> later (higher id), as SE_find_nearest_node does (spec_grid.f90:423-429)
        best = dd;
        e = ee;
        i = k;
      }
    }
}

}  // namespace s2d

using namespace s2d;

#define CART_GUARD_BEGIN                                              \
  if (!h || !engine_impl(h) || !engine_cart(h)) return S2D_EINVAL;    \
  EngineBase* Eb = engine_impl(h);                                    \
  CartState& S = *(CartState*)engine_cart(h);                         \
  (void)S;                                                            \
  try {                                                               \
    S2D_CUDA(cudaSetDevice(Eb->device));
#define CART_GUARD_END                                                \
    return S2D_OK;                                                    \
  } catch (const ArgError& e) {                                       \
    set_error(h, e.what());                                           \
    return S2D_EINVAL;                                                \
  } catch (const StateError& e) {                                     \
    set_error(h, e.what());                                           \
    return S2D_ESTATE;                                                \
  } catch (const std::exception& e) {                                 \
    set_error(h, e.what());                                           \
    return S2D_ECUDA;                                                 \
  }

static thread_local std::string g_cart_err;

extern "C" {

int s2d_cart_create(s2d_handle* out, const s2d_cart_desc* D) {
  if (!out || !D) return S2D_EINVAL;
  *out = nullptr;
  if (D->ngll < 3 || D->ngll > 10 || (D->ndof != 1 && D->ndof != 2) || D->nx < 1 || D->nz < 1 || D->ezflt < 0 ||
      D->ezflt >= D->nz || !(D->x1 > D->x0) || !(D->z1 > D->z0) || (D->precision != 8 && D->precision != 4) ||
      D->scheme.kind < 0 || D->scheme.kind > 3)
    return S2D_EINVAL;
  std::string err;
  const int dev = select_device(D->device, err);
  if (dev < 0) return S2D_ENODEV;
  try {
    std::unique_ptr<CartState> S(new CartState());
    CartGeom& G = S->G;
    G.N = D->ngll;
    G.ndof = D->ndof;
    G.nx = D->nx;
    G.nz = D->nz;
    G.ezflt = D->ezflt;
    G.x0 = D->x0;
    G.z0 = D->z0;
    G.hx = (D->x1 - D->x0) / (double)D->nx;
    G.hz = (D->z1 - D->z0) / (double)D->nz;
    G.seed = D->seed;
    G.ix0 = D->ix0;
    G.iz0 = D->iz0;
    G.rho = D->rho;
    G.cp = D->cp;
    G.cs = D->cs;
    G.halo_left = D->halo_left;
    G.halo_right = D->halo_right;
    gll_tables(G.N, G.xgll, G.wgll, S->H);
    // strip decomposition of the z-marching kernel (strip_kernels.cuh)
    G.S = make_strip_geom(G.N, G.ndof, G.nx, G.nz, G.ezflt, strip_default_seg(G.N, G.nx, G.nz), G.halo_left != 0, G.halo_right != 0);
    const long long npoin = cart_npoin(G);
    const long long nelem = (long long)G.nx * G.nz;
    if (npoin != (long long)G.S.LX * G.S.LZ) {
      g_cart_err = "internal error: lattice size does not match the node count";
      return S2D_EINVAL;
    }
    const long long nlat = (long long)G.S.LXP * G.S.LZ;  // device node count: lattice rows of pitch LXP
    if (nlat > 2147483647LL || nelem * G.N * G.N > (1LL << 40)) {
      g_cart_err = "mesh too large for 32-bit node ids";
      return S2D_EINVAL;
    }
    S->renumber = D->renumber != 0;
    if (S->renumber) {
      if (G.halo_left || G.halo_right) {
        g_cart_err = "renumber: an x-strip with neighbours keeps the natural element order";
        return S2D_EINVAL;
      }
      S->perm = rcm_box_perm(G.nx, G.nz);
      cart_number_nodes(*S);
    }
    S->scheme = D->scheme;
    S->courant = D->courant;
    S->coef_mode = D->coef_mode != 0 ? 1 : (env_int("S2D_COEF_FULL", 0) != 0 ? 1 : 0);
    // dt from the Courant number (init.f90:187-225, time.f90:334-341)
    S->grid_cfl = cart_grid_cfl(*S);
    S->dt_given = S->scheme.dt > 0.0;
    if (!(S->scheme.dt > 0.0)) S->scheme.dt = D->courant / S->grid_cfl;
    S->dt = S->scheme.dt;
    std::unique_ptr<EngineBase> impl;
    if (D->precision == 8) {
      auto* E = new Engine<double>(Engine<double>::Raw(), G.N, G.ndof, (int)nelem, (size_t)nlat, (size_t)npoin, S->H, S->scheme, dev);
      impl.reset(E);
      cart_build<double>(*E, *S);
    } else {
      auto* E = new Engine<float>(Engine<float>::Raw(), G.N, G.ndof, (int)nelem, (size_t)nlat, (size_t)npoin, S->H, S->scheme, dev);
      impl.reset(E);
      cart_build<float>(*E, *S);
    }
    *out = wrap_engine(std::move(impl), S.release());
    return S2D_OK;
  } catch (const std::exception& e) {
    g_cart_err = e.what();
    fprintf(stderr, "s2d_cart_create: %s\n", e.what());
    return S2D_ECUDA;
  }
}

// MAT_init_prop's result for the box (mat_gen.f90:204-303): the caller evaluated its materials (per-tag constants,
// DIST_* fields) at the GLL points of every element; the operator data are rebuilt from them
int s2d_cart_set_material(s2d_handle h, const double* rho, const double* cp, const double* cs) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(rho && cp && cs, "cart_set_material: null pointer");
  S2D_REQUIRE(!Eb->committed && !S.bc_added, "cart_set_material: must precede the boundary conditions (BC_init reads the mass)");
  S2D_REQUIRE(!S.G.halo_left && !S.G.halo_right, "cart_set_material: not available on an x-strip with neighbours");
  const size_t n = (size_t)Eb->nelem * S.G.N * S.G.N;
  const double* src[3] = {rho, cp, cs};
  for (size_t q = 0; q < n; ++q)
    S2D_REQUIRE(rho[q] > 0.0 && cp[q] > 0.0 && cs[q] > 0.0, "cart_set_material: rho, cp, cs must be positive");
  for (int k = 0; k < 3; ++k) {
    S.h_mat[k].assign(src[k], src[k] + n);
    S.d_mat[k].upload(S.h_mat[k]);
    S.G.mat[k] = S.h_mat[k].data();
  }
  S.grid_cfl = cart_grid_cfl(S);
  if (!S.dt_given) {  // time.f90:334-341
    S.scheme.dt = S.courant / S.grid_cfl;
    S.dt = S.scheme.dt;
    Eb->scheme.dt = S.dt;
  }
  if (Eb->prec == 8) cart_operator<double>(*as_engine<double>(Eb), S);
  else cart_operator<float>(*as_engine<float>(Eb), S);
  CART_GUARD_END
}

// matwrk_kv_type%eta of the Kelvin-Voigt elements (mat_kelvin_voigt.f90:117-150), elements in natural order
int s2d_cart_set_kv_elems(s2d_handle h, int32_t nkv, const int32_t* elem_ids, const double* eta) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(nkv >= 1 && elem_ids && eta, "cart_set_kv_elems: bad arguments");
  S2D_REQUIRE(!Eb->committed, "cart_set_kv_elems after commit");
  const CartGeom& G = S.G;
  const int N = G.N, n2 = N * N;
  std::vector<double> pe((size_t)Eb->nelem * n2, 0.0);
  for (int k = 0; k < nkv; ++k) {
    const int e = elem_ids[k] - 1;
    S2D_REQUIRE(e >= 0 && e < Eb->nelem, "cart_set_kv_elems: element id out of range");
    const int ix = e % G.nx, iz = e / G.nx;
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) pe[strip_scalar_index(G.S, ix, iz, i, j)] = eta[(size_t)k * n2 + i + N * j];
  }
  Eb->set_strip_eta(pe.data(), pe.size());
  CART_GUARD_END
}

int s2d_cart_set_plastic(s2d_handle h, int32_t nsets, const double* par, const int32_t* elem_set) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(nsets >= 1 && par && elem_set, "cart_set_plastic: bad arguments");
  S2D_REQUIRE(!Eb->committed, "cart_set_plastic after commit");
  const CartGeom& G = S.G;
  std::vector<unsigned char> ps((size_t)Eb->nelem, 0);
  for (int e = 0; e < Eb->nelem; ++e) {
    S2D_REQUIRE(elem_set[e] >= 0 && elem_set[e] <= nsets, "cart_set_plastic: material set out of range");
    ps[strip_elem_slot(G.S, e % G.nx, e / G.nx)] = (unsigned char)elem_set[e];
  }
  Eb->set_strip_plastic(ps.data(), ps.size(), nsets, par);
  CART_GUARD_END
}

int s2d_cart_set_visco(s2d_handle h, int32_t nsets, const int32_t* nbody, const double* moduli, const double* wbody,
                       const double* theta, const int32_t* elem_set) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(nsets >= 1 && nbody && moduli && wbody && theta && elem_set, "cart_set_visco: bad arguments");
  S2D_REQUIRE(!Eb->committed, "cart_set_visco after commit");
  const CartGeom& G = S.G;
  std::vector<unsigned char> ps((size_t)Eb->nelem, 0);
  for (int e = 0; e < Eb->nelem; ++e) {
    S2D_REQUIRE(elem_set[e] >= 0 && elem_set[e] <= nsets, "cart_set_visco: material set out of range");
    ps[strip_elem_slot(G.S, e % G.nx, e / G.nx)] = (unsigned char)elem_set[e];
  }
  Eb->set_strip_visco(ps.data(), ps.size(), nsets, nbody, moduli, wbody, theta);
  CART_GUARD_END
}

int s2d_cart_set_damage(s2d_handle h, int32_t nsets, const double* par, const int32_t* elem_set) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(nsets >= 1 && par && elem_set, "cart_set_damage: bad arguments");
  S2D_REQUIRE(!Eb->committed, "cart_set_damage after commit");
  const CartGeom& G = S.G;
  const int N = G.N, n2 = N * N;
  std::vector<unsigned char> ps((size_t)Eb->nelem, 0);
  bool nonzero = false;
  for (int k = 0; k < nsets; ++k)
    for (int q : {3, 10, 11, 12}) nonzero = nonzero || par[(size_t)13 * k + q] != 0.0;
  std::vector<double> st;
  if (nonzero) st.assign((size_t)Eb->nelem * 4 * n2, 0.0);
  for (int e = 0; e < Eb->nelem; ++e) {
    S2D_REQUIRE(elem_set[e] >= 0 && elem_set[e] <= nsets, "cart_set_damage: material set out of range");
    const int ix = e % G.nx, iz = e / G.nx;
    ps[strip_elem_slot(G.S, ix, iz)] = (unsigned char)elem_set[e];
    if (nonzero && elem_set[e] > 0) {
      const double* p = par + (size_t)13 * (elem_set[e] - 1);
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
          st[strip_plane_index(G.S, ix, iz, i, j, 0, 4)] = p[3];
          for (int c = 0; c < 3; ++c) st[strip_plane_index(G.S, ix, iz, i, j, 1 + c, 4)] = p[10 + c];
        }
    }
  }
  Eb->set_strip_damage(ps.data(), ps.size(), nsets, par, nonzero ? st.data() : nullptr);
  CART_GUARD_END
}

int s2d_cart_get_damage_state(s2d_handle h, double* state) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(state, "cart_get_damage_state: null output");
  const CartGeom& G = S.G;
  const int N = G.N, n2 = N * N;
  std::vector<double> st((size_t)Eb->nelem * 4 * n2);
  Eb->get_strip_damage_state(st.data());
  for (int e = 0; e < Eb->nelem; ++e) {
    const int ix = e % G.nx, iz = e / G.nx;
    for (int k = 0; k < 4; ++k)
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) state[(size_t)e * 4 * n2 + (size_t)k * n2 + i + N * j] = st[strip_plane_index(G.S, ix, iz, i, j, k, 4)];
  }
  CART_GUARD_END
}

int s2d_cart_get_plastic_strain(s2d_handle h, double* ep) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(ep, "cart_get_plastic_strain: null output");
  const CartGeom& G = S.G;
  const int N = G.N, n2 = N * N;
  std::vector<double> st((size_t)Eb->nelem * 3 * n2);
  Eb->get_strip_plastic_strain(st.data());
  for (int e = 0; e < Eb->nelem; ++e) {
    const int ix = e % G.nx, iz = e / G.nx;
    for (int k = 0; k < 3; ++k)
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) ep[(size_t)e * 3 * n2 + (size_t)k * n2 + i + N * j] = st[strip_ep_index(G.S, ix, iz, i, j, k)];
  }
  CART_GUARD_END
}

int s2d_cart_set_w25d(s2d_handle h, double W) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(W > 0.0, "GENERAL input block: W must be positive");
  S2D_REQUIRE(!Eb->committed, "cart_set_w25d after commit");
  S.W25d = W;
  if (Eb->prec == 8) cart_operator<double>(*as_engine<double>(Eb), S);
  else cart_operator<float>(*as_engine<float>(Eb), S);
  CART_GUARD_END
}

int s2d_cart_info(s2d_handle h, int64_t* npoin, int64_t* nelem, double* dt) {
  CART_GUARD_BEGIN
  if (npoin) *npoin = (int64_t)Eb->npoin_ref;
  if (nelem) *nelem = (int64_t)Eb->nelem;
  if (dt) *dt = S.dt;
  CART_GUARD_END
}

int s2d_cart_set_dt(s2d_handle h, double dt) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(dt > 0.0, "cart_set_dt: dt must be positive");
  S2D_REQUIRE(!Eb->committed && !S.bc_added, "cart_set_dt: must precede the first boundary condition");
  S.scheme.dt = dt;
  S.dt = dt;
  Eb->scheme.dt = dt;
  CART_GUARD_END
}

// BC_ABSO_init for one flat side of the box (bc_abso.f90:115-266)
int s2d_cart_add_abso(s2d_handle h, int32_t side, int32_t stacey) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(side >= 1 && side <= 4, "cart_add_abso: side tag must be 1..4");
  S2D_REQUIRE(!Eb->committed, "cart_add_abso after commit");
  S.bc_added = true;
  const CartGeom& G = S.G;
  S2D_REQUIRE(!(side == 4 && G.halo_left) && !(side == 2 && G.halo_right),
              "cart_add_abso: that side is a strip interface, not a physical boundary");
  const int N = G.N, ndof = G.ndof;
  const bool horiz = (side == 1 || side == 3);
  const int ne = horiz ? G.nx : G.nz;
  // a vertical side crosses the split-node fault: its two coincident nodes are both on the boundary, the
  // lower one first (stable sort of BC_set_bulk_node, spec_grid.f90:801-818)
  const int split = (!horiz && G.ezflt > 0) ? 1 : 0;
  const int np = ne * (N - 1) + 1 + split;
  std::vector<int> node(np), bibool((size_t)N * ne);
  std::vector<double> C((size_t)np * ndof, 0.0), K;
  const bool st = stacey && ndof == 2;
  if (st) K.assign((size_t)N * 2 * ne, 0.0);
  // GeoDimTan/Nor (bc_abso.f90:146-160): horizontal boundary -> tangent x (1), normal z (2)
  const int GeoDimTan = horiz ? 1 : 2, GeoDimNor = horiz ? 2 : 1;
  const double jac1d = horiz ? 0.5 * G.hx : 0.5 * G.hz;
  const double dloc_dglob = 1.0 / jac1d;  // DLocDGlob(LocDimTan,GeoDimTan)
  // virtual boundary elements of neighbour strips complete C at the two interface corner nodes
  const int e_lo = (horiz && G.halo_left) ? -1 : 0, e_hi = (horiz && G.halo_right) ? ne : ne - 1;
  for (int e = e_lo; e <= e_hi; ++e) {
    for (int k = 0; k < N; ++k) {  // k-th point of the edge in counter-clockwise order
      int ix, iz, i, j, pos;       // pos = position along the sorted (ascending coordinate) node list
      switch (side) {
        case 1: ix = e; iz = 0; i = k; j = 0; pos = e * (N - 1) + i; break;                 // edge_D
        case 2: ix = G.nx - 1; iz = e; i = N - 1; j = k; pos = e * (N - 1) + j + ((split && e >= G.ezflt) ? 1 : 0); break;  // edge_R
        case 3: ix = e; iz = G.nz - 1; i = N - 1 - k; j = N - 1; pos = e * (N - 1) + i; break;  // edge_U
        default: ix = 0; iz = e; i = 0; j = N - 1 - k; pos = e * (N - 1) + j + ((split && e >= G.ezflt) ? 1 : 0); break;  // edge_L
      }
      if (pos < 0 || pos >= np) continue;  // virtual element: only its node on the interface counts
      double rho, cp, cs;
      cart_material(G, ix, iz, i, j, rho, cp, cs);
      const double CoefIntegr = G.wgll[k] * jac1d;
      double c[3];
      c[GeoDimNor] = cp;
      c[GeoDimTan] = cs;
      if (ndof == 1) {
        C[pos] += rho * cs * CoefIntegr;
      } else {
        C[pos] += rho * c[1] * CoefIntegr;
        C[pos + np] += rho * c[2] * CoefIntegr;
      }
      if (e >= 0 && e < ne) {
        node[pos] = (int)cart_lat_id(G, ix, iz, i, j);
        bibool[k + (size_t)N * e] = pos + 1;
        if (st) {
          const double kv = CoefIntegr * dloc_dglob * rho * c[GeoDimTan] * (2.0 * c[GeoDimTan] - c[GeoDimNor]);
          for (int cc = 0; cc < 2; ++cc)
            K[k + (size_t)N * (cc + 2 * (size_t)e)] = (cc == GeoDimTan - 1) ? -kv : kv;
        }
      }
    }
  }
  // the boundary meets a periodic one at both ends (BC_PERIO_intersects): bc_abso.f90:226-230
  if ((horiz && S.perio_lr) || (!horiz && S.perio_bt)) {
    S2D_REQUIRE(!st, "cart_add_abso: a Stacey boundary that meets a periodic one is not provided (bc_abso.f90:328-331)");
    for (int c = 0; c < ndof; ++c) {
      C[0 + (size_t)np * c] = C[0 + (size_t)np * c] + C[(np - 1) + (size_t)np * c];
      C[(np - 1) + (size_t)np * c] = C[0 + (size_t)np * c];
    }
  }
  Eb->add_abso(np, node.data(), C.data(), 1, nullptr, st ? 1 : 0, ne, bibool.data(), st ? K.data() : nullptr);
  // bc_abso.f90:243: the implicit treatment of C*v augments the mass
  DevBuf<int> dn;
  DevBuf<double> dC;
  dn.upload(node);
  dC.upload(C);
  if (Eb->prec == 8)
    k_add_mass<double><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<double>(Eb)->rmass.p, Eb->npoin, ndof, np, dn.p, dC.p, S.CoefA2Vrhs());
  else
    k_add_mass<float><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<float>(Eb)->rmass.p, Eb->npoin, ndof, np, dn.p, dC.p, S.CoefA2Vrhs());
  S2D_CUDA(cudaStreamSynchronize(Eb->stream));
  CART_GUARD_END
}

// BC_PERIO_init (bc_periodic.f90:44-74) between two opposite sides of the box
int s2d_cart_add_periodic(s2d_handle h, int32_t master_tag, int32_t slave_tag) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(!Eb->committed, "cart_add_periodic after commit");
  S2D_REQUIRE(!S.bc_added, "cart_add_periodic: periodic boundaries are initialised before the others (bc_gen.f90:221-228)");
  const CartGeom& G = S.G;
  const bool lr = (master_tag == 4 && slave_tag == 2) || (master_tag == 2 && slave_tag == 4);
  const bool bt = (master_tag == 1 && slave_tag == 3) || (master_tag == 3 && slave_tag == 1);
  S2D_REQUIRE(lr || bt, "cart_add_periodic: tags must be two opposite sides of the box");
  S2D_REQUIRE(!(lr && (G.halo_left || G.halo_right)), "cart_add_periodic: not available across x-strips");
  const int LX = G.S.LX, LZ = G.S.LZ, LXP = G.S.LXP;
  const int np = lr ? LZ : LX;
  std::vector<int> m(np), sl(np);
  for (int k = 0; k < np; ++k) {
    int a, b;  // lattice ids (1-based) of the pair, in the order of the sorted boundary node lists
    if (lr) {
      a = k * LXP + 1;           // tag 4, left
      b = k * LXP + LX;          // tag 2, right
      if (master_tag == 2) std::swap(a, b);
    } else {
      a = k + 1;                  // tag 1, bottom
      b = (LZ - 1) * LXP + k + 1; // tag 3, top
      if (master_tag == 3) std::swap(a, b);
    }
    m[k] = a;
    sl[k] = b;
  }
  Eb->add_periodic(np, m.data(), sl.data());
  // the mass of the pairs is summed (bc_periodic.f90:73); rmass still holds the mass here
  DevBuf<int> dm, ds;
  dm.upload(m);
  ds.upload(sl);
  if (Eb->prec == 8)
    k_periodic<double><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<double>(Eb)->rmass.p, Eb->npoin, G.ndof, np, dm.p, ds.p);
  else
    k_periodic<float><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<float>(Eb)->rmass.p, Eb->npoin, G.ndof, np, dm.p, ds.p);
  S2D_CUDA(cudaStreamSynchronize(Eb->stream));
  if (lr) S.perio_lr = 1;
  else S.perio_bt = 1;
  CART_GUARD_END
}

// Topology of a fault on the box, as BC_DYNFLT_init builds it (bc_dynflt.f90:258-354): tags (5,6) = the
// split-node row of MESH_CART ezflt (5 = lower side, 6 = upper side); tag 1 or 3 alone = the bottom / top
// side as a one-sided fault with the symmetry assumption (tags(2) = 0, :700-716).  Nodes in ascending x.
struct FaultTopo {
  int np = 0;
  bool two = false;
  std::vector<int> node1, node2;
  std::vector<double> n1, B, coord;
};
static void cart_fault_topology(const CartState& S, int tag1, int tag2, FaultTopo& F) {
  const CartGeom& G = S.G;
  const int N = G.N, ndof = G.ndof;
  F.two = (tag1 == 5 && tag2 == 6);
  S2D_REQUIRE(F.two || ((tag1 == 1 || tag1 == 3) && tag2 == 0),
              "cart fault: tags must be (5,6) = the ezflt split-node row, or (1,0) / (3,0) = a one-sided fault on the bottom / top side");
  if (F.two) S2D_REQUIRE(G.ezflt > 0, "cart fault: the mesh has no split-node row (ezflt = 0)");
  const int np = G.nx * (N - 1) + 1;
  F.np = np;
  F.node1.assign(np, 0);
  F.node2.assign(F.two ? np : 0, 0);
  F.n1.assign((size_t)np * 2, 0.0);
  F.B.assign((size_t)np * ndof, 0.0);
  F.coord.assign((size_t)np * 2, 0.0);
  // element row and local row of side 1, outward normal (t_z,-t_x) of that side, fault depth
  int iz1, j1, iz2 = 0, j2 = 0;
  double nz, zf;
  if (F.two) {
    iz1 = G.ezflt - 1; j1 = N - 1; iz2 = G.ezflt; j2 = 0; nz = 1.0; zf = G.z0 + G.hz * G.ezflt;   // edge_U of the lower side
  } else if (tag1 == 1) {
    iz1 = 0; j1 = 0; nz = -1.0; zf = G.z0;                                                         // edge_D
  } else {
    iz1 = G.nz - 1; j1 = N - 1; nz = 1.0; zf = G.z0 + G.hz * G.nz;                                 // edge_U
  }
  const int e_lo = G.halo_left ? -1 : 0, e_hi = G.halo_right ? G.nx : G.nx - 1;  // virtual elements of neighbour strips
  for (int e = e_lo; e <= e_hi; ++e)
    for (int i = 0; i < N; ++i) {
      const int pos = e * (N - 1) + i;
      if (pos < 0 || pos >= np) continue;
      F.B[pos] += G.wgll[i] * (0.5 * G.hx);  // BC_get_normal_and_weights (spec_grid.f90:961-1011)
      if (e < 0 || e >= G.nx) continue;
      F.node1[pos] = (int)cart_lat_id(G, e, iz1, i, j1);
      if (F.two) F.node2[pos] = (int)cart_lat_id(G, e, iz2, i, j2);
      F.coord[2 * pos] = gx_of(G, e, i);
      F.coord[2 * pos + 1] = zf;
      F.n1[pos] = 0.0;
      F.n1[pos + np] = nz;
    }
  if (S.perio_lr) {  // BC_get_normal_and_weights(..., periodic) (spec_grid.f90:1000-1005)
    F.B[0] = F.B[0] + F.B[np - 1];
    F.B[np - 1] = F.B[0];
  }
  if (ndof == 2)
    for (int k = 0; k < np; ++k) F.B[k + np] = F.B[k];
}

// BC_DYNFLT_init (bc_dynflt.f90:231-520) on the box.  `law` carries what the host read and evaluated at the
// fault nodes (T0, cohesion, V0, the friction-law arrays, the normal-stress law, the output strides); the
// topology members (node1, node2, n1, B, invM1, invM2, Z, coord, CoefA2V, CoefA2D) are filled here.
static int cart_add_dynflt(EngineBase* Eb, CartState& S, int tag1, int tag2, const s2d_dynflt_desc& law) {
  const CartGeom& G = S.G;
  S2D_REQUIRE(!Eb->committed, "cart_add_dynflt after commit");
  S.bc_added = true;
  const int ndof = G.ndof;
  FaultTopo F;
  cart_fault_topology(S, tag1, tag2, F);
  const int np = F.np;
  S2D_REQUIRE(law.np == np, "cart_add_dynflt: np does not match the fault (use s2d_cart_fault_nodes)");
  S2D_REQUIRE(law.T0, "cart_add_dynflt: T0 missing");
  // invM at the fault nodes from the current mass (bc_dynflt.f90:338-343)
  std::vector<double> m1((size_t)np * ndof), m2((size_t)np * ndof, 0.0);
  {
    DevBuf<int> dn1, dn2;
    DevBuf<double> o1, o2;
    dn1.upload(F.node1);
    o1.alloc((size_t)np * ndof);
    if (F.two) {
      dn2.upload(F.node2);
      o2.alloc((size_t)np * ndof);
    }
    if (Eb->prec == 8) {
      k_gather_nodes<double><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<double>(Eb)->rmass.p, Eb->npoin, ndof, np, dn1.p, o1.p);
      if (F.two) k_gather_nodes<double><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<double>(Eb)->rmass.p, Eb->npoin, ndof, np, dn2.p, o2.p);
    } else {
      k_gather_nodes<float><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<float>(Eb)->rmass.p, Eb->npoin, ndof, np, dn1.p, o1.p);
      if (F.two) k_gather_nodes<float><<<ceil_div(np, 128), 128, 0, Eb->stream>>>(as_engine<float>(Eb)->rmass.p, Eb->npoin, ndof, np, dn2.p, o2.p);
    }
    S2D_CUDA(cudaStreamSynchronize(Eb->stream));
    o1.download(m1.data());
    if (F.two) o2.download(m2.data());
  }
  std::vector<double> invM1(m1.size()), invM2(m1.size(), 0.0), Z(m1.size());
  const double A2V = S.CoefA2V();
  for (size_t q = 0; q < m1.size(); ++q) {
    invM1[q] = 1.0 / m1[q];
    if (F.two) {
      invM2[q] = 1.0 / m2[q];
      Z[q] = 1.0 / (A2V * F.B[q] * (invM1[q] + invM2[q]));  // bc_dynflt.f90:349-354
    } else {
      Z[q] = 0.5 / (A2V * F.B[q] * invM1[q]);
    }
  }
  s2d_dynflt_desc d = law;
  d.node1 = F.node1.data();
  d.node2 = F.two ? F.node2.data() : nullptr;
  d.n1 = F.n1.data();
  d.B = F.B.data();
  d.invM1 = invM1.data();
  d.invM2 = F.two ? invM2.data() : nullptr;
  d.Z = Z.data();
  d.coord = F.coord.data();
  std::vector<double> coh;
  if (!d.cohesion) {
    coh.assign(np, 0.0);
    d.cohesion = coh.data();
  }
  d.CoefA2V = A2V;
  d.CoefA2D = S.CoefA2D();
  d.oix1 = std::max(d.oix1, 1);
  d.oixn = std::min(d.oixn, np);
  const int id = Eb->add_dynflt(d);
  CartState::FaultInfo fi;
  fi.coord = F.coord;
  fi.T0.assign(law.T0, law.T0 + (size_t)np * 2);
  fi.B.assign(F.B.begin(), F.B.begin() + np);
  if ((int)S.faults.size() <= id) S.faults.resize(id + 1);
  S.faults[id] = fi;
  return id;
}

int s2d_cart_fault_nodes(s2d_handle h, int32_t tag1, int32_t tag2, int32_t* np, double* coord) {
  CART_GUARD_BEGIN
  FaultTopo F;
  cart_fault_topology(S, tag1, tag2, F);
  if (np) *np = F.np;
  if (coord) std::copy(F.coord.begin(), F.coord.end(), coord);
  CART_GUARD_END
}

int s2d_cart_add_dynflt(s2d_handle h, int32_t tag1, int32_t tag2, const s2d_dynflt_desc* law, int32_t* fault_id) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(law, "cart_add_dynflt: null descriptor");
  const int id = cart_add_dynflt(Eb, S, tag1, tag2, *law);
  if (fault_id) *fault_id = id;
  CART_GUARD_END
}

// the two-sided fault of ezflt with linear slip weakening, uniform parameters and a nucleation patch
int s2d_cart_add_fault_swf(s2d_handle h, double Dc, double MuS, double MuD, double Tn, double Tt, double Tt_nuc,
                           double x_nuc, double half_nuc, int32_t oixd, int32_t oitd, int32_t nt_max,
                           int32_t* fault_id) {
  CART_GUARD_BEGIN
  FaultTopo F;
  cart_fault_topology(S, 5, 6, F);
  const int np = F.np;
  std::vector<double> T0((size_t)np * 2), dc(np, Dc), mus(np, MuS), mud(np, MuD), pw(np, 3.0), alpha(np, 0.0);
  for (int k = 0; k < np; ++k) {
    const bool nuc = std::fabs(F.coord[2 * k] - x_nuc) <= half_nuc;  // DIST_PWCONR with one radius (distribution_pwconr.f90:69-85)
    T0[k] = nuc ? Tt_nuc : Tt;                                       // bc_dynflt.f90:392-400 with no background stress
    T0[k + np] = Tn;
  }
  s2d_dynflt_desc d;
  std::memset(&d, 0, sizeof(d));
  d.np = np;
  d.T0 = T0.data();
  d.allow_opening = 1;
  d.swf_kind = 1;
  d.swf_dc = dc.data();
  d.swf_mus = mus.data();
  d.swf_mud = mud.data();
  d.swf_p = pw.data();
  d.swf_alpha = alpha.data();
  d.normal_kind = 1;
  d.normal_T = d.normal_L = d.normal_V = 1.0;
  d.oix1 = 1;
  d.oixn = np;
  d.oixd = oixd;
  d.oit = 0;
  d.oitd = oitd;
  d.nt_max = nt_max;
  const int id = cart_add_dynflt(Eb, S, 5, 6, d);
  if (fault_id) *fault_id = id;
  CART_GUARD_END
}

// bc_DIRNEU_init (bc_dirneu.f90:117-146) on a side of the box: kinds 1 = Neumann, 2 = Dirichlet per component
int s2d_cart_add_dirneu(s2d_handle h, int32_t side, int32_t kind_h, int32_t kind_v) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(side >= 1 && side <= 4, "cart_add_dirneu: side tag must be 1..4");
  S2D_REQUIRE(!Eb->committed, "cart_add_dirneu after commit");
  const CartGeom& G = S.G;
  S2D_REQUIRE(!(side == 4 && G.halo_left) && !(side == 2 && G.halo_right),
              "cart_add_dirneu: that side is a strip interface, not a physical boundary");
  S.bc_added = true;
  const int LX = G.S.LX, LZ = G.S.LZ, LXP = G.S.LXP;
  const bool horiz = (side == 1 || side == 3);
  const int np = horiz ? LX : LZ;
  std::vector<int> node(np);
  for (int k = 0; k < np; ++k)
    node[k] = side == 1 ? k + 1 : side == 3 ? (LZ - 1) * LXP + k + 1 : side == 4 ? k * LXP + 1 : k * LXP + LX;
  Eb->add_dirneu(np, node.data(), kind_h, kind_v, nullptr, nullptr);
  CART_GUARD_END
}

int s2d_cart_add_force(s2d_handle h, double x, double z, const double* dir, int32_t* src_id) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(dir, "cart_add_force: null dir");
  const CartGeom& G = S.G;
  int ex, i, ez, j;
  nearest_1d(G, x, G.x0, G.hx, G.nx, ex, i);
  nearest_1d(G, z, G.z0, G.hz, G.nz, ez, j);
  const int id = Eb->add_force((int)cart_lat_id(G, ex, ez, i, j), dir);
  if (src_id) *src_id = id;
  CART_GUARD_END
}

// SRC_MOMENT_init (src_moment.f90:129-180) on the box: for every element that holds the source node, in
// ascending element order (SE_node_belongs_to), the N nodes of its xi-line and of its eta-line through
// the node with G = M * transpose(jac_inv) (P-SV) or jac_inv * M (SH); jac_inv = diag(2/hx, 2/hz).
int s2d_cart_add_moment(s2d_handle h, double x, double z, const double* M, int32_t* src_id) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(M, "cart_add_moment: null M");
  const CartGeom& G = S.G;
  const int N = G.N, ndof = G.ndof;
  int ex, i, ez, j;
  nearest_1d(G, x, G.x0, G.hx, G.nx, ex, i);
  nearest_1d(G, z, G.z0, G.hz, G.nz, ez, j);
  const double dxi = 2.0 / G.hx, deta = 2.0 / G.hz;
  // elements containing lattice node (ex,i ; ez,j): the node may also be the last point of the element
  // to the left / below (not across the split fault row)
  std::vector<std::array<int, 4>> holders;  // ix, iz, i, j
  for (int dz = -1; dz <= 0; ++dz)
    for (int dx = -1; dx <= 0; ++dx) {
      int ix = ex, li = i, iz = ez, lj = j;
      if (dx == -1) {
        if (!(i == 0 && ex > 0)) continue;
        ix = ex - 1;
        li = N - 1;
      }
      if (dz == -1) {
        if (!(j == 0 && ez > 0 && !row_detached(G, ez))) continue;
        iz = ez - 1;
        lj = N - 1;
      }
      holders.push_back({ix, iz, li, lj});
    }
  // nearest_1d prefers the higher element on ties, so also look right / up
  if (i == N - 1 && ex + 1 < G.nx) {
    const size_t n0 = holders.size();
    for (size_t q = 0; q < n0; ++q) holders.push_back({holders[q][0] + 1, holders[q][1], 0, holders[q][3]});
  }
  if (j == N - 1 && ez + 1 < G.nz && !row_detached(G, ez + 1)) {
    const size_t n0 = holders.size();
    for (size_t q = 0; q < n0; ++q) holders.push_back({holders[q][0], holders[q][1] + 1, holders[q][2], 0});
  }
  std::sort(holders.begin(), holders.end(), [&](const std::array<int, 4>& a, const std::array<int, 4>& b) {
    return (long long)a[1] * G.nx + a[0] < (long long)b[1] * G.nx + b[0];
  });
  const int nt = (int)holders.size() * 2 * N;
  std::vector<int> node(nt);
  std::vector<double> coef((size_t)nt * ndof);
  int t = 0;
  for (auto& hd : holders) {
    double gx1[2], ge1[2];  // G(c,1), G(c,2) per component
    if (ndof == 2) {  // G = M * transpose(jac_inv): G(c,1) = M(c,1)*dxi, G(c,2) = M(c,2)*deta   (M column-major (2,2))
      for (int c = 0; c < 2; ++c) {
        gx1[c] = M[c + 2 * 0] * dxi;
        ge1[c] = M[c + 2 * 1] * deta;
      }
    } else {  // G(:,1) = jac_inv * M(:,1)
      gx1[0] = dxi * M[0];
      ge1[0] = deta * M[1];
    }
    for (int k = 0; k < N; ++k, ++t) {  // iglob_xi(:,k) = ibool(:,j,e), coef_xi = G(c,1)*hprime(:,i)
      node[t] = (int)cart_lat_id(G, hd[0], hd[1], k, hd[3]);
      for (int c = 0; c < ndof; ++c) coef[t + (size_t)nt * c] = gx1[c] * S.H[k + N * hd[2]];
    }
    for (int k = 0; k < N; ++k, ++t) {  // iglob_eta(:,k) = ibool(i,:,e), coef_eta = G(c,2)*hprime(:,j)
      node[t] = (int)cart_lat_id(G, hd[0], hd[1], hd[2], k);
      for (int c = 0; c < ndof; ++c) coef[t + (size_t)nt * c] = ge1[c] * S.H[k + N * hd[3]];
    }
  }
  const int id = Eb->add_moment(nt, node.data(), coef.data());
  if (src_id) *src_id = id;
  CART_GUARD_END
}

int s2d_cart_add_receivers(s2d_handle h, int32_t nx, double xa, double za, double xb, double zb, char field,
                           int32_t isamp, int32_t nt_rec) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(nx >= 1, "cart_add_receivers: nx < 1");
  const CartGeom& G = S.G;
  std::vector<int> ig;
  for (int n = 0; n < nx; ++n) {  // REC_LINE (receivers.f90:120-130) + nearest node, duplicates dropped (:255)
    const double w = nx > 1 ? (double)n / (double)(nx - 1) : 0.0;
    int ex, i, ez, j;
    nearest_1d(G, xa + w * (xb - xa), G.x0, G.hx, G.nx, ex, i);
    nearest_1d(G, za + w * (zb - za), G.z0, G.hz, G.nz, ez, j);
    const int id = (int)cart_lat_id(G, ex, ez, i, j);
    bool dup = false;
    if (ig.size() > 1)
      for (int v : ig) dup = dup || (v == id);
    if (!dup) {
      ig.push_back(id);
      S.rec_coord.push_back(gx_of(G, ex, i));
      S.rec_coord.push_back(gz_of(G, ez, j));
    }
  }
  Eb->add_receivers((int)ig.size(), field, isamp, nt_rec, 1, ig.data(), nullptr, nullptr);
  CART_GUARD_END
}

// Kelvin-Voigt viscosity of the box (mat_kelvin_voigt.f90): eta per GLL node in the caller's numbering,
// already multiplied by dt when ETAxDT (:127)
int s2d_cart_set_kv(s2d_handle h, const double* eta_node) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(eta_node, "cart_set_kv: null eta");
  Eb->set_node_kv(eta_node);
  CART_GUARD_END
}

// REC_posit with AtNode = F (receivers.f90:262-300): the nearest node picks the element (the first one that
// holds it, SE_node_belongs_to), the point is located inside it (FE_find_point: affine for a rectangle) and
// the Lagrange weights of SE_init_interpol (spec_grid.f90:385-409) multiply the N*N nodes of that element
int s2d_cart_add_receivers_interp(s2d_handle h, int32_t nx, double xa, double za, double xb, double zb, char field,
                                  int32_t isamp, int32_t nt_rec) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(nx >= 1, "cart_add_receivers: nx < 1");
  const CartGeom& G = S.G;
  const int N = G.N, N2 = N * N;
  std::vector<int> nodes((size_t)N2 * nx);
  std::vector<double> interp((size_t)N2 * nx);
  S.rec_coord.clear();
  auto lagrange = [&](double xi, double* hq) {
    for (int i = 0; i < N; ++i) {
      double p = 1.0;
      for (int m = 0; m < N; ++m)
        if (m != i) p *= (xi - G.xgll[m]) / (G.xgll[i] - G.xgll[m]);
      hq[i] = p;
    }
  };
  for (int n = 0; n < nx; ++n) {
    const double w = nx > 1 ? (double)n / (double)(nx - 1) : 0.0;
    const double x = xa + w * (xb - xa), z = za + w * (zb - za);
    int ex, i, ez, j;
    nearest_1d(G, x, G.x0, G.hx, G.nx, ex, i);
    nearest_1d(G, z, G.z0, G.hz, G.nz, ez, j);
    // first element (ascending index) that holds the node
    if (i == 0 && ex > 0) ex -= 1;
    if (j == 0 && ez > 0 && !row_detached(G, ez)) ez -= 1;
    const double xi = 2.0 * (x - (G.x0 + G.hx * ex)) / G.hx - 1.0, eta = 2.0 * (z - (G.z0 + G.hz * ez)) / G.hz - 1.0;
    std::vector<double> hx(N), hz(N);
    lagrange(xi, hx.data());
    lagrange(eta, hz.data());
    for (int jj = 0; jj < N; ++jj)
      for (int ii = 0; ii < N; ++ii) {
        nodes[(size_t)N2 * n + ii + N * jj] = (int)cart_lat_id(G, ex, ez, ii, jj);
        interp[(size_t)N2 * n + ii + N * jj] = hx[ii] * hz[jj];
      }
    S.rec_coord.push_back(x);  // the station keeps its position (rec%coord = newcoord of FE_find_point)
    S.rec_coord.push_back(z);
  }
  Eb->add_receivers_nodes(nx, field, isamp, nt_rec, nodes.data(), interp.data());
  CART_GUARD_END
}

int s2d_cart_fault_info(s2d_handle h, int32_t fault_id, int32_t* np, double* coord, double* T0, double* B) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(fault_id >= 0 && fault_id < (int)S.faults.size() && !S.faults[fault_id].coord.empty(), "cart_fault_info: unknown fault");
  const CartState::FaultInfo& fi = S.faults[fault_id];
  if (np) *np = (int32_t)(fi.coord.size() / 2);
  if (coord) std::copy(fi.coord.begin(), fi.coord.end(), coord);
  if (T0) std::copy(fi.T0.begin(), fi.T0.end(), T0);
  if (B) std::copy(fi.B.begin(), fi.B.end(), B);
  CART_GUARD_END
}

int s2d_cart_receiver_info(s2d_handle h, int32_t* nx, double* coord) {
  CART_GUARD_BEGIN
  if (nx) *nx = (int32_t)(S.rec_coord.size() / 2);
  if (coord) std::copy(S.rec_coord.begin(), S.rec_coord.end(), coord);
  CART_GUARD_END
}

int s2d_cart_get_gll(s2d_handle h, double* xgll, double* wgll, double* hprime) {
  CART_GUARD_BEGIN
  const int N = S.G.N;
  if (xgll) std::copy(S.G.xgll, S.G.xgll + N, xgll);
  if (wgll) std::copy(S.G.wgll, S.G.wgll + N, wgll);
  if (hprime) std::copy(S.H, S.H + N * N, hprime);
  CART_GUARD_END
}

int s2d_cart_get(s2d_handle h, int32_t* ibool, double* a, double* rmass, double* coord) {
  CART_GUARD_BEGIN
  const CartGeom& G = S.G;
  const int N = G.N, N2 = N * N;
  const long long tot = (long long)Eb->nelem * N2;
  // element e of the caller's order sits at (ix,iz) of the box
  auto elem_pos = [&](size_t e, int& ix, int& iz) {
    const size_t old = S.renumber ? (size_t)S.perm[e] : e;
    ix = (int)(old % G.nx);
    iz = (int)(old / G.nx);
  };
  if (ibool && S.renumber) {
    for (size_t e = 0; e < (size_t)Eb->nelem; ++e) {
      int ix, iz;
      elem_pos(e, ix, iz);
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) ibool[e * N2 + i + N * j] = (int32_t)cart_ref_id(S, ix, iz, i, j);
    }
  } else if (ibool) {
    DevBuf<int> ib;
    ib.alloc((size_t)tot);
    k_cart_ibool<<<(unsigned)((tot + 255) / 256), 256, 0, Eb->stream>>>(G, ib.p);
    S2D_CUDA(cudaStreamSynchronize(Eb->stream));
    ib.download(ibool);
  }
  if (a) {
    // reference layout a(ngll,ngll,nelast,nelem) from the strip layout
    const int nelast = (G.ndof == 1) ? 2 : 6;
    std::vector<double> pc;
    int compact = 0;
    if (Eb->prec == 8) {
      pc = as_engine<double>(Eb)->p_coef.to_host();
      compact = as_engine<double>(Eb)->cart_compact;
    } else {
      std::vector<float> t = as_engine<float>(Eb)->p_coef.to_host();
      pc.assign(t.begin(), t.end());
      compact = as_engine<float>(Eb)->cart_compact;
    }
    const double cdx = 2.0 / G.hx, cdz = 2.0 / G.hz, det = (0.5 * G.hx) * (0.5 * G.hz);
    for (size_t e = 0; e < (size_t)Eb->nelem; ++e) {
      {
        int ix, iz;
        elem_pos(e, ix, iz);
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i) {
            double av[6];
            if (compact) {  // the planes as the strip kernel forms them from (lambda, mu)
              const double la = pc[strip_coef_index(G.S, 2, ix, iz, i, j, 0)];
              const double mu = pc[strip_coef_index(G.S, 2, ix, iz, i, j, 1)];
              const double kx = la + 2.0 * mu, nw = -(det * (G.wgll[i] * G.wgll[j]));
              av[0] = nw * ((kx * cdx) * cdx);
              av[1] = nw * ((la * cdx) * cdz);
              av[2] = nw * ((kx * cdz) * cdz);
              av[3] = nw * ((mu * cdz) * cdz);
              av[4] = nw * ((mu * cdx) * cdz);
              av[5] = nw * ((mu * cdx) * cdx);
            } else {
              for (int pl = 0; pl < nelast; ++pl) av[pl] = pc[strip_coef_index(G.S, nelast, ix, iz, i, j, pl)];
            }
            for (int pl = 0; pl < nelast; ++pl) a[(e * nelast + pl) * N2 + i + N * j] = av[pl];
          }
      }
    }
  }
  if (rmass) {
    const size_t nd = Eb->npoin_ref * G.ndof;
    DevBuf<double> tmp;
    tmp.alloc(nd);
    if (Eb->prec == 8) as_engine<double>(Eb)->cart_to_ref(as_engine<double>(Eb)->rmass.p, tmp.p);
    else as_engine<float>(Eb)->cart_to_ref(as_engine<float>(Eb)->rmass.p, tmp.p);
    S2D_CUDA(cudaStreamSynchronize(Eb->stream));
    tmp.download(rmass);
    if (!Eb->committed)
      for (size_t q = 0; q < nd; ++q) rmass[q] = 1.0 / rmass[q];
  }
  if (coord) {
    for (int iz = 0; iz < G.nz; ++iz)
      for (int ix = 0; ix < G.nx; ++ix)
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i) {
            const size_t nd = (size_t)(cart_ref_id(S, ix, iz, i, j) - 1);
            coord[2 * nd] = gx_of(G, ix, i);
            coord[2 * nd + 1] = gz_of(G, iz, j);
          }
  }
  CART_GUARD_END
}

}  // extern "C"
template <typename T>
static void cart_fill(Engine<T>& E, const CartGeom& G, uint64_t seed, double amp_d, double amp_v) {
  const long long tot = (long long)G.S.LX * G.S.LZ;
  E.a.zero(E.stream);
  k_cart_fill<T><<<(unsigned)((tot + 255) / 256), 256, 0, E.stream>>>(G, E.dbuf().p, E.v.p, E.npoin, seed, amp_d, amp_v);
  S2D_CUDA(cudaGetLastError());
  E.pred_valid = false;
  S2D_CUDA(cudaStreamSynchronize(E.stream));
}
extern "C" int s2d_cart_fill_fields(s2d_handle h, uint64_t seed, double amp_d, double amp_v) {
  CART_GUARD_BEGIN
  if (Eb->prec == 8) cart_fill<double>(*as_engine<double>(Eb), S.G, seed, amp_d, amp_v);
  else cart_fill<float>(*as_engine<float>(Eb), S.G, seed, amp_d, amp_v);
  CART_GUARD_END
}

template <typename T>
static void cart_window(Engine<T>& E, const CartGeom& G, int gx0, int gz0, int nwx, int nwz, double* d, double* v,
                        double* a) {
  const size_t n = (size_t)nwx * nwz * G.ndof;
  DevBuf<double> tmp;
  tmp.alloc(n);
  if (a) E.ensure_accel();
  const T* src[3] = {E.dbuf().p, E.v.p, E.a.p};
  double* dst[3] = {d, v, a};
  for (int k = 0; k < 3; ++k) {
    if (!dst[k]) continue;
    k_cart_window<T><<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(src[k], tmp.p, E.npoin, G.S.LXP, G.ndof, gx0, gz0, nwx, nwz);
    S2D_CUDA(cudaGetLastError());
    S2D_CUDA(cudaStreamSynchronize(E.stream));
    tmp.download(dst[k]);
  }
}
template <typename T>
static void cart_snapshot(Engine<T>& E, CartState& S, char what, float* out) {
  const CartGeom G = S.dev_geom();
  const int N2 = G.N * G.N, ncomp = (what == 'E' || what == 'S') ? G.ndof + 1 : 1;
  const size_t n = (size_t)ncomp * E.nelem * N2;
  DevBuf<float> buf;
  buf.alloc(n);
  DevBuf<int> dperm;
  if (S.renumber) dperm.upload(S.perm);
  SnapH Hm;
  std::memcpy(Hm.H, S.H, sizeof(Hm.H));
  const long long tot = (long long)E.nelem * N2;
  S2D_CUDA(cudaStreamSynchronize(E.stream));
  k_cart_snap<T><<<(unsigned)((tot + 127) / 128), 128, 0, E.stream>>>(G, Hm, E.dbuf().p, E.v.p, E.strip_eta.n ? E.strip_eta.p : nullptr,
                                                                       E.pl_ep.n ? E.pl_ep.p : nullptr, E.npoin, (int)what, S.renumber ? dperm.p : nullptr, buf.p);
  S2D_CUDA(cudaGetLastError());
  S2D_CUDA(cudaStreamSynchronize(E.stream));
  buf.download(out);
}
extern "C" int s2d_cart_snapshot_elem(s2d_handle h, char what, float* out) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(out != nullptr, "cart_snapshot_elem: null pointer");
  S2D_REQUIRE(what == 'E' || what == 'S' || what == 'd' || what == 'c', "cart_snapshot_elem: field must be E, S, d or c");
  S2D_REQUIRE(!((what == 'd' || what == 'c') && S.G.ndof != 2), "mat_gen:MAT_divcurl_gen: only for P-SV");
  S2D_REQUIRE(Eb->committed, "cart_snapshot_elem before commit");
  if (Eb->prec == 8) cart_snapshot<double>(*as_engine<double>(Eb), S, what, out);
  else cart_snapshot<float>(*as_engine<float>(Eb), S, what, out);
  CART_GUARD_END
}
extern "C" {
int s2d_cart_get_window(s2d_handle h, int32_t gx0, int32_t gz0, int32_t nwx, int32_t nwz, double* d, double* v, double* a) {
  CART_GUARD_BEGIN
  const CartGeom& G = S.G;
  S2D_REQUIRE(gx0 >= 0 && gz0 >= 0 && nwx >= 1 && nwz >= 1 && gx0 + nwx <= G.S.LX && gz0 + nwz <= G.S.LZ,
              "cart_get_window: window outside the lattice");
  if (Eb->prec == 8) cart_window<double>(*as_engine<double>(Eb), G, gx0, gz0, nwx, nwz, d, v, a);
  else cart_window<float>(*as_engine<float>(Eb), G, gx0, gz0, nwx, nwz, d, v, a);
  CART_GUARD_END
}

int s2d_halo_info(s2d_handle h, int64_t* count, void** send_dev, void** recv_dev) {
  CART_GUARD_BEGIN
  Eb->halo_info(count, send_dev, recv_dev);
  CART_GUARD_END
}
int s2d_halo_set_exchange(s2d_handle h, s2d_exchange_fn fn, void* user) {
  CART_GUARD_BEGIN
  Eb->halo_set_exchange(fn, user);
  CART_GUARD_END
}
int s2d_halo_set_peers(s2d_handle h, void* left_recv_dev, void* right_recv_dev, void* left_flag_dev,
                       void* right_flag_dev) {
  CART_GUARD_BEGIN
  Eb->halo_set_peers(left_recv_dev, right_recv_dev, left_flag_dev, right_flag_dev);
  CART_GUARD_END
}

// CUDA IPC plumbing for one process per GPU: blob = handles of {recv[0], recv[1], flags}
int s2d_halo_ipc_export(s2d_handle h, void* blob) {
  CART_GUARD_BEGIN
  S2D_REQUIRE(blob, "halo_ipc_export: null blob");
  void* recv[2] = {nullptr, nullptr};
  void* flags = nullptr;
  Eb->halo_peer_buffers(recv, &flags);
  cudaIpcMemHandle_t hd[3];
  std::memset(hd, 0, sizeof(hd));
  if (recv[0]) S2D_CUDA(cudaIpcGetMemHandle(&hd[0], recv[0]));
  if (recv[1]) S2D_CUDA(cudaIpcGetMemHandle(&hd[1], recv[1]));
  S2D_CUDA(cudaIpcGetMemHandle(&hd[2], flags));
  std::memcpy(blob, hd, sizeof(hd));
  CART_GUARD_END
}

int s2d_halo_ipc_open(s2d_handle h, const void* left_blob, const void* right_blob) {
  CART_GUARD_BEGIN
  void *lr = nullptr, *lf = nullptr, *rr = nullptr, *rf = nullptr;
  cudaIpcMemHandle_t hd[3];
  if (left_blob) {  // I am the left neighbour's RIGHT side: its recv[1] and flags[1]
    std::memcpy(hd, left_blob, sizeof(hd));
    S2D_CUDA(cudaIpcOpenMemHandle(&lr, hd[1], cudaIpcMemLazyEnablePeerAccess));
    S2D_CUDA(cudaIpcOpenMemHandle(&lf, hd[2], cudaIpcMemLazyEnablePeerAccess));
    lf = (unsigned long long*)lf + 1;
  }
  if (right_blob) {  // I am the right neighbour's LEFT side: its recv[0] and flags[0]
    std::memcpy(hd, right_blob, sizeof(hd));
    S2D_CUDA(cudaIpcOpenMemHandle(&rr, hd[0], cudaIpcMemLazyEnablePeerAccess));
    S2D_CUDA(cudaIpcOpenMemHandle(&rf, hd[2], cudaIpcMemLazyEnablePeerAccess));
  }
  Eb->halo_set_peers(lr, rr, lf, rf);
  CART_GUARD_END
}

int s2d_halo_peer_buffers(s2d_handle h, void** recv_dev, void** flags_dev) {
  CART_GUARD_BEGIN
  Eb->halo_peer_buffers(recv_dev, flags_dev);
  CART_GUARD_END
}

}  // extern "C"
