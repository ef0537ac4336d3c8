// Element-force + assembly (+ node update) kernel for structured (MESH_CART) grids: the z-marching
// strip kernel.
//
// Replaces, for the flat Cartesian box, compute_Fint's element loop (SRC/solver.f90:291-299), the
// gather/scatter of SRC/fields.f90:173-189,113-129 and ELAST_KD1/KD2_{SH,PSV} with mxm/My_MATMUL
// folded in (SRC/mat_elastic.f90:464-775, SRC/mxmlib.f90) and -- in its fused form -- the
// mass-inverse update of solve_leapfrog (SRC/solver.f90:151,157-158).  Same operator as
// elem_kernels.cuh:   Uxi = Ht U, Ueta = U H,  f = H (tH) + (tHt) Ht      (H(i,j) = h'_i(x_j))
//
// Layout.  Fields live on the GLL LATTICE: node (gx,gz) at gz*LXP + gx, LX = nx*(N-1)+1 columns in rows of
// pitch LXP >= LX + 1, a multiple of 8 elements (every row, hence every strip boundary, starts on a 32-byte
// DRAM sector; the pad columns hold zeros and are never written); the row of split fault nodes is stored
// twice (lower side, then upper side).  The box is cut into vertical
// strips EPW = floor(32/N) elements wide and horizontal bands of SEG element rows.  One WARP owns one
// (band, strip) and marches through it upward, one element row per iteration; one CTA is a GROUP of
// GW adjacent strips of the same band:
//   * lane (el,i) owns lattice column i of element el and keeps that column's N values of the
//     current element row in registers, so eta-contractions (along z) are register-local with
//     hprime as constant-bank operands, and xi-contractions go through a warp-private
//     shared-memory tile read back as 128-bit rows (__syncwarp only);
//   * the top node of a row is the bottom node of the next: its displacement and its partial
//     force sum are carried in registers (vertical assembly costs one add, no memory);
//   * columns shared by two elements of a strip are merged with one warp shuffle, columns shared
//     by two strips of a group are handed over through shared memory (one CTA barrier per row);
//   * every lattice node is written exactly once.  Only a group's right edge column and a band's
//     top row leave partial sums in small halo arrays, folded in by k_strip_fold in a fixed order
//     (deterministic, no atomics).
// Fused form: a node whose force is complete when its owner lane holds it -- every node except the
// "deferred" ones (halo columns / rows, boundary-condition and source rows or columns, flagged in
// rowflag / colflag) -- is advanced on the spot: a = rmass*f, v += dt*a, d_next = d + dt*v, so the
// force array is never written or re-read for it.  Deferred nodes get their force stored and are
// advanced by k_strip_deferred after the fold, the sources and the boundary conditions.
// Coefficient planes are stored per (band, strip, element row) as [plane pair][j][lane] 16-byte
// vectors: every warp load is one contiguous run, the whole array is read exactly once per step.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "elem_kernels.cuh"
#include "tensor_map.hpp"

// strips (warps) per CTA of the strip kernel; 4 is the measured best (see strip_warps())
#ifndef S2D_STRIP_WARPS
#define S2D_STRIP_WARPS 4
#endif

namespace s2d {

struct StripGeom {
  int N, ndof;
  int nx, nz, ezflt;       // elements; split-node fault after element row ezflt (0 = none)
  int EPW, W, WL;          // elements per strip, lattice columns per full strip, W+1
  int nstrips, SEG, nseg_lo, nseg;
  int LX, LZ;              // lattice extent (LZ counts the duplicated fault row)
  int LXP;                 // row pitch of the lattice arrays (>= LX + 1, multiple of 8)
  // groups of strips (one CTA each): [strip 0 alone if g_lead] [runs of GW strips] [last strip alone if g_tail]
  int GW, g_lead, g_tail, ngroups;
  // subset of groups handled by one launch: group = it_g0 + (k % it_ng) * it_step, band = k / it_ng
  int it_g0, it_ng, it_step;
  long long nitems;        // CTAs of this launch
  // x-strip interfaces with neighbour GPUs: the fold leaves lattice column 0 / LX-1 to the exchange
  int xhalo_left, xhalo_right;
};

__host__ __device__ inline void strip_group(const StripGeom& G, int g, int& first, int& count) {
  if (G.g_lead && g == 0) {
    first = 0;
    count = 1;
    return;
  }
  const int s0 = G.g_lead ? 1 : 0;
  const int nmid = G.nstrips - s0 - (G.g_tail ? 1 : 0);
  const int gm = g - s0;
  if (gm * G.GW < nmid) {
    first = s0 + gm * G.GW;
    count = min(G.GW, nmid - gm * G.GW);
  } else {
    first = G.nstrips - 1;
    count = 1;
  }
}
__host__ __device__ inline int strip_group_of(const StripGeom& G, int strip) {
  if (G.g_lead && strip == 0) return 0;
  const int s0 = G.g_lead ? 1 : 0;
  if (G.g_tail && strip == G.nstrips - 1) return G.ngroups - 1;
  return s0 + (strip - s0) / G.GW;
}
inline void strip_set_groups(StripGeom& G, int GW, bool lead, bool tail) {
  G.GW = GW;
  G.g_lead = (lead && G.nstrips > 1) ? 1 : 0;
  G.g_tail = (tail && G.nstrips > 1) ? 1 : 0;
  const int nmid = G.nstrips - G.g_lead - G.g_tail;
  G.ngroups = G.g_lead + (nmid + GW - 1) / GW + G.g_tail;
}

constexpr int STRIP_MASK_WORDS = 128;  // a band has at most 32*128 lattice rows (make_strip_geom clamps SEG)
// lattice / strip decomposition of an nx x nz box (split-node row after element row ezflt); one CTA = a group of
// adjacent strips, the strips next to a GPU interface form groups of their own
inline StripGeom make_strip_geom(int N, int ndof, int nx, int nz, int ezflt, int seg, bool halo_left, bool halo_right,
                                 int warps = S2D_STRIP_WARPS) {
  StripGeom Q{};
  Q.N = N;
  Q.ndof = ndof;
  Q.nx = nx;
  Q.nz = nz;
  Q.ezflt = ezflt;
  Q.EPW = 32 / N;
  Q.W = Q.EPW * (N - 1);
  Q.WL = Q.W + 1;
  Q.nstrips = (nx + Q.EPW - 1) / Q.EPW;
  Q.SEG = seg > 1 ? seg : 1;
  if (Q.SEG > (32 * STRIP_MASK_WORDS - 2) / (N - 1)) Q.SEG = (32 * STRIP_MASK_WORDS - 2) / (N - 1);  // the kernel's row mask
  Q.nseg_lo = ezflt > 0 ? (ezflt + Q.SEG - 1) / Q.SEG : 0;
  Q.nseg = Q.nseg_lo + (nz - ezflt + Q.SEG - 1) / Q.SEG;
  Q.LX = nx * (N - 1) + 1;
  Q.LXP = (Q.LX + 1 + 7) / 8 * 8;  // rows start on 32-byte sectors (FP32: 8 elements), one spare column for 16-byte bulk copies
  Q.LZ = nz * (N - 1) + 1 + (ezflt > 0 ? 1 : 0);
  Q.xhalo_left = halo_left ? 1 : 0;
  Q.xhalo_right = halo_right ? 1 : 0;
  strip_set_groups(Q, warps, halo_left, halo_right);
  Q.it_g0 = 0;
  Q.it_ng = Q.ngroups;
  Q.it_step = 1;
  Q.nitems = (long long)Q.nseg * Q.ngroups;
  return Q;
}

// band below the shared row gz, or -1 when gz is not the bottom row of a band that shares it
__host__ __device__ inline int strip_shared_row_seg(const StripGeom& G, int gz) {
  int g = gz, lower = 1;
  if (G.ezflt > 0 && gz >= G.ezflt * (G.N - 1) + 1) {
    g = gz - 1;
    lower = 0;
  }
  if (g % (G.N - 1) != 0) return -1;
  const int ez = g / (G.N - 1);
  if (G.ezflt > 0 && lower) {
    if (ez > 0 && ez < G.ezflt && ez % G.SEG == 0) return ez / G.SEG - 1;
  } else {
    if (ez > G.ezflt && ez < G.nz && (ez - G.ezflt) % G.SEG == 0) return G.nseg_lo + (ez - G.ezflt) / G.SEG - 1;
  }
  return -1;
}
__host__ __device__ inline void strip_seg_rows(const StripGeom& G, int seg, int& ez0, int& ez1) {
  if (seg < G.nseg_lo) {
    ez0 = seg * G.SEG;
    ez1 = min(ez0 + G.SEG, G.ezflt);
  } else {
    ez0 = G.ezflt + (seg - G.nseg_lo) * G.SEG;
    ez1 = min(ez0 + G.SEG, G.nz);
  }
}
__host__ __device__ inline int strip_seg_of(const StripGeom& G, int ez) {
  if (G.ezflt > 0 && ez < G.ezflt) return ez / G.SEG;
  return G.nseg_lo + (ez - G.ezflt) / G.SEG;
}
__host__ __device__ inline int strip_lat_row(const StripGeom& G, int ez, int j) {
  return ez * (G.N - 1) + j + ((G.ezflt > 0 && ez >= G.ezflt) ? 1 : 0);
}
// element columns [gex0, gex0 + gcx) of the group (CTA) that holds `strip`
__host__ __device__ inline void strip_group_cols(const StripGeom& G, int strip, int& gex0, int& gcx) {
  int gfirst, gcount;
  strip_group(G, strip_group_of(G, strip), gfirst, gcount);
  gex0 = gfirst * G.EPW;
  gcx = min((gfirst + gcount) * G.EPW, G.nx) - gex0;
}
// first element (in units of elements) of the coefficient block of element row ez of (seg, strip).  Per band the
// groups follow one another, per group the element rows, per row the strips of the group: the blocks one CTA needs
// for one element row are ONE contiguous run (one bulk copy), each strip's block inside it is contiguous too.
__host__ __device__ inline long long strip_elem_off(const StripGeom& G, int seg, int strip, int ez) {
  int ez0, ez1, gex0, gcx;
  strip_seg_rows(G, seg, ez0, ez1);
  strip_group_cols(G, strip, gex0, gcx);
  return (long long)ez0 * G.nx + (long long)gex0 * (ez1 - ez0) + (long long)(ez - ez0) * gcx + (strip * G.EPW - gex0);
}
// position (in scalars) of a(i,j,plane) of element (ix,iz) inside the strip layout
__host__ __device__ inline size_t strip_coef_index(const StripGeom& G, int nelast, int ix, int iz, int i, int j,
                                                   int pl) {
  const int N = G.N;
  const int seg = strip_seg_of(G, iz), strip = ix / G.EPW;
  const int ex0 = strip * G.EPW, el = ix - ex0;
  const int cx = min(G.EPW, G.nx - ex0);
  const size_t base = (size_t)strip_elem_off(G, seg, strip, iz) * nelast * N * N;
  return base + 2 * ((size_t)((pl >> 1) * N + j) * (cx * N) + el * N + i) + (pl & 1);
}

// position of a per-GLL-point scalar (e.g. the Kelvin-Voigt eta) of element (ix,iz): [j][lane] per element row
__host__ __device__ inline size_t strip_scalar_index(const StripGeom& G, int ix, int iz, int i, int j) {
  const int N = G.N;
  const int seg = strip_seg_of(G, iz), strip = ix / G.EPW;
  const int ex0 = strip * G.EPW, el = ix - ex0;
  const int cx = min(G.EPW, G.nx - ex0);
  return (size_t)strip_elem_off(G, seg, strip, iz) * N * N + (size_t)j * (cx * N) + el * N + i;
}

// lattice column of the boundary between group hb and group hb+1
__host__ __device__ inline int strip_halo_col(const StripGeom& G, int hb, int& strip_right) {
  int first, count;
  strip_group(G, hb + 1, first, count);
  strip_right = first;
  return first * G.W;
}

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

constexpr int STRIP_VS_MAXB = 8, STRIP_VS_TAB = 3 + 4 * STRIP_VS_MAXB;  // visco: mechanisms per material, table row
constexpr int STRIP_DM_TAB = 16;
constexpr int STRIP_PL_SETS = 8;  // plastic material sets per problem (set 0 = elastic elements)
// position of plastic-strain component k at GLL point (i,j) of element (ix,iz)
__host__ __device__ inline size_t strip_ep_index(const StripGeom& G, int ix, int iz, int i, int j, int k) {
  const int N = G.N;
  const int seg = strip_seg_of(G, iz), strip = ix / G.EPW;
  const int ex0 = strip * G.EPW, el = ix - ex0;
  const int cx = min(G.EPW, G.nx - ex0);
  return (size_t)strip_elem_off(G, seg, strip, iz) * 3 * N * N + (size_t)(k * N + j) * (cx * N) + el * N + i;
}
// position of plane k (of nplanes per element) at GLL point (i,j) of element (ix,iz): per-point state of the rheologies
__host__ __device__ inline size_t strip_plane_index(const StripGeom& G, int ix, int iz, int i, int j, int k, int nplanes) {
  const int N = G.N;
  const int seg = strip_seg_of(G, iz), strip = ix / G.EPW;
  const int ex0 = strip * G.EPW, el = ix - ex0;
  const int cx = min(G.EPW, G.nx - ex0);
  return (size_t)strip_elem_off(G, seg, strip, iz) * nplanes * N * N + (size_t)(k * N + j) * (cx * N) + el * N + i;
}
__host__ __device__ inline size_t strip_elem_slot(const StripGeom& G, int ix, int iz) {
  const int strip = ix / G.EPW;
  return (size_t)strip_elem_off(G, strip_seg_of(G, iz), strip, iz) + (ix - strip * G.EPW);
}

template <typename T, int N>
struct StripArgs {
  // TMA tensor maps of the lattice fields (TENS variant of the kernel): d[n] (LXP, LZ, ndof), v likewise, rmass
  // (LXP, LZ); boxes of strip_box_width() columns x (N-1) rows (x ndof)
  alignas(64) CUtensorMap tm_d;
  alignas(64) CUtensorMap tm_v;
  alignas(64) CUtensorMap tm_r;
  alignas(64) CUtensorMap tm_a;   // a[n-1] (explicit Newmark)
  StripGeom G;
  const T* coef;
  const T* d;
  T* f;
  T* halo_x;   // [c][ngroups-1][LZ]      partial sums of the column shared with the group to the right
  T* halo_z;   // [c][nseg][nstrips][WL]  partial sums of the row shared with the band above
  size_t npoin;
  // fused leapfrog update of the nodes that are not deferred
  const T* v_in;
  T* v_out;
  const T* rmass;
  T* d_next;
  T* a_out;                 // accelerations (may be null: not materialised)
  const uint8_t* rowflag;   // (LZ) != 0: every node of the lattice row is deferred
  const uint8_t* colflag;   // (LX) 1: every node of the lattice column is deferred; 2: group-boundary column
  int* meet;                // [nseg][ngroups-1] arrival counters of the group-boundary columns (zero between launches)
  T dt;
  // explicit Newmark (solver.f90:59-60,78-82 with beta = 0): c1 = dt^2/2, c2 = (1-gamma) dt, c3 = gamma dt;
  // leapfrog is the same update with c1 = c2 = 0, c3 = dt
  T c1, c2, c3;
  const T* a_in;            // accelerations of the previous step (Newmark predictor); aliases f / a_out
  // Kelvin-Voigt elements (MAT_KV_add_etav, mat_kelvin_voigt.f90:137-150): forces from d + eta*v, eta per GLL
  // point of every element in the strip layout (strip_scalar_index; zero on elements without KV)
  const T* eta;
  const T* v_kv;
  // 2.5D term of MAT_ELAST_add_25D_f (mat_elastic.f90:447-459, mat_gen.f90:440): f = f - beta*d element by element,
  // beta per GLL point of every element in the strip layout (strip_scalar_index), or nullptr
  const T* beta;
  // Coulomb plasticity (MAT_PLAST_stress, mat_plastic.f90:281-387; PLAST instantiation): material set of every
  // element in strip order (strip_elem_off + el; 0 = elastic), plastic strain per element GLL point
  // ([strain component][j][lane] per element row of a strip), per-set yield_co, yield_mu, vp_factor, e0(3)
  const unsigned char* pl_set;
  T* pl_ep;
  const T* pl_tab;          // [STRIP_PL_SETS][6] on the device
  // visco-elasticity (MAT_VISCO_stress, mat_visco.f90:206-248; the same instantiation, vs_state != null): memory
  // variables el(Nbody,3) and the strain of the previous evaluation per element GLL point, planes
  // [3 b + c | 3 vs_nb + c][j][lane] per element row of a strip; per set lambda_inf, mu_inf, Nbody, RK(8), theta(8,3)
  T* vs_state;
  const T* vs_tab;          // [STRIP_PL_SETS][STRIP_VS_TAB]
  int vs_nb;                // memory-variable planes per component in the state layout (max Nbody over the sets)
  // damage rheology (MAT_DMG_stress, mat_damage.f90:337-491; the same instantiation, dm_state != null): damage
  // variable alpha and plastic strain ep(3) per element GLL point, planes [alpha | ep11 | ep22 | ep12][j][lane];
  // per set lambda, mu, xi_0, gamma_r, beta, Cd, Cv, e0(3), s0(3), dt; *dm_err is set when the loss-of-convexity
  // checks of compute_stress (:478-489) fail (the reference aborts)
  T* dm_state;
  const T* dm_tab;          // [STRIP_PL_SETS][STRIP_DM_TAB]
  int* dm_err;
  int prefetch;             // L2 prefetch of what is not staged
  T H[N * N];               // hprime, column-major (constant bank)
  // compact coefficient mode (isotropic flat grids): only (lambda, mu) are stored per GLL point and
  // the six planes of MAT_ELAST_init_a (mat_elastic.f90:334-340,355-357) are formed in registers,
  // a_k = -weights * ((c * m1) * m2) with the same order of roundings as the stored planes
  T cdx, cdz, cdet;         // DxiDx = 2/hx, DetaDz = 2/hz, |J| = (hx/2)(hz/2)
  T wg[N];                  // GLL weights: weights(i,j) = |J| * (wg[i] * wg[j])
  T Hz[N * N];              // DetaDz * hprime, and DetaDz / DxiDx: the metric folded into the contractions
  T rzx;
};

__device__ __forceinline__ double rsqrt_any(double x) { return rsqrt(x); }
__device__ __forceinline__ float rsqrt_any(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rcp_any(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float rcp_any(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }

template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// asynchronous global -> shared copies (LDGSTS): the loads of the next element row are in flight
// while the current one is computed, without holding registers
template <int BYTES>
__device__ __forceinline__ void stage_copy(void* smem, const void* g) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
#ifndef S2D_STAGE16_CA
#define S2D_STAGE16_CA 0
#endif
  if constexpr (BYTES == 16 && !S2D_STAGE16_CA)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sa), "l"(g), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

#ifndef S2D_STRIP_STAGE
#define S2D_STRIP_STAGE 7
#endif
#ifndef S2D_STRIP_ROT
#define S2D_STRIP_ROT 0
#endif
// Measurement only (results are WRONG when non-zero): parts of the kernel switched off to see what each costs.
// bit 0: no warp-tile traffic (xi-contractions from the lane's own column), 1: no coefficient loads, 2: no
// displacement loads, 3: no v / rmass loads, 4: no v / d_next stores, 5: no per-row CTA barrier.
// profiles/r2/ablation.md has the numbers.
#ifndef S2D_ABLATE
#define S2D_ABLATE 0
#endif
// Compact coefficient mode, how the planes of MAT_ELAST_init_a enter.  0: the six planes are formed per GLL point
// with the reference's sequence of roundings, a_k = -w*((c*m1)*m2) (mat_elastic.f90:334-340,355-357), then the
// pointwise stage of ELAST_KD*_PSV: 26 FP64 operations per point, bitwise equal to the stored-plane mode.
// 1 (default): the constant metric factors of the flat grid (DxiDx, DetaDz) are folded into the derivative
// matrices of the four contractions, so the point only sees lambda, mu and -w: 14 operations per point (a sixth of
// the kernel's FP64 instructions less); equal to the stored-plane mode to rounding (a few ulp), not bit for bit.
#ifndef S2D_COMPACT_FOLD
#define S2D_COMPACT_FOLD 1
#endif
// The coefficient block of an element row is one contiguous run of the strip layout: with all six planes
// stored (7200 B per row of a P-SV strip) it is brought in by ONE TMA bulk copy per warp and row
// (cp.async.bulk + mbarrier, SASS UBLKCP) instead of 15 per-lane LDGSTS.128, which cost 15.5 L1 wavefronts
// each (ncu).  Measured on B200, 4096^2 FP64, ms per launch, LDGSTS -> TMA: fused full planes 8.26 -> 6.99
// (73 % -> 86 % of the HBM roofline), plain force evaluation 5.61 -> 5.09 (78 % -> 86 %); with the compact
// (lambda, mu) block (2400 B per row) 5.80 -> 5.93, so that mode keeps the per-lane copies.
// 0: never, 1 (default): full planes only, 2: both modes.
#ifndef S2D_STRIP_TMA
#define S2D_STRIP_TMA 1
#endif
// Strips per CTA.  Measured on B200 (4096^2, FP64, compact, fused; ms per launch): 4 warps at 168
// registers, 3 CTAs/SM: 5.82;  5 warps (128 registers, 3 CTAs/SM): 6.02;  6 warps (168, 2 CTAs): 6.41;
// 8 warps (128, 2 CTAs): 6.58 -- registers (instruction-level parallelism) beat resident warps here.
constexpr int strip_warps() { return S2D_STRIP_WARPS; }
// element rows per band: S2D_SEG if set, else 64 when that still leaves every SM a dozen waves of CTAs, else 32.
// Measured with the tensor-map kernel (FP64 compact fused, ms per step): 4096^2 5.92 (32) / 5.80 (48) / 5.78 (64) /
// 5.98 (128); 8192^2 24.74-24.95 (32) / 24.37-24.40 (64) / 24.27 (96) / 24.33 (128).
inline int strip_default_seg(int N, int nx, int nz) {
  const char* v = std::getenv("S2D_SEG");
  if (v && *v) return std::atoi(v);
  const long long groups = ((long long)nx + (32 / N) * strip_warps() - 1) / ((32 / N) * strip_warps());
  return ((long long)((nz + 63) / 64) * groups >= 12LL * 148 * 3) ? 64 : 32;
}
// bytes of the staging area of one warp: coefficient vectors and displacement rows of one element
// row; in the fused form also the velocities and inverse masses of the nodes it will advance
// (fused: 0 plain force evaluation, 1 leapfrog update, 2 explicit Newmark update: also the old accelerations)
// plast: also the plastic strain of the row's elements (3 N values per lane) when S2D_PLAST_STAGE
#ifndef S2D_PLAST_STAGE
#define S2D_PLAST_STAGE 1
#endif
constexpr size_t strip_stage_bytes(int N, int NDOF, int tsize, int fused, bool compact, bool plast = false) {
  const size_t npl = compact ? 2 : (NDOF == 1 ? 2 : 6);
  const size_t c = (npl / 2) * N * 32 * (2 * tsize), u = (size_t)NDOF * (N - 1) * 32 * tsize, r = (size_t)(N - 1) * 32 * tsize;
  return c + u + (fused ? u + r : 0) + (fused == 2 ? u : 0) + (plast && S2D_PLAST_STAGE ? (size_t)3 * N * 32 * tsize : 0);
}
// TENS variant (S2D_STRIP_TENSOR): the displacement rows, velocities and inverse masses of an element row arrive
// CTA-wide by THREE tensor-map TMA copies (cp.async.bulk.tensor, SASS UTMALDG) instead of 20 per-lane LDGSTS per
// warp: a box of strip_box_width() lattice columns (the GW strips of the group + the shared column, rounded up to
// a 16-byte multiple) by N-1 rows by ndof components.  Two stages; only the coefficient vectors keep the per-lane
// copies (their block is per strip).
#ifndef S2D_STRIP_TENSOR
#define S2D_STRIP_TENSOR 1
#endif
// 1: the coefficient blocks of the group's strips (one contiguous run per element row, 9.6 kB in the compact mode)
// also arrive CTA-wide, by one bulk copy on the displacement box's mbarrier.  Measured on B200, 4096^2 FP64 compact
// fused, ms per launch: per-lane LDGSTS everywhere 5.84-5.94, boxes for d / v / rmass + per-lane coefficients 5.70-5.75,
// boxes + CTA-wide coefficient copy 6.18 -- so 0 is the default.
// largest ngll with a tensor-map instantiation
#ifndef S2D_STRIP_TENSOR_MAXN
#define S2D_STRIP_TENSOR_MAXN 10  // NGLL 9 (Lamb x24): 1.169 -> 1.090 ms per step with the boxes
#endif
#ifndef S2D_STRIP_TENSOR_COEF
#define S2D_STRIP_TENSOR_COEF 0
#endif
constexpr int strip_box_width(int N, int tsize) {
  const int w = strip_warps() * (32 / N) * (N - 1) + 1, q = 16 / tsize;
  return (w + q - 1) / q * q;
}
constexpr size_t align128(size_t x) { return (x + 127) & ~(size_t)127; }
constexpr size_t strip_tens_a_bytes(int N, int NDOF, int tsize) { return align128((size_t)NDOF * (N - 1) * strip_box_width(N, tsize) * tsize); }
constexpr size_t strip_tens_r_bytes(int N, int tsize) { return align128((size_t)(N - 1) * strip_box_width(N, tsize) * tsize); }
constexpr size_t strip_tens_c_bytes(int N, int NDOF, int tsize, bool compact) {  // the group's coefficient blocks of one row
  const size_t npl = compact ? 2 : (NDOF == 1 ? 2 : 6);
  return S2D_STRIP_TENSOR_COEF ? align128((size_t)strip_warps() * (32 / N) * (npl / 2) * N * N * (2 * tsize)) : 0;
}
// dynamic shared memory of the TENS variant: [per-warp coefficient staging unless they come CTA-wide], then
// 2 x (d box | v box | rmass box [| coefficient blocks of the group])
constexpr size_t strip_tens_smem(int N, int NDOF, int tsize, bool compact, int fused = 1) {
  const size_t npl = compact ? 2 : (NDOF == 1 ? 2 : 6);
  const size_t c = S2D_STRIP_TENSOR_COEF ? 0 : align128((size_t)strip_warps() * (npl / 2) * N * 32 * (2 * tsize));
  return c + 2 * ((fused == 2 ? 3 : 2) * strip_tens_a_bytes(N, NDOF, tsize) + strip_tens_r_bytes(N, tsize) +
                  strip_tens_c_bytes(N, NDOF, tsize, compact));
}
// cp.async.bulk.tensor global -> shared (tile mode), completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

// mbarrier wait that gives up (trap -> launch failure reported to the host) instead of spinning for ever if a copy
// never lands, e.g. a malformed tensor map: a hung GPU is worse than a failed call
__device__ __forceinline__ void mbar_wait_or_trap(unsigned long long* bar, unsigned parity) {
#ifndef S2D_MBAR_HINT_NS
#define S2D_MBAR_HINT_NS 20000
#endif
  unsigned ok;
  for (unsigned spins = 0;; ++spins) {
    // the suspend-time hint lets the warp sleep in hardware until the phase completes instead of polling
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.b32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)S2D_MBAR_HINT_NS)
        : "memory");
    if (ok) return;
    if (spins > (1u << 22)) __trap();
  }
}

constexpr int strip_min_ctas(int N, int tsize, bool compact = false) {
  return (N <= 6 ? (tsize == 4 ? 4 : 3) : (tsize == 4 ? 2 : 1)) * 4 / strip_warps();
}

// Kelvin-Voigt instantiations carry the velocities of the element as well: the FP64 P-SV ones spill at 168 registers
// (100-320 B), so they run 2 CTAs per SM with the full register file (S2D_KV_MINB to measure the other choice)
#ifndef S2D_KV_MINB
#define S2D_KV_MINB 2
#endif
constexpr int strip_min_ctas_kv(int N, int tsize, int ndof) {
  return (tsize == 8 && ndof == 2 && N >= 5 && N <= 6) ? S2D_KV_MINB : strip_min_ctas(N, tsize);
}

template <typename T, int N, int NDOF, int FUSED, bool COMPACT, int MINB = strip_min_ctas(N, sizeof(T), COMPACT), bool KV = false,
          bool TENS = false, int RHEO = 0>
__global__ void __launch_bounds__(strip_warps() * 32, MINB)
    k_elem_strip(const __grid_constant__ StripArgs<T, N> A) {
  // RHEO: stateful rheology of MAT_Fint's strain -> stress -> force branch: 1 Coulomb plasticity, 2 visco-elasticity,
  // 3 damage (one instantiation each: a run-time switch between them cost the plastic kernel 200 B of spills and 20 %)
  constexpr bool PLAST = RHEO != 0;
  static_assert(!PLAST || (COMPACT && NDOF == 2 && !TENS && S2D_COMPACT_FOLD != 0),
                "stateful rheologies: P-SV, (lambda, mu) coefficient stream, folded metric");
  static_assert(!TENS || !KV, "tensor-map staging: not with Kelvin-Voigt elements");
  static_assert(!COMPACT || NDOF == 2, "compact coefficients: P-SV only");
  static_assert(!KV || !TENS, "Kelvin-Voigt elements: per-lane staging only");
  constexpr int WARPS = strip_warps();
  constexpr int NEL = NDOF == 1 ? 2 : 6;
  constexpr int NPL = COMPACT ? 2 : NEL;  // planes stored per GLL point
  constexpr int KD2 = (N == 5) ? 1 : 0;  // OPT_NGLL (constants.f90:6, mat_elastic.f90:412)
  constexpr int EPW = 32 / N;
  constexpr int NP = N;                  // tile rows: N scalars, read back one broadcast LDS per value
  constexpr unsigned FULL = 0xffffffffu;
  using V2 = typename Vec2<T>::type;
  __shared__ __align__(16) T tile[WARPS][NDOF][N][EPW * NP];
  __shared__ T hand[2][WARPS][NDOF][N];  // right-edge column of a strip, handed to the strip on its right
  __shared__ unsigned rowmask[STRIP_MASK_WORDS + 1];  // bit r: lattice row (band's first row + r) is deferred
  extern __shared__ __align__(128) unsigned char stage_raw[];
  constexpr int NU = NDOF * (N - 1);
  constexpr size_t SZ_C = (size_t)(NPL / 2) * N * 32 * sizeof(V2), SZ_U = (size_t)NU * 32 * sizeof(T);
  const StripGeom& G = A.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // every lane only ever touches its own slots of the staging area: no barrier guards it
  unsigned char* wstage = stage_raw + (size_t)warp * (TENS ? SZ_C : strip_stage_bytes(N, NDOF, sizeof(T), FUSED, COMPACT, RHEO == 1));
  // TENS: CTA-wide boxes behind the per-warp coefficient staging: stage s at tbase + s * TSZ =
  // [d box | v box | (Newmark: a box) | rmass box | (coefficient blocks)]
  constexpr int BW = strip_box_width(N, sizeof(T));
  constexpr size_t TSZ_A = strip_tens_a_bytes(N, NDOF, sizeof(T)), TSZ_R = strip_tens_r_bytes(N, sizeof(T));
  constexpr size_t TSZ_C = strip_tens_c_bytes(N, NDOF, sizeof(T), COMPACT);
  constexpr size_t TOFF_R = (FUSED == 2 ? 3 : 2) * TSZ_A;   // rmass box behind d, v (, a)
  constexpr size_t TSZ = TOFF_R + TSZ_R + TSZ_C;
  constexpr bool TCOEF = TENS && S2D_STRIP_TENSOR_COEF != 0;
  unsigned char* tbase = stage_raw + (TCOEF ? 0 : align128((size_t)WARPS * SZ_C));
  __shared__ __align__(8) unsigned long long tbarA[2], tbarB[2];
  V2* st_c = reinterpret_cast<V2*>(wstage) + lane;             // [plane pair * N + j][32]
  T* st_u = reinterpret_cast<T*>(wstage + SZ_C) + lane;        // [c * (N-1) + j-1][32]
  T* st_v = reinterpret_cast<T*>(wstage + SZ_C + SZ_U) + lane; // [c * (N-1) + j][32]   (fused)
  T* st_r = st_v + NU * 32;                                    // [j][32]               (fused)
  T* st_a = st_r + (N - 1) * 32;                               // [c * (N-1) + j][32]   (fused Newmark: a[n-1])
  T* st_e = reinterpret_cast<T*>(wstage + strip_stage_bytes(N, NDOF, sizeof(T), FUSED, COMPACT, false)) + lane;  // [k * N + j][32] (plasticity)
  constexpr bool NM = FUSED == 2;
  // Kelvin-Voigt inside the fused step: every lane reads the velocities of its element's nodes for d + eta*v, so the
  // update must not overwrite them in place -- v (and, for Newmark, a) are double-buffered (v_in != v_out) and the
  // lane that stores a deferred node's force always leaves that node's (predicted) velocity in v_out
  constexpr bool VDB = KV && FUSED != 0;
  auto kv_vel = [&](size_t q) -> T {  // the velocity MAT_KV_add_etav sees (solver.f90:293-295)
    if constexpr (FUSED != 0) {
      T x = A.v_in[q];
      if (NM) x = x + A.c2 * A.a_in[q];
      return x;
    } else {
      return A.v_kv[q];
    }
  };
  constexpr bool tma_c = (S2D_STRIP_TMA == 2 || (S2D_STRIP_TMA == 1 && !COMPACT)) && ((S2D_STRIP_STAGE & 2) != 0) &&
                         sizeof(T) == 8;  // 16-byte vectors: block address and size are multiples of 16
  __shared__ __align__(8) unsigned long long cbar[WARPS];
  V2* st_cb = reinterpret_cast<V2*>(wstage);                   // TMA: the block as it lies in HBM, [pp*N+j][cx*N]
  unsigned cphase = 0;
  const long long cta = blockIdx.x;
  const int seg = (int)(cta / G.it_ng);
  const int grp = G.it_g0 + (int)(cta - (long long)seg * G.it_ng) * G.it_step;
  int gfirst, gcount;
  strip_group(G, grp, gfirst, gcount);
  const bool wact = warp < gcount;            // warps beyond the group only keep the barriers company
  const int strip = gfirst + (wact ? warp : 0);
  int ez0, ez1;
  strip_seg_rows(G, seg, ez0, ez1);
  const int ex0 = strip * EPW;
  const int cx = min(EPW, G.nx - ex0);
  int el = lane / N;
  const int i = lane - el * N;
  const bool real = wact && el < cx;
  if (el >= cx) el = cx - 1;  // shadow lanes mirror the last element (same cache lines), never store
  const int lanep = el * N + i;
  const bool dup = (i == N - 1) && (el < cx - 1);     // column owned by lane+1 (i = 0 of the next element)
  const bool redge = (i == N - 1) && (el == cx - 1);  // strip's right edge
  const bool merge = real && (i == 0) && (el > 0);
  const bool give = real && redge && (warp < gcount - 1);            // handed to the next warp of the group
  const bool take = real && (lane == 0) && (warp > 0);               // receives the previous warp's edge
  const bool to_halo = redge && !give && (strip < G.nstrips - 1);    // group's right edge: partial sum to halo_x
  const bool st_ok = real && !dup && !give;
  const size_t LX = (size_t)G.LXP;  // row pitch
  const int gx = (ex0 + el) * (N - 1) + i;
  const T* up = A.d + gx;
  T* sp;
  size_t rstride, cstride;
  if (to_halo) {
    sp = A.halo_x + (size_t)grp * G.LZ;
    rstride = 1;
    cstride = (size_t)(G.ngroups - 1) * G.LZ;
  } else {
    sp = A.f + gx;
    rstride = LX;
    cstride = A.npoin;
  }
  bool coldef = true;
  const int band_row0 = strip_lat_row(G, ez0, 0);
  if (FUSED) {
    coldef = to_halo || (A.colflag[gx] != 0);
    const int nrows_band = strip_lat_row(G, ez1 - 1, N - 1) - band_row0 + 1;
    for (int w = warp; w <= STRIP_MASK_WORDS; w += WARPS) {
      const int r = 32 * w + lane;
      const unsigned m = __ballot_sync(FULL, r < nrows_band && A.rowflag[band_row0 + r] != 0);
      if (lane == 0) rowmask[w] = m;
      if (32 * w >= nrows_band) break;  // one word past the band's last one (zeros) is enough
    }
    __syncthreads();
  }
  T(*tlw)[N][EPW * NP] = tile[warp];
  // xi-contractions read the other lanes' columns from the warp tile.  S2D_STRIP_ROT: a lane's own column is
  // already in its registers, so it reads only the other N-1, in a per-lane rotation m_k = (i + k) mod N
  // (k = 1..N-1): a fifth fewer shared-memory wavefronts, the contraction's sum starts with the own term.
  constexpr bool ROT = S2D_STRIP_ROT != 0;
  constexpr int NR = ROT ? N - 1 : N;
  constexpr bool FOLD = COMPACT && (S2D_COMPACT_FOLD != 0);
  const T* const HZ = FOLD ? A.Hz : A.H;   // operand of the eta-contractions (constant bank)
  const T sx = FOLD ? A.cdx : (T)1;        // xi-contractions: the lane's hprime column / row carries DxiDx
  T Hi[NR], HTi[NR], Hii = 0, HTii = 0;
  int mo[ROT ? N - 1 : 1];
  if (ROT) {
    Hii = A.H[i + N * i] * sx;
    HTii = Hii;
#pragma unroll
    for (int k = 1; k < N; ++k) {
      const int m = (i + k) % N;
      Hi[k - 1] = A.H[m + N * i] * sx;   // H(m,i)
      HTi[k - 1] = A.H[i + N * m] * sx;  // H(i,m)
      mo[k - 1] = el * N + m;
    }
  } else {
#pragma unroll
    for (int m = 0; m < NR; ++m) {
      Hi[m] = A.H[m + N * i] * sx;   // H(m,i)
      HTi[m] = A.H[i + N * m] * sx;  // H(i,m)
    }
  }
  T U[NDOF][N], Fc[NDOF];
  T Vc[KV ? NDOF : 1];  // KV: velocity of the row shared with the next element row, carried like U[c][0]
  {
    const size_t r0 = (size_t)strip_lat_row(G, ez0, 0) * LX;
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      U[c][0] = up[A.npoin * c + r0];
      Fc[c] = 0;
      if (KV) Vc[c] = kv_vel(gx + A.npoin * c + r0);
    }
  }
  const T* etap = KV ? A.eta + (size_t)strip_elem_off(G, seg, strip, ez0) * (N * N) + lanep : nullptr;
  T* epp = PLAST ? A.pl_ep + (size_t)strip_elem_off(G, seg, strip, ez0) * (3 * N * N) + lanep : nullptr;
  const unsigned char* plp = PLAST ? A.pl_set + strip_elem_off(G, seg, strip, ez0) + el : nullptr;
  int pset_next = (PLAST && wact) ? (int)*plp : 0;
  T* dmp = (RHEO == 3) ? A.dm_state + (size_t)strip_elem_off(G, seg, strip, ez0) * 4 * (N * N) + lanep : nullptr;
  T* vsp = (RHEO == 2) ? A.vs_state + (size_t)strip_elem_off(G, seg, strip, ez0) * (3 * (A.vs_nb + 1)) * (N * N) + lanep : nullptr;
  const int cxN = cx * N;
  const V2* cp = reinterpret_cast<const V2*>(A.coef) +
                 (size_t)strip_elem_off(G, seg, strip, ez0) * (NPL * N * N / 2) + lanep;
  int gex0, gcx;
  strip_group_cols(G, strip, gex0, gcx);
  const size_t cp_blk = (size_t)cx * (NPL * N * N / 2);    // this strip's block of one element row
  const size_t cp_row = (size_t)gcx * (NPL * N * N / 2);   // from one element row to the next (the group's blocks)
  T nW[COMPACT ? N : 1];  // -weights(i,j) of this lane's column
  if (COMPACT) {
#pragma unroll
    for (int j = 0; j < N; ++j) nW[j] = -mul_rn(A.cdet, mul_rn(A.wg[i], A.wg[j]));
  }

  // loads of one element row: displacement rows j = 1..N-1 and the coefficient vectors.
  // S2D_STRIP_STAGE (compile time) selects what goes through the staging area (bit 0:
  // displacements, 1: coefficients, 2: v and rmass of the fused update); the rest is loaded straight
  // into registers when it is needed.  Measured on B200, 4096^2 FP64 compact: 7 -> 6.45 ms,
  // 5 -> 6.59, 4 -> 6.83, 1 -> 6.88, 0 -> 7.11 ms per launch.
  constexpr bool stg_u = S2D_STRIP_STAGE & 1, stg_c = S2D_STRIP_STAGE & 2, stg_v = S2D_STRIP_STAGE & 4;
  auto issue_row = [&](int ezr, const V2* cpr) {
    const size_t rb = (size_t)strip_lat_row(G, ezr, 0) * LX;
    if (stg_u && !TENS && !(S2D_ABLATE & 4)) {
#pragma unroll
      for (int c = 0; c < NDOF; ++c)
#pragma unroll
        for (int j = 1; j < N; ++j)
          stage_copy<sizeof(T)>(st_u + (c * (N - 1) + j - 1) * 32, up + A.npoin * c + rb + (size_t)j * LX);
    }
    if constexpr (VDB) {  // the row's velocities (Newmark: and accelerations) travel with its displacements
      if (stg_u && stg_v) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 1; j < N; ++j) {
            const size_t q = gx + A.npoin * c + rb + (size_t)j * LX;
            stage_copy<sizeof(T)>(st_v + (c * (N - 1) + j - 1) * 32, A.v_in + q);
            if (NM) stage_copy<sizeof(T)>(st_a + (c * (N - 1) + j - 1) * 32, A.a_in + q);
          }
      }
    }
    if constexpr (RHEO == 1 && S2D_PLAST_STAGE != 0) {  // the plastic strain of the row's elements travels with its displacements
      const T* en = A.pl_ep + (size_t)strip_elem_off(G, seg, strip, ezr) * (3 * N * N) + lanep;
#pragma unroll
      for (int k = 0; k < 3 * N; ++k) stage_copy<sizeof(T)>(st_e + k * 32, en + (size_t)k * cxN);
    }
    (void)rb;
    if (TCOEF || (S2D_ABLATE & 2)) {
    } else if (tma_c) {
      if (lane == 0) {
        const unsigned bytes = (unsigned)(cp_blk * sizeof(V2));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warp's reads of the block come first
        mbar_expect_tx(&cbar[warp], bytes);
        bulk_g2s(st_cb, cpr - lanep, bytes, &cbar[warp]);
      }
    } else if (stg_c) {
#pragma unroll
      for (int pp = 0; pp < NPL / 2; ++pp)
#pragma unroll
        for (int j = 0; j < N; ++j)
          stage_copy<sizeof(V2)>(st_c + (pp * N + j) * 32, cpr + (size_t)(pp * N + j) * cxN);
    } else if (A.prefetch) {  // pull the row's coefficient block into L2
      const char* nb = reinterpret_cast<const char*>(cpr - lanep);
      const int nlines = (int)((cp_blk * sizeof(V2) + 127) / 128);
      for (int l = lane; l < nlines; l += 32) l2_prefetch_line(nb + (size_t)l * 128);
    }
  };
  if (tma_c && wact) {
    if (lane == 0) mbar_init(&cbar[warp], 1);
    __syncwarp();
  }
  // TENS: box origin of this group and the two issue helpers (one elected thread of the CTA)
  const int bx0 = gfirst * G.W;                 // first lattice column of the group
  const int bx = gx - bx0;                      // this lane's column inside the boxes
  auto tens_issue_A = [&](int ezr) {            // displacement rows 1..N-1 of element row ezr -> stage (ezr - ez0) & 1
    const int sg = (ezr - ez0) & 1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // ... and (S2D_STRIP_TENSOR_COEF) the coefficient blocks of the group's strips for that row: one contiguous run
    int g0, gc;
    strip_group_cols(G, gfirst, g0, gc);
    const unsigned cbytes = TCOEF ? (unsigned)((size_t)gc * (NPL * N * N / 2) * sizeof(V2)) : 0u;
    mbar_expect_tx(&tbarA[sg], (unsigned)((size_t)NDOF * (N - 1) * BW * sizeof(T)) + cbytes);
    tma_load_3d(tbase + sg * TSZ, &A.tm_d, bx0, strip_lat_row(G, ezr, 0) + 1, 0, &tbarA[sg]);
    if (TCOEF)
      bulk_g2s(tbase + sg * TSZ + TOFF_R + TSZ_R,
               reinterpret_cast<const V2*>(A.coef) + (size_t)strip_elem_off(G, seg, gfirst, ezr) * (NPL * N * N / 2), cbytes,
               &tbarA[sg]);
  };
  auto tens_issue_B = [&](int ezr) {            // v and rmass of rows 0..N-2 of element row ezr
    const int sg = (ezr - ez0) & 1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&tbarB[sg], (unsigned)((size_t)((FUSED == 2 ? 2 : 1) * NDOF + 1) * (N - 1) * BW * sizeof(T)));
    tma_load_3d(tbase + sg * TSZ + TSZ_A, &A.tm_v, bx0, strip_lat_row(G, ezr, 0), 0, &tbarB[sg]);
    if (FUSED == 2) tma_load_3d(tbase + sg * TSZ + 2 * TSZ_A, &A.tm_a, bx0, strip_lat_row(G, ezr, 0), 0, &tbarB[sg]);
    tma_load_2d(tbase + sg * TSZ + TOFF_R, &A.tm_r, bx0, strip_lat_row(G, ezr, 0), &tbarB[sg]);
  };
  if constexpr (TENS) {
    if (threadIdx.x == 0) {
      mbar_init(&tbarA[0], 1);
      mbar_init(&tbarA[1], 1);
      mbar_init(&tbarB[0], 1);
      mbar_init(&tbarB[1], 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tens_issue_A(ez0);
      if (ez0 + 1 < ez1) tens_issue_A(ez0 + 1);
      if constexpr (FUSED != 0) tens_issue_B(ez0);
    }
    __syncthreads();  // the barriers exist before anybody waits on them
  }
  if (wact) {
    issue_row(ez0, cp);
    stage_commit();
  }

  for (int ez = ez0; ez < ez1; ++ez, cp += cp_row) {
    const size_t grow = (size_t)strip_lat_row(G, ez, 0);
    const size_t rowbase = grow * LX;
    T f[NDOF][N];
    unsigned defer = 0;  // bit j: the node of row grow+j is deferred (its force is stored instead)
    T vk[VDB ? NDOF : 1][VDB ? N - 1 : 1];  // fused Kelvin-Voigt: velocities of rows 0..N-2, reused by the node update
    if (wact) {
      // ---- this element row has landed in the staging area: move it to registers, then start the
      // copies of what the end of this iteration needs (v, rmass) and of the whole next row
      stage_wait<0>();
      V2 a2[NPL / 2][N];
      // plasticity: the element's plastic strain and material set, requested before anything else of this row
      T epr[PLAST ? 3 : 1][PLAST ? N : 1], ppar[PLAST ? 6 : 1];
      bool yielded = false;
      int vset = 0;
      if constexpr (PLAST) {
        // shadow lanes (el >= cx) mirror the last element -- they write the same tile slots, so they must see
        // the same state; only real lanes store it back
        const int pset = pset_next;           // read one row ahead: the parameter loads below depend on it
        plp += gcx;
        if (ez + 1 < ez1) pset_next = (int)*plp;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < N; ++j)
            epr[k][j] = RHEO != 1 ? (T)0 : (S2D_PLAST_STAGE ? st_e[(k * N + j) * 32] : __ldcs(epp + (size_t)(k * N + j) * cxN));
#pragma unroll
        for (int q = 0; q < 6; ++q) ppar[q] = RHEO != 1 ? (T)0 : __ldg(A.pl_tab + pset * 6 + q);
        vset = pset;
      }
      if constexpr (TENS) {
        const int kk = ez - ez0;
        mbar_wait_or_trap(&tbarA[kk & 1], (unsigned)((kk >> 1) & 1));
        const T* bA = reinterpret_cast<const T*>(tbase + (kk & 1) * TSZ) + bx;
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 1; j < N; ++j) U[c][j] = bA[(c * (N - 1) + j - 1) * BW];
      } else if (S2D_ABLATE & 4) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 1; j < N; ++j) U[c][j] = U[c][0] + (T)j;
      } else if (stg_u) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 1; j < N; ++j) U[c][j] = st_u[(c * (N - 1) + j - 1) * 32];
      } else {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 1; j < N; ++j) U[c][j] = up[A.npoin * c + rowbase + (size_t)j * LX];
      }
      T Vn[VDB ? NDOF : 1][VDB ? N - 1 : 1];  // fused Kelvin-Voigt: (predicted) velocities of rows 1..N-1
      if constexpr (VDB) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 1; j < N; ++j) {
            if (stg_u && stg_v) {
              Vn[c][j - 1] = st_v[(c * (N - 1) + j - 1) * 32];
              if (NM) Vn[c][j - 1] = Vn[c][j - 1] + A.c2 * st_a[(c * (N - 1) + j - 1) * 32];
            } else {
              Vn[c][j - 1] = kv_vel(gx + A.npoin * c + rowbase + (size_t)j * LX);
            }
          }
      }
      if constexpr (TCOEF) {  // landed with the displacement box (same mbarrier, waited for above)
        const V2* cb = reinterpret_cast<const V2*>(tbase + ((ez - ez0) & 1) * TSZ + TOFF_R + TSZ_R) +
                       (size_t)(ex0 - gex0) * (NPL * N * N / 2) + lanep;
#pragma unroll
        for (int pp = 0; pp < NPL / 2; ++pp)
#pragma unroll
          for (int j = 0; j < N; ++j) a2[pp][j] = cb[(pp * N + j) * cxN];
      } else if (S2D_ABLATE & 2) {
#pragma unroll
        for (int pp = 0; pp < NPL / 2; ++pp)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            a2[pp][j].x = U[0][0] + (T)(pp + j);
            a2[pp][j].y = (T)3e10 + U[0][0];
          }
      } else if (tma_c) {
        mbar_wait(&cbar[warp], cphase);
        cphase ^= 1u;
#pragma unroll
        for (int pp = 0; pp < NPL / 2; ++pp)
#pragma unroll
          for (int j = 0; j < N; ++j) a2[pp][j] = st_cb[(pp * N + j) * cxN + lanep];
        __syncwarp();  // every lane has its vectors in registers before the next row's copy may land
      } else if (stg_c) {
#pragma unroll
        for (int pp = 0; pp < NPL / 2; ++pp)
#pragma unroll
          for (int j = 0; j < N; ++j) a2[pp][j] = st_c[(pp * N + j) * 32];
      } else {
#pragma unroll
        for (int pp = 0; pp < NPL / 2; ++pp)
#pragma unroll
          for (int j = 0; j < N; ++j) a2[pp][j] = ld_stream(cp + (size_t)(pp * N + j) * cxN);
      }
      if (FUSED) {
        {
          const int o = (int)grow - band_row0;
          const unsigned long long two = ((unsigned long long)rowmask[(o >> 5) + 1] << 32) | rowmask[o >> 5];
          defer = coldef ? ~0u : (unsigned)(two >> (o & 31));
        }
        if (st_ok && !TENS && !(S2D_ABLATE & 8)) {
          if (stg_v) {
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
              const size_t q = rowbase + (size_t)j * LX + gx;
              stage_copy<sizeof(T)>(st_r + j * 32, A.rmass + q);
#pragma unroll
              if (!VDB) {  // Kelvin-Voigt: the (predicted) velocities are already in registers (vk)
#pragma unroll
                for (int c = 0; c < NDOF; ++c) {
                  stage_copy<sizeof(T)>(st_v + (c * (N - 1) + j) * 32, A.v_in + A.npoin * c + q);
                  if (NM) stage_copy<sizeof(T)>(st_a + (c * (N - 1) + j) * 32, A.a_in + A.npoin * c + q);
                }
              }
            }
          } else if (A.prefetch && i == 0) {  // read at the end of the row: pull the lines into L2 now
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
              const size_t q = rowbase + (size_t)j * LX + gx;
              l2_prefetch_line(A.rmass + q);
#pragma unroll
              for (int c = 0; c < NDOF; ++c) l2_prefetch_line(A.v_in + A.npoin * c + q);
            }
          }
        }
      }
      stage_commit();
      if (ez + 1 < ez1) issue_row(ez + 1, cp + cp_row);
      stage_commit();
      // ---- gradients: xi through the warp tile, eta in registers.
      // Measured alternative (B200, 4096^2, FP64, ms per launch, fused compact / fused full / plain
      // full): this tile 5.82 / 8.26 / 5.61;  __shfl_sync rotations among the N lanes of an element
      // instead of the tile (no shared memory, no __syncwarp) 6.24 / 7.72 / 6.08.  The tile costs 2
      // L1 wavefronts per broadcast LDS.64 (ncu), the shuffles cost issue slots of the same MIO queue.
      // Kelvin-Voigt: the element sees d + eta*v (mat_kelvin_voigt.f90:147); eta belongs to the element, so the
      // row shared with the next element row is combined again there, from the carried d and v
      T Ue[KV ? NDOF : 1][KV ? N : 1];
      if constexpr (KV) {
        T et[N];
#pragma unroll
        for (int j = 0; j < N; ++j) et[j] = ld_stream(etap + (size_t)j * cxN);
        etap += (size_t)gcx * (N * N);
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          T vr[N];
          vr[0] = Vc[c];
#pragma unroll
          for (int j = 1; j < N; ++j) {
            if constexpr (VDB) vr[j] = Vn[c][j - 1];
            else vr[j] = kv_vel(gx + A.npoin * c + rowbase + (size_t)j * LX);
          }
#pragma unroll
          for (int j = 0; j < N; ++j) Ue[c][j] = U[c][j] + et[j] * vr[j];
          if constexpr (VDB) {
#pragma unroll
            for (int j = 0; j < N - 1; ++j) vk[c][j] = vr[j];
          }
          Vc[c] = vr[N - 1];
        }
      }
      auto& Ut = [&]() -> T(&)[NDOF][N] {
        if constexpr (KV) return Ue;
        else return U;
      }();
      if (!(S2D_ABLATE & 1)) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 0; j < N; ++j) tlw[c][j][lanep] = Ut[c][j];
      }
      __syncwarp();
      T gxi[NDOF][N], get[NDOF][N];
#pragma unroll
      for (int c = 0; c < NDOF; ++c)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          T row[NR];  // the 5 lanes of an element read the same word: one wavefront per value
#pragma unroll
          for (int m = 0; m < NR; ++m) row[m] = (S2D_ABLATE & 1) ? Ut[c][(j + m) % N] : tlw[c][j][ROT ? mo[m] : el * N + m];
          T s1 = ROT ? Hii * Ut[c][j] : (T)0, s2 = 0;
#pragma unroll
          for (int m = 0; m < NR; ++m) s1 += Hi[m] * row[m];                   // (Ht U)(i,j)
#pragma unroll
          for (int m = 0; m < N; ++m) s2 += Ut[c][m] * HZ[m + N * j];          // (U H)(i,j)
          gxi[c][j] = s1;
          get[c][j] = s2;
        }
      __syncwarp();
      // ---- pointwise stage (mat_elastic.f90:600-619 / :484-496 / :751-762)
      T tH[NDOF][N], tHt[NDOF][N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if constexpr (RHEO == 3) {
          // damage rheology: moduli degraded by alpha (mat_damage.f90:372-378), stress and strain invariants of the
          // elastic strain e0 + e - ep (compute_stress), damage growth dalpha = dt Cd i2 [xi alpha^beta - xi_0]+ and
          // the damage-related plastic strain increment Cv dalpha (s - s_mean) (:393-412), relative stress out
          const T* tb = A.dm_tab + vset * STRIP_DM_TAB;
          const T e1t = gxi[0][j], e2t = get[1][j], e3t = T(0.5) * (get[0][j] + gxi[1][j]);
          if (vset == 0) {  // elastic element of a damage problem
            const T la = a2[0][j].x, two_mu = T(2) * a2[0][j].y;
            const T s1 = (la + two_mu) * e1t + la * e2t, s2 = la * e1t + (la + two_mu) * e2t, s3 = two_mu * e3t;
            tH[0][j] = nW[j] * s1;
            tHt[0][j] = nW[j] * s3;
            tH[1][j] = nW[j] * s3;
            tHt[1][j] = nW[j] * s2;
            continue;
          }
          const T rl = __ldg(tb), mu0 = __ldg(tb + 1), xi0 = __ldg(tb + 2), gr = __ldg(tb + 3), beta = __ldg(tb + 4),
                  Cd = __ldg(tb + 5), Cv = __ldg(tb + 6), dtl = __ldg(tb + 13);
          const size_t pstr = (size_t)N * cxN;
          T* sp = dmp + (size_t)j * cxN;
          T al = sp[0], p1 = sp[pstr], p2 = sp[2 * pstr], p3 = sp[3 * pstr];
          const T e1 = (e1t + __ldg(tb + 7)) - p1, e2 = (e2t + __ldg(tb + 8)) - p2, e3 = (e3t + __ldg(tb + 9)) - p3;
          const T rm = mu0 + xi0 * gr * al;
          const T rg = beta == T(0) ? gr * al : gr * pow(al, T(1) + beta) / (T(1) + beta);
          const T i1 = e1 + e2, i2 = e1 * e1 + e2 * e2 + T(2) * e3 * e3;
          const T si2 = sqrt(i2);
          const T xi = si2 < T(1e-10) ? T(0) : i1 / si2;
          const T two_mue = T(2) * rm - rg * xi;
          T s1 = rl * i1 - rg * si2 + two_mue * e1;
          T s2 = rl * i1 - rg * si2 + two_mue * e2;
          T s3 = two_mue * e3;
          {
            const T pp = -(T(4) * rm + T(2) * rl - T(3) * rg * xi);
            const T qq = two_mue * two_mue + two_mue * (T(2) * rl - rg * xi) + rg * (rl * xi - rg) * (T(2) - xi * xi);
            const T dd = pp * pp / T(4) - qq;
            if (real && (dd <= T(0) || pp / T(2) + sqrt(dd) >= T(0) || two_mue <= T(0))) *A.dm_err = 4;
          }
          T dal = dtl * Cd * i2 * fmax((beta == T(0) ? xi : xi * pow(al, beta)) - xi0, T(0));
          if (real && dal > T(0)) {
            const T sm = T(0.5) * (s1 + s2);
            const T da = Cv * dal;
            sp[0] = al + dal;
            sp[pstr] = p1 + (s1 - sm) * da;
            sp[2 * pstr] = p2 + (s2 - sm) * da;
            sp[3 * pstr] = p3 + s3 * da;
          }
          s1 = s1 - __ldg(tb + 10);
          s2 = s2 - __ldg(tb + 11);
          s3 = s3 - __ldg(tb + 12);
          tH[0][j] = nW[j] * s1;
          tHt[0][j] = nW[j] * s3;
          tH[1][j] = nW[j] * s3;
          tHt[1][j] = nW[j] * s2;
          continue;
        }
        if constexpr (RHEO == 2) {
          // generalized Maxwell body: the memory variables of every mechanism relax towards the strain of the
          // PREVIOUS evaluation (4th-order expansion of 1 - exp(-w dt), mat_visco.f90:221-229), the strain is kept
          // for the next one, the anelastic stress is taken off the unrelaxed elastic one (:236-246)
          const T* tb = A.vs_tab + vset * STRIP_VS_TAB;
          const T la = vset ? __ldg(tb) : a2[0][j].x, two_mu = T(2) * (vset ? __ldg(tb + 1) : a2[0][j].y);
          const int nb = vset ? (int)__ldg(tb + 2) : 0;
          const T e1 = gxi[0][j], e2 = get[1][j], e3 = T(0.5) * (get[0][j] + gxi[1][j]);
          const size_t pstr = (size_t)N * cxN;   // one plane of the element row's block
          T* sp = vsp + (size_t)j * cxN;
          T* so = sp + (size_t)(3 * A.vs_nb) * pstr;
          const T o1 = so[0], o2 = so[pstr], o3 = so[2 * pstr];
          T sa1 = 0, sa2 = 0, sa3 = 0;
          for (int b = 0; b < nb; ++b) {
            const T rk = __ldg(tb + 3 + b), th1 = __ldg(tb + 3 + STRIP_VS_MAXB + b), th2 = __ldg(tb + 3 + 2 * STRIP_VS_MAXB + b),
                    th3 = __ldg(tb + 3 + 3 * STRIP_VS_MAXB + b);
            T* sb = sp + (size_t)(3 * b) * pstr;
            T q1 = sb[0], q2 = sb[pstr], q3 = sb[2 * pstr];
            q1 = q1 + rk * (o1 - q1);
            q2 = q2 + rk * (o2 - q2);
            q3 = q3 + rk * (o3 - q3);
            if (real) {
              sb[0] = q1;
              sb[pstr] = q2;
              sb[2 * pstr] = q3;
            }
            sa1 = sa1 + th1 * q1 + th2 * q2;
            sa2 = sa2 + th2 * q1 + th1 * q2;
            sa3 = sa3 + th3 * q3;
          }
          if (real && vset) {
            so[0] = e1;
            so[pstr] = e2;
            so[2 * pstr] = e3;
          }
          const T s1 = (la + two_mu) * e1 + la * e2 - sa1;
          const T s2 = la * e1 + (la + two_mu) * e2 - sa2;
          const T s3 = two_mu * e3 - sa3;
          tH[0][j] = nW[j] * s1;
          tHt[0][j] = nW[j] * s3;
          tH[1][j] = nW[j] * s3;
          tHt[1][j] = nW[j] * s2;
          continue;
        }
        if constexpr (RHEO == 1) {
          // MAT_strain_PSV (mat_gen.f90:752-775) on the flat box: e11 = Ux,x  e22 = Uz,z  e12 = (Ux,z + Uz,x)/2;
          // MAT_PLAST_stress with update (mat_plastic.f90:297-377): trial stress from the absolute elastic strain,
          // visco-plastic return of the deviatoric part towards the Coulomb yield stress (Andrews 2005), plastic
          // strain advanced, stress relative to the initial one; MAT_forces (mat_gen.f90:851-860) with the metric
          // factors carried by the second contractions
          const T la = a2[0][j].x, two_mu = T(2) * a2[0][j].y;
          const T px = gxi[0][j], pz = gxi[1][j], qx = get[0][j], qz = get[1][j];
          const T e1 = (px - epr[0][j]) + ppar[3];
          const T e2 = (qz - epr[1][j]) + ppar[4];
          const T e3 = (T(0.5) * (qx + pz) - epr[2][j]) + ppar[5];
          T s1 = (la + two_mu) * e1 + la * e2;
          T s2 = la * e1 + (la + two_mu) * e2;
          T s3 = two_mu * e3;
          // Y / tau as Y * rsqrt(tau^2) and the three divisions by 2 mu as one reciprocal: one special-function
          // sequence each instead of a square root and four divisions (21.9 -> see DESIGN.md G DOF/s); the results
          // differ from the reference's expressions by an ulp, far inside the parity tolerance
          const T tau2 = T(0.25) * ((s1 - s2) * (s1 - s2)) + s3 * s3;
          const T sm = T(0.5) * (s1 + s2);
          const T Y = ppar[0] - ppar[1] * sm;
          const T t1 = s1 - sm, t2 = s2 - sm, t3 = s3;
          const T factor = T(1) - fmax(T(1) - Y * rsqrt_any(tau2), T(0)) * ppar[2];
          const T d1 = factor * t1, d2 = factor * t2, d3 = factor * t3;
          const T i2m = rcp_any(two_mu);
          epr[0][j] = epr[0][j] + (t1 - d1) * i2m;
          epr[1][j] = epr[1][j] + (t2 - d2) * i2m;
          epr[2][j] = epr[2][j] + (t3 - d3) * i2m;
          yielded = yielded || factor < T(1);
          s1 = (d1 + sm) - ((la + two_mu) * ppar[3] + la * ppar[4]);
          s2 = (d2 + sm) - (la * ppar[3] + (la + two_mu) * ppar[4]);
          s3 = d3 - two_mu * ppar[5];
          tH[0][j] = nW[j] * s1;
          tHt[0][j] = nW[j] * s3;
          tH[1][j] = nW[j] * s3;
          tHt[1][j] = nW[j] * s2;
          continue;
        }
        if constexpr (FOLD) {
          // gxi carries DxiDx, get carries DetaDz; the second contractions carry the other factor of each plane:
          //   fx = (DxiDx H)(-w [Kx px + la qz]) + (-w mu [qx + DetaDz Uz,xi]) (DetaDz Ht)      (KD2: a4*(Ux,eta + Uz,xi),
          //   fz = (DxiDx H)(-w mu [qx + pz])    + (-w [la px + Kx qz]) (DetaDz Ht)              mat_elastic.f90:612)
          const T la = a2[0][j].x, mu = a2[0][j].y;
          const T kx = la + T(2) * mu;
          const T px = gxi[0][j], pz = gxi[NDOF - 1][j], qx = get[0][j], qz = get[NDOF - 1][j];
          const T shear = mu * (qx + pz);
          tH[0][j] = nW[j] * fma(kx, px, la * qz);
          tHt[0][j] = KD2 ? nW[j] * (mu * fma(A.rzx, pz, qx)) : nW[j] * shear;
          tH[NDOF - 1][j] = nW[j] * shear;
          tHt[NDOF - 1][j] = nW[j] * fma(la, px, kx * qz);
          continue;
        }
        T ar[NEL], g1[NDOF], g2[NDOF], o1[NDOF], o2[NDOF];
        if constexpr (COMPACT) {
          const T la = a2[0][j].x, mu = a2[0][j].y;
          const T kx = la + T(2) * mu;  // 2*mu is exact: one rounding with or without contraction
          const T kdx = mul_rn(kx, A.cdx), ldx = mul_rn(la, A.cdx), mdx = mul_rn(mu, A.cdx);
          const T kdz = mul_rn(kx, A.cdz), mdz = mul_rn(mu, A.cdz);
          ar[0] = mul_rn(nW[j], mul_rn(kdx, A.cdx));
          ar[1] = mul_rn(nW[j], mul_rn(ldx, A.cdz));
          ar[2] = mul_rn(nW[j], mul_rn(kdz, A.cdz));
          ar[3] = mul_rn(nW[j], mul_rn(mdz, A.cdz));
          ar[4] = mul_rn(nW[j], mul_rn(mdx, A.cdz));
          ar[5] = mul_rn(nW[j], mul_rn(mdx, A.cdx));
        } else {
#pragma unroll
          for (int pp = 0; pp < NPL / 2; ++pp) {
            ar[2 * pp] = a2[pp][j].x;
            ar[2 * pp + 1] = a2[pp][j].y;
          }
        }
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          g1[c] = gxi[c][j];
          g2[c] = get[c][j];
        }
        pointwise_stage<T, NDOF>(ar, NEL, KD2, g1, g2, o1, o2);
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          tH[c][j] = o1[c];
          tHt[c][j] = o2[c];
        }
      }
      if constexpr (PLAST) {
        if constexpr (RHEO == 2) vsp += (size_t)gcx * (3 * (A.vs_nb + 1)) * (N * N);
        if constexpr (RHEO == 3) dmp += (size_t)gcx * 4 * (N * N);
        if (real && yielded) {  // an element that did not yield leaves its plastic strain as it is in HBM
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < N; ++j) epp[(size_t)(k * N + j) * cxN] = epr[k][j];
        }
        epp += (size_t)gcx * (3 * N * N);
      }
      // ---- second contractions: H tH through the tile, tHt Ht in registers
      if (!(S2D_ABLATE & 1)) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 0; j < N; ++j) tlw[c][j][lanep] = tH[c][j];
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < NDOF; ++c)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          T row[NR];
#pragma unroll
          for (int m = 0; m < NR; ++m) row[m] = (S2D_ABLATE & 1) ? tH[c][(j + m) % N] : tlw[c][j][ROT ? mo[m] : el * N + m];
          T s1 = ROT ? HTii * tH[c][j] : (T)0, s2 = 0;
#pragma unroll
          for (int m = 0; m < NR; ++m) s1 += HTi[m] * row[m];                  // (H tH)(i,j)
#pragma unroll
          for (int m = 0; m < N; ++m) s2 += tHt[c][m] * HZ[j + N * m];         // (tHt Ht)(i,j)
          f[c][j] = s1 + s2;
        }
      if (A.beta != nullptr) {  // finite seismogenic width: - beta*d with the element's (KV-modified) d
        const T* bp = A.beta + (size_t)strip_elem_off(G, seg, strip, ez) * (N * N) + lanep;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const T b = ld_stream(bp + (size_t)j * cxN);
#pragma unroll
          for (int c = 0; c < NDOF; ++c) f[c][j] = f[c][j] - b * Ut[c][j];
        }
      }
      __syncwarp();
      // ---- assembly inside the strip: merge the column shared with the element to the left
#pragma unroll
      for (int c = 0; c < NDOF; ++c)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const T t = __shfl_up_sync(FULL, f[c][j], 1);
          if (merge) f[c][j] += t;
        }
      if (give) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 0; j < N; ++j) hand[ez & 1][warp][c][j] = f[c][j];
      }
    }
    // (producer / consumer named barriers between neighbouring warps instead of this CTA barrier were
    // tried with two slots and hung: two bar.arrive of a producer that runs a row ahead complete a 64-thread
    // phase on their own.  Strict alternation would work but couples the pair as tightly as this barrier.)
    if ((WARPS > 1 && !(S2D_ABLATE & 32)) || TENS) __syncthreads();
    if constexpr (TENS) {
      // every warp has read the displacement box of this row (top of the iteration) and the v / rmass boxes of the
      // previous row (end of the previous iteration): both stages are free for rows ez + 2 and ez + 1
      if (threadIdx.x == 0) {
        if (ez + 2 < ez1) tens_issue_A(ez + 2);
        if constexpr (FUSED != 0) {
          if (ez + 1 < ez1) tens_issue_B(ez + 1);
        }
      }
    }
    if (wact) {
      if (take) {  // column shared with the strip on the left (same group)
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 0; j < N; ++j) f[c][j] += hand[ez & 1][warp - 1][c][j];
      }
      // ---- vertical carry, then every node of rows j = 0..N-2 is final for this strip
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        f[c][0] += Fc[c];
        Fc[c] = f[c][N - 1];
      }
      if (st_ok) {
        // The inverse mass of a node that no boundary condition touches is the same for every
        // component (mat_mass.f90:56-57; only bc_abso.f90:243 makes the columns differ, on deferred
        // nodes), so one read of component 1 serves all of them.
        T vv[NDOF][N - 1], rm[N - 1];
        if constexpr (TENS && FUSED != 0) {
          const int kk = ez - ez0;
          mbar_wait_or_trap(&tbarB[kk & 1], (unsigned)((kk >> 1) & 1));
          const T* bV = reinterpret_cast<const T*>(tbase + (kk & 1) * TSZ + TSZ_A) + bx;
          const T* bR = reinterpret_cast<const T*>(tbase + (kk & 1) * TSZ + TOFF_R) + bx;
#pragma unroll
          for (int j = 0; j < N - 1; ++j) {
            rm[j] = bR[j * BW];
#pragma unroll
            for (int c = 0; c < NDOF; ++c) {
              vv[c][j] = bV[(c * (N - 1) + j) * BW];
              if (NM) vv[c][j] = vv[c][j] + A.c2 * bV[TSZ_A / sizeof(T) + (c * (N - 1) + j) * BW];  // predictor, solver.f90:60
            }
          }
        } else if (FUSED && (S2D_ABLATE & 8)) {
#pragma unroll
          for (int j = 0; j < N - 1; ++j) {
            rm[j] = (T)1e-9;
#pragma unroll
            for (int c = 0; c < NDOF; ++c) vv[c][j] = U[c][j];
          }
        } else if (FUSED) {  // requested at the top of this iteration; the next row's copies may still be in flight
          if (stg_v) {
            stage_wait<1>();
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
              rm[j] = st_r[j * 32];
#pragma unroll
              for (int c = 0; c < NDOF; ++c) {
                if constexpr (VDB) {
                  vv[c][j] = vk[c][j];
                } else {
                  vv[c][j] = st_v[(c * (N - 1) + j) * 32];
                  if (NM) vv[c][j] = vv[c][j] + A.c2 * st_a[(c * (N - 1) + j) * 32];  // predictor, solver.f90:60
                }
              }
            }
          } else {  // one batch of independent loads (deferred nodes included)
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
              const size_t q = rowbase + (size_t)j * LX + gx;
              rm[j] = A.rmass[q];
#pragma unroll
              for (int c = 0; c < NDOF; ++c) {
                if constexpr (VDB) {
                  vv[c][j] = vk[c][j];
                } else {
                  vv[c][j] = A.v_in[A.npoin * c + q];
                  if (NM) vv[c][j] = vv[c][j] + A.c2 * A.a_in[A.npoin * c + q];
                }
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
#pragma unroll
          for (int j = 0; j < N - 1; ++j) {
            const size_t q = A.npoin * c + rowbase + (size_t)j * LX + gx;
            if (!FUSED || ((defer >> j) & 1u)) {
              sp[cstride * c + (grow + j) * rstride] = f[c][j];
              // Newmark: the lane that stores into f owns the node and leaves the predicted velocity
              // for the boundary conditions and the deferred update
              if ((NM || VDB) && !to_halo) A.v_out[q] = vv[c][j];
            } else {  // solver.f90:157-158 (leapfrog) / :78-81 (Newmark), then the predictor of the next step
              const T acc = rm[j] * f[c][j];
              const T vn = vv[c][j] + A.c3 * acc;
              T dn = U[c][j] + A.dt * vn;
              if (NM) dn = dn + A.c1 * acc;
              if ((S2D_ABLATE & 16) && dn != (T)12345.678) continue;  // keeps the arithmetic alive, never stores
              A.v_out[q] = vn;
              A.d_next[q] = dn;
              if (A.a_out) A.a_out[q] = acc;
            }
          }
      }
#pragma unroll
      for (int c = 0; c < NDOF; ++c) U[c][0] = U[c][N - 1];
    }
  }
  // ---- top row of the band: final when nothing above shares it, else a partial sum for the band above
  if (st_ok) {
    if (ez1 == G.nz || (G.ezflt > 0 && ez1 == G.ezflt)) {
      const size_t gt = (size_t)strip_lat_row(G, ez1 - 1, N - 1);
      if (!FUSED || coldef || A.rowflag[gt]) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          const size_t q = A.npoin * c + gt * LX + gx;
          T vp = 0;
          if (NM && !to_halo) vp = A.v_in[q] + A.c2 * A.a_in[q];  // before f overwrites a
          else if (VDB && !to_halo) vp = A.v_in[q];
          sp[cstride * c + gt * rstride] = Fc[c];
          if ((NM || VDB) && !to_halo) A.v_out[q] = vp;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          const size_t q = A.npoin * c + gt * LX + gx;
          const T acc = A.rmass[gt * LX + gx] * Fc[c];
          T vp = A.v_in[q];
          if (NM) vp = vp + A.c2 * A.a_in[q];
          const T vn = vp + A.c3 * acc;
          A.v_out[q] = vn;
          T dn = U[c][0] + A.dt * vn;
          if (NM) dn = dn + A.c1 * acc;
          A.d_next[q] = dn;
          if (A.a_out) A.a_out[q] = acc;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < NDOF; ++c)
        A.halo_z[((size_t)(c * G.nseg + seg) * G.nstrips + strip) * G.WL + el * (N - 1) + i] = Fc[c];
    }
  }
  // ---- columns shared by two groups: the left group left its partial sums in halo_x, the right
  // group (the column's owner) in f.  Whichever of the two CTAs gets there second adds them -- the
  // same two addends either way, so the result does not depend on who it is -- and, in the fused
  // form, advances the nodes.  No CTA ever waits for another.  Rows shared by two bands stay with
  // k_strip_fold (they also need the band below).
  if (G.ngroups > 1) {
    __shared__ int second[2];
    __threadfence();
    __syncthreads();
    if (threadIdx.x < 2) {
      const int hb = grp - 1 + (int)threadIdx.x;  // boundary with the group on the left / on the right
      int sec = 0;
      if (hb >= 0 && hb < G.ngroups - 1) {
        int* m = A.meet + (size_t)seg * (G.ngroups - 1) + hb;
        sec = atomicAdd(m, 1);
        if (sec) *m = 0;  // armed again for the next evaluation
      }
      second[threadIdx.x] = sec;
    }
    __syncthreads();
    int ra = strip_lat_row(G, ez0, 0), rb = strip_lat_row(G, ez1 - 1, N - 1);
    if (strip_shared_row_seg(G, ra) >= 0) ++ra;
    if (!(ez1 == G.nz || (G.ezflt > 0 && ez1 == G.ezflt))) --rb;
    const int nrows = rb - ra + 1;
    const size_t hx_c = (size_t)(G.ngroups - 1) * G.LZ;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      if (!second[side]) continue;
      __threadfence();
      const int hb = grp - 1 + side;
      int sr;
      const int hx = strip_halo_col(G, hb, sr);
      const bool bcol = FUSED && A.colflag[hx] == 1;
      // two nodes per thread and pass, every load issued before the first use
#pragma unroll 1
      for (int k0 = threadIdx.x; k0 < nrows * NDOF; k0 += 2 * blockDim.x) {
        size_t q[2];
        T tot[2], rm[2], vv[2], dd[2];
        bool on[2], store_f[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = k0 + u * blockDim.x;
          on[u] = k < nrows * NDOF;
          const int kk = on[u] ? k : k0;
          const int c = kk / nrows, gz = ra + (kk - c * nrows);
          q[u] = A.npoin * c + (size_t)gz * LX + hx;
          tot[u] = __ldcg(A.f + q[u]) + __ldcg(A.halo_x + hx_c * c + (size_t)hb * G.LZ + gz);
          store_f[u] = !FUSED || bcol || A.rowflag[gz];
          if (FUSED) {
            rm[u] = A.rmass[q[u]];
            vv[u] = __ldcg((NM && VDB ? A.v_out : A.v_in) + q[u]);  // Newmark: already predicted by the column's owner lane
            dd[u] = A.d[q[u]];
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!on[u]) continue;
          if (store_f[u]) {
            A.f[q[u]] = tot[u];
          } else {
            const T acc = rm[u] * tot[u];
            const T vn = vv[u] + A.c3 * acc;
            A.v_out[q[u]] = vn;
            T dn = dd[u] + A.dt * vn;
            if (NM) dn = dn + A.c1 * acc;
            A.d_next[q[u]] = dn;
            if (A.a_out) A.a_out[q[u]] = acc;
          }
        }
      }
    }
  }
}


// Adds the partial sums that meet on the rows shared by two bands.  One thread per node, fixed order of
// additions ((f + left group) + band below + band below of the left strip): deterministic.
//   part A: the nodes of those rows that also lie on a group-boundary column
//   part B: the other nodes of those rows
// (group-boundary columns away from these rows are added inside k_elem_strip)
__host__ __device__ inline int strip_shared_upper_seg(const StripGeom& G, int r) {
  const int nsh_lo = G.nseg_lo > 0 ? G.nseg_lo - 1 : 0;
  return r < nsh_lo ? r + 1 : G.nseg_lo + 1 + (r - nsh_lo);
}
template <typename T>
__global__ void k_strip_fold(StripGeom G, T* __restrict__ f, const T* __restrict__ halo_x,
                             const T* __restrict__ halo_z, size_t npoin, StepCtl* tick) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tick && w == 0) tick->it += 1;  // the step counter rides here (the strip kernel before does not read it)
  const int nhb = G.ngroups - 1;
  const int nsh_lo = G.nseg_lo > 0 ? G.nseg_lo - 1 : 0;
  const int nsh = nsh_lo + (G.nseg - G.nseg_lo - 1);
  const long long nA = (long long)nhb * nsh;
  const size_t hz_c = (size_t)G.nseg * G.nstrips * G.WL;
  const size_t hx_c = (size_t)nhb * G.LZ;
  if (w < nA) {
    const int r = (int)(w / nhb), hb = (int)(w - (long long)r * nhb);
    const int seg_u = strip_shared_upper_seg(G, r);
    int ez0, ez1;
    strip_seg_rows(G, seg_u, ez0, ez1);
    const int gz = strip_lat_row(G, ez0, 0);
    int sr;
    const int gx = strip_halo_col(G, hb, sr);
    const size_t node = (size_t)gz * G.LXP + gx;
    const int seg_l = seg_u - 1;
    for (int c = 0; c < G.ndof; ++c) {
      T acc = f[node + npoin * c] + halo_x[hx_c * c + (size_t)hb * G.LZ + gz];
      acc += halo_z[hz_c * c + ((size_t)seg_l * G.nstrips + sr) * G.WL];
      acc += halo_z[hz_c * c + ((size_t)seg_l * G.nstrips + (sr - 1)) * G.WL + G.W];
      f[node + npoin * c] = acc;
    }
    return;
  }
  const long long w2 = w - nA;
  if (w2 >= (long long)nsh * G.LX) return;
  const int r = (int)(w2 / G.LX), gx = (int)(w2 - (long long)r * G.LX);
  const int seg_u = strip_shared_upper_seg(G, r);
  int strip = gx / G.W, lc = gx - strip * G.W;
  if (strip >= G.nstrips) {
    strip = G.nstrips - 1;
    lc = gx - strip * G.W;
  } else if (lc == 0 && strip > 0) {
    if (strip_group_of(G, strip - 1) != strip_group_of(G, strip)) return;  // group-boundary column: part A
  }
  if ((gx == 0 && G.xhalo_left) || (gx == G.LX - 1 && G.xhalo_right)) return;  // folded by k_xhalo_unpack
  int ez0, ez1;
  strip_seg_rows(G, seg_u, ez0, ez1);
  const size_t node = (size_t)strip_lat_row(G, ez0, 0) * G.LXP + gx;
  for (int c = 0; c < G.ndof; ++c)
    f[node + npoin * c] += halo_z[hz_c * c + ((size_t)(seg_u - 1) * G.nstrips + strip) * G.WL + lc];
}

// Leapfrog update of the deferred nodes once their force is complete (fold, sources, boundary
// conditions done): a = rmass*f, v += dt*a, d_next = d + dt*v  (solver.f90:157-158,151).
//   part A: flagged rows (contiguous);  part B: flagged columns, minus the nodes of flagged rows
template <typename T>
__global__ void k_strip_deferred(int LX, int LXP, int LZ, int ndof, size_t npoin, const int* __restrict__ drows, int ndrows,
                                 const int* __restrict__ dcols, int ndcols, const uint8_t* __restrict__ rowflag,
                                 T* __restrict__ fa, T* __restrict__ v, const T* __restrict__ rmass,
                                 const T* __restrict__ d, T* __restrict__ d_next, T dt, T c1, T c3) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nA = (long long)ndrows * LX;
  size_t node;
  if (w < nA) {
    const int r = (int)(w / LX);
    node = (size_t)drows[r] * LXP + (size_t)(w - (long long)r * LX);
  } else {
    const long long w2 = w - nA;
    if (w2 >= (long long)ndcols * LZ) return;
    const int gz = (int)(w2 / ndcols), k = (int)(w2 - (long long)gz * ndcols);
    if (rowflag[gz]) return;
    node = (size_t)gz * LXP + dcols[k];
  }
  for (int c = 0; c < ndof; ++c) {
    const size_t q = node + npoin * c;
    const T acc = rmass[q] * fa[q];
    const T vn = v[q] + c3 * acc;  // Newmark: v already holds the predicted velocity
    fa[q] = acc;
    v[q] = vn;
    T dn = d[q] + dt * vn;
    if (c1 != (T)0) dn = dn + c1 * acc;
    d_next[q] = dn;
  }
}

// a = rmass * f on the nodes the fused kernel advanced itself (everything that is not a deferred row / column)
template <typename T>
__global__ void k_strip_accel_fill(int LX, int LXP, int LZ, int ndof, size_t npoin, const uint8_t* __restrict__ rowflag,
                                   const uint8_t* __restrict__ colflag, const T* __restrict__ f, const T* __restrict__ rmass,
                                   T* __restrict__ a) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (long long)LX * LZ) return;
  const int gz = (int)(w / LX), gx = (int)(w - (long long)gz * LX);
  if (rowflag[gz] || colflag[gx] == 1) return;
  const size_t node = (size_t)gz * LXP + gx;
  for (int c = 0; c < ndof; ++c) a[node + npoin * c] = rmass[node + npoin * c] * f[node + npoin * c];
}

// x-strip interface columns (multi-GPU): this GPU's complete partial sum of lattice column 0 / LX-1,
// i.e. the stored force plus the band-top partial of the same strip.  buf[c][gz].
template <typename T>
__device__ __forceinline__ T xhalo_own(const StripGeom& G, const T* f, const T* halo_z, size_t npoin, int side,
                                       int c, int gz) {
  const int strip = side ? G.nstrips - 1 : 0;
  const int gx = side ? G.LX - 1 : 0;
  T acc = f[(size_t)gz * G.LXP + gx + npoin * c];
  const int seg_l = strip_shared_row_seg(G, gz);
  if (seg_l >= 0) {
    const size_t hz_c = (size_t)G.nseg * G.nstrips * G.WL;
    const int lc = side ? (G.LX - 1 - strip * G.W) : 0;
    acc += halo_z[hz_c * c + ((size_t)seg_l * G.nstrips + strip) * G.WL + lc];
  }
  return acc;
}
template <typename T>
__global__ void k_xhalo_pack(StripGeom G, const T* __restrict__ f, const T* __restrict__ halo_z, size_t npoin,
                             T* __restrict__ send_l, T* __restrict__ send_r) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = G.LZ * G.ndof;
  if (w >= 2 * n) return;
  const int side = w / n, q = w - side * n;
  T* dst = side ? send_r : send_l;
  if (!dst) return;
  dst[q] = xhalo_own(G, f, halo_z, npoin, side, q / G.LZ, q % G.LZ);
}
// Peer-memory exchange (NVLink): k_xhalo_pack writes straight into the neighbour GPU's receive slot;
// then one thread publishes the sequence number of this force evaluation in the neighbour's flag,
// and the neighbour's stream spins on its own flag before it unpacks.  Two receive slots (seq & 1):
// a GPU sends evaluation seq+1 only after it has unpacked seq, so slot seq & 1 is rewritten (at
// seq+2) only after its owner has read it.
static __global__ void k_xhalo_signal(unsigned long long* peer_flag_l, unsigned long long* peer_flag_r, unsigned long long seq) {
  __threadfence_system();
  if (peer_flag_l) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag_l), "l"(seq) : "memory");
  if (peer_flag_r) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag_r), "l"(seq) : "memory");
}
// flags[0] is written by the left neighbour, flags[1] by the right one.  Gives up after ~10 s with
// ctl->err = 3 instead of hanging the GPU (a neighbour that died never signals).
static __global__ void k_xhalo_wait(const unsigned long long* flags, int wait_l, int wait_r, unsigned long long seq, StepCtl* ctl) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int sd = 0; sd < 2; ++sd) {
    if (!(sd ? wait_r : wait_l)) continue;
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + sd) : "memory");
      if (v >= seq) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 10000000000ull) {
        ctl->err = 3;
        return;
      }
      __nanosleep(200);
    } while (true);
  }
  __threadfence_system();
}

// f = own + neighbour's; a + b == b + a bit for bit, so both GPUs hold the same value afterwards
template <typename T>
__global__ void k_xhalo_unpack(StripGeom G, T* __restrict__ f, const T* __restrict__ halo_z, size_t npoin,
                               const T* __restrict__ recv_l, const T* __restrict__ recv_r) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = G.LZ * G.ndof;
  if (w >= 2 * n) return;
  const int side = w / n, q = w - side * n;
  const T* src = side ? recv_r : recv_l;
  if (!src) return;
  const int c = q / G.LZ, gz = q % G.LZ;
  const int gx = side ? G.LX - 1 : 0;
  f[(size_t)gz * G.LXP + gx + npoin * c] = xhalo_own(G, f, halo_z, npoin, side, c, gz) + src[q];
}

// node-wise eta (a function of position, s2d_cart_set_kv) spread over the elements in the strip layout
template <typename T>
__global__ void k_strip_eta_from_nodes(StripGeom G, const T* __restrict__ eta_node, T* __restrict__ eta_strip) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = G.N, N2 = N * N;
  if (w >= (long long)G.nx * G.nz * N2) return;
  const long long e = w / N2;
  const int k = (int)(w - e * N2), i = k % N, j = k / N;
  const int ix = (int)(e % G.nx), iz = (int)(e / G.nx);
  eta_strip[strip_scalar_index(G, ix, iz, i, j)] = eta_node[(size_t)strip_lat_row(G, iz, j) * G.LXP + (size_t)ix * (N - 1) + i];
}

// E_W of energy_compute (energy.f90:84-104): sum over the elements of beta * |d|^2 (2.5D term), per-CTA partial sums
template <typename T>
__global__ void __launch_bounds__(256) k_strip_energy_w(StripGeom G, const T* __restrict__ beta, const T* __restrict__ d,
                                                        size_t nlat, double* __restrict__ partial) {
  __shared__ double red[256];
  const int N = G.N, N2 = N * N;
  const long long total = (long long)G.nx * G.nz * N2;
  double acc = 0.0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const long long e = w / N2;
    const int k = (int)(w - e * N2), i = k % N, j = k / N;
    const int ix = (int)(e % G.nx), iz = (int)(e / G.nx);
    const size_t q = (size_t)strip_lat_row(G, iz, j) * G.LXP + (size_t)ix * (N - 1) + i;
    double d2 = 0.0;
    for (int c = 0; c < G.ndof; ++c) d2 += (double)d[q + nlat * c] * (double)d[q + nlat * c];
    acc += (double)beta[strip_scalar_index(G, ix, iz, i, j)] * d2;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if ((int)threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// everything a strip launch needs besides the geometry
template <typename T>
struct StripIO {
  const T* coef;
  const T* d;
  T* f;
  T* halo_x;
  T* halo_z;
  size_t npoin;
  const double* hprime;
  // fused update (null v_in = plain force evaluation)
  const T* v_in = nullptr;
  T* v_out = nullptr;
  const T* rmass = nullptr;
  T* d_next = nullptr;
  T* a_out = nullptr;
  const uint8_t* rowflag = nullptr;
  const uint8_t* colflag = nullptr;
  int* meet = nullptr;
  double dt = 0.0;
  int newmark = 0;          // fused explicit Newmark instead of leapfrog
  double c1 = 0.0, c2 = 0.0, c3 = 0.0;
  const T* a_in = nullptr;
  const T* eta = nullptr;   // Kelvin-Voigt: eta per element GLL point (strip layout) and the velocity field
  const T* v_kv = nullptr;
  const unsigned char* pl_set = nullptr;  // Coulomb plasticity (see StripArgs)
  T* pl_ep = nullptr;
  const T* pl_tab = nullptr;
  T* vs_state = nullptr;                  // visco-elasticity (see StripArgs)
  const T* vs_tab = nullptr;
  int vs_nb = 0;
  T* dm_state = nullptr;                  // damage rheology (see StripArgs)
  const T* dm_tab = nullptr;
  int* dm_err = nullptr;
  const T* beta = nullptr;  // 2.5D: beta per element GLL point (strip layout)
  // tensor-map staging of the fused leapfrog kernel (null: per-lane copies)
  const CUtensorMap* tm_d = nullptr;
  const CUtensorMap* tm_v = nullptr;
  const CUtensorMap* tm_r = nullptr;
  const CUtensorMap* tm_a = nullptr;
  int prefetch = 1;
  // compact coefficient mode (coef holds lambda, mu only)
  int compact = 0;
  int occ = 0;              // resident CTAs per SM the kernel is compiled for (0 = default)
  double cdx = 0.0, cdz = 0.0, cdet = 0.0;
  const double* wgll = nullptr;
};

// largest ngll whose Kelvin-Voigt instantiation also carries the fused node update (compile time of the library)
constexpr int STRIP_KV_FUSED_MAXN = 6;
constexpr int STRIP_PLAST_MAXN = 6;  // plasticity instantiations
#ifndef S2D_PLAST_MINB
// CTAs per SM of the FP64 plasticity instantiation.  With the plastic strain loaded directly, 3 CTAs at 168 registers
// and ~300 B of spills beat 2 CTAs without spills (4.04 vs 4.36 ms, 2560^2); once the strain was staged one row ahead
// the order reversed: 2 CTAs 3.90 ms, 3 CTAs 4.31 ms (heavy yielding 4.35 vs 4.87).  Visco / damage keep 3 (not measured).
#define S2D_PLAST_MINB 2
#endif
template <typename T, int N, int NDOF, int FUSED, bool COMPACT, int MINB = strip_min_ctas(N, sizeof(T), COMPACT), bool KV = false,
          bool TENS = false, int RHEO = 0>
inline void strip_launch(unsigned nb, const StripArgs<T, N>& A, cudaStream_t s) {
  constexpr size_t smem = TENS ? strip_tens_smem(N, NDOF, sizeof(T), COMPACT, FUSED)
                               : strip_warps() * strip_stage_bytes(N, NDOF, sizeof(T), FUSED, COMPACT, RHEO == 1);
  // Opt in to the dynamic shared memory once per device and instantiation.  Never on the step path
  // afterwards: cudaFuncSetAttribute can serialise with running kernels, and a strip that is waiting
  // on its neighbour's flag must not keep the neighbour's host thread from launching.
  static bool done[64] = {};
  int dev = 0;
  S2D_CUDA(cudaGetDevice(&dev));
  if (!done[dev & 63]) {
    S2D_CUDA(cudaFuncSetAttribute(k_elem_strip<T, N, NDOF, FUSED, COMPACT, MINB, KV, TENS, RHEO>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done[dev & 63] = true;
  }
  k_elem_strip<T, N, NDOF, FUSED, COMPACT, MINB, KV, TENS, RHEO><<<nb, strip_warps() * 32, smem, s>>>(A);
}

// Element-force launch of one NGLL (all its instantiations of k_elem_strip).  Defined only in the translation units that
// set S2D_STRIP_CASES (strip_cases.inc: one TU per precision and NGLL range, so that the ~580 kernel instantiations
// compile in parallel and only once); everybody else sees the declaration.
template <typename T, int NN>
void launch_strip_case(const StripGeom& G, const StripIO<T>& io, cudaStream_t s);
#ifdef S2D_STRIP_CASES
template <typename T, int NN>
void launch_strip_case(const StripGeom& G, const StripIO<T>& io, cudaStream_t s) {
  const bool fused = io.v_in != nullptr;

    StripArgs<T, NN> A{};
    A.G = G;
    A.coef = io.coef;
    A.d = io.d;
    A.f = io.f;
    A.halo_x = io.halo_x;
    A.halo_z = io.halo_z;
    A.npoin = io.npoin;
    A.v_in = io.v_in;
    A.v_out = io.v_out;
    A.rmass = io.rmass;
    A.d_next = io.d_next;
    A.a_out = io.a_out;
    A.rowflag = io.rowflag;
    A.colflag = io.colflag;
    A.meet = io.meet;
    A.dt = (T)io.dt;
    A.c1 = (T)io.c1;
    A.c2 = (T)io.c2;
    A.c3 = (T)(fused && !io.newmark ? io.dt : io.c3);
    A.a_in = io.a_in;
    A.eta = io.eta;
    A.v_kv = io.v_kv;
    A.beta = io.beta;
    A.prefetch = io.prefetch;
    for (int k = 0; k < NN * NN; ++k) A.H[k] = (T)io.hprime[k];
    A.cdx = (T)io.cdx;
    A.cdz = (T)io.cdz;
    A.cdet = (T)io.cdet;
    for (int k = 0; k < NN; ++k) A.wg[k] = io.wgll ? (T)io.wgll[k] : (T)0;
    for (int k = 0; k < NN * NN; ++k) A.Hz[k] = (T)(io.cdz * io.hprime[k]);
    A.rzx = (T)(io.cdx != 0.0 ? io.cdz / io.cdx : 0.0);
    const unsigned nb = (unsigned)G.nitems;
    const int mode = !fused ? 0 : (io.newmark ? 2 : 1);
    if constexpr (NN <= S2D_STRIP_TENSOR_MAXN) { /* CTA-wide tensor-map staging of the fused step (S2D_STRIP_TENSOR) */
      if (mode == 0 && io.tm_d && !io.eta && !io.pl_set && (G.ndof == 1 || io.compact)) {
        // plain force evaluation: the displacement boxes alone (S2D_STRIP_TENSOR_PLAIN)
        constexpr int MB = strip_min_ctas(NN, sizeof(T));
        A.tm_d = *io.tm_d;
        if (G.ndof == 1) strip_launch<T, NN, 1, 0, false, MB, false, true>(nb, A, s);
        else strip_launch<T, NN, 2, 0, true, MB, false, true>(nb, A, s);
        return;
      }
      if (mode >= 1 && io.tm_d && !io.eta && !io.pl_set) {
        constexpr int MB = strip_min_ctas(NN, sizeof(T));
        A.tm_d = *io.tm_d;
        A.tm_v = *io.tm_v;
        A.tm_r = *io.tm_r;
        if (mode == 2) A.tm_a = *io.tm_a;
        if (G.ndof == 1) {
          if (mode == 2) strip_launch<T, NN, 1, 2, false, MB, false, true>(nb, A, s);
          else strip_launch<T, NN, 1, 1, false, MB, false, true>(nb, A, s);
        } else if (io.compact) {
          if (mode == 2) strip_launch<T, NN, 2, 2, true, MB, false, true>(nb, A, s);
          else strip_launch<T, NN, 2, 1, true, MB, false, true>(nb, A, s);
        } else {
          if (mode == 2) strip_launch<T, NN, 2, 2, false, MB, false, true>(nb, A, s);
          else strip_launch<T, NN, 2, 1, false, MB, false, true>(nb, A, s);
        }
        return;
      }
    }
    if (io.pl_set) { /* stateful rheologies: plasticity / visco-elasticity / damage, state per element GLL point */
      if constexpr (NN <= STRIP_PLAST_MAXN) {
        if (!io.compact || G.ndof != 2) throw ArgError("stateful rheologies: isotropic P-SV boxes");
        A.pl_set = io.pl_set;
        A.pl_ep = io.pl_ep;
        A.pl_tab = io.pl_tab;
        A.vs_state = io.vs_state;
        A.vs_tab = io.vs_tab;
        A.vs_nb = io.vs_nb;
        A.dm_state = io.dm_state;
        A.dm_tab = io.dm_tab;
        A.dm_err = io.dm_err;
        constexpr int MP1 = sizeof(T) == 8 ? S2D_PLAST_MINB : 3;   // plasticity
        constexpr int MP = 3;                                      // visco-elasticity, damage
        constexpr int MK = sizeof(T) == 8 ? 2 : 3;
        if (io.eta) { /* a Kelvin-Voigt layer on top (the one non-exclusive material; EXAMPLES/Damage: kind='DMG','KV') */
          if (mode != 0 && (io.v_in == io.v_out || (mode == 2 && (const T*)io.f == io.a_in)))
            throw ArgError("Kelvin-Voigt elements: the fused update needs separate output buffers");
          if (io.dm_state) {
            if (mode == 2) strip_launch<T, NN, 2, 2, true, MK, true, false, 3>(nb, A, s);
            else if (mode == 1) strip_launch<T, NN, 2, 1, true, MK, true, false, 3>(nb, A, s);
            else strip_launch<T, NN, 2, 0, true, MK, true, false, 3>(nb, A, s);
          } else if (io.vs_state) {
            if (mode == 2) strip_launch<T, NN, 2, 2, true, MK, true, false, 2>(nb, A, s);
            else if (mode == 1) strip_launch<T, NN, 2, 1, true, MK, true, false, 2>(nb, A, s);
            else strip_launch<T, NN, 2, 0, true, MK, true, false, 2>(nb, A, s);
          } else {
            if (mode == 2) strip_launch<T, NN, 2, 2, true, MK, true, false, 1>(nb, A, s);
            else if (mode == 1) strip_launch<T, NN, 2, 1, true, MK, true, false, 1>(nb, A, s);
            else strip_launch<T, NN, 2, 0, true, MK, true, false, 1>(nb, A, s);
          }
        } else if (io.dm_state) {
          if (mode == 2) strip_launch<T, NN, 2, 2, true, MP, false, false, 3>(nb, A, s);
          else if (mode == 1) strip_launch<T, NN, 2, 1, true, MP, false, false, 3>(nb, A, s);
          else strip_launch<T, NN, 2, 0, true, MP, false, false, 3>(nb, A, s);
        } else if (io.vs_state) {
          if (mode == 2) strip_launch<T, NN, 2, 2, true, MP, false, false, 2>(nb, A, s);
          else if (mode == 1) strip_launch<T, NN, 2, 1, true, MP, false, false, 2>(nb, A, s);
          else strip_launch<T, NN, 2, 0, true, MP, false, false, 2>(nb, A, s);
        } else {
          if (mode == 2) strip_launch<T, NN, 2, 2, true, MP1, false, false, 1>(nb, A, s);
          else if (mode == 1) strip_launch<T, NN, 2, 1, true, MP1, false, false, 1>(nb, A, s);
          else strip_launch<T, NN, 2, 0, true, MP1, false, false, 1>(nb, A, s);
        }
        return;
      }
      throw ArgError("stateful rheologies: ngll <= 6 only");
    }
    if (io.eta) { /* Kelvin-Voigt elements: plain force evaluation from d + eta*v */
      constexpr int MB = strip_min_ctas_kv(NN, sizeof(T), 1), M2 = strip_min_ctas_kv(NN, sizeof(T), 2);
      if (mode != 0) { /* fused with the node update: v (and a) double-buffered by the caller */
        if constexpr (NN <= STRIP_KV_FUSED_MAXN) {
          if (io.v_in == io.v_out || (mode == 2 && (const T*)io.f == io.a_in))
            throw ArgError("Kelvin-Voigt elements: the fused update needs separate output buffers");
          if (G.ndof == 1) {
            if (mode == 2) strip_launch<T, NN, 1, 2, false, MB, true>(nb, A, s);
            else strip_launch<T, NN, 1, 1, false, MB, true>(nb, A, s);
          } else if (io.compact) {
            if (mode == 2) strip_launch<T, NN, 2, 2, true, M2, true>(nb, A, s);
            else strip_launch<T, NN, 2, 1, true, M2, true>(nb, A, s);
          } else {
            if (mode == 2) strip_launch<T, NN, 2, 2, false, M2, true>(nb, A, s);
            else strip_launch<T, NN, 2, 1, false, M2, true>(nb, A, s);
          }
          return;
        }
        throw ArgError("Kelvin-Voigt elements: the node update is fused for ngll <= 6 only");
      }
      if (G.ndof == 1) strip_launch<T, NN, 1, 0, false, MB, true>(nb, A, s);
      else if (io.compact) strip_launch<T, NN, 2, 0, true, M2, true>(nb, A, s);
      else strip_launch<T, NN, 2, 0, false, M2, true>(nb, A, s);
      return;
    }
    if (G.ndof == 1) {
      if (io.compact) throw ArgError("compact coefficients need ndof = 2");
      if (mode == 2) strip_launch<T, NN, 1, 2, false>(nb, A, s);
      else if (mode == 1) strip_launch<T, NN, 1, 1, false>(nb, A, s);
      else strip_launch<T, NN, 1, 0, false>(nb, A, s);
    } else if (io.compact) {
      if constexpr (NN == 5 && sizeof(T) == 8) {  /* measured alternative: 4 CTAs/SM, spills */
        if (io.occ == 4 && mode < 2) {
          if (mode == 1) strip_launch<T, NN, 2, 1, true, 4>(nb, A, s);
          else strip_launch<T, NN, 2, 0, true, 4>(nb, A, s);
          return;
        }
      }
      if (mode == 2) strip_launch<T, NN, 2, 2, true>(nb, A, s);
      else if (mode == 1) strip_launch<T, NN, 2, 1, true>(nb, A, s);
      else strip_launch<T, NN, 2, 0, true>(nb, A, s);
    } else {
      if (mode == 2) strip_launch<T, NN, 2, 2, false>(nb, A, s);
      else if (mode == 1) strip_launch<T, NN, 2, 1, false>(nb, A, s);
      else strip_launch<T, NN, 2, 0, false>(nb, A, s);
    }

}
#endif

// element-force launch over the groups selected by G.it_* (no halo fold)
template <typename T>
inline void launch_elem_strip_items(const StripGeom& G, const StripIO<T>& io, cudaStream_t s) {
  if (G.nitems <= 0) return;
  switch (G.N) {
#ifndef S2D_ONLY_N5
    case 3: launch_strip_case<T, 3>(G, io, s); break;
    case 4: launch_strip_case<T, 4>(G, io, s); break;
#endif
    case 5: launch_strip_case<T, 5>(G, io, s); break;
#ifndef S2D_ONLY_N5
    case 6: launch_strip_case<T, 6>(G, io, s); break;
    case 7: launch_strip_case<T, 7>(G, io, s); break;
    case 8: launch_strip_case<T, 8>(G, io, s); break;
    case 9: launch_strip_case<T, 9>(G, io, s); break;
    case 10: launch_strip_case<T, 10>(G, io, s); break;
#endif
    default:
      throw ArgError("ngll must be in 3..10");
  }
  S2D_CUDA(cudaGetLastError());
}
template <typename T>
inline int launch_strip_fold(const StripGeom& G, T* f, const T* halo_x, const T* halo_z, size_t npoin,
                             cudaStream_t s, StepCtl* tick = nullptr) {
  const int nsh = (G.nseg_lo > 0 ? G.nseg_lo - 1 : 0) + (G.nseg - G.nseg_lo - 1);
  const long long nh = (long long)(G.ngroups - 1) * nsh + (long long)nsh * G.LX;
  if (nh <= 0 && !tick) return 0;
  k_strip_fold<T><<<(unsigned)((std::max(nh, 1LL) + 255) / 256), 256, 0, s>>>(G, f, halo_x, halo_z, npoin, tick);
  S2D_CUDA(cudaGetLastError());
  return 1;
}
// all groups of the box
inline StripGeom strip_all_groups(const StripGeom& G0) {
  StripGeom G = G0;
  G.it_g0 = 0;
  G.it_ng = G.ngroups;
  G.it_step = 1;
  G.nitems = (long long)G.nseg * G.ngroups;
  return G;
}

}  // namespace s2d
