// Element internal-force kernels: replace compute_Fint's element loop (SRC/solver.f90:291-299),
// FIELD_get_elem_sub / FIELD_add_elem (SRC/fields.f90:173-189,113-129), MAT_KV_add_etav
// (SRC/mat_kelvin_voigt.f90:137-150) and ELAST_KD1/KD2_{SH,PSV} (SRC/mat_elastic.f90:464-775)
// with mxm / My_MATMUL (SRC/mxmlib.f90:12-184, mat_elastic.f90:779-799) folded in.
//
// Operator (mat_elastic.f90:600-619 for P-SV flat; the planes a(:,:,q) already hold -w*|J|*c*metric):
//   Uxi = Ht U, Ueta = U H               (H(i,j) = h'_i(x_j), Ht its transpose)
//   f_c = H (tH_c) + (tHt_c) Ht          with tH_c, tHt_c pointwise combinations of a and gradients.
#pragma once
#include "common.cuh"

namespace s2d {

template <typename T>
struct ElemArgs {
  const int* elist;     // element ids (0-based) handled by this launch, or nullptr = 0..ne-1
  int ne;               // number of elements of this launch
  const int* ibool;     // (N*N, nelem), 1-based node ids as received (spec_grid.f90:52)
  const T* a;           // (N*N, nelast, ncoefsets)
  const int* elem2set;  // (nelem) 0-based
  const int* elem2kv;   // (nelem) -1 | 0-based KV slot, or nullptr
  const T* eta;         // (N*N, nkv)
  const T* beta;        // (N*N, ncoefsets) 2.5D term of MAT_ELAST_add_25D_f (mat_elastic.f90:447-459), or nullptr
  const T* H;           // (N,N) column-major hprime
  const T* d;
  const T* v;
  T* f;
  size_t npoin;
  int nelast;
  int kd2;
};

// pointwise stage: from the 2*NDOF gradients at one GLL point and the planes of that point, the
// values to be left-multiplied by H (tH) and right-multiplied by Ht (tHt).
template <typename T, int NDOF>
__device__ __forceinline__ void pointwise_stage(const T* __restrict__ a, int nelast, int kd2,
                                                const T gxi[NDOF], const T get[NDOF], T tH[NDOF],
                                                T tHt[NDOF]) {
  if (NDOF == 1) {
    if (nelast == 2) {  // ELAST_KD*_SH flat (mat_elastic.f90:751-762)
      tH[0] = a[0] * gxi[0];
      tHt[0] = a[1] * get[0];
    } else {  // general SH (mat_elastic.f90:663-676)
      tH[0] = a[0] * gxi[0] + a[2] * get[0];
      tHt[0] = a[2] * gxi[0] + a[1] * get[0];
    }
  } else {
    const T uxx = gxi[0], uzx = gxi[NDOF - 1], uxe = get[0], uze = get[NDOF - 1];
    if (nelast == 6) {  // P-SV flat (mat_elastic.f90:484-496 KD1, :600-619 KD2)
      // explicit fused multiply-adds: every kernel (and both coefficient modes of the strip kernel)
      // rounds these sums the same way, whatever the compiler would have contracted
      tH[0] = fma(a[0], uxx, a[1] * uze);
      tHt[0] = kd2 ? a[3] * (uxe + uzx) : fma(a[3], uxe, a[4] * uzx);
      tH[NDOF - 1] = fma(a[4], uxe, a[5] * uzx);
      tHt[NDOF - 1] = fma(a[1], uxx, a[2] * uze);
    } else {  // general P-SV, 10 planes (mat_elastic.f90:497-515)
      tH[0] = a[0] * uxx + a[6] * uxe + a[7] * uzx + a[1] * uze;
      tHt[0] = a[6] * uxx + a[3] * uxe + a[4] * uzx + a[8] * uze;
      tH[NDOF - 1] = a[7] * uxx + a[4] * uxe + a[5] * uzx + a[9] * uze;
      tHt[NDOF - 1] = a[1] * uxx + a[8] * uxe + a[9] * uzx + a[2] * uze;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// v1: one thread per GLL node, EPB elements per CTA, scatter straight to the global force array.
// Used with one launch per mesh colour (deterministic) or in a single launch with atomicAdd.
template <typename T, int N, int NDOF, bool ATOMIC>
__global__ void __launch_bounds__(256) k_elem_node(ElemArgs<T> A) {
  constexpr int N2 = N * N;
  constexpr int EPB = 256 / N2;
  __shared__ T sH[N2];
  __shared__ T sU[EPB][NDOF][N2];
  __shared__ T sP[EPB][NDOF][N2];  // tH
  __shared__ T sQ[EPB][NDOF][N2];  // tHt
  const int t = threadIdx.x;
  if (t < N2) sH[t] = A.H[t];
  const int slot = t / N2;
  const int k = t - slot * N2;
  const int i = k % N, j = k / N;
  const long long eidx = (long long)blockIdx.x * EPB + slot;
  const bool active = (slot < EPB) && (eidx < A.ne);
  size_t node = 0;
  T ar[10];
  if (active) {
    const int e = A.elist ? A.elist[eidx] : (int)eidx;
    node = (size_t)(A.ibool[(size_t)e * N2 + k] - 1);
    const T* ap = A.a + (size_t)A.elem2set[e] * A.nelast * N2 + k;
#pragma unroll
    for (int q = 0; q < 10; ++q)
      if (q < A.nelast) ar[q] = ap[(size_t)q * N2];
    T etav = 0;
    const int ikv = A.elem2kv ? A.elem2kv[e] : -1;
    if (ikv >= 0) etav = A.eta[(size_t)ikv * N2 + k];
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      T u = A.d[node + A.npoin * c];
      if (ikv >= 0) u = u + etav * A.v[node + A.npoin * c];  // mat_kelvin_voigt.f90:147
      sU[slot][c][k] = u;
    }
  }
  __syncthreads();
  if (active) {
    T gxi[NDOF], get[NDOF], tH[NDOF], tHt[NDOF];
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      T s1 = 0, s2 = 0;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        s1 += sH[m + N * i] * sU[slot][c][m + N * j];  // (Ht U)(i,j)
        s2 += sU[slot][c][i + N * m] * sH[m + N * j];  // (U H)(i,j)
      }
      gxi[c] = s1;
      get[c] = s2;
    }
    pointwise_stage<T, NDOF>(ar, A.nelast, A.kd2, gxi, get, tH, tHt);
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      sP[slot][c][k] = tH[c];
      sQ[slot][c][k] = tHt[c];
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      T s1 = 0, s2 = 0;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        s1 += sH[i + N * m] * sP[slot][c][m + N * j];  // (H tH)(i,j)
        s2 += sQ[slot][c][i + N * m] * sH[j + N * m];  // (tHt Ht)(i,j)
      }
      T val = s1 + s2;
      if (A.beta) {  // f = f - beta*d with the element's (KV-modified) d (mat_gen.f90:440)
        const int e = A.elist ? A.elist[eidx] : (int)eidx;
        val = val - A.beta[(size_t)A.elem2set[e] * N2 + k] * sU[slot][c][k];
      }
      T* dst = A.f + node + A.npoin * c;
      if (ATOMIC)
        atomicAdd(dst, val);
      else
        *dst += val;
    }
  }
}

template <typename T, int NDOF, bool ATOMIC>
inline void launch_elem_node(int ngll, const ElemArgs<T>& A, cudaStream_t s) {
  if (A.ne <= 0) return;
#define S2D_CASE(NN)                                                        \
  case NN: {                                                                \
    constexpr int EPB = 256 / (NN * NN);                                    \
    k_elem_node<T, NN, NDOF, ATOMIC><<<ceil_div(A.ne, EPB), 256, 0, s>>>(A); \
  } break;
  switch (ngll) {
#ifndef S2D_ONLY_N5  // kernel-experiment builds (make EXTRA=-DS2D_ONLY_N5) instantiate NGLL = 5 only
    S2D_CASE(3)
    S2D_CASE(4)
#endif
    S2D_CASE(5)
#ifndef S2D_ONLY_N5
    S2D_CASE(6)
    S2D_CASE(7)
    S2D_CASE(8)
    S2D_CASE(9)
    S2D_CASE(10)
#endif
    default:
      throw ArgError("ngll must be in 3..10");
  }
#undef S2D_CASE
}


// =============================================================================================
// v2: CTA-patch kernel (default).  One CTA owns a patch of up to EP clustered elements.
//   stage   : the patch's unique nodes are read once from HBM into shared memory (sD);
//   compute : one thread owns one GLL column (fixed j) of one element, i.e. N nodes, so the
//             xi-contractions run out of registers with hprime as constant-bank operands and only
//             the eta-contractions go through a shared-memory tile;
//   assemble: element forces are summed per patch in shared memory in the order of the in-patch
//             greedy colours (no atomics); nodes private to the patch are stored straight to the
//             global force array, partial sums of nodes shared with other patches go to a halo
//             slot, and a second small kernel adds the slots of each shared node in ascending
//             patch order.  Every node is written exactly once per step: no zero fill, no RMW.
// Index tables are split into a per-patch part (pnode: the global node of every patch-local node)
// and a per-SHAPE part (local indices, colours, slot offsets) that structured meshes share between
// all congruent patches, so it stays cache resident.
__device__ __forceinline__ void l2_prefetch(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// bulk L2 prefetch of a byte range (cp.async.bulk.prefetch: 16-byte aligned address and size)
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, size_t bytes) {
  const size_t a0 = (size_t)p & ~(size_t)15;
  size_t n = ((size_t)p + bytes - a0 + 15) & ~(size_t)15;
  while (n > 0) {
    const unsigned chunk = (unsigned)(n > 65536 ? 65536 : n);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0 + ((((size_t)p + bytes - a0 + 15) & ~(size_t)15) - n)), "r"(chunk) : "memory");
    n -= chunk;
  }
}

template <typename T, int N>
struct PatchArgs {
  int npatch, EP, max_nloc, max_colors;
  const int* pelem_start;     // (npatch+1) offsets into the patch-major element order
  const int* pshape;          // (npatch) shape id
  const uint16_t* sh_lidx;    // [shape][EP*N*N] patch-local node of each element node
  const uint8_t* sh_ecolor;   // [shape][EP] in-patch colour
  const int* sh_slot;         // [shape][max_nloc] -1 = private node, else slot offset
  const long long* pslot_base;  // (npatch) first halo slot of the patch (added to sh_slot)
  const long long* pnode_start; // (npatch+1)
  const int* pnode;           // global node (0-based) of each patch-local node
  const int* eset;            // (nelem patch-major) coefficient set when !hetero
  const int* ekv;             // (nelem patch-major) KV slot | -1, or nullptr
  const T* a;                 // hetero: per patch [nelast][N(i)][cnt*N(t)] ; else (N*N,nelast,nsets)
  const T* eta;               // (N*N, nkv)
  const T* beta;              // (N*N, ncoefsets) 2.5D term (mat_elastic.f90:447-459) indexed through eset, or nullptr
  const T* d;
  const T* v;
  T* f;
  T* fhalo;                   // (nslots, ndof)
  size_t npoin, nslots;
  int nelast, kd2, hetero;
  int pf_dist;                // software prefetch distance in patches (0 = off)
  int use_bulk;               // stream the coefficient block into shared memory with one TMA bulk copy
  T H[N * N];                 // hprime, column-major: read as constant-bank operands
};

// elements per patch: as many as keep the CTA at <= 384 threads (one thread per GLL column)
constexpr int patch_ep(int ngll) { return (384 / ngll) > 64 ? 64 : ((384 / ngll) < 8 ? 8 : (384 / ngll)); }
constexpr int patch_min_ctas(int ngll, int ndof, int tsize) {
  return (ngll <= 6 || tsize == 4) ? 2 : 1;
}

// ---- asynchronous copy helpers (sm_90+/sm_100a) ---------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src) {  // SASS: LDGSTS
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T, int N, int NDOF>
__global__ void __launch_bounds__(patch_ep(N) * N, patch_min_ctas(N, NDOF, sizeof(T)))
    k_elem_patch(const __grid_constant__ PatchArgs<T, N> A) {
  constexpr int N2 = N * N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  // shared-memory map (every region 16-byte aligned)
  T* sU = reinterpret_cast<T*>(smem_raw);                        // [EP][NDOF][N2] displacement, then tHt tiles
  T* sF = sU + align16((size_t)A.EP * NDOF * N2 * sizeof(T)) / sizeof(T);      // [NDOF][max_nloc] force sums
  T* sD = sF + align16((size_t)NDOF * A.max_nloc * sizeof(T)) / sizeof(T);     // [NDOF][max_nloc] staged displ
  T* sH = sD + align16((size_t)NDOF * A.max_nloc * sizeof(T)) / sizeof(T);     // [N2]
  T* sA = sH + align16((size_t)N2 * sizeof(T)) / sizeof(T);                     // [nelast][N][cnt*N] planes
  const int p = blockIdx.x;
  const int t = threadIdx.x;
  const int es = A.pelem_start[p];
  const int cnt = A.pelem_start[p + 1] - es;
  const int shape = A.pshape[p];
  const int el = t / N;
  const int j = t - el * N;
  const bool active = el < cnt;
  const int q = es + el;
  const long long ps = A.pnode_start[p];
  const int nloc = (int)(A.pnode_start[p + 1] - ps);
  const bool kv = A.ekv != nullptr;

  // (1) the patch's coefficient block -- the bulk of the bytes -- starts streaming into shared
  // memory through the TMA unit before anything else; it is consumed after the gradient stage.
  const T* ablock = A.a + (size_t)es * A.nelast * N2;
  const unsigned abytes = (unsigned)((size_t)cnt * A.nelast * N2 * sizeof(T));
  const bool bulk = A.use_bulk && A.hetero && (((size_t)ablock & 15) == 0) && ((abytes & 15) == 0);
  if (t == 0) {
    if (bulk) {
      mbar_init(&bar, 1);
      mbar_expect_tx(&bar, abytes);
      for (unsigned off = 0; off < abytes; off += 32768u) {
        const unsigned n = (abytes - off) < 32768u ? (abytes - off) : 32768u;
        bulk_g2s(reinterpret_cast<unsigned char*>(sA) + off, reinterpret_cast<const unsigned char*>(ablock) + off, n, &bar);
      }
    }
    // software prefetch into L2 for the patch that will run pf_dist CTAs from now
    if (A.pf_dist > 0 && p + A.pf_dist < A.npatch) {
      const int pf = p + A.pf_dist;
      if (A.hetero) {
        const int e0 = A.pelem_start[pf], e1 = A.pelem_start[pf + 1];
        l2_prefetch_bulk(A.a + (size_t)e0 * A.nelast * N2, (size_t)(e1 - e0) * A.nelast * N2 * sizeof(T));
      }
      if (p + 2 * A.pf_dist < A.npatch) {
        const int pf2 = p + 2 * A.pf_dist;
        l2_prefetch_bulk(A.pnode + A.pnode_start[pf2], (size_t)(A.pnode_start[pf2 + 1] - A.pnode_start[pf2]) * sizeof(int));
      }
    }
  }
  // (2) stage the patch's unique nodes: index loads first, then asynchronous copies (LDGSTS)
  if (t < N2) sH[t] = A.H[t];
  for (int l0 = t; l0 < nloc; l0 += 4 * blockDim.x) {
    int g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + u * blockDim.x;
      g[u] = l < nloc ? A.pnode[ps + l] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + u * blockDim.x;
      if (g[u] >= 0) {
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          cp_async<sizeof(T)>(&sD[c * A.max_nloc + l], A.d + (size_t)g[u] + A.npoin * c);
          if (kv) cp_async<sizeof(T)>(&sF[c * A.max_nloc + l], A.v + (size_t)g[u] + A.npoin * c);
          else sF[c * A.max_nloc + l] = (T)0;
        }
      }
    }
  }
  int li[N];
  if (active) {
    const uint16_t* lp = A.sh_lidx + ((size_t)shape * A.EP + el) * N2 + N * j;
#pragma unroll
    for (int i = 0; i < N; ++i) li[i] = lp[i];
  }
  cp_async_wait_all();
  __syncthreads();

  T* myU = sU + (size_t)el * NDOF * N2;
  T ucol[NDOF][N];
  if (active) {
    const int ikv = kv ? A.ekv[q] : -1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T etav = 0;
      if (ikv >= 0) etav = A.eta[(size_t)ikv * N2 + i + N * j];
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        T u = sD[c * A.max_nloc + li[i]];
        if (ikv >= 0) u = u + etav * sF[c * A.max_nloc + li[i]];  // mat_kelvin_voigt.f90:147
        ucol[c][i] = u;
        myU[c * N2 + i + N * j] = u;
      }
    }
  }
  __syncthreads();
  if (kv)
    for (int l = t; l < NDOF * A.max_nloc; l += blockDim.x) sF[l] = 0;

  T fcol[NDOF][N];   // force column accumulated in registers
  T tHt[NDOF][N];    // values that go through the tile exchange
  T HTj[N];          // H(j,m)
  if (active) {
    T Hj[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      Hj[m] = sH[m + N * j];   // H(m,j)
      HTj[m] = sH[j + N * m];  // H(j,m)
    }
    T gxi[NDOF][N], get[NDOF][N];
#pragma unroll
    for (int c = 0; c < NDOF; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s1 = 0, s2 = 0;
#pragma unroll
        for (int m = 0; m < N; ++m) {
          s1 += A.H[m + N * i] * ucol[c][m];      // (Ht U)(i,j), column in registers
          s2 += myU[c * N2 + i + N * m] * Hj[m];  // (U H)(i,j), row from the tile
        }
        gxi[c][i] = s1;
        get[c][i] = s2;
      }
    // pointwise stage with the coefficient planes of this column
    const T* ap;
    size_t qstride, istride;
    if (A.hetero) {
      ap = (bulk ? sA : ablock) + t;
      istride = (size_t)cnt * N;
      qstride = (size_t)N * istride;
    } else {
      ap = A.a + (size_t)A.eset[q] * A.nelast * N2 + N * j;
      qstride = N2;
      istride = 1;
    }
    if (bulk) mbar_wait(&bar, 0);
    T tH[NDOF][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T ar[10];
#pragma unroll
      for (int pl = 0; pl < 10; ++pl)
        if (pl < A.nelast) ar[pl] = ap[pl * qstride + i * istride];
      T g1[NDOF], g2[NDOF], o1[NDOF], o2[NDOF];
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        g1[c] = gxi[c][i];
        g2[c] = get[c][i];
      }
      pointwise_stage<T, NDOF>(ar, A.nelast, A.kd2, g1, g2, o1, o2);
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        tH[c][i] = o1[c];
        tHt[c][i] = o2[c];
      }
    }
#pragma unroll
    for (int c = 0; c < NDOF; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s1 = 0;
#pragma unroll
        for (int m = 0; m < N; ++m) s1 += A.H[i + N * m] * tH[c][m];  // (H tH)(i,j)
        fcol[c][i] = s1;
      }
    if (A.beta) {  // MAT_ELAST_add_25D_f: - beta*d with the element's (KV-modified) d (mat_gen.f90:440); the
      // (tHt Ht) half of the elastic force is added below, the sum per node is the same three terms
      const T* bp = A.beta + (size_t)A.eset[q] * N2 + N * j;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const T b = bp[i];
#pragma unroll
        for (int c = 0; c < NDOF; ++c) fcol[c][i] = fcol[c][i] - b * ucol[c][i];
      }
    }
  }
  __syncthreads();  // all reads of the displacement tiles are done
  if (active) {
#pragma unroll
    for (int c = 0; c < NDOF; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) myU[c * N2 + i + N * j] = tHt[c][i];
  }
  __syncthreads();
  int mycol = -1;
  if (active) {
#pragma unroll
    for (int c = 0; c < NDOF; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s2 = 0;
#pragma unroll
        for (int m = 0; m < N; ++m) s2 += myU[c * N2 + i + N * m] * HTj[m];  // (tHt Ht)(i,j)
        fcol[c][i] += s2;
      }
    mycol = A.sh_ecolor[(size_t)shape * A.EP + el];
  }
  // in-patch assembly, one colour at a time (elements of one colour share no node)
  for (int col = 0; col < A.max_colors; ++col) {
    if (mycol == col) {
#pragma unroll
      for (int c = 0; c < NDOF; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) sF[c * A.max_nloc + li[i]] += fcol[c][i];
    }
    __syncthreads();
  }
  const int* slotp = A.sh_slot + (size_t)shape * A.max_nloc;
  const long long sbase = A.pslot_base[p];
  for (int l = t; l < nloc; l += blockDim.x) {
    const int so = slotp[l];
    const size_t g = (size_t)A.pnode[ps + l];
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      const T val = sF[c * A.max_nloc + l];
      if (so < 0)
        A.f[g + A.npoin * c] = val;
      else
        A.fhalo[(size_t)(sbase + so) + A.nslots * c] = val;
    }
  }
  if (A.pf_dist > 0 && p + A.pf_dist < A.npatch) {
    const long long pfs = A.pnode_start[p + A.pf_dist];
    const int pfn = (int)(A.pnode_start[p + A.pf_dist + 1] - pfs);
    for (int l = t; l < pfn; l += blockDim.x) {
      const size_t g = (size_t)A.pnode[pfs + l];
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        l2_prefetch(A.d + g + A.npoin * c);
        if (kv) l2_prefetch(A.v + g + A.npoin * c);
      }
    }
  }
}


// adds the halo slots of each node shared between patches, in ascending patch order (generic plan)
template <typename T>
__global__ void k_halo_sum(T* __restrict__ f, const T* __restrict__ fhalo, const int* __restrict__ snode,
                           const int* __restrict__ sstart, int nshared, size_t npoin, size_t nslots,
                           int ndof) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nshared) return;
  const size_t g = (size_t)snode[s];
  const int b = sstart[s], e = sstart[s + 1];
  for (int c = 0; c < ndof; ++c) {
    T acc = fhalo[(size_t)b + nslots * c];
    for (int k = b + 1; k < e; ++k) acc += fhalo[(size_t)k + nslots * c];
    f[g + npoin * c] = acc;
  }
}

template <typename T, int N, int NDOF>
inline void launch_elem_patch_n(const PatchArgs<T, N>& A0, cudaStream_t s) {
  if (A0.npatch <= 0) return;
  PatchArgs<T, N> A = A0;
  const size_t base = align16((size_t)A.EP * NDOF * N * N * sizeof(T)) + 2 * align16((size_t)NDOF * A.max_nloc * sizeof(T)) +
                      align16((size_t)N * N * sizeof(T));
  const size_t blk = align16((size_t)A.EP * A.nelast * N * N * sizeof(T));
  // the TMA-staged coefficient block is optional: drop it when it does not fit beside two resident CTAs
  if (A.use_bulk && A.hetero && base + blk > 100 * 1024) A.use_bulk = 0;
  const size_t smem = base + ((A.use_bulk && A.hetero) ? blk : 0);
  static size_t configured = 0;
  if (smem > configured) {
    S2D_CUDA(cudaFuncSetAttribute(k_elem_patch<T, N, NDOF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = smem;
  }
  k_elem_patch<T, N, NDOF><<<A.npatch, A.EP * N, smem, s>>>(A);
}

}  // namespace s2d
