// Node-wise kernels of the time step: predictor / corrector (SRC/solver.f90:59-60,78-82,151-158),
// SO_add (SRC/src_gen.f90:290-317, src_force.f90:77-90), BC_ABSO_apply (SRC/bc_abso.f90:286-336),
// bc_DIRNEU_apply (SRC/bc_dirneu.f90:148-169), BC_DYNFLT_apply (SRC/bc_dynflt.f90:569-689) with the
// friction laws of bc_dynflt_swf/_rsf/_twf/_normal.f90, REC_store (SRC/receivers.f90:309-344),
// BC_DYNFLT_write (SRC/bc_dynflt.f90:751-778,832-855), progress maxima (SRC/main.f90:73-76) and
// kinetic energy (SRC/energy.f90:49-106).
// Boundary state is FP64 whatever the field precision T: these kernels touch O(boundary) nodes.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace s2d {

// ------------------------------------------------------------------------------------------
// step control
static __global__ void k_tick(StepCtl* ctl) { ctl->it += 1; }

// ------------------------------------------------------------------------------------------
// whole-array updates.  f aliases accel (solver.f90:52).
template <typename T>
__global__ void k_predict_leapfrog(T* __restrict__ d, const T* __restrict__ v, T* __restrict__ f,
                                   size_t n, T dt, int zero_f) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    d[q] = d[q] + dt * v[q];  // solver.f90:151
    if (zero_f) f[q] = 0;     // solver.f90:286
  }
}
// out-of-place displacement predictor of the fused step: dst = src + dt*v (+ c1*a for Newmark, solver.f90:59)
template <typename T>
__global__ void k_predict_to(T* __restrict__ dst, const T* __restrict__ src, const T* __restrict__ v,
                             const T* __restrict__ a, size_t n, T dt, T c1) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    T x = src[q] + dt * v[q];
    if (c1 != (T)0) x = x + c1 * a[q];
    dst[q] = x;
  }
}
template <typename T>
__global__ void k_predict_newmark(T* __restrict__ d, T* __restrict__ v, T* __restrict__ a, size_t n,
                                  T dt, T c1, T c2, int zero_f) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const T aq = a[q], vq = v[q];
    d[q] = d[q] + dt * vq + c1 * aq;  // solver.f90:59
    v[q] = vq + c2 * aq;              // solver.f90:60
    if (zero_f) a[q] = 0;
  }
}
// a = f*rmass ; v += c3*a ; d += c4*a   (solver.f90:78-82 ; leapfrog :157-158 with c3=dt,c4=0)
template <typename T>
__global__ void k_correct(T* __restrict__ d, T* __restrict__ v, T* __restrict__ a,
                          const T* __restrict__ rmass, size_t n, T c3, T c4) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const T aq = a[q] * rmass[q];
    a[q] = aq;
    v[q] = v[q] + c3 * aq;
    if (c4 != (T)0) d[q] = d[q] + c4 * aq;
  }
}

template <typename T>
__global__ void k_fill_n(T* x, size_t n, T val) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) x[q] = val;
}
// caller's node numbering <-> GLL lattice through an index table (generic handles routed to the strip kernel):
// to_ref: ref[k] = lat[lat_of[k]], else lat[lat_of[k]] = ref[k]
template <typename TS, typename TD>
__global__ void k_lat_permute(const TS* __restrict__ src, TD* __restrict__ dst, const int* __restrict__ lat_of,
                              size_t np_ref, size_t np_lat, int ncomp, int to_ref) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < np_ref; k += stride) {
    const size_t l = (size_t)lat_of[k];
    for (int c = 0; c < ncomp; ++c) {
      if (to_ref) dst[k + np_ref * c] = (TD)src[l + np_lat * c];
      else dst[l + np_lat * c] = (TD)src[k + np_ref * c];
    }
  }
}
static __global__ void k_remap_ids(int* ids, size_t n, const int* __restrict__ lat_of) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    if (ids[q] > 0) ids[q] = lat_of[ids[q] - 1] + 1;
}

// rmass = 1/M once every boundary condition has had its say on M (init.f90:112-116)
template <typename T>
__global__ void k_invert(T* x, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) x[q] = (T)1 / x[q];
}

// ------------------------------------------------------------------------------------------
// SO_add (src_gen.f90:290-317): point forces f(iglob,:) += dir*ampli (src_force.f90:77-90) and the terms of
// SRC_MOMENT_add (src_moment.f90:183-197).  The reference adds the sources one after the other; here every
// target node is owned by ONE thread that walks the terms landing on it in that same order (CSR built at
// commit), so sources that share a node neither race nor change the order of the additions.
template <typename T>
__global__ void k_sources(T* f, size_t npoin, int ndof, int nnodes, const int* __restrict__ tnode,
                          const int* __restrict__ tstart, const int* __restrict__ tsrc,
                          const double* __restrict__ tcoef /*(nterms,ndof)*/, int nterms, int nsrc,
                          const double* __restrict__ ampli /*[steps*nstages][nsrc]*/, const StepCtl* ctl, int stage,
                          int nstages) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnodes) return;
  const double* row = ampli + (size_t)(((ctl->it - ctl->it0) * nstages + stage) % ctl->nrows) * nsrc;
  const size_t node = (size_t)(tnode[k] - 1);
  for (int c = 0; c < ndof; ++c) {
    double acc = (double)f[node + npoin * c];
    for (int t = tstart[k]; t < tstart[k + 1]; ++t) acc = (double)(T)(acc + row[tsrc[t]] * tcoef[t + (size_t)nterms * c]);
    f[node + npoin * c] = (T)acc;
  }
}

// periodic boundary (BC_PERIO_set_field, bc_periodic.f90:77-86)
template <typename T>
__global__ void k_periodic(T* f, size_t npoin, int ndof, int np, const int* master, const int* slave) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= np) return;
  for (int c = 0; c < ndof; ++c) {
    const size_t m = (size_t)(master[k] - 1) + npoin * c, s = (size_t)(slave[k] - 1) + npoin * c;
    const T sum = f[m] + f[s];
    f[m] = sum;
    f[s] = sum;
  }
}

// out = d + eta*v with a node-wise eta (Kelvin-Voigt on structured boxes)
template <typename T>
__global__ void k_kv_combine(T* __restrict__ out, const T* __restrict__ d, const T* __restrict__ v,
                             const T* __restrict__ eta, size_t npoin, int ndof) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n = npoin * ndof;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) out[q] = d[q] + eta[q % npoin] * v[q];
}

// y += c*x  (symplectic stages, solver.f90:185,197)
template <typename T>
__global__ void k_axpy(T* __restrict__ y, const T* __restrict__ x, T* __restrict__ f, size_t n, T c, int zero_f) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    y[q] = y[q] + c * x[q];
    if (zero_f) f[q] = 0;
  }
}
// HHT-alpha predictor (solver.f90:108-113)
template <typename T>
__global__ void k_predict_hht(T* __restrict__ d, T* __restrict__ v, T* __restrict__ a, T* __restrict__ d_alpha,
                              T* __restrict__ v_alpha, size_t n, T dt, T c1, T c2, T alpha, int zero_f) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const T aq = a[q], v0 = v[q], d0 = d[q];
    const T d1 = d0 + dt * v0 + c1 * aq;
    const T v1 = v0 + c2 * aq;
    d[q] = d1;
    v[q] = v1;
    d_alpha[q] = alpha * d1 + ((T)1 - alpha) * d0;
    v_alpha[q] = alpha * v1 + ((T)1 - alpha) * v0;
    if (zero_f) a[q] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// absorbing boundary
struct AbsoDev {
  int np, ndof, is_flat, stacey, ngll;
  const int* node;     // (np) 1-based
  const double* C;     // (np,ndof)
  const double* n;     // (np,2)
  // Stacey: per boundary node, its (element, local index) incidences in ascending element order
  const int* st_start;  // (np+1)
  const int* st_elem;   // boundary element (0-based)
  const int* st_loc;    // local index i (0-based)
  const int* bibool;    // (ngll,nbe) 1-based boundary node index
  const double* K;      // (ngll,2,nbe)
  const double* Ht;     // (ngll,ngll) col-major, Ht(i,k) = H(k,i)
};

// entry k of one absorbing boundary (bc_abso.f90:286-336): only MxA of its own node is modified
template <typename T>
__device__ __forceinline__ void abso_entry(const AbsoDev& A, int k, const T* __restrict__ D, const T* __restrict__ V,
                                           T* MxA, size_t npoin) {
  const size_t node = (size_t)(A.node[k] - 1);
  if (A.is_flat || A.ndof == 1) {  // bc_abso.f90:309
    for (int c = 0; c < A.ndof; ++c) {
      const size_t q = node + npoin * c;
      MxA[q] = (T)((double)MxA[q] - A.C[k + (size_t)A.np * c] * (double)V[q]);
    }
  } else {  // bc_abso.f90:311-317
    const double v1 = V[node], v2 = V[node + npoin];
    const double nx = A.n[k], nz = A.n[k + A.np];
    const double vn = v1 * nx + v2 * nz;
    const double vn1 = vn * nx, vn2 = vn * nz;
    MxA[node] = (T)((double)MxA[node] - A.C[k] * vn1 - A.C[k + A.np] * (v1 - vn1));
    MxA[node + npoin] = (T)((double)MxA[node + npoin] - A.C[k] * vn2 - A.C[k + A.np] * (v2 - vn2));
  }
  if (A.stacey) {  // bc_abso.f90:321-334, gathered per boundary node in the reference's element order
    for (int c = 0; c < 2; ++c) {
      double kxd = 0.0;
      for (int p = A.st_start[k]; p < A.st_start[k + 1]; ++p) {
        const int e = A.st_elem[p], i = A.st_loc[p];
        double s = 0.0;
        for (int kk = 0; kk < A.ngll; ++kk) {
          const int bn = A.bibool[kk + (size_t)A.ngll * e] - 1;
          s += A.Ht[i + A.ngll * kk] * (double)D[(size_t)(A.node[bn] - 1) + npoin * c];
        }
        kxd = kxd + A.K[i + (size_t)A.ngll * (c + 2 * (size_t)e)] * s;
      }
      const size_t q = node + npoin * c;
      MxA[q] = (T)((double)MxA[q] - kxd);
    }
  }
}
template <typename T>
__global__ void k_abso(AbsoDev A, const T* __restrict__ D, const T* __restrict__ V, T* MxA,
                       size_t npoin) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.np) return;
  abso_entry<T>(A, k, D, V, MxA, npoin);
}

// SO_add and every BC_ABSO_apply of a step in ONE launch.  The sources and the absorbing sides meet on a few
// nodes (box corners belong to two sides; a source may sit on a side), where the reference applies them one after
// the other.  Every touched node is owned by one thread that walks ITS operations in that order: source terms
// first (solve: SO_add before BC_apply, solver.f90:70-75,156), then the absorbing boundaries in the order of
// bc(:) (bc_gen.f90:283-290).  op >= 0: entry (op >> 3) of absorbing boundary (op & 7); op < 0: source term -op-1.
template <typename T>
__global__ void k_node_ops(int nnodes, const int* __restrict__ onode, const int* __restrict__ ostart,
                           const int* __restrict__ ops, const AbsoDev* __restrict__ absos, const T* __restrict__ D,
                           const T* __restrict__ V, T* f, size_t npoin, int ndof, const int* __restrict__ tsrc,
                           const double* __restrict__ tcoef, int nterms, int nsrc, const double* __restrict__ ampli,
                           const StepCtl* ctl) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nnodes) return;
  const size_t node = (size_t)(onode[w] - 1);
  for (int p = ostart[w]; p < ostart[w + 1]; ++p) {
    const int op = ops[p];
    if (op < 0) {
      const int t = -op - 1;
      const double amp = ampli[(size_t)((ctl->it - ctl->it0) % ctl->nrows) * nsrc + tsrc[t]];
      for (int c = 0; c < ndof; ++c)
        f[node + npoin * c] = (T)((double)f[node + npoin * c] + amp * tcoef[t + (size_t)nterms * c]);
    } else {
      abso_entry<T>(absos[op & 7], op >> 3, D, V, f, npoin);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Dirichlet / Neumann (bc_dirneu.f90:148-169)
template <typename T>
__global__ void k_dirneu(T* f, size_t npoin, int ndof, int np, const int* node, int kind_h,
                         int kind_v, const double* B_h, const double* B_v, const double* ampli,
                         int nampl, int slot, const StepCtl* ctl) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= np) return;
  const size_t nd = (size_t)(node[k] - 1);
  const double* row = ampli ? ampli + (size_t)((ctl->it - ctl->it0) % ctl->nrows) * nampl : nullptr;
  if (kind_h == 2)
    f[nd] = 0;
  else if (B_h && row)
    f[nd] = (T)((double)f[nd] + row[slot] * B_h[k]);
  if (ndof == 1) return;
  if (kind_v == 2)
    f[nd + npoin] = 0;
  else if (B_v && row)
    f[nd + npoin] = (T)((double)f[nd + npoin] + row[slot + 1] * B_v[k]);
}

// ------------------------------------------------------------------------------------------
// dynamic fault
struct FaultDev {
  int np, ndof, two_sides, allow_opening;
  double CoefA2V, CoefA2D, dt, tshift;
  const int *node1, *node2;
  const double *n1, *B, *invM1, *invM2, *Z, *T0, *cohesion, *coord;
  double *T, *Tstick, *V, *D, *MU, *sigma;
  int swf_kind, swf_healing;
  const double *swf_dc, *swf_mus, *swf_mud, *swf_p, *swf_alpha;
  double* swf_theta;
  int rsf_kind;
  const double *rsf_dc, *rsf_mus, *rsf_a, *rsf_b, *rsf_Vstar, *rsf_Vc, *rsf_Tc, *rsf_coeft;
  double* rsf_theta;
  int twf_kind;
  double twf_X, twf_Z, twf_mus, twf_mud, twf_mu0, twf_L, twf_V, twf_T, twf_Dc;
  int normal_kind;
  double normal_V, normal_coef;
  // outputs
  int oix1, oixn, oixd, oitd, onx, nrec_max, ncall_max;
  int* ostate;     // [0] oit (next output step), [1] nout, [2] ncalls
  float* records;  // [nrec_max][6][onx]
  double* potency; // [ncall_max][2*(ndof+1)]
};

// swf_mu (bc_dynflt_swf.f90:142-160)
__device__ inline double swf_mu_dev(const FaultDev& F, int k, double theta) {
  double mu = 0.0;
  const double mus = F.swf_mus[k], mud = F.swf_mud[k], dc = F.swf_dc[k];
  if (F.swf_kind == 1)
    mu = mus - (mus - mud) * fmin(theta / dc, 1.0);
  else if (F.swf_kind == 2)
    mu = mud - (mud - mus) * exp(-theta / dc);
  else if (F.swf_kind == 3)
    mu = mud + (mus - mud) / pow(1.0 + theta / dc, F.swf_p[k]);
  return mu + F.swf_alpha[k] * theta;
}
// twf_mu (bc_dynflt_twf.f90:117-184)
__device__ inline double twf_mu_dev(const FaultDev& F, double x, double z, double time, double d) {
  const double BIG = DBL_MAX;
  double t, r = 0.0, mu = BIG;
  if (F.twf_kind == 1) {
    t = time + (F.twf_mus - F.twf_mu0) * F.twf_L / ((F.twf_mus - F.twf_mud) * F.twf_V);
    if (t > F.twf_T) t = 0.0;
    r = F.twf_V * t;
  } else if (F.twf_kind == 2) {
    t = time + 0.5 * F.twf_T *
                   (1.0 - sqrt(1.0 - 4.0 * (F.twf_mus - F.twf_mu0) * F.twf_L /
                                         ((F.twf_mus - F.twf_mud) * F.twf_T * F.twf_V)));
    t = fmin(t, F.twf_T);
    r = F.twf_V * t * (1.0 - t / F.twf_T);
  }
  const double dist = sqrt((x - F.twf_X) * (x - F.twf_X) + (z - F.twf_Z) * (z - F.twf_Z));
  if (F.twf_kind == 1 || F.twf_kind == 2) {
    const double rr = dist - r;
    if (rr < -F.twf_L)
      mu = F.twf_mud;
    else if (rr <= F.twf_L)
      mu = F.twf_mus + (F.twf_mus - F.twf_mud) / F.twf_L * rr;
  } else {
    t = time;
    if (dist <= F.twf_V * F.twf_T && d <= F.twf_Dc) {
      if (dist < F.twf_V * t - F.twf_L)
        mu = F.twf_mud;
      else if (dist <= F.twf_V * t)
        mu = F.twf_mus + (F.twf_mus - F.twf_mud) / F.twf_L * (dist - F.twf_V * t);
    }
  }
  return mu;
}
// rsf_mu (bc_dynflt_rsf.f90:165-184)
__device__ inline double rsf_mu_dev(const FaultDev& F, int k, double v, double theta) {
  const double av = fabs(v);
  if (F.rsf_kind == 1)
    return F.rsf_mus[k] + F.rsf_a[k] * av / (av + F.rsf_Vstar[k]) -
           F.rsf_b[k] * theta / (theta + F.rsf_dc[k]);
  double arg;
  if (F.rsf_kind == 4)
    arg = F.rsf_Vc[k] * theta / F.rsf_dc[k] + 1.0;
  else
    arg = F.rsf_Vstar[k] * theta / F.rsf_dc[k];
  return F.rsf_a[k] *
         asinh(av / (2.0 * F.rsf_Vstar[k]) * exp((F.rsf_mus[k] + F.rsf_b[k] * log(arg)) / F.rsf_a[k]));
}
// rsf_update_theta (bc_dynflt_rsf.f90:273-308)
__device__ inline double rsf_theta_dev(const FaultDev& F, int k, double theta, double v) {
  double tn = 0.0;
  if (F.rsf_kind == 1) {
    tn = theta * F.rsf_coeft[k] + F.rsf_Tc[k] * fabs(v) * (1.0 - F.rsf_coeft[k]);
    if (tn < 1.0e-12) tn = 0.0;
  } else if (F.rsf_kind == 2 || F.rsf_kind == 4) {
    const double x = fabs(v) / F.rsf_dc[k];
    const double ex = exp(-F.dt * x);
    if (F.dt * x > 1e-8)
      tn = theta * ex + (1.0 - ex) / x;
    else
      tn = theta * ex + F.dt * (1.0 - 0.5 * F.dt * x);
  } else {
    tn = F.rsf_dc[k] / fabs(v);
    tn = tn * pow(theta / tn, exp(-F.dt / tn));
  }
  return tn;
}
// nr_fric_func_tau (bc_dynflt_rsf.f90:533-568)
__device__ inline void nr_func_dev(const FaultDev& F, int k, double tau, double theta,
                                   double tau_stick, double sigma, double Z, double& func,
                                   double& dfunc, double& v) {
  double tmp;
  if (F.rsf_kind == 4)
    tmp = F.rsf_mus[k] + F.rsf_b[k] * log(F.rsf_Vc[k] * theta / F.rsf_dc[k] + 1.0);
  else
    tmp = F.rsf_mus[k] + F.rsf_b[k] * log(F.rsf_Vstar[k] * theta / F.rsf_dc[k]);
  tmp = 2.0 * F.rsf_Vstar[k] * exp(-tmp / F.rsf_a[k]);
  const double s = -sigma * F.rsf_a[k];
  v = sinh(tau / s) * tmp;
  func = tau_stick - Z * v - tau;
  dfunc = -Z * (cosh(tau / s) * tmp / s) - 1.0;
}
// nr_solver (bc_dynflt_rsf.f90:369-468): returns the velocity of the last function evaluation
__device__ inline double nr_solver_dev(const FaultDev& F, int k, double xL, double xR, double x_acc,
                                       double theta, double tau_stick, double sigma, double Z,
                                       int* err) {
  double v = 0, dfunc, f_low, f_high, func, x_est, dx, dx_old, x_high, x_low;
  nr_func_dev(F, k, xL, theta, tau_stick, sigma, Z, f_low, dfunc, v);
  nr_func_dev(F, k, xR, theta, tau_stick, sigma, Z, f_high, dfunc, v);
  double xLeft = xL, xRight = xR;
  int guard = 0;
  while (f_low * f_high > 0) {
    xLeft = xLeft / 2.0;
    xRight = xRight * 2.0;
    nr_func_dev(F, k, xLeft, theta, tau_stick, sigma, Z, f_low, dfunc, v);
    nr_func_dev(F, k, xRight, theta, tau_stick, sigma, Z, f_high, dfunc, v);
    if (++guard > 2000) {  // the reference would spin forever here; report instead
      atomicExch(err, 2);
      return v;
    }
  }
  if (f_low == 0) return v;
  if (f_high == 0) return v;
  if (f_low < 0) {
    x_low = xLeft;
    x_high = xRight;
  } else {
    x_high = xLeft;
    x_low = xRight;
  }
  x_est = 0.5 * (xLeft + xRight);
  dx_old = fabs(xRight - xLeft);
  dx = dx_old;
  nr_func_dev(F, k, x_est, theta, tau_stick, sigma, Z, func, dfunc, v);
  for (int is = 1; is <= 200; ++is) {
    if (((x_est - x_high) * dfunc - func) * ((x_est - x_low) * dfunc - func) > 0 ||
        fabs(2 * func) > fabs(dx_old * dfunc)) {
      dx_old = dx;
      dx = 0.5 * (x_high - x_low);
      x_est = x_low + dx;
      if (x_low == x_est) return v;
    } else {
      dx_old = dx;
      dx = func / dfunc;
      const double temp = x_est;
      x_est = x_est - dx;
      if (temp == x_est) return v;
    }
    if (fabs(dx) < fabs(x_acc)) {
      nr_func_dev(F, k, x_est, theta, tau_stick, sigma, Z, func, dfunc, v);
      return v;
    }
    nr_func_dev(F, k, x_est, theta, tau_stick, sigma, Z, func, dfunc, v);
    if (func < 0)
      x_low = x_est;
    else
      x_high = x_est;
  }
  atomicExch(err, 1);  // IO_abort('NR_Solver has exceeded the maximum iterations (200)')
  return v;
}
// rsf_update_V (bc_dynflt_rsf.f90:325-358); theta_stored = f%theta (kind 1 uses it, :338)
__device__ inline double rsf_update_V_dev(const FaultDev& F, int k, double tau_stick, double sigma,
                                          double theta, double theta_stored, double Z, int* err) {
  if (F.rsf_kind == 1) {
    const double mu_nd = F.rsf_mus[k] - F.rsf_b[k] * theta_stored / (theta_stored + F.rsf_dc[k]);
    double v = (tau_stick + sigma * mu_nd) / Z;
    const double tmp = v - F.rsf_Vstar[k] + sigma * F.rsf_a[k] / Z;
    v = 0.5 * (tmp + sqrt(tmp * tmp + 4.0 * v * F.rsf_Vstar[k]));
    v = fmax(0.0, v);
    if (v < 1.0e-12) v = 0.0;
    return v;
  }
  const double tol = -0.001 * F.rsf_a[k] * sigma;
  return nr_solver_dev(F, k, fmin(0.0, tau_stick), fmax(0.0, tau_stick), tol, theta, tau_stick,
                       sigma, Z, err);
}

__device__ inline void rot_fwd(double nx, double nz, double& a, double& b) {  // bc_dynflt.f90:731-733
  const double v1 = a, v2 = b;
  a = nz * v1 - nx * v2;
  b = nx * v1 + nz * v2;
}

template <typename T>
__global__ void k_dynflt(FaultDev F, T* MxA, const T* __restrict__ Vf, const T* __restrict__ Df,
                         size_t npoin, StepCtl* ctl) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = F.np;
  if (k >= np) return;
  const int ndof = F.ndof;
  const double time = (double)ctl->it * F.dt + F.tshift;  // HHT-alpha: t_alpha (solver.f90:116-121)
  double dD[2] = {0, 0}, dV[2] = {0, 0}, dA[2] = {0, 0}, Tt[2] = {0, 0}, Ts[2] = {0, 0};
  size_t i1[2], i2[2] = {0, 0};
  for (int c = 0; c < ndof; ++c) {
    const size_t q = k + (size_t)np * c;
    i1[c] = (size_t)(F.node1[k] - 1) + npoin * c;
    if (F.two_sides) {  // get_jump / get_weighted_jump (bc_dynflt.f90:693-719)
      i2[c] = (size_t)(F.node2[k] - 1) + npoin * c;
      dD[c] = (double)Df[i2[c]] - (double)Df[i1[c]];
      dV[c] = (double)Vf[i2[c]] - (double)Vf[i1[c]];
      dA[c] = F.invM2[q] * (double)MxA[i2[c]] - F.invM1[q] * (double)MxA[i1[c]];
    } else {
      dD[c] = -2.0 * (double)Df[i1[c]];
      dV[c] = -2.0 * (double)Vf[i1[c]];
      dA[c] = -2.0 * F.invM1[q] * (double)MxA[i1[c]];
    }
    Tt[c] = F.Z[q] * (dV[c] + F.CoefA2V * dA[c]);  // :593
  }
  const double nx = F.n1[k], nz = F.n1[k + np];
  if (ndof == 2) {  // :596-601
    rot_fwd(nx, nz, dD[0], dD[1]);
    rot_fwd(nx, nz, dV[0], dV[1]);
    rot_fwd(nx, nz, dA[0], dA[1]);
    rot_fwd(nx, nz, Tt[0], Tt[1]);
  }
  if (!F.two_sides || ndof == 1) Tt[1] = 0.0;  // :604
  const double T0a = F.T0[k], T0b = F.T0[k + np];
  Tt[0] += T0a;
  Tt[1] += T0b;
  if (F.allow_opening) Tt[1] = fmin(Tt[1], 0.0);  // :611
  // normal_update (bc_dynflt_normal.f90:118-136)
  double sigma = F.sigma[k];
  switch (F.normal_kind) {
    case 1: sigma = Tt[1]; break;
    case 2: sigma = Tt[1] + F.normal_coef * (sigma - Tt[1]); break;
    case 3: sigma = Tt[1] + exp(-(fabs(dV[0]) + F.normal_V) * F.normal_coef) * (sigma - Tt[1]); break;
    default: break;
  }
  F.sigma[k] = sigma;
  const double cx = F.coord[2 * k], cz = F.coord[2 * k + 1];
  double MU = F.MU[k];
  if (F.rsf_kind) {  // rsf_solver (bc_dynflt_rsf.f90:229-249)
    const double v_old = F.V[k], th_old = F.rsf_theta[k], Z1 = F.Z[k];
    double th = rsf_theta_dev(F, k, th_old, v_old);
    double vn = rsf_update_V_dev(F, k, Tt[0], sigma, th, th_old, Z1, &ctl->err);
    th = rsf_theta_dev(F, k, th_old, 0.5 * (v_old + vn));
    vn = rsf_update_V_dev(F, k, Tt[0], sigma, th, th_old, Z1, &ctl->err);
    F.rsf_theta[k] = th;
    MU = rsf_mu_dev(F, k, vn, th);
    if (F.twf_kind) MU = fmin(MU, twf_mu_dev(F, cx, cz, time, F.D[k]));
    const double strength = -MU * sigma;
    Tt[0] = copysign(fabs(strength), Tt[0]);  // sign(strength,T)
    // Tstick is never assigned in this branch of the reference (:620-636)
  } else {
    if (F.swf_kind) {
      double theta = F.swf_theta[k];
      if (F.CoefA2D == 0.0) {  // swf_update_state (bc_dynflt_swf.f90:163-181)
        if (F.swf_healing) {
          theta = theta + fabs(dV[0]) * F.dt;
          if (fabs(dV[0]) < 1e-14) theta = 0.0;
        } else {
          theta = fabs(dD[0]);
        }
      } else {
        theta = fabs(F.D[k]);  // swf_set_state
      }
      F.swf_theta[k] = theta;
      MU = swf_mu_dev(F, k, theta);
      if (F.twf_kind) MU = fmin(MU, twf_mu_dev(F, cx, cz, time, F.D[k]));
    } else if (F.twf_kind) {
      MU = twf_mu_dev(F, cx, cz, time, F.D[k]);
    }
    const double strength = F.cohesion[k] - MU * sigma;
    Ts[0] = Tt[0];
    Ts[1] = Tt[1];
    Tt[0] = copysign(fabs(fmin(fabs(Tt[0]), strength)), Tt[0]);  // :664
  }
  F.MU[k] = MU;
  Tt[0] -= T0a;
  Tt[1] -= T0b;
  Ts[0] -= T0a;
  Ts[1] -= T0b;
  F.T[k] = Tt[0];
  F.T[k + np] = Tt[1];
  F.Tstick[k] = Ts[0];
  F.Tstick[k + np] = Ts[1];
  double Tx = Tt[0], Tz = Tt[1];
  if (ndof == 2) {  // rotate back (:735-737)
    Tx = nz * Tt[0] + nx * Tt[1];
    Tz = -nx * Tt[0] + nz * Tt[1];
  }
  const double Tg[2] = {Tx, Tz};
  for (int c = 0; c < ndof; ++c) {
    const size_t q = k + (size_t)np * c;
    MxA[i1[c]] = (T)((double)MxA[i1[c]] + F.B[q] * Tg[c]);                    // :681
    if (F.two_sides) MxA[i2[c]] = (T)((double)MxA[i2[c]] - F.B[q] * Tg[c]);  // :682
    const double dAc = dA[c] - Tt[c] / (F.Z[q] * F.CoefA2V);                  // :685
    F.D[q] = dD[c] + F.CoefA2D * dAc;
    F.V[q] = dV[c] + F.CoefA2V * dAc;
  }
}

// ------------------------------------------------------------------------------------------
// receivers
struct RecDev {
  int nx, ndof, isamp, nt, at_node, ngll;
  const int* iglob;    // (nx) 1-based
  const int* einterp;  // (nx) 1-based element
  const double* interp;  // (ngll*ngll, nx)
  const int* ibool;
  float* sis;  // (nt,nx,ndof)
};
// REC_store (receivers.f90:309-344) for trace q = station + nx * component
template <typename T>
__device__ __forceinline__ void rec_store_one(const RecDev& R, const T* __restrict__ field, size_t npoin, int it, int q) {
  if (it % R.isamp != 0) return;
  const int itsis = it / R.isamp;  // 0-based row
  if (itsis >= R.nt || q >= R.nx * R.ndof) return;
  const int n = q % R.nx, c = q / R.nx;
  double val;
  if (R.at_node) {
    val = (double)field[(size_t)(R.iglob[n] - 1) + npoin * c];
  } else {
    const int n2 = R.ngll * R.ngll;
    const int* ib = R.ibool + (size_t)(R.einterp[n] - 1) * n2;
    double s = 0.0;
    for (int k = 0; k < n2; ++k)
      s += R.interp[k + (size_t)n2 * n] * (double)field[(size_t)(ib[k] - 1) + npoin * c];
    val = s;
  }
  R.sis[(size_t)itsis + (size_t)R.nt * (n + (size_t)R.nx * c)] = (float)val;
}
template <typename T>
__global__ void k_rec_store(RecDev R, const T* __restrict__ field, size_t npoin, const StepCtl* ctl) {
  rec_store_one<T>(R, field, npoin, ctl->it, blockIdx.x * blockDim.x + threadIdx.x);
}

// BC_DYNFLT_write.  Every CTA reduces its slice of the fault in a fixed tree order and copies its
// share of the output records; the CTA that finishes last (ticket) adds the per-CTA partial sums
// in ascending CTA order, writes the potency line and advances the output state: deterministic.
constexpr int DYNW_THREADS = 256;
constexpr int DYNW_MAX_CTAS = 128;
template <typename T>
__global__ void __launch_bounds__(DYNW_THREADS) k_dynflt_write(FaultDev F, const T* __restrict__ d,
                                                               const T* __restrict__ v, size_t npoin,
                                                               const StepCtl* ctl, double* __restrict__ part,
                                                               unsigned* __restrict__ ticket, int nfb, RecDev R,
                                                               const T* __restrict__ rec_field) {
  __shared__ double red[6][DYNW_THREADS];
  __shared__ bool last;
  // REC_store rides in the CTAs beyond the first nfb (one launch for the outputs of a step)
  if ((int)blockIdx.x >= nfb) {
    rec_store_one<T>(R, rec_field, npoin, ctl->it, ((int)blockIdx.x - nfb) * DYNW_THREADS + (int)threadIdx.x);
    return;
  }
  const int t = threadIdx.x, np = F.np, ndof = F.ndof;
  const int stride = nfb * DYNW_THREADS;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int k = blockIdx.x * DYNW_THREADS + t; k < np; k += stride) {
    const double nx = F.n1[k], nz = F.n1[k + np], B = F.B[k];
    for (int w = 0; w < 2; ++w) {
      const T* fld = w ? v : d;
      double j0, j1 = 0.0;
      const size_t a1 = (size_t)(F.node1[k] - 1);
      if (F.two_sides) {
        const size_t a2 = (size_t)(F.node2[k] - 1);
        j0 = (double)fld[a2] - (double)fld[a1];
        if (ndof == 2) j1 = (double)fld[a2 + npoin] - (double)fld[a1 + npoin];
      } else {
        j0 = -2.0 * (double)fld[a1];
        if (ndof == 2) j1 = -2.0 * (double)fld[a1 + npoin];
      }
      if (ndof == 2) {  // bc_dynflt.f90:843-847
        acc[3 * w + 0] += nx * j0 * B;
        acc[3 * w + 1] += nz * j1 * B;
        acc[3 * w + 2] += (nx * j1 + nz * j0) * B;
      } else {  // :849-851
        acc[2 * w + 0] += nx * j0 * B;
        acc[2 * w + 1] += nz * j0 * B;
      }
    }
  }
  for (int q = 0; q < 6; ++q) red[q][t] = acc[q];
  __syncthreads();
  for (int s = DYNW_THREADS / 2; s > 0; s >>= 1) {
    if (t < s)
      for (int q = 0; q < 6; ++q) red[q][t] += red[q][t + s];
    __syncthreads();
  }
  const int oit = F.ostate[0], nout = F.ostate[1], ncall = F.ostate[2];
  const bool out = (ctl->it >= oit) && nout < F.nrec_max;
  if (out) {
    float* r = F.records + (size_t)nout * 6 * F.onx;
    for (int m = blockIdx.x * DYNW_THREADS + t; m < F.onx; m += stride) {
      const int k = F.oix1 - 1 + m * F.oixd;
      r[m] = (float)F.D[k];
      r[F.onx + m] = (float)F.V[k];
      r[2 * F.onx + m] = (float)F.T[k];
      r[3 * F.onx + m] = (float)F.T[k + np];
      r[4 * F.onx + m] = (float)F.MU[k];
      r[5 * F.onx + m] = (float)F.Tstick[k];
    }
  }
  if (t < 6) part[blockIdx.x * 6 + t] = red[t][0];
  __threadfence();
  __syncthreads();
  if (t == 0) last = (atomicAdd(ticket, 1u) == (unsigned)nfb - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (t == 0) {
    double tot[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nfb; ++b)
      for (int q = 0; q < 6; ++q) tot[q] += __ldcg(&part[b * 6 + q]);
    const int npot = 2 * (ndof + 1);
    if (ncall < F.ncall_max) {
      double* p = F.potency + (size_t)ncall * npot;
      if (ndof == 2) {
        p[0] = tot[0]; p[1] = tot[1]; p[2] = 0.5 * tot[2];
        p[3] = tot[3]; p[4] = tot[4]; p[5] = 0.5 * tot[5];
      } else {
        p[0] = 0.5 * tot[0]; p[1] = 0.5 * tot[1];
        p[2] = 0.5 * tot[2]; p[3] = 0.5 * tot[3];
      }
    }
    F.ostate[2] = min(ncall + 1, F.ncall_max);  // calls beyond nt_max are not recorded (get_fault copies ostate[2] rows)
    if (ctl->it >= oit) {
      F.ostate[0] = oit + F.oitd;
      F.ostate[1] = nout + (out ? 1 : 0);
    }
    *ticket = 0;
  }
}

// ------------------------------------------------------------------------------------------
// reductions: max|x| over an array (two launches), sum(m*|v|^2)
template <typename T>
__global__ void __launch_bounds__(256) k_absmax(const T* __restrict__ x, size_t n, double* partial) {
  __shared__ double red[256];
  double m = 0.0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    m = fmax(m, fabs((double)x[q]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
template <typename T>
__global__ void __launch_bounds__(256) k_kinetic(const T* __restrict__ v, const double* __restrict__ mass,
                                                 size_t npoin, int ndof, double* partial) {
  __shared__ double red[256];
  double m = 0.0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < npoin; q += stride) {
    double v2 = 0.0;
    for (int c = 0; c < ndof; ++c) v2 += (double)v[q + npoin * c] * (double)v[q + npoin * c];
    m += mass[q] * v2;
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

}  // namespace s2d
