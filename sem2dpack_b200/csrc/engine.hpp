// Engine core: device-resident state of one SEM2DPACK problem and the per-step launch sequence
// of solve_leapfrog / solve_Newmark (SRC/solver.f90:42-84,140-160) + REC_store + BC_write.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <memory>

#include "bc_kernels.cuh"
#include "common.cuh"
#include "elem_kernels.cuh"
#include "plan.hpp"
#include "strip_kernels.cuh"
#include "structured.hpp"

namespace s2d {

inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

enum BcKind { BC_ABSO = 1, BC_DIRNEU = 2, BC_DYNFLT = 3 };

struct AbsoBc {
  AbsoDev dev{};
  DevBuf<int> node, st_start, st_elem, st_loc, bibool;
  DevBuf<double> C, n, K, Ht;
};
struct DirneuBc {
  int np = 0, kind_h = 1, kind_v = 1, slot = 0;
  DevBuf<int> node;
  DevBuf<double> B_h, B_v;
};
struct FaultBc {
  FaultDev dev{};
  DevBuf<int> node1, node2, ostate;
  DevBuf<double> n1, B, invM1, invM2, Z, T0, cohesion, coord, T, Tstick, V, D, MU, sigma;
  DevBuf<double> swf_dc, swf_mus, swf_mud, swf_p, swf_alpha, swf_theta;
  DevBuf<double> rsf_dc, rsf_mus, rsf_a, rsf_b, rsf_Vstar, rsf_Vc, rsf_Tc, rsf_coeft, rsf_theta;
  DevBuf<float> records;
  DevBuf<double> potency, wpart;
  DevBuf<unsigned> wticket;
  int wctas = 1;
};
struct BcRef {
  int kind;
  int index;
};

struct Receivers {
  bool present = false;
  char field = 'V';
  RecDev dev{};
  DevBuf<int> iglob, einterp, nodes;
  DevBuf<double> interp;
  DevBuf<float> sis;
};

class EngineBase {
 public:
  virtual ~EngineBase() {}
  int ngll = 0, ndof = 0, nelem = 0, prec = 8;
  size_t npoin = 0;      // device node count = component stride (builder-made engines: lattice rows of pitch LXP)
  size_t npoin_ref = 0;  // node count of the caller's numbering (what host arrays are sized with)
  s2d_scheme scheme{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool committed = false;
  int variant = S2D_ASM_PATCH;
  int it = 0;  // host mirror of ctl.it
  int64_t launches = 0;

  virtual void set_elastic(int nelast, int ncoefsets, const double* a, const int32_t* elem2set, const double* beta25d,
                           int kd2) = 0;
  virtual void set_kv(int nkv, const int32_t* elem_ids, const double* eta) = 0;
  virtual void set_mass(const double* mass) = 0;
  virtual void add_abso(int np, const int32_t* node, const double* C, int is_flat, const double* n,
                        int stacey, int nbe, const int32_t* bibool, const double* K) = 0;
  virtual void add_dirneu(int np, const int32_t* node, int kind_h, int kind_v, const double* B_h,
                          const double* B_v) = 0;
  virtual int add_dynflt(const s2d_dynflt_desc& d) = 0;
  virtual void add_periodic(int np, const int32_t* master, const int32_t* slave) = 0;
  virtual int add_force(int iglob, const double dir[2]) = 0;
  virtual int add_moment(int nterms, const int32_t* node, const double* coef) = 0;
  virtual void add_receivers(int nx, char field, int isamp, int nt_rec, int at_node, const int32_t* iglob,
                             const int32_t* einterp, const double* interp) = 0;
  virtual void add_receivers_nodes(int nx, char field, int isamp, int nt_rec, const int32_t* nodes, const double* interp) = 0;
  virtual void set_node_kv(const double* eta_ref) = 0;
  virtual void set_strip_eta(const double* eta_strip, size_t n) = 0;
  virtual void set_strip_plastic(const unsigned char* set_strip, size_t n, int nsets, const double* raw) = 0;
  virtual void get_strip_plastic_strain(double* ep_strip) = 0;
  virtual void set_strip_damage(const unsigned char* set_strip, size_t n, int nsets, const double* par, const double* state_strip) = 0;
  virtual void get_strip_damage_state(double* state_strip) = 0;
  virtual void set_strip_visco(const unsigned char* set_strip, size_t n, int nsets, const int32_t* nbody, const double* moduli,
                               const double* wbody, const double* theta) = 0;
  virtual void commit(int variant) = 0;
  virtual void set_fields(const double* d, const double* v, const double* a) = 0;
  virtual void get_fields(double* d, double* v, double* a) = 0;
  virtual void step(int nsteps, const double* src_ampli, const double* bc_ampli) = 0;
  virtual void compute_fint(double* fint) = 0;
  virtual void get_seis(float* sis) = 0;
  virtual void get_seis_row(int it, float* row) = 0;
  virtual void get_fault(int id, float* records, int32_t* nout, double* potency, int32_t* ncalls) = 0;
  virtual void get_fault_state(int id, double* D, double* V, double* T, double* Tstick, double* MU,
                               double* theta, double* sigma) = 0;
  virtual void progress(double* vmax, double* dmax) = 0;
  virtual double energy() = 0;
  virtual double energy_w25d() = 0;
  virtual void get_coloring(int32_t* ncolors, int32_t* color) = 0;
  virtual float time_fint(int reps) = 0;
  virtual float time_steps(int nsteps) = 0;
  virtual void time_phases(int nsteps, float* ms) = 0;
  float last_kernel_ms_base = 0.f;
  virtual float kernel_ms() = 0;
  virtual int route() = 0;
  virtual void halo_info(int64_t* count, void** send_dev, void** recv_dev) = 0;
  virtual void halo_set_exchange(s2d_exchange_fn fn, void* user) = 0;
  virtual void halo_peer_buffers(void** recv_dev, void** flags_dev) = 0;
  virtual void halo_set_peers(void* left_recv, void* right_recv, void* left_flag, void* right_flag) = 0;
};

template <typename T>
class Engine : public EngineBase {
 public:
  // grid
  std::vector<int32_t> h_ibool;  // kept on the host for planning
  DevBuf<int> ibool;
  DevBuf<T> H;
  std::vector<double> h_H;
  // fields
  DevBuf<T> d, v, a, rmass, scratch;
  // material
  int nelast = 0, ncoefsets = 0, kd2 = 0, nkv = 0;
  DevBuf<T> coef, eta, beta;      // beta: 2.5D term (N*N, ncoefsets), empty when W is infinite
  std::vector<double> h_beta;
  DevBuf<T> strip_beta;           // ... per element GLL point in the strip layout (strip kernel)
  DevBuf<int> elem2set, elem2kv;
  std::vector<int32_t> h_elem2set, h_elem2kv;
  std::vector<double> h_coef;
  DevBuf<double> mass;
  // boundary conditions in registration order
  std::vector<std::unique_ptr<AbsoBc>> abso;
  std::vector<std::unique_ptr<DirneuBc>> dirneu;
  std::vector<std::unique_ptr<FaultBc>> faults;
  std::vector<BcRef> bc_order;
  // periodic boundaries: pairs of nodes whose forces are summed first (bc_gen.f90:271-275)
  struct PerioBc {
    int np = 0;
    DevBuf<int> master, slave;
  };
  std::vector<std::unique_ptr<PerioBc>> perio;
  int n_neumann_slots = 0;
  // sources
  std::vector<int32_t> h_src_iglob;
  std::vector<double> h_src_dir;
  DevBuf<double> src_ampli, bc_ampli;
  // source terms grouped by target node, in the order SO_add applies them (k_sources)
  DevBuf<int> st_node, st_start, st_src;
  DevBuf<double> st_coef;
  int st_nnodes = 0, st_nterms = 0;
  // moment-tensor sources (src_moment.f90): terms of SRC_MOMENT_add per source
  std::vector<int32_t> h_mom_src, h_mom_start{0}, h_mom_node;
  std::vector<double> h_mom_coef;
  // HHT-alpha work fields (fields%displ_alpha, veloc_alpha; fields.f90:10-11)
  DevBuf<T> d_alpha, v_alpha;
  int nstages() const { return scheme.kind == 3 ? scheme.nstages : 1; }
  size_t src_ampli_cap = 0, bc_ampli_cap = 0;
  // pinned host staging of the per-call inputs / outputs (the caller's buffers are ordinary host memory)
  struct Pinned {
    void* p = nullptr;
    size_t bytes = 0;
    void* need(size_t n) {
      if (n > bytes) {
        if (p) cudaFreeHost(p);
        p = nullptr;
        if (cudaMallocHost(&p, n) != cudaSuccess) throw StateError("cudaMallocHost failed");
        bytes = n;
      }
      return p;
    }
    ~Pinned() {
      if (p) cudaFreeHost(p);
    }
  } pin_src, pin_bc, pin_row;
  Receivers rec;
  // control
  DevBuf<StepCtl> ctl;
  DevBuf<double> partial;
  // colour plan
  int ncolors = 0;
  std::vector<int32_t> h_color;
  std::vector<int32_t> color_start;  // (ncolors+1)
  DevBuf<int> color_elems;
  // patch plan (device tables of the CTA-patch kernel)
  PatchPlan plan;
  int pp_npatch = 0, pp_EP = 0, pp_max_nloc = 0, pp_max_colors = 0;
  size_t pp_nslots = 0;
  DevBuf<int> p_pelem_start, p_pshape, p_sh_slot, p_pnode, p_eset, p_ekv, p_snode, p_sstart;
  DevBuf<long long> p_pslot_base, p_pnode_start;
  DevBuf<uint16_t> p_sh_lidx;
  DevBuf<uint8_t> p_sh_ecolor;
  DevBuf<T> p_coef, fhalo;
  bool p_hetero = false;
  int pf_dist = env_int("S2D_PF_DIST", 444);  // L2 software-prefetch distance of the patch kernel
  int use_bulk = env_int("S2D_BULK", 1);      // TMA bulk copy of the coefficient block into shared memory
  // structured builder (cart.cu): fields live on the GLL lattice, forces come from the z-marching
  // strip kernel; the hooks translate to / from the caller's node numbering at the API boundary
  bool cart_mode = false;
  StripGeom cart_S{};
  DevBuf<T> cart_hx, cart_hz;
  DevBuf<int> cart_meet;  // arrival counters of the group-boundary columns (strip_kernels.cuh)
  // Kelvin-Voigt damping on a structured box: eta (already * dt) is a function of position, so the
  // element-wise d_loc + eta*v_loc of MAT_KV_add_etav (mat_kelvin_voigt.f90:137-150) is the node field d + eta*v
  DevBuf<T> cart_kv_eta;  // node-wise eta as s2d_cart_set_kv hands it over, until commit spreads it over the elements
  // compact coefficient mode: p_coef holds (lambda, mu) per GLL point, the strip kernel forms the planes
  int cart_compact = 0;
  double cart_cdx = 0.0, cart_cdz = 0.0, cart_cdet = 0.0;
  std::vector<double> cart_wgll;
  std::function<void(const T*, double*)> cart_to_ref;    // lattice (T) -> reference numbering (FP64), device to device
  std::function<void(const double*, T*)> cart_from_ref;  // reference numbering (FP64) -> lattice (T)
  std::function<void(const double*, T*)> cart_from_ref1; // the same for a one-component node array
  std::function<void(const double*, double*)> cart_from_ref1d;  // ... kept in FP64 (the mass of s2d_energy)
  // fused leapfrog step of the strip kernel: two displacement buffers (the kernel reads d[n] and
  // writes the predicted d[n+1] of the next step), deferred-node tables (strip_kernels.cuh)
  DevBuf<T> d2;
  int dsel = 0;             // which buffer holds the displacement the caller sees
  bool pred_valid = false;  // the other buffer holds d + dt*v of the next step
  bool fused = false;
  // accelerations of the fused step: 1 = written every step, 0 = never, 2 (default) = when somebody can
  // see them: on the last step of every s2d_step call, and on every step if receivers record 'A'
  int store_accel = env_int("S2D_STORE_ACCEL", 2);
  int strip_prefetch = env_int("S2D_STRIP_PF", 1);
  int strip_occ = env_int("S2D_STRIP_OCC", 0);
  // tensor-map staging of the compact fused leapfrog kernel (S2D_STRIP_TENSOR=0 in the environment: per-lane copies)
  int strip_tensor = env_int("S2D_STRIP_TENSOR", S2D_STRIP_TENSOR);
  bool tm_ready = false;
  int strip_tensor_plain = env_int("S2D_STRIP_TENSOR_PLAIN", 1);
  CUtensorMap tm_d[2], tm_v, tm_r, tm_a;   // d[n] is in either displacement buffer
  std::vector<uint8_t> h_rowflag, h_colflag;
  std::vector<std::vector<int32_t>> h_bc_nodes;  // node lists of every boundary condition (for the flags)
  DevBuf<uint8_t> rowflag, colflag;
  DevBuf<int> drows, dcols;
  int ndrows = 0, ndcols = 0;
  T* dn() { return dsel ? d2.p : d.p; }
  T* dalt() { return dsel ? d.p : d2.p; }
  DevBuf<T>& dbuf() { return dsel ? d2 : d; }
  // per-launch timing of the dominant kernel inside s2d_time_steps
  std::vector<cudaEvent_t> kev;
  bool kev_on = false;
  size_t kev_n = 0;
  float last_kernel_ms = 0.f;
  // phase marks of s2d_time_phases: the time between two consecutive marks on the engine stream is charged to
  // the phase of the first one.  0 tick / predictor, 1 element-force kernel, 2 halo folds + interface exchange,
  // 3 sources, 4 boundary conditions, 5 node update (deferred nodes / corrector), 6 outputs, 7 idle
  enum { PH_PRED = 0, PH_FORCE, PH_FOLD, PH_SRC, PH_BC, PH_UPDATE, PH_OUT, PH_COUNT };
  std::vector<cudaEvent_t> pev;
  std::vector<int> pev_phase;
  bool pev_on = false;
  void phase(int ph) {
    if (!pev_on) return;
    cudaEvent_t e;
    if (pev_phase.size() < pev.size()) {
      e = pev[pev_phase.size()];
    } else {
      S2D_CUDA(cudaEventCreate(&e));
      pev.push_back(e);
    }
    S2D_CUDA(cudaEventRecord(e, stream));
    pev_phase.push_back(ph);
  }
  void kev_mark() {
    if (!kev_on) return;
    if (kev_n >= kev.size()) {
      cudaEvent_t e;
      S2D_CUDA(cudaEventCreate(&e));
      kev.push_back(e);
    }
    S2D_CUDA(cudaEventRecord(kev[kev_n++], stream));
  }
  // x-strip interfaces with neighbour GPUs: partial sums of lattice column 0 / LX-1 are
  // packed right after the two boundary strips are done, exchanged on a side stream while the
  // interior strips are still computing, and added as (own + neighbour's) on both sides
  DevBuf<T> xh_send[2], xh_recv[2];
  s2d_exchange_fn xh_fn = nullptr;
  void* xh_user = nullptr;
  // direct peer-memory exchange: the neighbours' receive slots and flags (device pointers that are
  // valid on this GPU: CUDA IPC mappings or peer-enabled allocations), my own flags, the evaluation count
  bool xh_peer = false;
  T* xh_peer_recv[2] = {nullptr, nullptr};
  unsigned long long* xh_peer_flag[2] = {nullptr, nullptr};
  DevBuf<unsigned long long> xh_flags;
  unsigned long long xh_seq = 0;
  cudaStream_t xstream = nullptr;
  cudaEvent_t xh_ev_b = nullptr, xh_ev_x = nullptr;
  bool xhalo() const { return cart_mode && (cart_S.xhalo_left || cart_S.xhalo_right); }
  void xhalo_setup() {
    if (!xhalo() || xstream) return;
    const size_t n = (size_t)cart_S.LZ * ndof;
    for (int sd = 0; sd < 2; ++sd)
      if (sd ? cart_S.xhalo_right : cart_S.xhalo_left) {
        xh_send[sd].alloc(n);
        xh_recv[sd].alloc(2 * n);  // two slots for the peer-memory exchange; the hook exchange uses the first
        xh_send[sd].zero();
        xh_recv[sd].zero();
      }
    xh_flags.alloc(2);
    xh_flags.zero();
    S2D_CUDA(cudaStreamCreateWithFlags(&xstream, cudaStreamNonBlocking));
    S2D_CUDA(cudaEventCreateWithFlags(&xh_ev_b, cudaEventDisableTiming));
    S2D_CUDA(cudaEventCreateWithFlags(&xh_ev_x, cudaEventDisableTiming));
  }
  StripIO<T> strip_io(const T* dd, T* ff) {
    StripIO<T> io;
    io.coef = p_coef.p;
    io.d = dd;
    io.f = ff;
    io.halo_x = cart_hx.p;
    io.halo_z = cart_hz.p;
    io.meet = cart_meet.p;
    io.npoin = npoin;
    io.hprime = h_H.data();
    io.prefetch = strip_prefetch;
    io.compact = cart_compact;
    io.occ = strip_occ;
    io.cdx = cart_cdx;
    io.cdz = cart_cdz;
    io.cdet = cart_cdet;
    io.wgll = cart_wgll.empty() ? nullptr : cart_wgll.data();
    io.beta = strip_beta.p;
    if (pl_set.n) {
      io.pl_set = pl_set.p;
      io.pl_ep = pl_ep.p;
      io.pl_tab = pl_tab.p;
      io.vs_state = vs_state.n ? vs_state.p : nullptr;
      io.vs_tab = vs_tab.p;
      io.vs_nb = vs_nb;
      io.dm_state = dm_state.n ? dm_state.p : nullptr;
      io.dm_tab = dm_tab.p;
      io.dm_err = &ctl.p->err;
    }
    return io;
  }
  // plain force evaluation of one of the two displacement buffers: the kernel can take its rows as tensor-map boxes
  void plain_tensor_map(StripIO<T>& io, const T* dd) {
    if (!tm_ready || !strip_tensor_plain) return;
    if (dd == d.p) io.tm_d = &tm_d[0];
    else if (dd == d2.p) io.tm_d = &tm_d[1];
  }
  // strip kernel over the whole box (+ halo fold, + interface exchange); io.v_in != null = fused update
  // tick: the step counter is advanced by the fold kernel (fused step: one launch less)
  void launch_strips(const StripIO<T>& io, StepCtl* tick = nullptr) {
    const StripGeom& S0 = cart_S;
    if (!xhalo()) {
      kev_mark();
      phase(PH_FORCE);
      launch_elem_strip_items<T>(strip_all_groups(S0), io, stream);
      kev_mark();
      phase(PH_FOLD);
      launches += 1 + launch_strip_fold<T>(S0, io.f, cart_hx.p, cart_hz.p, npoin, stream, tick);
      return;
    }
    phase(PH_FORCE);
    if (!xh_fn && !xh_peer)
      throw StateError("this x-strip has neighbours: attach a halo exchange (s2d_halo_set_exchange or s2d_halo_set_peers) first");
    // boundary groups = the single-strip groups next to the interfaces
    const int nb = S0.g_lead + S0.g_tail + ((S0.nstrips == 1) ? 1 : 0);
    StripGeom B = S0, I = S0;
    B.it_g0 = S0.g_lead ? 0 : S0.ngroups - 1;
    B.it_step = S0.ngroups - 1;
    B.it_ng = nb;
    if (nb == 1) B.it_step = 0;
    B.nitems = (long long)S0.nseg * nb;
    I.it_g0 = S0.g_lead;
    I.it_ng = S0.ngroups - nb;
    I.it_step = 1;
    I.nitems = (long long)S0.nseg * I.it_ng;
    T* ff = io.f;
    launch_elem_strip_items<T>(B, io, stream);
    S2D_CUDA(cudaEventRecord(xh_ev_b, stream));
    launches++;
    if (I.nitems > 0) {
      kev_mark();
      launch_elem_strip_items<T>(I, io, stream);
      kev_mark();
      launches++;
    }
    phase(PH_FOLD);
    S2D_CUDA(cudaStreamWaitEvent(xstream, xh_ev_b, 0));
    const int n2 = 2 * S0.LZ * ndof;
    const size_t nh = (size_t)S0.LZ * ndof;
    const T *rl = xh_recv[0].p, *rr = xh_recv[1].p;
    if (xh_peer) {  // write my partial sums into the neighbours' slots over NVLink, then raise their flags
      ++xh_seq;
      const size_t slot = (size_t)(xh_seq & 1) * nh;
      k_xhalo_pack<T><<<ceil_div(n2, 256), 256, 0, xstream>>>(S0, ff, cart_hz.p, npoin,
                                                             xh_peer_recv[0] ? xh_peer_recv[0] + slot : nullptr,
                                                             xh_peer_recv[1] ? xh_peer_recv[1] + slot : nullptr);
      k_xhalo_signal<<<1, 1, 0, xstream>>>(xh_peer_flag[0], xh_peer_flag[1], xh_seq);
      launches += 2;
      if (rl) rl += slot;
      if (rr) rr += slot;
    } else {
      k_xhalo_pack<T><<<ceil_div(n2, 256), 256, 0, xstream>>>(S0, ff, cart_hz.p, npoin, xh_send[0].p, xh_send[1].p);
      launches++;
      const int rc = xh_fn(xh_user, (void*)xstream);
      if (rc != 0) throw StateError("halo exchange hook failed with code " + std::to_string(rc));
    }
    S2D_CUDA(cudaEventRecord(xh_ev_x, xstream));
    launches += launch_strip_fold<T>(S0, ff, cart_hx.p, cart_hz.p, npoin, stream, tick);
    S2D_CUDA(cudaStreamWaitEvent(stream, xh_ev_x, 0));
    if (xh_peer) {
      k_xhalo_wait<<<1, 1, 0, stream>>>(xh_flags.p, S0.xhalo_left, S0.xhalo_right, xh_seq, ctl.p);
      launches++;
    }
    k_xhalo_unpack<T><<<ceil_div(n2, 256), 256, 0, stream>>>(S0, ff, cart_hz.p, npoin, rl, rr);
    launches++;
    S2D_CUDA(cudaGetLastError());
  }

  // bare engine for the structured builder (cart.cu fills the tables on the device)
  struct Raw {};
  Engine(Raw, int ngll_, int ndof_, int nelem_, size_t npoin_, size_t npoin_ref_, const double* hprime,
         const s2d_scheme& sch, int dev) {
    ngll = ngll_;
    ndof = ndof_;
    nelem = nelem_;
    npoin = npoin_;
    npoin_ref = npoin_ref_;
    prec = (int)sizeof(T);
    scheme = sch;
    device = dev;
    cart_mode = true;
    S2D_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    const size_t n2 = (size_t)ngll * ngll;
    h_H.assign(hprime, hprime + n2);
    upload_as(H, hprime, n2);
    const size_t nd = npoin * ndof;
    d.alloc(nd);
    v.alloc(nd);
    a.alloc(nd);
    d.zero();
    v.zero();
    a.zero();
    rmass.alloc(nd);
    StepCtl c0{0, 1, 0, 1};
    ctl.upload(&c0, 1);
    partial.alloc(1024);
  }

  Engine(int ngll_, int ndof_, int nelem_, size_t npoin_, const int32_t* ibool_, const double* hprime,
         const double* rmass_, const s2d_scheme& sch, int dev) {
    ngll = ngll_;
    ndof = ndof_;
    nelem = nelem_;
    npoin = npoin_;
    npoin_ref = npoin_;
    prec = (int)sizeof(T);
    scheme = sch;
    device = dev;
    S2D_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    const size_t n2 = (size_t)ngll * ngll;
    h_ibool.assign(ibool_, ibool_ + n2 * nelem);
    for (size_t q = 0; q < h_ibool.size(); ++q)
      S2D_REQUIRE(h_ibool[q] >= 1 && (size_t)h_ibool[q] <= npoin, "ibool entry out of range");
    ibool.upload(h_ibool);
    h_H.assign(hprime, hprime + n2);
    upload_as(H, hprime, n2);
    const size_t nd = npoin * ndof;
    d.alloc(nd);
    v.alloc(nd);
    a.alloc(nd);
    d.zero();
    v.zero();
    a.zero();
    upload_as(rmass, rmass_, nd);
    StepCtl c0{0, 1, 0, 1};
    ctl.upload(&c0, 1);
    partial.alloc(1024);
  }
  ~Engine() override {
    for (auto e : kev) cudaEventDestroy(e);
    for (auto e : pev) cudaEventDestroy(e);
    if (xh_ev_b) cudaEventDestroy(xh_ev_b);
    if (xh_ev_x) cudaEventDestroy(xh_ev_x);
    if (xstream) cudaStreamDestroy(xstream);
    if (stream) cudaStreamDestroy(stream);
  }

  // ---- configuration -------------------------------------------------------------------
  void set_elastic(int nelast_, int ncoefsets_, const double* a_, const int32_t* e2s, const double* beta_,
                   int kd2_) override {
    S2D_REQUIRE(!committed, "set_elastic after commit");
    S2D_REQUIRE((ndof == 1 && (nelast_ == 2 || nelast_ == 3)) || (ndof == 2 && (nelast_ == 6 || nelast_ == 10)),
                "nelast does not match ndof (mat_elastic.f90:255-268)");
    S2D_REQUIRE(ncoefsets_ >= 1, "ncoefsets < 1");
    nelast = nelast_;
    ncoefsets = ncoefsets_;
    kd2 = kd2_ ? 1 : 0;
    const size_t n2 = (size_t)ngll * ngll;
    h_coef.assign(a_, a_ + n2 * nelast * ncoefsets);
    upload_as(coef, a_, h_coef.size());
    h_beta.clear();
    beta.release();
    if (beta_) {  // matwrk_elast_type%beta (mat_elastic.f90:280-284): finite seismogenic width W
      h_beta.assign(beta_, beta_ + n2 * ncoefsets);
      upload_as(beta, beta_, h_beta.size());
    }
    h_elem2set.resize(nelem);
    for (int e = 0; e < nelem; ++e) {
      S2D_REQUIRE(e2s[e] >= 1 && e2s[e] <= ncoefsets, "elem2set entry out of range");
      h_elem2set[e] = e2s[e] - 1;
    }
    elem2set.upload(h_elem2set);
  }
  void set_kv(int nkv_, const int32_t* elem_ids, const double* eta_) override {
    S2D_REQUIRE(!committed, "set_kv after commit");
    nkv = nkv_;
    h_elem2kv.assign(nelem, -1);
    for (int k = 0; k < nkv; ++k) {
      S2D_REQUIRE(elem_ids[k] >= 1 && elem_ids[k] <= nelem, "KV element id out of range");
      h_elem2kv[elem_ids[k] - 1] = k;
    }
    if (nkv > 0) {
      elem2kv.upload(h_elem2kv);
      upload_as(eta, eta_, (size_t)ngll * ngll * nkv);
      h_eta.assign(eta_, eta_ + (size_t)ngll * ngll * nkv);
    }
  }
  // mass(npoin) in the caller's numbering; builder-made engines keep it on the lattice like every field
  // (the builder also fills it itself, before the boundary conditions touch the mass: cart.cu)
  void set_mass(const double* m) override {
    if (!cart_mode) {
      mass.upload(m, npoin);
      return;
    }
    DevBuf<double> tmp;
    tmp.upload(m, npoin_ref);
    mass.alloc(npoin);
    mass.zero(stream);
    cart_from_ref1d(tmp.p, mass.p);
    S2D_CUDA(cudaStreamSynchronize(stream));
  }

  void check_nodes(int np, const int32_t* node, const char* what) {
    for (int k = 0; k < np; ++k)
      if (node[k] < 1 || (size_t)node[k] > npoin) throw ArgError(std::string(what) + ": node id out of range");
  }

  void add_abso(int np, const int32_t* node, const double* C, int is_flat, const double* n, int stacey,
                int nbe, const int32_t* bibool, const double* K) override {
    S2D_REQUIRE(!committed, "add_abso after commit");
    S2D_REQUIRE(np > 0 && node && C, "add_abso: empty boundary");
    check_nodes(np, node, "add_abso");
    auto b = std::make_unique<AbsoBc>();
    h_bc_nodes.emplace_back(node, node + np);
    b->node.upload(node, np);
    b->C.upload(C, (size_t)np * ndof);
    if (!is_flat && ndof == 2) {
      S2D_REQUIRE(n != nullptr, "add_abso: non-flat boundary needs normals");
      b->n.upload(n, (size_t)np * 2);
    }
    AbsoDev& A = b->dev;
    A.np = np;
    A.ndof = ndof;
    A.is_flat = is_flat;
    A.stacey = (stacey && ndof == 2) ? 1 : 0;
    A.ngll = ngll;
    if (A.stacey) {
      S2D_REQUIRE(nbe > 0 && bibool && K, "add_abso: Stacey needs bibool and K");
      std::vector<int> start(np + 1, 0), el, lc;
      for (size_t q = 0; q < (size_t)ngll * nbe; ++q) {
        S2D_REQUIRE(bibool[q] >= 1 && bibool[q] <= np, "add_abso: bibool out of range");
        start[bibool[q]]++;
      }
      for (int k = 0; k < np; ++k) start[k + 1] += start[k];
      el.resize(start[np]);
      lc.resize(start[np]);
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (int e = 0; e < nbe; ++e)  // ascending element, then local index: the reference's order
        for (int i = 0; i < ngll; ++i) {
          const int bn = bibool[i + (size_t)ngll * e] - 1;
          el[fill[bn]] = e;
          lc[fill[bn]] = i;
          fill[bn]++;
        }
      b->st_start.upload(start);
      b->st_elem.upload(el);
      b->st_loc.upload(lc);
      b->bibool.upload(bibool, (size_t)ngll * nbe);
      b->K.upload(K, (size_t)ngll * 2 * nbe);
      std::vector<double> Ht((size_t)ngll * ngll);
      for (int i = 0; i < ngll; ++i)
        for (int k = 0; k < ngll; ++k) Ht[i + ngll * k] = h_H[k + ngll * i];
      b->Ht.upload(Ht);
    }
    A.node = b->node.p;
    A.C = b->C.p;
    A.n = b->n.p;
    A.st_start = b->st_start.p;
    A.st_elem = b->st_elem.p;
    A.st_loc = b->st_loc.p;
    A.bibool = b->bibool.p;
    A.K = b->K.p;
    A.Ht = b->Ht.p;
    abso.push_back(std::move(b));
    bc_order.push_back({BC_ABSO, (int)abso.size() - 1});
  }

  void add_dirneu(int np, const int32_t* node, int kind_h, int kind_v, const double* B_h,
                  const double* B_v) override {
    S2D_REQUIRE(!committed, "add_dirneu after commit");
    S2D_REQUIRE(np > 0 && node, "add_dirneu: empty boundary");
    check_nodes(np, node, "add_dirneu");
    auto b = std::make_unique<DirneuBc>();
    h_bc_nodes.emplace_back(node, node + np);
    b->np = np;
    b->kind_h = kind_h;
    b->kind_v = kind_v;
    b->node.upload(node, np);
    if (B_h) b->B_h.upload(B_h, np);
    if (B_v) b->B_v.upload(B_v, np);
    b->slot = n_neumann_slots;
    n_neumann_slots += 2;
    dirneu.push_back(std::move(b));
    bc_order.push_back({BC_DIRNEU, (int)dirneu.size() - 1});
  }

  int add_dynflt(const s2d_dynflt_desc& D) override {
    S2D_REQUIRE(!committed, "add_dynflt after commit");
    const int np = D.np;
    S2D_REQUIRE(np > 0 && D.node1 && D.n1 && D.B && D.invM1 && D.Z && D.T0 && D.cohesion && D.coord,
                "add_dynflt: missing array");
    check_nodes(np, D.node1, "add_dynflt");
    if (D.node2) check_nodes(np, D.node2, "add_dynflt");
    S2D_REQUIRE(!D.node2 || D.invM2, "add_dynflt: two-sided fault needs invM2");
    S2D_REQUIRE(!(D.swf_kind && D.rsf_kind), "add_dynflt: SWF and RSF are exclusive (bc_dynflt.f90:168-183)");
    auto b = std::make_unique<FaultBc>();
    FaultDev& F = b->dev;
    std::memset(&F, 0, sizeof(F));
    F.np = np;
    F.ndof = ndof;
    F.two_sides = D.node2 ? 1 : 0;
    F.allow_opening = D.allow_opening;
    F.CoefA2V = D.CoefA2V;
    F.CoefA2D = D.CoefA2D;
    F.dt = scheme.dt;
    F.tshift = scheme.kind == 2 ? (scheme.alpha - 1.0) * scheme.dt : 0.0;
    auto up = [&](DevBuf<double>& buf, const double* src, size_t n) -> const double* {
      if (!src) return nullptr;
      buf.upload(src, n);
      return buf.p;
    };
    b->node1.upload(D.node1, np);
    F.node1 = b->node1.p;
    h_fault_node1.push_back(D.node1[0]);
    h_bc_nodes.emplace_back(D.node1, D.node1 + np);
    if (D.node2) h_bc_nodes.emplace_back(D.node2, D.node2 + np);
    if (D.node2) {
      b->node2.upload(D.node2, np);
      F.node2 = b->node2.p;
    }
    F.n1 = up(b->n1, D.n1, (size_t)np * 2);
    F.B = up(b->B, D.B, (size_t)np * ndof);
    F.invM1 = up(b->invM1, D.invM1, (size_t)np * ndof);
    F.invM2 = up(b->invM2, D.invM2, (size_t)np * ndof);
    F.Z = up(b->Z, D.Z, (size_t)np * ndof);
    F.T0 = up(b->T0, D.T0, (size_t)np * 2);
    F.cohesion = up(b->cohesion, D.cohesion, np);
    F.coord = up(b->coord, D.coord, (size_t)np * 2);
    std::vector<double> zero((size_t)np * 2, 0.0);
    b->T.upload(zero);
    b->Tstick.upload(zero);
    b->D.upload(zero.data(), (size_t)np * ndof);
    if (D.V0)
      b->V.upload(D.V0, (size_t)np * ndof);
    else
      b->V.upload(zero.data(), (size_t)np * ndof);
    F.T = b->T.p;
    F.Tstick = b->Tstick.p;
    F.D = b->D.p;
    F.V = b->V.p;
    // normal stress (bc_dynflt_normal.f90:96-114)
    F.normal_kind = D.normal_kind;
    F.normal_V = D.normal_V;
    F.normal_coef = 0.0;
    if (D.normal_kind == 2) F.normal_coef = std::exp(-scheme.dt / D.normal_T);
    if (D.normal_kind == 3) F.normal_coef = scheme.dt / D.normal_L;
    std::vector<double> sigma(D.T0 + np, D.T0 + 2 * (size_t)np);
    b->sigma.upload(sigma);
    F.sigma = b->sigma.p;
    std::vector<double> MU(np, 0.0);
    // friction
    F.swf_kind = D.swf_kind;
    F.swf_healing = D.swf_healing;
    F.rsf_kind = D.rsf_kind;
    F.twf_kind = D.twf_kind;
    F.twf_X = D.twf_X; F.twf_Z = D.twf_Z; F.twf_mus = D.twf_mus; F.twf_mud = D.twf_mud;
    F.twf_mu0 = D.twf_mu0; F.twf_L = D.twf_L; F.twf_V = D.twf_V; F.twf_T = D.twf_T; F.twf_Dc = D.twf_Dc;
    if (D.swf_kind) {
      S2D_REQUIRE(D.swf_dc && D.swf_mus && D.swf_mud && D.swf_p && D.swf_alpha, "add_dynflt: SWF arrays missing");
      F.swf_dc = up(b->swf_dc, D.swf_dc, np);
      F.swf_mus = up(b->swf_mus, D.swf_mus, np);
      F.swf_mud = up(b->swf_mud, D.swf_mud, np);
      F.swf_p = up(b->swf_p, D.swf_p, np);
      F.swf_alpha = up(b->swf_alpha, D.swf_alpha, np);
      if (D.swf_theta)
        b->swf_theta.upload(D.swf_theta, np);
      else
        b->swf_theta.upload(zero.data(), np);
      F.swf_theta = b->swf_theta.p;
    }
    if (D.rsf_kind) {
      S2D_REQUIRE(D.rsf_dc && D.rsf_mus && D.rsf_a && D.rsf_b && D.rsf_Vstar && D.rsf_theta && D.rsf_Vc,
                  "add_dynflt: RSF arrays missing");
      F.rsf_dc = up(b->rsf_dc, D.rsf_dc, np);
      F.rsf_mus = up(b->rsf_mus, D.rsf_mus, np);
      F.rsf_a = up(b->rsf_a, D.rsf_a, np);
      F.rsf_b = up(b->rsf_b, D.rsf_b, np);
      F.rsf_Vstar = up(b->rsf_Vstar, D.rsf_Vstar, np);
      F.rsf_Vc = up(b->rsf_Vc, D.rsf_Vc, np);
      b->rsf_theta.upload(D.rsf_theta, np);
      F.rsf_theta = b->rsf_theta.p;
      std::vector<double> Tc(np), coeft(np);  // rsf_init (bc_dynflt_rsf.f90:152-156)
      for (int k = 0; k < np; ++k) {
        Tc[k] = D.rsf_dc[k] / D.rsf_Vstar[k];
        coeft[k] = std::exp(-scheme.dt / Tc[k]);
      }
      b->rsf_Tc.upload(Tc);
      b->rsf_coeft.upload(coeft);
      F.rsf_Tc = b->rsf_Tc.p;
      F.rsf_coeft = b->rsf_coeft.p;
    }
    // initial friction coefficient (bc_dynflt.f90:404-417) -- evaluated on the host once
    for (int k = 0; k < np; ++k) {
      const double BIG = 1.7976931348623157e308;
      auto twf0 = [&]() {  // twf_mu at time 0, slip 0
        double t, r = 0, mu = BIG;
        const double dist = std::sqrt((D.coord[2 * k] - D.twf_X) * (D.coord[2 * k] - D.twf_X) +
                                      (D.coord[2 * k + 1] - D.twf_Z) * (D.coord[2 * k + 1] - D.twf_Z));
        if (D.twf_kind == 1) {
          t = (D.twf_mus - D.twf_mu0) * D.twf_L / ((D.twf_mus - D.twf_mud) * D.twf_V);
          if (t > D.twf_T) t = 0.0;
          r = D.twf_V * t;
        } else if (D.twf_kind == 2) {
          t = 0.5 * D.twf_T * (1.0 - std::sqrt(1.0 - 4.0 * (D.twf_mus - D.twf_mu0) * D.twf_L /
                                                         ((D.twf_mus - D.twf_mud) * D.twf_T * D.twf_V)));
          t = std::min(t, D.twf_T);
          r = D.twf_V * t * (1.0 - t / D.twf_T);
        }
        if (D.twf_kind == 1 || D.twf_kind == 2) {
          const double rr = dist - r;
          if (rr < -D.twf_L) mu = D.twf_mud;
          else if (rr <= D.twf_L) mu = D.twf_mus + (D.twf_mus - D.twf_mud) / D.twf_L * rr;
        } else {
          if (dist <= D.twf_V * D.twf_T && 0.0 <= D.twf_Dc) {
            if (dist < -D.twf_L) mu = D.twf_mud;
            else if (dist <= 0.0) mu = D.twf_mus + (D.twf_mus - D.twf_mud) / D.twf_L * dist;
          }
        }
        return mu;
      };
      if (D.swf_kind) {
        const double th = D.swf_theta ? D.swf_theta[k] : 0.0;
        double mu = 0;
        if (D.swf_kind == 1) mu = D.swf_mus[k] - (D.swf_mus[k] - D.swf_mud[k]) * std::min(th / D.swf_dc[k], 1.0);
        else if (D.swf_kind == 2) mu = D.swf_mud[k] - (D.swf_mud[k] - D.swf_mus[k]) * std::exp(-th / D.swf_dc[k]);
        else if (D.swf_kind == 3) mu = D.swf_mud[k] + (D.swf_mus[k] - D.swf_mud[k]) / std::pow(1.0 + th / D.swf_dc[k], D.swf_p[k]);
        MU[k] = mu + D.swf_alpha[k] * th;
        if (D.twf_kind) MU[k] = std::min(MU[k], twf0());
      } else if (D.rsf_kind) {
        const double vv = std::fabs(D.V0 ? D.V0[k] : 0.0), th = D.rsf_theta[k];
        if (D.rsf_kind == 1)
          MU[k] = D.rsf_mus[k] + D.rsf_a[k] * vv / (vv + D.rsf_Vstar[k]) - D.rsf_b[k] * th / (th + D.rsf_dc[k]);
        else {
          const double arg = (D.rsf_kind == 4) ? D.rsf_Vc[k] * th / D.rsf_dc[k] + 1.0 : D.rsf_Vstar[k] * th / D.rsf_dc[k];
          MU[k] = D.rsf_a[k] * std::asinh(vv / (2.0 * D.rsf_Vstar[k]) *
                                          std::exp((D.rsf_mus[k] + D.rsf_b[k] * std::log(arg)) / D.rsf_a[k]));
        }
        if (D.twf_kind) MU[k] = std::min(MU[k], twf0());
      } else if (D.twf_kind) {
        MU[k] = twf0();
      }
    }
    b->MU.upload(MU);
    F.MU = b->MU.p;
    // outputs
    F.oix1 = std::max(D.oix1, 1);
    F.oixn = std::min(D.oixn, np);
    F.oixd = std::max(D.oixd, 1);
    F.oitd = std::max(D.oitd, 1);
    F.onx = (F.oixn >= F.oix1) ? (F.oixn - F.oix1) / F.oixd + 1 : 0;
    const int nt_max = std::max(D.nt_max, 0);
    F.ncall_max = nt_max + 1;
    F.nrec_max = nt_max / F.oitd + 2;
    b->records.alloc((size_t)F.nrec_max * 6 * std::max(F.onx, 1));
    b->records.zero();
    b->potency.alloc((size_t)F.ncall_max * 2 * (ndof + 1));
    b->potency.zero();
    b->wctas = std::max(1, std::min(DYNW_MAX_CTAS, (np + 4 * DYNW_THREADS - 1) / (4 * DYNW_THREADS)));
    b->wpart.alloc((size_t)DYNW_MAX_CTAS * 6);
    b->wpart.zero();
    b->wticket.alloc(1);
    b->wticket.zero();
    std::vector<int> ost = {D.oit, 0, 0};
    b->ostate.upload(ost);
    F.ostate = b->ostate.p;
    F.records = b->records.p;
    F.potency = b->potency.p;
    faults.push_back(std::move(b));
    bc_order.push_back({BC_DYNFLT, (int)faults.size() - 1});
    return (int)faults.size() - 1;
  }

  int add_force(int iglob, const double dir[2]) override {
    S2D_REQUIRE(!committed, "add_force after commit");
    S2D_REQUIRE(iglob >= 1 && (size_t)iglob <= npoin, "add_force: node id out of range");
    h_src_iglob.push_back(iglob);
    h_src_dir.push_back(dir[0]);
    h_src_dir.push_back(dir[1]);
    return (int)h_src_iglob.size() - 1;
  }

  void add_periodic(int np, const int32_t* master, const int32_t* slave) override {
    S2D_REQUIRE(!committed, "add_periodic after commit");
    S2D_REQUIRE(np > 0 && master && slave, "add_periodic: empty boundary");
    check_nodes(np, master, "add_periodic");
    check_nodes(np, slave, "add_periodic");
    std::unique_ptr<PerioBc> b(new PerioBc());
    b->np = np;
    b->master.upload(std::vector<int32_t>(master, master + np));
    b->slave.upload(std::vector<int32_t>(slave, slave + np));
    std::vector<int32_t> both(master, master + np);
    both.insert(both.end(), slave, slave + np);
    h_bc_nodes.emplace_back(master, master + np);  // two lists: each side flags its own row / column
    h_bc_nodes.emplace_back(slave, slave + np);
    perio.push_back(std::move(b));
  }

  // interpolated stations given by the N*N nodes of their element (structured builder: there is no ibool table)
  void add_receivers_nodes(int nx, char field, int isamp, int nt_rec, const int32_t* nodes, const double* interp) override {
    S2D_REQUIRE(!committed, "add_receivers after commit");
    S2D_REQUIRE(nx > 0 && isamp > 0 && nt_rec > 0 && nodes && interp, "add_receivers_nodes: bad arguments");
    S2D_REQUIRE(field == 'D' || field == 'V' || field == 'A', "add_receivers: field must be D, V or A");
    const size_t n2 = (size_t)ngll * ngll;
    check_nodes((int)(n2 * nx), nodes, "add_receivers_nodes");
    std::vector<int32_t> eidx(nx);
    for (int n = 0; n < nx; ++n) eidx[n] = n + 1;
    rec.present = true;
    rec.field = field;
    RecDev& R = rec.dev;
    R.nx = nx;
    R.ndof = ndof;
    R.isamp = isamp;
    R.nt = nt_rec;
    R.at_node = 0;
    R.ngll = ngll;
    rec.einterp.upload(eidx.data(), nx);
    rec.interp.upload(interp, n2 * nx);
    rec.nodes.upload(nodes, n2 * nx);
    rec.sis.alloc((size_t)nt_rec * nx * ndof);
    rec.sis.zero();
    R.iglob = nullptr;
    R.einterp = rec.einterp.p;
    R.interp = rec.interp.p;
    R.ibool = rec.nodes.p;
    R.sis = rec.sis.p;
  }
  void set_node_kv(const double* eta_ref) override {
    S2D_REQUIRE(cart_mode, "set_node_kv: structured builder only (the generic engine takes s2d_set_kv)");
    S2D_REQUIRE(!committed, "set_node_kv after commit");
    DevBuf<double> tmp;
    tmp.upload(eta_ref, npoin_ref);
    cart_kv_eta.alloc(npoin);
    cart_kv_eta.zero(stream);
    cart_from_ref1(tmp.p, cart_kv_eta.p);
    S2D_CUDA(cudaStreamSynchronize(stream));
  }

  // eta per element GLL point, already in the strip layout (s2d_cart_set_kv_elems)
  void set_strip_plastic(const unsigned char* set_strip, size_t n, int nsets, const double* raw) override {
    S2D_REQUIRE(cart_mode && !committed, "set_strip_plastic: builder-made engines only, before commit");
    S2D_REQUIRE(n == (size_t)nelem && nsets >= 1 && nsets < STRIP_PL_SETS, "set_strip_plastic: 1..7 plastic material sets");
    S2D_REQUIRE(ndof == 2 && cart_compact && ngll <= STRIP_PLAST_MAXN,
                "MAT_init_work: plasticity requires ndof=2 (P-SV), an isotropic box and ngll <= 6");
    pl_set.alloc(n);
    h2d_sync(pl_set.p, set_strip, n);
    pl_ep.alloc((size_t)nelem * 3 * ngll * ngll);
    pl_ep.zero();
    for (int k = 0; k < nsets; ++k)
      for (int q = 0; q < 6; ++q) pl_raw[k + 1][q] = raw[(size_t)6 * k + q];
  }
  void set_strip_visco(const unsigned char* set_strip, size_t n, int nsets, const int32_t* nbody, const double* moduli,
                       const double* wbody, const double* theta) override {
    S2D_REQUIRE(cart_mode && !committed, "set_strip_visco: builder-made engines only, before commit");
    S2D_REQUIRE(n == (size_t)nelem && nsets >= 1 && nsets < STRIP_PL_SETS, "set_strip_visco: 1..7 visco-elastic material sets");
    S2D_REQUIRE(ndof == 2 && cart_compact && ngll <= STRIP_PLAST_MAXN,
                "MAT_init_work: visco-elasticity requires ndof=2 (P-SV), an isotropic box and ngll <= 6");
    S2D_REQUIRE(pl_ep.n == 0, "plastic and visco-elastic elements in one problem: not provided");
    vs_nb = 0;
    for (int k = 0; k < nsets; ++k) {
      S2D_REQUIRE(nbody[k] >= 1 && nbody[k] <= STRIP_VS_MAXB, "set_strip_visco: Nbody must be in 1..8");
      vs_nb = std::max(vs_nb, (int)nbody[k]);
      double* r = vs_raw[k + 1];
      r[0] = moduli[2 * k];
      r[1] = moduli[2 * k + 1];
      r[2] = (double)nbody[k];
      for (int b = 0; b < nbody[k]; ++b) {
        r[3 + b] = wbody[(size_t)STRIP_VS_MAXB * k + b];
        for (int c = 0; c < 3; ++c) r[3 + (c + 1) * STRIP_VS_MAXB + b] = theta[((size_t)k * 3 + c) * STRIP_VS_MAXB + b];
      }
    }
    pl_set.alloc(n);
    h2d_sync(pl_set.p, set_strip, n);
    vs_state.alloc((size_t)nelem * 3 * (vs_nb + 1) * ngll * ngll);
    vs_state.zero();
  }
  // par(13,nsets): lambda, mu, phi [deg], alpha0, Cd, beta, R, e0(3), ep0(3); state_strip: initial alpha, ep(3) of every
  // element GLL point in the strip layout, or null when every set starts from alpha = 0, ep = 0
  void set_strip_damage(const unsigned char* set_strip, size_t n, int nsets, const double* par, const double* state_strip) override {
    S2D_REQUIRE(cart_mode && !committed, "set_strip_damage: builder-made engines only, before commit");
    S2D_REQUIRE(n == (size_t)nelem && nsets >= 1 && nsets < STRIP_PL_SETS, "set_strip_damage: 1..7 damage material sets");
    S2D_REQUIRE(ndof == 2 && cart_compact && ngll <= STRIP_PLAST_MAXN,
                "MAT_init_work: the damage rheology requires ndof=2 (P-SV), an isotropic box and ngll <= 6");
    S2D_REQUIRE(pl_ep.n == 0 && vs_state.n == 0, "damage together with plastic or visco-elastic elements: not provided");
    for (int k = 0; k < nsets; ++k) {
      const double* p = par + (size_t)13 * k;
      const double lam = p[0], mu1 = p[1], phi = p[2], alpha0 = p[3], Cd = p[4], beta = p[5], R = p[6];
      S2D_REQUIRE(mu1 > 0.0 && beta >= 0.0, "set_strip_damage: mu must be positive, beta non-negative");
      const double q = std::sin(phi * 3.141592653589793 / 180.0);                      // xi_zero_2d (mat_damage.f90:295-305)
      const double xi0 = -std::sqrt(2.0) / std::sqrt(q * q * ((lam / mu1 + 1.0) * (lam / mu1 + 1.0)) + 1.0);
      const double qq = 2.0 * (mu1 + lam) / (2.0 - xi0 * xi0);                         // gamma_r_2d (:308-317)
      const double pp = 0.5 * xi0 * (qq + lam);
      const double gr = pp + std::sqrt(pp * pp + 2.0 * mu1 * qq);
      double* r = dm_raw[k + 1];
      r[0] = lam; r[1] = mu1; r[2] = xi0; r[3] = gr; r[4] = beta; r[5] = Cd; r[6] = R / mu1;
      for (int c = 0; c < 3; ++c) r[7 + c] = p[7 + c];
      // initial stress (:262-265): compute_stress of e0 - ep0 with the moduli damaged by alpha0
      const double mud = mu1 + xi0 * gr * alpha0, rg = gr * std::pow(alpha0, 1.0 + beta) / (1.0 + beta);
      const double e[3] = {p[7] - p[10], p[8] - p[11], p[9] - p[12]};
      const double i1 = e[0] + e[1], i2 = e[0] * e[0] + e[1] * e[1] + 2.0 * e[2] * e[2], si2 = std::sqrt(i2);
      const double xi = si2 < 1e-10 ? 0.0 : i1 / si2, two_mue = 2.0 * mud - rg * xi;
      r[10] = lam * i1 - rg * si2 + two_mue * e[0];
      r[11] = lam * i1 - rg * si2 + two_mue * e[1];
      r[12] = two_mue * e[2];
    }
    pl_set.alloc(n);
    h2d_sync(pl_set.p, set_strip, n);
    const size_t ns = (size_t)nelem * 4 * ngll * ngll;
    if (state_strip) upload_as(dm_state, state_strip, ns);
    else {
      dm_state.alloc(ns);
      dm_state.zero();
    }
  }
  void get_strip_damage_state(double* state_strip) override {
    S2D_REQUIRE(dm_state.n > 0, "no damage elements");
    S2D_CUDA(cudaStreamSynchronize(stream));
    std::vector<T> tmp = dm_state.to_host();
    for (size_t q = 0; q < tmp.size(); ++q) state_strip[q] = (double)tmp[q];
  }
  void get_strip_plastic_strain(double* ep_strip) override {
    S2D_REQUIRE(pl_ep.n > 0, "no plastic elements");
    S2D_CUDA(cudaStreamSynchronize(stream));
    std::vector<T> tmp = pl_ep.to_host();
    for (size_t q = 0; q < tmp.size(); ++q) ep_strip[q] = (double)tmp[q];
  }

  void set_strip_eta(const double* eta_strip, size_t n) override {
    S2D_REQUIRE(cart_mode && !committed, "set_strip_eta: builder-made engines only, before commit");
    upload_as(strip_eta, eta_strip, n);
    cart_kv_eta.release();
  }

  int add_moment(int nterms, const int32_t* node, const double* coef_) override {
    S2D_REQUIRE(!committed, "add_moment after commit");
    S2D_REQUIRE(nterms > 0 && node && coef_, "add_moment: empty source");
    for (int t = 0; t < nterms; ++t) S2D_REQUIRE(node[t] >= 1 && (size_t)node[t] <= npoin, "add_moment: node id out of range");
    h_src_iglob.push_back(0);  // keeps the source index space of the amplitude table
    h_src_dir.push_back(0.0);
    h_src_dir.push_back(0.0);
    h_mom_src.push_back((int)h_src_iglob.size() - 1);
    h_mom_node.insert(h_mom_node.end(), node, node + nterms);
    h_mom_coef.insert(h_mom_coef.end(), coef_, coef_ + (size_t)nterms * ndof);
    h_mom_start.push_back((int)h_mom_node.size());
    return (int)h_src_iglob.size() - 1;
  }

  void add_receivers(int nx, char field, int isamp, int nt_rec, int at_node, const int32_t* iglob,
                     const int32_t* einterp, const double* interp) override {
    S2D_REQUIRE(!committed, "add_receivers after commit");
    S2D_REQUIRE(nx > 0 && isamp > 0 && nt_rec > 0, "add_receivers: bad sizes");
    S2D_REQUIRE(field == 'D' || field == 'V' || field == 'A', "add_receivers: field must be D, V or A");
    rec.present = true;
    rec.field = field;
    RecDev& R = rec.dev;
    R.nx = nx;
    R.ndof = ndof;
    R.isamp = isamp;
    R.nt = nt_rec;
    R.at_node = at_node;
    R.ngll = ngll;
    if (at_node) {
      S2D_REQUIRE(iglob != nullptr, "add_receivers: iglob missing");
      check_nodes(nx, iglob, "add_receivers");
      rec.iglob.upload(iglob, nx);
    } else {
      S2D_REQUIRE(einterp && interp, "add_receivers: einterp/interp missing");
      for (int n = 0; n < nx; ++n) S2D_REQUIRE(einterp[n] >= 1 && einterp[n] <= nelem, "add_receivers: element out of range");
      rec.einterp.upload(einterp, nx);
      rec.interp.upload(interp, (size_t)ngll * ngll * nx);
    }
    rec.sis.alloc((size_t)nt_rec * nx * ndof);
    rec.sis.zero();
    R.iglob = rec.iglob.p;
    R.einterp = rec.einterp.p;
    R.interp = rec.interp.p;
    R.ibool = ibool.p;
    R.sis = rec.sis.p;
  }

  // ---- strip-kernel tables ---------------------------------------------------------------
  // halo arrays of the strip kernel and the deferred nodes of the fused step that come from the decomposition
  // itself: rows shared by two bands, columns shared by two groups, GPU interface columns
  void init_strip_tables(const StripGeom& S) {
    cart_S = S;
    cart_hx.alloc((size_t)ndof * std::max(S.ngroups - 1, 0) * S.LZ + 1);
    cart_hz.alloc((size_t)ndof * S.nseg * S.nstrips * S.WL + 1);
    cart_hx.zero(stream);
    cart_hz.zero(stream);
    cart_meet.alloc((size_t)S.nseg * std::max(S.ngroups - 1, 1));
    cart_meet.zero(stream);
    h_rowflag.assign(S.LZ, 0);
    h_colflag.assign(S.LX, 0);
    for (int gz = 0; gz < S.LZ; ++gz)
      if (strip_shared_row_seg(S, gz) >= 0) h_rowflag[gz] = 1;
    for (int hb = 0; hb + 1 < S.ngroups; ++hb) {
      int sr;
      h_colflag[strip_halo_col(S, hb, sr)] = 2;  // not finished by its owner lane, but inside the strip kernel
    }
    if (S.xhalo_left) h_colflag[0] = 1;
    if (S.xhalo_right) h_colflag[S.LX - 1] = 1;
  }

  // A structured box handed over through the generic API (s2d_create with the Fortran host's RCM-ordered ibool,
  // s2d_set_elastic with its coefficient blocks): recognised from the topology (structured.hpp) and moved onto the
  // GLL lattice, so that it runs on the z-marching strip kernel like a builder-made box.  Everything that was
  // uploaded in the caller's node numbering -- fields, inverse mass, boundary node lists, sources, receivers -- is
  // re-indexed once, here; the API keeps the caller's numbering (cart_to_ref / cart_from_ref).
  bool routed = false;      // generic handle running on the strip kernel
  bool have_colors = false;
  DevBuf<int> lat_of;       // (npoin_ref) 0-based lattice index of every caller node
  std::vector<int32_t> h_lat_of;
  bool try_route_to_strips() {
    if (cart_mode || variant != S2D_ASM_PATCH || env_int("S2D_ROUTE_STRIP", 1) == 0) return false;
    if (!(nelast == 2 || nelast == 6)) return false;          // flat-grid planes only (mat_elastic.f90:255-268)
    if (ndof == 2 && kd2 != (ngll == 5 ? 1 : 0)) return false;  // the strip kernel follows OPT_NGLL (mat_elastic.f90:412)
    if (nkv > 0 && !strip_kv_ok()) return false;
    int32_t hint = 0;
    for (auto& r : bc_order)
      if (r.kind == BC_DYNFLT && faults[r.index]->dev.two_sides && !h_fault_node1.empty()) {
        hint = h_fault_node1[r.index];
        break;
      }
    StructuredBox B = detect_structured(h_ibool.data(), ngll, nelem, npoin, hint);
    if (!B.ok) return false;
    const StripGeom S = make_strip_geom(ngll, ndof, B.nx, B.nz, B.ezflt, strip_default_seg(ngll, B.nx, B.nz), false, false);
    const size_t nlat = (size_t)S.LXP * S.LZ, nref = npoin;
    if (nlat > 2147483647ull) return false;
    // The fused update reads ONE inverse mass per node for the nodes no boundary condition touches (the reference's
    // mass has equal columns there, mat_mass.f90:56-57; only BC_ABSO_init / BC_PERIO_init change single columns).
    // A caller whose rmass differs between components elsewhere keeps the any-mesh kernels.
    if (ndof == 2) {
      std::vector<T> hr = rmass.to_host();
      std::vector<uint8_t> touched(nref, 0);
      for (auto& L : h_bc_nodes)
        for (int32_t nd : L) touched[(size_t)nd - 1] = 1;
      for (size_t k = 0; k < nref; ++k)
        if (!touched[k] && hr[k] != hr[k + nref]) return false;
    }
    h_lat_of.resize(nref);
    for (size_t k = 0; k < nref; ++k) h_lat_of[k] = (int32_t)((size_t)B.gz[k] * S.LXP + B.gx[k]);
    lat_of.upload(h_lat_of);
    // fields and inverse mass onto the lattice (pad columns: zero fields, unit inverse mass)
    auto relayout = [&](DevBuf<T>& buf, T padval) {
      DevBuf<T> nb;
      nb.alloc(nlat * ndof);
      k_fill_n<T><<<grid_for(nb.n), 256, 0, stream>>>(nb.p, nb.n, padval);
      k_lat_permute<T, T><<<grid_for(nref), 256, 0, stream>>>(buf.p, nb.p, lat_of.p, nref, nlat, ndof, 0);
      S2D_CUDA(cudaStreamSynchronize(stream));
      buf = std::move(nb);
    };
    relayout(d, (T)0);
    relayout(v, (T)0);
    relayout(a, (T)0);
    relayout(rmass, (T)1);
    if (mass.n == nref) {
      DevBuf<double> nb;
      nb.alloc(nlat);
      nb.zero(stream);
      k_lat_permute<double, double><<<grid_for(nref), 256, 0, stream>>>(mass.p, nb.p, lat_of.p, nref, nlat, 1, 0);
      S2D_CUDA(cudaStreamSynchronize(stream));
      mass = std::move(nb);
    }
    // coefficient planes in the strip layout, all nelast planes stored (what matwrk_elast_type%a holds)
    {
      const int N = ngll, n2 = N * N;
      std::vector<T> pc((size_t)nelem * nelast * n2);
      for (int e = 0; e < nelem; ++e) {
        const double* src = h_coef.data() + (size_t)h_elem2set[e] * nelast * n2;
        for (int pl = 0; pl < nelast; ++pl)
          for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i)
              pc[strip_coef_index(S, nelast, B.ex[e], B.ez[e], i, j, pl)] = (T)src[(size_t)pl * n2 + i + N * j];
      }
      p_coef.upload(pc);
      if (!h_beta.empty()) {  // 2.5D: beta(ngll,ngll) of every element through its coefficient set
        std::vector<T> pb((size_t)nelem * n2);
        for (int e = 0; e < nelem; ++e)
          for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i)
              pb[strip_scalar_index(S, B.ex[e], B.ez[e], i, j)] = (T)h_beta[(size_t)h_elem2set[e] * n2 + i + N * j];
        strip_beta.upload(pb);
      }
      if (nkv > 0) {  // eta(ngll,ngll) of the Kelvin-Voigt elements, zero elsewhere (mat_kelvin_voigt.f90:137-150)
        std::vector<T> pe((size_t)nelem * n2, (T)0);
        for (int e = 0; e < nelem; ++e) {
          if (h_elem2kv[e] < 0) continue;
          for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i)
              pe[strip_scalar_index(S, B.ex[e], B.ez[e], i, j)] = (T)h_eta[(size_t)h_elem2kv[e] * n2 + i + N * j];
        }
        strip_eta.upload(pe);
      }
    }
    // node ids held by the boundary conditions, sources and receivers
    auto remap_dev = [&](DevBuf<int>& ids) {
      if (ids.n == 0) return;
      k_remap_ids<<<grid_for(ids.n), 256, 0, stream>>>(ids.p, ids.n, lat_of.p);
    };
    auto remap_host = [&](std::vector<int32_t>& ids) {
      for (auto& x : ids)
        if (x > 0) x = h_lat_of[(size_t)x - 1] + 1;
    };
    for (auto& b : abso) remap_dev(b->node);
    for (auto& b : dirneu) remap_dev(b->node);
    for (auto& b : faults) {
      remap_dev(b->node1);
      remap_dev(b->node2);
    }
    for (auto& b : perio) {
      remap_dev(b->master);
      remap_dev(b->slave);
    }
    remap_dev(rec.iglob);
    remap_dev(rec.nodes);
    remap_dev(ibool);  // interpolated receivers read the nodes of their element from it
    for (auto& L : h_bc_nodes) remap_host(L);
    remap_host(h_src_iglob);
    remap_host(h_mom_node);
    S2D_CUDA(cudaStreamSynchronize(stream));
    npoin_ref = nref;
    npoin = nlat;
    cart_mode = true;
    routed = true;
    rmass_is_inverse = true;
    cart_compact = 0;
    p_hetero = true;
    init_strip_tables(S);
    install_lat_table(h_lat_of);
    return true;
  }
  // the caller's node numbering as a table: lat_of[k] = 0-based lattice index of the caller's node k + 1
  void install_lat_table(const std::vector<int32_t>& table) {
    h_lat_of = table;
    lat_of.upload(h_lat_of);
    Engine<T>* Ep = this;
    cart_to_ref = [Ep](const T* lat, double* ref) {
      k_lat_permute<T, double><<<Ep->grid_for(Ep->npoin_ref), 256, 0, Ep->stream>>>(lat, ref, Ep->lat_of.p, Ep->npoin_ref, Ep->npoin, Ep->ndof, 1);
      S2D_CUDA(cudaGetLastError());
    };
    cart_from_ref = [Ep](const double* ref, T* lat) {
      k_lat_permute<double, T><<<Ep->grid_for(Ep->npoin_ref), 256, 0, Ep->stream>>>(ref, lat, Ep->lat_of.p, Ep->npoin_ref, Ep->npoin, Ep->ndof, 0);
      S2D_CUDA(cudaGetLastError());
    };
    cart_from_ref1 = [Ep](const double* ref, T* lat) {
      k_lat_permute<double, T><<<Ep->grid_for(Ep->npoin_ref), 256, 0, Ep->stream>>>(ref, lat, Ep->lat_of.p, Ep->npoin_ref, Ep->npoin, 1, 0);
      S2D_CUDA(cudaGetLastError());
    };
    cart_from_ref1d = [Ep](const double* ref, double* lat) {
      k_lat_permute<double, double><<<Ep->grid_for(Ep->npoin_ref), 256, 0, Ep->stream>>>(ref, lat, Ep->lat_of.p, Ep->npoin_ref, Ep->npoin, 1, 0);
      S2D_CUDA(cudaGetLastError());
    };
  }
  bool rmass_is_inverse = false;
  std::vector<int32_t> h_fault_node1;  // first node1 of every fault (which side of a split-node row is the lower one)
  std::vector<double> h_eta;
  DevBuf<T> strip_eta;                 // Kelvin-Voigt eta per element GLL point in the strip layout (zero off the KV elements)
  DevBuf<T> v_alt, a_alt;              // second velocity / acceleration buffers of the fused Kelvin-Voigt step
  // Coulomb plasticity (mat_plastic.f90): material set per element (strip order), plastic strain per element GLL
  // point, per set (coh, phi [deg], Tv, e0(3)) as read and (yield_co, yield_mu, vp_factor, e0(3)) as the kernel uses them
  DevBuf<unsigned char> pl_set;
  DevBuf<T> pl_ep, pl_tab;
  // visco-elasticity (mat_visco.f90): the element sets share pl_set; memory variables + previous strain per element
  // GLL point; per set lambda_inf, mu_inf, Nbody, wbody(8) [RK factors once dt is known], theta(8,3)
  // damage rheology (mat_damage.f90): alpha + ep(3) per element GLL point; per set the 16-entry row of StripArgs::dm_tab
  DevBuf<T> dm_state, dm_tab;
  double dm_raw[STRIP_PL_SETS][STRIP_DM_TAB] = {};
  DevBuf<T> vs_state, vs_tab;
  int vs_nb = 0;
  double vs_raw[STRIP_PL_SETS][STRIP_VS_TAB] = {};
  bool a_stale = false;                // the fused leapfrog step left the accelerations of the nodes it advanced unwritten
  bool accel_lazy = env_int("S2D_ACCEL_LAZY", 1) != 0;
  double pl_raw[STRIP_PL_SETS][6] = {}, pl_par[STRIP_PL_SETS][6] = {};
  static bool strip_kv_ok() { return true; }   // k_elem_strip<KV>: d + eta*v element by element

  // ---- planning ------------------------------------------------------------------------
  // Deferred nodes of the fused step: halo rows / columns (flagged by the builder) plus every node a
  // boundary condition or a source touches.  A list that runs along one lattice column flags that
  // column, anything else flags the rows of its nodes.
  void build_deferred_tables() {
    const int LX = cart_S.LX, LZ = cart_S.LZ, LXP = cart_S.LXP;
    h_rowflag.resize(LZ, 0);
    h_colflag.resize(LX, 0);
    std::vector<std::vector<int32_t>> lists = h_bc_nodes;
    std::vector<int32_t> forces;
    for (int nd : h_src_iglob)
      if (nd > 0) forces.push_back(nd);
    lists.push_back(forces);
    for (size_t m = 0; m + 1 < h_mom_start.size(); ++m)  // one list per moment source: a row and a column
      lists.emplace_back(h_mom_node.begin() + h_mom_start[m], h_mom_node.begin() + h_mom_start[m + 1]);
    for (auto& L : lists) {
      std::vector<int> colcount(LX, 0), rowcount(LZ, 0);
      for (int nd : L) {
        colcount[(nd - 1) % LXP]++;
        rowcount[(nd - 1) / LXP]++;
      }
      for (int nd : L) {
        const int gx = (nd - 1) % LXP, gz = (nd - 1) / LXP;
        if (colcount[gx] > rowcount[gz]) h_colflag[gx] = 1;  // overrides 2 (group-boundary column)
        else h_rowflag[gz] = 1;
      }
    }
    std::vector<int> rows, cols;
    for (int r = 0; r < LZ; ++r)
      if (h_rowflag[r]) rows.push_back(r);
    for (int c = 0; c < LX; ++c)
      if (h_colflag[c] == 1) cols.push_back(c);  // 2 = group-boundary column: advanced inside k_elem_strip
    ndrows = (int)rows.size();
    ndcols = (int)cols.size();
    rowflag.upload(h_rowflag);
    colflag.upload(h_colflag);
    if (ndrows) drows.upload(rows);
    if (ndcols) dcols.upload(cols);
    d2.alloc(npoin * ndof);
    d2.zero();
    tm_ready = false;
    // measured on B200 (4096^2 FP64, ms per launch, per-lane copies -> boxes): compact leapfrog 5.96 -> 5.74, compact
    // Newmark 7.32 -> 6.81, six stored planes 7.39 -> 7.43 (its coefficient block already comes by TMA): not there
    if (strip_tensor && ngll <= S2D_STRIP_TENSOR_MAXN && (cart_compact || ndof == 1)) {
      const int bw = strip_box_width(ngll, (int)sizeof(T));
      // the TENS variant keeps its CTAs per SM only while the shared memory fits (static part: tile, hand-over, masks)
      const int fusedk = scheme.kind == 1 ? 2 : 1;
      const size_t smem = strip_tens_smem(ngll, ndof, (int)sizeof(T), cart_compact != 0, fusedk) +
                          (size_t)strip_warps() * ndof * ngll * (32 / ngll) * ngll * sizeof(T) + 2048;
      if (smem * strip_min_ctas(ngll, (int)sizeof(T)) <= 227 * 1024) {
        const StripGeom& S = cart_S;
        tm_d[0] = lattice_tmap(d.p, (int)sizeof(T), S.LXP, S.LZ, ndof, npoin, bw, ngll - 1);
        tm_d[1] = lattice_tmap(d2.p, (int)sizeof(T), S.LXP, S.LZ, ndof, npoin, bw, ngll - 1);
        tm_v = lattice_tmap(v.p, (int)sizeof(T), S.LXP, S.LZ, ndof, npoin, bw, ngll - 1);
        tm_a = lattice_tmap(a.p, (int)sizeof(T), S.LXP, S.LZ, ndof, npoin, bw, ngll - 1);
        tm_r = lattice_tmap(rmass.p, (int)sizeof(T), S.LXP, S.LZ, 0, npoin, bw, ngll - 1);
        tm_ready = true;
      }
    }
  }
  void build_color_plan() {
    const int n2 = ngll * ngll;
    ncolors = greedy_coloring(h_ibool.data(), n2, nelem, npoin, h_color);
    color_start.assign(ncolors + 1, 0);
    for (int e = 0; e < nelem; ++e) color_start[h_color[e] + 1]++;
    for (int c = 0; c < ncolors; ++c) color_start[c + 1] += color_start[c];
    std::vector<int> list(nelem), fill(color_start.begin(), color_start.end() - 1);
    for (int e = 0; e < nelem; ++e) list[fill[h_color[e]]++] = e;
    color_elems.upload(list);
  }
  static int patch_EP(int ngll) {
    const int cap = patch_ep(ngll);
    const int want = env_int("S2D_EP", ngll <= 6 ? 32 : cap);
    return std::max(4, std::min(cap, want));
  }
  void build_patch_plan_dev() {
    const int n2 = ngll * ngll;
    build_patch_plan(h_ibool.data(), n2, nelem, npoin, patch_EP(ngll), plan);
    upload_patch_plan();
  }
  void upload_patch_plan() {
    const int n2 = ngll * ngll;
    PatchPlan& P = plan;
    const int EP = P.EP;
    pp_npatch = P.npatch;
    pp_EP = EP;
    pp_max_nloc = P.max_nloc;
    pp_max_colors = P.max_colors;
    pp_nslots = std::max<size_t>(P.nslots, 1);
    std::vector<int> eset(nelem), ekv(nelem, -1), pshape(P.npatch);
    std::vector<uint16_t> lidx((size_t)P.npatch * EP * n2, 0);
    std::vector<uint8_t> ecol((size_t)P.npatch * EP, 0);
    std::vector<int> slot((size_t)P.npatch * P.max_nloc, -1);
    std::vector<long long> sbase(P.npatch, 0), pns(P.npatch + 1);
    for (int p = 0; p < P.npatch; ++p) {
      pshape[p] = p;  // unstructured plan: every patch is its own shape
      const int es = P.pelem_start[p], cnt = P.pelem_start[p + 1] - es;
      for (int el = 0; el < cnt; ++el) {
        const int e = P.elems[es + el];
        eset[es + el] = h_elem2set[e];
        if (nkv > 0) ekv[es + el] = h_elem2kv[e];
        ecol[(size_t)p * EP + el] = (uint8_t)P.ecolor[es + el];
        for (int k = 0; k < n2; ++k) lidx[((size_t)p * EP + el) * n2 + k] = P.lidx[(size_t)(es + el) * n2 + k];
      }
      const int ns = P.pnode_start[p], nl = P.pnode_start[p + 1] - ns;
      for (int l = 0; l < nl; ++l) slot[(size_t)p * P.max_nloc + l] = P.pslot[ns + l];
    }
    for (int p = 0; p <= P.npatch; ++p) pns[p] = P.pnode_start[p];
    p_pelem_start.upload(P.pelem_start);
    p_pshape.upload(pshape);
    p_sh_lidx.upload(lidx);
    p_sh_ecolor.upload(ecol);
    p_sh_slot.upload(slot);
    p_pslot_base.upload(sbase);
    p_pnode_start.upload(pns);
    p_pnode.upload(P.pnode);
    p_eset.upload(eset);
    if (nkv > 0) p_ekv.upload(ekv);
    p_snode.upload(P.snode);
    p_sstart.upload(P.sstart);
    fhalo.alloc(pp_nslots * ndof);
    fhalo.zero();
    // heterogeneous media (one coefficient block per element): re-lay the planes patch-major so
    // that each thread of the patch kernel streams them with unit stride
    p_hetero = (ncoefsets == nelem);
    if (p_hetero) {
      std::vector<T> pc((size_t)nelem * nelast * n2);
      for (int p = 0; p < P.npatch; ++p) {
        const int es = P.pelem_start[p], cnt = P.pelem_start[p + 1] - es;
        const size_t base = (size_t)es * nelast * n2;
        const size_t nthr = (size_t)cnt * ngll;
        for (int el = 0; el < cnt; ++el) {
          const double* src = h_coef.data() + (size_t)h_elem2set[P.elems[es + el]] * nelast * n2;
          for (int pl = 0; pl < nelast; ++pl)
            for (int j = 0; j < ngll; ++j)
              for (int i = 0; i < ngll; ++i)
                pc[base + ((size_t)pl * ngll + i) * nthr + (size_t)el * ngll + j] =
                    (T)src[(size_t)pl * n2 + i + ngll * j];
        }
      }
      p_coef.upload(pc);
    }
  }

  void commit(int variant_) override {
    S2D_REQUIRE(!committed, "commit called twice");
    S2D_REQUIRE(nelast > 0, "commit: s2d_set_elastic was not called");
    S2D_REQUIRE(variant_ >= 0 && variant_ <= 2, "commit: unknown assembly variant");
    variant = variant_;
    if (!cart_mode) {
      if (nkv == 0) h_elem2kv.assign(nelem, -1);
      build_color_plan();  // s2d_get_coloring (and S2D_ASM_COLOR)
      have_colors = true;
      try_route_to_strips();
    }
    if (cart_mode) {
      S2D_REQUIRE(variant == S2D_ASM_PATCH, "commit: the structured builder only provides the strip kernel");
      if (!rmass_is_inverse) k_invert<T><<<grid_for(rmass.n), 256, 0, stream>>>(rmass.p, rmass.n);
      if (cart_kv_eta.n) {  // s2d_cart_set_kv: eta at the nodes -> eta(ngll,ngll) of every element
        const size_t tot = (size_t)nelem * ngll * ngll;
        strip_eta.alloc(tot);
        k_strip_eta_from_nodes<T><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(cart_S, cart_kv_eta.p, strip_eta.p);
        S2D_CUDA(cudaStreamSynchronize(stream));
        cart_kv_eta.release();
      }
      if (dm_state.n) {
        for (int k = 1; k < STRIP_PL_SETS; ++k) dm_raw[k][13] = scheme.dt;
        upload_as(dm_tab, &dm_raw[0][0], (size_t)STRIP_PL_SETS * STRIP_DM_TAB);
      }
      if (vs_state.n) {  // RK_factor of MAT_VISCO_stress (mat_visco.f90:222-224), now that dt is known
        double tab[STRIP_PL_SETS][STRIP_VS_TAB];
        std::memcpy(tab, vs_raw, sizeof(tab));
        for (int k = 1; k < STRIP_PL_SETS; ++k)
          for (int b = 0; b < (int)vs_raw[k][2]; ++b) {
            const double x = vs_raw[k][3 + b] * scheme.dt;
            tab[k][3 + b] = x - (x * x) / 2.0 + (x * x * x) / 6.0 - (x * x * x * x) / 24.0;
          }
        upload_as(vs_tab, &tab[0][0], (size_t)STRIP_PL_SETS * STRIP_VS_TAB);
      }
      if (pl_set.n && !vs_state.n && !dm_state.n) {  // MAT_PLAST_init_elem_work (mat_plastic.f90:165-184); set 0 (elastic elements) never yields
        for (int k = 1; k < STRIP_PL_SETS; ++k) {
          const double phi = 3.141592653589793 / 180.0 * pl_raw[k][1];
          pl_par[k][0] = pl_raw[k][0] * std::cos(phi);
          pl_par[k][1] = std::sin(phi);
          pl_par[k][2] = pl_raw[k][2] > 0.0 ? 1.0 - std::exp(-scheme.dt / pl_raw[k][2]) : 1.0;
          for (int q = 3; q < 6; ++q) pl_par[k][q] = pl_raw[k][q];
        }
        upload_as(pl_tab, &pl_par[0][0], (size_t)STRIP_PL_SETS * 6);
      }
      // the node update rides in the strip kernel for leapfrog and for the explicit Newmark scheme (beta = 0)
      fused = (scheme.kind == 0 || (scheme.kind == 1 && scheme.beta == 0.0)) && env_int("S2D_FUSED", 1) != 0;
      // Kelvin-Voigt elements read their neighbours' velocities: the fused update then writes v (Newmark: and a)
      // into a second buffer (S2D_KV_FUSED=0: the separate predictor / corrector passes of round 1)
      if (strip_eta.n && (ngll > STRIP_KV_FUSED_MAXN || env_int("S2D_KV_FUSED", 1) == 0)) fused = false;
      if (fused && strip_eta.n) {
        v_alt.alloc(npoin * ndof);
        v_alt.zero(stream);
        if (scheme.kind == 1) {
          a_alt.alloc(npoin * ndof);
          a_alt.zero(stream);
        }
      }
      if (fused) build_deferred_tables();
    } else {
      if (variant == S2D_ASM_PATCH) build_patch_plan_dev();
    }
    build_source_terms();
    build_node_ops();
    if (scheme.kind == 2) {
      d_alpha.alloc(npoin * ndof);
      v_alpha.alloc(npoin * ndof);
    }
    if (scheme.kind == 3)
      S2D_REQUIRE(scheme.nstages >= 1 && scheme.nstages <= S2D_MAX_STAGES, "commit: symplectic scheme needs 1..8 stages");
    committed = true;
    // it = 0 outputs: BC_write(bc,0) at the end of BC_init (bc_gen.f90:249), REC_store(rec,0) (main.f90:35)
    launch_outputs();
    S2D_CUDA(cudaStreamSynchronize(stream));
  }

  // ---- launches ------------------------------------------------------------------------
  int grid_for(size_t n, int block = 256) const {
    size_t g = (n + block - 1) / block;
    return (int)std::min<size_t>(g, 148 * 16);
  }

  bool needs_zero_f() const { return variant != S2D_ASM_PATCH; }

  // f = -K d  (compute_Fint, solver.f90:273-320); f must be zero on entry unless the patch variant
  void launch_fint(const T* dd, const T* vv, T* ff) {
    if (cart_mode) {
      StripIO<T> io = strip_io(dd, ff);
      plain_tensor_map(io, dd);
      if (strip_eta.n) {  // forces from d + eta*v, element by element (solver.f90:293-295, mat_kelvin_voigt.f90:147)
        io.eta = strip_eta.p;
        io.v_kv = vv;
      }
      launch_strips(io);
      return;
    }
    if (variant == S2D_ASM_PATCH) {
      launch_patch(dd, vv, ff);
      return;
    }
    ElemArgs<T> A{};
    A.ibool = ibool.p;
    A.a = coef.p;
    A.elem2set = elem2set.p;
    A.elem2kv = nkv > 0 ? elem2kv.p : nullptr;
    A.eta = eta.p;
    A.beta = beta.p;
    A.H = H.p;
    A.d = dd;
    A.v = vv;
    A.f = ff;
    A.npoin = npoin;
    A.nelast = nelast;
    A.kd2 = kd2;
    if (variant == S2D_ASM_ATOMIC) {
      A.elist = nullptr;
      A.ne = nelem;
      if (ndof == 1) launch_elem_node<T, 1, true>(ngll, A, stream);
      else launch_elem_node<T, 2, true>(ngll, A, stream);
      launches++;
    } else {
      for (int c = 0; c < ncolors; ++c) {
        A.elist = color_elems.p + color_start[c];
        A.ne = color_start[c + 1] - color_start[c];
        if (ndof == 1) launch_elem_node<T, 1, false>(ngll, A, stream);
        else launch_elem_node<T, 2, false>(ngll, A, stream);
        launches++;
      }
    }
    S2D_CUDA(cudaGetLastError());
  }

  template <int N>
  void launch_patch_n(const T* dd, const T* vv, T* ff) {
    PatchArgs<T, N> A{};
    A.npatch = pp_npatch;
    A.EP = pp_EP;
    A.max_nloc = pp_max_nloc;
    A.max_colors = pp_max_colors;
    A.pelem_start = p_pelem_start.p;
    A.pshape = p_pshape.p;
    A.sh_lidx = p_sh_lidx.p;
    A.sh_ecolor = p_sh_ecolor.p;
    A.sh_slot = p_sh_slot.p;
    A.pslot_base = p_pslot_base.p;
    A.pnode_start = p_pnode_start.p;
    A.pnode = p_pnode.p;
    A.eset = p_eset.p;
    A.ekv = nkv > 0 ? p_ekv.p : nullptr;
    A.a = p_hetero ? p_coef.p : coef.p;
    A.eta = eta.p;
    A.beta = beta.p;
    A.d = dd;
    A.v = vv;
    A.f = ff;
    A.fhalo = fhalo.p;
    A.npoin = npoin;
    A.nslots = pp_nslots;
    A.nelast = nelast;
    A.kd2 = kd2;
    A.hetero = p_hetero ? 1 : 0;
    A.pf_dist = pf_dist;
    A.use_bulk = use_bulk;
    for (int k = 0; k < N * N; ++k) A.H[k] = (T)h_H[k];
    if (ndof == 1) launch_elem_patch_n<T, N, 1>(A, stream);
    else launch_elem_patch_n<T, N, 2>(A, stream);
    launches++;
    const int ns = (int)plan.snode.size();
    if (ns > 0) {
      k_halo_sum<T><<<ceil_div(ns, 256), 256, 0, stream>>>(ff, fhalo.p, p_snode.p, p_sstart.p, ns, npoin,
                                                           pp_nslots, ndof);
      launches++;
    }
    S2D_CUDA(cudaGetLastError());
  }
  void launch_patch(const T* dd, const T* vv, T* ff) {
    switch (ngll) {
#ifndef S2D_ONLY_N5
      case 3: launch_patch_n<3>(dd, vv, ff); break;
      case 4: launch_patch_n<4>(dd, vv, ff); break;
#endif
      case 5: launch_patch_n<5>(dd, vv, ff); break;
#ifndef S2D_ONLY_N5
      case 6: launch_patch_n<6>(dd, vv, ff); break;
      case 7: launch_patch_n<7>(dd, vv, ff); break;
      case 8: launch_patch_n<8>(dd, vv, ff); break;
      case 9: launch_patch_n<9>(dd, vv, ff); break;
      case 10: launch_patch_n<10>(dd, vv, ff); break;
#endif
      default: throw ArgError("ngll must be in 3..10");
    }
  }

  void launch_bcs(const T* D) {
    const T* V = v.p;
    T* f = a.p;
    // bc_gen.f90:271-281: periodic boundaries first, then absorbing, then the others in input order
    for (auto& b : perio) {
      k_periodic<T><<<ceil_div(b->np, 128), 128, 0, stream>>>(f, npoin, ndof, b->np, b->master.p, b->slave.p);
      launches++;
    }
    if (node_ops_ready) {  // SO_add + every BC_ABSO_apply, one launch (k_node_ops)
      k_node_ops<T><<<ceil_div(no_nnodes, 128), 128, 0, stream>>>(
          no_nnodes, no_node.p, no_start.p, no_ops.p, no_absos.p, D, V, f, npoin, ndof, st_src.p, st_coef.p, st_nterms,
          (int)h_src_iglob.size(), src_ampli.p, ctl.p);
      launches++;
    } else {
      for (auto& r : bc_order)
        if (r.kind == BC_ABSO) {
          AbsoBc& b = *abso[r.index];
          k_abso<T><<<ceil_div(b.dev.np, 128), 128, 0, stream>>>(b.dev, D, V, f, npoin);
          launches++;
        }
    }
    for (auto& r : bc_order) {
      if (r.kind == BC_DIRNEU) {
        DirneuBc& b = *dirneu[r.index];
        const bool has_stf = b.B_h.p || b.B_v.p;
        if (b.kind_h == 2 || b.kind_v == 2 || has_stf) {
          k_dirneu<T><<<ceil_div(b.np, 128), 128, 0, stream>>>(f, npoin, ndof, b.np, b.node.p, b.kind_h, b.kind_v,
                                                               b.B_h.p, b.B_v.p, bc_ampli.p, n_neumann_slots,
                                                               b.slot, ctl.p);
          launches++;
        }
      } else if (r.kind == BC_DYNFLT) {
        FaultBc& b = *faults[r.index];
        k_dynflt<T><<<ceil_div(b.dev.np, 64), 64, 0, stream>>>(b.dev, f, V, D, npoin, ctl.p);
        launches++;
      }
    }
  }

  // SO_add (src_gen.f90:290-317): the terms of every source (one for a point force, src_force.f90:84; those of
  // SRC_MOMENT_add, src_moment.f90:183-197, for a moment tensor) in source order, grouped by target node
  void build_source_terms() {
    const int ns = (int)h_src_iglob.size();
    if (ns == 0) return;
    struct Term {
      int node, src;
      double c[2];
    };
    std::vector<Term> terms;
    size_t m = 0;
    for (int s = 0; s < ns; ++s) {
      if (h_src_iglob[s] > 0) {
        terms.push_back({h_src_iglob[s], s, {ndof == 1 ? 1.0 : h_src_dir[2 * s], ndof == 1 ? 0.0 : h_src_dir[2 * s + 1]}});
      } else {  // the m-th moment source
        const int t0 = h_mom_start[m], nt = h_mom_start[m + 1] - t0;
        for (int t = 0; t < nt; ++t)
          terms.push_back({h_mom_node[t0 + t], s,
                           {h_mom_coef[(size_t)t0 * ndof + t], ndof == 2 ? h_mom_coef[(size_t)t0 * ndof + t + nt] : 0.0}});
        ++m;
      }
    }
    std::stable_sort(terms.begin(), terms.end(), [](const Term& a, const Term& b) { return a.node < b.node; });
    const int nt = (int)terms.size();
    std::vector<int> node, start, src(nt);
    std::vector<double> coef((size_t)nt * ndof);
    for (int t = 0; t < nt; ++t) {
      if (t == 0 || terms[t].node != terms[t - 1].node) {
        node.push_back(terms[t].node);
        start.push_back(t);
      }
      src[t] = terms[t].src;
      for (int c = 0; c < ndof; ++c) coef[t + (size_t)nt * c] = terms[t].c[c];
    }
    start.push_back(nt);
    st_nnodes = (int)node.size();
    st_nterms = nt;
    st_node.upload(node);
    st_start.upload(start);
    st_src.upload(src);
    st_coef.upload(coef);
  }
  // Sources and absorbing boundaries of a step as per-node operation lists (k_node_ops).  Not with periodic
  // boundaries (they go between the two, bc_gen.f90:271-281), not for multi-stage schemes (one source row per stage,
  // no boundary conditions), at most 8 absorbing boundaries.
  bool node_ops_ready = false;
  int no_nnodes = 0;
  DevBuf<int> no_node, no_start, no_ops;
  DevBuf<AbsoDev> no_absos;
  void build_node_ops() {
    node_ops_ready = false;
    if (abso.empty() || !perio.empty() || scheme.kind == 3 || abso.size() > 8 || env_int("S2D_NODE_OPS", 1) == 0) return;
    struct Op {
      int node, op;
    };
    std::vector<Op> all;
    // source terms first, in source order (the table of build_source_terms is sorted by node, stable)
    {
      std::vector<int> tnode = st_node.n ? st_node.to_host() : std::vector<int>();
      std::vector<int> tstart = st_start.n ? st_start.to_host() : std::vector<int>();
      for (int k = 0; k < st_nnodes; ++k)
        for (int t = tstart[k]; t < tstart[k + 1]; ++t) all.push_back({tnode[k], -t - 1});
    }
    std::vector<AbsoDev> devs;
    std::vector<std::vector<int>> nodes;
    for (auto& r : bc_order)
      if (r.kind == BC_ABSO) {
        devs.push_back(abso[r.index]->dev);
        nodes.push_back(abso[r.index]->node.to_host());
      }
    for (size_t b = 0; b < devs.size(); ++b)
      for (int k = 0; k < devs[b].np; ++k) all.push_back({nodes[b][k], (k << 3) | (int)b});
    std::stable_sort(all.begin(), all.end(), [](const Op& x, const Op& y) { return x.node < y.node; });
    std::vector<int> node, start, ops(all.size());
    for (size_t q = 0; q < all.size(); ++q) {
      if (q == 0 || all[q].node != all[q - 1].node) {
        node.push_back(all[q].node);
        start.push_back((int)q);
      }
      ops[q] = all[q].op;
    }
    start.push_back((int)all.size());
    no_nnodes = (int)node.size();
    no_node.upload(node);
    no_start.upload(start);
    no_ops.upload(ops);
    no_absos.upload(devs);
    node_ops_ready = true;
  }
  void launch_sources(T* f, int stage = 0) {
    if (st_nnodes == 0) return;
    k_sources<T><<<ceil_div(st_nnodes, 64), 64, 0, stream>>>(f, npoin, ndof, st_nnodes, st_node.p, st_start.p, st_src.p,
                                                            st_coef.p, st_nterms, (int)h_src_iglob.size(), src_ampli.p,
                                                            ctl.p, stage, nstages());
    launches++;
  }

  void launch_outputs() {
    const T* fld = rec.field == 'D' ? dn() : (rec.field == 'V' ? v.p : a.p);
    const int nrb = rec.present ? ceil_div((long long)rec.dev.nx * ndof, DYNW_THREADS) : 0;
    bool rec_done = !rec.present;
    for (auto& b : faults) {  // REC_store rides with the first fault's BC_DYNFLT_write: one launch for the outputs
      const int extra = rec_done ? 0 : nrb;
      k_dynflt_write<T><<<b->wctas + extra, DYNW_THREADS, 0, stream>>>(b->dev, dn(), v.p, npoin, ctl.p, b->wpart.p,
                                                                       b->wticket.p, b->wctas, rec.dev, fld);
      rec_done = true;
      launches++;
    }
    if (!rec_done) {
      k_rec_store<T><<<ceil_div((long long)rec.dev.nx * ndof, 128), 128, 0, stream>>>(rec.dev, fld, npoin, ctl.p);
      launches++;
    }
  }

  // one leapfrog step with the node update fused into the strip kernel (strip_kernels.cuh)
  void launch_step_fused(bool last_of_call) {
    const size_t nd = npoin * ndof;
    const T dt = (T)scheme.dt;
    phase(PH_PRED);
    const bool nmk = scheme.kind == 1;
    const T c1 = nmk ? (T)(0.5 * scheme.dt * scheme.dt) : (T)0;            // (1/2 - beta) dt^2, beta = 0
    const T c2 = nmk ? (T)((1.0 - scheme.gamma) * scheme.dt) : (T)0;
    const T c3 = nmk ? (T)(scheme.gamma * scheme.dt) : dt;
    if (!pred_valid) {  // first step after the fields were set: the predictor (solver.f90:151 / :59)
      k_predict_to<T><<<grid_for(nd), 256, 0, stream>>>(dalt(), dn(), v.p, a.p, nd, dt, c1);
      launches++;
    }
    T* dc = dalt();   // d[n]
    T* dnx = dn();    // receives the prediction d[n+1]; d[n-1] is dead
    const bool kvf = strip_eta.n != 0;  // Kelvin-Voigt: v (Newmark: and a) go to the second buffer, swapped below
    StripIO<T> io = strip_io(dc, kvf && nmk ? a_alt.p : a.p);
    io.v_in = v.p;
    io.v_out = kvf ? v_alt.p : v.p;
    if (kvf) io.eta = strip_eta.p;
    io.rmass = rmass.p;
    io.d_next = dnx;
    // accelerations are written when somebody will read them: the explicit Newmark scheme (next step), receivers
    // recording 'A', S2D_STORE_ACCEL=1.  Otherwise (leapfrog) they are formed on demand -- ensure_accel(), at
    // s2d_get_fields / s2d_cart_get_window -- from one plain force evaluation of the displacement the step used: no
    // 8 B/DOF of stores per s2d_step call.  Where that evaluation is not repeatable (state advanced by every
    // evaluation: the stateful rheologies; Kelvin-Voigt: v already updated), the last step of every call stores
    // them as before.
    const bool lazy_ok = store_accel == 2 && accel_lazy && !nmk && strip_eta.n == 0 && pl_set.n == 0;  // pl_set: any stateful rheology
    const bool want_a = nmk || store_accel == 1 ||
                        (store_accel == 2 && ((last_of_call && !lazy_ok) || (rec.present && rec.field == 'A')));
    a_stale = !want_a;
    io.a_out = want_a ? io.f : nullptr;
    io.rowflag = rowflag.p;
    io.colflag = colflag.p;
    io.dt = scheme.dt;
    io.newmark = nmk ? 1 : 0;
    io.c1 = c1;
    io.c2 = c2;
    io.c3 = c3;
    io.a_in = a.p;
    if (tm_ready && !kvf) {  // d[n] lives in the buffer the caller does not see at this point
      io.tm_d = &tm_d[dc == d.p ? 0 : 1];
      io.tm_v = &tm_v;
      io.tm_r = &tm_r;
      io.tm_a = &tm_a;
    }
    launch_strips(io, ctl.p);  // ... and the step counter (k_strip_fold)
    if (kvf) {  // from here on v / a name the buffers this step wrote
      std::swap(v, v_alt);
      if (nmk) std::swap(a, a_alt);
    }
    phase(PH_SRC);
    if (!node_ops_ready) launch_sources(a.p);
    phase(PH_BC);
    launch_bcs(dc);
    phase(PH_UPDATE);
    const long long nw = (long long)ndrows * cart_S.LX + (long long)ndcols * cart_S.LZ;
    if (nw > 0) {
      k_strip_deferred<T><<<(unsigned)((nw + 255) / 256), 256, 0, stream>>>(
          cart_S.LX, cart_S.LXP, cart_S.LZ, ndof, npoin, drows.p, ndrows, dcols.p, ndcols, rowflag.p, a.p, v.p, rmass.p, dc, dnx, dt,
          c1, c3);
      launches++;
    }
    dsel ^= 1;  // the caller now sees d[n]; the other buffer holds the prediction
    pred_valid = true;
    phase(PH_OUT);
    launch_outputs();
    phase(PH_COUNT);
  }

  void launch_step(bool last_of_call = true) {
    if (fused) {
      launch_step_fused(last_of_call);
      return;
    }
    const size_t nd = npoin * ndof;
    const T dt = (T)scheme.dt;
    phase(PH_PRED);
    k_tick<<<1, 1, 0, stream>>>(ctl.p);
    launches++;
    const int zf = needs_zero_f() ? 1 : 0;
    if (scheme.kind == 3) {  // solve_symplectic (solver.f90:169-199): no boundary conditions, as in the reference
      for (int k = 0; k < scheme.nstages; ++k) {
        k_axpy<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, nd, (T)(scheme.dt * scheme.coa[k]), zf);
        launch_fint(dn(), v.p, a.p);
        launch_sources(a.p, k);
        k_correct<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, rmass.p, nd, (T)(scheme.dt * scheme.cob[k]), (T)0);
        launches += 2;
      }
      k_axpy<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, nd, (T)(scheme.dt * scheme.coa[scheme.nstages]), 0);
      launches++;
      launch_outputs();
      return;
    }
    const T* dforce = dn();
    const T* vforce = v.p;
    if (scheme.kind == 0) {
      k_predict_leapfrog<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, nd, dt, zf);
    } else {
      const T c1 = (T)((0.5 - scheme.beta) * scheme.dt * scheme.dt), c2 = (T)((1.0 - scheme.gamma) * scheme.dt);
      if (scheme.kind == 2) {  // solve_HHT_alpha (solver.f90:89-128): forces from the alpha-weighted fields
        k_predict_hht<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, d_alpha.p, v_alpha.p, nd, dt, c1, c2,
                                                           (T)scheme.alpha, zf);
        dforce = d_alpha.p;
        vforce = v_alpha.p;
      } else {
        k_predict_newmark<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, nd, dt, c1, c2, zf);
      }
    }
    launches++;
    if (!cart_mode) phase(PH_FORCE);
    launch_fint(dforce, vforce, a.p);
    phase(PH_SRC);
    if (!node_ops_ready) launch_sources(a.p);
    phase(PH_BC);
    launch_bcs(dn());  // BC_apply sees fields%displ / veloc (solver.f90:121), not the alpha-weighted copies
    phase(PH_UPDATE);
    T c3, c4;
    if (scheme.kind == 0) {
      c3 = dt;
      c4 = 0;
    } else {
      c3 = (T)(scheme.gamma * scheme.dt);
      c4 = (T)(scheme.beta * scheme.dt * scheme.dt);
    }
    k_correct<T><<<grid_for(nd), 256, 0, stream>>>(dn(), v.p, a.p, rmass.p, nd, c3, c4);
    launches++;
    phase(PH_OUT);
    launch_outputs();
    phase(PH_COUNT);
  }

  // Device-side aborts (ctl.err) are reported once: the flag is cleared, so that the next call is judged on its
  // own.  After S2D_ESOLVER the engine state is undefined (the reference stops the program, bc_dynflt_rsf.f90:463-466).
  void check_device_error() {
    StepCtl c;
    S2D_CUDA(cudaMemcpyAsync(&c, ctl.p, sizeof(c), cudaMemcpyDeviceToHost, stream));
    S2D_CUDA(cudaStreamSynchronize(stream));
    if (c.err == 0) return;
    const int zero = 0;
    S2D_CUDA(cudaMemcpyAsync(&ctl.p->err, &zero, sizeof(int), cudaMemcpyHostToDevice, stream));
    S2D_CUDA(cudaStreamSynchronize(stream));
    if (c.err == 1) throw SolverError("NR_Solver has exceeded the maximum iterations (200)");
    if (c.err == 2) throw SolverError("NR_Solver could not bracket a root");
    if (c.err == 3) throw StateError("halo exchange: a neighbour GPU did not signal within 10 s");
    if (c.err == 4) throw SolverError("MAT_DMG: damage exceeded critical value");
    throw StateError("device-side abort, code " + std::to_string(c.err));
  }

  void step(int nsteps, const double* srca, const double* bca) override {
    S2D_REQUIRE(committed, "step before commit");
    S2D_REQUIRE(nsteps >= 0, "nsteps < 0");
    if (nsteps == 0) return;
    const size_t ns = h_src_iglob.size();
    if (ns > 0) {
      S2D_REQUIRE(srca != nullptr, "step: src_ampli missing");
      const size_t need = ns * nsteps * nstages();
      if (need > src_ampli_cap) {
        S2D_CUDA(cudaStreamSynchronize(stream));
        src_ampli.alloc(need);
        src_ampli_cap = need;
      }
      S2D_CUDA(cudaStreamSynchronize(stream));  // the staging buffer of the previous call is free again
      std::memcpy(pin_src.need(need * sizeof(double)), srca, need * sizeof(double));
      S2D_CUDA(cudaMemcpyAsync(src_ampli.p, pin_src.p, need * sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    bool any_stf = false;
    for (auto& b : dirneu) any_stf = any_stf || b->B_h.p || b->B_v.p;
    if (any_stf) {
      S2D_REQUIRE(bca != nullptr, "step: bc_ampli missing");
      const size_t need = (size_t)n_neumann_slots * nsteps;
      if (need > bc_ampli_cap) {
        S2D_CUDA(cudaStreamSynchronize(stream));
        bc_ampli.alloc(need);
        bc_ampli_cap = need;
      }
      S2D_CUDA(cudaStreamSynchronize(stream));
      std::memcpy(pin_bc.need(need * sizeof(double)), bca, need * sizeof(double));
      S2D_CUDA(cudaMemcpyAsync(bc_ampli.p, pin_bc.p, need * sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    const int it0 = it + 1;
    S2D_CUDA(cudaMemcpyAsync(&ctl.p->it0, &it0, sizeof(int), cudaMemcpyHostToDevice, stream));
    const int nrows = nsteps * nstages();  // one row per stage for the symplectic schemes
    S2D_CUDA(cudaMemcpyAsync(&ctl.p->nrows, &nrows, sizeof(int), cudaMemcpyHostToDevice, stream));
    for (int k = 0; k < nsteps; ++k) launch_step(k == nsteps - 1);
    it += nsteps;
    S2D_CUDA(cudaGetLastError());
    check_device_error();
  }

  void compute_fint(double* out) override {
    S2D_REQUIRE(committed, "compute_fint before commit");
    const size_t nd = npoin * ndof;
    if (scratch.n != nd) scratch.alloc(nd);
    scratch.zero(stream);
    launch_fint(dn(), v.p, scratch.p);
    download(scratch, out);
  }

  // ---- field transfer ------------------------------------------------------------------
  void upload_field(DevBuf<T>& dst, const double* src) {
    if (!src) return;
    const size_t nd = npoin * ndof;
    S2D_CUDA(cudaStreamSynchronize(stream));
    if (cart_mode) {
      DevBuf<double> tmp;
      tmp.upload(src, npoin_ref * ndof);
      cart_from_ref(tmp.p, dst.p);
      S2D_CUDA(cudaStreamSynchronize(stream));
      return;
    }
    if (sizeof(T) == 8) {
      h2d_sync(dst.p, src, nd * 8);
    } else {
      std::vector<T> tmp(nd);
      for (size_t q = 0; q < nd; ++q) tmp[q] = (T)src[q];
      h2d_sync(dst.p, tmp.data(), nd * sizeof(T));
    }
  }
  void download(const DevBuf<T>& src, double* dst) {
    if (!dst) return;
    const size_t nd = npoin * ndof;
    S2D_CUDA(cudaStreamSynchronize(stream));
    if (cart_mode) {
      DevBuf<double> tmp;
      tmp.alloc(npoin_ref * ndof);
      cart_to_ref(src.p, tmp.p);
      S2D_CUDA(cudaStreamSynchronize(stream));
      tmp.download(dst);
      return;
    }
    if (sizeof(T) == 8) {
      S2D_CUDA(cudaMemcpy(dst, src.p, nd * 8, cudaMemcpyDeviceToHost));
    } else {
      std::vector<T> tmp(nd);
      S2D_CUDA(cudaMemcpy(tmp.data(), src.p, nd * sizeof(T), cudaMemcpyDeviceToHost));
      for (size_t q = 0; q < nd; ++q) dst[q] = (double)tmp[q];
    }
  }
  void set_fields(const double* dd, const double* vv, const double* aa) override {
    if (dd || vv || aa) pred_valid = false;  // the cached prediction d + dt*v (+ dt^2/2 a) is stale
    if (aa) a_stale = false;
    else if (dd) ensure_accel();  // a new displacement: form the accelerations of the last step while they still can be
    upload_field(dbuf(), dd);
    upload_field(v, vv);
    upload_field(a, aa);
  }
  // accelerations of the nodes the fused leapfrog kernel advanced without storing them: a = rmass * Fint(d[n])
  // (deferred nodes -- boundary conditions, sources, band rows -- always have theirs stored by k_strip_deferred)
  void ensure_accel() {
    if (!a_stale) return;
    S2D_REQUIRE(cart_mode && fused, "ensure_accel: fused structured engines only");
    const size_t nd = npoin * ndof;
    if (scratch.n != nd) scratch.alloc(nd);
    StripIO<T> io = strip_io(dbuf().p, scratch.p);
    plain_tensor_map(io, dbuf().p);
    // no interface exchange on x-strips: the interface columns are deferred nodes (their accelerations are stored),
    // every other node's force is complete with this GPU's elements alone -- so this is not a collective call
    launch_elem_strip_items<T>(strip_all_groups(cart_S), io, stream);
    launches += 1 + launch_strip_fold<T>(cart_S, io.f, cart_hx.p, cart_hz.p, npoin, stream, nullptr);
    const long long tot = (long long)cart_S.LX * cart_S.LZ;
    k_strip_accel_fill<T><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(cart_S.LX, cart_S.LXP, cart_S.LZ, ndof, npoin, rowflag.p,
                                                                        colflag.p, scratch.p, rmass.p, a.p);
    launches++;
    S2D_CUDA(cudaGetLastError());
    S2D_CUDA(cudaStreamSynchronize(stream));
    a_stale = false;
  }

  void get_fields(double* dd, double* vv, double* aa) override {
    if (aa) ensure_accel();
    download(dbuf(), dd);
    download(v, vv);
    download(a, aa);
  }

  void get_seis(float* sis) override {
    S2D_REQUIRE(rec.present, "get_seis: no receivers");
    S2D_CUDA(cudaStreamSynchronize(stream));
    rec.sis.download(sis);
  }
  void get_seis_row(int it_, float* row) override {
    S2D_REQUIRE(rec.present, "get_seis_row: no receivers");
    S2D_REQUIRE(it_ >= 0 && it_ % rec.dev.isamp == 0 && it_ / rec.dev.isamp < rec.dev.nt, "get_seis_row: step not sampled");
    const size_t r = (size_t)(it_ / rec.dev.isamp);
    const size_t nrow = (size_t)rec.dev.nx * ndof;
    float* stage = (float*)pin_row.need(nrow * sizeof(float));
    S2D_CUDA(cudaMemcpy2DAsync(stage, sizeof(float), rec.sis.p + r, (size_t)rec.dev.nt * sizeof(float), sizeof(float),
                               nrow, cudaMemcpyDeviceToHost, stream));
    S2D_CUDA(cudaStreamSynchronize(stream));
    std::memcpy(row, stage, nrow * sizeof(float));
  }
  void get_fault(int id, float* records, int32_t* nout, double* potency, int32_t* ncalls) override {
    S2D_REQUIRE(id >= 0 && id < (int)faults.size(), "get_fault: bad fault id");
    FaultBc& b = *faults[id];
    S2D_CUDA(cudaStreamSynchronize(stream));
    std::vector<int> ost = b.ostate.to_host();
    if (nout) *nout = ost[1];
    if (ncalls) *ncalls = ost[2];
    if (records && ost[1] > 0)
      S2D_CUDA(cudaMemcpy(records, b.records.p, (size_t)ost[1] * 6 * b.dev.onx * sizeof(float), cudaMemcpyDeviceToHost));
    if (potency && ost[2] > 0)
      S2D_CUDA(cudaMemcpy(potency, b.potency.p, (size_t)ost[2] * 2 * (ndof + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  }
  void get_fault_state(int id, double* D, double* V, double* T_, double* Tstick, double* MU, double* theta,
                       double* sigma) override {
    S2D_REQUIRE(id >= 0 && id < (int)faults.size(), "get_fault_state: bad fault id");
    FaultBc& b = *faults[id];
    S2D_CUDA(cudaStreamSynchronize(stream));
    if (D) b.D.download(D);
    if (V) b.V.download(V);
    if (T_) b.T.download(T_);
    if (Tstick) b.Tstick.download(Tstick);
    if (MU) b.MU.download(MU);
    if (sigma) b.sigma.download(sigma);
    if (theta) {
      if (b.dev.rsf_kind) b.rsf_theta.download(theta);
      else if (b.dev.swf_kind) b.swf_theta.download(theta);
      else std::fill(theta, theta + b.dev.np, 0.0);
    }
  }

  double reduce_absmax(const DevBuf<T>& x) {
    const int g = std::min<int>(1024, grid_for(x.n));
    k_absmax<T><<<g, 256, 0, stream>>>(x.p, x.n, partial.p);
    launches++;
    std::vector<double> h(g);
    S2D_CUDA(cudaMemcpyAsync(h.data(), partial.p, g * sizeof(double), cudaMemcpyDeviceToHost, stream));
    S2D_CUDA(cudaStreamSynchronize(stream));
    double m = 0;
    for (double x_ : h) m = std::max(m, x_);
    return m;
  }
  void progress(double* vmax, double* dmax) override {
    if (vmax) *vmax = reduce_absmax(v);
    if (dmax) *dmax = reduce_absmax(dbuf());
  }
  double energy() override {
    S2D_REQUIRE(mass.n == npoin, "energy: s2d_set_mass was not called");
    const int g = std::min<int>(1024, grid_for(npoin));
    k_kinetic<T><<<g, 256, 0, stream>>>(v.p, mass.p, npoin, ndof, partial.p);
    launches++;
    std::vector<double> h(g);
    S2D_CUDA(cudaMemcpyAsync(h.data(), partial.p, g * sizeof(double), cudaMemcpyDeviceToHost, stream));
    S2D_CUDA(cudaStreamSynchronize(stream));
    double s = 0;
    for (double x_ : h) s += x_;
    return 0.5 * s;
  }
  // E_W = 1/2 sum(beta * d.d) (energy.f90:84-104), the elastic energy of the 2.5D term; 0 when W is infinite
  double energy_w25d() override {
    if (!strip_beta.n) {
      S2D_REQUIRE(h_beta.empty(), "energy_w25d: available on the strip-kernel route (structured boxes)");
      return 0.0;
    }
    const int g = 512;
    k_strip_energy_w<T><<<g, 256, 0, stream>>>(cart_S, strip_beta.p, dbuf().p, npoin, partial.p);
    launches++;
    std::vector<double> h(g);
    S2D_CUDA(cudaMemcpyAsync(h.data(), partial.p, g * sizeof(double), cudaMemcpyDeviceToHost, stream));
    S2D_CUDA(cudaStreamSynchronize(stream));
    double s = 0;
    for (double x_ : h) s += x_;
    return 0.5 * s;
  }
  void get_coloring(int32_t* nc, int32_t* color) override {
    S2D_REQUIRE(committed, "get_coloring before commit");
    S2D_REQUIRE(have_colors, "get_coloring: not available for builder-made meshes");
    if (nc) *nc = ncolors;
    if (color) std::copy(h_color.begin(), h_color.end(), color);
  }

  float kernel_ms() override { return last_kernel_ms; }
  int route() override { return cart_mode ? 1 : 0; }
  void halo_info(int64_t* count, void** send_dev, void** recv_dev) override {
    S2D_REQUIRE(cart_mode, "halo_info: only x-strips made by the structured builder have halos");
    xhalo_setup();
    if (count) *count = (int64_t)cart_S.LZ * ndof;
    for (int sd = 0; sd < 2; ++sd) {
      if (send_dev) send_dev[sd] = xh_send[sd].p;
      if (recv_dev) recv_dev[sd] = xh_recv[sd].p;
    }
  }
  void halo_peer_buffers(void** recv_dev, void** flags_dev) override {
    S2D_REQUIRE(cart_mode, "halo_peer_buffers: only x-strips made by the structured builder have halos");
    xhalo_setup();
    for (int sd = 0; sd < 2; ++sd)
      if (recv_dev) recv_dev[sd] = xh_recv[sd].p;
    if (flags_dev) *flags_dev = xh_flags.p;
  }
  void halo_set_peers(void* left_recv, void* right_recv, void* left_flag, void* right_flag) override {
    S2D_REQUIRE(cart_mode, "halo_set_peers: only x-strips made by the structured builder have halos");
    xhalo_setup();
    S2D_REQUIRE(!cart_S.xhalo_left || (left_recv && left_flag), "halo_set_peers: the left neighbour's buffers are missing");
    S2D_REQUIRE(!cart_S.xhalo_right || (right_recv && right_flag), "halo_set_peers: the right neighbour's buffers are missing");
    xh_peer_recv[0] = cart_S.xhalo_left ? (T*)left_recv : nullptr;
    xh_peer_recv[1] = cart_S.xhalo_right ? (T*)right_recv : nullptr;
    xh_peer_flag[0] = cart_S.xhalo_left ? (unsigned long long*)left_flag : nullptr;
    xh_peer_flag[1] = cart_S.xhalo_right ? (unsigned long long*)right_flag : nullptr;
    xh_peer = true;
  }
  void halo_set_exchange(s2d_exchange_fn fn, void* user) override {
    S2D_REQUIRE(cart_mode, "halo_set_exchange: only x-strips made by the structured builder have halos");
    xhalo_setup();
    xh_fn = fn;
    xh_user = user;
  }

  float time_fint(int reps) override {
    S2D_REQUIRE(committed && reps > 0, "time_fint: not committed or reps <= 0");
    const size_t nd = npoin * ndof;
    if (scratch.n != nd) scratch.alloc(nd);
    scratch.zero(stream);
    cudaEvent_t e0, e1;
    S2D_CUDA(cudaEventCreate(&e0));
    S2D_CUDA(cudaEventCreate(&e1));
    launch_fint(dn(), v.p, scratch.p);  // warm
    S2D_CUDA(cudaStreamSynchronize(stream));
    S2D_CUDA(cudaEventRecord(e0, stream));
    for (int r = 0; r < reps; ++r) launch_fint(dn(), v.p, scratch.p);
    S2D_CUDA(cudaEventRecord(e1, stream));
    S2D_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    S2D_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / reps;
  }
  // average milliseconds per step of every phase (see the PH_* list) over nsteps steps
  void time_phases(int nsteps, float* ms) override {
    S2D_REQUIRE(committed && nsteps > 0 && ms, "time_phases: not committed, nsteps <= 0 or null pointer");
    S2D_REQUIRE(h_src_iglob.empty() || src_ampli_cap > 0, "time_phases: call s2d_step once first to load the stf table");
    S2D_CUDA(cudaStreamSynchronize(stream));
    pev_phase.clear();
    pev_on = true;
    for (int k = 0; k < nsteps; ++k) launch_step(k == nsteps - 1);
    pev_on = false;
    it += nsteps;
    S2D_CUDA(cudaStreamSynchronize(stream));
    for (int q = 0; q < PH_COUNT; ++q) ms[q] = 0.f;
    for (size_t q = 0; q + 1 < pev_phase.size(); ++q) {
      if (pev_phase[q] >= PH_COUNT) continue;
      float t = 0;
      S2D_CUDA(cudaEventElapsedTime(&t, pev[q], pev[q + 1]));
      ms[pev_phase[q]] += t;
    }
    for (int q = 0; q < PH_COUNT; ++q) ms[q] /= (float)nsteps;
    check_device_error();
  }
  float time_steps(int nsteps) override {
    S2D_REQUIRE(committed && nsteps > 0, "time_steps: not committed or nsteps <= 0");
    S2D_REQUIRE(h_src_iglob.empty() || src_ampli_cap > 0, "time_steps: call s2d_step once first to load the stf table");
    cudaEvent_t e0, e1;
    S2D_CUDA(cudaEventCreate(&e0));
    S2D_CUDA(cudaEventCreate(&e1));
    // the stf tables of the last s2d_step call are replayed cyclically (row = (it-it0) mod nrows)
    S2D_CUDA(cudaStreamSynchronize(stream));
    kev_on = cart_mode;
    kev_n = 0;
    S2D_CUDA(cudaEventRecord(e0, stream));
    for (int k = 0; k < nsteps; ++k) launch_step(k == nsteps - 1);
    S2D_CUDA(cudaEventRecord(e1, stream));
    S2D_CUDA(cudaEventSynchronize(e1));
    kev_on = false;
    it += nsteps;
    float ms = 0;
    S2D_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    last_kernel_ms = 0.f;
    for (size_t q = 0; q + 1 < kev_n; q += 2) {
      float t = 0;
      S2D_CUDA(cudaEventElapsedTime(&t, kev[q], kev[q + 1]));
      last_kernel_ms += t;
    }
    if (kev_n >= 2) last_kernel_ms /= (float)(kev_n / 2);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    check_device_error();
    return ms;
  }
};

// One translation unit per precision instantiates the engine (engine.cu: FP64, engine_f32.cu: FP32); every
// other one only refers to it.
#ifndef S2D_INSTANTIATE_F64
extern template class Engine<double>;
#endif
#ifndef S2D_INSTANTIATE_F32
extern template class Engine<float>;
#endif

}  // namespace s2d
