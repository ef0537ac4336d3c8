// C-ABI of the sem2d_b200 engine (include/sem2d_b200.h).
#define S2D_INSTANTIATE_F64
#include "engine.hpp"
#include "rcm_box.hpp"

using namespace s2d;
template class s2d::Engine<double>;

struct s2d_engine {
  std::unique_ptr<EngineBase> impl;
  std::string err;
  void* cart = nullptr;  // builder state (cart.cu)
};

namespace s2d {
void cart_free(void* cart);
}

static thread_local std::string g_create_err;

template <typename F>
static int guard(s2d_handle h, F&& fn) {
  if (!h || !h->impl) return S2D_EINVAL;
  try {
    S2D_CUDA(cudaSetDevice(h->impl->device));
    fn(*h->impl);
    return S2D_OK;
  } catch (const ArgError& e) {
    h->err = e.what();
    return S2D_EINVAL;
  } catch (const SolverError& e) {
    h->err = e.what();
    return S2D_ESOLVER;
  } catch (const StateError& e) {
    h->err = e.what();
    return S2D_ESTATE;
  } catch (const CudaError& e) {
    h->err = e.what();
    return S2D_ECUDA;
  } catch (const std::bad_alloc&) {
    h->err = "out of host memory";
    return S2D_ENOMEM;
  } catch (const std::exception& e) {
    h->err = e.what();
    return S2D_ECUDA;
  }
}

namespace s2d {
// used by s2d_create and by the structured builder
int select_device(int device, std::string& err) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    err = std::string("no usable CUDA device: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return -1;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev) {
    err = "device index out of range";
    return -1;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    err = "cudaSetDevice failed";
    return -1;
  }
  return device;
}
s2d_handle wrap_engine(std::unique_ptr<EngineBase> impl, void* cart) {
  s2d_engine* h = new s2d_engine();
  h->impl = std::move(impl);
  h->cart = cart;
  return h;
}
void* engine_cart(s2d_handle h) { return h->cart; }
EngineBase* engine_impl(s2d_handle h) { return h->impl.get(); }
void set_error(s2d_handle h, const std::string& m) { h->err = m; }
}  // namespace s2d

extern "C" {

const char* s2d_version(void) { return "sem2d_b200 0.1 (sm_100a)"; }

const char* s2d_last_error(s2d_handle h) {
  if (!h) return g_create_err.c_str();
  return h->err.c_str();
}

int s2d_create(s2d_handle* out, int32_t ngll, int32_t ndof, int32_t nelem, int32_t npoin,
               const int32_t* ibool, const double* hprime, const double* rmass, int32_t precision,
               const s2d_scheme* scheme, int32_t device) {
  if (!out) return S2D_EINVAL;
  *out = nullptr;
  if (!ibool || !hprime || !rmass || !scheme || ngll < 3 || ngll > 10 || (ndof != 1 && ndof != 2) ||
      nelem < 1 || npoin < 1 || (precision != 8 && precision != 4) || scheme->kind < 0 || scheme->kind > 3 ||
      !(scheme->dt > 0.0)) {
    g_create_err = "s2d_create: invalid argument";
    return S2D_EINVAL;
  }
  int dev = select_device(device, g_create_err);
  if (dev < 0) return S2D_ENODEV;
  try {
    std::unique_ptr<EngineBase> impl;
    if (precision == 8)
      impl.reset(new Engine<double>(ngll, ndof, nelem, (size_t)npoin, ibool, hprime, rmass, *scheme, dev));
    else
      impl.reset(new Engine<float>(ngll, ndof, nelem, (size_t)npoin, ibool, hprime, rmass, *scheme, dev));
    *out = wrap_engine(std::move(impl), nullptr);
    return S2D_OK;
  } catch (const ArgError& e) {
    g_create_err = e.what();
    return S2D_EINVAL;
  } catch (const std::exception& e) {
    g_create_err = e.what();
    return S2D_ECUDA;
  }
}

int s2d_destroy(s2d_handle h) {
  if (!h) return S2D_EINVAL;
  if (h->impl) cudaSetDevice(h->impl->device);
  if (h->cart) cart_free(h->cart);
  h->impl.reset();
  delete h;
  return S2D_OK;
}

int s2d_set_elastic(s2d_handle h, int32_t nelast, int32_t ncoefsets, const double* a, const int32_t* elem2set,
                    const double* beta25d, int32_t kd2) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(a && elem2set, "s2d_set_elastic: null pointer");
    E.set_elastic(nelast, ncoefsets, a, elem2set, beta25d, kd2);
  });
}
int s2d_set_kv(s2d_handle h, int32_t nkv, const int32_t* elem_ids, const double* eta) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(nkv == 0 || (elem_ids && eta), "s2d_set_kv: null pointer");
    E.set_kv(nkv, elem_ids, eta);
  });
}
int s2d_set_mass(s2d_handle h, const double* mass) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(mass, "s2d_set_mass: null pointer");
    E.set_mass(mass);
  });
}
int s2d_add_abso(s2d_handle h, int32_t np, const int32_t* node, const double* C, int32_t is_flat, const double* n,
                 int32_t stacey, int32_t nbe, const int32_t* bibool, const double* K) {
  return guard(h, [&](EngineBase& E) { E.add_abso(np, node, C, is_flat, n, stacey, nbe, bibool, K); });
}
int s2d_add_dirneu(s2d_handle h, int32_t np, const int32_t* node, int32_t kind_h, int32_t kind_v,
                   const double* B_h, const double* B_v) {
  return guard(h, [&](EngineBase& E) { E.add_dirneu(np, node, kind_h, kind_v, B_h, B_v); });
}
int s2d_add_dynflt(s2d_handle h, const s2d_dynflt_desc* desc, int32_t* fault_id) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(desc, "s2d_add_dynflt: null descriptor");
    int id = E.add_dynflt(*desc);
    if (fault_id) *fault_id = id;
  });
}
int s2d_add_force(s2d_handle h, int32_t iglob, const double dir[2], int32_t* src_id) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(dir, "s2d_add_force: null dir");
    int id = E.add_force(iglob, dir);
    if (src_id) *src_id = id;
  });
}
int s2d_add_periodic(s2d_handle h, int32_t np, const int32_t* master, const int32_t* slave) {
  return guard(h, [&](EngineBase& E) { E.add_periodic(np, master, slave); });
}
int s2d_add_moment(s2d_handle h, int32_t nterms, const int32_t* node, const double* coef, int32_t* src_id) {
  return guard(h, [&](EngineBase& E) {
    int id = E.add_moment(nterms, node, coef);
    if (src_id) *src_id = id;
  });
}
int s2d_add_receivers(s2d_handle h, int32_t nx, char field, int32_t isamp, int32_t nt_rec, int32_t at_node,
                      const int32_t* iglob, const int32_t* einterp, const double* interp) {
  return guard(h, [&](EngineBase& E) { E.add_receivers(nx, field, isamp, nt_rec, at_node, iglob, einterp, interp); });
}
int s2d_commit(s2d_handle h, int32_t assembly_variant) {
  return guard(h, [&](EngineBase& E) { E.commit(assembly_variant); });
}
int s2d_set_fields(s2d_handle h, const double* d, const double* v, const double* a) {
  return guard(h, [&](EngineBase& E) { E.set_fields(d, v, a); });
}
int s2d_get_fields(s2d_handle h, double* d, double* v, double* a) {
  return guard(h, [&](EngineBase& E) { E.get_fields(d, v, a); });
}
int s2d_step(s2d_handle h, int32_t nsteps, const double* src_ampli, const double* bc_ampli) {
  return guard(h, [&](EngineBase& E) { E.step(nsteps, src_ampli, bc_ampli); });
}
int s2d_compute_fint(s2d_handle h, double* fint) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(fint, "s2d_compute_fint: null pointer");
    E.compute_fint(fint);
  });
}
int s2d_get_it(s2d_handle h, int32_t* it) {
  return guard(h, [&](EngineBase& E) {
    if (it) *it = E.it;
  });
}
int s2d_get_seis(s2d_handle h, float* sis) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(sis, "s2d_get_seis: null pointer");
    E.get_seis(sis);
  });
}
int s2d_get_seis_row(s2d_handle h, int32_t it, float* row) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(row, "s2d_get_seis_row: null pointer");
    E.get_seis_row(it, row);
  });
}
int s2d_get_fault(s2d_handle h, int32_t fault_id, float* records, int32_t* nout, double* potency, int32_t* ncalls) {
  return guard(h, [&](EngineBase& E) { E.get_fault(fault_id, records, nout, potency, ncalls); });
}
int s2d_get_fault_state(s2d_handle h, int32_t fault_id, double* D, double* V, double* T, double* Tstick,
                        double* MU, double* theta, double* sigma) {
  return guard(h, [&](EngineBase& E) { E.get_fault_state(fault_id, D, V, T, Tstick, MU, theta, sigma); });
}
int s2d_progress(s2d_handle h, double* vmax, double* dmax) {
  return guard(h, [&](EngineBase& E) { E.progress(vmax, dmax); });
}
int s2d_energy(s2d_handle h, double* E_k) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(E_k, "s2d_energy: null pointer");
    *E_k = E.energy();
  });
}
int s2d_energy_w25d(s2d_handle h, double* E_W) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(E_W, "s2d_energy_w25d: null pointer");
    *E_W = E.energy_w25d();
  });
}
int s2d_get_coloring(s2d_handle h, int32_t* ncolors, int32_t* color) {
  return guard(h, [&](EngineBase& E) { E.get_coloring(ncolors, color); });
}
int s2d_time_fint(s2d_handle h, int32_t reps, float* ms_avg) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(ms_avg, "s2d_time_fint: null pointer");
    *ms_avg = E.time_fint(reps);
  });
}
int s2d_time_steps(s2d_handle h, int32_t nsteps, float* ms_total) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(ms_total, "s2d_time_steps: null pointer");
    *ms_total = E.time_steps(nsteps);
  });
}
int s2d_detect_structured(int32_t ngll, int32_t nelem, int32_t npoin, const int32_t* ibool, int32_t lower_hint,
                          int32_t* nx, int32_t* nz, int32_t* ezflt, int32_t* ex, int32_t* ez, int32_t* gx, int32_t* gz) {
  if (!ibool || ngll < 3 || ngll > 10 || nelem < 1 || npoin < 1) return S2D_EINVAL;
  for (size_t q = 0; q < (size_t)nelem * ngll * ngll; ++q)
    if (ibool[q] < 1 || ibool[q] > npoin) return S2D_EINVAL;
  StructuredBox B = detect_structured(ibool, ngll, nelem, (size_t)npoin, lower_hint);
  if (nx) *nx = B.ok ? B.nx : 0;
  if (nz) *nz = B.ok ? B.nz : 0;
  if (ezflt) *ezflt = B.ok ? B.ezflt : 0;
  if (!B.ok) return S2D_OK;
  if (ex) std::copy(B.ex.begin(), B.ex.end(), ex);
  if (ez) std::copy(B.ez.begin(), B.ez.end(), ez);
  if (gx) std::copy(B.gx.begin(), B.gx.end(), gx);
  if (gz) std::copy(B.gz.begin(), B.gz.end(), gz);
  return S2D_OK;
}
int s2d_rcm_box(int32_t nx, int32_t nz, int32_t* perm) {
  if (nx < 1 || nz < 1 || !perm || (long long)nx * nz > 2147483647LL) return S2D_EINVAL;
  const std::vector<int32_t> p = rcm_box_perm(nx, nz);
  for (size_t k = 0; k < p.size(); ++k) perm[k] = p[k] + 1;
  return S2D_OK;
}
int s2d_kernel_route(s2d_handle h, int32_t* route) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(route != nullptr, "s2d_kernel_route: null pointer");
    *route = E.route();
  });
}
int s2d_kernel_ms(s2d_handle h, float* ms) {
  return guard(h, [&](EngineBase& E) {
    S2D_REQUIRE(ms != nullptr, "s2d_kernel_ms: null pointer");
    *ms = E.kernel_ms();
  });
}
int s2d_time_phases(s2d_handle h, int32_t nsteps, float* ms_phase) {
  return guard(h, [&](EngineBase& E) { E.time_phases(nsteps, ms_phase); });
}
int s2d_launch_count(s2d_handle h, int64_t* n) {
  return guard(h, [&](EngineBase& E) {
    if (n) *n = E.launches;
  });
}
int s2d_stream(s2d_handle h, void** stream) {
  return guard(h, [&](EngineBase& E) {
    if (stream) *stream = (void*)E.stream;
  });
}

}  // extern "C"
