// placeholder until the structured builder lands
#include "engine.hpp"
namespace s2d {
void cart_free(void*) {}
}
extern "C" {
int s2d_cart_create(s2d_handle*, const s2d_cart_desc*) { return S2D_ESTATE; }
int s2d_cart_add_abso(s2d_handle, int32_t, int32_t) { return S2D_ESTATE; }
int s2d_cart_add_fault_swf(s2d_handle, double, double, double, double, double, double, double, double, int32_t,
                           int32_t, int32_t, int32_t*) { return S2D_ESTATE; }
int s2d_cart_add_force(s2d_handle, double, double, const double*, int32_t*) { return S2D_ESTATE; }
int s2d_cart_add_receivers(s2d_handle, int32_t, double, double, double, double, char, int32_t, int32_t) { return S2D_ESTATE; }
int s2d_cart_info(s2d_handle, int64_t*, int64_t*, double*) { return S2D_ESTATE; }
int s2d_cart_get(s2d_handle, int32_t*, double*, double*, double*) { return S2D_ESTATE; }
int s2d_halo_info(s2d_handle, int64_t*, void**, void**) { return S2D_ESTATE; }
int s2d_halo_set_exchange(s2d_handle, s2d_exchange_fn, void*) { return S2D_ESTATE; }
int s2d_halo_set_peers(s2d_handle, void*, void*, void*, void*) { return S2D_ESTATE; }
}
