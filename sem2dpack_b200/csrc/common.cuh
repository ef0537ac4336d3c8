// Shared helpers of the sem2d_b200 engine (device buffers, error plumbing, launch accounting).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sem2d_b200.h"

namespace s2d {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
struct ArgError : std::runtime_error {
  explicit ArgError(const std::string& m) : std::runtime_error(m) {}
};
struct StateError : std::runtime_error {
  explicit StateError(const std::string& m) : std::runtime_error(m) {}
};
struct SolverError : std::runtime_error {  // device-side abort of the friction solver -> S2D_ESOLVER
  explicit SolverError(const std::string& m) : std::runtime_error(m) {}
};

#define S2D_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      throw s2d::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                           ":" + std::to_string(__LINE__) + ")");                               \
  } while (0)

#define S2D_REQUIRE(cond, msg)                \
  do {                                        \
    if (!(cond)) throw s2d::ArgError(msg);    \
  } while (0)

// Host -> device copy that is complete on the DEVICE when it returns.  cudaMemcpy from pageable host memory may
// return as soon as the data sit in the driver's staging buffer, with the DMA still in flight on the legacy default
// stream; the engine's kernels run on a NON-BLOCKING stream, which that stream does not order -- a kernel launched
// right after an upload could read the old contents of the tail of the buffer (seen: the highest-numbered nodes of
// s2d_set_fields).  Waiting for the default stream closes the window.
inline void h2d_sync(void* dst, const void* src, size_t bytes) {
  cudaError_t e_ = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  if (e_ == cudaSuccess) e_ = cudaStreamSynchronize(0);
  if (e_ != cudaSuccess) throw std::runtime_error(std::string("host to device copy: ") + cudaGetErrorString(e_));
}

// Owning device array.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) S2D_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  // on a given stream: ordered there; without one: complete on the device at return (the legacy default stream
  // does not order the engine's non-blocking stream)
  void zero(cudaStream_t s = 0) {
    if (!n) return;
    S2D_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    if (s == 0) S2D_CUDA(cudaStreamSynchronize(0));
  }
  void upload(const T* host, size_t count) {
    alloc(count);
    if (count) h2d_sync(p, host, count * sizeof(T));
  }
  void upload(const std::vector<T>& v) { upload(v.data(), v.size()); }
  void download(T* host) const {
    if (n) S2D_CUDA(cudaMemcpy(host, p, n * sizeof(T), cudaMemcpyDeviceToHost));
  }
  std::vector<T> to_host() const {
    std::vector<T> v(n);
    download(v.data());
    return v;
  }
};

// Upload a host double array converted to T.
template <typename T>
inline void upload_as(DevBuf<T>& dst, const double* src, size_t count) {
  if (sizeof(T) == sizeof(double)) {
    dst.upload(reinterpret_cast<const T*>(src), count);
  } else {
    std::vector<T> tmp(count);
    for (size_t i = 0; i < count; ++i) tmp[i] = (T)src[i];
    dst.upload(tmp);
  }
}

// device-resident step control block, read by every per-step kernel (keeps the step sequence
// identical from launch to launch, so it can be replayed from a CUDA graph)
struct StepCtl {
  int it;        // current time step (main.f90:51-53)
  int it0;       // step number of the first row of the uploaded stf tables
  int err;       // != 0: device-side abort (NR_Solver)
  int nrows;     // rows in the uploaded stf tables (row = (it - it0) mod nrows)
};

__host__ __device__ inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace s2d
