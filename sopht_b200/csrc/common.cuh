// Shared helpers for libsopht_b200: view descriptors, argument checks, launch accounting.
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "sopht_b200.h"

namespace sopht {

// ---- error reporting (thread-local text for sopht_last_error) -------------------------------
void set_error(const char* fmt, ...);
extern int64_t g_launch_count;

#define SOPHT_FAIL(code, ...)      \
  do {                             \
    ::sopht::set_error(__VA_ARGS__); \
    return (code);                 \
  } while (0)

#define SOPHT_CHECK_DTYPE(dtype)                                          \
  do {                                                                    \
    if ((dtype) != SOPHT_F32 && (dtype) != SOPHT_F64)                     \
      SOPHT_FAIL(SOPHT_ERR_DTYPE, "%s: invalid dtype %d", __func__, (int)(dtype)); \
  } while (0)

#define SOPHT_CHECK_LAUNCH()                                                          \
  do {                                                                                \
    ::sopht::g_launch_count++;                                                        \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess)                                                           \
      SOPHT_FAIL(SOPHT_ERR_CUDA, "%s: kernel launch failed: %s", __func__,            \
                 cudaGetErrorString(e__));                                            \
  } while (0)

#define SOPHT_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess)                                                           \
      SOPHT_FAIL(SOPHT_ERR_CUDA, "%s: %s failed: %s", __func__, #call,                \
                 cudaGetErrorString(e__));                                            \
  } while (0)

// ---- optional per-kernel event timers (profile.cu) ---------------------------------------------
extern bool g_prof_on;
struct ProfScope {
  ProfScope(const char* label, cudaStream_t st);
  ~ProfScope();
  cudaStream_t st_;
  int idx_;
};
#define SOPHT_PROF(label, st) ::sopht::ProfScope prof_scope__(label, st)

// fused_step3d.cu: w += p * curl_c(f) with the 16-byte-vector register-marching kernel; false = views not
// eligible, nothing launched
bool ns3d_try_forcing_curl_vec(int dtype, const sopht_field_t* vorticity_field,
                               const sopht_field_t* velocity_forcing_field, double prefactor, cudaStream_t st);

// max that keeps a NaN once it has seen one (numpy's amax does; `a > m ? a : m` drops it): a diverged velocity field must
// show up as dt = NaN in compute_stable_timestep like in the reference (passive_transport_flow_simulators.py:150-155).
// A NaN is returned with the sign bit clear so that the integer atomicMax of AtomicMaxNonNeg ranks it above every float.
template <typename T>
__host__ __device__ __forceinline__ T nanmax(T a, T m) {
  if (a != a) return fabs(a);
  return (a > m) ? a : m;  // m NaN: comparison false, m kept
}

// ---- device-side views ----------------------------------------------------------------------
// 3-D scalar view (z, y, x); strides in elements.
template <typename T>
struct View3 {
  T* p;
  int64_t sz, sy, sx;
  __host__ __device__ __forceinline__ T& operator()(int k, int j, int i) const {
    return p[k * sz + j * sy + i * sx];
  }
  __host__ __device__ __forceinline__ T* at(int k, int j, int i) const {
    return p + (k * sz + j * sy + i * sx);
  }
};

// 2-D scalar view (y, x)
template <typename T>
struct View2 {
  T* p;
  int64_t sy, sx;
  __host__ __device__ __forceinline__ T& operator()(int j, int i) const {
    return p[j * sy + i * sx];
  }
};

// ---- host-side helpers on sopht_field_t --------------------------------------------------------
inline bool same_shape(const sopht_field_t* a, const sopht_field_t* b) {
  if (a->ndim != b->ndim) return false;
  for (int d = 0; d < a->ndim; ++d)
    if (a->shape[d] != b->shape[d]) return false;
  return true;
}

inline int64_t numel(const sopht_field_t* a) {
  int64_t n = 1;
  for (int d = 0; d < a->ndim; ++d) n *= a->shape[d];
  return n;
}

inline bool is_contiguous(const sopht_field_t* a) {
  int64_t expect = 1;
  for (int d = a->ndim - 1; d >= 0; --d) {
    if (a->shape[d] != 1 && a->stride[d] != expect) return false;
    expect *= a->shape[d];
  }
  return true;
}

inline bool valid_field(const sopht_field_t* a, int min_dim, int max_dim) {
  if (!a || !a->data) return false;
  if (a->ndim < min_dim || a->ndim > max_dim) return false;
  for (int d = 0; d < a->ndim; ++d)
    if (a->shape[d] < 0 || a->stride[d] < 0) return false;
  return true;
}

// component `c` of a vector field as a 3-D view (drops the leading axis)
template <typename T>
inline View3<T> comp3(const sopht_field_t* f, int c) {
  View3<T> v;
  v.p = reinterpret_cast<T*>(f->data) + c * f->stride[0];
  v.sz = f->stride[1];
  v.sy = f->stride[2];
  v.sx = f->stride[3];
  return v;
}
template <typename T>
inline View3<T> scalar3(const sopht_field_t* f) {
  View3<T> v;
  v.p = reinterpret_cast<T*>(f->data);
  v.sz = f->stride[0];
  v.sy = f->stride[1];
  v.sx = f->stride[2];
  return v;
}
template <typename T>
inline View2<T> comp2(const sopht_field_t* f, int c) {
  View2<T> v;
  v.p = reinterpret_cast<T*>(f->data) + c * f->stride[0];
  v.sy = f->stride[1];
  v.sx = f->stride[2];
  return v;
}
template <typename T>
inline View2<T> scalar2(const sopht_field_t* f) {
  View2<T> v;
  v.p = reinterpret_cast<T*>(f->data);
  v.sy = f->stride[0];
  v.sx = f->stride[1];
  return v;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL), opt-in: SOPHT_PDL=1 ---------------------------------------------------------
// A kernel launched through launch_pdl with the attribute may become resident while its predecessor in the stream is
// still draining: it runs its prologue (shared-memory tables, barrier initialisation - nothing that reads or writes what
// the predecessor produces) and then blocks in pdl_wait() until the predecessor grid has completed and its writes are
// visible; a kernel calls pdl_launch_dependents() when it enters its last tile. Both are no-ops in a launch without the
// attribute. MEASURED SLOWER on the Poisson chain (B200, gpurun_out/r2b_pdl_timings.txt): 512^3 step 14.1 -> 16.4 ms,
// C2 0.505 -> 0.510 ms, C3 0.819 -> 0.846 ms - the early-resident CTAs of the next persistent kernel hold shared memory
// and registers that the single-wave kernel in front of them still needs (the z pass wants a whole SM per CTA). Off by
// default; kept for grids where the chain is launch-bound.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
  static const int v = [] {
    const char* e = getenv("SOPHT_PDL");
    return e ? atoi(e) : 0;
  }();
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// grid for "one thread per cell, x fastest" kernels
struct Grid3 {
  dim3 grid, block;
};
inline Grid3 cell_grid(int nz, int ny, int nx) {
  Grid3 g;
  int bx = nx >= 128 ? 128 : (nx >= 64 ? 64 : 32);
  int by = 256 / bx;
  g.block = dim3(bx, by, 1);
  g.grid = dim3((nx + bx - 1) / bx, (ny + by - 1) / by, nz);
  return g;
}

}  // namespace sopht
