// Building blocks of the register-marching stencil kernels (fused_step3d.cu).
//
// A thread owns W = 16 / sizeof(T) consecutive x cells (one 16-byte vector: float4 / double2) of one (y, x)
// column and marches along z. z neighbours ride in registers, y neighbours are 16-byte read-only loads of the
// rows above / below (served by L1/L2: the same lines are the centre loads of the neighbouring warps of the
// CTA), x neighbours come from the adjacent lanes by warp shuffle; only the two edge lanes of a warp issue a
// scalar load. There is no shared memory and no barrier, so every warp keeps several independent 16-byte
// loads in flight (the loop is unrolled and all loads are ld.global.nc, which lets ptxas hoist the next
// planes' loads above the current plane's stores) - the bytes-in-flight a B200 SM needs to saturate HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sopht {
namespace sv {

template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  using type = float4;
};
template <>
struct VecOf<double> {
  using type = double2;
};

template <typename T>
struct alignas(16) Vec {
  static constexpr int W = 16 / sizeof(T);
  T v[W];
};

template <typename T>
__device__ __forceinline__ Vec<T> vzero() {
  Vec<T> r;
#pragma unroll
  for (int m = 0; m < Vec<T>::W; ++m) r.v[m] = T(0);
  return r;
}

// 16-byte read-only load (p must be 16-byte aligned)
__device__ __forceinline__ Vec<float> vload(const float* p) {
  const float4 q = __ldg(reinterpret_cast<const float4*>(p));
  Vec<float> r;
  r.v[0] = q.x, r.v[1] = q.y, r.v[2] = q.z, r.v[3] = q.w;
  return r;
}
__device__ __forceinline__ Vec<double> vload(const double* p) {
  const double2 q = __ldg(reinterpret_cast<const double2*>(p));
  Vec<double> r;
  r.v[0] = q.x, r.v[1] = q.y;
  return r;
}
// 16-byte load of data this kernel also writes (plain ld.global, not the read-only path)
__device__ __forceinline__ Vec<float> vload_rw(const float* p) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  Vec<float> r;
  r.v[0] = q.x, r.v[1] = q.y, r.v[2] = q.z, r.v[3] = q.w;
  return r;
}
__device__ __forceinline__ Vec<double> vload_rw(const double* p) {
  const double2 q = *reinterpret_cast<const double2*>(p);
  Vec<double> r;
  r.v[0] = q.x, r.v[1] = q.y;
  return r;
}
template <typename T>
__device__ __forceinline__ Vec<T> vload_if(bool pred, const T* p) {
  return pred ? vload(p) : vzero<T>();
}
template <typename T>
__device__ __forceinline__ T sload_if(bool pred, const T* p) {
  return pred ? __ldg(p) : T(0);
}
__device__ __forceinline__ void vstore(float* p, const Vec<float>& a) {
  *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ void vstore(double* p, const Vec<double>& a) {
  *reinterpret_cast<double2*>(p) = make_double2(a.v[0], a.v[1]);
}

// x neighbours of the W cells of `c`: xl[m] = value at x-1, xr[m] = value at x+1. edge_l / edge_r are the
// scalars lane 0 / lane 31 loaded themselves. Must be called by all 32 lanes.
template <typename T>
__device__ __forceinline__ void x_neighbours(const Vec<T>& c, T edge_l, T edge_r, int lane, Vec<T>& xl,
                                             Vec<T>& xr) {
  constexpr int W = Vec<T>::W;
  const T from_l = __shfl_up_sync(0xffffffffu, c.v[W - 1], 1);
  const T from_r = __shfl_down_sync(0xffffffffu, c.v[0], 1);
  xl.v[0] = lane == 0 ? edge_l : from_l;
  xr.v[W - 1] = lane == 31 ? edge_r : from_r;
#pragma unroll
  for (int m = 1; m < W; ++m) xl.v[m] = c.v[m - 1];
#pragma unroll
  for (int m = 0; m < W - 1; ++m) xr.v[m] = c.v[m + 1];
}

}  // namespace sv
}  // namespace sopht
