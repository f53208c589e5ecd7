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
__device__ __forceinline__ void x_neighbours(const Vec<T>& c, T edge_l, T edge_r, bool use_l, bool use_r,
                                             Vec<T>& xl, Vec<T>& xr) {
  constexpr int W = Vec<T>::W;
  const T from_l = __shfl_up_sync(0xffffffffu, c.v[W - 1], 1);
  const T from_r = __shfl_down_sync(0xffffffffu, c.v[0], 1);
  xl.v[0] = use_l ? edge_l : from_l;
  xr.v[W - 1] = use_r ? edge_r : from_r;
#pragma unroll
  for (int m = 1; m < W; ++m) xl.v[m] = c.v[m - 1];
#pragma unroll
  for (int m = 0; m < W - 1; ++m) xr.v[m] = c.v[m + 1];
}

template <typename T>
__device__ __forceinline__ void x_neighbours(const Vec<T>& c, T edge_l, T edge_r, int lane, Vec<T>& xl,
                                             Vec<T>& xr) {
  x_neighbours(c, edge_l, edge_r, lane == 0, lane == 31, xl, xr);
}

// Neighbour addressing of a marching kernel for one thread: offsets (in elements) from a cell of the thread's own
// vector to the rows above / below and to the scalar cells left / right of the warp's span, plus the predicates that
// say whether those loads happen. PXY = false: the reference's ghost-ring rule (neighbours outside the array do not
// exist, the ring is not updated). PXY = true: the x and y directions wrap around (periodic box); z keeps the ring
// rule - a periodic z direction is provided by halo planes the host layer fills (a wrap copy on one GPU, the
// neighbour rank's planes in a slab decomposition).
template <bool PXY>
struct Nbr {
  int64_t up, dn;       // row j + 1, j - 1
  int left, right;      // cell i0 - 1, i0 + W
  bool has_up, has_dn;  // row loads issued
  bool has_l, has_r;    // this lane issues the scalar edge load / uses it instead of the shuffled value
  bool jin;             // row is updated
  __device__ __forceinline__ Nbr(bool act, int lane, int i0, int w, int j, int ny, int nx, int64_t sy) {
    if (PXY) {
      up = j + 1 < ny ? sy : -(int64_t)(ny - 1) * sy;
      dn = j >= 1 ? -sy : (int64_t)(ny - 1) * sy;
      left = i0 > 0 ? -1 : nx - 1;
      right = i0 + w < nx ? w : -i0;
      has_up = has_dn = act;
      has_l = act && lane == 0;
      has_r = act && (lane == 31 || i0 + w >= nx);
      jin = true;
    } else {
      up = sy, dn = -sy, left = -1, right = w;
      has_up = act && j + 1 < ny, has_dn = act && j >= 1;
      has_l = lane == 0 && act && i0 > 0;
      has_r = lane == 31 && act && i0 + w < nx;
      jin = j >= 1 && j < ny - 1;
    }
  }
  // x interior test of cell i (periodic: every cell of the row)
  __device__ __forceinline__ bool iin(int i, int nx) const { return PXY ? true : (i >= 1 && i < nx - 1); }
  __device__ __forceinline__ bool use_l(int lane) const { return PXY ? has_l : lane == 0; }
  __device__ __forceinline__ bool use_r(int lane) const { return PXY ? has_r : lane == 31; }
};

}  // namespace sv
}  // namespace sopht
