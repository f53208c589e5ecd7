// Elementwise kernels on strided views (HBM-bound, no reuse -> no shared memory).
// Two code paths per op: a 16-byte vectorised grid-stride kernel when every operand is a dense,
// 16B-aligned array, and a generic strided kernel (x fastest, coalesced when stride_x == 1) for the
// views the reference passes around (vector components, boundary slabs, buffer corners).
#include <stdarg.h>

#include "common.cuh"

namespace sopht {

thread_local char g_err[512] = "";
int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
template <int NIN>
struct EwDesc {
  int n[4];  // shape, leading dims padded with 1
  int64_t so[4];
  int64_t si[NIN > 0 ? NIN : 1][4];
};

template <typename T, int NIN, typename Op>
__global__ void __launch_bounds__(256)
    ew_strided_kernel(T* __restrict__ out, const T* in0, const T* in1, const T* in2,
                      EwDesc<NIN> d, Op op) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= d.n[3] || j >= d.n[2]) return;
  for (int cz = blockIdx.z; cz < d.n[0] * d.n[1]; cz += gridDim.z) {
    const int c = cz / d.n[1];
    const int k = cz - c * d.n[1];
    T a0 = T(), a1 = T(), a2 = T();
    if (NIN > 0) a0 = in0[c * d.si[0][0] + k * d.si[0][1] + j * d.si[0][2] + i * d.si[0][3]];
    if (NIN > 1) a1 = in1[c * d.si[1 % (NIN > 0 ? NIN : 1)][0] + k * d.si[1 % (NIN > 0 ? NIN : 1)][1] +
                         j * d.si[1 % (NIN > 0 ? NIN : 1)][2] + i * d.si[1 % (NIN > 0 ? NIN : 1)][3]];
    if (NIN > 2) a2 = in2[c * d.si[2 % (NIN > 0 ? NIN : 1)][0] + k * d.si[2 % (NIN > 0 ? NIN : 1)][1] +
                         j * d.si[2 % (NIN > 0 ? NIN : 1)][2] + i * d.si[2 % (NIN > 0 ? NIN : 1)][3]];
    out[c * d.so[0] + k * d.so[1] + j * d.so[2] + i * d.so[3]] = op(a0, a1, a2);
  }
}

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int N = 4;
};
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int N = 2;
};

template <typename T, int NIN, typename Op>
__global__ void __launch_bounds__(256)
    ew_contig_kernel(T* out, const T* in0, const T* in1, const T* in2, int64_t n, Op op) {
  using V = typename Vec16<T>::type;
  constexpr int VN = Vec16<T>::N;
  const int64_t nvec = n / VN;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    V x0, x1, x2, r;
    if (NIN > 0) x0 = reinterpret_cast<const V*>(in0)[v];
    if (NIN > 1) x1 = reinterpret_cast<const V*>(in1)[v];
    if (NIN > 2) x2 = reinterpret_cast<const V*>(in2)[v];
    T* r_ = reinterpret_cast<T*>(&r);
    const T* p0 = reinterpret_cast<const T*>(&x0);
    const T* p1 = reinterpret_cast<const T*>(&x1);
    const T* p2 = reinterpret_cast<const T*>(&x2);
#pragma unroll
    for (int e = 0; e < VN; ++e)
      r_[e] = op(NIN > 0 ? p0[e] : T(), NIN > 1 ? p1[e] : T(), NIN > 2 ? p2[e] : T());
    reinterpret_cast<V*>(out)[v] = r;
  }
  // tail
  const int64_t t = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    out[t] = op(NIN > 0 ? in0[t] : T(), NIN > 1 ? in1[t] : T(), NIN > 2 ? in2[t] : T());
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static void fill_strides(const sopht_field_t* f, int64_t* s4) {
  const int pad = 4 - f->ndim;
  for (int d = 0; d < 4; ++d) s4[d] = d < pad ? 0 : f->stride[d - pad];
}

template <typename T, int NIN, typename Op>
static int launch_ew(const char* name, const sopht_field_t* out, const sopht_field_t* in0,
                     const sopht_field_t* in1, const sopht_field_t* in2, Op op, cudaStream_t st) {
  const sopht_field_t* ins[3] = {in0, in1, in2};
  if (!valid_field(out, 1, 4)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: bad output field", name);
  for (int q = 0; q < NIN; ++q) {
    if (!valid_field(ins[q], 1, 4)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: bad input field %d", name, q);
    if (!same_shape(out, ins[q]))
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: input %d shape differs from output", name, q);
  }
  const int64_t n = numel(out);
  if (n == 0) return SOPHT_OK;
  T* po = reinterpret_cast<T*>(out->data);
  const T* p[3] = {nullptr, nullptr, nullptr};
  bool dense = is_contiguous(out) && aligned16(out->data);
  for (int q = 0; q < NIN; ++q) {
    p[q] = reinterpret_cast<const T*>(ins[q]->data);
    dense = dense && is_contiguous(ins[q]) && aligned16(ins[q]->data);
  }
  if (dense) {
    const int64_t nvec = n / Vec16<T>::N + 1;
    int64_t blocks = (nvec + 255) / 256;
    const int64_t cap = 148 * 16;  // grid-stride: a few resident waves of the 148 SMs
    if (blocks > cap) blocks = cap;
    ew_contig_kernel<T, NIN, Op><<<(unsigned)blocks, 256, 0, st>>>(po, p[0], p[1], p[2], n, op);
    SOPHT_CHECK_LAUNCH();
    return SOPHT_OK;
  }
  EwDesc<NIN> d;
  const int pad = 4 - out->ndim;
  for (int q = 0; q < 4; ++q) {
    const int64_t s = q < pad ? 1 : out->shape[q - pad];
    if (s > 0x7fffffff) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: dimension too large", name);
    d.n[q] = (int)s;
  }
  fill_strides(out, d.so);
  for (int q = 0; q < NIN; ++q) fill_strides(ins[q], d.si[q]);
  Grid3 g = cell_grid(1, d.n[2], d.n[3]);
  int64_t gz = (int64_t)d.n[0] * d.n[1];
  if (gz > 65535) gz = 65535;
  g.grid.z = (unsigned)gz;
  if (g.grid.y > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: y dimension too large", name);
  ew_strided_kernel<T, NIN, Op><<<g.grid, g.block, 0, st>>>(po, p[0], p[1], p[2], d, op);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

// ---- functors --------------------------------------------------------------------------------
template <typename T>
struct OpSet {
  T v;
  __device__ __forceinline__ T operator()(T, T, T) const { return v; }
};
template <typename T>
struct OpCopy {
  __device__ __forceinline__ T operator()(T a, T, T) const { return a; }
};
template <typename T>
struct OpSum {
  __device__ __forceinline__ T operator()(T a, T b, T) const { return a + b; }
};
template <typename T>
struct OpSaxpby {
  T pa, pb;
  __device__ __forceinline__ T operator()(T a, T b, T) const { return pa * a + pb * b; }
};
template <typename T>
struct OpAddVal {
  T v;
  __device__ __forceinline__ T operator()(T a, T, T) const { return a + v; }
};
template <typename T>
struct OpBrinkmann {  // (field, char, penalty)
  T f;
  __device__ __forceinline__ T operator()(T a, T chi, T pen) const {
    return (a + f * chi * pen) / (T(1) + f * chi);
  }
};
template <typename T>
struct OpBrinkmannFixed {  // (field, char)
  T f, pen;
  __device__ __forceinline__ T operator()(T a, T chi, T) const {
    return (a + f * chi * pen) / (T(1) + f * chi);
  }
};
template <typename T>
struct OpCharFunc {
  T blend, sine_prefactor, inv_pi;
  __device__ __forceinline__ T operator()(T phi, T, T) const {
    const T step = phi > blend ? T(1) : T(0);
    const T smooth =
        fabs(phi) > blend ? T(0)
                          : T(0.5) * (T(1) + phi / blend + sin(sine_prefactor * phi) * inv_pi);
    return step + smooth;
  }
};

#define DISPATCH_EW(NIN, OPT, OPINIT, out, a, b, c)                                             \
  do {                                                                                          \
    SOPHT_CHECK_DTYPE(dtype);                                                                   \
    if (dtype == SOPHT_F32) {                                                                   \
      typedef float T;                                                                          \
      OPT<T> op OPINIT;                                                                         \
      return launch_ew<T, NIN, OPT<T>>(__func__, out, a, b, c, op, as_stream(stream));          \
    } else {                                                                                    \
      typedef double T;                                                                         \
      OPT<T> op OPINIT;                                                                         \
      return launch_ew<T, NIN, OPT<T>>(__func__, out, a, b, c, op, as_stream(stream));          \
    }                                                                                           \
  } while (0)

// ---- complex product ----------------------------------------------------------------------------
template <typename T2>
__global__ void __launch_bounds__(256)
    complex_product_kernel(T2* out, const T2* a, const T2* b, EwDesc<2> d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= d.n[3] || j >= d.n[2]) return;
  for (int cz = blockIdx.z; cz < d.n[0] * d.n[1]; cz += gridDim.z) {
    const int c = cz / d.n[1];
    const int k = cz - c * d.n[1];
    const T2 x = a[c * d.si[0][0] + k * d.si[0][1] + j * d.si[0][2] + i * d.si[0][3]];
    const T2 y = b[c * d.si[1][0] + k * d.si[1][1] + j * d.si[1][2] + i * d.si[1][3]];
    T2 r;
    r.x = x.x * y.x - x.y * y.y;
    r.y = x.x * y.y + x.y * y.x;
    out[c * d.so[0] + k * d.so[1] + j * d.so[2] + i * d.so[3]] = r;
  }
}

// ---- cross product ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    cross_product_kernel(View3<T> rx, View3<T> ry, View3<T> rz, View3<const T> ax, View3<const T> ay,
                         View3<const T> az, View3<const T> bx, View3<const T> by, View3<const T> bz,
                         int nz, int ny, int nx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  for (int k = blockIdx.z; k < nz; k += gridDim.z) {
    const T a0 = ax(k, j, i), a1 = ay(k, j, i), a2 = az(k, j, i);
    const T b0 = bx(k, j, i), b1 = by(k, j, i), b2 = bz(k, j, i);
    rx(k, j, i) = a1 * b2 - b1 * a2;
    ry(k, j, i) = a2 * b0 - b2 * a0;
    rz(k, j, i) = a0 * b1 - b0 * a1;
  }
}

template <typename T>
static View3<const T> ccomp3(const sopht_field_t* f, int c) {
  View3<T> v = comp3<T>(f, c);
  View3<const T> r{v.p, v.sz, v.sy, v.sx};
  return r;
}

template <typename T>
static int cross_impl(const sopht_field_t* r, const sopht_field_t* a, const sopht_field_t* b,
                      cudaStream_t st) {
  const int nz = (int)r->shape[1], ny = (int)r->shape[2], nx = (int)r->shape[3];
  if ((int64_t)nz * ny * nx == 0) return SOPHT_OK;
  Grid3 g = cell_grid(nz, ny, nx);
  if (g.grid.z > 65535) g.grid.z = 65535;
  cross_product_kernel<T><<<g.grid, g.block, 0, st>>>(
      comp3<T>(r, 0), comp3<T>(r, 1), comp3<T>(r, 2), ccomp3<T>(a, 0), ccomp3<T>(a, 1),
      ccomp3<T>(a, 2), ccomp3<T>(b, 0), ccomp3<T>(b, 1), ccomp3<T>(b, 2), nz, ny, nx);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

// ---- boundary ring setter -----------------------------------------------------------------------
// One launch per face pair: the face axis is squeezed to 2*width indices (front then back slab).
template <typename T>
__global__ void __launch_bounds__(256)
    ring_face_kernel(T* p, int64_t s_face, int64_t s_a, int64_t s_b, int n_face, int n_a, int n_b,
                     int width, T val) {
  // thread x -> axis b (fastest iterated), thread y -> face index (0..2w-1), block z -> axis a
  const int ib = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y * blockDim.y + threadIdx.y;
  if (ib >= n_b || f >= 2 * width) return;
  const int face_idx = f < width ? f : n_face - 2 * width + f;
  if (face_idx < 0 || face_idx >= n_face) return;
  for (int ia = blockIdx.z; ia < n_a; ia += gridDim.z)
    p[face_idx * s_face + ia * s_a + ib * s_b] = val;
}

template <typename T>
static int ring_set_scalar(T* p, const int64_t* shape, const int64_t* stride, int ndim, int width,
                           T val, cudaStream_t st) {
  // ndim is 2 or 3; for each axis launch over (face, a, b)
  for (int ax = 0; ax < ndim; ++ax) {
    int oa = -1, ob = -1;  // the other axes; ob = the one with the smallest stride (fastest)
    for (int d = 0; d < ndim; ++d) {
      if (d == ax) continue;
      if (oa < 0)
        oa = d;
      else
        ob = d;
    }
    int n_a, n_b;
    int64_t s_a, s_b;
    if (ob < 0) {  // 2-D: only one other axis
      n_a = 1;
      s_a = 0;
      n_b = (int)shape[oa];
      s_b = stride[oa];
    } else {
      n_a = (int)shape[oa];
      s_a = stride[oa];
      n_b = (int)shape[ob];
      s_b = stride[ob];
    }
    const int n_face = (int)shape[ax];
    if (n_face == 0 || n_a == 0 || n_b == 0) continue;
    const int w = width < n_face ? width : n_face;  // slabs clip like numpy slices
    (void)w;
    dim3 block(64, 4, 1);
    dim3 grid((n_b + 63) / 64, (2 * width + 3) / 4, n_a > 65535 ? 65535 : n_a);
    ring_face_kernel<T><<<grid, block, 0, st>>>(p, stride[ax], s_a, s_b, n_face, n_a, n_b, width,
                                                 val);
    g_launch_count++;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    SOPHT_FAIL(SOPHT_ERR_CUDA, "set_fixed_val_at_boundaries: %s", cudaGetErrorString(e));
  return SOPHT_OK;
}

// ---- sum |u_c| + max ----------------------------------------------------------------------------
template <typename T>
struct MaxBits;
template <>
struct MaxBits<float> {
  __device__ static void atomic_max(float* addr, float v) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));  // valid for v >= 0
  }
};
template <>
struct MaxBits<double> {
  __device__ static void atomic_max(double* addr, double v) {
    atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
  }
};

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
    abs_sum_max_kernel(T* mag, int64_t ms0, int64_t ms1, int64_t ms2, const T* vel, int64_t vc,
                       int64_t vs0, int64_t vs1, int64_t vs2, int n0, int n1, int n2, T* max_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  T m = T(0);
  if (i < n2 && j < n1) {
    for (int k = blockIdx.z; k < n0; k += gridDim.z) {
      const int64_t o = k * vs0 + j * vs1 + i * vs2;
      T s = T(0);
#pragma unroll
      for (int c = 0; c < DIM; ++c) s += fabs(vel[c * vc + o]);
      mag[k * ms0 + j * ms1 + i * ms2] = s;
      m = nanmax(s, m);
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const T o = __shfl_xor_sync(0xffffffffu, m, off);
    m = nanmax(o, m);
  }
  __shared__ T wm[8];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0) wm[tid >> 5] = m;
  __syncthreads();
  if (tid < 32) {
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    m = tid < nw ? wm[tid] : T(0);
    for (int off = 4; off > 0; off >>= 1) {
      const T o = __shfl_xor_sync(0xffffffffu, m, off);
      m = nanmax(o, m);
    }
    if (tid == 0) MaxBits<T>::atomic_max(max_out, m);
  }
}

template <typename T>
static int abs_sum_max_impl(const sopht_field_t* mag, const sopht_field_t* vel, void* max_out,
                            cudaStream_t st) {
  const int dim = (int)vel->shape[0];
  int n0, n1, n2;
  int64_t ms[3], vs[3];
  if (dim == 3) {
    n0 = (int)vel->shape[1];
    n1 = (int)vel->shape[2];
    n2 = (int)vel->shape[3];
    for (int d = 0; d < 3; ++d) {
      ms[d] = mag->stride[d];
      vs[d] = vel->stride[d + 1];
    }
  } else {
    n0 = 1;
    n1 = (int)vel->shape[1];
    n2 = (int)vel->shape[2];
    ms[0] = 0;
    vs[0] = 0;
    for (int d = 0; d < 2; ++d) {
      ms[d + 1] = mag->stride[d];
      vs[d + 1] = vel->stride[d + 1];
    }
  }
  SOPHT_CUDA(cudaMemsetAsync(max_out, 0, sizeof(T), st));
  if ((int64_t)n0 * n1 * n2 == 0) return SOPHT_OK;
  Grid3 g = cell_grid(n0, n1, n2);
  // a few planes per block keep the atomic count low
  unsigned gz = (n0 + 7) / 8;
  g.grid.z = gz < 1 ? 1 : gz;
  T* m = reinterpret_cast<T*>(mag->data);
  const T* v = reinterpret_cast<const T*>(vel->data);
  if (dim == 3)
    abs_sum_max_kernel<T, 3><<<g.grid, g.block, 0, st>>>(m, ms[0], ms[1], ms[2], v, vel->stride[0],
                                                          vs[0], vs[1], vs[2], n0, n1, n2,
                                                          reinterpret_cast<T*>(max_out));
  else
    abs_sum_max_kernel<T, 2><<<g.grid, g.block, 0, st>>>(m, ms[0], ms[1], ms[2], v, vel->stride[0],
                                                          vs[0], vs[1], vs[2], n0, n1, n2,
                                                          reinterpret_cast<T*>(max_out));
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

}  // namespace sopht

using namespace sopht;

extern "C" {

const char* sopht_last_error(void) { return g_err; }
int sopht_version(void) { return 100; }
int64_t sopht_launch_count(void) { return g_launch_count; }

int sopht_set_fixed_val(int dtype, const sopht_field_t* field, double fixed_val, void* stream) {
  DISPATCH_EW(0, OpSet, {(T)fixed_val}, field, nullptr, nullptr, nullptr);
}

static int sub_component(const sopht_field_t* f, int c, sopht_field_t* out, size_t elem) {
  out->ndim = f->ndim - 1;
  out->data = reinterpret_cast<char*>(f->data) + (size_t)c * f->stride[0] * elem;
  for (int d = 0; d < out->ndim; ++d) {
    out->shape[d] = f->shape[d + 1];
    out->stride[d] = f->stride[d + 1];
  }
  return 0;
}

int sopht_set_fixed_vals_vector(int dtype, const sopht_field_t* vector_field,
                                const double* fixed_vals, int n_vals, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(vector_field, 2, 5) || !fixed_vals || n_vals < vector_field->shape[0])
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: need one value per leading-axis component", __func__);
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  for (int c = 0; c < vector_field->shape[0]; ++c) {
    sopht_field_t sub;
    sub_component(vector_field, c, &sub, elem);
    int rc = sopht_set_fixed_val(dtype, &sub, fixed_vals[c], stream);
    if (rc) return rc;
  }
  return SOPHT_OK;
}

int sopht_elementwise_copy(int dtype, const sopht_field_t* field, const sopht_field_t* rhs_field,
                           void* stream) {
  DISPATCH_EW(1, OpCopy, {}, field, rhs_field, nullptr, nullptr);
}

int sopht_elementwise_sum(int dtype, const sopht_field_t* sum_field, const sopht_field_t* field_1,
                          const sopht_field_t* field_2, void* stream) {
  DISPATCH_EW(2, OpSum, {}, sum_field, field_1, field_2, nullptr);
}

int sopht_elementwise_saxpby(int dtype, const sopht_field_t* sum_field,
                             const sopht_field_t* field_1, const sopht_field_t* field_2,
                             double field_1_prefac, double field_2_prefac, void* stream) {
  DISPATCH_EW(2, OpSaxpby, ({(T)field_1_prefac, (T)field_2_prefac}), sum_field, field_1, field_2,
              nullptr);
}

int sopht_add_fixed_val(int dtype, const sopht_field_t* sum_field, const sopht_field_t* field,
                        double fixed_val, void* stream) {
  DISPATCH_EW(1, OpAddVal, {(T)fixed_val}, sum_field, field, nullptr, nullptr);
}

int sopht_add_fixed_vals_vector(int dtype, const sopht_field_t* sum_field,
                                const sopht_field_t* vector_field, const double* fixed_vals,
                                int n_vals, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(vector_field, 2, 5) || !valid_field(sum_field, 2, 5) ||
      !same_shape(sum_field, vector_field) || !fixed_vals || n_vals < vector_field->shape[0])
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: shape mismatch or missing component values", __func__);
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  for (int c = 0; c < vector_field->shape[0]; ++c) {
    sopht_field_t so, si;
    sub_component(sum_field, c, &so, elem);
    sub_component(vector_field, c, &si, elem);
    int rc = sopht_add_fixed_val(dtype, &so, &si, fixed_vals[c], stream);
    if (rc) return rc;
  }
  return SOPHT_OK;
}

int sopht_elementwise_complex_product(int dtype, const sopht_field_t* product_field,
                                      const sopht_field_t* field_1, const sopht_field_t* field_2,
                                      void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(product_field, 1, 4) || !valid_field(field_1, 1, 4) ||
      !valid_field(field_2, 1, 4) || !same_shape(product_field, field_1) ||
      !same_shape(product_field, field_2))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: operand shapes differ", __func__);
  if (numel(product_field) == 0) return SOPHT_OK;
  EwDesc<2> d;
  const int pad = 4 - product_field->ndim;
  for (int q = 0; q < 4; ++q) d.n[q] = q < pad ? 1 : (int)product_field->shape[q - pad];
  fill_strides(product_field, d.so);
  fill_strides(field_1, d.si[0]);
  fill_strides(field_2, d.si[1]);
  Grid3 g = cell_grid(1, d.n[2], d.n[3]);
  int64_t gz = (int64_t)d.n[0] * d.n[1];
  g.grid.z = (unsigned)(gz > 65535 ? 65535 : gz);
  if (dtype == SOPHT_F32)
    complex_product_kernel<float2><<<g.grid, g.block, 0, as_stream(stream)>>>(
        (float2*)product_field->data, (const float2*)field_1->data, (const float2*)field_2->data, d);
  else
    complex_product_kernel<double2><<<g.grid, g.block, 0, as_stream(stream)>>>(
        (double2*)product_field->data, (const double2*)field_1->data,
        (const double2*)field_2->data, d);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_elementwise_cross_product_3d(int dtype, const sopht_field_t* result_field,
                                       const sopht_field_t* field_1, const sopht_field_t* field_2,
                                       void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(result_field, 4, 4) || !valid_field(field_1, 4, 4) ||
      !valid_field(field_2, 4, 4) || result_field->shape[0] != 3 ||
      !same_shape(result_field, field_1) || !same_shape(result_field, field_2))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: operands must all be (3, nz, ny, nx)", __func__);
  return dtype == SOPHT_F32
             ? cross_impl<float>(result_field, field_1, field_2, as_stream(stream))
             : cross_impl<double>(result_field, field_1, field_2, as_stream(stream));
}

int sopht_set_fixed_val_at_boundaries(int dtype, const sopht_field_t* field, int width,
                                      const double* fixed_vals, int is_vector, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (width <= 0) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: width must be a positive integer", __func__);
  if (!fixed_vals) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: fixed_vals is null", __func__);
  const int gdim = field ? field->ndim - (is_vector ? 1 : 0) : 0;
  if (!valid_field(field, 2, 4) || (gdim != 2 && gdim != 3))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a 2-D/3-D grid field", __func__);
  const int ncomp = is_vector ? (int)field->shape[0] : 1;
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  for (int c = 0; c < ncomp; ++c) {
    char* base = reinterpret_cast<char*>(field->data) +
                 (is_vector ? (size_t)c * field->stride[0] * elem : 0);
    const int64_t* shp = field->shape + (is_vector ? 1 : 0);
    const int64_t* str = field->stride + (is_vector ? 1 : 0);
    int rc = dtype == SOPHT_F32
                 ? ring_set_scalar<float>((float*)base, shp, str, gdim, width, (float)fixed_vals[c],
                                          as_stream(stream))
                 : ring_set_scalar<double>((double*)base, shp, str, gdim, width, fixed_vals[c],
                                           as_stream(stream));
    if (rc) return rc;
  }
  return SOPHT_OK;
}

int sopht_brinkmann_penalise(int dtype, const sopht_field_t* penalised_field,
                             const sopht_field_t* field, const sopht_field_t* char_field,
                             const sopht_field_t* penalty_field, double penalty_factor,
                             void* stream) {
  DISPATCH_EW(3, OpBrinkmann, {(T)penalty_factor}, penalised_field, field, char_field,
              penalty_field);
}

int sopht_brinkmann_penalise_vs_fixed_val(int dtype, const sopht_field_t* penalised_field,
                                          const sopht_field_t* field,
                                          const sopht_field_t* char_field, double penalty_val,
                                          double penalty_factor, void* stream) {
  DISPATCH_EW(2, OpBrinkmannFixed, ({(T)penalty_factor, (T)penalty_val}), penalised_field, field,
              char_field, nullptr);
}

int sopht_char_func_from_level_set(int dtype, const sopht_field_t* char_func_field,
                                   const sopht_field_t* level_set_field, double blend_width,
                                   void* stream) {
  const double pi = 3.14159265358979323846;
  DISPATCH_EW(1, OpCharFunc, ({(T)blend_width, (T)(pi / blend_width), (T)(1.0 / pi)}),
              char_func_field, level_set_field, nullptr, nullptr);
}

int sopht_abs_sum_max(int dtype, const sopht_field_t* velocity_magnitude_field,
                      const sopht_field_t* velocity_field, void* max_out, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(velocity_field, 3, 4) || !valid_field(velocity_magnitude_field, 2, 3) ||
      velocity_field->ndim != velocity_magnitude_field->ndim + 1 ||
      velocity_field->shape[0] != velocity_field->ndim - 1 || !max_out)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected (dim, grid...) velocity and (grid...) magnitude",
               __func__);
  for (int d = 0; d < velocity_magnitude_field->ndim; ++d)
    if (velocity_magnitude_field->shape[d] != velocity_field->shape[d + 1])
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid shapes differ", __func__);
  return dtype == SOPHT_F32 ? abs_sum_max_impl<float>(velocity_magnitude_field, velocity_field,
                                                      max_out, as_stream(stream))
                            : abs_sum_max_impl<double>(velocity_magnitude_field, velocity_field,
                                                       max_out, as_stream(stream));
}

}  // extern "C"
