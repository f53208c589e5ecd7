// Unbounded (free-space) Poisson solve by Hockney-Eastwood domain doubling — GENERIC path.
//
// Works for any grid size and both precisions; cuFFT does only batched 2-D and strided 1-D
// transforms, everything around them is hand-written:
//   pad    : rhs (nz,ny,nx) -> zero-padded planes (nz,2ny,2nx)   [z padding is never materialised in real space]
//   fft    : batched 2-D R2C over the nz non-zero planes, then one strided 1-D C2C along z (2nz)
//   green  : spectrum *= Re(G_hat) * dx^dim / (doubled cell count)  (G is real and even => G_hat is real)
//   ifft   : strided 1-D C2C inverse along z, batched 2-D C2R over the first nz planes only
//   crop   : low corner -> solution
// The power-of-two fp32 fast path (poisson_pow2.cu) replaces all of this with fused shared-memory FFT
// kernels; this file is the reference-faithful any-size path and also builds G_hat for both.
//
// ref: sopht/numeric/eulerian_grid_ops/poisson_solver_3d/UnboundedPoissonSolverPYFFTW3D.py:9-172,
//      poisson_solver_2d/UnboundedPoissonSolverPYFFTW2D.py:8-129
#include <cufft.h>

#include <vector>

#include "common.cuh"
#include "poisson.cuh"

namespace sopht {

#define SOPHT_CUFFT(call)                                                                     \
  do {                                                                                        \
    cufftResult r__ = (call);                                                                 \
    if (r__ != CUFFT_SUCCESS)                                                                 \
      SOPHT_FAIL(SOPHT_ERR_CUFFT, "%s: %s failed with cufftResult %d", __func__, #call, (int)r__); \
  } while (0)

template <typename T>
struct FFTTypes;
template <>
struct FFTTypes<float> {
  using C = cufftComplex;
  static constexpr cufftType R2C = CUFFT_R2C, C2R = CUFFT_C2R, C2C = CUFFT_C2C;
  static cufftResult r2c(cufftHandle p, float* in, C* out) { return cufftExecR2C(p, in, out); }
  static cufftResult c2r(cufftHandle p, C* in, float* out) { return cufftExecC2R(p, in, out); }
  static cufftResult c2c(cufftHandle p, C* in, C* out, int dir) { return cufftExecC2C(p, in, out, dir); }
};
template <>
struct FFTTypes<double> {
  using C = cufftDoubleComplex;
  static constexpr cufftType R2C = CUFFT_D2Z, C2R = CUFFT_Z2D, C2C = CUFFT_Z2Z;
  static cufftResult r2c(cufftHandle p, double* in, C* out) { return cufftExecD2Z(p, in, out); }
  static cufftResult c2r(cufftHandle p, C* in, double* out) { return cufftExecZ2D(p, in, out); }
  static cufftResult c2c(cufftHandle p, C* in, C* out, int dir) { return cufftExecZ2Z(p, in, out, dir); }
};

// ---- hand-written kernels around the transforms --------------------------------------------------------
// planes (np, 2ny, 2nx) <- rhs (np, ny, nx) zero padded; one thread per padded cell, x fastest
template <typename T>
__global__ void __launch_bounds__(256)
    pad_planes_kernel(T* __restrict__ dst, View3<const T> src, int np, int ny, int nx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= 2 * nx || j >= 2 * ny) return;
  for (int k = blockIdx.z; k < np; k += gridDim.z) {
    T v = T(0);
    if (i < nx && j < ny) v = src(k, j, i);
    dst[((int64_t)k * 2 * ny + j) * 2 * nx + i] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    crop_planes_kernel(View3<T> dst, const T* __restrict__ src, int np, int ny, int nx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  for (int k = blockIdx.z; k < np; k += gridDim.z)
    dst(k, j, i) = src[((int64_t)k * 2 * ny + j) * 2 * nx + i];
}

// spectrum[n] *= g[n] (g real); 1-D grid-stride
template <typename T, typename C>
__global__ void __launch_bounds__(256) green_multiply_kernel(C* spec, const T* __restrict__ g, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    C v = spec[q];
    const T s = g[q];
    v.x *= s;
    v.y *= s;
    spec[q] = v;
  }
}

// g_real[n] = Re(spec_double[n]) * scale  (setup only)
template <typename T>
__global__ void __launch_bounds__(256)
    take_real_scaled_kernel(T* g, const cufftDoubleComplex* spec, double scale, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    g[q] = (T)(spec[q].x * scale);
}

// Free-space Green's function on the doubled grid, evaluated in T with the reference's operation order
// (each numpy ufunc rounds once, so no FMA contraction here), output widened to double for the setup FFT.
// 3-D: (1/sqrt(mx^2+my^2+mz^2))/(4 pi); 2-D: -log(sqrt(mx^2+my^2))/(2 pi); origin regularised by the caller's value.
template <typename T>
struct RnOps;
template <>
struct RnOps<float> {
  __device__ static float mul(float a, float b) { return __fmul_rn(a, b); }
  __device__ static float add(float a, float b) { return __fadd_rn(a, b); }
  __device__ static float sqrt_(float a) { return __fsqrt_rn(a); }
  __device__ static float div(float a, float b) { return __fdiv_rn(a, b); }
  __device__ static float log_(float a) { return logf(a); }
};
template <>
struct RnOps<double> {
  __device__ static double mul(double a, double b) { return __dmul_rn(a, b); }
  __device__ static double add(double a, double b) { return __dadd_rn(a, b); }
  __device__ static double sqrt_(double a) { return __dsqrt_rn(a); }
  __device__ static double div(double a, double b) { return __ddiv_rn(a, b); }
  __device__ static double log_(double a) { return log(a); }
};

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
    greens_function_kernel(double* g, const T* mz, const T* my, const T* mx, int n2z, int n2y, int n2x,
                           T denom, T origin_value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n2x || j >= n2y) return;
  using R = RnOps<T>;
  for (int k = blockIdx.z; k < n2z; k += gridDim.z) {
    T r2 = R::add(R::mul(mx[i], mx[i]), R::mul(my[j], my[j]));
    if (DIM == 3) r2 = R::add(r2, R::mul(mz[k], mz[k]));
    const T r = R::sqrt_(r2);
    T v;
    if (DIM == 3)
      v = R::div(R::div(T(1), r), denom);
    else
      v = R::div(-R::log_(r), denom);
    if (i == 0 && j == 0 && k == 0) v = origin_value;
    g[((int64_t)k * n2y + j) * n2x + i] = (double)v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// G_hat (real part, scaled) on the doubled half-spectrum (n2z, n2y, nx+1), computed in double precision.
template <typename T>
int build_green_hat(T** g_out, int dim, int nz, int ny, int nx, double dx, const double* mz_h,
                           const double* my_h, const double* mx_h, double origin_value,
                           cudaStream_t st) {
  const int n2z = dim == 3 ? 2 * nz : 1, n2y = 2 * ny, n2x = 2 * nx, nkx = nx + 1;
  const int64_t nreal = (int64_t)n2z * n2y * n2x, nspec = (int64_t)n2z * n2y * nkx;
  std::vector<T> hz(n2z), hy(n2y), hx(n2x);
  for (int q = 0; q < n2z; ++q) hz[q] = dim == 3 ? (T)mz_h[q] : T(0);
  for (int q = 0; q < n2y; ++q) hy[q] = (T)my_h[q];
  for (int q = 0; q < n2x; ++q) hx[q] = (T)mx_h[q];
  T *dz = nullptr, *dy = nullptr, *dxp = nullptr, *ghat = nullptr;
  double* greal = nullptr;
  cufftDoubleComplex* gspec = nullptr;
  cufftHandle p2d = 0, p1d = 0;
  int rc = SOPHT_OK;
  auto cleanup = [&]() {
    cudaFree(dz);
    cudaFree(dy);
    cudaFree(dxp);
    cudaFree(greal);
    cudaFree(gspec);
    if (p2d) cufftDestroy(p2d);
    if (p1d) cufftDestroy(p1d);
  };
#define SETUP_TRY(expr)       \
  do {                        \
    rc = [&]() -> int {       \
      expr;                   \
      return SOPHT_OK;        \
    }();                      \
    if (rc) {                 \
      cleanup();              \
      cudaFree(ghat);         \
      return rc;              \
    }                         \
  } while (0)
  SETUP_TRY(SOPHT_CUDA(cudaMalloc(&dz, sizeof(T) * n2z)));
  SETUP_TRY(SOPHT_CUDA(cudaMalloc(&dy, sizeof(T) * n2y)));
  SETUP_TRY(SOPHT_CUDA(cudaMalloc(&dxp, sizeof(T) * n2x)));
  SETUP_TRY(SOPHT_CUDA(cudaMalloc(&greal, sizeof(double) * nreal)));
  SETUP_TRY(SOPHT_CUDA(cudaMalloc(&gspec, sizeof(cufftDoubleComplex) * nspec)));
  SETUP_TRY(SOPHT_CUDA(cudaMalloc(&ghat, sizeof(T) * nspec)));
  SETUP_TRY(SOPHT_CUDA(cudaMemcpyAsync(dz, hz.data(), sizeof(T) * n2z, cudaMemcpyHostToDevice, st)));
  SETUP_TRY(SOPHT_CUDA(cudaMemcpyAsync(dy, hy.data(), sizeof(T) * n2y, cudaMemcpyHostToDevice, st)));
  SETUP_TRY(SOPHT_CUDA(cudaMemcpyAsync(dxp, hx.data(), sizeof(T) * n2x, cudaMemcpyHostToDevice, st)));
  {
    Grid3 g = cell_grid(n2z, n2y, n2x);
    if (g.grid.z > 1024) g.grid.z = 1024;
    const double pi = 3.14159265358979323846;
    if (dim == 3)
      greens_function_kernel<T, 3><<<g.grid, g.block, 0, st>>>(greal, dz, dy, dxp, n2z, n2y, n2x,
                                                               (T)(4 * pi), (T)origin_value);
    else
      greens_function_kernel<T, 2><<<g.grid, g.block, 0, st>>>(greal, dz, dy, dxp, n2z, n2y, n2x,
                                                               (T)(2 * pi), (T)origin_value);
    SETUP_TRY(SOPHT_CHECK_LAUNCH());
  }
  {
    int n2[2] = {n2y, n2x};
    int inembed[2] = {n2y, n2x}, onembed[2] = {n2y, nkx};
    SETUP_TRY(SOPHT_CUFFT(cufftPlanMany(&p2d, 2, n2, inembed, 1, n2y * n2x, onembed, 1, n2y * nkx,
                                        CUFFT_D2Z, n2z)));
    SETUP_TRY(SOPHT_CUFFT(cufftSetStream(p2d, st)));
    SETUP_TRY(SOPHT_CUFFT(cufftExecD2Z(p2d, greal, gspec)));
    if (dim == 3) {
      int n1[1] = {n2z};
      const int S = n2y * nkx;
      SETUP_TRY(SOPHT_CUFFT(cufftPlanMany(&p1d, 1, n1, n1, S, 1, n1, S, 1, CUFFT_Z2Z, S)));
      SETUP_TRY(SOPHT_CUFFT(cufftSetStream(p1d, st)));
      SETUP_TRY(SOPHT_CUFFT(cufftExecZ2Z(p1d, gspec, gspec, CUFFT_FORWARD)));
    }
  }
  {
    // G_hat * dx^dim (in T, like the reference) / doubled cell count (cuFFT's inverse is unnormalised)
    const T dxT = (T)dx;
    const T dxp_ = dim == 3 ? dxT * dxT * dxT : dxT * dxT;
    const double scale = (double)dxp_ / (double)nreal;
    take_real_scaled_kernel<T><<<148 * 8, 256, 0, st>>>(ghat, gspec, scale, nspec);
    SETUP_TRY(SOPHT_CHECK_LAUNCH());
  }
  SETUP_TRY(SOPHT_CUDA(cudaStreamSynchronize(st)));
#undef SETUP_TRY
  cleanup();
  *g_out = ghat;
  return SOPHT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Folded G_hat of ONE rank's kx range, built without ever holding the doubled domain (z-slab decomposed solve:
// at 1024^3 the full double-precision transform above needs 137 GB, this one ~9 GB per rank):
//   x: planes z = 0..nz of the (even) Green's function in batches, batched 1-D D2Z along x, keep the bins
//      [kx0, kx0 + nxl) and kx = nx          -> S[z][y][k], k <= nxl
//   y: Z2Z along y per plane;  mirror planes z -> 2nz - z (G is even in z);  z: one batched Z2Z along z
//   fold: gm[fz][fy][k] = 2 scale Re S[fz][fy][k], gn[fz][fy] = 2 scale Re S[fz][fy][nxl]   (fz <= nz, fy <= ny)
// Same arithmetic as build_green_hat (Green's function in float with the reference's operation order, transform in
// double), so the two agree to the rounding of the FFT factorisation.
__global__ void __launch_bounds__(256)
    greens_rows_kernel(double* g, const float* mz, const float* my, const float* mx, int z0, int nzb, int n2y,
                       int n2x, float denom, float origin_value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n2x || j >= n2y) return;
  using R = RnOps<float>;
  for (int k = blockIdx.z; k < nzb; k += gridDim.z) {
    const float zc = mz[z0 + k];
    const float r2 = R::add(R::add(R::mul(mx[i], mx[i]), R::mul(my[j], my[j])), R::mul(zc, zc));
    float v = R::div(R::div(1.0f, R::sqrt_(r2)), denom);
    if (i == 0 && j == 0 && z0 + k == 0) v = origin_value;
    g[((int64_t)k * n2y + j) * n2x + i] = (double)v;
  }
}
// S[z0 + k][y][q] = row_spectrum[k][y][kx0 + q] (q < nxl), S[..][nxl] = row_spectrum[..][nx]
__global__ void __launch_bounds__(256)
    keep_bins_kernel(cufftDoubleComplex* S, const cufftDoubleComplex* rows, int64_t nrows, int nkx, int kx0,
                     int nxl, int nx) {
  const int64_t total = nrows * (nxl + 1);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(q % (nxl + 1));
    const int64_t r = q / (nxl + 1);
    S[q] = rows[r * nkx + (k < nxl ? kx0 + k : nx)];
  }
}
__global__ void __launch_bounds__(256)
    fold_slice_kernel(float* gm, float* gn, const cufftDoubleComplex* S, int nz, int ny, int nxl, double scale2) {
  const int64_t total = (int64_t)(nz + 1) * (ny + 1) * (nxl + 1);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(q % (nxl + 1));
    const int64_t r = q / (nxl + 1);
    const int fy = (int)(r % (ny + 1)), fz = (int)(r / (ny + 1));
    // same rounding as the single-GPU path: scale in double, narrow to float, then the fold's factor 2
    const float v = 2.0f * (float)(S[((int64_t)fz * 2 * ny + fy) * (nxl + 1) + k].x * scale2);
    if (k == nxl)
      gn[(int64_t)fz * (ny + 1) + fy] = v;
    else
      gm[((int64_t)fz * (ny + 1) + fy) * nxl + k] = v;
  }
}

int build_green_folded_slice(float* gm, float* gn, int nz, int ny, int nx, int kx0, int nxl, double dx,
                             const double* mz_h, const double* my_h, const double* mx_h, double origin_value,
                             cudaStream_t st) {
  const int n2z = 2 * nz, n2y = 2 * ny, n2x = 2 * nx, nkx = nx + 1, nk = nxl + 1;
  std::vector<float> hz(n2z), hy(n2y), hx(n2x);
  for (int q = 0; q < n2z; ++q) hz[q] = (float)mz_h[q];
  for (int q = 0; q < n2y; ++q) hy[q] = (float)my_h[q];
  for (int q = 0; q < n2x; ++q) hx[q] = (float)mx_h[q];
  // z planes per batch: about 256 MB of real rows
  int zb = (int)(((int64_t)32 << 20) / ((int64_t)n2y * n2x));
  if (zb < 1) zb = 1;
  if (zb > nz + 1) zb = nz + 1;
  float *dz = nullptr, *dy = nullptr, *dxp = nullptr;
  double* rows = nullptr;
  cufftDoubleComplex *rspec = nullptr, *S = nullptr;
  cufftHandle px = 0, py = 0, pz = 0;
  int rc = SOPHT_OK;
  auto cleanup = [&]() {
    cudaFree(dz);
    cudaFree(dy);
    cudaFree(dxp);
    cudaFree(rows);
    cudaFree(rspec);
    cudaFree(S);
    if (px) cufftDestroy(px);
    if (py) cufftDestroy(py);
    if (pz) cufftDestroy(pz);
  };
#define SLICE_TRY(expr)       \
  do {                        \
    rc = [&]() -> int {       \
      expr;                   \
      return SOPHT_OK;        \
    }();                      \
    if (rc) {                 \
      cleanup();              \
      return rc;              \
    }                         \
  } while (0)
  const int64_t plane = (int64_t)n2y * nk;  // complex elements of one z plane of S
  SLICE_TRY(SOPHT_CUDA(cudaMalloc(&dz, sizeof(float) * n2z)));
  SLICE_TRY(SOPHT_CUDA(cudaMalloc(&dy, sizeof(float) * n2y)));
  SLICE_TRY(SOPHT_CUDA(cudaMalloc(&dxp, sizeof(float) * n2x)));
  SLICE_TRY(SOPHT_CUDA(cudaMalloc(&rows, sizeof(double) * (size_t)zb * n2y * n2x)));
  SLICE_TRY(SOPHT_CUDA(cudaMalloc(&rspec, sizeof(cufftDoubleComplex) * (size_t)zb * n2y * nkx)));
  SLICE_TRY(SOPHT_CUDA(cudaMalloc(&S, sizeof(cufftDoubleComplex) * (size_t)n2z * plane)));
  SLICE_TRY(SOPHT_CUDA(cudaMemcpyAsync(dz, hz.data(), sizeof(float) * n2z, cudaMemcpyHostToDevice, st)));
  SLICE_TRY(SOPHT_CUDA(cudaMemcpyAsync(dy, hy.data(), sizeof(float) * n2y, cudaMemcpyHostToDevice, st)));
  SLICE_TRY(SOPHT_CUDA(cudaMemcpyAsync(dxp, hx.data(), sizeof(float) * n2x, cudaMemcpyHostToDevice, st)));
  {
    int n1[1] = {n2x};
    SLICE_TRY(SOPHT_CUFFT(cufftPlanMany(&px, 1, n1, nullptr, 1, n2x, nullptr, 1, nkx, CUFFT_D2Z, zb * n2y)));
    SLICE_TRY(SOPHT_CUFFT(cufftSetStream(px, st)));
    int ny1[1] = {n2y};
    SLICE_TRY(SOPHT_CUFFT(cufftPlanMany(&py, 1, ny1, ny1, nk, 1, ny1, nk, 1, CUFFT_Z2Z, nk)));
    SLICE_TRY(SOPHT_CUFFT(cufftSetStream(py, st)));
    int nz1[1] = {n2z};
    SLICE_TRY(SOPHT_CUFFT(cufftPlanMany(&pz, 1, nz1, nz1, (int)plane, 1, nz1, (int)plane, 1, CUFFT_Z2Z, (int)plane)));
    SLICE_TRY(SOPHT_CUFFT(cufftSetStream(pz, st)));
  }
  const float four_pi = (float)(4 * 3.14159265358979323846);
  for (int z0 = 0; z0 <= nz; z0 += zb) {
    const int nzb = z0 + zb <= nz + 1 ? zb : nz + 1 - z0;
    Grid3 g = cell_grid(nzb, n2y, n2x);
    greens_rows_kernel<<<g.grid, g.block, 0, st>>>(rows, dz, dy, dxp, z0, nzb, n2y, n2x, four_pi,
                                                   (float)origin_value);
    SLICE_TRY(SOPHT_CHECK_LAUNCH());
    // the plan transforms zb planes; a short last batch leaves stale rows behind that are not kept
    SLICE_TRY(SOPHT_CUFFT(cufftExecD2Z(px, rows, rspec)));
    keep_bins_kernel<<<148 * 4, 256, 0, st>>>(S + (int64_t)z0 * plane, rspec, (int64_t)nzb * n2y, nkx, kx0, nxl, nx);
    SLICE_TRY(SOPHT_CHECK_LAUNCH());
  }
  for (int z = 0; z <= nz; ++z) SLICE_TRY(SOPHT_CUFFT(cufftExecZ2Z(py, S + z * plane, S + z * plane, CUFFT_FORWARD)));
  for (int z = 1; z < nz; ++z)
    SLICE_TRY(SOPHT_CUDA(cudaMemcpyAsync(S + (int64_t)(n2z - z) * plane, S + (int64_t)z * plane,
                                         sizeof(cufftDoubleComplex) * plane, cudaMemcpyDeviceToDevice, st)));
  SLICE_TRY(SOPHT_CUFFT(cufftExecZ2Z(pz, S, S, CUFFT_FORWARD)));
  {
    const float dxT = (float)dx;
    const float dx3 = dxT * dxT * dxT;
    const double scale = (double)dx3 / ((double)n2z * n2y * n2x);
    fold_slice_kernel<<<148 * 4, 256, 0, st>>>(gm, gn, S, nz, ny, nxl, scale);
    SLICE_TRY(SOPHT_CHECK_LAUNCH());
  }
  SLICE_TRY(SOPHT_CUDA(cudaStreamSynchronize(st)));
#undef SLICE_TRY
  cleanup();
  return SOPHT_OK;
}

// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct GenericPoisson : PoissonImpl {
  using C = typename FFTTypes<T>::C;
  int dim, nz, ny, nx;
  T* ghat = nullptr;       // (n2z, n2y, nx+1) real, scaled
  T* planes = nullptr;     // (nz, 2ny, 2nx) padded real planes
  C* spec = nullptr;       // (n2z, n2y, nx+1)
  cufftHandle p_r2c = 0, p_c2r = 0, p_z = 0;

  ~GenericPoisson() override {
    cudaFree(ghat);
    cudaFree(planes);
    cudaFree(spec);
    if (p_r2c) cufftDestroy(p_r2c);
    if (p_c2r) cufftDestroy(p_c2r);
    if (p_z) cufftDestroy(p_z);
  }

  int init(int dim_, int nz_, int ny_, int nx_, double dx, const double* mz, const double* my,
           const double* mx, double origin, cudaStream_t st) {
    dim = dim_;
    nz = dim == 3 ? nz_ : 1;
    ny = ny_;
    nx = nx_;
    const int n2z = dim == 3 ? 2 * nz : 1, n2y = 2 * ny, n2x = 2 * nx, nkx = nx + 1;
    if ((int64_t)n2z * n2y * n2x > 0x7fffffffLL)
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "poisson(generic): doubled grid exceeds 2^31 cells");
    int rc = build_green_hat<T>(&ghat, dim, nz, ny, nx, dx, mz, my, mx, origin, st);
    if (rc) return rc;
    SOPHT_CUDA(cudaMalloc(&planes, sizeof(T) * (size_t)nz * n2y * n2x));
    SOPHT_CUDA(cudaMalloc(&spec, sizeof(C) * (size_t)n2z * n2y * nkx));
    int n2[2] = {n2y, n2x};
    int rembed[2] = {n2y, n2x}, cembed[2] = {n2y, nkx};
    SOPHT_CUFFT(cufftPlanMany(&p_r2c, 2, n2, rembed, 1, n2y * n2x, cembed, 1, n2y * nkx,
                              FFTTypes<T>::R2C, nz));
    SOPHT_CUFFT(cufftPlanMany(&p_c2r, 2, n2, cembed, 1, n2y * nkx, rembed, 1, n2y * n2x,
                              FFTTypes<T>::C2R, nz));
    if (dim == 3) {
      int n1[1] = {n2z};
      const int S = n2y * nkx;
      SOPHT_CUFFT(cufftPlanMany(&p_z, 1, n1, n1, S, 1, n1, S, 1, FFTTypes<T>::C2C, S));
    }
    return SOPHT_OK;
  }

  int solve_scalar(View3<T> sol, View3<const T> rhs, cudaStream_t st) {
    const int n2z = dim == 3 ? 2 * nz : 1, n2y = 2 * ny, n2x = 2 * nx, nkx = nx + 1;
    const int64_t plane_spec = (int64_t)n2y * nkx;
    {
      Grid3 g = cell_grid(nz, n2y, n2x);
      if (g.grid.z > 4096) g.grid.z = 4096;
      pad_planes_kernel<T><<<g.grid, g.block, 0, st>>>(planes, rhs, nz, ny, nx);
      SOPHT_CHECK_LAUNCH();
    }
    SOPHT_CUFFT(cufftSetStream(p_r2c, st));
    SOPHT_CUFFT(FFTTypes<T>::r2c(p_r2c, planes, spec));
    g_launch_count++;
    if (dim == 3) {
      SOPHT_CUDA(cudaMemsetAsync(spec + plane_spec * nz, 0, sizeof(C) * plane_spec * (n2z - nz), st));
      SOPHT_CUFFT(cufftSetStream(p_z, st));
      SOPHT_CUFFT(FFTTypes<T>::c2c(p_z, spec, spec, CUFFT_FORWARD));
      g_launch_count++;
    }
    green_multiply_kernel<T, C><<<148 * 8, 256, 0, st>>>(spec, ghat, plane_spec * n2z);
    SOPHT_CHECK_LAUNCH();
    if (dim == 3) {
      SOPHT_CUFFT(FFTTypes<T>::c2c(p_z, spec, spec, CUFFT_INVERSE));
      g_launch_count++;
    }
    SOPHT_CUFFT(cufftSetStream(p_c2r, st));
    SOPHT_CUFFT(FFTTypes<T>::c2r(p_c2r, spec, planes));
    g_launch_count++;
    {
      Grid3 g = cell_grid(nz, ny, nx);
      if (g.grid.z > 4096) g.grid.z = 4096;
      crop_planes_kernel<T><<<g.grid, g.block, 0, st>>>(sol, planes, nz, ny, nx);
      SOPHT_CHECK_LAUNCH();
    }
    return SOPHT_OK;
  }

  int solve(const sopht_field_t* sol, const sopht_field_t* rhs, cudaStream_t st) override {
    const int gd = dim;
    const bool vec = sol->ndim == gd + 1;
    const int ncomp = vec ? (int)sol->shape[0] : 1;
    for (int c = 0; c < ncomp; ++c) {
      View3<T> s;
      View3<const T> r;
      const int o = vec ? 1 : 0;
      s.p = reinterpret_cast<T*>(sol->data) + (vec ? c * sol->stride[0] : 0);
      r.p = reinterpret_cast<const T*>(rhs->data) + (vec ? c * rhs->stride[0] : 0);
      if (gd == 3) {
        s.sz = sol->stride[o], s.sy = sol->stride[o + 1], s.sx = sol->stride[o + 2];
        r.sz = rhs->stride[o], r.sy = rhs->stride[o + 1], r.sx = rhs->stride[o + 2];
      } else {
        s.sz = 0, s.sy = sol->stride[o], s.sx = sol->stride[o + 1];
        r.sz = 0, r.sy = rhs->stride[o], r.sx = rhs->stride[o + 1];
      }
      int rc = solve_scalar(s, r, st);
      if (rc) return rc;
    }
    return SOPHT_OK;
  }

  const void* green_hat() const override { return ghat; }
};

template int build_green_hat<float>(float**, int, int, int, int, double, const double*, const double*,
                                     const double*, double, cudaStream_t);
template int build_green_hat<double>(double**, int, int, int, int, double, const double*, const double*,
                                      const double*, double, cudaStream_t);

PoissonImpl* make_generic_poisson(int dtype, int dim, int nz, int ny, int nx, double dx,
                                  const double* mz, const double* my, const double* mx,
                                  double origin, cudaStream_t st, int* rc) {
  if (dtype == SOPHT_F32) {
    auto* p = new GenericPoisson<float>();
    *rc = p->init(dim, nz, ny, nx, dx, mz, my, mx, origin, st);
    if (*rc) {
      delete p;
      return nullptr;
    }
    return p;
  }
  auto* p = new GenericPoisson<double>();
  *rc = p->init(dim, nz, ny, nx, dx, mz, my, mx, origin, st);
  if (*rc) {
    delete p;
    return nullptr;
  }
  return p;
}

}  // namespace sopht
