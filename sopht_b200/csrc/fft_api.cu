// Plain real <-> half-spectrum transforms over all axes of a 2-D / 3-D grid: the device counterpart of the
// reference's FFTPyFFTW{2,3}D plan objects and of its scipy rfftn / irfftn helper. cuFFT does batched 2-D R2C / C2R
// over the z planes and one strided 1-D C2C along z (the same plan shapes as the generic Poisson path); the inverse
// is normalised by a kernel of ours (pyfftw's default) and, like an FFTW c2r plan, may destroy its input.
//
// ref: sopht/numeric/eulerian_grid_ops/poisson_solver_3d/FFTPyFFTW3D.py:7-67, poisson_solver_2d/FFTPyFFTW2D.py:7-65,
//      poisson_solver_3d/scipy_fft_3d.py:7-15, poisson_solver_2d/scipy_fft_2d.py
#include <cufft.h>

#include "common.cuh"

namespace sopht {
namespace {

#define FFT_TRY(call)                                                                             \
  do {                                                                                            \
    cufftResult r__ = (call);                                                                     \
    if (r__ != CUFFT_SUCCESS)                                                                     \
      SOPHT_FAIL(SOPHT_ERR_CUFFT, "%s: %s failed with cufftResult %d", __func__, #call, (int)r__); \
  } while (0)

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* a, T s, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    a[q] *= s;
}

bool contiguous(const sopht_field_t* f, int dim, const int64_t* shape) {
  if (!f || !f->data || f->ndim != dim) return false;
  int64_t expect = 1;
  for (int d = dim - 1; d >= 0; --d) {
    if (f->shape[d] != shape[d] || (shape[d] > 1 && f->stride[d] != expect)) return false;
    expect *= shape[d];
  }
  return true;
}

}  // namespace
}  // namespace sopht

using namespace sopht;

struct sopht_fft {
  int dtype = SOPHT_F32, dim = 3;
  int nz = 1, ny = 1, nx = 1;
  cufftHandle p_r2c = 0, p_c2r = 0, p_z = 0;
  ~sopht_fft() {
    if (p_r2c) cufftDestroy(p_r2c);
    if (p_c2r) cufftDestroy(p_c2r);
    if (p_z) cufftDestroy(p_z);
  }
  int make_plans() {
    const int nkx = nx / 2 + 1;
    int n2[2] = {ny, nx};
    int rembed[2] = {ny, nx}, cembed[2] = {ny, nkx};
    const bool f32 = dtype == SOPHT_F32;
    FFT_TRY(cufftPlanMany(&p_r2c, 2, n2, rembed, 1, ny * nx, cembed, 1, ny * nkx, f32 ? CUFFT_R2C : CUFFT_D2Z, nz));
    FFT_TRY(cufftPlanMany(&p_c2r, 2, n2, cembed, 1, ny * nkx, rembed, 1, ny * nx, f32 ? CUFFT_C2R : CUFFT_Z2D, nz));
    if (dim == 3 && nz > 1) {
      int n1[1] = {nz};
      const int S = ny * nkx;
      FFT_TRY(cufftPlanMany(&p_z, 1, n1, n1, S, 1, n1, S, 1, f32 ? CUFFT_C2C : CUFFT_Z2Z, S));
    }
    return SOPHT_OK;
  }
  int check(const char* fn, const sopht_field_t* real_f, const sopht_field_t* cplx_f) const {
    const int64_t rs[3] = {nz, ny, nx}, cs[3] = {nz, ny, nx / 2 + 1};
    const int o = dim == 3 ? 0 : 1;
    if (!contiguous(real_f, dim, rs + o) || !contiguous(cplx_f, dim, cs + o))
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected contiguous (%s%d, %d) real and (..., %d) complex arrays of the plan",
                 fn, dim == 3 ? "nz, " : "", ny, nx, nx / 2 + 1);
    return SOPHT_OK;
  }
};

extern "C" {

int sopht_fft_create(sopht_fft_t* handle, int dtype, int dim, int nz, int ny, int nx) {
  if (!handle) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null handle pointer", __func__);
  SOPHT_CHECK_DTYPE(dtype);
  if ((dim != 2 && dim != 3) || ny < 1 || nx < 2 || (dim == 3 && nz < 1))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: dim must be 2 or 3 and the grid non-empty", __func__);
  auto* h = new sopht_fft();
  h->dtype = dtype, h->dim = dim, h->nz = dim == 3 ? nz : 1, h->ny = ny, h->nx = nx;
  const int rc = h->make_plans();
  if (rc) {
    delete h;
    return rc;
  }
  *handle = h;
  return SOPHT_OK;
}

int sopht_fft_forward(sopht_fft_t h, const sopht_field_t* field, const sopht_field_t* fourier_field, void* stream) {
  if (!h) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  int rc = h->check(__func__, field, fourier_field);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  FFT_TRY(cufftSetStream(h->p_r2c, st));
  if (h->dtype == SOPHT_F32)
    FFT_TRY(cufftExecR2C(h->p_r2c, reinterpret_cast<float*>(field->data),
                         reinterpret_cast<cufftComplex*>(fourier_field->data)));
  else
    FFT_TRY(cufftExecD2Z(h->p_r2c, reinterpret_cast<double*>(field->data),
                         reinterpret_cast<cufftDoubleComplex*>(fourier_field->data)));
  g_launch_count++;
  if (h->p_z) {
    FFT_TRY(cufftSetStream(h->p_z, st));
    if (h->dtype == SOPHT_F32) {
      auto* c = reinterpret_cast<cufftComplex*>(fourier_field->data);
      FFT_TRY(cufftExecC2C(h->p_z, c, c, CUFFT_FORWARD));
    } else {
      auto* c = reinterpret_cast<cufftDoubleComplex*>(fourier_field->data);
      FFT_TRY(cufftExecZ2Z(h->p_z, c, c, CUFFT_FORWARD));
    }
    g_launch_count++;
  }
  return SOPHT_OK;
}

int sopht_fft_inverse(sopht_fft_t h, const sopht_field_t* fourier_field, const sopht_field_t* field, void* stream) {
  if (!h) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  int rc = h->check(__func__, field, fourier_field);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (h->p_z) {  // in place along z: the input spectrum is consumed, as by an FFTW c2r plan
    FFT_TRY(cufftSetStream(h->p_z, st));
    if (h->dtype == SOPHT_F32) {
      auto* c = reinterpret_cast<cufftComplex*>(fourier_field->data);
      FFT_TRY(cufftExecC2C(h->p_z, c, c, CUFFT_INVERSE));
    } else {
      auto* c = reinterpret_cast<cufftDoubleComplex*>(fourier_field->data);
      FFT_TRY(cufftExecZ2Z(h->p_z, c, c, CUFFT_INVERSE));
    }
    g_launch_count++;
  }
  FFT_TRY(cufftSetStream(h->p_c2r, st));
  const int64_t n = (int64_t)h->nz * h->ny * h->nx;
  if (h->dtype == SOPHT_F32) {
    FFT_TRY(cufftExecC2R(h->p_c2r, reinterpret_cast<cufftComplex*>(fourier_field->data),
                         reinterpret_cast<float*>(field->data)));
    scale_kernel<float><<<148 * 4, 256, 0, st>>>(reinterpret_cast<float*>(field->data), 1.0f / (float)n, n);
  } else {
    FFT_TRY(cufftExecZ2D(h->p_c2r, reinterpret_cast<cufftDoubleComplex*>(fourier_field->data),
                         reinterpret_cast<double*>(field->data)));
    scale_kernel<double><<<148 * 4, 256, 0, st>>>(reinterpret_cast<double*>(field->data), 1.0 / (double)n, n);
  }
  g_launch_count++;
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_fft_destroy(sopht_fft_t h) {
  delete h;
  return SOPHT_OK;
}

}  // extern "C"
