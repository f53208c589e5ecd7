// Poisson solve with homogeneous Neumann walls on the cell-centred grid: the reference's
// FastDiagPoissonSolver{2,3}D, whose dense eigen-transforms (three tensordots each way, O(N n) flops) are replaced
// by their closed form. The matrices it diagonalises, tridiag(-1, 2, -1) / dx^2 with the two corner entries set to
// 1 / dx^2, are the second-difference operator with mirror ghost cells, so
//   * their eigenvectors are the DCT-II basis cos(pi k (j + 1/2) / n), eigenvalues (2 - 2 cos(pi k / n)) / dx^2,
//   * V diag(1 / lambda) V^-1 does not depend on the sign / scaling / order la.eig happens to return, and
//   * the solve equals a PERIODIC solve with the same three-point symbol on the grid mirrored about every wall
//     (period 2 n per axis; even data stay even), with the mean mode dropped like the reference's inf eigenvalue.
// Dataflow, sharing the unbounded generic path's structure (poisson_generic.cu), O(N log n), HBM-bound:
//   mirror : rhs (nz, ny, nx) -> even extension in y and x, planes (nz, 2ny, 2nx)
//   fft    : batched 2-D R2C over the nz planes; the z mirror images are copies of the plane SPECTRA
//            (plane 2nz-1-z = plane z), then one strided 1-D C2C along z (2nz)
//   symbol : spectrum *= 1 / (lz[kz] + ly[ky] + lx[kx]) / (8 nz ny nx), 0 for the mean mode; the symbol is rebuilt
//            from three 1-D tables, never stored on the grid
//   ifft   : C2C inverse along z, batched 2-D C2R over the first nz planes, crop the low corner
// The same pipeline without the mirror step is the PERIODIC solve (BASELINE config 4, an extension: the reference has
// nothing periodic): transform sizes n instead of 2 n, symbol either spectral ((2 pi m / L)^2, the
// default) or the three-point one (exact inverse of the 7-point Laplacian with wrap-around neighbours).
// ref: sopht/numeric/eulerian_grid_ops/poisson_solver_3d/FastDiagPoissonSolver3D.py:15-208,
//      poisson_solver_2d/FastDiagPoissonSolver2D.py:13-119
#include <cufft.h>
#include <math.h>

#include <vector>

#include "common.cuh"
#include "poisson.cuh"

namespace sopht {

namespace {

#define NEUMANN_CUFFT(call)                                                                   \
  do {                                                                                        \
    cufftResult r__ = (call);                                                                 \
    if (r__ != CUFFT_SUCCESS)                                                                 \
      SOPHT_FAIL(SOPHT_ERR_CUFFT, "%s: %s failed with cufftResult %d", __func__, #call, (int)r__); \
  } while (0)

template <typename T>
struct Fft;
template <>
struct Fft<float> {
  using C = cufftComplex;
  static constexpr cufftType R2C = CUFFT_R2C, C2R = CUFFT_C2R, C2C = CUFFT_C2C;
  static cufftResult r2c(cufftHandle p, float* in, C* out) { return cufftExecR2C(p, in, out); }
  static cufftResult c2r(cufftHandle p, C* in, float* out) { return cufftExecC2R(p, in, out); }
  static cufftResult c2c(cufftHandle p, C* in, C* out, int dir) { return cufftExecC2C(p, in, out, dir); }
};
template <>
struct Fft<double> {
  using C = cufftDoubleComplex;
  static constexpr cufftType R2C = CUFFT_D2Z, C2R = CUFFT_Z2D, C2C = CUFFT_Z2Z;
  static cufftResult r2c(cufftHandle p, double* in, C* out) { return cufftExecD2Z(p, in, out); }
  static cufftResult c2r(cufftHandle p, C* in, double* out) { return cufftExecZ2D(p, in, out); }
  static cufftResult c2c(cufftHandle p, C* in, C* out, int dir) { return cufftExecZ2Z(p, in, out, dir); }
};

// planes (np, 2ny, 2nx) <- rhs (np, ny, nx) mirrored about the y and x walls; x fastest, one thread per cell
template <typename T>
__global__ void __launch_bounds__(256)
    mirror_planes_kernel(T* __restrict__ dst, View3<const T> src, int np, int ny, int nx, int my, int mx) {
  // my = 2 ny, mx = 2 nx: even extension; my = ny, mx = nx (periodic solve): a plain gather into contiguous planes
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= mx || j >= my) return;
  const int si = i < nx ? i : 2 * nx - 1 - i, sj = j < ny ? j : 2 * ny - 1 - j;
  for (int k = blockIdx.z; k < np; k += gridDim.z) dst[((int64_t)k * my + j) * mx + i] = src(k, sj, si);
}

// spectrum plane 2nz-1-z <- plane z (the 2-D transform of the mirror image in z is the transform of the plane)
template <typename C>
__global__ void __launch_bounds__(256) mirror_z_spectrum_kernel(C* spec, int nz, int64_t plane) {
  const int64_t total = (int64_t)nz * plane;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t z = q / plane, r = q - z * plane;
    spec[(2 * nz - 1 - z) * plane + r] = spec[q];
  }
}

// spectrum *= norm / (lz[kz] + ly[ky] + lx[kx]); the mean mode (all three zero) -> 0 (FastDiagPoissonSolver3D.py:
// 143-146: its eigenvalue is set to inf before the reciprocal)
template <typename T, typename C>
__global__ void __launch_bounds__(256)
    neumann_symbol_kernel(C* spec, const T* __restrict__ lz, const T* __restrict__ ly, const T* __restrict__ lx,
                          int n2z, int n2y, int nkx, T norm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nkx || j >= n2y) return;
  const T lyx = ly[j] + lx[i];
  for (int k = blockIdx.z; k < n2z; k += gridDim.z) {
    const int64_t q = ((int64_t)k * n2y + j) * nkx + i;
    const T lam = lz[k] + lyx;
    // modes with k_d = n_d carry no energy for mirrored data (their symbol is finite anyway); only the mean is null
    const T s = (i == 0 && j == 0 && k == 0) ? T(0) : norm / lam;
    C v = spec[q];
    v.x *= s;
    v.y *= s;
    spec[q] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    crop_corner_kernel(View3<T> dst, const T* __restrict__ src, int np, int ny, int nx, int my, int mx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  for (int k = blockIdx.z; k < np; k += gridDim.z) dst(k, j, i) = src[((int64_t)k * my + j) * mx + i];
}

template <typename T>
struct NeumannPoisson : PoissonImpl {
  using C = typename Fft<T>::C;
  int dim = 3, nz = 1, ny = 0, nx = 0;
  int kind = 0;  // 0: Neumann walls (mirror), 1: periodic + spectral symbol, 2: periodic + three-point symbol
  int tz = 1, ty = 0, tx = 0;  // transform sizes: 2 n (mirror) or n (periodic)
  T *lz = nullptr, *ly = nullptr, *lx = nullptr;  // three-point symbol per axis on the mirrored period
  T* planes = nullptr;                            // (nz, 2ny, 2nx)
  C* spec = nullptr;                              // (n2z, 2ny, nx+1)
  T norm = T(1);
  cufftHandle p_r2c = 0, p_c2r = 0, p_z = 0;

  ~NeumannPoisson() override {
    cudaFree(lz);
    cudaFree(ly);
    cudaFree(lx);
    cudaFree(planes);
    cudaFree(spec);
    if (p_r2c) cufftDestroy(p_r2c);
    if (p_c2r) cufftDestroy(p_c2r);
    if (p_z) cufftDestroy(p_z);
  }
  const char* path_name() const override {
    return kind == 0 ? "neumann_mirror_fft" : kind == 1 ? "periodic_fft_spectral" : "periodic_fft_three_point";
  }

  int upload_symbol(T** dst, int period, int count, double dx, cudaStream_t st) {
    // three-point symbol on a period of `period` cells: (2 - 2 cos(2 pi k / period)) / dx^2 = 4 sin^2(pi k / period) / dx^2
    // spectral symbol: (2 pi m / (period dx))^2, m the signed wavenumber; both evaluated in double
    const double pi = 3.14159265358979323846;
    std::vector<T> h(count);
    for (int k = 0; k < count; ++k) {
      if (kind == 1) {
        const int m = k <= period / 2 ? k : k - period;
        const double w = 2.0 * pi * m / (period * dx);
        h[k] = (T)(w * w);
      } else {
        const double s = sin(pi * k / period);
        h[k] = (T)(4.0 * s * s / (dx * dx));
      }
    }
    SOPHT_CUDA(cudaMalloc(dst, sizeof(T) * count));
    SOPHT_CUDA(cudaMemcpyAsync(*dst, h.data(), sizeof(T) * count, cudaMemcpyHostToDevice, st));
    SOPHT_CUDA(cudaStreamSynchronize(st));  // h goes out of scope
    return SOPHT_OK;
  }

  int init(int kind_, int dim_, int nz_, int ny_, int nx_, double dx, cudaStream_t st) {
    kind = kind_;
    dim = dim_;
    nz = dim == 3 ? nz_ : 1;
    ny = ny_;
    nx = nx_;
    const int f = kind == 0 ? 2 : 1;
    tz = dim == 3 ? f * nz : 1, ty = f * ny, tx = f * nx;
    const int n2z = tz, n2y = ty, n2x = tx, nkx = tx / 2 + 1;
    if ((int64_t)n2z * n2y * n2x > 0x7fffffffLL)
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "poisson(neumann / periodic): transform grid exceeds 2^31 cells");
    int rc;
    if ((rc = upload_symbol(&lz, tz, n2z, dx, st))) return rc;  // dim 2: one entry, k = 0 -> 0
    if ((rc = upload_symbol(&ly, ty, n2y, dx, st))) return rc;
    if ((rc = upload_symbol(&lx, tx, nkx, dx, st))) return rc;
    norm = (T)(1.0 / ((double)n2z * n2y * n2x));  // cuFFT's inverse is unnormalised
    SOPHT_CUDA(cudaMalloc(&planes, sizeof(T) * (size_t)nz * n2y * n2x));
    SOPHT_CUDA(cudaMalloc(&spec, sizeof(C) * (size_t)n2z * n2y * nkx));
    int n2[2] = {n2y, n2x};
    int rembed[2] = {n2y, n2x}, cembed[2] = {n2y, nkx};
    NEUMANN_CUFFT(cufftPlanMany(&p_r2c, 2, n2, rembed, 1, n2y * n2x, cembed, 1, n2y * nkx, Fft<T>::R2C, nz));
    NEUMANN_CUFFT(cufftPlanMany(&p_c2r, 2, n2, cembed, 1, n2y * nkx, rembed, 1, n2y * n2x, Fft<T>::C2R, nz));
    if (dim == 3) {
      int n1[1] = {n2z};
      const int S = n2y * nkx;
      NEUMANN_CUFFT(cufftPlanMany(&p_z, 1, n1, n1, S, 1, n1, S, 1, Fft<T>::C2C, S));
    }
    return SOPHT_OK;
  }

  int solve_scalar(View3<T> sol, View3<const T> rhs, cudaStream_t st) {
    const int n2z = tz, n2y = ty, n2x = tx, nkx = tx / 2 + 1;
    const int64_t plane_spec = (int64_t)n2y * nkx;
    {
      Grid3 g = cell_grid(nz, n2y, n2x);
      if (g.grid.z > 4096) g.grid.z = 4096;
      SOPHT_PROF("poisson_neumann.mirror", st);
      mirror_planes_kernel<T><<<g.grid, g.block, 0, st>>>(planes, rhs, nz, ny, nx, n2y, n2x);
      SOPHT_CHECK_LAUNCH();
    }
    NEUMANN_CUFFT(cufftSetStream(p_r2c, st));
    NEUMANN_CUFFT(Fft<T>::r2c(p_r2c, planes, spec));
    g_launch_count++;
    if (dim == 3) {
      if (kind == 0) {
        mirror_z_spectrum_kernel<C><<<148 * 8, 256, 0, st>>>(spec, nz, plane_spec);
        SOPHT_CHECK_LAUNCH();
      }
      NEUMANN_CUFFT(cufftSetStream(p_z, st));
      NEUMANN_CUFFT(Fft<T>::c2c(p_z, spec, spec, CUFFT_FORWARD));
      g_launch_count++;
    }
    {
      Grid3 g = cell_grid(n2z, n2y, nkx);
      if (g.grid.z > 4096) g.grid.z = 4096;
      SOPHT_PROF("poisson_neumann.symbol", st);
      neumann_symbol_kernel<T, C><<<g.grid, g.block, 0, st>>>(spec, lz, ly, lx, n2z, n2y, nkx, norm);
      SOPHT_CHECK_LAUNCH();
    }
    if (dim == 3) {
      NEUMANN_CUFFT(Fft<T>::c2c(p_z, spec, spec, CUFFT_INVERSE));
      g_launch_count++;
    }
    NEUMANN_CUFFT(cufftSetStream(p_c2r, st));
    NEUMANN_CUFFT(Fft<T>::c2r(p_c2r, spec, planes));
    g_launch_count++;
    {
      Grid3 g = cell_grid(nz, ny, nx);
      if (g.grid.z > 4096) g.grid.z = 4096;
      SOPHT_PROF("poisson_neumann.crop", st);
      crop_corner_kernel<T><<<g.grid, g.block, 0, st>>>(sol, planes, nz, ny, nx, n2y, n2x);
      SOPHT_CHECK_LAUNCH();
    }
    return SOPHT_OK;
  }

  int solve(const sopht_field_t* sol, const sopht_field_t* rhs, cudaStream_t st) override {
    const bool vec = sol->ndim == dim + 1;
    const int ncomp = vec ? (int)sol->shape[0] : 1;
    const int o = vec ? 1 : 0;
    for (int c = 0; c < ncomp; ++c) {
      View3<T> s;
      View3<const T> r;
      s.p = reinterpret_cast<T*>(sol->data) + (vec ? c * sol->stride[0] : 0);
      r.p = reinterpret_cast<const T*>(rhs->data) + (vec ? c * rhs->stride[0] : 0);
      if (dim == 3) {
        s.sz = sol->stride[o], s.sy = sol->stride[o + 1], s.sx = sol->stride[o + 2];
        r.sz = rhs->stride[o], r.sy = rhs->stride[o + 1], r.sx = rhs->stride[o + 2];
      } else {
        s.sz = 0, s.sy = sol->stride[o], s.sx = sol->stride[o + 1];
        r.sz = 0, r.sy = rhs->stride[o], r.sx = rhs->stride[o + 1];
      }
      const int rc = solve_scalar(s, r, st);
      if (rc) return rc;
    }
    return SOPHT_OK;
  }
};

template <typename T>
PoissonImpl* make_neumann(int kind, int dim, int nz, int ny, int nx, double dx, cudaStream_t st, int* rc) {
  auto* p = new NeumannPoisson<T>();
  *rc = p->init(kind, dim, nz, ny, nx, dx, st);
  if (*rc) {
    delete p;
    return nullptr;
  }
  return p;
}

}  // namespace

PoissonImpl* make_neumann_poisson(int dtype, int dim, int nz, int ny, int nx, double dx, cudaStream_t st, int* rc) {
  if (neumann_dct_eligible(dim, nz, ny, nx)) return make_neumann_dct_poisson(dtype, nz, ny, nx, dx, st, rc);
  return dtype == SOPHT_F32 ? make_neumann<float>(0, dim, nz, ny, nx, dx, st, rc)
                            : make_neumann<double>(0, dim, nz, ny, nx, dx, st, rc);
}

PoissonImpl* make_periodic_poisson(int dtype, int three_point_symbol, int dim, int nz, int ny, int nx, double dx,
                                   cudaStream_t st, int* rc) {
  const int kind = three_point_symbol ? 2 : 1;
  return dtype == SOPHT_F32 ? make_neumann<float>(kind, dim, nz, ny, nx, dx, st, rc)
                            : make_neumann<double>(kind, dim, nz, ny, nx, dx, st, rc);
}

}  // namespace sopht
