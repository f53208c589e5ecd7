// Neumann-wall Poisson solve (the reference's FastDiagPoissonSolver3D) through SAME-LENGTH transforms.
//
// poisson_neumann.cu evaluates the closed form of the reference's eigen-decomposition - the DCT-II basis - as a periodic
// solve on the grid mirrored about every wall: transforms of (2nz, 2ny, 2nx), eight times the cells. Here the DCT-II of
// length n is taken from ONE real FFT of length n (Makhoul 1980): with the even / odd reordering
//     v[m] = x[2m],  v[n-1-m] = x[2m+1]   (m < n/2),     V = FFT_n(v),    w = exp(-i pi / (2n)),
//     X[k] = sum_j x[j] cos(pi (2j+1) k / (2n)) = Re(w^k V[k]),          X[n-k] = -Im(w^k V[k]),
// and back:  V[k] = conj(w^k) (X[k] - i X[n-k])  (X[n] := 0),  v = IFFT_n(V).
// Dataflow for one scalar field, every array real (nz, ny, nx) or its half spectrum:
//   reorder (y, x)            rhs -> R1                                        [reorder_yx_kernel]
//   2-D R2C, batch nz         R1 -> S1 (nz, ny, nx/2+1)                        [cuFFT]
//   2-D post-twiddle          S1 -> R2, plane z written at its reordered z     [post_yx_kernel]
//        X[ky,kx]    = Re(wy^ky (P + Q)) / 2,   X[ky,nx-kx] = Re(i wy^ky (P - Q)) / 2,
//        P = wx^kx V[ky,kx],  Q = conj(wx^kx) conj(V[ny-ky,kx])   (the 1-D rule applied along x, then along y)
//   1-D R2C along z (stride ny nx, batch ny nx)   R2 -> S2 (nz/2+1, ny, nx)    [cuFFT]
//   z post-twiddle x 1 / lambda x z pre-twiddle, in place on S2                [symbol_z_kernel]
//        the pair (kz, nz-kz) lives in ONE complex element, so the forward rule, the division by the eigenvalue
//        lz[kz] + ly[ky] + lx[kx] (mean mode -> 0) and the backward rule are one pass
//   1-D C2R along z           S2 -> R2
//   2-D pre-twiddle           R2 (plane at its reordered z) -> S1              [pre_yx_kernel]
//        V[ky,kx] = conj(wy^ky wx^kx) (X[ky,kx] - X[ny-ky,nx-kx] - i (X[ky,nx-kx] + X[ny-ky,kx]))
//   2-D C2R, batch nz         S1 -> R1
//   undo the (y, x) reordering  R1 -> solution                                 [unreorder_yx_kernel]
// Nine passes over n^3-sized arrays (72 B per cell in fp32) instead of the mirrored volume; needs even nz, ny, nx.
// ref: sopht/numeric/eulerian_grid_ops/poisson_solver_3d/FastDiagPoissonSolver3D.py:15-208
#include <cufft.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "poisson.cuh"

namespace sopht {

namespace {

#define DCT_CUFFT(call)                                                                       \
  do {                                                                                        \
    cufftResult r__ = (call);                                                                 \
    if (r__ != CUFFT_SUCCESS)                                                                 \
      SOPHT_FAIL(SOPHT_ERR_CUFFT, "%s: %s failed with cufftResult %d", __func__, #call, (int)r__); \
  } while (0)

template <typename T>
struct DctFft;
template <>
struct DctFft<float> {
  using C = cufftComplex;
  static constexpr cufftType R2C = CUFFT_R2C, C2R = CUFFT_C2R;
  static cufftResult r2c(cufftHandle p, float* in, C* out) { return cufftExecR2C(p, in, out); }
  static cufftResult c2r(cufftHandle p, C* in, float* out) { return cufftExecC2R(p, in, out); }
};
template <>
struct DctFft<double> {
  using C = cufftDoubleComplex;
  static constexpr cufftType R2C = CUFFT_D2Z, C2R = CUFFT_Z2D;
  static cufftResult r2c(cufftHandle p, double* in, C* out) { return cufftExecD2Z(p, in, out); }
  static cufftResult c2r(cufftHandle p, C* in, double* out) { return cufftExecZ2D(p, in, out); }
};

// position of source index s in the reordered sequence of length n, and its inverse
__device__ __forceinline__ int reorder_dst(int s, int n) { return (s & 1) ? n - 1 - (s >> 1) : (s >> 1); }
__device__ __forceinline__ int reorder_src(int d, int n) { return d < n / 2 ? 2 * d : 2 * (n - 1 - d) + 1; }

template <typename C, typename T>
__device__ __forceinline__ C cmulc(C a, T bx, T by) {  // a * (bx + i by)
  C r;
  r.x = a.x * bx - a.y * by;
  r.y = a.x * by + a.y * bx;
  return r;
}

// R1[z][j'][i'] = rhs(z, src(j'), src(i'))
template <typename T>
__global__ void __launch_bounds__(256) reorder_yx_kernel(T* __restrict__ dst, View3<const T> src, int nz, int ny, int nx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  const int si = reorder_src(i, nx), sj = reorder_src(j, ny);
  for (int k = blockIdx.z; k < nz; k += gridDim.z) dst[((int64_t)k * ny + j) * nx + i] = src(k, sj, si);
}

// sol(z, j, i) = R1[z][dst(j)][dst(i)]
template <typename T>
__global__ void __launch_bounds__(256) unreorder_yx_kernel(View3<T> dst, const T* __restrict__ src, int nz, int ny, int nx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  const int di = reorder_dst(i, nx), dj = reorder_dst(j, ny);
  for (int k = blockIdx.z; k < nz; k += gridDim.z) dst(k, j, i) = src[((int64_t)k * ny + dj) * nx + di];
}

// half spectrum of the reordered planes -> the planes' 2-D DCT-II, plane z stored at its reordered z position
template <typename T, typename C>
__global__ void __launch_bounds__(256)
    post_yx_kernel(T* __restrict__ out, const C* __restrict__ spec, const C* __restrict__ wy, const C* __restrict__ wx,
                   int nz, int ny, int nx) {
  const int nkx = nx / 2 + 1;
  const int kx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ky = blockIdx.y * blockDim.y + threadIdx.y;
  if (kx >= nkx || ky >= ny) return;
  const C w2 = wx[kx], w1 = wy[ky];
  const int kyc = ky ? ny - ky : 0;
  for (int z = blockIdx.z; z < nz; z += gridDim.z) {
    const C* sp = spec + (int64_t)z * ny * nkx;
    const C a = sp[(int64_t)ky * nkx + kx];
    C b = sp[(int64_t)kyc * nkx + kx];
    b.y = -b.y;
    const C p = cmulc(a, w2.x, w2.y), q = cmulc(b, w2.x, -w2.y);
    C s, d;  // (P + Q) / 2, i (P - Q) / 2
    s.x = T(0.5) * (p.x + q.x), s.y = T(0.5) * (p.y + q.y);
    d.x = T(-0.5) * (p.y - q.y), d.y = T(0.5) * (p.x - q.x);
    T* o = out + ((int64_t)reorder_dst(z, nz) * ny + ky) * nx;
    o[kx] = s.x * w1.x - s.y * w1.y;                               // Re(wy^ky (P + Q) / 2)
    if (kx > 0 && kx < nx - kx) o[nx - kx] = d.x * w1.x - d.y * w1.y;  // Re(wy^ky i (P - Q) / 2)
  }
}

// the inverse of post_yx_kernel: planes of DCT coefficients (at their reordered z) -> half spectrum for the 2-D C2R
template <typename T, typename C>
__global__ void __launch_bounds__(256)
    pre_yx_kernel(C* __restrict__ spec, const T* __restrict__ in, const C* __restrict__ wy, const C* __restrict__ wx,
                  int nz, int ny, int nx) {
  const int nkx = nx / 2 + 1;
  const int kx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ky = blockIdx.y * blockDim.y + threadIdx.y;
  if (kx >= nkx || ky >= ny) return;
  const C w2 = wx[kx], w1 = wy[ky];
  // conj(wy^ky wx^kx)
  const T cr = w1.x * w2.x - w1.y * w2.y, ci = -(w1.x * w2.y + w1.y * w2.x);
  for (int z = blockIdx.z; z < nz; z += gridDim.z) {
    const T* x = in + (int64_t)reorder_dst(z, nz) * ny * nx;
    const T x00 = x[(int64_t)ky * nx + kx];
    const T x01 = kx ? x[(int64_t)ky * nx + (nx - kx)] : T(0);                        // X[ky, nx - kx]
    const T x10 = ky ? x[(int64_t)(ny - ky) * nx + kx] : T(0);                        // X[ny - ky, kx]
    const T x11 = (ky && kx) ? x[(int64_t)(ny - ky) * nx + (nx - kx)] : T(0);         // X[ny - ky, nx - kx]
    C v;
    v.x = x00 - x11, v.y = -(x01 + x10);
    spec[((int64_t)z * ny + ky) * nkx + kx] = cmulc(v, cr, ci);
  }
}

// in place on the z half spectrum (nz/2+1, ny, nx): DCT coefficients of the pair (kz, nz - kz), division by the
// eigenvalue, and the spectrum of the scaled coefficients for the C2R back
template <typename T, typename C>
__global__ void __launch_bounds__(256)
    symbol_z_kernel(C* __restrict__ spec, const C* __restrict__ wz, const T* __restrict__ lz, const T* __restrict__ ly,
                    const T* __restrict__ lx, int nz, int ny, int nx, T norm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  const T lyx = ly[j] + lx[i];
  for (int k = blockIdx.z; k <= nz / 2; k += gridDim.z) {
    const int64_t q = ((int64_t)k * ny + j) * nx + i;
    const C w = wz[k];
    const C t = cmulc(spec[q], w.x, w.y);  // wz^kz V: X[kz] = Re, X[nz - kz] = -Im
    const T lam0 = lz[k] + lyx;
    const T s0 = (i == 0 && j == 0 && k == 0) ? T(0) : norm / lam0;  // the mean mode is dropped
    const T s1 = k ? norm / (lz[nz - k] + lyx) : T(0);               // X[nz] does not exist
    C v;
    v.x = t.x * s0, v.y = t.y * s1;  // X'[kz] - i X'[nz - kz] = Re t s0 + i Im t s1
    spec[q] = cmulc(v, w.x, -w.y);
  }
}

template <typename T>
struct NeumannDctPoisson : PoissonImpl {
  using C = typename DctFft<T>::C;
  int nz = 0, ny = 0, nx = 0;
  T *lz = nullptr, *ly = nullptr, *lx = nullptr;  // (2 - 2 cos(pi k / n)) / dx^2, k < n
  C *wz = nullptr, *wy = nullptr, *wx = nullptr;  // exp(-i pi k / (2n))
  T *r1 = nullptr, *r2 = nullptr;
  C *s1 = nullptr, *s2 = nullptr;
  T norm = T(1);
  cufftHandle p_r2c = 0, p_c2r = 0, p_zf = 0, p_zb = 0;

  ~NeumannDctPoisson() override {
    cudaFree(lz), cudaFree(ly), cudaFree(lx);
    cudaFree(wz), cudaFree(wy), cudaFree(wx);
    cudaFree(r1), cudaFree(r2), cudaFree(s1), cudaFree(s2);
    if (p_r2c) cufftDestroy(p_r2c);
    if (p_c2r) cufftDestroy(p_c2r);
    if (p_zf) cufftDestroy(p_zf);
    if (p_zb) cufftDestroy(p_zb);
  }
  const char* path_name() const override { return "neumann_dct"; }

  int upload_tables(T** l, C** w, int n, int nw, double dx, cudaStream_t st) {
    const double pi = 3.14159265358979323846;
    std::vector<T> hl(n);
    std::vector<C> hw(nw);
    for (int k = 0; k < n; ++k) {
      const double s = sin(pi * k / (2.0 * n));
      hl[k] = (T)(4.0 * s * s / (dx * dx));
    }
    for (int k = 0; k < nw; ++k) {
      hw[k].x = (T)cos(pi * k / (2.0 * n));
      hw[k].y = (T)(-sin(pi * k / (2.0 * n)));
    }
    SOPHT_CUDA(cudaMalloc(l, sizeof(T) * n));
    SOPHT_CUDA(cudaMalloc(w, sizeof(C) * nw));
    SOPHT_CUDA(cudaMemcpyAsync(*l, hl.data(), sizeof(T) * n, cudaMemcpyHostToDevice, st));
    SOPHT_CUDA(cudaMemcpyAsync(*w, hw.data(), sizeof(C) * nw, cudaMemcpyHostToDevice, st));
    SOPHT_CUDA(cudaStreamSynchronize(st));  // the host vectors go out of scope
    return SOPHT_OK;
  }

  int init(int nz_, int ny_, int nx_, double dx, cudaStream_t st) {
    nz = nz_, ny = ny_, nx = nx_;
    const int nkx = nx / 2 + 1, nkz = nz / 2 + 1;
    int rc;
    if ((rc = upload_tables(&lz, &wz, nz, nkz, dx, st))) return rc;
    if ((rc = upload_tables(&ly, &wy, ny, ny, dx, st))) return rc;
    if ((rc = upload_tables(&lx, &wx, nx, nkx, dx, st))) return rc;
    norm = (T)(1.0 / ((double)nz * ny * nx));  // the three unnormalised inverse transforms
    const size_t cells = (size_t)nz * ny * nx;
    if (cudaMalloc(&r1, sizeof(T) * cells) != cudaSuccess || cudaMalloc(&r2, sizeof(T) * cells) != cudaSuccess ||
        cudaMalloc(&s1, sizeof(C) * (size_t)nz * ny * nkx) != cudaSuccess ||
        cudaMalloc(&s2, sizeof(C) * (size_t)nkz * ny * nx) != cudaSuccess)
      SOPHT_FAIL(SOPHT_ERR_ALLOC, "poisson(neumann, dct): out of device memory for the work arrays");
    int n2[2] = {ny, nx};
    int rembed[2] = {ny, nx}, cembed[2] = {ny, nkx};
    DCT_CUFFT(cufftPlanMany(&p_r2c, 2, n2, rembed, 1, ny * nx, cembed, 1, ny * nkx, DctFft<T>::R2C, nz));
    DCT_CUFFT(cufftPlanMany(&p_c2r, 2, n2, cembed, 1, ny * nkx, rembed, 1, ny * nx, DctFft<T>::C2R, nz));
    int n1[1] = {nz};
    const int S = ny * nx;
    int re1[1] = {nz}, ce1[1] = {nkz};
    DCT_CUFFT(cufftPlanMany(&p_zf, 1, n1, re1, S, 1, ce1, S, 1, DctFft<T>::R2C, S));
    DCT_CUFFT(cufftPlanMany(&p_zb, 1, n1, ce1, S, 1, re1, S, 1, DctFft<T>::C2R, S));
    return SOPHT_OK;
  }

  int solve_scalar(View3<T> sol, View3<const T> rhs, cudaStream_t st) {
    const int nkx = nx / 2 + 1;
    Grid3 g = cell_grid(nz, ny, nx), gk = cell_grid(nz, ny, nkx), gz = cell_grid(nz / 2 + 1, ny, nx);
    if (g.grid.z > 1024) g.grid.z = 1024;
    if (gk.grid.z > 1024) gk.grid.z = 1024;
    {
      SOPHT_PROF("poisson_neumann.reorder", st);
      reorder_yx_kernel<T><<<g.grid, g.block, 0, st>>>(r1, rhs, nz, ny, nx);
      SOPHT_CHECK_LAUNCH();
    }
    DCT_CUFFT(cufftSetStream(p_r2c, st));
    DCT_CUFFT(DctFft<T>::r2c(p_r2c, r1, s1));
    {
      SOPHT_PROF("poisson_neumann.post_yx", st);
      post_yx_kernel<T, C><<<gk.grid, gk.block, 0, st>>>(r2, s1, wy, wx, nz, ny, nx);
      SOPHT_CHECK_LAUNCH();
    }
    DCT_CUFFT(cufftSetStream(p_zf, st));
    DCT_CUFFT(DctFft<T>::r2c(p_zf, r2, s2));
    {
      SOPHT_PROF("poisson_neumann.symbol", st);
      symbol_z_kernel<T, C><<<gz.grid, gz.block, 0, st>>>(s2, wz, lz, ly, lx, nz, ny, nx, norm);
      SOPHT_CHECK_LAUNCH();
    }
    DCT_CUFFT(cufftSetStream(p_zb, st));
    DCT_CUFFT(DctFft<T>::c2r(p_zb, s2, r2));
    {
      SOPHT_PROF("poisson_neumann.pre_yx", st);
      pre_yx_kernel<T, C><<<gk.grid, gk.block, 0, st>>>(s1, r2, wy, wx, nz, ny, nx);
      SOPHT_CHECK_LAUNCH();
    }
    DCT_CUFFT(cufftSetStream(p_c2r, st));
    DCT_CUFFT(DctFft<T>::c2r(p_c2r, s1, r1));
    g_launch_count += 4;
    {
      SOPHT_PROF("poisson_neumann.unreorder", st);
      unreorder_yx_kernel<T><<<g.grid, g.block, 0, st>>>(sol, r1, nz, ny, nx);
      SOPHT_CHECK_LAUNCH();
    }
    return SOPHT_OK;
  }

  int solve(const sopht_field_t* sol, const sopht_field_t* rhs, cudaStream_t st) override {
    const bool vec = sol->ndim == 4;
    const int ncomp = vec ? (int)sol->shape[0] : 1;
    const int o = vec ? 1 : 0;
    for (int c = 0; c < ncomp; ++c) {
      View3<T> s;
      View3<const T> r;
      s.p = reinterpret_cast<T*>(sol->data) + (vec ? c * sol->stride[0] : 0);
      r.p = reinterpret_cast<const T*>(rhs->data) + (vec ? c * rhs->stride[0] : 0);
      s.sz = sol->stride[o], s.sy = sol->stride[o + 1], s.sx = sol->stride[o + 2];
      r.sz = rhs->stride[o], r.sy = rhs->stride[o + 1], r.sx = rhs->stride[o + 2];
      const int rc = solve_scalar(s, r, st);
      if (rc) return rc;
    }
    return SOPHT_OK;
  }
};

template <typename T>
PoissonImpl* make_dct(int nz, int ny, int nx, double dx, cudaStream_t st, int* rc) {
  auto* p = new NeumannDctPoisson<T>();
  *rc = p->init(nz, ny, nx, dx, st);
  if (*rc) {
    delete p;
    return nullptr;
  }
  return p;
}

}  // namespace

// 3-D grids with even extents; SOPHT_NEUMANN_DCT=0 keeps the mirrored-grid form
bool neumann_dct_eligible(int dim, int nz, int ny, int nx) {
  static const int want = [] {
    const char* e = getenv("SOPHT_NEUMANN_DCT");
    return e ? atoi(e) : 1;
  }();
  return want && dim == 3 && nz >= 2 && ny >= 2 && nx >= 2 && nz % 2 == 0 && ny % 2 == 0 && nx % 2 == 0 &&
         (int64_t)nz * ny * nx <= 0x7fffffffLL;
}

PoissonImpl* make_neumann_dct_poisson(int dtype, int nz, int ny, int nx, double dx, cudaStream_t st, int* rc) {
  return dtype == SOPHT_F32 ? make_dct<float>(nz, ny, nx, dx, st, rc) : make_dct<double>(nz, ny, nx, dx, st, rc);
}

}  // namespace sopht
