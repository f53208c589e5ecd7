// 2-D stencil kernels behind the gen_*_pyst_kernel_2d factories. Same ghost-ring rule as 3-D
// (see stencils3d.cu). C1-sized 2-D grids (512x256) are L2 resident, so these are latency bound:
// one thread per cell, x fastest, nothing fancier.
#include "common.cuh"

namespace sopht {

template <typename T>
using CView2 = View2<const T>;
template <typename T>
static CView2<T> cview2(const View2<T>& v) {
  return CView2<T>{v.p, v.sy, v.sx};
}

#define CELL2D_PROLOGUE(ny, nx)                        \
  const int i = blockIdx.x * blockDim.x + threadIdx.x; \
  const int j = blockIdx.y * blockDim.y + threadIdx.y; \
  if (i >= (nx) || j >= (ny)) return;

#define IN_RING2(j, i, ny, nx, r) ((j) < (r) || (j) >= (ny) - (r) || (i) < (r) || (i) >= (nx) - (r))

template <typename T>
__global__ void __launch_bounds__(256)
    diffusion_flux2d_kernel(View2<T> flux, CView2<T> f, T p, int ny, int nx, int reset) {
  CELL2D_PROLOGUE(ny, nx)
  if (IN_RING2(j, i, ny, nx, 1)) {
    if (reset) flux(j, i) = T(0);
    return;
  }
  flux(j, i) = p * (f(j, i + 1) + f(j, i - 1) + f(j + 1, i) + f(j - 1, i) - T(4) * f(j, i));
}

template <typename T>
__device__ __forceinline__ T eno3_axis2d(T acc, T inv_dx, const T* f, const T* v) {
  const T c13 = T(1.0 / 3.0), c56 = T(5.0 / 6.0), c16 = T(1.0 / 6.0);
  const T front = (v[2] > -v[3]) ? (c13 * f[3] * v[3] + c56 * f[2] * v[2] - c16 * f[1] * v[1])
                                 : (c13 * f[2] * v[2] + c56 * f[3] * v[3] - c16 * f[4] * v[4]);
  acc = acc + inv_dx * front;
  const T back = (v[2] > -v[1]) ? (c13 * f[2] * v[2] + c56 * f[1] * v[1] - c16 * f[0] * v[0])
                                : (c13 * f[1] * v[1] + c56 * f[2] * v[2] - c16 * f[3] * v[3]);
  acc = acc - inv_dx * back;
  return acc;
}

template <typename T>
__global__ void __launch_bounds__(256)
    advection_flux_eno3_2d_kernel(View2<T> flux, CView2<T> f, CView2<T> vx, CView2<T> vy, T inv_dx,
                                  int ny, int nx) {
  CELL2D_PROLOGUE(ny, nx)
  if (IN_RING2(j, i, ny, nx, 2)) return;
  T fs[5], vs[5];
  T acc = flux(j, i);
#pragma unroll
  for (int o = 0; o < 5; ++o) {
    fs[o] = f(j, i + o - 2);
    vs[o] = vx(j, i + o - 2);
  }
  acc = eno3_axis2d(acc, inv_dx, fs, vs);
#pragma unroll
  for (int o = 0; o < 5; ++o) {
    fs[o] = f(j + o - 2, i);
    vs[o] = vy(j + o - 2, i);
  }
  acc = eno3_axis2d(acc, inv_dx, fs, vs);
  flux(j, i) = acc;
}

template <typename T>
__global__ void __launch_bounds__(256)
    outplane_curl2d_kernel(View2<T> cx, View2<T> cy, CView2<T> f, T p, int ny, int nx, int reset) {
  CELL2D_PROLOGUE(ny, nx)
  if (IN_RING2(j, i, ny, nx, 1)) {
    if (reset) {
      cx(j, i) = T(0);
      cy(j, i) = T(0);
    }
    return;
  }
  cx(j, i) = (f(j + 1, i) - f(j - 1, i)) * p;
  cy(j, i) = (f(j, i - 1) - f(j, i + 1)) * p;
}

// MODE 0: curl = c*p ; MODE 1: out += c*p
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
    inplane_curl2d_kernel(View2<T> out, CView2<T> fx, CView2<T> fy, T p, int ny, int nx) {
  CELL2D_PROLOGUE(ny, nx)
  if (IN_RING2(j, i, ny, nx, 1)) return;
  const T c = fy(j, i + 1) - fy(j, i - 1) - fx(j + 1, i) + fx(j - 1, i);
  if (MODE == 0)
    out(j, i) = c * p;
  else
    out(j, i) = out(j, i) + c * p;
}

template <typename T>
__global__ void __launch_bounds__(256)
    penalised_velocity_update2d_kernel(View2<T> w, CView2<T> px, CView2<T> py, CView2<T> ux,
                                       CView2<T> uy, T p, int ny, int nx) {
  CELL2D_PROLOGUE(ny, nx)
  if (IN_RING2(j, i, ny, nx, 1)) return;
  const T c = py(j, i + 1) - uy(j, i + 1) - py(j, i - 1) + uy(j, i - 1) - px(j + 1, i) +
              ux(j + 1, i) + px(j - 1, i) - ux(j - 1, i);
  w(j, i) = w(j, i) + c * p;
}

struct RampTable2 {
  double v[2][16];
};

// one thread per source cell on the shell of the inner box [w-1, n-w]^2 (see stencils3d.cu)
template <typename T>
__global__ void __launch_bounds__(128)
    penalise_shell2d_kernel(View2<T> f, int ny, int nx, int w, RampTable2 ramps) {
  const int ly = ny - 2 * w + 2, lx = nx - 2 * w + 2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int side = blockIdx.y;  // 0: y-front row, 1: y-back row, 2: x-front col, 3: x-back col
  int b, c;
  if (side < 2) {
    if (t >= lx) return;
    if (side == 1 && ly == 1) return;
    b = side ? ly - 1 : 0;
    c = t;
  } else {
    if (t >= ly - 2) return;
    if (side == 3 && lx == 1) return;
    b = t + 1;
    c = side == 3 ? lx - 1 : 0;
  }
  const int sj = b + w - 1, si = c + w - 1;
  const T val = f(sj, si);
  int jy0 = sj, jy1 = sj + 1, ix0 = si, ix1 = si + 1;
  if (sj == w - 1) jy0 = 0;
  if (sj == ny - w) jy1 = ny;
  if (si == w - 1) ix0 = 0;
  if (si == nx - w) ix1 = nx;
  for (int j = jy0; j < jy1; ++j) {
    const bool yr = j < w || j >= ny - w;
    const T ry = yr ? (T)ramps.v[1][j < w ? j : j - (ny - 2 * w)] : T(1);
    for (int i2 = ix0; i2 < ix1; ++i2) {
      const bool xr = i2 < w || i2 >= nx - w;
      T r = val;
      if (xr) r = r * (T)ramps.v[0][i2 < w ? i2 : i2 - (nx - 2 * w)];
      if (yr) r = r * ry;
      f(j, i2) = r;
    }
  }
}

static int chk_scalar2(const char* fn, const sopht_field_t* a) {
  if (!valid_field(a, 2, 2)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (ny, nx) field", fn);
  return SOPHT_OK;
}
static int chk_vector2(const char* fn, const sopht_field_t* a) {
  if (!valid_field(a, 3, 3) || a->shape[0] != 2)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (2, ny, nx) field", fn);
  return SOPHT_OK;
}
static bool same_grid2(const sopht_field_t* a, const sopht_field_t* b) {
  return a->shape[a->ndim - 1] == b->shape[b->ndim - 1] &&
         a->shape[a->ndim - 2] == b->shape[b->ndim - 2];
}

}  // namespace sopht

using namespace sopht;

#define RETURN_IF(rc_expr) \
  do {                     \
    int rc__ = (rc_expr);  \
    if (rc__) return rc__; \
  } while (0)

#define GRID2(f)                                    \
  const int ny = (int)(f)->shape[(f)->ndim - 2];    \
  const int nx = (int)(f)->shape[(f)->ndim - 1];    \
  if ((int64_t)ny * nx == 0) return SOPHT_OK;       \
  Grid3 g = cell_grid(1, ny, nx);                   \
  cudaStream_t st = as_stream(stream);              \
  if (g.grid.y > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: ny too large", __func__);

extern "C" {

int sopht_diffusion_flux_2d(int dtype, const sopht_field_t* diffusion_flux,
                            const sopht_field_t* field, double prefactor, int reset_ghost_zone,
                            void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(chk_scalar2(__func__, diffusion_flux));
  RETURN_IF(chk_scalar2(__func__, field));
  if (!same_shape(diffusion_flux, field)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: shapes differ", __func__);
  GRID2(field)
  if (dtype == SOPHT_F32)
    diffusion_flux2d_kernel<float><<<g.grid, g.block, 0, st>>>(
        scalar2<float>(diffusion_flux), cview2(scalar2<float>(field)), (float)prefactor, ny, nx,
        reset_ghost_zone);
  else
    diffusion_flux2d_kernel<double><<<g.grid, g.block, 0, st>>>(
        scalar2<double>(diffusion_flux), cview2(scalar2<double>(field)), prefactor, ny, nx,
        reset_ghost_zone);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_advection_flux_eno3_2d(int dtype, const sopht_field_t* advection_flux,
                                 const sopht_field_t* field, const sopht_field_t* velocity,
                                 double inv_dx, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(chk_scalar2(__func__, advection_flux));
  RETURN_IF(chk_scalar2(__func__, field));
  RETURN_IF(chk_vector2(__func__, velocity));
  if (!same_shape(advection_flux, field) || !same_grid2(field, velocity))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: shapes differ", __func__);
  GRID2(field)
  if (dtype == SOPHT_F32)
    advection_flux_eno3_2d_kernel<float><<<g.grid, g.block, 0, st>>>(
        scalar2<float>(advection_flux), cview2(scalar2<float>(field)),
        cview2(comp2<float>(velocity, 0)), cview2(comp2<float>(velocity, 1)), (float)inv_dx, ny, nx);
  else
    advection_flux_eno3_2d_kernel<double><<<g.grid, g.block, 0, st>>>(
        scalar2<double>(advection_flux), cview2(scalar2<double>(field)),
        cview2(comp2<double>(velocity, 0)), cview2(comp2<double>(velocity, 1)), inv_dx, ny, nx);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_outplane_field_curl_2d(int dtype, const sopht_field_t* curl, const sopht_field_t* field,
                                 double prefactor, int reset_ghost_zone, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(chk_vector2(__func__, curl));
  RETURN_IF(chk_scalar2(__func__, field));
  if (!same_grid2(curl, field)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: shapes differ", __func__);
  GRID2(field)
  if (dtype == SOPHT_F32)
    outplane_curl2d_kernel<float><<<g.grid, g.block, 0, st>>>(
        comp2<float>(curl, 0), comp2<float>(curl, 1), cview2(scalar2<float>(field)),
        (float)prefactor, ny, nx, reset_ghost_zone);
  else
    outplane_curl2d_kernel<double><<<g.grid, g.block, 0, st>>>(
        comp2<double>(curl, 0), comp2<double>(curl, 1), cview2(scalar2<double>(field)), prefactor,
        ny, nx, reset_ghost_zone);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

static int inplane_dispatch(int dtype, int mode, const sopht_field_t* out,
                            const sopht_field_t* field, double prefactor, void* stream,
                            const char* fn) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(chk_scalar2(fn, out));
  RETURN_IF(chk_vector2(fn, field));
  if (!same_grid2(out, field)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: shapes differ", fn);
  GRID2(field)
#define LAUNCH_IC(T, M)                                                                       \
  inplane_curl2d_kernel<T, M><<<g.grid, g.block, 0, st>>>(scalar2<T>(out),                    \
                                                          cview2(comp2<T>(field, 0)),         \
                                                          cview2(comp2<T>(field, 1)),         \
                                                          (T)prefactor, ny, nx)
  if (dtype == SOPHT_F32) {
    if (mode == 0)
      LAUNCH_IC(float, 0);
    else
      LAUNCH_IC(float, 1);
  } else {
    if (mode == 0)
      LAUNCH_IC(double, 0);
    else
      LAUNCH_IC(double, 1);
  }
#undef LAUNCH_IC
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_inplane_field_curl_2d(int dtype, const sopht_field_t* curl, const sopht_field_t* field,
                                double prefactor, void* stream) {
  return inplane_dispatch(dtype, 0, curl, field, prefactor, stream, __func__);
}

int sopht_update_vorticity_from_velocity_forcing_2d(int dtype, const sopht_field_t* vorticity_field,
                                                    const sopht_field_t* velocity_forcing_field,
                                                    double prefactor, void* stream) {
  return inplane_dispatch(dtype, 1, vorticity_field, velocity_forcing_field, prefactor, stream,
                          __func__);
}

int sopht_update_vorticity_from_penalised_velocity_2d(int dtype,
                                                      const sopht_field_t* vorticity_field,
                                                      const sopht_field_t* penalised_velocity_field,
                                                      const sopht_field_t* velocity_field,
                                                      double prefactor, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(chk_scalar2(__func__, vorticity_field));
  RETURN_IF(chk_vector2(__func__, penalised_velocity_field));
  RETURN_IF(chk_vector2(__func__, velocity_field));
  if (!same_grid2(vorticity_field, velocity_field) ||
      !same_shape(penalised_velocity_field, velocity_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: shapes differ", __func__);
  GRID2(vorticity_field)
  if (dtype == SOPHT_F32)
    penalised_velocity_update2d_kernel<float><<<g.grid, g.block, 0, st>>>(
        scalar2<float>(vorticity_field), cview2(comp2<float>(penalised_velocity_field, 0)),
        cview2(comp2<float>(penalised_velocity_field, 1)), cview2(comp2<float>(velocity_field, 0)),
        cview2(comp2<float>(velocity_field, 1)), (float)prefactor, ny, nx);
  else
    penalised_velocity_update2d_kernel<double><<<g.grid, g.block, 0, st>>>(
        scalar2<double>(vorticity_field), cview2(comp2<double>(penalised_velocity_field, 0)),
        cview2(comp2<double>(penalised_velocity_field, 1)),
        cview2(comp2<double>(velocity_field, 0)), cview2(comp2<double>(velocity_field, 1)),
        prefactor, ny, nx);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_penalise_field_boundary_2d(int dtype, const sopht_field_t* field, int width,
                                     const double* ramp_x, const double* ramp_y, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (width < 0 || width > 8) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: width must be in [0, 8]", __func__);
  if (width == 0) return SOPHT_OK;
  RETURN_IF(chk_scalar2(__func__, field));
  if (!ramp_x || !ramp_y) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null ramp table", __func__);
  const int ny = (int)field->shape[0], nx = (int)field->shape[1];
  if (ny < 2 * width || nx < 2 * width)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid smaller than twice the penalisation width", __func__);
  RampTable2 tab;
  for (int q = 0; q < 2 * width; ++q) {
    tab.v[0][q] = ramp_x[q];
    tab.v[1][q] = ramp_y[q];
  }
  const int ly = ny - 2 * width + 2, lx = nx - 2 * width + 2;
  const int span = lx > ly ? lx : ly;
  dim3 block(128, 1, 1), grid((span + 127) / 128, 4, 1);
  if (dtype == SOPHT_F32)
    penalise_shell2d_kernel<float><<<grid, block, 0, as_stream(stream)>>>(scalar2<float>(field), ny,
                                                                          nx, width, tab);
  else
    penalise_shell2d_kernel<double><<<grid, block, 0, as_stream(stream)>>>(scalar2<double>(field),
                                                                           ny, nx, width, tab);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

}  // extern "C"
