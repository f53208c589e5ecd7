// 3-D stencil kernels behind the gen_*_pyst_kernel_3d factories (one kernel per reference op).
// Ghost-ring rule of the reference's code generator: a kernel writes only cells whose whole stencil
// lies inside the array, with ONE ring width (max |offset| over all accesses) used on every axis.
// Cells outside are left untouched, or zeroed in the same launch when `reset_ghost_zone` is set
// (the reference does that with 6/18 extra sliced launches).
//
// Mapping: x fastest across the warp (coalesced 128B rows), 4 rows per CTA, CTA marches over a
// chunk of z planes so the +-z neighbours are L1/L2 hits. These are the API-level kernels; the
// simulator-level step uses the fused kernels in fused_step3d.cu.
#include "common.cuh"

namespace sopht {

template <typename T>
using CView3 = View3<const T>;

template <typename T>
static CView3<T> cview(const View3<T>& v) {
  return CView3<T>{v.p, v.sz, v.sy, v.sx};
}

#define CELL_LOOP_PROLOGUE(nz, ny, nx)                         \
  const int i = blockIdx.x * blockDim.x + threadIdx.x;         \
  const int j = blockIdx.y * blockDim.y + threadIdx.y;         \
  if (i >= (nx) || j >= (ny)) return;                          \
  const int kchunk = ((nz) + gridDim.z - 1) / gridDim.z;       \
  const int k0 = blockIdx.z * kchunk;                          \
  const int k1 = min(k0 + kchunk, (nz));

#define IN_RING(k, j, i, nz, ny, nx, r) \
  ((k) < (r) || (k) >= (nz) - (r) || (j) < (r) || (j) >= (ny) - (r) || (i) < (r) || (i) >= (nx) - (r))

static Grid3 stencil_grid(int nz, int ny, int nx) {
  Grid3 g = cell_grid(nz, ny, nx);
  // aim for >= ~8 CTAs per SM overall while keeping z-chunks long enough for plane reuse
  const int64_t xy = (int64_t)g.grid.x * g.grid.y;
  int64_t gz = (148 * 8 + xy - 1) / xy;
  if (gz < 1) gz = 1;
  if (gz > nz) gz = nz;
  g.grid.z = (unsigned)gz;
  return g;
}

// ---- diffusion flux -------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    diffusion_flux_kernel(View3<T> flux, CView3<T> f, T p, int nz, int ny, int nx, int reset) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 1)) {
      if (reset) flux(k, j, i) = T(0);
      continue;
    }
    flux(k, j, i) = p * (f(k + 1, j, i) + f(k - 1, j, i) + f(k, j + 1, i) + f(k, j - 1, i) +
                         f(k, j, i + 1) + f(k, j, i - 1) - T(6) * f(k, j, i));
  }
}

// ---- curl / forcing update ------------------------------------------------------------------
// MODE 0: curl = p*c (ring optional zero); MODE 1: out += p*c (interior only)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
    curl_kernel(View3<T> ox, View3<T> oy, View3<T> oz, CView3<T> fx, CView3<T> fy, CView3<T> fz,
                T p, int nz, int ny, int nx, int reset) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 1)) {
      if (MODE == 0 && reset) {
        ox(k, j, i) = T(0);
        oy(k, j, i) = T(0);
        oz(k, j, i) = T(0);
      }
      continue;
    }
    const T cx = fz(k, j + 1, i) - fz(k, j - 1, i) - fy(k + 1, j, i) + fy(k - 1, j, i);
    const T cy = fx(k + 1, j, i) - fx(k - 1, j, i) - fz(k, j, i + 1) + fz(k, j, i - 1);
    const T cz = fy(k, j, i + 1) - fy(k, j, i - 1) - fx(k, j + 1, i) + fx(k, j - 1, i);
    if (MODE == 0) {
      ox(k, j, i) = p * cx;
      oy(k, j, i) = p * cy;
      oz(k, j, i) = p * cz;
    } else {
      ox(k, j, i) = ox(k, j, i) + p * cx;
      oy(k, j, i) = oy(k, j, i) + p * cy;
      oz(k, j, i) = oz(k, j, i) + p * cz;
    }
  }
}

// vorticity += p * curl_c(u_pen - u), term order as in the reference stencil
template <typename T>
__global__ void __launch_bounds__(256)
    penalised_velocity_update_kernel(View3<T> ox, View3<T> oy, View3<T> oz, CView3<T> px,
                                     CView3<T> py, CView3<T> pz, CView3<T> ux, CView3<T> uy,
                                     CView3<T> uz, T p, int nz, int ny, int nx) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 1)) continue;
    const T cx = pz(k, j + 1, i) - uz(k, j + 1, i) - pz(k, j - 1, i) + uz(k, j - 1, i) -
                 py(k + 1, j, i) + uy(k + 1, j, i) + py(k - 1, j, i) - uy(k - 1, j, i);
    const T cy = px(k + 1, j, i) - ux(k + 1, j, i) - px(k - 1, j, i) + ux(k - 1, j, i) -
                 pz(k, j, i + 1) + uz(k, j, i + 1) + pz(k, j, i - 1) - uz(k, j, i - 1);
    const T cz = py(k, j, i + 1) - uy(k, j, i + 1) - py(k, j, i - 1) + uy(k, j, i - 1) -
                 px(k, j + 1, i) + ux(k, j + 1, i) + px(k, j - 1, i) - ux(k, j - 1, i);
    ox(k, j, i) = ox(k, j, i) + p * cx;
    oy(k, j, i) = oy(k, j, i) + p * cy;
    oz(k, j, i) = oz(k, j, i) + p * cz;
  }
}

// ---- divergence ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    divergence_kernel(View3<T> div, CView3<T> fx, CView3<T> fy, CView3<T> fz, T inv_dx, int nz,
                      int ny, int nx, int reset) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 1)) {
      if (reset) div(k, j, i) = T(0);
      continue;
    }
    div(k, j, i) = T(0.5) * inv_dx *
                   (fx(k, j, i + 1) - fx(k, j, i - 1) + fy(k, j + 1, i) - fy(k, j - 1, i) +
                    fz(k + 1, j, i) - fz(k - 1, j, i));
  }
}

// ---- vorticity stretching flux --------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    stretching_flux_kernel(View3<T> qx, View3<T> qy, View3<T> qz, CView3<T> wx, CView3<T> wy,
                           CView3<T> wz, CView3<T> ux, CView3<T> uy, CView3<T> uz, T p, int nz,
                           int ny, int nx) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 1)) {
      qx(k, j, i) = T(0);
      qy(k, j, i) = T(0);
      qz(k, j, i) = T(0);
      continue;
    }
    const T a = wx(k, j, i), b = wy(k, j, i), c = wz(k, j, i);
    qx(k, j, i) = p * (a * (ux(k, j, i + 1) - ux(k, j, i - 1)) + b * (ux(k, j + 1, i) - ux(k, j - 1, i)) +
                       c * (ux(k + 1, j, i) - ux(k - 1, j, i)));
    qy(k, j, i) = p * (a * (uy(k, j, i + 1) - uy(k, j, i - 1)) + b * (uy(k, j + 1, i) - uy(k, j - 1, i)) +
                       c * (uy(k + 1, j, i) - uy(k - 1, j, i)));
    qz(k, j, i) = p * (a * (uz(k, j, i + 1) - uz(k, j, i - 1)) + b * (uz(k, j + 1, i) - uz(k, j - 1, i)) +
                       c * (uz(k + 1, j, i) - uz(k - 1, j, i)));
  }
}

// ---- ENO3 conservative advection flux ---------------------------------------------------------------
// One axis: f/v sampled at offsets -2..+2 along that axis. Returns (front face flux - back face flux)
// accumulated exactly in the reference's order: acc = acc + inv_dx*F_front; acc = acc - inv_dx*F_back.
template <typename T>
__device__ __forceinline__ T eno3_axis_accumulate(T acc, T inv_dx, const T* f, const T* v) {
  // index 2 is the centre cell
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
    advection_flux_eno3_kernel(View3<T> flux, CView3<T> f, CView3<T> vx, CView3<T> vy, CView3<T> vz,
                               T inv_dx, int nz, int ny, int nx) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 2)) continue;
    T fs[5], vs[5];
    T acc = flux(k, j, i);
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      fs[o] = f(k, j, i + o - 2);
      vs[o] = vx(k, j, i + o - 2);
    }
    acc = eno3_axis_accumulate(acc, inv_dx, fs, vs);
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      fs[o] = f(k, j + o - 2, i);
      vs[o] = vy(k, j + o - 2, i);
    }
    acc = eno3_axis_accumulate(acc, inv_dx, fs, vs);
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      fs[o] = f(k + o - 2, j, i);
      vs[o] = vz(k + o - 2, j, i);
    }
    acc = eno3_axis_accumulate(acc, inv_dx, fs, vs);
    flux(k, j, i) = acc;
  }
}

// ---- 1-D Laplacian filter flux -----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    laplacian_filter_flux_kernel(View3<T> flux, CView3<T> f, int dk, int dj, int di, int nz, int ny,
                                 int nx) {
  CELL_LOOP_PROLOGUE(nz, ny, nx)
  for (int k = k0; k < k1; ++k) {
    if (IN_RING(k, j, i, nz, ny, nx, 1)) continue;
    flux(k, j, i) =
        T(0.25) * (-f(k + dk, j + dj, i + di) - f(k - dk, j - dj, i - di) + T(2) * f(k, j, i));
  }
}

// ---- boundary penalisation ---------------------------------------------------------------------------
// Separable restatement of the reference's sequential x, y, z copy-and-scale: every ring cell takes the
// value of its clamped source cell (coordinates clamped into [w-1, n-w]) times the ramps s_x*s_y*s_z in
// that order. One thread per SOURCE cell on the shell of the inner box writes all its targets, so the
// in-place update has no read/write hazard between threads.
struct RampTable {
  double v[3][16];  // per axis: front ramp (w values) then back ramp (w values); w <= 8
};

// z_faces: bit 0 = the low-z face of this view is a global boundary, bit 1 = the high-z face is (both set
// for a whole grid; a z-slab of a decomposed grid sets only the faces it owns, its other z side is interior).
// One launch covers every face of every component: blockIdx.z = (component * 3 + face) * 2 + side.
template <typename T>
struct ShellViews {
  View3<T> f[3];
};
template <typename T>
__global__ void __launch_bounds__(128)
    penalise_shell_kernel(ShellViews<T> views, int nz, int ny, int nx, int w, RampTable ramps, int z_faces) {
  // face 0: z-shell (the owned z faces of the inner box, all y,x in the box)
  // face 1: y-shell excluding z-shell cells; face 2: x-shell excluding z- and y-shell cells
  const int face = (blockIdx.z >> 1) % 3;
  const int comp = (blockIdx.z >> 1) / 3;
  const View3<T> f = comp == 0 ? views.f[0] : (comp == 1 ? views.f[1] : views.f[2]);
  const bool zlo = z_faces & 1, zhi = z_faces & 2;
  const int za = zlo ? w - 1 : 0, zb = zhi ? nz - w : nz - 1;  // inner box z extent [za, zb]
  const int ly = ny - 2 * w + 2, lx = nx - 2 * w + 2;           // inner box extent in y, x
  const int zin0 = za + (zlo ? 1 : 0), zin1 = zb - (zhi ? 1 : 0);  // z planes not on an owned z-shell
  int sk, b, c;                                                  // source z index, box-local (y,x)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int u = blockIdx.y;
  const int s = blockIdx.z & 1;  // 0: front, 1: back
  if (face == 0) {
    if (t >= lx || u >= ly) return;
    if (!z_faces) return;
    if (s == 0 && !zlo) return;
    if (s == 1 && (!zhi || (zlo && zb == za))) return;
    sk = s ? zb : za;
    b = u;
    c = t;
  } else if (face == 1) {
    if (t >= lx || u > zin1 - zin0) return;
    if (s == 1 && ly == 1) return;
    sk = zin0 + u;
    b = s ? ly - 1 : 0;
    c = t;
  } else {
    if (t >= ly - 2 || u > zin1 - zin0) return;
    if (s == 1 && lx == 1) return;
    sk = zin0 + u;
    b = t + 1;
    c = s ? lx - 1 : 0;
  }
  const int sj = b + w - 1, si = c + w - 1;  // source cell
  const T val = f(sk, sj, si);
  // target ranges along each axis
  int kz0 = sk, kz1 = sk + 1, jy0 = sj, jy1 = sj + 1, ix0 = si, ix1 = si + 1;
  if (zlo && sk == w - 1) kz0 = 0;
  if (zhi && sk == nz - w) kz1 = nz;
  if (sj == w - 1) jy0 = 0;
  if (sj == ny - w) jy1 = ny;
  if (si == w - 1) ix0 = 0;
  if (si == nx - w) ix1 = nx;
  for (int k = kz0; k < kz1; ++k) {
    const bool zr = (zlo && k < w) || (zhi && k >= nz - w);
    const T rz = zr ? (T)ramps.v[2][k < w ? k : k - (nz - 2 * w)] : T(1);
    for (int j = jy0; j < jy1; ++j) {
      const bool yr = j < w || j >= ny - w;
      const T ry = yr ? (T)ramps.v[1][j < w ? j : j - (ny - 2 * w)] : T(1);
      for (int i2 = ix0; i2 < ix1; ++i2) {
        const bool xr = i2 < w || i2 >= nx - w;
        T r = val;
        if (xr) r = r * (T)ramps.v[0][i2 < w ? i2 : i2 - (nx - 2 * w)];
        if (yr) r = r * ry;
        if (zr) r = r * rz;
        f(k, j, i2) = r;
      }
    }
  }
}

template <typename T>
static int penalise3d_impl(const ShellViews<T>& views, int ncomp, int nz, int ny, int nx, int w,
                           const double* rx, const double* ry, const double* rz, int z_faces, cudaStream_t st) {
  RampTable tab;
  for (int q = 0; q < 2 * w; ++q) {
    tab.v[0][q] = rx[q];
    tab.v[1][q] = ry[q];
    tab.v[2][q] = rz[q];
  }
  const bool zlo = z_faces & 1, zhi = z_faces & 2;
  const int za = zlo ? w - 1 : 0, zb = zhi ? nz - w : nz - 1;
  const int nin = (zb - (zhi ? 1 : 0)) - (za + (zlo ? 1 : 0)) + 1;  // z planes off the owned z-shells
  const int ly = ny - 2 * w + 2, lx = nx - 2 * w + 2;
  // thread extents of the three faces: (lx, ly), (lx, nin), (ly - 2, nin); the kernel drops the excess
  const int ext_t = lx > ly - 2 ? lx : ly - 2;
  int ext_u = z_faces ? ly : 1;
  if (nin > ext_u) ext_u = nin;
  SOPHT_PROF("penalise_field_boundary", st);
  dim3 block(128, 1, 1), grid((ext_t + 127) / 128, ext_u, ncomp * 6);
  penalise_shell_kernel<T><<<grid, block, 0, st>>>(views, nz, ny, nx, w, tab, z_faces);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

// ---- host-side checks ----------------------------------------------------------------------------------
static int check_scalar3(const char* fn, const sopht_field_t* a) {
  if (!valid_field(a, 3, 3)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (nz, ny, nx) field", fn);
  return SOPHT_OK;
}
static int check_vector3(const char* fn, const sopht_field_t* a) {
  if (!valid_field(a, 4, 4) || a->shape[0] != 3)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (3, nz, ny, nx) field", fn);
  return SOPHT_OK;
}
static bool same_grid(const sopht_field_t* a, const sopht_field_t* b) {
  // compares the trailing three axes
  for (int d = 1; d <= 3; ++d)
    if (a->shape[a->ndim - d] != b->shape[b->ndim - d]) return false;
  return true;
}
static bool fits_int(const sopht_field_t* a) {
  for (int d = 0; d < a->ndim; ++d)
    if (a->shape[d] > 0x7fffffff) return false;
  return true;
}

#define GRID_DIMS(f)                               \
  const int nz = (int)(f)->shape[(f)->ndim - 3];   \
  const int ny = (int)(f)->shape[(f)->ndim - 2];   \
  const int nx = (int)(f)->shape[(f)->ndim - 1];   \
  if ((int64_t)nz * ny * nx == 0) return SOPHT_OK; \
  Grid3 g = stencil_grid(nz, ny, nx);              \
  if (g.grid.y > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: ny too large", __func__);

template <typename T>
static int diffusion_flux_impl(const sopht_field_t* flux, const sopht_field_t* field, double p,
                               int reset, cudaStream_t st) {
  GRID_DIMS(field)
  if (field->ndim == 3) {
    diffusion_flux_kernel<T><<<g.grid, g.block, 0, st>>>(scalar3<T>(flux), cview(scalar3<T>(field)),
                                                         (T)p, nz, ny, nx, reset);
    SOPHT_CHECK_LAUNCH();
  } else {
    for (int c = 0; c < 3; ++c) {
      diffusion_flux_kernel<T><<<g.grid, g.block, 0, st>>>(
          comp3<T>(flux, c), cview(comp3<T>(field, c)), (T)p, nz, ny, nx, reset);
      SOPHT_CHECK_LAUNCH();
    }
  }
  return SOPHT_OK;
}

template <typename T, int MODE>
static int curl_impl(const sopht_field_t* out, const sopht_field_t* field, double p, int reset,
                     cudaStream_t st) {
  GRID_DIMS(field)
  SOPHT_PROF(MODE == 1 ? "update_vorticity_from_velocity_forcing" : "curl", st);
  if (MODE == 1 && ns3d_try_forcing_curl_vec(sizeof(T) == 4 ? SOPHT_F32 : SOPHT_F64, out, field, p, st)) {
    SOPHT_CHECK_LAUNCH();
    return SOPHT_OK;
  }
  curl_kernel<T, MODE><<<g.grid, g.block, 0, st>>>(
      comp3<T>(out, 0), comp3<T>(out, 1), comp3<T>(out, 2), cview(comp3<T>(field, 0)),
      cview(comp3<T>(field, 1)), cview(comp3<T>(field, 2)), (T)p, nz, ny, nx, reset);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

}  // namespace sopht

using namespace sopht;

#define RETURN_IF(rc_expr)   \
  do {                       \
    int rc__ = (rc_expr);    \
    if (rc__) return rc__;   \
  } while (0)

extern "C" {

int sopht_diffusion_flux_3d(int dtype, const sopht_field_t* diffusion_flux,
                            const sopht_field_t* field, double prefactor, int reset_ghost_zone,
                            void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(field, 3, 4) || !valid_field(diffusion_flux, 3, 4) ||
      !same_shape(field, diffusion_flux) || (field->ndim == 4 && field->shape[0] != 3) ||
      !fits_int(field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected matching (nz,ny,nx) or (3,nz,ny,nx) fields", __func__);
  return dtype == SOPHT_F32 ? diffusion_flux_impl<float>(diffusion_flux, field, prefactor,
                                                         reset_ghost_zone, as_stream(stream))
                            : diffusion_flux_impl<double>(diffusion_flux, field, prefactor,
                                                          reset_ghost_zone, as_stream(stream));
}

int sopht_curl_3d(int dtype, const sopht_field_t* curl, const sopht_field_t* field,
                  double prefactor, int reset_ghost_zone, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vector3(__func__, curl));
  RETURN_IF(check_vector3(__func__, field));
  if (!same_shape(curl, field) || !fits_int(field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: curl and field shapes differ", __func__);
  return dtype == SOPHT_F32
             ? curl_impl<float, 0>(curl, field, prefactor, reset_ghost_zone, as_stream(stream))
             : curl_impl<double, 0>(curl, field, prefactor, reset_ghost_zone, as_stream(stream));
}

int sopht_update_vorticity_from_velocity_forcing_3d(int dtype, const sopht_field_t* vorticity_field,
                                                    const sopht_field_t* velocity_forcing_field,
                                                    double prefactor, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vector3(__func__, vorticity_field));
  RETURN_IF(check_vector3(__func__, velocity_forcing_field));
  if (!same_shape(vorticity_field, velocity_forcing_field) || !fits_int(vorticity_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", __func__);
  return dtype == SOPHT_F32 ? curl_impl<float, 1>(vorticity_field, velocity_forcing_field, prefactor,
                                                  0, as_stream(stream))
                            : curl_impl<double, 1>(vorticity_field, velocity_forcing_field,
                                                   prefactor, 0, as_stream(stream));
}

int sopht_update_vorticity_from_penalised_velocity_3d(int dtype,
                                                      const sopht_field_t* vorticity_field,
                                                      const sopht_field_t* penalised_velocity_field,
                                                      const sopht_field_t* velocity_field,
                                                      double prefactor, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vector3(__func__, vorticity_field));
  RETURN_IF(check_vector3(__func__, penalised_velocity_field));
  RETURN_IF(check_vector3(__func__, velocity_field));
  if (!same_shape(vorticity_field, penalised_velocity_field) ||
      !same_shape(vorticity_field, velocity_field) || !fits_int(vorticity_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", __func__);
  GRID_DIMS(vorticity_field)
  cudaStream_t st = as_stream(stream);
#define LAUNCH_PV(T)                                                                              \
  penalised_velocity_update_kernel<T><<<g.grid, g.block, 0, st>>>(                                \
      comp3<T>(vorticity_field, 0), comp3<T>(vorticity_field, 1), comp3<T>(vorticity_field, 2),   \
      cview(comp3<T>(penalised_velocity_field, 0)), cview(comp3<T>(penalised_velocity_field, 1)), \
      cview(comp3<T>(penalised_velocity_field, 2)), cview(comp3<T>(velocity_field, 0)),           \
      cview(comp3<T>(velocity_field, 1)), cview(comp3<T>(velocity_field, 2)), (T)prefactor, nz,   \
      ny, nx)
  if (dtype == SOPHT_F32)
    LAUNCH_PV(float);
  else
    LAUNCH_PV(double);
#undef LAUNCH_PV
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_divergence_3d(int dtype, const sopht_field_t* divergence, const sopht_field_t* field,
                        double inv_dx, int reset_ghost_zone, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_scalar3(__func__, divergence));
  RETURN_IF(check_vector3(__func__, field));
  if (!same_grid(divergence, field) || !fits_int(field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid shapes differ", __func__);
  GRID_DIMS(field)
  cudaStream_t st = as_stream(stream);
  if (dtype == SOPHT_F32)
    divergence_kernel<float><<<g.grid, g.block, 0, st>>>(
        scalar3<float>(divergence), cview(comp3<float>(field, 0)), cview(comp3<float>(field, 1)),
        cview(comp3<float>(field, 2)), (float)inv_dx, nz, ny, nx, reset_ghost_zone);
  else
    divergence_kernel<double><<<g.grid, g.block, 0, st>>>(
        scalar3<double>(divergence), cview(comp3<double>(field, 0)), cview(comp3<double>(field, 1)),
        cview(comp3<double>(field, 2)), inv_dx, nz, ny, nx, reset_ghost_zone);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_vorticity_stretching_flux_3d(int dtype, const sopht_field_t* flux_field,
                                       const sopht_field_t* vorticity_field,
                                       const sopht_field_t* velocity_field, double prefactor,
                                       void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vector3(__func__, flux_field));
  RETURN_IF(check_vector3(__func__, vorticity_field));
  RETURN_IF(check_vector3(__func__, velocity_field));
  if (!same_shape(flux_field, vorticity_field) || !same_shape(flux_field, velocity_field) ||
      !fits_int(flux_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", __func__);
  GRID_DIMS(flux_field)
  cudaStream_t st = as_stream(stream);
#define LAUNCH_VS(T)                                                                           \
  stretching_flux_kernel<T><<<g.grid, g.block, 0, st>>>(                                       \
      comp3<T>(flux_field, 0), comp3<T>(flux_field, 1), comp3<T>(flux_field, 2),               \
      cview(comp3<T>(vorticity_field, 0)), cview(comp3<T>(vorticity_field, 1)),                \
      cview(comp3<T>(vorticity_field, 2)), cview(comp3<T>(velocity_field, 0)),                 \
      cview(comp3<T>(velocity_field, 1)), cview(comp3<T>(velocity_field, 2)), (T)prefactor, nz, \
      ny, nx)
  if (dtype == SOPHT_F32)
    LAUNCH_VS(float);
  else
    LAUNCH_VS(double);
#undef LAUNCH_VS
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_advection_flux_eno3_3d(int dtype, const sopht_field_t* advection_flux,
                                 const sopht_field_t* field, const sopht_field_t* velocity,
                                 double inv_dx, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_scalar3(__func__, advection_flux));
  RETURN_IF(check_scalar3(__func__, field));
  RETURN_IF(check_vector3(__func__, velocity));
  if (!same_shape(advection_flux, field) || !same_grid(field, velocity) || !fits_int(field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid shapes differ", __func__);
  GRID_DIMS(field)
  cudaStream_t st = as_stream(stream);
  if (dtype == SOPHT_F32)
    advection_flux_eno3_kernel<float><<<g.grid, g.block, 0, st>>>(
        scalar3<float>(advection_flux), cview(scalar3<float>(field)),
        cview(comp3<float>(velocity, 0)), cview(comp3<float>(velocity, 1)),
        cview(comp3<float>(velocity, 2)), (float)inv_dx, nz, ny, nx);
  else
    advection_flux_eno3_kernel<double><<<g.grid, g.block, 0, st>>>(
        scalar3<double>(advection_flux), cview(scalar3<double>(field)),
        cview(comp3<double>(velocity, 0)), cview(comp3<double>(velocity, 1)),
        cview(comp3<double>(velocity, 2)), inv_dx, nz, ny, nx);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_laplacian_filter_flux_3d(int dtype, const sopht_field_t* filter_flux,
                                   const sopht_field_t* field, int axis, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_scalar3(__func__, filter_flux));
  RETURN_IF(check_scalar3(__func__, field));
  if (axis < 0 || axis > 2) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: axis must be 0 (x), 1 (y) or 2 (z)", __func__);
  if (!same_shape(filter_flux, field) || !fits_int(field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", __func__);
  GRID_DIMS(field)
  const int di = axis == 0, dj = axis == 1, dk = axis == 2;
  cudaStream_t st = as_stream(stream);
  if (dtype == SOPHT_F32)
    laplacian_filter_flux_kernel<float><<<g.grid, g.block, 0, st>>>(
        scalar3<float>(filter_flux), cview(scalar3<float>(field)), dk, dj, di, nz, ny, nx);
  else
    laplacian_filter_flux_kernel<double><<<g.grid, g.block, 0, st>>>(
        scalar3<double>(filter_flux), cview(scalar3<double>(field)), dk, dj, di, nz, ny, nx);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

static int penalise3d_entry(const char* fn, int dtype, const sopht_field_t* field, int width,
                            const double* ramp_x, const double* ramp_y, const double* ramp_z, int z_faces,
                            void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (width < 0 || width > 8) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: width must be in [0, 8]", fn);
  if (z_faces < 0 || z_faces > 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: z_faces must be in [0, 3]", fn);
  if (width == 0) return SOPHT_OK;
  if (!valid_field(field, 3, 4) || (field->ndim == 4 && field->shape[0] != 3) || !fits_int(field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected (nz,ny,nx) or (3,nz,ny,nx)", fn);
  if (!ramp_x || !ramp_y || !ramp_z) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null ramp table", fn);
  const int nz = (int)field->shape[field->ndim - 3];
  const int ny = (int)field->shape[field->ndim - 2];
  const int nx = (int)field->shape[field->ndim - 1];
  const int zneed = ((z_faces & 1) ? width : 0) + ((z_faces & 2) ? width : 0);
  if (nz < zneed || nz < 1 || ny < 2 * width || nx < 2 * width)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid smaller than twice the penalisation width", fn);
  cudaStream_t st = as_stream(stream);
  const int ncomp = field->ndim == 4 ? 3 : 1;
  if (dtype == SOPHT_F32) {
    ShellViews<float> v;
    for (int c = 0; c < 3; ++c)
      v.f[c] = field->ndim == 4 ? comp3<float>(field, c) : scalar3<float>(field);
    return penalise3d_impl<float>(v, ncomp, nz, ny, nx, width, ramp_x, ramp_y, ramp_z, z_faces, st);
  }
  ShellViews<double> v;
  for (int c = 0; c < 3; ++c) v.f[c] = field->ndim == 4 ? comp3<double>(field, c) : scalar3<double>(field);
  return penalise3d_impl<double>(v, ncomp, nz, ny, nx, width, ramp_x, ramp_y, ramp_z, z_faces, st);
}

int sopht_penalise_field_boundary_3d(int dtype, const sopht_field_t* field, int width,
                                     const double* ramp_x, const double* ramp_y,
                                     const double* ramp_z, void* stream) {
  return penalise3d_entry(__func__, dtype, field, width, ramp_x, ramp_y, ramp_z, 3, stream);
}

int sopht_penalise_field_boundary_3d_slab(int dtype, const sopht_field_t* field, int width,
                                          const double* ramp_x, const double* ramp_y,
                                          const double* ramp_z, int z_faces, void* stream) {
  return penalise3d_entry(__func__, dtype, field, width, ramp_x, ramp_y, ramp_z, z_faces, stream);
}

}  // extern "C"
