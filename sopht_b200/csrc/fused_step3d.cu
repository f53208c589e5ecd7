// Fused stencil passes of the 3-D Navier-Stokes step (the simulator-level path; the one-op-per-kernel
// versions behind the public factories live in stencils3d.cu / elementwise.cu).
//
//   advect   : out = w + p * curl_c(u x w)            (cross product recomputed on the fly at the 6 neighbours,
//              the reference's buffer_vector_field round trip disappears)   24 B read + 12 B written per cell
//   diffuse  : out = f + q * Lap_7pt(f), optionally forcing_field <- 0 in the same pass   4+4 (+4) B per cell/comp
//   velocity : u = p * curl_c(psi) (ring <- 0) + U_inf, max_cells sum_c |u_c| reduced on the device  12 + 12 B
//
// All three are z-marching kernels: a CTA owns a TX x TY patch of (x, y) columns and walks a chunk of z
// planes. The raw fields of the incoming plane (patch + 1-cell halo) are staged in a double-buffered
// shared-memory tile with coalesced loads; x/y neighbours come from that tile, z neighbours of a thread's own
// column ride in registers. Ghost-ring rule of the reference's generated kernels: only cells with all
// indices in [1, n-2] are updated, ring cells keep (advect, diffuse) or zero (velocity) their value.
//
// ref: sopht/simulator/flow/navier_stokes_flow_simulators.py:449-498 (step order),
//      stencil_ops_3d/{elementwise_ops_3d.py:390-449, update_vorticity_from_velocity_forcing_3d.py:12-132,
//      diffusion_timestep_3d.py:12-80, curl_3d.py:13-132}, passive_transport_flow_simulators.py:139-155
#include <stdlib.h>

#include "common.cuh"
#include "stream_vec.cuh"

namespace sopht {

namespace {

// output patch per CTA (= threads), staged tile = patch + 1-cell halo; fp64 halves the x extent to stay
// inside the 48 KB static shared-memory window
template <typename T>
struct Tile {
  static constexpr int TX = sizeof(T) == 4 ? 64 : 32, TY = 8;
  static constexpr int SX = TX + 2, SY = TY + 2, THREADS = TX * TY;
};
#define FTX (Tile<T>::TX)
#define FTY (Tile<T>::TY)
#define FSX (Tile<T>::SX)
#define FSY (Tile<T>::SY)
#define FTHREADS (Tile<T>::THREADS)

template <typename T>
struct Vec3View {
  const T* p[3];
  int64_t sz, sy;  // x stride is 1 (checked on the host)
};
template <typename T>
struct Vec3Out {
  T* p[3];
  int64_t sz, sy;
};

// stage plane k of NF fields (patch + halo, zero outside the grid) into s[NF][FSY][FSX]
template <typename T, int NF>
__device__ __forceinline__ void load_plane(T (*s)[FSY][FSX], const T* const* f, int64_t sz, int64_t sy,
                                           int k, int x0, int y0, int ny, int nx) {
  const int tid = threadIdx.y * FTX + threadIdx.x;
#pragma unroll 1
  for (int q = tid; q < FSX * FSY; q += FTHREADS) {
    const int ly = q / FSX, lx = q - ly * FSX;
    const int gy = y0 - 1 + ly, gx = x0 - 1 + lx;
    const bool in = gy >= 0 && gy < ny && gx >= 0 && gx < nx;
    const int64_t o = (int64_t)k * sz + (int64_t)gy * sy + gx;
#pragma unroll
    for (int c = 0; c < NF; ++c) s[c][ly][lx] = in ? __ldg(f[c] + o) : T(0);
  }
}

template <typename T>
struct AtomicMaxNonNeg;  // max of non-negative floats through their (order-preserving) bit patterns
template <>
struct AtomicMaxNonNeg<float> {
  __device__ static void apply(float* a, float v) { atomicMax(reinterpret_cast<int*>(a), __float_as_int(v)); }
};
template <>
struct AtomicMaxNonNeg<double> {
  __device__ static void apply(double* a, double v) {
    atomicMax(reinterpret_cast<long long*>(a), __double_as_longlong(v));
  }
};

// ---- advect: out = w + p * curl_c(u x w) --------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(FTHREADS, 2)
    advect_rotational_kernel(Vec3Out<T> out, Vec3View<T> w, Vec3View<T> u, T p, int nz, int ny, int nx,
                             int kchunk) {
  __shared__ T s[2][6][FSY][FSX];  // fields: 0..2 = u, 3..5 = w
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY;
  const int i = x0 + tx, j = y0 + ty;
  const int k0 = blockIdx.z * kchunk, k1 = min(k0 + kchunk, nz);
  const T* f[6] = {u.p[0], u.p[1], u.p[2], w.p[0], w.p[1], w.p[2]};
  const bool own = i < nx && j < ny;
  const bool inner_xy = i >= 1 && i < nx - 1 && j >= 1 && j < ny - 1;
  const int cy = ty + 1, cx = tx + 1;

  T bprev_x = T(0), bprev_y = T(0);
  if (k0 > 0 && own) {
    const int64_t o = (int64_t)(k0 - 1) * w.sz + (int64_t)j * w.sy + i;
    const T ux = __ldg(u.p[0] + o), uy = __ldg(u.p[1] + o), uz = __ldg(u.p[2] + o);
    const T wx = __ldg(w.p[0] + o), wy = __ldg(w.p[1] + o), wz = __ldg(w.p[2] + o);
    bprev_x = uy * wz - wy * uz;
    bprev_y = uz * wx - wz * ux;
  }
  int cur = 0;
  load_plane<T, 6>(s[cur], f, w.sz, w.sy, k0, x0, y0, ny, nx);
  __syncthreads();
  T bcur_x, bcur_y;
  {
    const T ux = s[cur][0][cy][cx], uy = s[cur][1][cy][cx], uz = s[cur][2][cy][cx];
    const T wx = s[cur][3][cy][cx], wy = s[cur][4][cy][cx], wz = s[cur][5][cy][cx];
    bcur_x = uy * wz - wy * uz;
    bcur_y = uz * wx - wz * ux;
  }
  for (int k = k0; k < k1; ++k) {
    const int nxt = cur ^ 1;
    T bnext_x = T(0), bnext_y = T(0);
    if (k + 1 < nz) {
      load_plane<T, 6>(s[nxt], f, w.sz, w.sy, k + 1, x0, y0, ny, nx);
      __syncthreads();
      const T ux = s[nxt][0][cy][cx], uy = s[nxt][1][cy][cx], uz = s[nxt][2][cy][cx];
      const T wx = s[nxt][3][cy][cx], wy = s[nxt][4][cy][cx], wz = s[nxt][5][cy][cx];
      bnext_x = uy * wz - wy * uz;
      bnext_y = uz * wx - wz * ux;
    }
    if (own) {
      T ox = s[cur][3][cy][cx], oy = s[cur][4][cy][cx], oz = s[cur][5][cy][cx];
      if (inner_xy && k >= 1 && k < nz - 1) {
        // b = u x w at the four in-plane neighbours (only the components the curl needs)
#define U_(c, yy, xx) s[cur][c][yy][xx]
#define W_(c, yy, xx) s[cur][3 + c][yy][xx]
#define BX_(yy, xx) (U_(1, yy, xx) * W_(2, yy, xx) - W_(1, yy, xx) * U_(2, yy, xx))
#define BY_(yy, xx) (U_(2, yy, xx) * W_(0, yy, xx) - W_(2, yy, xx) * U_(0, yy, xx))
#define BZ_(yy, xx) (U_(0, yy, xx) * W_(1, yy, xx) - W_(0, yy, xx) * U_(1, yy, xx))
        const T bz_jp = BZ_(cy + 1, cx), bz_jm = BZ_(cy - 1, cx);
        const T bz_ip = BZ_(cy, cx + 1), bz_im = BZ_(cy, cx - 1);
        const T by_ip = BY_(cy, cx + 1), by_im = BY_(cy, cx - 1);
        const T bx_jp = BX_(cy + 1, cx), bx_jm = BX_(cy - 1, cx);
#undef U_
#undef W_
#undef BX_
#undef BY_
#undef BZ_
        const T ccx = bz_jp - bz_jm - bnext_y + bprev_y;
        const T ccy = bnext_x - bprev_x - bz_ip + bz_im;
        const T ccz = by_ip - by_im - bx_jp + bx_jm;
        ox = ox + p * ccx;
        oy = oy + p * ccy;
        oz = oz + p * ccz;
      }
      const int64_t o = (int64_t)k * out.sz + (int64_t)j * out.sy + i;
      out.p[0][o] = ox;
      out.p[1][o] = oy;
      out.p[2][o] = oz;
    }
    bprev_x = bcur_x, bprev_y = bcur_y;
    bcur_x = bnext_x, bcur_y = bnext_y;
    cur = nxt;
    __syncthreads();
  }
}

// ---- diffuse: out = f + q * Lap(f); blockIdx.z = chunk * ncomp + comp ------------------------------------
template <typename T>
__global__ void __launch_bounds__(FTHREADS, 2)
    diffuse_kernel(Vec3Out<T> out, Vec3View<T> in, Vec3Out<T> zero, int has_zero, int ncomp, T q, int nz,
                   int ny, int nx, int kchunk) {
  __shared__ T s[2][1][FSY][FSX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY;
  const int i = x0 + tx, j = y0 + ty;
  const int comp = blockIdx.z % ncomp, chunk = blockIdx.z / ncomp;
  const int k0 = chunk * kchunk, k1 = min(k0 + kchunk, nz);
  const T* f[1] = {comp == 0 ? in.p[0] : (comp == 1 ? in.p[1] : in.p[2])};
  T* o_p = comp == 0 ? out.p[0] : (comp == 1 ? out.p[1] : out.p[2]);
  T* z_p = comp == 0 ? zero.p[0] : (comp == 1 ? zero.p[1] : zero.p[2]);
  const bool own = i < nx && j < ny;
  const bool inner_xy = i >= 1 && i < nx - 1 && j >= 1 && j < ny - 1;
  const int cy = ty + 1, cx = tx + 1;
  const int64_t col = (int64_t)j * in.sy + i;

  T fprev = (k0 > 0 && own) ? __ldg(f[0] + (int64_t)(k0 - 1) * in.sz + col) : T(0);
  int cur = 0;
  load_plane<T, 1>(s[cur], f, in.sz, in.sy, k0, x0, y0, ny, nx);
  __syncthreads();
  T fcur = s[cur][0][cy][cx];
  for (int k = k0; k < k1; ++k) {
    const int nxt = cur ^ 1;
    T fnext = T(0);
    if (k + 1 < nz) {
      load_plane<T, 1>(s[nxt], f, in.sz, in.sy, k + 1, x0, y0, ny, nx);
      __syncthreads();
      fnext = s[nxt][0][cy][cx];
    }
    if (own) {
      T v = fcur;
      if (inner_xy && k >= 1 && k < nz - 1) {
        const T flux = q * (fnext + fprev + s[cur][0][cy + 1][cx] + s[cur][0][cy - 1][cx] +
                            s[cur][0][cy][cx + 1] + s[cur][0][cy][cx - 1] - T(6) * fcur);
        v = fcur + flux;
      }
      o_p[(int64_t)k * out.sz + (int64_t)j * out.sy + i] = v;
      if (has_zero) z_p[(int64_t)k * zero.sz + (int64_t)j * zero.sy + i] = T(0);
    }
    fprev = fcur;
    fcur = fnext;
    cur = nxt;
    __syncthreads();
  }
}

// ---- velocity: u = p * curl_c(psi) on the interior, 0 on the ring, + U_inf; max sum_c |u_c| --------------
template <typename T>
__global__ void __launch_bounds__(FTHREADS, 2)
    velocity_from_psi_kernel(Vec3Out<T> out, Vec3View<T> psi, T p, T fx, T fy, T fz, T* max_out, int nz,
                             int ny, int nx, int kchunk) {
  __shared__ T s[2][3][FSY][FSX];
  __shared__ T wmax[FTHREADS / 32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY;
  const int i = x0 + tx, j = y0 + ty;
  const int k0 = blockIdx.z * kchunk, k1 = min(k0 + kchunk, nz);
  const T* f[3] = {psi.p[0], psi.p[1], psi.p[2]};
  const bool own = i < nx && j < ny;
  const bool inner_xy = i >= 1 && i < nx - 1 && j >= 1 && j < ny - 1;
  const int cy = ty + 1, cx = tx + 1;
  const int64_t col = (int64_t)j * psi.sy + i;

  T pprev_x = T(0), pprev_y = T(0);
  if (k0 > 0 && own) {
    pprev_x = __ldg(f[0] + (int64_t)(k0 - 1) * psi.sz + col);
    pprev_y = __ldg(f[1] + (int64_t)(k0 - 1) * psi.sz + col);
  }
  int cur = 0;
  load_plane<T, 3>(s[cur], f, psi.sz, psi.sy, k0, x0, y0, ny, nx);
  __syncthreads();
  T pcur_x = s[cur][0][cy][cx], pcur_y = s[cur][1][cy][cx];
  T m = T(0);
  for (int k = k0; k < k1; ++k) {
    const int nxt = cur ^ 1;
    T pnext_x = T(0), pnext_y = T(0);
    if (k + 1 < nz) {
      load_plane<T, 3>(s[nxt], f, psi.sz, psi.sy, k + 1, x0, y0, ny, nx);
      __syncthreads();
      pnext_x = s[nxt][0][cy][cx];
      pnext_y = s[nxt][1][cy][cx];
    }
    if (own) {
      T ux = T(0), uy = T(0), uz = T(0);
      if (inner_xy && k >= 1 && k < nz - 1) {
        const T ccx = s[cur][2][cy + 1][cx] - s[cur][2][cy - 1][cx] - pnext_y + pprev_y;
        const T ccy = pnext_x - pprev_x - s[cur][2][cy][cx + 1] + s[cur][2][cy][cx - 1];
        const T ccz = s[cur][1][cy][cx + 1] - s[cur][1][cy][cx - 1] - s[cur][0][cy + 1][cx] + s[cur][0][cy - 1][cx];
        ux = p * ccx;
        uy = p * ccy;
        uz = p * ccz;
      }
      ux = ux + fx;
      uy = uy + fy;
      uz = uz + fz;
      const int64_t o = (int64_t)k * out.sz + (int64_t)j * out.sy + i;
      out.p[0][o] = ux;
      out.p[1][o] = uy;
      out.p[2][o] = uz;
      const T a = fabs(ux) + fabs(uy) + fabs(uz);
      m = nanmax(a, m);
    }
    pprev_x = pcur_x, pprev_y = pcur_y;
    pcur_x = pnext_x, pcur_y = pnext_y;
    cur = nxt;
    __syncthreads();
  }
  if (max_out) {
    for (int off = 16; off > 0; off >>= 1) {
      const T o = __shfl_xor_sync(0xffffffffu, m, off);
      m = nanmax(o, m);
    }
    const int tid = ty * FTX + tx;
    if ((tid & 31) == 0) wmax[tid >> 5] = m;
    __syncthreads();
    if (tid < 32) {
      m = tid < FTHREADS / 32 ? wmax[tid] : T(0);
      for (int off = 16; off > 0; off >>= 1) {
        const T o = __shfl_xor_sync(0xffffffffu, m, off);
        m = nanmax(o, m);
      }
      if (tid == 0) AtomicMaxNonNeg<T>::apply(max_out, m);
    }
  }
}

// ======================================================================================================
// Register-marching versions (stream_vec.cuh): 16-byte vectors along x, no shared memory, no barriers.
// Taken whenever nx, the row/plane strides and the base addresses are multiples of the vector width;
// the shared-memory kernels above remain for every other view.
// ======================================================================================================
using sv::Vec;

constexpr int VBY = 8;  // rows of a CTA: block = (32 lanes, VBY)

template <typename T>
__device__ __forceinline__ Vec<T> cross_x(const Vec<T>* u, const Vec<T>* w) {
  Vec<T> r;
#pragma unroll
  for (int m = 0; m < Vec<T>::W; ++m) r.v[m] = u[1].v[m] * w[2].v[m] - w[1].v[m] * u[2].v[m];
  return r;
}
template <typename T>
__device__ __forceinline__ Vec<T> cross_y(const Vec<T>* u, const Vec<T>* w) {
  Vec<T> r;
#pragma unroll
  for (int m = 0; m < Vec<T>::W; ++m) r.v[m] = u[2].v[m] * w[0].v[m] - w[2].v[m] * u[0].v[m];
  return r;
}
template <typename T>
__device__ __forceinline__ Vec<T> cross_z(const Vec<T>* u, const Vec<T>* w) {
  Vec<T> r;
#pragma unroll
  for (int m = 0; m < Vec<T>::W; ++m) r.v[m] = u[0].v[m] * w[1].v[m] - w[0].v[m] * u[1].v[m];
  return r;
}

// out = w + p * curl_c(u x w). 24 B read + 12 B written per cell (fp32).
template <typename T, bool PXY>
__global__ void __launch_bounds__(32 * VBY, 2)
    advect_vec_kernel(Vec3Out<T> out, Vec3View<T> w, Vec3View<T> u, T p, int nz, int ny, int nx, int kchunk) {
  constexpr int W = Vec<T>::W;
  const int lane = threadIdx.x;
  const int i0 = (blockIdx.x * 32 + lane) * W;
  const int j = blockIdx.y * VBY + threadIdx.y;
  if (j >= ny) return;  // warp-uniform
  const int k0 = blockIdx.z * kchunk, k1 = min(k0 + kchunk, nz);
  const bool act = i0 < nx;
  const sv::Nbr<PXY> nb(act, lane, i0, W, j, ny, nx, w.sy);
  const bool jin = nb.jin, has_l = nb.has_l, has_r = nb.has_r;
  const int64_t col = (int64_t)j * w.sy + i0;  // u shares sz / sy with w (checked on the host)
  const int64_t ocol = (int64_t)j * out.sy + i0;

  Vec<T> uc[3], wc[3];
  Vec<T> bprev_x = sv::vzero<T>(), bprev_y = sv::vzero<T>();
  if (k0 > 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uc[c] = sv::vload_if(act, u.p[c] + (int64_t)(k0 - 1) * w.sz + col);
      wc[c] = sv::vload_if(act, w.p[c] + (int64_t)(k0 - 1) * w.sz + col);
    }
    bprev_x = cross_x(uc, wc);
    bprev_y = cross_y(uc, wc);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uc[c] = sv::vload_if(act, u.p[c] + (int64_t)k0 * w.sz + col);
    wc[c] = sv::vload_if(act, w.p[c] + (int64_t)k0 * w.sz + col);
  }
  Vec<T> bcur_x = cross_x(uc, wc), bcur_y = cross_y(uc, wc), bcur_z = cross_z(uc, wc);

#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    const int64_t pl = (int64_t)k * w.sz + col;
    // issue every load of this iteration first: next plane (centre), rows j+1 / j-1 and the two x edges
    Vec<T> un[3], wn[3], uu[3], wu[3], ud[3], wd[3];
    T ul[3], wl[3], ur[3], wr[3];
    const bool kn = act && k + 1 < nz, up = nb.has_up, dn = nb.has_dn;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      un[c] = sv::vload_if(kn, u.p[c] + pl + w.sz);
      wn[c] = sv::vload_if(kn, w.p[c] + pl + w.sz);
      uu[c] = sv::vload_if(up, u.p[c] + pl + nb.up);
      wu[c] = sv::vload_if(up, w.p[c] + pl + nb.up);
      ud[c] = sv::vload_if(dn, u.p[c] + pl + nb.dn);
      wd[c] = sv::vload_if(dn, w.p[c] + pl + nb.dn);
      ul[c] = sv::sload_if(has_l, u.p[c] + pl + nb.left);
      wl[c] = sv::sload_if(has_l, w.p[c] + pl + nb.left);
      ur[c] = sv::sload_if(has_r, u.p[c] + pl + nb.right);
      wr[c] = sv::sload_if(has_r, w.p[c] + pl + nb.right);
    }
    const Vec<T> bnext_x = cross_x(un, wn), bnext_y = cross_y(un, wn), bnext_z = cross_z(un, wn);
    const Vec<T> bz_jp = cross_z(uu, wu), bx_jp = cross_x(uu, wu);
    const Vec<T> bz_jm = cross_z(ud, wd), bx_jm = cross_x(ud, wd);
    const T ebz_l = ul[0] * wl[1] - wl[0] * ul[1], eby_l = ul[2] * wl[0] - wl[2] * ul[0];
    const T ebz_r = ur[0] * wr[1] - wr[0] * ur[1], eby_r = ur[2] * wr[0] - wr[2] * ur[0];
    Vec<T> bz_im, bz_ip, by_im, by_ip;
    sv::x_neighbours(bcur_z, ebz_l, ebz_r, nb.use_l(lane), nb.use_r(lane), bz_im, bz_ip);
    sv::x_neighbours(bcur_y, eby_l, eby_r, nb.use_l(lane), nb.use_r(lane), by_im, by_ip);
    if (act) {
      Vec<T> ox = wc[0], oy = wc[1], oz = wc[2];
      if (jin && k >= 1 && k < nz - 1) {
#pragma unroll
        for (int m = 0; m < W; ++m) {
          const int i = i0 + m;
          if (nb.iin(i, nx)) {
            const T ccx = bz_jp.v[m] - bz_jm.v[m] - bnext_y.v[m] + bprev_y.v[m];
            const T ccy = bnext_x.v[m] - bprev_x.v[m] - bz_ip.v[m] + bz_im.v[m];
            const T ccz = by_ip.v[m] - by_im.v[m] - bx_jp.v[m] + bx_jm.v[m];
            ox.v[m] = ox.v[m] + p * ccx;
            oy.v[m] = oy.v[m] + p * ccy;
            oz.v[m] = oz.v[m] + p * ccz;
          }
        }
      }
      const int64_t o = (int64_t)k * out.sz + ocol;
      sv::vstore(out.p[0] + o, ox);
      sv::vstore(out.p[1] + o, oy);
      sv::vstore(out.p[2] + o, oz);
    }
    bprev_x = bcur_x, bprev_y = bcur_y;
    bcur_x = bnext_x, bcur_y = bnext_y, bcur_z = bnext_z;
#pragma unroll
    for (int c = 0; c < 3; ++c) wc[c] = wn[c];
  }
}

// out = f + q * Lap_7pt(f) [, zero <- 0]; blockIdx.z = chunk * ncomp + comp. 4 + 4 (+4) B per cell (fp32).
// RAMP: the result is multiplied by rx[i] * ry[j] * rz[k] (in that order) on the way out - the sine penalisation of the
// boundary ring for the default width 2, where the reference's copy-and-scale (penalise_field_boundary_3d.py:182-208)
// degenerates to a per-axis factor {0, sin(pi/4), 1, ..., 1, sin(pi/4), 0}: the boundary cell takes the value of its
// neighbour times sin(0) = 0, the neighbour is scaled in place.
template <typename T, bool PXY, bool RAMP = false>
__global__ void __launch_bounds__(32 * VBY, sizeof(T) == 4 ? 4 : 1)  // fp32: 64 registers, four CTAs per SM
    diffuse_vec_kernel(Vec3Out<T> out, Vec3View<T> in, Vec3Out<T> zero, int has_zero, int ncomp, T q, int nz,
                       int ny, int nx, int kchunk, const T* __restrict__ rx = nullptr,
                       const T* __restrict__ ry = nullptr, const T* __restrict__ rz = nullptr) {
  constexpr int W = Vec<T>::W;
  const int lane = threadIdx.x;
  const int i0 = (blockIdx.x * 32 + lane) * W;
  const int j = blockIdx.y * VBY + threadIdx.y;
  if (j >= ny) return;
  const int comp = blockIdx.z % ncomp, chunk = blockIdx.z / ncomp;
  const int k0 = chunk * kchunk, k1 = min(k0 + kchunk, nz);
  const T* f = comp == 0 ? in.p[0] : (comp == 1 ? in.p[1] : in.p[2]);
  T* o_p = comp == 0 ? out.p[0] : (comp == 1 ? out.p[1] : out.p[2]);
  T* z_p = comp == 0 ? zero.p[0] : (comp == 1 ? zero.p[1] : zero.p[2]);
  const bool act = i0 < nx;
  const sv::Nbr<PXY> nb(act, lane, i0, W, j, ny, nx, in.sy);
  const bool jin = nb.jin, has_l = nb.has_l, has_r = nb.has_r, up = nb.has_up, dn = nb.has_dn;
  f += (int64_t)j * in.sy + i0;
  o_p += (int64_t)j * out.sy + i0;
  z_p += (int64_t)j * zero.sy + i0;

  Vec<T> fprev = sv::vload_if(act && k0 > 0, f + (int64_t)(k0 - 1) * in.sz);
  Vec<T> fcur = sv::vload_if(act, f + (int64_t)k0 * in.sz);
  Vec<T> rampx = sv::vzero<T>();
  T rampy = T(1);
  if (RAMP) {
    if (act) rampx = sv::vload(rx + i0);
    rampy = ry[j];
  }
#pragma unroll 2
  for (int k = k0; k < k1; ++k) {
    const T* pk = f + (int64_t)k * in.sz;
    const T rk = RAMP ? __ldg(rz + k) : T(1);  // issued with the plane's other loads
    const Vec<T> fnext = sv::vload_if(act && k + 1 < nz, pk + in.sz);
    const Vec<T> fup = sv::vload_if(up, pk + nb.up);
    const Vec<T> fdn = sv::vload_if(dn, pk + nb.dn);
    const T el = sv::sload_if(has_l, pk + nb.left), er = sv::sload_if(has_r, pk + nb.right);
    Vec<T> xl, xr;
    sv::x_neighbours(fcur, el, er, nb.use_l(lane), nb.use_r(lane), xl, xr);
    if (act) {
      Vec<T> v = fcur;
      if (jin && k >= 1 && k < nz - 1) {
#pragma unroll
        for (int m = 0; m < W; ++m) {
          const int i = i0 + m;
          if (nb.iin(i, nx)) {
            const T flux =
                q * (fnext.v[m] + fprev.v[m] + fup.v[m] + fdn.v[m] + xr.v[m] + xl.v[m] - T(6) * fcur.v[m]);
            v.v[m] = fcur.v[m] + flux;
          }
        }
      }
      if (RAMP) {
#pragma unroll
        for (int m = 0; m < W; ++m) v.v[m] = ((v.v[m] * rampx.v[m]) * rampy) * rk;
      }
      sv::vstore(o_p + (int64_t)k * out.sz, v);
      if (has_zero) sv::vstore(z_p + (int64_t)k * zero.sz, sv::vzero<T>());
    }
    fprev = fcur;
    fcur = fnext;
  }
}

// The same pass with the three components of the vector field in ONE thread (fp32): three times the independent 16-byte
// loads in flight per thread and one set of neighbour predicates / addresses for the three - the shape of the curl
// kernel below, which moves the same bytes at 5.8 TB/s where the one-component form reaches 4.5 (ncu: 478 M warp
// instructions, issue 58 %, nine long-scoreboard stalls per issue). 512^3: 0.702 -> 0.643 ms. The periodic (wrap-around)
// variant measured 3 % slower in this form and keeps the one-component kernel.
template <typename T, bool PXY, bool RAMP>
__global__ void __launch_bounds__(32 * VBY, 3)  // 80 registers, a few spilled words; two CTAs (100 registers) are no faster than the old form
    diffuse_vec3_kernel(Vec3Out<T> out, Vec3View<T> in, Vec3Out<T> zero, int has_zero, T q, int nz, int ny, int nx,
                        int kchunk, const T* __restrict__ rx, const T* __restrict__ ry, const T* __restrict__ rz) {
  constexpr int W = Vec<T>::W;
  const int lane = threadIdx.x;
  const int i0 = (blockIdx.x * 32 + lane) * W;
  const int j = blockIdx.y * VBY + threadIdx.y;
  if (j >= ny) return;
  const int k0 = blockIdx.z * kchunk, k1 = min(k0 + kchunk, nz);
  const bool act = i0 < nx;
  const sv::Nbr<PXY> nb(act, lane, i0, W, j, ny, nx, in.sy);
  const bool jin = nb.jin, has_l = nb.has_l, has_r = nb.has_r, up = nb.has_up, dn = nb.has_dn;
  const int64_t ioff = (int64_t)j * in.sy + i0, ooff = (int64_t)j * out.sy + i0, zoff = (int64_t)j * zero.sy + i0;
  Vec<T> fprev[3], fcur[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    fprev[c] = sv::vload_if(act && k0 > 0, in.p[c] + ioff + (int64_t)(k0 - 1) * in.sz);
    fcur[c] = sv::vload_if(act, in.p[c] + ioff + (int64_t)k0 * in.sz);
  }
  Vec<T> rampx = sv::vzero<T>();
  T rampy = T(1);
  if (RAMP) {
    if (act) rampx = sv::vload(rx + i0);
    rampy = ry[j];
  }
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    const T rk = RAMP ? __ldg(rz + k) : T(1);  // issued with the plane's other loads
    const bool kin = k >= 1 && k < nz - 1;
    Vec<T> fnext[3], fup[3], fdn[3];
    T el[3], er[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // all loads of the plane first
      const T* pk = in.p[c] + ioff + (int64_t)k * in.sz;
      fnext[c] = sv::vload_if(act && k + 1 < nz, pk + in.sz);
      fup[c] = sv::vload_if(up, pk + nb.up);
      fdn[c] = sv::vload_if(dn, pk + nb.dn);
      el[c] = sv::sload_if(has_l, pk + nb.left), er[c] = sv::sload_if(has_r, pk + nb.right);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Vec<T> xl, xr;
      sv::x_neighbours(fcur[c], el[c], er[c], nb.use_l(lane), nb.use_r(lane), xl, xr);
      if (act) {
        Vec<T> v = fcur[c];
        if (jin && kin) {
#pragma unroll
          for (int m = 0; m < W; ++m) {
            if (nb.iin(i0 + m, nx)) {
              const T flux = q * (fnext[c].v[m] + fprev[c].v[m] + fup[c].v[m] + fdn[c].v[m] + xr.v[m] + xl.v[m] -
                                  T(6) * fcur[c].v[m]);
              v.v[m] = fcur[c].v[m] + flux;
            }
          }
        }
        if (RAMP) {
#pragma unroll
          for (int m = 0; m < W; ++m) v.v[m] = ((v.v[m] * rampx.v[m]) * rampy) * rk;
        }
        sv::vstore(out.p[c] + ooff + (int64_t)k * out.sz, v);
        if (has_zero) sv::vstore(zero.p[c] + zoff + (int64_t)k * zero.sz, sv::vzero<T>());
      }
      fprev[c] = fcur[c];
      fcur[c] = fnext[c];
    }
  }
}

// ACCUM = false: u = p * curl_c(psi) (ring <- 0) + U_inf; max_cells sum_c |u_c|. 12 B read + 12 B written per
//                cell (fp32).
// ACCUM = true : out += p * curl_c(psi) on the interior, ring cells untouched (the forcing update
//                w += p curl(f), update_vorticity_from_velocity_forcing_3d.py:12-132). 24 B read + 12 B written.
template <typename T, bool ACCUM, bool PXY = false>
// fp32 curl(psi): three CTAs per SM (80 registers, a few spilled words) instead of two (106): 0.637 -> 0.556 ms at 512^3
// (5.8 TB/s); the accumulating form and fp64 keep their register budget. Measured the same way: advect at three CTAs
// spills 300 bytes (0.91 -> 2.66 ms), diffuse at five / six CTAs loses 2 - 9 %.
__global__ void __launch_bounds__(32 * VBY, (sizeof(T) == 4 && !ACCUM) ? 3 : 1)
    velocity_vec_kernel(Vec3Out<T> out, Vec3View<T> psi, T p, T fx, T fy, T fz, T* max_out, int nz, int ny,
                        int nx, int kchunk) {
  constexpr int W = Vec<T>::W;
  __shared__ T wmax[VBY];
  const int lane = threadIdx.x;
  const int i0 = (blockIdx.x * 32 + lane) * W;
  const int j = blockIdx.y * VBY + threadIdx.y;
  const int k0 = blockIdx.z * kchunk, k1 = min(k0 + kchunk, nz);
  const bool row = j < ny;
  const bool act = row && i0 < nx;
  const sv::Nbr<PXY> nb(act, lane, i0, W, j, ny, nx, psi.sy);
  const bool jin = nb.jin, has_l = nb.has_l, has_r = nb.has_r, up = nb.has_up, dn = nb.has_dn;
  const int64_t col = (int64_t)j * psi.sy + i0;
  const int64_t ocol = (int64_t)j * out.sy + i0;
  T m_acc = T(0);

  // psi_x, psi_y are needed at k +- 1; psi_z only in-plane
  Vec<T> pxm = sv::vload_if(act && k0 > 0, psi.p[0] + (int64_t)(k0 - 1) * psi.sz + col);
  Vec<T> pym = sv::vload_if(act && k0 > 0, psi.p[1] + (int64_t)(k0 - 1) * psi.sz + col);
  Vec<T> pxc = sv::vload_if(act, psi.p[0] + (int64_t)k0 * psi.sz + col);
  Vec<T> pyc = sv::vload_if(act, psi.p[1] + (int64_t)k0 * psi.sz + col);
#pragma unroll 2
  for (int k = k0; k < k1; ++k) {
    const int64_t pl = (int64_t)k * psi.sz + col;
    const bool kn = act && k + 1 < nz;
    const Vec<T> pxn = sv::vload_if(kn, psi.p[0] + pl + psi.sz);
    const Vec<T> pyn = sv::vload_if(kn, psi.p[1] + pl + psi.sz);
    const Vec<T> pzc = sv::vload_if(act, psi.p[2] + pl);
    const Vec<T> pz_jp = sv::vload_if(up, psi.p[2] + pl + nb.up);
    const Vec<T> pz_jm = sv::vload_if(dn, psi.p[2] + pl + nb.dn);
    const Vec<T> px_jp = sv::vload_if(up, psi.p[0] + pl + nb.up);
    const Vec<T> px_jm = sv::vload_if(dn, psi.p[0] + pl + nb.dn);
    const T ezl = sv::sload_if(has_l, psi.p[2] + pl + nb.left), ezr = sv::sload_if(has_r, psi.p[2] + pl + nb.right);
    const T eyl = sv::sload_if(has_l, psi.p[1] + pl + nb.left), eyr = sv::sload_if(has_r, psi.p[1] + pl + nb.right);
    Vec<T> pz_im, pz_ip, py_im, py_ip;
    sv::x_neighbours(pzc, ezl, ezr, nb.use_l(lane), nb.use_r(lane), pz_im, pz_ip);
    sv::x_neighbours(pyc, eyl, eyr, nb.use_l(lane), nb.use_r(lane), py_im, py_ip);
    const bool kin = jin && k >= 1 && k < nz - 1;
    if (act && (!ACCUM || kin)) {  // accumulate mode: a ring plane / ring row keeps its values (warp-uniform)
      Vec<T> ux, uy, uz;
      const int64_t o = (int64_t)k * out.sz + ocol;
      if (ACCUM) ux = sv::vload_rw(out.p[0] + o), uy = sv::vload_rw(out.p[1] + o), uz = sv::vload_rw(out.p[2] + o);
#pragma unroll
      for (int m = 0; m < W; ++m) {
        const int i = i0 + m;
        if (ACCUM) {
          if (nb.iin(i, nx)) {
            ux.v[m] += p * (pz_jp.v[m] - pz_jm.v[m] - pyn.v[m] + pym.v[m]);
            uy.v[m] += p * (pxn.v[m] - pxm.v[m] - pz_ip.v[m] + pz_im.v[m]);
            uz.v[m] += p * (py_ip.v[m] - py_im.v[m] - px_jp.v[m] + px_jm.v[m]);
          }
          continue;
        }
        T vx = T(0), vy = T(0), vz = T(0);
        if (kin && nb.iin(i, nx)) {
          const T ccx = pz_jp.v[m] - pz_jm.v[m] - pyn.v[m] + pym.v[m];
          const T ccy = pxn.v[m] - pxm.v[m] - pz_ip.v[m] + pz_im.v[m];
          const T ccz = py_ip.v[m] - py_im.v[m] - px_jp.v[m] + px_jm.v[m];
          vx = p * ccx;
          vy = p * ccy;
          vz = p * ccz;
        }
        vx = vx + fx;
        vy = vy + fy;
        vz = vz + fz;
        ux.v[m] = vx, uy.v[m] = vy, uz.v[m] = vz;
        const T a = fabs(vx) + fabs(vy) + fabs(vz);
        m_acc = nanmax(a, m_acc);
      }
      sv::vstore(out.p[0] + o, ux);
      sv::vstore(out.p[1] + o, uy);
      sv::vstore(out.p[2] + o, uz);
    }
    pxm = pxc, pym = pyc;
    pxc = pxn, pyc = pyn;
  }
  if (!ACCUM && max_out) {
    for (int off = 16; off > 0; off >>= 1) {
      const T o = __shfl_xor_sync(0xffffffffu, m_acc, off);
      m_acc = nanmax(o, m_acc);
    }
    if (lane == 0) wmax[threadIdx.y] = m_acc;
    __syncthreads();
    if (threadIdx.y == 0 && lane == 0) {
      T m = wmax[0];
#pragma unroll
      for (int r = 1; r < VBY; ++r) m = nanmax(wmax[r], m);
      AtomicMaxNonNeg<T>::apply(max_out, m);
    }
  }
}

// eligibility of the register-marching kernels for a (3, nz, ny, nx) view
bool vec_ok(const sopht_field_t* f, int dtype) {
  const int64_t w = dtype == SOPHT_F32 ? 4 : 2;
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  if (f->stride[3] != 1 || f->shape[3] % w) return false;
  if (f->stride[0] % w || f->stride[1] % w || f->stride[2] % w) return false;
  return (reinterpret_cast<uintptr_t>(f->data) % (w * elem)) == 0;
}

// z chunking of the marching kernels. A launch of n CTAs on `slots` resident CTA slots (SMs x CTAs per SM) takes
// ceil(n / slots) waves, and every chunk re-reads two neighbour planes, so the chunk count is chosen to minimise
//   (ceil(waves) / waves) * (1 + 4/3 / planes_per_chunk)        (reads are about 2/3 of a stencil pass's traffic)
// instead of a fixed "6 CTAs per SM": at 512^3 the old rule gave 3.46 waves (13 % of the last wave idle).
template <class Kernel>
int resident_ctas(Kernel kernel, int threads) {
  int dev = 0, num_sm = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return num_sm * per_sm;
}
int pick_vec_kchunk(int nz, int ny, int nx, int ncomp_grids, int w, int slots) {
  static const char* force = getenv("SOPHT_VEC_KCHUNK");  // experiments
  if (force && atoi(force) > 0) return atoi(force) < nz ? atoi(force) : nz;
  const int64_t xy = (int64_t)((nx + 32 * w - 1) / (32 * w)) * ((ny + VBY - 1) / VBY) * ncomp_grids;
  int best_k = nz;
  double best_cost = 1e30;
  for (int chunks = 1; chunks <= nz; ++chunks) {
    const int k = (nz + chunks - 1) / chunks;
    if (k < 8 && chunks > 1) break;
    const int real_chunks = (nz + k - 1) / k;
    const double waves = (double)(xy * real_chunks) / slots;
    const double full = waves < 1.0 ? 1.0 : (double)(int64_t)(waves + 0.999999);
    // below one wave the machine is not filled at all: count the idle slots in full
    const double cost = (full / waves) * (1.0 + (real_chunks > 1 ? 4.0 / 3.0 / k : 0.0));
    if (cost < best_cost - 1e-9) best_cost = cost, best_k = k;
  }
  return best_k;
}

// ---- host side --------------------------------------------------------------------------------------
int check_vec3(const char* fn, const sopht_field_t* f) {
  if (!valid_field(f, 4, 4) || f->shape[0] != 3)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected (3, nz, ny, nx) fields", fn);
  if (f->stride[3] != 1 && f->shape[3] > 1)
    SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: fused kernels need unit x-stride fields", fn);
  for (int d = 1; d < 4; ++d)
    if (f->shape[d] > 0x7fffffff) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
  return SOPHT_OK;
}

template <typename T>
Vec3View<T> in_view(const sopht_field_t* f) {
  Vec3View<T> v;
  for (int c = 0; c < 3; ++c) v.p[c] = reinterpret_cast<const T*>(f->data) + c * f->stride[0];
  v.sz = f->stride[1];
  v.sy = f->stride[2];
  return v;
}
template <typename T>
Vec3Out<T> out_view(const sopht_field_t* f) {
  Vec3Out<T> v;
  for (int c = 0; c < 3; ++c) v.p[c] = reinterpret_cast<T*>(f->data) + c * f->stride[0];
  v.sz = f->stride[1];
  v.sy = f->stride[2];
  return v;
}

bool overlaps(const sopht_field_t* a, const sopht_field_t* b, size_t elem) {
  auto span = [&](const sopht_field_t* f, const char** lo, const char** hi) {
    int64_t last = 0;
    for (int d = 0; d < f->ndim; ++d) last += (f->shape[d] - 1) * f->stride[d];
    *lo = reinterpret_cast<const char*>(f->data);
    *hi = *lo + (last + 1) * elem;
  };
  const char *alo, *ahi, *blo, *bhi;
  span(a, &alo, &ahi);
  span(b, &blo, &bhi);
  return alo < bhi && blo < ahi;
}

// z chunking: enough CTAs for >= ~4 waves of 2 CTAs/SM, chunks of at least 8 planes
int pick_kchunk(int nz, int ny, int nx, int ncomp_grids, int tx, int ty) {
  const int64_t xy = (int64_t)((nx + tx - 1) / tx) * ((ny + ty - 1) / ty) * ncomp_grids;
  int64_t chunks = (148 * 2 * 4 + xy - 1) / xy;
  if (chunks < 1) chunks = 1;
  int kchunk = (int)((nz + chunks - 1) / chunks);
  if (kchunk < 8) kchunk = 8;
  if (kchunk > nz) kchunk = nz;
  return kchunk;
}

#undef FTX
#undef FTY
#undef FSX
#undef FSY
#undef FTHREADS

}  // namespace

// w += p * curl_c(f) with the register-marching kernel when both views are vector-aligned and distinct;
// returns false (nothing launched) otherwise and the caller falls back to the one-thread-per-cell kernel.
bool ns3d_try_forcing_curl_vec(int dtype, const sopht_field_t* vorticity_field,
                               const sopht_field_t* velocity_forcing_field, double prefactor, cudaStream_t st) {
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  if (!vec_ok(vorticity_field, dtype) || !vec_ok(velocity_forcing_field, dtype) ||
      overlaps(vorticity_field, velocity_forcing_field, elem))
    return false;
  const int nz = (int)vorticity_field->shape[1], ny = (int)vorticity_field->shape[2],
            nx = (int)vorticity_field->shape[3];
  const int w = dtype == SOPHT_F32 ? 4 : 2;
  static const int slots32 = resident_ctas(velocity_vec_kernel<float, true>, 32 * VBY);
  static const int slots64 = resident_ctas(velocity_vec_kernel<double, true>, 32 * VBY);
  const int kchunk = pick_vec_kchunk(nz, ny, nx, 1, w, dtype == SOPHT_F32 ? slots32 : slots64);
  dim3 grid((nx + 32 * w - 1) / (32 * w), (ny + VBY - 1) / VBY, (nz + kchunk - 1) / kchunk), block(32, VBY, 1);
  if (grid.y > 65535 || grid.z > 65535) return false;
  if (dtype == SOPHT_F32)
    velocity_vec_kernel<float, true><<<grid, block, 0, st>>>(
        out_view<float>(vorticity_field), in_view<float>(velocity_forcing_field), (float)prefactor, 0.f, 0.f,
        0.f, nullptr, nz, ny, nx, kchunk);
  else
    velocity_vec_kernel<double, true><<<grid, block, 0, st>>>(
        out_view<double>(vorticity_field), in_view<double>(velocity_forcing_field), prefactor, 0.0, 0.0, 0.0,
        nullptr, nz, ny, nx, kchunk);
  return true;
}

}  // namespace sopht

using namespace sopht;

#define RETURN_IF(rc_expr) \
  do {                     \
    int rc__ = (rc_expr);  \
    if (rc__) return rc__; \
  } while (0)

static int ns3d_advect(const char* fn, bool pxy, int dtype, const sopht_field_t* out_vorticity_field,
                       const sopht_field_t* vorticity_field, const sopht_field_t* velocity_field,
                       double prefactor, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vec3(fn, out_vorticity_field));
  RETURN_IF(check_vec3(fn, vorticity_field));
  RETURN_IF(check_vec3(fn, velocity_field));
  if (!same_shape(out_vorticity_field, vorticity_field) || !same_shape(out_vorticity_field, velocity_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", fn);
  if (vorticity_field->stride[1] != velocity_field->stride[1] ||
      vorticity_field->stride[2] != velocity_field->stride[2])
    SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: vorticity and velocity must share their plane/row strides", fn);
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  if (overlaps(out_vorticity_field, vorticity_field, elem) || overlaps(out_vorticity_field, velocity_field, elem))
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the output must not alias an input (neighbours are read)", fn);
  const int nz = (int)vorticity_field->shape[1], ny = (int)vorticity_field->shape[2],
            nx = (int)vorticity_field->shape[3];
  if ((int64_t)nz * ny * nx == 0) return SOPHT_OK;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ns3d.advect", st);
  if (vec_ok(out_vorticity_field, dtype) && vec_ok(vorticity_field, dtype) && vec_ok(velocity_field, dtype)) {
    const int w = dtype == SOPHT_F32 ? 4 : 2;
    static const int slots32 = resident_ctas(advect_vec_kernel<float, false>, 32 * VBY);
    static const int slots64 = resident_ctas(advect_vec_kernel<double, false>, 32 * VBY);
    const int kchunk = pick_vec_kchunk(nz, ny, nx, 1, w, dtype == SOPHT_F32 ? slots32 : slots64);
    dim3 grid((nx + 32 * w - 1) / (32 * w), (ny + VBY - 1) / VBY, (nz + kchunk - 1) / kchunk), block(32, VBY, 1);
    if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
#define SOPHT_LAUNCH_ADVECT(T, P)                                                                          \
  advect_vec_kernel<T, P><<<grid, block, 0, st>>>(out_view<T>(out_vorticity_field), in_view<T>(vorticity_field), \
                                                  in_view<T>(velocity_field), (T)prefactor, nz, ny, nx, kchunk)
    if (dtype == SOPHT_F32) {
      if (pxy) SOPHT_LAUNCH_ADVECT(float, true); else SOPHT_LAUNCH_ADVECT(float, false);
    } else {
      if (pxy) SOPHT_LAUNCH_ADVECT(double, true); else SOPHT_LAUNCH_ADVECT(double, false);
    }
#undef SOPHT_LAUNCH_ADVECT
    SOPHT_CHECK_LAUNCH();
    return SOPHT_OK;
  }
  if (pxy) SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: the periodic kernels need 16-byte aligned, unit-stride rows", fn);
  const int tx = dtype == SOPHT_F32 ? Tile<float>::TX : Tile<double>::TX, ty = Tile<float>::TY;
  const int kchunk = pick_kchunk(nz, ny, nx, 1, tx, ty);
  dim3 grid((nx + tx - 1) / tx, (ny + ty - 1) / ty, (nz + kchunk - 1) / kchunk), block(tx, ty, 1);
  if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
  if (dtype == SOPHT_F32)
    advect_rotational_kernel<float><<<grid, block, 0, st>>>(
        out_view<float>(out_vorticity_field), in_view<float>(vorticity_field), in_view<float>(velocity_field),
        (float)prefactor, nz, ny, nx, kchunk);
  else
    advect_rotational_kernel<double><<<grid, block, 0, st>>>(
        out_view<double>(out_vorticity_field), in_view<double>(vorticity_field),
        in_view<double>(velocity_field), prefactor, nz, ny, nx, kchunk);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

static int ns3d_diffuse(const char* fn, bool pxy, int dtype, const sopht_field_t* out_field,
                        const sopht_field_t* field, double nu_dt_by_dx2, const sopht_field_t* zero_field,
                        void* stream, const void* ramp_x = nullptr, const void* ramp_y = nullptr,
                        const void* ramp_z = nullptr) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vec3(fn, out_field));
  RETURN_IF(check_vec3(fn, field));
  if (!same_shape(out_field, field)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", fn);
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  if (overlaps(out_field, field, elem))
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the output must not alias the input (neighbours are read)", fn);
  if (zero_field) {
    RETURN_IF(check_vec3(fn, zero_field));
    if (!same_shape(zero_field, field)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", fn);
    if (overlaps(zero_field, field, elem) || overlaps(zero_field, out_field, elem))
      SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the field to zero must not alias input or output", fn);
  }
  const int nz = (int)field->shape[1], ny = (int)field->shape[2], nx = (int)field->shape[3];
  if ((int64_t)nz * ny * nx == 0) return SOPHT_OK;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ns3d.diffuse", st);
  if (vec_ok(out_field, dtype) && vec_ok(field, dtype) && (!zero_field || vec_ok(zero_field, dtype))) {
    const int w = dtype == SOPHT_F32 ? 4 : 2;
    static const bool three = [] {  // SOPHT_DIFFUSE_VEC3=0: the one-component-per-thread kernel for fp32 too
      const char* e = getenv("SOPHT_DIFFUSE_VEC3");
      return !e || atoi(e) != 0;
    }();
    if (dtype == SOPHT_F32 && three && !pxy) {
      static const int slots = resident_ctas(diffuse_vec3_kernel<float, false, false>, 32 * VBY);
      const int kchunk = pick_vec_kchunk(nz, ny, nx, 1, w, slots);
      dim3 grid((nx + 32 * w - 1) / (32 * w), (ny + VBY - 1) / VBY, (nz + kchunk - 1) / kchunk), block(32, VBY, 1);
      if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
      const Vec3Out<float> o = out_view<float>(out_field), z = zero_field ? out_view<float>(zero_field) : o;
      const Vec3View<float> f = in_view<float>(field);
      const float q = (float)nu_dt_by_dx2;
      const int hz = zero_field != nullptr;
      if (ramp_x)
        diffuse_vec3_kernel<float, false, true><<<grid, block, 0, st>>>(o, f, z, hz, q, nz, ny, nx, kchunk, (const float*)ramp_x,
                                                                       (const float*)ramp_y, (const float*)ramp_z);
      else
        diffuse_vec3_kernel<float, false, false><<<grid, block, 0, st>>>(o, f, z, hz, q, nz, ny, nx, kchunk, nullptr, nullptr,
                                                                        nullptr);
      SOPHT_CHECK_LAUNCH();
      return SOPHT_OK;
    }
    static const int slots32 = resident_ctas(diffuse_vec_kernel<float, false>, 32 * VBY);
    static const int slots64 = resident_ctas(diffuse_vec_kernel<double, false>, 32 * VBY);
    const int kchunk = pick_vec_kchunk(nz, ny, nx, 3, w, dtype == SOPHT_F32 ? slots32 : slots64);
    const int nchunk = (nz + kchunk - 1) / kchunk;
    dim3 grid((nx + 32 * w - 1) / (32 * w), (ny + VBY - 1) / VBY, nchunk * 3), block(32, VBY, 1);
    if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
#define SOPHT_LAUNCH_DIFFUSE(T, P)                                                                       \
  diffuse_vec_kernel<T, P><<<grid, block, 0, st>>>(                                                     \
      out_view<T>(out_field), in_view<T>(field), zero_field ? out_view<T>(zero_field) : out_view<T>(out_field), \
      zero_field != nullptr, 3, (T)nu_dt_by_dx2, nz, ny, nx, kchunk)
    if (ramp_x) {
      if (pxy) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: no boundary ramps in a periodic box", fn);
      if (dtype == SOPHT_F32)
        diffuse_vec_kernel<float, false, true><<<grid, block, 0, st>>>(
            out_view<float>(out_field), in_view<float>(field),
            zero_field ? out_view<float>(zero_field) : out_view<float>(out_field), zero_field != nullptr, 3,
            (float)nu_dt_by_dx2, nz, ny, nx, kchunk, (const float*)ramp_x, (const float*)ramp_y, (const float*)ramp_z);
      else
        diffuse_vec_kernel<double, false, true><<<grid, block, 0, st>>>(
            out_view<double>(out_field), in_view<double>(field),
            zero_field ? out_view<double>(zero_field) : out_view<double>(out_field), zero_field != nullptr, 3,
            nu_dt_by_dx2, nz, ny, nx, kchunk, (const double*)ramp_x, (const double*)ramp_y, (const double*)ramp_z);
    } else if (dtype == SOPHT_F32) {
      if (pxy) SOPHT_LAUNCH_DIFFUSE(float, true); else SOPHT_LAUNCH_DIFFUSE(float, false);
    } else {
      if (pxy) SOPHT_LAUNCH_DIFFUSE(double, true); else SOPHT_LAUNCH_DIFFUSE(double, false);
    }
#undef SOPHT_LAUNCH_DIFFUSE
    SOPHT_CHECK_LAUNCH();
    return SOPHT_OK;
  }
  if (pxy) SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: the periodic kernels need 16-byte aligned, unit-stride rows", fn);
  if (ramp_x) SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: the fused penalisation needs 16-byte aligned, unit-stride rows", fn);
  const int tx = dtype == SOPHT_F32 ? Tile<float>::TX : Tile<double>::TX, ty = Tile<float>::TY;
  const int kchunk = pick_kchunk(nz, ny, nx, 3, tx, ty);
  const int nchunk = (nz + kchunk - 1) / kchunk;
  dim3 grid((nx + tx - 1) / tx, (ny + ty - 1) / ty, nchunk * 3), block(tx, ty, 1);
  if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
  if (dtype == SOPHT_F32)
    diffuse_kernel<float><<<grid, block, 0, st>>>(
        out_view<float>(out_field), in_view<float>(field),
        zero_field ? out_view<float>(zero_field) : out_view<float>(out_field), zero_field != nullptr, 3,
        (float)nu_dt_by_dx2, nz, ny, nx, kchunk);
  else
    diffuse_kernel<double><<<grid, block, 0, st>>>(
        out_view<double>(out_field), in_view<double>(field),
        zero_field ? out_view<double>(zero_field) : out_view<double>(out_field), zero_field != nullptr, 3,
        nu_dt_by_dx2, nz, ny, nx, kchunk);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

static int ns3d_velocity(const char* fn, bool pxy, int dtype, const sopht_field_t* velocity_field,
                         const sopht_field_t* stream_func_field, double prefactor,
                         const double* free_stream_velocity, void* max_abs_sum_out, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  RETURN_IF(check_vec3(fn, velocity_field));
  RETURN_IF(check_vec3(fn, stream_func_field));
  if (!same_shape(velocity_field, stream_func_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shapes differ", fn);
  const size_t elem = dtype == SOPHT_F32 ? 4 : 8;
  if (overlaps(velocity_field, stream_func_field, elem))
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the output must not alias the input (neighbours are read)", fn);
  const int nz = (int)velocity_field->shape[1], ny = (int)velocity_field->shape[2],
            nx = (int)velocity_field->shape[3];
  if ((int64_t)nz * ny * nx == 0) return SOPHT_OK;
  const double f0 = free_stream_velocity ? free_stream_velocity[0] : 0.0;
  const double f1 = free_stream_velocity ? free_stream_velocity[1] : 0.0;
  const double f2 = free_stream_velocity ? free_stream_velocity[2] : 0.0;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ns3d.velocity", st);
  if (max_abs_sum_out) SOPHT_CUDA(cudaMemsetAsync(max_abs_sum_out, 0, elem, st));
  if (vec_ok(velocity_field, dtype) && vec_ok(stream_func_field, dtype)) {
    const int w = dtype == SOPHT_F32 ? 4 : 2;
    static const int slots32 = resident_ctas(velocity_vec_kernel<float, false>, 32 * VBY);
    static const int slots64 = resident_ctas(velocity_vec_kernel<double, false>, 32 * VBY);
    const int kchunk = pick_vec_kchunk(nz, ny, nx, 1, w, dtype == SOPHT_F32 ? slots32 : slots64);
    dim3 grid((nx + 32 * w - 1) / (32 * w), (ny + VBY - 1) / VBY, (nz + kchunk - 1) / kchunk), block(32, VBY, 1);
    if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
#define SOPHT_LAUNCH_VELOCITY(T, P)                                                                      \
  velocity_vec_kernel<T, false, P><<<grid, block, 0, st>>>(out_view<T>(velocity_field), in_view<T>(stream_func_field), \
                                                           (T)prefactor, (T)f0, (T)f1, (T)f2,           \
                                                           reinterpret_cast<T*>(max_abs_sum_out), nz, ny, nx, kchunk)
    if (dtype == SOPHT_F32) {
      if (pxy) SOPHT_LAUNCH_VELOCITY(float, true); else SOPHT_LAUNCH_VELOCITY(float, false);
    } else {
      if (pxy) SOPHT_LAUNCH_VELOCITY(double, true); else SOPHT_LAUNCH_VELOCITY(double, false);
    }
#undef SOPHT_LAUNCH_VELOCITY
    SOPHT_CHECK_LAUNCH();
    return SOPHT_OK;
  }
  if (pxy) SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: the periodic kernels need 16-byte aligned, unit-stride rows", fn);
  const int tx = dtype == SOPHT_F32 ? Tile<float>::TX : Tile<double>::TX, ty = Tile<float>::TY;
  const int kchunk = pick_kchunk(nz, ny, nx, 1, tx, ty);
  dim3 grid((nx + tx - 1) / tx, (ny + ty - 1) / ty, (nz + kchunk - 1) / kchunk), block(tx, ty, 1);
  if (grid.y > 65535 || grid.z > 65535) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
  if (dtype == SOPHT_F32)
    velocity_from_psi_kernel<float><<<grid, block, 0, st>>>(
        out_view<float>(velocity_field), in_view<float>(stream_func_field), (float)prefactor, (float)f0,
        (float)f1, (float)f2, reinterpret_cast<float*>(max_abs_sum_out), nz, ny, nx, kchunk);
  else
    velocity_from_psi_kernel<double><<<grid, block, 0, st>>>(
        out_view<double>(velocity_field), in_view<double>(stream_func_field), prefactor, f0, f1, f2,
        reinterpret_cast<double*>(max_abs_sum_out), nz, ny, nx, kchunk);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}


namespace sopht {
// plane 0 <- plane nz, plane nz + 1 <- plane 1 of every component; 16-byte vectors when rows allow it
template <typename V>
static __global__ void __launch_bounds__(256)
    wrap_z_kernel(char* base, int64_t comp_stride_b, int64_t plane_stride_b, int64_t row_stride_b, int ncomp,
                  int nzp, int ny, int row_vecs) {
  const int64_t per = (int64_t)ny * row_vecs, total = per * ncomp * 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(q % row_vecs);
    const int j = (int)((q / row_vecs) % ny);
    const int side = (int)((q / per) & 1), c = (int)(q / per / 2);
    char* comp = base + c * comp_stride_b + j * row_stride_b + (int64_t)v * sizeof(V);
    const int src = side ? 1 : nzp - 2, dst = side ? nzp - 1 : 0;
    *reinterpret_cast<V*>(comp + dst * plane_stride_b) = *reinterpret_cast<const V*>(comp + src * plane_stride_b);
  }
}
}  // namespace sopht

extern "C" {

int sopht_ns3d_advect_rotational(int dtype, const sopht_field_t* out_vorticity_field,
                                 const sopht_field_t* vorticity_field, const sopht_field_t* velocity_field,
                                 double prefactor, void* stream) {
  return ns3d_advect(__func__, false, dtype, out_vorticity_field, vorticity_field, velocity_field, prefactor, stream);
}
int sopht_ns3d_diffuse(int dtype, const sopht_field_t* out_field, const sopht_field_t* field, double nu_dt_by_dx2,
                       const sopht_field_t* zero_field, void* stream) {
  return ns3d_diffuse(__func__, false, dtype, out_field, field, nu_dt_by_dx2, zero_field, stream);
}
int sopht_ns3d_velocity_from_stream_function(int dtype, const sopht_field_t* velocity_field,
                                             const sopht_field_t* stream_func_field, double prefactor,
                                             const double* free_stream_velocity, void* max_abs_sum_out,
                                             void* stream) {
  return ns3d_velocity(__func__, false, dtype, velocity_field, stream_func_field, prefactor, free_stream_velocity,
                       max_abs_sum_out, stream);
}


int sopht_wrap_z_halos(int dtype, const sopht_field_t* field, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(field, 3, 4)) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (C, nz + 2, ny, nx) or (nz + 2, ny, nx) field", __func__);
  const int o = field->ndim == 4 ? 1 : 0;
  const int ncomp = o ? (int)field->shape[0] : 1;
  const int nzp = (int)field->shape[o], ny = (int)field->shape[o + 1], nx = (int)field->shape[o + 2];
  if (nzp < 3) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: at least one owned plane between the two halo planes", __func__);
  if (field->stride[o + 2] != 1 && nx > 1) SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: unit x stride required", __func__);
  if ((int64_t)ny * nx == 0) return SOPHT_OK;
  const int64_t elem = dtype == SOPHT_F32 ? 4 : 8;
  const int64_t cs = (o ? field->stride[0] : 0) * elem, ps = field->stride[o] * elem, rs = field->stride[o + 1] * elem;
  const bool vec = ((nx * elem) % 16 == 0) && (cs % 16 == 0) && (ps % 16 == 0) && (rs % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(field->data) % 16 == 0);
  const int row_vecs = vec ? (int)(nx * elem / 16) : nx;
  const int64_t total = (int64_t)ny * row_vecs * ncomp * 2;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ns3d.wrap_z", st);
  char* base = reinterpret_cast<char*>(field->data);
  if (vec)
    wrap_z_kernel<uint4><<<blocks, 256, 0, st>>>(base, cs, ps, rs, ncomp, nzp, ny, row_vecs);
  else if (elem == 4)
    wrap_z_kernel<float><<<blocks, 256, 0, st>>>(base, cs, ps, rs, ncomp, nzp, ny, row_vecs);
  else
    wrap_z_kernel<double><<<blocks, 256, 0, st>>>(base, cs, ps, rs, ncomp, nzp, ny, row_vecs);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

/* diffuse + the width-2 sine penalisation of the boundary ring in one pass: ramp_{x,y,z} are DEVICE arrays of nx / ny / nz
 * factors of `dtype` (1 in the interior; the ramp values of penalise_field_boundary_3d.py:62-94 on the two cells next to
 * each face, i.e. {0, sin(pi/4)} for width 2). ref: navier_stokes_flow_simulators.py:466-478 */
int sopht_ns3d_diffuse_penalise(int dtype, const sopht_field_t* out_field, const sopht_field_t* field,
                                double nu_dt_by_dx2, const sopht_field_t* zero_field, const void* ramp_x,
                                const void* ramp_y, const void* ramp_z, void* stream) {
  if (!ramp_x || !ramp_y || !ramp_z) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null ramp array", __func__);
  return ns3d_diffuse(__func__, false, dtype, out_field, field, nu_dt_by_dx2, zero_field, stream, ramp_x, ramp_y,
                      ramp_z);
}

/* periodic box: x and y wrap around inside the kernels; z keeps the ghost-ring rule, so the arrays carry one halo plane
 * per z side that the caller fills (wrap copy on one GPU, neighbour planes in a slab decomposition) */
int sopht_ns3d_advect_rotational_periodic_xy(int dtype, const sopht_field_t* out_vorticity_field,
                                             const sopht_field_t* vorticity_field,
                                             const sopht_field_t* velocity_field, double prefactor, void* stream) {
  return ns3d_advect(__func__, true, dtype, out_vorticity_field, vorticity_field, velocity_field, prefactor, stream);
}
int sopht_ns3d_diffuse_periodic_xy(int dtype, const sopht_field_t* out_field, const sopht_field_t* field,
                                   double nu_dt_by_dx2, const sopht_field_t* zero_field, void* stream) {
  return ns3d_diffuse(__func__, true, dtype, out_field, field, nu_dt_by_dx2, zero_field, stream);
}
int sopht_ns3d_velocity_from_stream_function_periodic_xy(int dtype, const sopht_field_t* velocity_field,
                                                         const sopht_field_t* stream_func_field, double prefactor,
                                                         const double* free_stream_velocity, void* max_abs_sum_out,
                                                         void* stream) {
  return ns3d_velocity(__func__, true, dtype, velocity_field, stream_func_field, prefactor, free_stream_velocity,
                       max_abs_sum_out, stream);
}

}  // extern "C"
