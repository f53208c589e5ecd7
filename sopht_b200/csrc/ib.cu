// Immersed-boundary Eulerian <-> Lagrangian transfer with 4-point discrete delta kernels.
//
// ref: sopht/numeric/immersed_boundary_ops/EulerianLagrangianGridCommunicator3D.py:7-518 (and ...2D.py),
//      VirtualBoundaryForcing.py:132-283.
//
// Mapping: one WARP per Lagrangian node. The 4^dim taps of a node are spread over the lanes (3-D: two
// taps per lane, x fastest so four consecutive lanes touch 16 contiguous bytes of a grid row); the
// gather is a warp-shuffle reduction, the spread a warp of reductions (red.global.add) into the grid.
// The API-level kernels fill the reference's (dim,4,4,4,N) support and (4,4,4,N) weight arrays; the fused
// virtual-boundary kernel does support + weights + gather + penalty force + spread for a node in ONE
// launch, computing the separable weights in registers (and still storing the public buffers).
// Taps that fall outside the grid are skipped (the reference indexes out of bounds there).
#include <stdlib.h>

#include "common.cuh"

namespace sopht {

template <int DIM>
struct Taps {
  static constexpr int N = DIM == 3 ? 64 : 16;
};

// numpy / numba floor division of floats (npy_divmod): used for the nearest-index computation so the
// integer index is bit-identical to the reference's `(X - shift) // dx`.
template <typename P>
__device__ __forceinline__ P np_floor_divide(P a, P b) {
  P mod = fmod(a, b);
  P div = (a - mod) / b;
  if (mod != P(0)) {
    if ((b < P(0)) != (mod < P(0))) div -= P(1);
  }
  P fl;
  if (div != P(0)) {
    fl = floor(div);
    if (div - fl > P(0.5)) fl += P(1);
  } else {
    fl = copysign(P(0), a / b);
  }
  return fl;
}

struct EulView {
  void* p;
  int64_t sc, s[3];  // component stride, then (z, y, x) / (y, x) strides
  int n[3];          // grid extents, slowest first
  int ncomp;
};

template <int DIM>
__device__ __forceinline__ void tap_offsets(int tap, int off[3]) {
  // tap = (kz*4 + ky)*4 + kx ; component d (0 = x) varies along the LAST stencil axis
  off[0] = (tap & 3) - 1;
  off[1] = ((tap >> 2) & 3) - 1;
  off[2] = DIM == 3 ? ((tap >> 4) & 3) - 1 : 0;
}

// returns the element offset of tap cell, or -1 when outside the grid. idx[d]: d = 0 is x.
template <int DIM>
__device__ __forceinline__ int64_t tap_cell(const EulView& e, const int64_t idx[3], const int off[3]) {
  int64_t o = 0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const int64_t c = idx[d] + off[d];
    const int ax = DIM - 1 - d;  // array axis of coordinate d
    if (c < 0 || c >= e.n[ax]) return -1;
    o += c * e.s[ax];
  }
  return o;
}

// ---- 1. nearest index + local support ---------------------------------------------------------------
template <typename T, typename P, int DIM>
__global__ void __launch_bounds__(256)
    ib_support_kernel(T* support, int64_t* nearest, const P* pos, int64_t pos_sd, int64_t pos_sn,
                      int64_t n_lag, P dx_p, P shift_p, double dx, double shift) {
  constexpr int NT = Taps<DIM>::N;
  const int64_t total = (int64_t)NT * n_lag;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(e / n_lag);
    const int64_t i = e - (int64_t)tap * n_lag;
    int off[3];
    tap_offsets<DIM>(tap, off);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const P x = pos[d * pos_sd + i * pos_sn];
      const int64_t id = (int64_t)np_floor_divide<P>(x - shift_p, dx_p);
      if (tap == 0) nearest[d * n_lag + i] = id;
      support[((int64_t)d * NT + tap) * n_lag + i] = (T)((double)(id + off[d]) * dx + shift - (double)x);
    }
  }
}

// ---- 2. interpolation weights (mutates support like the reference) -----------------------------------
template <typename T>
__device__ __forceinline__ double peskin_phi(T r_t) {
  const double r = (double)r_t;
  double v = 0.0;
  if (r < 1.0) v = 3.0 - 2 * r + sqrt(fabs(1 + 4 * r - 4 * r * r));
  if (r >= 1.0 && r < 2.0) v = 5.0 - 2 * r - sqrt(fabs(-7 + 12 * r - 4 * r * r));
  return v;
}

template <typename T, int DIM, int KIND>  // KIND 0 cosine, 1 peskin
__global__ void __launch_bounds__(256)
    ib_weights_kernel(T* weights, T* support, int64_t n_lag, T dx, double prefactor) {
  constexpr int NT = Taps<DIM>::N;
  const int64_t total = (int64_t)NT * n_lag;
  const T half_pi = (T)(0.5 * 3.14159265358979323846);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (KIND == 0) {
      T w = (T)prefactor;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        T* sp = support + (int64_t)d * total + e;
        const T r = *sp / dx;
        *sp = r;
        w = w * (T(1) + cos(half_pi * r));
      }
      weights[e] = w;
    } else {
      double w = prefactor;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        T* sp = support + (int64_t)d * total + e;
        const T r = fabs(*sp) / dx;
        *sp = r;
        w = w * peskin_phi<T>(r);
      }
      weights[e] = (T)w;
    }
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 3. Eulerian -> Lagrangian gather ------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
    ib_gather_kernel(T* lag, int64_t lag_sc, EulView eul, const T* weights, const int64_t* nearest,
                     int64_t n_lag, T vol) {
  constexpr int NT = Taps<DIM>::N;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const T* ep = reinterpret_cast<const T*>(eul.p);
  for (int64_t i = warp; i < n_lag; i += nwarps) {
    int64_t idx[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) idx[d] = nearest[d * n_lag + i];
    T acc[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int t = lane; t < NT; t += 32) {
      int off[3];
      tap_offsets<DIM>(t, off);
      const int64_t cell = tap_cell<DIM>(eul, idx, off);
      if (cell >= 0) {
        const T w = weights[(int64_t)t * n_lag + i];
        for (int c = 0; c < eul.ncomp; ++c) acc[c] += ep[c * eul.sc + cell] * w;
      }
    }
    for (int c = 0; c < eul.ncomp; ++c) {
      const T s = warp_sum(acc[c]);
      if (lane == 0) lag[c * lag_sc + i] = s * vol;
    }
  }
}

// ---- 4. Lagrangian -> Eulerian spread ------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
    ib_spread_kernel(EulView eul, const T* lag, int64_t lag_sc, const T* weights, const int64_t* nearest,
                     int64_t n_lag) {
  constexpr int NT = Taps<DIM>::N;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  T* ep = reinterpret_cast<T*>(eul.p);
  for (int64_t i = warp; i < n_lag; i += nwarps) {
    int64_t idx[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) idx[d] = nearest[d * n_lag + i];
    T f[3] = {T(0), T(0), T(0)};
    for (int c = 0; c < eul.ncomp; ++c) f[c] = lag[c * lag_sc + i];
#pragma unroll
    for (int t = lane; t < NT; t += 32) {
      int off[3];
      tap_offsets<DIM>(t, off);
      const int64_t cell = tap_cell<DIM>(eul, idx, off);
      if (cell >= 0) {
        const T w = weights[(int64_t)t * n_lag + i];
        for (int c = 0; c < eul.ncomp; ++c) atomicAdd(ep + c * eul.sc + cell, f[c] * w);
      }
    }
  }
}

// ---- 5. fused virtual-boundary forcing -------------------------------------------------------------------
// For each node: index, support, cosine weights (registers + public buffers), velocity gather,
// dv = U - V_body, F = k dX + c dv, spread of F. VirtualBoundaryForcing.py:187-253 in one launch.
struct VbfArgs {
  void *support, *weights;                 // (dim, NT, N), (NT, N)
  int64_t* nearest;                        // (dim, N)
  void *flow_vel, *vel_mismatch, *pos_mismatch, *forcing;  // (dim, N) contiguous, real_t
  const void *pos, *body_vel;              // (dim, N), P
  int64_t pos_sd, pos_sn, bv_sd, bv_sn;
  int64_t n_lag;
  double dx, shift, prefactor, vol, stiffness, damping;
};

template <typename T, typename P, int DIM>
__global__ void __launch_bounds__(256) vbf_fused_kernel(VbfArgs a, EulView vel, EulView force) {
  constexpr int NT = Taps<DIM>::N;
  constexpr int TPL = NT / 32 > 0 ? NT / 32 : 1;  // taps per lane
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_lag = a.n_lag;
  const T* vp = reinterpret_cast<const T*>(vel.p);
  T* fp = reinterpret_cast<T*>(force.p);
  T* support = reinterpret_cast<T*>(a.support);
  T* weights = reinterpret_cast<T*>(a.weights);
  const P* pos = reinterpret_cast<const P*>(a.pos);
  const P* bvel = reinterpret_cast<const P*>(a.body_vel);
  const T dxT = (T)a.dx;
  const T half_pi = (T)(0.5 * 3.14159265358979323846);
  for (int64_t i = warp; i < n_lag; i += nwarps) {
    int64_t idx[3] = {0, 0, 0};
    P x[3] = {P(0), P(0), P(0)};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      x[d] = pos[d * a.pos_sd + i * a.pos_sn];
      idx[d] = (int64_t)np_floor_divide<P>(x[d] - (P)a.shift, (P)a.dx);
      if (lane == 0) a.nearest[d * n_lag + i] = idx[d];
    }
    T w[TPL];
    int64_t cell[TPL];
    T acc[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int q = 0; q < TPL; ++q) {
      const int t = lane + 32 * q;
      w[q] = T(0);
      cell[q] = -1;
      if (t < NT) {
        int off[3];
        tap_offsets<DIM>(t, off);
        T ww = (T)a.prefactor;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          const T s = (T)((double)(idx[d] + off[d]) * a.dx + a.shift - (double)x[d]);
          const T r = s / dxT;
          support[((int64_t)d * NT + t) * n_lag + i] = r;
          ww = ww * (T(1) + cos(half_pi * r));
        }
        weights[(int64_t)t * n_lag + i] = ww;
        w[q] = ww;
        cell[q] = tap_cell<DIM>(vel, idx, off);
        if (cell[q] >= 0) {
#pragma unroll
          for (int c = 0; c < DIM; ++c) acc[c] += vp[c * vel.sc + cell[q]] * w[q];
        }
      }
    }
    T f[3];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      const T u = warp_sum(acc[c]) * (T)a.vol;
      // dv = U_flow - V_body (mixed precision promotes like numpy: real_t - P evaluated in the wider type)
      const T dv = (T)((double)u - (double)bvel[c * a.bv_sd + i * a.bv_sn]);
      const T dxm = reinterpret_cast<const T*>(a.pos_mismatch)[c * n_lag + i];
      f[c] = (T)(a.stiffness * (double)dxm + a.damping * (double)dv);
      if (lane == 0) {
        reinterpret_cast<T*>(a.flow_vel)[c * n_lag + i] = u;
        reinterpret_cast<T*>(a.vel_mismatch)[c * n_lag + i] = dv;
        reinterpret_cast<T*>(a.forcing)[c * n_lag + i] = f[c];
      }
    }
    if (force.p) {
#pragma unroll
      for (int q = 0; q < TPL; ++q) {
        if (cell[q] >= 0) {
          // vel and force views share the grid shape; recompute the offset with force strides
          const int t = lane + 32 * q;
          int off[3];
          tap_offsets<DIM>(t, off);
          const int64_t fc = tap_cell<DIM>(force, idx, off);
#pragma unroll
          for (int c = 0; c < DIM; ++c) atomicAdd(fp + c * force.sc + fc, f[c] * w[q]);
        }
      }
    }
  }
}

// ---- 6. Lagrangian -> Eulerian spread without global floating-point atomics -----------------------------------
// The reference spreads node by node in a serial loop (EulerianLagrangianGridCommunicator3D.py:348-380): every cell
// receives its contributions in node order. `red.global.add` from one warp per node gives the same sum in a
// run-dependent order (fp addition is not associative: results differ in the last bits from run to run) and
// serialises on the cells shared by neighbouring nodes. Here the scatter is turned into a per-tile GATHER:
//   bin:   a node belongs to the grid tile (8^3 cells / 16^2 in 2-D) that holds the low corner of its 4^dim support;
//          one integer atomicAdd on the tile's counter gives it a slot in the tile's node list (capacity CAP; the
//          rare node beyond it is a straggler and spreads with atomics right there);
//   tile:  one CTA per tile, one thread per cell. The supports that can reach the tile start in the tile itself or in
//          the previous tile along each axis (support width 4 <= tile size): the CTA merges those 2^dim lists in
//          shared memory, sorts them by node index, and every thread walks the sorted list accumulating F_i * w for
//          the nodes whose support covers its cell. The thread owns its cell: one plain read-modify-write, no atomics,
//          and the summation order is the reference's (ascending node index) - bit-reproducible.
// The tile counters are cleared by a cudaMemsetAsync in front of the bin kernel.
template <int DIM>
struct SpreadTile {
  static constexpr int TS = DIM == 3 ? 8 : 16;                 // tile edge in cells (>= support width 4)
  static constexpr int THREADS = DIM == 3 ? 512 : 256;         // one thread per cell
  static constexpr int NSRC = DIM == 3 ? 8 : 4;                // source tiles of an output tile
  static constexpr int CAP = 128;                              // list slots per tile (a body surface with one node per
                                                               // cell^2 starts ~64 supports in an 8^3 tile)
  static constexpr int MAXN = NSRC * CAP;                      // merged list: a power of two, two entries per thread
};
struct SpreadWork {
  int* counts = nullptr;  // [ntiles] list lengths, [ntiles] "on the work list" flags, [1] work-list length: one memset
  int* lists = nullptr;   // [ntiles][CAP]
  int* active = nullptr;  // [ntiles] work list of the gather kernel
  unsigned long long* stragglers = nullptr;  // running count of nodes that overflowed their tile's list
  int64_t ntiles = 0;
  int dev = -1;
};
static SpreadWork g_spread_work;

static int spread_workspace(int64_t ntiles, int cap, SpreadWork** out) {
  SpreadWork& w = g_spread_work;
  int dev = 0;
  SOPHT_CUDA(cudaGetDevice(&dev));
  if (w.dev != dev || w.ntiles < ntiles) {
    cudaFree(w.counts);
    cudaFree(w.lists);
    cudaFree(w.active);
    if (!w.stragglers || w.dev != dev) {
      cudaFree(w.stragglers);
      SOPHT_CUDA(cudaMalloc(&w.stragglers, sizeof(unsigned long long)));
      SOPHT_CUDA(cudaMemset(w.stragglers, 0, sizeof(unsigned long long)));
    }
    w.counts = w.lists = w.active = nullptr;
    SOPHT_CUDA(cudaMalloc(&w.counts, sizeof(int) * (2 * ntiles + 1)));
    SOPHT_CUDA(cudaMalloc(&w.lists, sizeof(int) * ntiles * cap));
    SOPHT_CUDA(cudaMalloc(&w.active, sizeof(int) * ntiles));
    w.ntiles = ntiles;
    w.dev = dev;
  }
  *out = &w;
  return SOPHT_OK;
}

struct TileGrid {
  int nt[3];  // tiles per array axis (slowest first), 1 for unused axes
};
template <int DIM>
__device__ __forceinline__ int tile_of_corner(const EulView& e, const TileGrid& tg, const int64_t idx[3], bool* any) {
  // idx[d]: d = 0 is x; support corner = idx - 1. A support that misses the grid entirely has no tile.
  int t = 0;
  *any = true;
#pragma unroll
  for (int ax = 0; ax < DIM; ++ax) {
    const int d = DIM - 1 - ax;
    const int64_t c = idx[d] - 1;
    if (c + 3 < 0 || c >= e.n[ax]) *any = false;
    const int64_t cc = c < 0 ? 0 : c;
    int ta = (int)(cc / SpreadTile<DIM>::TS);
    if (ta >= tg.nt[ax]) ta = tg.nt[ax] - 1;
    t = t * tg.nt[ax] + ta;
  }
  return t;
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
    ib_bin_kernel(EulView eul, TileGrid tg, const T* lag, int64_t lag_sc, const T* weights, const int64_t* nearest,
                  int64_t n_lag, int* counts, int* lists, unsigned long long* stragglers, int* flags, int* active,
                  int* nactive) {
  constexpr int NT = Taps<DIM>::N, CAP = SpreadTile<DIM>::CAP;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lag) return;
  int64_t idx[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < DIM; ++d) idx[d] = nearest[d * n_lag + i];
  bool any;
  const int tile = tile_of_corner<DIM>(eul, tg, idx, &any);
  if (!any) return;
  const int slot = atomicAdd(counts + tile, 1);
  if (slot == 0) {
    // first node of this tile: its supports reach this tile and the next one along each axis - put those output tiles
    // on the work list of the gather kernel (once each: flags)
    int tc[3] = {0, 0, 0};
    int r = tile;
#pragma unroll
    for (int ax = DIM - 1; ax >= 0; --ax) {
      tc[ax] = r % tg.nt[ax];
      r /= tg.nt[ax];
    }
    for (int k = 0; k < SpreadTile<DIM>::NSRC; ++k) {
      int t = 0;
      bool ok = true;
#pragma unroll
      for (int ax = 0; ax < DIM; ++ax) {
        const int ta = tc[ax] + ((k >> (DIM - 1 - ax)) & 1);
        if (ta >= tg.nt[ax]) ok = false;
        t = t * tg.nt[ax] + ta;
      }
      if (ok && atomicExch(flags + t, 1) == 0) active[atomicAdd(nactive, 1)] = t;
    }
  }
  if (slot < CAP) {
    lists[(int64_t)tile * CAP + slot] = (int)i;
    return;
  }
  // straggler: more nodes start in this tile than its list holds - spread with atomics (order not reproducible)
  atomicAdd(stragglers, 1ull);
  T* ep = reinterpret_cast<T*>(eul.p);
  for (int t = 0; t < NT; ++t) {
    int off[3];
    tap_offsets<DIM>(t, off);
    const int64_t cell = tap_cell<DIM>(eul, idx, off);
    if (cell >= 0) {
      const T w = weights[(int64_t)t * n_lag + i];
      for (int c = 0; c < eul.ncomp; ++c) atomicAdd(ep + c * eul.sc + cell, lag[c * lag_sc + i] * w);
    }
  }
}

template <typename T, int DIM>
__global__ void __launch_bounds__(SpreadTile<DIM>::THREADS)
    ib_tile_gather_kernel(EulView eul, TileGrid tg, const T* lag, int64_t lag_sc, const T* weights,
                          const int64_t* nearest, int64_t n_lag, const int* counts, const int* lists,
                          const int* active, const int* nactive) {
  using ST = SpreadTile<DIM>;
  constexpr int TS = ST::TS, CAP = ST::CAP, MAXN = ST::MAXN, NSRC = ST::NSRC;
  __shared__ int s_node[MAXN];
  __shared__ int s_corner[MAXN][3];  // per array axis (slowest first)
  __shared__ T s_force[MAXN][3];
  __shared__ int s_cnt[NSRC], s_src[NSRC];
  const int tid = threadIdx.x;
  const int na = *nactive;
  // the blocks walk the list of tiles some support reaches (a few hundred for a body surface; the order of the list
  // varies from run to run, a tile's result does not depend on it)
  for (int a = blockIdx.x; a < na; a += gridDim.x) {
    const int tile = active[a];
    int tc[3] = {0, 0, 0};
    {
      int r = tile;
#pragma unroll
      for (int ax = DIM - 1; ax >= 0; --ax) {
        tc[ax] = r % tg.nt[ax];
        r /= tg.nt[ax];
      }
    }
    __syncthreads();  // the previous tile's readers of the shared arrays are done
    if (tid < NSRC) {
      int t = 0;
      bool ok = true;
#pragma unroll
      for (int ax = 0; ax < DIM; ++ax) {
        const int ta = tc[ax] - ((tid >> (DIM - 1 - ax)) & 1);
        if (ta < 0) ok = false;
        t = t * tg.nt[ax] + ta;
      }
      const int c = ok ? counts[t] : 0;
      s_cnt[tid] = c < CAP ? c : CAP;
      s_src[tid] = ok ? t : -1;
    }
    __syncthreads();
    int total = 0;
    int src_start[NSRC];
#pragma unroll
    for (int k = 0; k < NSRC; ++k) {
      src_start[k] = total;
      total += s_cnt[k];
    }
    if (total == 0) continue;  // block-uniform
    // merged list, padded with INT_MAX up to the next power of two >= total, then a bitonic sort by node index
    int len = 2;
    while (len < total) len <<= 1;
    for (int e = tid; e < len; e += ST::THREADS) {
      int mine = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < NSRC; ++k)
        if (e >= src_start[k] && e < src_start[k] + s_cnt[k]) mine = lists[(int64_t)s_src[k] * CAP + (e - src_start[k])];
      s_node[e] = mine;
    }
    __syncthreads();
    for (int k = 2; k <= len; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < len / 2; i += ST::THREADS) {
          const int lo = 2 * i - (i & (j - 1)), hi = lo + j;  // the i-th compare-exchange pair of this stage
          const int x = s_node[lo], y = s_node[hi];
          const bool up = (lo & k) == 0;
          if ((x > y) == up) {
            s_node[lo] = y;
            s_node[hi] = x;
          }
        }
        __syncthreads();
      }
    for (int e = tid; e < total; e += ST::THREADS) {
      const int node = s_node[e];
#pragma unroll
      for (int ax = 0; ax < DIM; ++ax) s_corner[e][ax] = (int)nearest[(int64_t)(DIM - 1 - ax) * n_lag + node] - 1;
      for (int c = 0; c < eul.ncomp; ++c) s_force[e][c] = lag[c * lag_sc + node];
    }
    __syncthreads();
    // this thread's cell
    int cell[3] = {0, 0, 0};
    {
      int r = tid;
#pragma unroll
      for (int ax = DIM - 1; ax >= 0; --ax) {
        cell[ax] = tc[ax] * TS + r % TS;
        r /= TS;
      }
    }
    bool inside = true;
    int64_t o = 0;
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) {
      if (cell[ax] >= eul.n[ax]) inside = false;
      o += (int64_t)cell[ax] * eul.s[ax];
    }
    if (inside) {
      T acc[3] = {T(0), T(0), T(0)};
      bool hit = false;
      for (int m = 0; m < total; ++m) {
        int tap = 0;
        bool cover = true;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
          const int d = cell[ax] - s_corner[m][ax];
          if (d < 0 || d > 3) cover = false;
          tap = tap * 4 + d;  // (kz*4 + ky)*4 + kx: array-axis order, like tap_offsets
        }
        if (cover) {
          const T w = weights[(int64_t)tap * n_lag + s_node[m]];
          for (int c = 0; c < eul.ncomp; ++c) acc[c] += s_force[m][c] * w;
          hit = true;
        }
      }
      if (hit) {
        T* ep = reinterpret_cast<T*>(eul.p);
        for (int c = 0; c < eul.ncomp; ++c) ep[c * eul.sc + o] += acc[c];
      }
    }
  }
}

static bool spread_tiles_enabled() {
  static const int v = [] {
    const char* e = getenv("SOPHT_IB_SPREAD");
    return !(e && e[0] == 'a');  // "atomics": one warp per node with red.global.add (the round-1 path)
  }();
  return v != 0;
}

template <typename T, int DIM>
static int spread_tiles(const EulView& ev, const T* lag, int64_t lag_sc, const T* weights, const int64_t* nearest,
                        int64_t n, cudaStream_t st) {
  using ST = SpreadTile<DIM>;
  static_assert((ST::MAXN & (ST::MAXN - 1)) == 0, "bitonic sort length");
  TileGrid tg{{1, 1, 1}};
  int64_t ntiles = 1;
  for (int ax = 0; ax < DIM; ++ax) {
    tg.nt[ax] = (ev.n[ax] + ST::TS - 1) / ST::TS;
    ntiles *= tg.nt[ax];
  }
  if (ntiles > 0x7fffffff) SOPHT_FAIL(SOPHT_ERR_SHAPE, "ib spread: grid too large");
  SpreadWork* w;
  int rc = spread_workspace(ntiles, ST::CAP, &w);
  if (rc) return rc;
  int* cur = w->counts;
  int* flags = cur + ntiles;
  int* nactive = cur + 2 * ntiles;
  SOPHT_CUDA(cudaMemsetAsync(cur, 0, sizeof(int) * (2 * ntiles + 1), st));
  {
    SOPHT_PROF("ib.spread_bin", st);
    ib_bin_kernel<T, DIM><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ev, tg, lag, lag_sc, weights, nearest, n, cur,
                                                                      w->lists, w->stragglers, flags, w->active,
                                                                      nactive);
    SOPHT_CHECK_LAUNCH();
  }
  SOPHT_PROF("ib.spread_tiles", st);
  const int64_t blocks = ntiles < 148 * 4 ? ntiles : 148 * 4;
  ib_tile_gather_kernel<T, DIM><<<(unsigned)blocks, ST::THREADS, 0, st>>>(ev, tg, lag, lag_sc, weights, nearest, n, cur,
                                                                         w->lists, w->active, nactive);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

// ---- host side ---------------------------------------------------------------------------------------
static int make_eul_view(const char* fn, EulView* v, const sopht_field_t* f, int dim) {
  if (!valid_field(f, dim, dim + 1))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: Eulerian field must be a %d-D scalar or vector grid field", fn, dim);
  const bool vec = f->ndim == dim + 1;
  v->p = f->data;
  v->ncomp = vec ? (int)f->shape[0] : 1;
  if (vec && v->ncomp != dim)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: vector field must have %d components", fn, dim);
  v->sc = vec ? f->stride[0] : 0;
  const int o = vec ? 1 : 0;
  for (int a = 0; a < 3; ++a) {
    v->n[a] = 1;
    v->s[a] = 0;
  }
  for (int a = 0; a < dim; ++a) {
    if (f->shape[o + a] > 0x7fffffff) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large", fn);
    v->n[a] = (int)f->shape[o + a];
    v->s[a] = f->stride[o + a];
  }
  return SOPHT_OK;
}

static int warp_grid(int64_t n_lag) {
  int64_t blocks = (n_lag * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int flat_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static bool lag_matrix_ok(const sopht_field_t* f, int dim, int64_t n) {
  return valid_field(f, 2, 2) && f->shape[0] == dim && f->shape[1] == n;
}

}  // namespace sopht

using namespace sopht;

#define RETURN_IF(rc_expr) \
  do {                     \
    int rc__ = (rc_expr);  \
    if (rc__) return rc__; \
  } while (0)

#define DISPATCH_DIM_T(dim, dtype, CALL)          \
  do {                                            \
    if ((dim) == 3) {                             \
      if ((dtype) == SOPHT_F32) { CALL(float, 3); } else { CALL(double, 3); } \
    } else {                                      \
      if ((dtype) == SOPHT_F32) { CALL(float, 2); } else { CALL(double, 2); } \
    }                                             \
  } while (0)

extern "C" {

int sopht_ib_local_support(int dtype, int dim, const sopht_field_t* local_support,
                           const sopht_field_t* nearest_index, const sopht_field_t* lag_positions,
                           int pos_dtype, double dx, double eul_grid_coord_shift, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  SOPHT_CHECK_DTYPE(pos_dtype);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (!valid_field(lag_positions, 2, 2) || lag_positions->shape[0] != dim)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: lag_positions must be (dim, N)", __func__);
  const int64_t n = lag_positions->shape[1];
  if (!lag_matrix_ok(nearest_index, dim, n) || !is_contiguous(nearest_index))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: nearest index must be a contiguous (dim, N) int64 array", __func__);
  if (!valid_field(local_support, dim + 2, dim + 2) || !is_contiguous(local_support) ||
      local_support->shape[0] != dim || local_support->shape[dim + 1] != n)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: local support must be a contiguous (dim, 4, .., 4, N) array", __func__);
  for (int a = 1; a <= dim; ++a)
    if (local_support->shape[a] != 4)
      SOPHT_FAIL(SOPHT_ERR_ARG, "%s: interpolation kernel width must be 2 (4 taps per axis)", __func__);
  if (n == 0) return SOPHT_OK;
  cudaStream_t st = as_stream(stream);
  const int nt = dim == 3 ? 64 : 16;
  const int grid = flat_grid(n * nt);
#define CALL(T, D)                                                                                      \
  if (pos_dtype == SOPHT_F64)                                                                           \
    ib_support_kernel<T, double, D><<<grid, 256, 0, st>>>(                                              \
        (T*)local_support->data, (int64_t*)nearest_index->data, (const double*)lag_positions->data,     \
        lag_positions->stride[0], lag_positions->stride[1], n, dx, eul_grid_coord_shift, dx,            \
        eul_grid_coord_shift);                                                                          \
  else                                                                                                  \
    ib_support_kernel<T, float, D><<<grid, 256, 0, st>>>(                                               \
        (T*)local_support->data, (int64_t*)nearest_index->data, (const float*)lag_positions->data,      \
        lag_positions->stride[0], lag_positions->stride[1], n, (float)dx, (float)eul_grid_coord_shift,  \
        dx, eul_grid_coord_shift);
  DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_ib_interpolation_weights(int dtype, int dim, int kernel_type,
                                   const sopht_field_t* interp_weights,
                                   const sopht_field_t* local_support, double dx, double prefactor,
                                   void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (kernel_type != 0 && kernel_type != 1)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: kernel type must be 0 (cosine) or 1 (peskin)", __func__);
  if (!valid_field(interp_weights, dim + 1, dim + 1) || !is_contiguous(interp_weights) ||
      !valid_field(local_support, dim + 2, dim + 2) || !is_contiguous(local_support) ||
      local_support->shape[0] != dim)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected contiguous (4,..,4,N) weights and (dim,4,..,4,N) support", __func__);
  for (int a = 0; a <= dim; ++a)
    if (interp_weights->shape[a] != local_support->shape[a + 1] || (a < dim && interp_weights->shape[a] != 4))
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: weights / support shapes inconsistent", __func__);
  const int64_t n = interp_weights->shape[dim];
  if (n == 0) return SOPHT_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = flat_grid(n * (dim == 3 ? 64 : 16));
#define CALL(T, D)                                                                                   \
  if (kernel_type == 0)                                                                              \
    ib_weights_kernel<T, D, 0><<<grid, 256, 0, st>>>((T*)interp_weights->data, (T*)local_support->data, n, \
                                                     (T)dx, prefactor);                              \
  else                                                                                               \
    ib_weights_kernel<T, D, 1><<<grid, 256, 0, st>>>((T*)interp_weights->data, (T*)local_support->data, n, \
                                                     (T)dx, prefactor);
  DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

static int check_transfer_args(const char* fn, int dim, const sopht_field_t* lag, const EulView& ev,
                               const sopht_field_t* weights, const sopht_field_t* nearest, int64_t* n_out,
                               int64_t* lag_sc) {
  if (!valid_field(weights, dim + 1, dim + 1) || !is_contiguous(weights))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: weights must be a contiguous (4,..,4,N) array", fn);
  const int64_t n = weights->shape[dim];
  for (int a = 0; a < dim; ++a)
    if (weights->shape[a] != 4) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: interpolation kernel width must be 2", fn);
  if (!lag_matrix_ok(nearest, dim, n) || !is_contiguous(nearest))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: nearest index must be a contiguous (dim, N) int64 array", fn);
  if (ev.ncomp == 1) {
    if (!valid_field(lag, 1, 2) || lag->shape[lag->ndim - 1] != n || !(lag->ndim == 1 || lag->shape[0] == 1) ||
        lag->stride[lag->ndim - 1] != 1)
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: scalar Lagrangian field must be (N,)", fn);
    *lag_sc = 0;
  } else {
    if (!lag_matrix_ok(lag, dim, n) || lag->stride[1] != 1)
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: vector Lagrangian field must be (dim, N)", fn);
    *lag_sc = lag->stride[0];
  }
  *n_out = n;
  return SOPHT_OK;
}

int sopht_ib_eulerian_to_lagrangian(int dtype, int dim, const sopht_field_t* lag_grid_field,
                                    const sopht_field_t* eul_grid_field,
                                    const sopht_field_t* interp_weights,
                                    const sopht_field_t* nearest_index, double dx_pow_dim, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  EulView ev;
  RETURN_IF(make_eul_view(__func__, &ev, eul_grid_field, dim));
  int64_t n, lag_sc;
  RETURN_IF(check_transfer_args(__func__, dim, lag_grid_field, ev, interp_weights, nearest_index, &n, &lag_sc));
  if (n == 0) return SOPHT_OK;
  cudaStream_t st = as_stream(stream);
#define CALL(T, D)                                                                             \
  ib_gather_kernel<T, D><<<warp_grid(n), 256, 0, st>>>((T*)lag_grid_field->data, lag_sc, ev,   \
                                                       (const T*)interp_weights->data,         \
                                                       (const int64_t*)nearest_index->data, n, (T)dx_pow_dim);
  DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_ib_lagrangian_to_eulerian(int dtype, int dim, const sopht_field_t* eul_grid_field,
                                    const sopht_field_t* lag_grid_field,
                                    const sopht_field_t* interp_weights,
                                    const sopht_field_t* nearest_index, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  EulView ev;
  RETURN_IF(make_eul_view(__func__, &ev, eul_grid_field, dim));
  int64_t n, lag_sc;
  RETURN_IF(check_transfer_args(__func__, dim, lag_grid_field, ev, interp_weights, nearest_index, &n, &lag_sc));
  if (n == 0) return SOPHT_OK;
  cudaStream_t st = as_stream(stream);
  if (spread_tiles_enabled()) {
#define CALL(T, D)                                                                                    \
  return spread_tiles<T, D>(ev, (const T*)lag_grid_field->data, lag_sc, (const T*)interp_weights->data, \
                            (const int64_t*)nearest_index->data, n, st);
    DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
  }
#define CALL(T, D)                                                                                \
  ib_spread_kernel<T, D><<<warp_grid(n), 256, 0, st>>>(ev, (const T*)lag_grid_field->data, lag_sc, \
                                                       (const T*)interp_weights->data,            \
                                                       (const int64_t*)nearest_index->data, n);
  DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

/* running count of Lagrangian nodes that did not fit their tile's list in the privatised spread and fell back to
 * atomics (0 for every body in the reference's examples); synchronises. */
int sopht_ib_spread_stragglers(unsigned long long* count_out) {
  if (!count_out) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null pointer", __func__);
  *count_out = 0;
  if (g_spread_work.stragglers)
    SOPHT_CUDA(cudaMemcpy(count_out, g_spread_work.stragglers, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return SOPHT_OK;
}

int sopht_ib_virtual_boundary_forcing(int dtype, int dim, const sopht_field_t* eul_grid_forcing_field,
                                      const sopht_field_t* eul_grid_velocity_field,
                                      const sopht_field_t* lag_positions,
                                      const sopht_field_t* lag_body_velocity, int pos_dtype,
                                      const sopht_field_t* local_support,
                                      const sopht_field_t* interp_weights,
                                      const sopht_field_t* nearest_index,
                                      const sopht_field_t* lag_flow_velocity,
                                      const sopht_field_t* lag_velocity_mismatch,
                                      const sopht_field_t* lag_position_mismatch,
                                      const sopht_field_t* lag_forcing, double dx,
                                      double eul_grid_coord_shift, double weight_prefactor,
                                      double dx_pow_dim, double stiffness, double damping, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  SOPHT_CHECK_DTYPE(pos_dtype);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  EulView vv, fv;
  RETURN_IF(make_eul_view(__func__, &vv, eul_grid_velocity_field, dim));
  if (vv.ncomp != dim) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: velocity must be a vector field", __func__);
  fv.p = nullptr;
  if (eul_grid_forcing_field) {
    RETURN_IF(make_eul_view(__func__, &fv, eul_grid_forcing_field, dim));
    if (fv.ncomp != dim || fv.n[0] != vv.n[0] || fv.n[1] != vv.n[1] || fv.n[2] != vv.n[2])
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: forcing and velocity fields must have the same shape", __func__);
  }
  if (!valid_field(lag_positions, 2, 2) || lag_positions->shape[0] != dim)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: lag_positions must be (dim, N)", __func__);
  const int64_t n = lag_positions->shape[1];
  if (!lag_matrix_ok(lag_body_velocity, dim, n))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: body velocity must be (dim, N)", __func__);
  const sopht_field_t* mats[5] = {nearest_index, lag_flow_velocity, lag_velocity_mismatch,
                                  lag_position_mismatch, lag_forcing};
  for (const sopht_field_t* m : mats)
    if (!lag_matrix_ok(m, dim, n) || !is_contiguous(m))
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: Lagrangian buffers must be contiguous (dim, N)", __func__);
  if (!valid_field(local_support, dim + 2, dim + 2) || !is_contiguous(local_support) ||
      local_support->shape[0] != dim || local_support->shape[dim + 1] != n ||
      !valid_field(interp_weights, dim + 1, dim + 1) || !is_contiguous(interp_weights) ||
      interp_weights->shape[dim] != n)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: support / weights buffers have the wrong shape", __func__);
  for (int a = 0; a < dim; ++a)
    if (interp_weights->shape[a] != 4 || local_support->shape[a + 1] != 4)
      SOPHT_FAIL(SOPHT_ERR_ARG, "%s: interpolation kernel width must be 2", __func__);
  if (n == 0) return SOPHT_OK;
  VbfArgs a;
  a.support = local_support->data;
  a.weights = interp_weights->data;
  a.nearest = (int64_t*)nearest_index->data;
  a.flow_vel = lag_flow_velocity->data;
  a.vel_mismatch = lag_velocity_mismatch->data;
  a.pos_mismatch = lag_position_mismatch->data;
  a.forcing = lag_forcing->data;
  a.pos = lag_positions->data;
  a.body_vel = lag_body_velocity->data;
  a.pos_sd = lag_positions->stride[0];
  a.pos_sn = lag_positions->stride[1];
  a.bv_sd = lag_body_velocity->stride[0];
  a.bv_sn = lag_body_velocity->stride[1];
  a.n_lag = n;
  a.dx = dx;
  a.shift = eul_grid_coord_shift;
  a.prefactor = weight_prefactor;
  a.vol = dx_pow_dim;
  a.stiffness = stiffness;
  a.damping = damping;
  cudaStream_t st = as_stream(stream);
  const bool tiles = fv.p != nullptr && spread_tiles_enabled();
  EulView fv_kernel = fv;
  if (tiles) fv_kernel.p = nullptr;  // forces only; the spread is the privatised tile gather below
  {
    SOPHT_PROF("ib.virtual_boundary_fused", st);
#define CALL(T, D)                                                                      \
  if (pos_dtype == SOPHT_F64)                                                           \
    vbf_fused_kernel<T, double, D><<<warp_grid(n), 256, 0, st>>>(a, vv, fv_kernel);     \
  else                                                                                  \
    vbf_fused_kernel<T, float, D><<<warp_grid(n), 256, 0, st>>>(a, vv, fv_kernel);
    DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
    SOPHT_CHECK_LAUNCH();
  }
  if (tiles) {
#define CALL(T, D)                                                                                         \
  return spread_tiles<T, D>(fv, (const T*)lag_forcing->data, n, (const T*)interp_weights->data, a.nearest, n, st);
    DISPATCH_DIM_T(dim, dtype, CALL);
#undef CALL
  }
  return SOPHT_OK;
}

}  // extern "C"
