// Unbounded Poisson solve, fp32 power-of-two fast path: a pruned, zero-padding-free FFT pipeline of five
// hand-written shared-memory FFT kernels (no cuFFT on the solve path). See poisson_pow2_phases.cuh for the
// dataflow; this file holds the kernel launcher, the dispatch over transform lengths and the handle.
//
// HBM traffic per cell per component: x 4+8, y 8+16, z 16+16 (+ G_hat, folded by its even
// symmetry and reused across the three components), y^-1 16+8, x^-1 8+4  = 104 B + G_hat.
//
// ref: sopht/numeric/eulerian_grid_ops/poisson_solver_3d/UnboundedPoissonSolverPYFFTW3D.py:111-172
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "poisson.cuh"
#include "poisson_pow2_phases.cuh"
#include "poisson_zrow.cuh"

namespace sopht {

namespace {

// Barrier of a kernel's synchronisation group (K::SYNC_THREADS consecutive threads: the whole CTA, a named
// barrier per group of warps, or a single warp). Groups share nothing but the read-only tables.
template <class K>
__device__ __forceinline__ void group_sync() {
  if constexpr (K::SYNC_THREADS >= K::THREADS) {
    __syncthreads();
  } else if constexpr (K::SYNC_THREADS == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x / K::SYNC_THREADS) + 1), "n"(K::SYNC_THREADS)
                 : "memory");
  }
}

// phases 1..P of kernel K, a barrier between consecutive phases (phase 0 is issued by the kernel loop)
template <class K, int P>
struct DevPhases {
  __device__ __forceinline__ static void run(const typename K::Params& p, int bx, int by, int it,
                                             float2* smem, const float2* stage) {
    DevPhases<K, P - 1>::run(p, bx, by, it, smem, stage);
    if (P > 1) group_sync<K>();
    K::template phase<P>(p, bx, by, it, threadIdx.x, smem, stage);
  }
};
template <class K>
struct DevPhases<K, 0> {
  __device__ __forceinline__ static void run(const typename K::Params&, int, int, int, float2*, const float2*) {}
};

// Two tuning knobs per kernel family, chosen from measurements on B200 (profiles/): MINB = CTAs per SM the
// register allocation is bounded for, STAGED = cp.async prefetch of the next tile's first-phase inputs into a
// staging buffer. Staging pays when only one CTA fits an SM; with two resident CTAs the hardware scheduler
// overlaps one CTA's loads with the other's butterflies.
template <class K>
constexpr bool stage_fits(int ctas) {
  return sizeof(float2) * (size_t)(K::SMEM_ELEMS + K::EXTRA_ELEMS + K::STAGE_ELEMS) * ctas <= 226 * 1024;
}
template <class K>
constexpr int auto_min_ctas() {
  return 512 / K::THREADS > 1 ? 512 / K::THREADS : 1;
}

// Persistent kernel: CTA b owns tiles b, b + gridDim.x, ... (neighbouring CTAs work on neighbouring tiles at
// the same time, which keeps DRAM pages shared); all iterations (components) of a tile stay on one CTA.
// Tile s -> (bx, by): 2^zb_shift consecutive by first, then bx, then the remaining by (zb_shift = 0: bx fastest).
// The y inverse uses it: neighbouring by (z planes) are adjacent 64-byte segments of the kx-tile-major spectrum
// it reads, neighbouring bx adjacent segments of the x-major spectrum it writes.
struct TileOrder {
  int gx, zb_shift;
  __device__ __forceinline__ void operator()(int64_t s, int& bx, int& by) const {
    const int64_t q = s >> zb_shift;
    bx = (int)(q % gx);
    by = (int)((q / gx) << zb_shift) + (int)(s & ((1 << zb_shift) - 1));
  }
};

template <class K, int MINB, bool STAGED>
__global__ void __launch_bounds__(K::THREADS, MINB)
    p2_kernel(const typename K::Params p, int gx, int gy, int zb_shift) {
  extern __shared__ float2 p2_smem[];
  float2* stage = STAGED ? p2_smem + K::SMEM_ELEMS + K::EXTRA_ELEMS : nullptr;
  const int niter = K::niter(p);
  const int64_t ntile = (int64_t)gx * gy;
  int64_t s = blockIdx.x;
  int it = 0;
  if (s >= ntile) return;
  K::init(p, threadIdx.x, p2_smem);  // shared-memory tables (twiddles are read from phase 0 on)
  if (K::EXTRA_ELEMS) __syncthreads();
  pdl_wait();  // everything below reads what the previous kernel of the stream wrote
  const TileOrder order{gx, zb_shift};
  int bx, by;
  order(s, bx, by);
  if (STAGED) K::prefetch(p, bx, by, 0, threadIdx.x, stage);
  while (true) {
    if (STAGED) {
      fft::async_commit_wait_all();
      if (K::STAGE_SHARED) __syncthreads();
    }
    order(s, bx, by);
    int64_t ns = s;
    int nit = it + 1;
    if (nit == niter) {
      nit = 0;
      ns = s + gridDim.x;
    }
    const bool has_next = ns < ntile;
    if (!has_next) pdl_launch_dependents();  // last tile of this CTA: the next kernel may start filling the tail
    K::template phase<0>(p, bx, by, it, threadIdx.x, p2_smem, stage);
    group_sync<K>();
    if (has_next) {
      int nbx, nby;
      order(ns, nbx, nby);
      if (STAGED) K::prefetch(p, nbx, nby, nit, threadIdx.x, stage);
      if (!STAGED && K::L2_PREFETCH) K::l2_prefetch(p, nbx, nby, nit, threadIdx.x);
    }
    DevPhases<K, K::NPHASE - 1>::run(p, bx, by, it, p2_smem, stage);
    if (!has_next) break;
    group_sync<K>();
    s = ns;
    it = nit;
  }
}

template <class K, int MINB, bool STAGED>
int launch_variant(const typename K::Params& p, dim3 grid, const char* label, cudaStream_t st, int zb_shift) {
  const size_t smem = sizeof(float2) * (K::SMEM_ELEMS + K::EXTRA_ELEMS + (STAGED ? K::STAGE_ELEMS : 0));
  static int ctas_per_sm = 0, num_sm = 0;  // per kernel instantiation
  if (!ctas_per_sm) {
    if (smem > 48 * 1024)
      SOPHT_CUDA(cudaFuncSetAttribute(p2_kernel<K, MINB, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    int dev = 0;
    SOPHT_CUDA(cudaGetDevice(&dev));
    SOPHT_CUDA(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
    SOPHT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, p2_kernel<K, MINB, STAGED>,
                                                             K::THREADS, smem));
    if (ctas_per_sm < 1) SOPHT_FAIL(SOPHT_ERR_CUDA, "poisson(pow2): kernel does not fit on an SM");
  }
  const int64_t ntile = (int64_t)grid.x * grid.y;
  int64_t g = (int64_t)num_sm * ctas_per_sm;
  if (g > ntile) g = ntile;
  SOPHT_PROF(label, st);
  while (zb_shift > 0 && (grid.y & ((1u << zb_shift) - 1))) --zb_shift;
  SOPHT_CUDA(launch_pdl(p2_kernel<K, MINB, STAGED>, dim3((unsigned)g), dim3(K::THREADS), smem, st, p, (int)grid.x,
                        (int)grid.y, zb_shift));
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

// experiment switches (SOPHT_P2_MINB = 1 | 2, SOPHT_P2_STAGE = 0 | 1); unset = the tuned default
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <class K>
int launch(const typename K::Params& p, dim3 grid, const char* label, cudaStream_t st, int zb_shift = 0) {
  constexpr int AUTO = auto_min_ctas<K>();
  static const int minb = env_int("SOPHT_P2_MINB", 0);
  static const int stg = env_int("SOPHT_P2_STAGE", -1);
  if constexpr (!K::WANT_STAGE) {
    return launch_variant<K, AUTO, false>(p, grid, label, st, zb_shift);
  } else {
  // big radix-32 column kernels: 256 threads per CTA
  // measured defaults (512^3, L = 1024; profiles/): the y passes run best as two 128-register CTAs per SM
  // without staging, the fused z pass as one 255-register CTA with the next tile's inputs staged
  const bool two = minb ? minb >= 2 : K::DEFAULT_TWO_CTAS;
  bool staged = stg >= 0 ? stg != 0 : K::DEFAULT_STAGED;
  if (two) {
    if (staged && stage_fits<K>(AUTO)) return launch_variant<K, AUTO, true>(p, grid, label, st, zb_shift);
    return launch_variant<K, AUTO, false>(p, grid, label, st, zb_shift);
  }
  if (staged && stage_fits<K>(1)) return launch_variant<K, 1, true>(p, grid, label, st, zb_shift);
  return launch_variant<K, 1, false>(p, grid, label, st, zb_shift);
  }
}

constexpr int TX = 8;  // columns per CTA in the y / z passes (8 complex = 64 B segments)

template <int L>
constexpr int rows_per_cta() {  // x passes: aim for 128..256 threads per CTA
  return fft::Cfg<L>::T >= 128 ? 1 : (128 / fft::Cfg<L>::T < 1 ? 1 : 128 / fft::Cfg<L>::T);
}

#define P2_SWITCH_L(Lval, MACRO)                                   \
  switch (Lval) {                                                  \
    case 16: MACRO(16); break;                                     \
    case 32: MACRO(32); break;                                     \
    case 64: MACRO(64); break;                                     \
    case 128: MACRO(128); break;                                   \
    case 256: MACRO(256); break;                                   \
    case 512: MACRO(512); break;                                   \
    case 1024: MACRO(1024); break;                                 \
    case 2048: MACRO(2048); break;                                 \
    default: SOPHT_FAIL(SOPHT_ERR_SHAPE, "poisson(pow2): unsupported transform length %d", (int)(Lval)); \
  }

int launch_xfwd(int L, const p2::XParams& p, int64_t rows, cudaStream_t st) {
#define M(LL)                                                                         \
  {                                                                                   \
    constexpr int RX = rows_per_cta<LL>();                                            \
    return launch<p2::XFwd<LL, RX>>(p, dim3((unsigned)(rows / RX), 1, 1), "poisson.x_fwd", st);     \
  }
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
// With a peer spectrum (slab solve, pull transpose) the next tile's rows are staged with cp.async while the current
// tile is transformed, so the NVLink reads overlap the butterflies instead of preceding them.
int launch_xinv(int L, const p2::XParams& p, int64_t rows, cudaStream_t st) {
#define M(LL)                                                                         \
  {                                                                                   \
    constexpr int RX = rows_per_cta<LL>();                                            \
    using K = p2::XInv<LL, RX>;                                                       \
    const dim3 grid((unsigned)(rows / RX), 1, 1);                                     \
    if (p.peer_read && stage_fits<K>(1))                                              \
      return launch_variant<K, auto_min_ctas<K>(), true>(p, grid, "poisson.x_inv", st, 0); \
    return launch<K>(p, grid, "poisson.x_inv", st);                                   \
  }
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
// y inverse: 2^shift consecutive z planes of one kx tile run on neighbouring CTAs (TileOrder)
int y_zb_shift() {
  static const int v = env_int("SOPHT_P2_ZB_SHIFT", 3);
  return v;
}

int launch_yfwd(int L, const p2::ColParams& p, dim3 grid, cudaStream_t st) {
#define M(LL) return launch<p2::YFwd<LL, TX>>(p, grid, grid.y > 1 ? "poisson.y_fwd" : "poisson.y_fwd.nyquist", st);
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
int launch_yinv(int L, const p2::ColParams& p, dim3 grid, cudaStream_t st) {
#define M(LL) \
  return launch<p2::YInv<LL, TX>>(p, grid, grid.y > 1 ? "poisson.y_inv" : "poisson.y_inv.nyquist", st, y_zb_shift());
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
int launch_zconv(int L, const p2::ZParams& p, dim3 grid, cudaStream_t st) {
#define M(LL) return launch<p2::ZConv<LL, TX>>(p, grid, p.nyq ? "poisson.z_conv.nyquist" : "poisson.z_conv", st);
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}

// ---- periodic box: the same kernels with nothing padded and nothing dropped (FULL), z pass with the symbol -------------
int launch_xfwd_full(int L, const p2::XParams& p, int64_t rows, cudaStream_t st) {
#define M(LL)                                                                                          \
  {                                                                                                    \
    constexpr int RX = rows_per_cta<LL>();                                                             \
    return launch<p2::XFwd<LL, RX, true>>(p, dim3((unsigned)(rows / RX), 1, 1), "poisson.x_fwd", st);  \
  }
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
int launch_xinv_full(int L, const p2::XParams& p, int64_t rows, cudaStream_t st) {
#define M(LL)                                                                                   \
  {                                                                                             \
    constexpr int RX = rows_per_cta<LL>();                                                      \
    using K = p2::XInv<LL, RX, true>;                                                           \
    const dim3 grid((unsigned)(rows / RX), 1, 1);                                               \
    if (p.peer_read && stage_fits<K>(1))                                                        \
      return launch_variant<K, auto_min_ctas<K>(), true>(p, grid, "poisson.x_inv", st, 0);      \
    return launch<K>(p, grid, "poisson.x_inv", st);                                             \
  }
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
int launch_yfwd_full(int L, const p2::ColParams& p, dim3 grid, cudaStream_t st) {
#define M(LL) \
  return launch<p2::YFwd<LL, TX, true>>(p, grid, grid.y > 1 ? "poisson.y_fwd" : "poisson.y_fwd.nyquist", st);
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
int launch_yinv_full(int L, const p2::ColParams& p, dim3 grid, cudaStream_t st) {
#define M(LL) \
  return launch<p2::YInv<LL, TX, true>>(p, grid, grid.y > 1 ? "poisson.y_inv" : "poisson.y_inv.nyquist", st);
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}
// columns per CTA of the periodic z pass: its rows lie ny * nx/2 * 8 bytes apart and are read AND written in place, so
// wider rows (16 columns = one 128-byte line) pay where shared memory allows it (L <= 1024)
int zsym_tx(int L, int ncols) {
  static const int want = env_int("SOPHT_ZSYM_TX", 16);
  return want == 16 && L <= 1024 && ncols % 16 == 0 ? 16 : TX;
}
int launch_zsym(int L, const p2::ZSymParams& p, dim3 grid, cudaStream_t st, int tx = TX) {
  const char* label = p.nyq ? "poisson.z_sym.nyquist" : "poisson.z_sym";
  if (tx == 16) {
    switch (L) {
      case 16: return launch<p2::ZSym<16, 16>>(p, grid, label, st);
      case 32: return launch<p2::ZSym<32, 16>>(p, grid, label, st);
      case 64: return launch<p2::ZSym<64, 16>>(p, grid, label, st);
      case 128: return launch<p2::ZSym<128, 16>>(p, grid, label, st);
      case 256: return launch<p2::ZSym<256, 16>>(p, grid, label, st);
      case 512: return launch<p2::ZSym<512, 16>>(p, grid, label, st);
      case 1024: return launch<p2::ZSym<1024, 16>>(p, grid, label, st);
      default: break;
    }
  }
#define M(LL) return launch<p2::ZSym<LL, TX>>(p, grid, label, st);
  P2_SWITCH_L(L, M)
#undef M
  return SOPHT_OK;
}

// folded spectrum: gm[(fz*(ny+1) + fy)*nx + kx] (kx < nx), gn[fz*(ny+1) + fy] (kx = nx); x2 because the
// half-length x transform's unnormalised round trip is nx * 2ny * 2nz, half the doubled cell count.
// gm keeps the kx range [kx0, kx0 + g_row) only (the whole spectrum on one GPU, a rank's slice otherwise)
__global__ void __launch_bounds__(256)
    fold_green_kernel(float* gm, float* gn, const float* ghat, int nz, int ny, int nx, int kx0, int g_row) {
  const int64_t total = (int64_t)(nz + 1) * (ny + 1) * (nx + 1);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int kx = (int)(q % (nx + 1));
    const int64_t r = q / (nx + 1);
    const int fy = (int)(r % (ny + 1)), fz = (int)(r / (ny + 1));
    const float v = 2.0f * ghat[((int64_t)fz * 2 * ny + fy) * (nx + 1) + kx];
    if (kx == nx)
      gn[(int64_t)fz * (ny + 1) + fy] = v;
    else if (kx >= kx0 && kx < kx0 + g_row)
      gm[((int64_t)fz * (ny + 1) + fy) * g_row + (kx - kx0)] = v;
  }
}

bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// ---- row-mode z pass (poisson_zrow.cuh) -------------------------------------------------------------------------
// tile-major copy of the folded G_hat for it, columns in the consumer order of p2::ZRow (poisson_zrow.cuh):
//   gt[(((fy * ntx + kxt) * TX + col) * (K2/4) + kq) * (NBLK+1) * 4 + r * 4 + i] = gm[fz = r + NBLK (4 kq + i)][fy][kx]
__global__ void __launch_bounds__(256)
    green_tiles_kernel(float* gt, const float* gm, int nz, int ny, int g_row, int NBLK, int K2) {
  const int GP = (K2 / 4) * (NBLK + 1) * 4;
  const int64_t total = (int64_t)(ny + 1) * g_row * GP;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(q % GP);
    const int64_t rr = q / GP;      // (fy * ntx + kxt) * TX + col = fy * g_row + kx
    const int kx = (int)(rr % g_row), fy = (int)(rr / g_row);
    const int i = e & 3, r = (e >> 2) % (NBLK + 1), kq = (e >> 2) / (NBLK + 1);
    const int fz = r + NBLK * (4 * kq + i);
    gt[q] = fz <= nz ? gm[((int64_t)fz * (ny + 1) + fy) * g_row + kx] : 0.0f;
  }
}

// SOPHT_P2_ZQUAD (default 1): 2 nz = 1024 runs the warp-quartet form (poisson_zquad.cuh) instead of zrow_kernel
bool zquad_enabled() {
  static const int v = env_int("SOPHT_P2_ZQUAD", 1);
  return v != 0;
}
bool zrow_enabled() {
  static const int v = env_int("SOPHT_P2_ZROW", 1);
  return v != 0;
}
// transform lengths the row-mode kernel is built for; 0 = not eligible
int zrow_tx(int LZ, int ncomp, int nxl) {
  if (!zrow_enabled() || ncomp != 3) return 0;
  // measured on B200 (profiles/r02_zpass_experiments.txt): 2 nz = 1024: 5.99 ms against 6.32 ms for p2::ZConv at 512^3;
  // 2 nz = 512: 0.60 against 0.57 ms at 256^3 (p2::ZConv runs four 128-thread CTAs per SM there) - opt-in only
  static const int allow512 = env_int("SOPHT_P2_ZROW_512", 0);
  int tx = 0;
  if (LZ == 1024) tx = p2::ZRow<1024>::TX;
  if (LZ == 512 && allow512) tx = p2::ZRow<512>::TX;
  return tx && nxl % tx == 0 && nxl % 8 == 0 ? tx : 0;
}
// the warp-quartet kernel writes, and the y inverse pass then reads, the z-blocked form of the tile-major spectrum
// (poisson_pow2_phases.cuh: slab_y_params); SOPHT_P2_B2_Z8=0 keeps the plain form
bool b2_z_blocked(int LZ, int ncomp, int nxl) {
  static const int want = env_int("SOPHT_P2_B2_Z8", 1);
  return want && LZ == 1024 && zquad_enabled() && zrow_tx(LZ, ncomp, nxl) != 0;
}
int zrow_gp(int LZ) { return LZ == 1024 ? p2::ZRow<1024>::GP : p2::ZRow<512>::GP; }

int build_green_tiles(float** gt, const float* gm, int nz, int ny, int g_row, cudaStream_t st) {
  const int LZ = 2 * nz;
  if (!zrow_tx(LZ, 3, g_row)) return SOPHT_OK;
  const int GP = zrow_gp(LZ);
  if (cudaMalloc(gt, sizeof(float) * (size_t)(ny + 1) * g_row * GP) != cudaSuccess)
    SOPHT_FAIL(SOPHT_ERR_ALLOC, "poisson(pow2): out of device memory for the tile-major Green's function");
  const int NBLK = LZ == 1024 ? p2::ZRow<1024>::NBLK : p2::ZRow<512>::NBLK;
  const int K2 = LZ == 1024 ? p2::ZRow<1024>::K2 : p2::ZRow<512>::K2;
  green_tiles_kernel<<<148 * 8, 256, 0, st>>>(*gt, gm, nz, ny, g_row, NBLK, K2);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

template <int L>
int launch_zrow_L(const p2::ZRowParams& p, int nunits, cudaStream_t st) {
  using K = p2::ZRow<L>;
  if (L == 1024 && zquad_enabled()) return launch_zquad(p, nunits, st);
  static int num_sm = 0;
  if (!num_sm) {
    SOPHT_CUDA(cudaFuncSetAttribute(p2::zrow_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)K::SMEM_BYTES));
    int dev = 0;
    SOPHT_CUDA(cudaGetDevice(&dev));
    SOPHT_CUDA(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const int g = nunits < num_sm ? nunits : num_sm;
  SOPHT_PROF("poisson.z_conv", st);
  SOPHT_CUDA(launch_pdl(p2::zrow_kernel<L>, dim3(g), dim3(K::THREADS), K::SMEM_BYTES, st, p, nunits));
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}
// b: x-major (C, nz, 2ny, nxl) -> b2: kx-tile(8)-major, like slab_z_params
int launch_zrow(const p2::SlabDims& d, float2* b, float2* b2, const float* gt, const float2* tw, cudaStream_t st) {
  const int64_t nxl = d.nxl(), LY = 2 * d.ny;
  const int LZ = 2 * d.nz, tx = zrow_tx(LZ, d.C, (int)nxl);
  p2::ZRowParams p{};
  p.in = b;
  p.rs = LY * nxl, p.d_c = (int64_t)d.nz * LY * nxl, p.d_by = nxl;
  p.out = b2;
  p.o_by = (int64_t)d.nz * 8, p.o_bx8 = LY * p.o_by, p.o_c = LY * d.nz * nxl, p.o_bz8 = 64;
  if (b2_z_blocked(LZ, d.C, (int)nxl)) p.o_by = 64, p.o_bz8 = LY * 64;  // (C, nxl/8, nz/8, 2ny, 8 z, 8 kx)
  p.gt = gt;
  p.ntx = (int)(nxl / tx);
  p.n2y = (int)LY;
  p.tw = tw;
  const int nunits = (int)(p.ntx * LY);
  return LZ == 1024 ? launch_zrow_L<1024>(p, nunits, st) : launch_zrow_L<512>(p, nunits, st);
}

struct Pow2Poisson : PoissonImpl {
  int nz, ny, nx;
  double dx;
  std::vector<double> mz, my, mx;
  double origin;
  float* ghat_natural = nullptr;  // (2nz, 2ny, nx+1), kept for sopht_poisson_green_hat
  float *gm = nullptr, *gn = nullptr;
  float* gt = nullptr;  // tile-major copy of gm for the row-mode z pass (null: not eligible)
  float2 *A = nullptr, *nyqA = nullptr, *B = nullptr, *B2 = nullptr, *nyqB = nullptr;
  float2 *twx = nullptr, *twx2 = nullptr, *twy = nullptr, *twz = nullptr;
  PoissonImpl* generic = nullptr;  // built lazily for views this path cannot take (x-stride != 1, ...)
  // The kx = nx (Nyquist) plane's three small kernels form a dependency chain of latency-bound launches; they
  // run on a side stream, forked after the x forward pass and joined before the x inverse, and fill the tails
  // of the main y / z kernels (grids up to 2^25 cells; larger grids keep them on the main stream).
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  ~Pow2Poisson() override {
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    cudaFree(ghat_natural);
    cudaFree(gm);
    cudaFree(gn);
    cudaFree(gt);
    cudaFree(A);
    cudaFree(nyqA);
    cudaFree(B);
    cudaFree(B2);
    cudaFree(nyqB);
    cudaFree(twx);
    cudaFree(twx2);
    cudaFree(twy);
    cudaFree(twz);
    delete generic;
  }

  static int upload_twiddles(float2** dst, int L, int denom, cudaStream_t st) {
    std::vector<float2> h(L);
    for (int j = 0; j < L; ++j) {
      const double a = -2.0 * 3.14159265358979323846 * j / denom;
      h[j] = make_float2((float)cos(a), (float)sin(a));
    }
    SOPHT_CUDA(cudaMalloc(dst, sizeof(float2) * L));
    SOPHT_CUDA(cudaMemcpyAsync(*dst, h.data(), sizeof(float2) * L, cudaMemcpyHostToDevice, st));
    SOPHT_CUDA(cudaStreamSynchronize(st));
    return SOPHT_OK;
  }

  int init(cudaStream_t st) {
    int rc = build_green_hat<float>(&ghat_natural, 3, nz, ny, nx, dx, mz.data(), my.data(), mx.data(),
                                    origin, st);
    if (rc) return rc;
    const size_t rows = (size_t)3 * nz * ny;
    SOPHT_CUDA(cudaMalloc(&gm, sizeof(float) * (size_t)(nz + 1) * (ny + 1) * nx));
    SOPHT_CUDA(cudaMalloc(&gn, sizeof(float) * (size_t)(nz + 1) * (ny + 1)));
    fold_green_kernel<<<148 * 8, 256, 0, st>>>(gm, gn, ghat_natural, nz, ny, nx, 0, nx);
    SOPHT_CHECK_LAUNCH();
    if ((rc = build_green_tiles(&gt, gm, nz, ny, nx, st))) return rc;
    SOPHT_CUDA(cudaMalloc(&A, sizeof(float2) * rows * nx));
    SOPHT_CUDA(cudaMalloc(&nyqA, sizeof(float2) * rows));
    SOPHT_CUDA(cudaMalloc(&B, sizeof(float2) * rows * 2 * nx));
    SOPHT_CUDA(cudaMalloc(&B2, sizeof(float2) * rows * 2 * nx));
    SOPHT_CUDA(cudaMalloc(&nyqB, sizeof(float2) * rows * 2));
    if ((rc = upload_twiddles(&twx, nx, nx, st))) return rc;
    if ((rc = upload_twiddles(&twx2, nx, 2 * nx, st))) return rc;
    if ((rc = upload_twiddles(&twy, 2 * ny, 2 * ny, st))) return rc;
    if ((rc = upload_twiddles(&twz, 2 * nz, 2 * nz, st))) return rc;
    SOPHT_CUDA(cudaStreamSynchronize(st));
    int lo = 0, hi = 0;
    SOPHT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SOPHT_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
    SOPHT_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    SOPHT_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    return SOPHT_OK;
  }

  static bool view_ok(const sopht_field_t* f, int o) {
    // rows contiguous and 8-byte aligned so a real row can be moved as float2
    if (f->stride[o + 2] != 1) return false;
    if ((reinterpret_cast<uintptr_t>(f->data) & 7) != 0) return false;
    if ((f->stride[o] & 1) || (f->stride[o + 1] & 1)) return false;
    if (o == 1 && (f->stride[0] & 1)) return false;
    return true;
  }

  int solve(const sopht_field_t* sol, const sopht_field_t* rhs, cudaStream_t st) override {
    const bool vec = sol->ndim == 4;
    const int o = vec ? 1 : 0;
    const int C = vec ? (int)sol->shape[0] : 1;
    if (C > 3 || !view_ok(sol, o) || !view_ok(rhs, o)) {
      if (!generic) {
        int rc = SOPHT_OK;
        generic = make_generic_poisson(SOPHT_F32, 3, nz, ny, nx, dx, mz.data(), my.data(), mx.data(),
                                       origin, st, &rc);
        if (!generic) return rc;
      }
      return generic->solve(sol, rhs, st);
    }
    const int LY = 2 * ny, LZ = 2 * nz;
    const int64_t rows = (int64_t)C * nz * ny;
    const p2::SlabDims d{C, nz, ny, nx, 1, 0};
    int rc;
    p2::XParams xp = p2::slab_x_params(d, reinterpret_cast<const float*>(rhs->data), nullptr,
                                       vec ? rhs->stride[0] : 0, rhs->stride[o], rhs->stride[o + 1], A, nyqA,
                                       twx, twx2);
    if ((rc = launch_xfwd(nx, xp, rows, st))) return rc;
    // measured: -3 % per step at 128x128x256, -1.6 % at 256^3, but +4 % at 512^3 (the small high-priority kernels
    // disturb the persistent main kernels more than their own latency is worth there)
    static const int side_env = env_int("SOPHT_P2_SIDE_STREAM", -1);
    const bool use_side = side_env >= 0 ? side_env != 0 : (int64_t)nz * ny * nx <= ((int64_t)1 << 25);
    cudaStream_t side = use_side ? this->side : st;
    if (use_side) {
      SOPHT_CUDA(cudaEventRecord(ev_fork, st));
      SOPHT_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    }
    if ((rc = launch_yfwd(LY, p2::nyquist_y_params(d, TX, nyqA, nyqB, true, twy), dim3(C * nz / TX, 1, 1), side)))
      return rc;
    if ((rc = launch_zconv(LZ, p2::nyquist_z_params(d, TX, nyqB, gn, twz), dim3(LY / TX, 1, 1), side))) return rc;
    if ((rc = launch_yinv(LY, p2::nyquist_y_params(d, TX, nyqB, nyqA, false, twy), dim3(C * nz / TX, 1, 1), side)))
      return rc;
    if (use_side) SOPHT_CUDA(cudaEventRecord(ev_join, side));
    if ((rc = launch_yfwd(LY, p2::slab_y_params(d, TX, A, B, true, twy), dim3(nx / TX, C * nz, 1), st)))
      return rc;
    if (gt && zrow_tx(LZ, C, nx)) {
      if ((rc = launch_zrow(d, B, B2, gt, twz, st))) return rc;
    } else if ((rc = launch_zconv(LZ, p2::slab_z_params(d, TX, B, B2, gm, nx, 0, twz), dim3(nx / TX, LY, 1), st))) {
      return rc;
    }
    if ((rc = launch_yinv(LY, p2::slab_y_params(d, TX, B2, A, false, twy, gt && b2_z_blocked(LZ, C, nx)), dim3(nx / TX, C * nz, 1), st)))
      return rc;
    if (use_side) SOPHT_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    xp = p2::slab_x_params(d, nullptr, reinterpret_cast<float*>(sol->data), vec ? sol->stride[0] : 0,
                           sol->stride[o], sol->stride[o + 1], A, nyqA, twx, twx2);
    return launch_xinv(nx, xp, rows, st);
  }

  const void* green_hat() const override { return ghat_natural; }
  const char* path_name() const override { return "pow2"; }
};


// ---- periodic Poisson solve on the pow2 pipeline (BASELINE config 4; an extension, see poisson_neumann.cu) -----------
// -lap(psi) = rhs on the periodic box: x real-to-half-spectrum (complex length nx/2, Hermitian split), y, z forward,
// x norm / symbol, z, y inverse, x inverse - five kernels, every pass in place on ONE spectrum buffer (C, nz, ny, nx/2)
// + the kx = nx/2 plane: 40 B per cell and component, no cuFFT.
// With P > 1 ranks the same object runs the three local phases of the z-slab decomposed solve (x passes on nz / P
// planes, y / z passes on nx / 2 / P kx bins), the transposes fused into the x kernels like SlabPow2Poisson's.
struct PeriodicSymbols {
  float *lz = nullptr, *ly = nullptr, *lx = nullptr;
  float norm = 1.f;
  ~PeriodicSymbols() {
    cudaFree(lz);
    cudaFree(ly);
    cudaFree(lx);
  }
  static int upload(float** dst, int period, int count, double dx, bool three_point, cudaStream_t st) {
    const double pi = 3.14159265358979323846;
    std::vector<float> h(count);
    for (int k = 0; k < count; ++k) {
      if (!three_point) {
        const int m = k <= period / 2 ? k : k - period;
        const double w = 2.0 * pi * m / (period * dx);
        h[k] = (float)(w * w);
      } else {
        const double sn = sin(pi * k / period);
        h[k] = (float)(4.0 * sn * sn / (dx * dx));
      }
    }
    SOPHT_CUDA(cudaMalloc(dst, sizeof(float) * count));
    SOPHT_CUDA(cudaMemcpyAsync(*dst, h.data(), sizeof(float) * count, cudaMemcpyHostToDevice, st));
    SOPHT_CUDA(cudaStreamSynchronize(st));
    return SOPHT_OK;
  }
  int init(int nz, int ny, int nx, double dx, bool three_point, cudaStream_t st) {
    int rc;
    if ((rc = upload(&lz, nz, nz, dx, three_point, st))) return rc;
    if ((rc = upload(&ly, ny, ny, dx, three_point, st))) return rc;
    if ((rc = upload(&lx, nx, nx / 2 + 1, dx, three_point, st))) return rc;
    norm = (float)(1.0 / ((double)(nx / 2) * ny * nz));  // unnormalised round trip of the three transform lengths
    return SOPHT_OK;
  }
};

// y / z parameter blocks on a kx slab of nxl bins of a (C, nz, ny, nxl) spectrum, in place
p2::ColParams periodic_y_params(const p2::SlabDims& d, float2* a, const float2* tw) {
  const int64_t nxl = d.nxl();
  p2::ColParams yp{};
  yp.in = a, yp.out = a;
  yp.in_cs = 1, yp.out_cs = 1;
  yp.log2_bz = p2::ilog2(d.nz);
  yp.in_rs = yp.out_rs = nxl;
  yp.in_bx = yp.out_bx = TX;
  yp.in_by = yp.out_by = (int64_t)d.ny * nxl;
  yp.in_bc = yp.out_bc = (int64_t)d.nz * d.ny * nxl;
  yp.tw = tw;
  return yp;
}
p2::ColParams periodic_nyquist_y_params(const p2::SlabDims& d, float2* a, const float2* tw) {  // (C, nz, ny) plane
  p2::ColParams yn{};
  yn.in = a, yn.out = a;
  yn.in_rs = yn.out_rs = 1;
  yn.in_cs = yn.out_cs = d.ny;
  yn.in_bx = yn.out_bx = (int64_t)TX * d.ny;
  yn.tw = tw;
  return yn;
}
p2::ZSymParams periodic_z_params(const p2::SlabDims& d, float2* a, const PeriodicSymbols& sym, const float2* tw,
                                 int tx = TX) {
  const int64_t nxl = d.nxl();
  p2::ZSymParams zp{};
  zp.data = a;
  zp.rs = (int64_t)d.ny * nxl, zp.cs = 1, zp.d_bx = tx, zp.d_by = nxl, zp.d_c = (int64_t)d.nz * d.ny * nxl;
  zp.ncomp = d.C;
  zp.lz = sym.lz, zp.ly = sym.ly, zp.lx = sym.lx, zp.norm = sym.norm;
  zp.kx0 = d.rank * (int)nxl;
  zp.nyq = 0;
  zp.tw = tw;
  return zp;
}
p2::ZSymParams periodic_nyquist_z_params(const p2::SlabDims& d, float2* a, const PeriodicSymbols& sym,
                                         const float2* tw) {
  p2::ZSymParams zn{};
  zn.data = a;
  zn.rs = d.ny, zn.cs = 1, zn.d_bx = TX, zn.d_by = 0, zn.d_c = (int64_t)d.nz * d.ny;
  zn.ncomp = d.C;
  zn.lz = sym.lz, zn.ly = sym.ly, zn.lx = sym.lx, zn.norm = sym.norm;
  zn.nyq = 1;
  zn.kx_fixed = d.nx;  // SlabDims::nx counts complex bins per row = real nx / 2: the Nyquist bin
  zn.tw = tw;
  return zn;
}

// ---- z-slab decomposed solve: the three local phases between the all-to-all transposes ----------------------
// (the exchanges themselves are issued by the host layer on its process group; see SlabDims in
// poisson_pow2_phases.cuh and sopht_b200/parallel/slab_poisson.py)
struct SlabPow2Poisson {
  p2::SlabDims d{};       // periodic mode: d.nx counts the complex bins of a row (real nx / 2)
  bool periodic = false;  // periodic box: FULL kernels, symbol instead of G_hat, y / z passes in place
  PeriodicSymbols sym;
  float *gm = nullptr, *gn = nullptr;  // this rank's kx slice of the folded G_hat, and the Nyquist plane's
  float* gt = nullptr;                 // tile-major copy of gm for the row-mode z pass
  float2 *twx = nullptr, *twx2 = nullptr, *twy = nullptr, *twz = nullptr;
  // peer exchange: library-owned (cudaMalloc, so the IPC handle maps the exact base) buffers
  // (C, P, nzl, ny, nxl); peer_recv[q] / peer_send[q] are rank q's buffers mapped into this process
  float2 *xrecv = nullptr, *xsend = nullptr;
  float2* peer_recv[8] = {};
  float2* peer_send[8] = {};
  bool peers_open = false;

  size_t exchange_bytes() const { return sizeof(float2) * (size_t)d.C * d.nz * d.ny * d.nxl(); }
  // pipelined solve: the kx = nx (Nyquist) plane of every rank, (C, nz, ny), lives behind the spectrum in the xrecv
  // block so that the peers can fill their planes of it with a copy-engine copy
  size_t nyquist_bytes() const { return sizeof(float2) * (size_t)d.C * d.nz * d.ny; }
  float2* nyq_all_peer(int q) const { return peer_recv[q] + exchange_bytes() / sizeof(float2); }

  ~SlabPow2Poisson() {
    if (peers_open)
      for (int q = 0; q < d.P; ++q)
        if (q != d.rank) {
          cudaIpcCloseMemHandle(peer_recv[q]);
          cudaIpcCloseMemHandle(peer_send[q]);
        }
    cudaFree(xrecv);
    cudaFree(xsend);
    cudaFree(gm);
    cudaFree(gn);
    cudaFree(gt);
    cudaFree(twx);
    cudaFree(twx2);
    cudaFree(twy);
    cudaFree(twz);
  }

  int init_periodic(double dx, bool three_point, cudaStream_t st) {
    periodic = true;
    const int nz = d.nz, ny = d.ny, lx = d.nx;
    int rc;
    if ((rc = sym.init(nz, ny, 2 * lx, dx, three_point, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twx, lx, lx, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twx2, lx, 2 * lx, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twy, ny, ny, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twz, nz, nz, st))) return rc;
    return SOPHT_OK;
  }

  int init(double dx, const double* mz, const double* my, const double* mx, double origin, cudaStream_t st) {
    const int nz = d.nz, ny = d.ny, nx = d.nx, nxl = d.nxl();
    if (cudaMalloc(&gm, sizeof(float) * (size_t)(nz + 1) * (ny + 1) * nxl) != cudaSuccess ||
        cudaMalloc(&gn, sizeof(float) * (size_t)(nz + 1) * (ny + 1)) != cudaSuccess)
      SOPHT_FAIL(SOPHT_ERR_ALLOC, "poisson(slab): out of device memory for the Green's function slice");
    // only this rank's kx range of G_hat is ever formed (the full doubled-domain transform does not fit at 1024^3)
    int rc = build_green_folded_slice(gm, gn, nz, ny, nx, d.rank * nxl, nxl, dx, mz, my, mx, origin, st);
    if (rc) return rc;
    if (d.C == 3 && (rc = build_green_tiles(&gt, gm, nz, ny, nxl, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twx, nx, nx, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twx2, nx, 2 * nx, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twy, 2 * ny, 2 * ny, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twz, 2 * nz, 2 * nz, st))) return rc;
    return SOPHT_OK;
  }

  static bool real_view_ok(const sopht_field_t* f) {
    if (f->ndim != 4 || f->stride[3] != 1) return false;
    if ((reinterpret_cast<uintptr_t>(f->data) & 7) != 0) return false;
    return !((f->stride[0] | f->stride[1] | f->stride[2]) & 1);
  }

  int check_local(const char* fn, const sopht_field_t* f) const {
    const int nx_real = periodic ? 2 * d.nx : d.nx;
    if (!valid_field(f, 4, 4) || f->shape[0] != d.C || f->shape[1] != d.nzl() || f->shape[2] != d.ny ||
        f->shape[3] != nx_real)
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected this rank's (%d, %d, %d, %d) z-slab", fn, d.C, d.nzl(), d.ny,
                 nx_real);
    if (!real_view_ok(f))
      SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: rows must be contiguous, 8-byte aligned, even plane/row strides", fn);
    return SOPHT_OK;
  }

  int enable_peer_exchange(unsigned char* handles_out) {
    if (!xrecv) {
      if (cudaMalloc(&xrecv, exchange_bytes() + nyquist_bytes()) != cudaSuccess ||
          cudaMalloc(&xsend, exchange_bytes()) != cudaSuccess)
        SOPHT_FAIL(SOPHT_ERR_ALLOC, "poisson(slab): out of device memory for the exchange buffers");
    }
    cudaIpcMemHandle_t h0, h1;
    SOPHT_CUDA(cudaIpcGetMemHandle(&h0, xrecv));
    SOPHT_CUDA(cudaIpcGetMemHandle(&h1, xsend));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles_out, &h0, 64);
    memcpy(handles_out + 64, &h1, 64);
    return SOPHT_OK;
  }

  int open_peers(const unsigned char* all_handles) {
    if (!xrecv) SOPHT_FAIL(SOPHT_ERR_HANDLE, "poisson(slab): enable_peer_exchange first");
    for (int q = 0; q < d.P; ++q) {
      if (q == d.rank) {
        peer_recv[q] = xrecv;
        peer_send[q] = xsend;
        continue;
      }
      cudaIpcMemHandle_t h0, h1;
      memcpy(&h0, all_handles + (size_t)q * 128, 64);
      memcpy(&h1, all_handles + (size_t)q * 128 + 64, 64);
      void *p0 = nullptr, *p1 = nullptr;
      SOPHT_CUDA(cudaIpcOpenMemHandle(&p0, h0, cudaIpcMemLazyEnablePeerAccess));
      SOPHT_CUDA(cudaIpcOpenMemHandle(&p1, h1, cudaIpcMemLazyEnablePeerAccess));
      peer_recv[q] = reinterpret_cast<float2*>(p0);
      peer_send[q] = reinterpret_cast<float2*>(p1);
    }
    peers_open = true;
    return SOPHT_OK;
  }

  // x forward; send == nullptr: the spectrum chunks go straight into the peers' exchange buffers (NVLink stores)
  int forward_x(const sopht_field_t* rhs, float2* send, float2* nyq_local, cudaStream_t st) const {
    int rc = check_local(__func__, rhs);
    if (rc) return rc;
    if (!send && !peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    p2::XParams xp =
        p2::slab_x_params(d, reinterpret_cast<const float*>(rhs->data), nullptr, rhs->stride[0], rhs->stride[1],
                          rhs->stride[2], send ? send : xrecv, nyq_local, twx, twx2);
    if (!send) xp = p2::slab_x_params_peer(xp, d, peer_recv);
    if (periodic) return launch_xfwd_full(d.nx, xp, (int64_t)d.C * d.nzl() * d.ny, st);
    return launch_xfwd(d.nx, xp, (int64_t)d.C * d.nzl() * d.ny, st);
  }

  // recv == nullptr: read this rank's own exchange buffer (filled by the peers' x forward kernels) and leave the y
  // inverse's kx-slab (C, nz, ny, nxl) in this rank's second buffer, from where the peers' x inverse kernels pull it
  int yz(float2* recv, float2* nyq_all, float2* work, float2* nyq_work, cudaStream_t st) const {
    const int LY = 2 * d.ny, LZ = 2 * d.nz, nxl = d.nxl();
    const bool peer = recv == nullptr;
    if (peer && !peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    if (peer) recv = xrecv;
    int rc;
    if (periodic) {  // in place on the received kx slab; the y inverse leaves it where the reverse transpose reads it
      const dim3 gy(nxl / TX, d.C * d.nz, 1), gyn(d.C * d.nz / TX, 1, 1);
      if ((rc = launch_yfwd_full(d.ny, periodic_y_params(d, recv, twy), gy, st))) return rc;
      if ((rc = launch_yfwd_full(d.ny, periodic_nyquist_y_params(d, nyq_all, twy), gyn, st))) return rc;
      const int ztx = zsym_tx(d.nz, nxl);
      if ((rc = launch_zsym(d.nz, periodic_z_params(d, recv, sym, twz, ztx), dim3(nxl / ztx, d.ny, 1), st, ztx)))
        return rc;
      if ((rc = launch_zsym(d.nz, periodic_nyquist_z_params(d, nyq_all, sym, twz), dim3(d.ny / TX, 1, 1), st)))
        return rc;
      p2::ColParams yi = periodic_y_params(d, recv, twy);
      if (peer) yi.out = xsend;
      if ((rc = launch_yinv_full(d.ny, yi, gy, st))) return rc;
      return launch_yinv_full(d.ny, periodic_nyquist_y_params(d, nyq_all, twy), gyn, st);
    }
    if (!work || !nyq_work) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null work buffer", __func__);
    if ((rc = launch_yfwd(LY, p2::slab_y_params(d, TX, recv, work, true, twy), dim3(nxl / TX, d.C * d.nz, 1), st)))
      return rc;
    if ((rc = launch_yfwd(LY, p2::nyquist_y_params(d, TX, nyq_all, nyq_work, true, twy),
                          dim3(d.C * d.nz / TX, 1, 1), st)))
      return rc;
    float2* work2 = work + (int64_t)d.C * d.nz * LY * nxl;  // second half: the z pass's tile-major output
    if (gt && zrow_tx(LZ, d.C, nxl)) {
      if ((rc = launch_zrow(d, work, work2, gt, twz, st))) return rc;
    } else if ((rc = launch_zconv(LZ, p2::slab_z_params(d, TX, work, work2, gm, nxl, 0, twz), dim3(nxl / TX, LY, 1),
                                  st))) {
      return rc;
    }
    if ((rc = launch_zconv(LZ, p2::nyquist_z_params(d, TX, nyq_work, gn, twz), dim3(LY / TX, 1, 1), st)))
      return rc;
    const p2::ColParams yi =
        p2::slab_y_params(d, TX, work2, peer ? xsend : recv, false, twy, gt && b2_z_blocked(LZ, d.C, nxl));
    if ((rc = launch_yinv(LY, yi, dim3(nxl / TX, d.C * d.nz, 1), st))) return rc;
    return launch_yinv(LY, p2::nyquist_y_params(d, TX, nyq_work, nyq_all, false, twy),
                       dim3(d.C * d.nz / TX, 1, 1), st);
  }

  // ---- pipelined solve: one component at a time, transposes by the copy engines -------------------------------------
  // The fused transposes above keep the SMs busy with NVLink stores / loads for 2 x 4.3 ms of a 25 ms step at 1024^3 on
  // 8 GPUs while nothing else runs. Here every phase works on ONE component: the x forward pass writes a plain
  // row-major spectrum S[c] (the xsend block reinterpreted as (C, rows, nx)), cudaMemcpy2DAsync copies (DMA, no SM)
  // move its kx chunks into the peers' xrecv blocks while the SMs run the y / z passes of the previous component, and
  // the way back mirrors it. The host layer (parallel/slab_poisson.py) orders the phases with events and the
  // peer-arena barrier.
  p2::SlabDims comp_dims_x() const { return p2::SlabDims{1, d.nzl(), d.ny, d.nx, 1, 0}; }  // this rank's rows, whole rows
  int64_t rows_local() const { return (int64_t)d.nzl() * d.ny; }

  int pipe_forward_x(const sopht_field_t* rhs, int c, float2* nyq_local, cudaStream_t st) const {
    int rc = check_local(__func__, rhs);
    if (rc) return rc;
    if (!peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    const p2::SlabDims dx_ = comp_dims_x();
    p2::XParams xp = p2::slab_x_params(dx_, reinterpret_cast<const float*>(rhs->data) + (int64_t)c * rhs->stride[0],
                                       nullptr, 0, rhs->stride[1], rhs->stride[2], xsend + (int64_t)c * rows_local() * d.nx,
                                       nyq_local + (int64_t)c * rows_local(), twx, twx2);
    return periodic ? launch_xfwd_full(d.nx, xp, rows_local(), st) : launch_xfwd(d.nx, xp, rows_local(), st);
  }
  int pipe_inverse_x(const sopht_field_t* sol, int c, float2* nyq_local, cudaStream_t st) const {
    int rc = check_local(__func__, sol);
    if (rc) return rc;
    if (!peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    const p2::SlabDims dx_ = comp_dims_x();
    p2::XParams xp = p2::slab_x_params(dx_, nullptr, reinterpret_cast<float*>(sol->data) + (int64_t)c * sol->stride[0], 0,
                                       sol->stride[1], sol->stride[2], xsend + (int64_t)c * rows_local() * d.nx,
                                       nyq_local + (int64_t)c * rows_local(), twx, twx2);
    return periodic ? launch_xinv_full(d.nx, xp, rows_local(), st) : launch_xinv(d.nx, xp, rows_local(), st);
  }
  // forward: chunk q of my S[c] rows -> slot `rank` of rank q's recv[c]; my Nyquist bins -> everybody's nyq_all.
  // backward: z block p of my recv[c] -> kx chunk `rank` of rank p's S[c] rows. Copies go round robin over `streams`.
  int pipe_transpose(int c, int backward, cudaStream_t* streams, int nstreams, const float2* nyq_local) const {
    if (!peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    if (nstreams < 1 || !streams) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: at least one stream", __func__);
    const int64_t rows = rows_local(), nxl = d.nxl(), nx = d.nx;
    const int64_t comp_chunks = (int64_t)d.P * rows * nxl;  // elements of one component of an exchange buffer
    for (int k = 0; k < d.P; ++k) {
      const int q = (d.rank + k) % d.P;  // start with the self copy, then ring order: all links busy at once
      cudaStream_t st = streams[k % nstreams];
      if (!backward) {
        const float2* src = xsend + (int64_t)c * rows * nx + q * nxl;
        float2* dst = peer_recv[q] + c * comp_chunks + (int64_t)d.rank * rows * nxl;
        SOPHT_CUDA(cudaMemcpy2DAsync(dst, sizeof(float2) * nxl, src, sizeof(float2) * nx, sizeof(float2) * nxl, rows,
                                     cudaMemcpyDeviceToDevice, st));
        if (nyq_local)
          SOPHT_CUDA(cudaMemcpyAsync(nyq_all_peer(q) + ((int64_t)c * d.nz + (int64_t)d.rank * d.nzl()) * d.ny,
                                     nyq_local + (int64_t)c * rows, sizeof(float2) * rows, cudaMemcpyDeviceToDevice,
                                     st));
      } else {
        const float2* src = xrecv + c * comp_chunks + (int64_t)q * rows * nxl;
        float2* dst = peer_send[q] + (int64_t)c * rows * nx + (int64_t)d.rank * nxl;
        SOPHT_CUDA(cudaMemcpy2DAsync(dst, sizeof(float2) * nx, src, sizeof(float2) * nxl, sizeof(float2) * nxl, rows,
                                     cudaMemcpyDeviceToDevice, st));
      }
    }
    return SOPHT_OK;
  }
  // y forward, z, y inverse of component c on the received kx slab, in place in xrecv[c] (and the Nyquist plane)
  int pipe_yz(int c, float2* work, float2* nyq_work, cudaStream_t st) const {
    if (!peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    const int LY = 2 * d.ny, LZ = 2 * d.nz, nxl = d.nxl();
    p2::SlabDims d1 = d;
    d1.C = 1;
    float2* recv = xrecv + (int64_t)c * d.nz * d.ny * nxl;
    float2* nyq = nyq_all_peer(d.rank) + (int64_t)c * d.nz * d.ny;
    int rc;
    if (periodic) {
      const dim3 gy(nxl / TX, d.nz, 1), gyn(d.nz / TX, 1, 1);
      if ((rc = launch_yfwd_full(d.ny, periodic_y_params(d1, recv, twy), gy, st))) return rc;
      if ((rc = launch_yfwd_full(d.ny, periodic_nyquist_y_params(d1, nyq, twy), gyn, st))) return rc;
      const int ztx = zsym_tx(d.nz, nxl);
      if ((rc = launch_zsym(d.nz, periodic_z_params(d1, recv, sym, twz, ztx), dim3(nxl / ztx, d.ny, 1), st, ztx)))
        return rc;
      if ((rc = launch_zsym(d.nz, periodic_nyquist_z_params(d1, nyq, sym, twz), dim3(d.ny / TX, 1, 1), st))) return rc;
      if ((rc = launch_yinv_full(d.ny, periodic_y_params(d1, recv, twy), gy, st))) return rc;
      return launch_yinv_full(d.ny, periodic_nyquist_y_params(d1, nyq, twy), gyn, st);
    }
    if (!work || !nyq_work) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null work buffer", __func__);
    float2* work2 = work + (int64_t)d.nz * LY * nxl;  // second half: the z pass's tile-major output
    if ((rc = launch_yfwd(LY, p2::slab_y_params(d1, TX, recv, work, true, twy), dim3(nxl / TX, d.nz, 1), st))) return rc;
    if ((rc = launch_yfwd(LY, p2::nyquist_y_params(d1, TX, nyq, nyq_work, true, twy), dim3(d.nz / TX, 1, 1), st)))
      return rc;
    if ((rc = launch_zconv(LZ, p2::slab_z_params(d1, TX, work, work2, gm, nxl, 0, twz), dim3(nxl / TX, LY, 1), st)))
      return rc;
    if ((rc = launch_zconv(LZ, p2::nyquist_z_params(d1, TX, nyq_work, gn, twz), dim3(LY / TX, 1, 1), st))) return rc;
    if ((rc = launch_yinv(LY, p2::slab_y_params(d1, TX, work2, recv, false, twy), dim3(nxl / TX, d.nz, 1), st)))
      return rc;
    return launch_yinv(LY, p2::nyquist_y_params(d1, TX, nyq_work, nyq, false, twy), dim3(d.nz / TX, 1, 1), st);
  }
  // this rank's planes of the processed Nyquist plane of component c -> nyq_local[c] (the x inverse reads it)
  int pipe_nyquist_slice(int c, float2* nyq_local, cudaStream_t st) const {
    const float2* src = nyq_all_peer(d.rank) + ((int64_t)c * d.nz + (int64_t)d.rank * d.nzl()) * d.ny;
    SOPHT_CUDA(cudaMemcpyAsync(nyq_local + (int64_t)c * rows_local(), src, sizeof(float2) * rows_local(),
                               cudaMemcpyDeviceToDevice, st));
    return SOPHT_OK;
  }

  int inverse_x(const sopht_field_t* sol, float2* recv2, float2* nyq_local, cudaStream_t st) const {
    int rc = check_local(__func__, sol);
    if (rc) return rc;
    if (!recv2 && !peers_open) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: peer exchange is not open", __func__);
    const bool pull = recv2 == nullptr;  // chunk q of every spectrum row is read from rank q's buffer over NVLink
    if (pull) recv2 = xsend;
    p2::XParams xp = p2::slab_x_params(d, nullptr, reinterpret_cast<float*>(sol->data), sol->stride[0],
                                       sol->stride[1], sol->stride[2], recv2, nyq_local, twx, twx2);
    if (pull) xp = p2::slab_x_params_peer(xp, d, peer_send);
    if (periodic) return launch_xinv_full(d.nx, xp, (int64_t)d.C * d.nzl() * d.ny, st);
    return launch_xinv(d.nx, xp, (int64_t)d.C * d.nzl() * d.ny, st);
  }
};


struct PeriodicPow2Poisson : PoissonImpl {
  int nz, ny, nx;  // real grid
  bool three_point;
  PeriodicSymbols sym;
  float2 *A = nullptr, *nyqA = nullptr;
  float2 *twx = nullptr, *twx2 = nullptr, *twy = nullptr, *twz = nullptr;
  PoissonImpl* generic = nullptr;
  double dx;

  ~PeriodicPow2Poisson() override {
    cudaFree(A);
    cudaFree(nyqA);
    cudaFree(twx);
    cudaFree(twx2);
    cudaFree(twy);
    cudaFree(twz);
    delete generic;
  }
  int init(cudaStream_t st) {
    int rc;
    if ((rc = sym.init(nz, ny, nx, dx, three_point, st))) return rc;
    const size_t rows = (size_t)3 * nz * ny;
    SOPHT_CUDA(cudaMalloc(&A, sizeof(float2) * rows * (nx / 2)));
    SOPHT_CUDA(cudaMalloc(&nyqA, sizeof(float2) * rows));
    if ((rc = Pow2Poisson::upload_twiddles(&twx, nx / 2, nx / 2, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twx2, nx / 2, nx, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twy, ny, ny, st))) return rc;
    if ((rc = Pow2Poisson::upload_twiddles(&twz, nz, nz, st))) return rc;
    return SOPHT_OK;
  }
  int solve(const sopht_field_t* sol, const sopht_field_t* rhs, cudaStream_t st) override {
    const bool vec = sol->ndim == 4;
    const int o = vec ? 1 : 0;
    const int C = vec ? (int)sol->shape[0] : 1;
    if (C > 3 || !Pow2Poisson::view_ok(sol, o) || !Pow2Poisson::view_ok(rhs, o)) {
      if (!generic) {
        int rc = SOPHT_OK;
        generic = make_periodic_poisson(SOPHT_F32, three_point, 3, nz, ny, nx, dx, st, &rc);
        if (!generic) return rc;
      }
      return generic->solve(sol, rhs, st);
    }
    const int LX = nx / 2;
    const int64_t rows = (int64_t)C * nz * ny;
    const p2::SlabDims d{C, nz, ny, LX, 1, 0};
    int rc;
    p2::XParams xp = p2::slab_x_params(d, reinterpret_cast<const float*>(rhs->data), nullptr,
                                       vec ? rhs->stride[0] : 0, rhs->stride[o], rhs->stride[o + 1], A, nyqA, twx,
                                       twx2);
    if ((rc = launch_xfwd_full(LX, xp, rows, st))) return rc;
    if ((rc = launch_yfwd_full(ny, periodic_y_params(d, A, twy), dim3(LX / TX, C * nz, 1), st))) return rc;
    if ((rc = launch_yfwd_full(ny, periodic_nyquist_y_params(d, nyqA, twy), dim3(C * nz / TX, 1, 1), st))) return rc;
    const int ztx = zsym_tx(nz, LX);
    if ((rc = launch_zsym(nz, periodic_z_params(d, A, sym, twz, ztx), dim3(LX / ztx, ny, 1), st, ztx))) return rc;
    if ((rc = launch_zsym(nz, periodic_nyquist_z_params(d, nyqA, sym, twz), dim3(ny / TX, 1, 1), st))) return rc;
    if ((rc = launch_yinv_full(ny, periodic_y_params(d, A, twy), dim3(LX / TX, C * nz, 1), st))) return rc;
    if ((rc = launch_yinv_full(ny, periodic_nyquist_y_params(d, nyqA, twy), dim3(C * nz / TX, 1, 1), st))) return rc;
    xp = p2::slab_x_params(d, nullptr, reinterpret_cast<float*>(sol->data), vec ? sol->stride[0] : 0, sol->stride[o],
                           sol->stride[o + 1], A, nyqA, twx, twx2);
    return launch_xinv_full(LX, xp, rows, st);
  }
  const char* path_name() const override { return three_point ? "periodic_pow2_three_point" : "periodic_pow2_spectral"; }
};

}  // namespace

bool pow2_poisson_eligible(int dtype, int dim, int nz, int ny, int nx) {
  return dtype == SOPHT_F32 && dim == 3 && is_pow2(nx) && is_pow2(ny) && is_pow2(nz) && nx >= 16 &&
         nx <= 2048 && ny >= 8 && ny <= 1024 && nz >= 8 && nz <= 1024;
}

PoissonImpl* make_pow2_poisson(int nz, int ny, int nx, double dx, const double* mz, const double* my,
                               const double* mx, double origin, cudaStream_t st, int* rc) {
  auto* p = new Pow2Poisson();
  p->nz = nz, p->ny = ny, p->nx = nx, p->dx = dx, p->origin = origin;
  p->mz.assign(mz, mz + 2 * nz);
  p->my.assign(my, my + 2 * ny);
  p->mx.assign(mx, mx + 2 * nx);
  *rc = p->init(st);
  if (*rc) {
    delete p;
    return nullptr;
  }
  return p;
}

bool periodic_pow2_eligible(int dtype, int dim, int nz, int ny, int nx) {
  static const int off = env_int("SOPHT_PERIODIC_FORCE_CUFFT", 0);
  return !off && dtype == SOPHT_F32 && dim == 3 && is_pow2(nx) && is_pow2(ny) && is_pow2(nz) && nx >= 32 &&
         nx <= 4096 && ny >= 16 && ny <= 2048 && nz >= 16 && nz <= 2048;
}
PoissonImpl* make_periodic_pow2_poisson(int three_point_symbol, int nz, int ny, int nx, double dx, cudaStream_t st,
                                        int* rc) {
  auto* p = new PeriodicPow2Poisson();
  p->nz = nz, p->ny = ny, p->nx = nx, p->dx = dx, p->three_point = three_point_symbol != 0;
  *rc = p->init(st);
  if (*rc) {
    delete p;
    return nullptr;
  }
  return p;
}

}  // namespace sopht

using namespace sopht;

struct sopht_poisson_slab {
  SlabPow2Poisson impl;
};

extern "C" {

int sopht_poisson_slab_create(sopht_poisson_slab_t* handle, int ncomp, int nz, int ny, int nx, int nranks,
                              int rank, double dx, const double* mz, const double* my, const double* mx,
                              double origin_value, void* stream) {
  if (!handle) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null handle pointer", __func__);
  if (!pow2_poisson_eligible(SOPHT_F32, 3, nz, ny, nx))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: the slab solver needs a power-of-two fp32 3-D grid", __func__);
  if (ncomp < 1 || ncomp > 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: 1..3 components", __func__);
  if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: nranks must be a power of two and 0 <= rank < nranks", __func__);
  if (nz % nranks || nx % nranks || (nx / nranks) % TX || ((int64_t)ncomp * (nz / nranks) * ny) % 32 ||
      (ncomp * nz) % TX)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: nz/nranks planes and nx/nranks >= %d kx bins per rank are required",
               __func__, TX);
  if (!mz || !my || !mx) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null Green's function coordinate arrays", __func__);
  auto* h = new sopht_poisson_slab();
  h->impl.d = p2::SlabDims{ncomp, nz, ny, nx, nranks, rank};
  const int rc = h->impl.init(dx, mz, my, mx, origin_value, as_stream(stream));
  if (rc) {
    delete h;
    return rc;
  }
  *handle = h;
  return SOPHT_OK;
}

int sopht_poisson_slab_create_periodic(sopht_poisson_slab_t* handle, int ncomp, int nz, int ny, int nx, int nranks,
                                       int rank, double dx, int three_point_symbol, void* stream) {
  if (!handle) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null handle pointer", __func__);
  if (!periodic_pow2_eligible(SOPHT_F32, 3, nz, ny, nx))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: the slab solver needs a power-of-two fp32 3-D grid", __func__);
  if (ncomp < 1 || ncomp > 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: 1..3 components", __func__);
  if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: nranks must be a power of two and 0 <= rank < nranks", __func__);
  const int lx = nx / 2;
  if (nz % nranks || lx % nranks || (lx / nranks) % TX || ((int64_t)ncomp * (nz / nranks) * ny) % 32 ||
      (ncomp * nz) % TX || ny % TX)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: nz/nranks planes and nx/2/nranks >= %d kx bins per rank are required", __func__,
               TX);
  if (!(dx > 0)) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dx must be positive", __func__);
  auto* h = new sopht_poisson_slab();
  h->impl.d = p2::SlabDims{ncomp, nz, ny, lx, nranks, rank};
  const int rc = h->impl.init_periodic(dx, three_point_symbol != 0, as_stream(stream));
  if (rc) {
    delete h;
    return rc;
  }
  *handle = h;
  return SOPHT_OK;
}

int sopht_poisson_slab_forward_x(sopht_poisson_slab_t h, const sopht_field_t* rhs_field, void* send_buffer,
                                 void* nyquist_local, void* stream) {
  if (!h || !nyquist_local) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  return h->impl.forward_x(rhs_field, reinterpret_cast<float2*>(send_buffer),
                           reinterpret_cast<float2*>(nyquist_local), as_stream(stream));
}

int sopht_poisson_slab_yz(sopht_poisson_slab_t h, void* recv_buffer, void* nyquist_all, void* work_buffer,
                          void* nyquist_work, void* stream) {
  if (!h || !nyquist_all) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  return h->impl.yz(reinterpret_cast<float2*>(recv_buffer), reinterpret_cast<float2*>(nyquist_all),
                    reinterpret_cast<float2*>(work_buffer), reinterpret_cast<float2*>(nyquist_work),
                    as_stream(stream));
}

int sopht_poisson_slab_inverse_x(sopht_poisson_slab_t h, const sopht_field_t* solution_field, void* recv_buffer,
                                 void* nyquist_local, void* stream) {
  if (!h || !nyquist_local) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  return h->impl.inverse_x(solution_field, reinterpret_cast<float2*>(recv_buffer),
                           reinterpret_cast<float2*>(nyquist_local), as_stream(stream));
}

int sopht_poisson_slab_enable_peer_exchange(sopht_poisson_slab_t h, unsigned char* ipc_handles_out) {
  if (!h || !ipc_handles_out) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  return h->impl.enable_peer_exchange(ipc_handles_out);
}

int sopht_poisson_slab_open_peers(sopht_poisson_slab_t h, const unsigned char* all_ipc_handles) {
  if (!h || !all_ipc_handles) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  return h->impl.open_peers(all_ipc_handles);
}

/* ---- pipelined solve (one component per call, copy-engine transposes): see SlabPow2Poisson::pipe_* ---- */
int sopht_poisson_slab_pipe_forward_x(sopht_poisson_slab_t h, const sopht_field_t* rhs_field, int component,
                                      void* nyquist_local, void* stream) {
  if (!h || !nyquist_local) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  if (component < 0 || component >= h->impl.d.C) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: bad component", __func__);
  return h->impl.pipe_forward_x(rhs_field, component, reinterpret_cast<float2*>(nyquist_local), as_stream(stream));
}
int sopht_poisson_slab_pipe_transpose(sopht_poisson_slab_t h, int component, int backward, void** streams,
                                      int nstreams, const void* nyquist_local) {
  if (!h) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  if (component < 0 || component >= h->impl.d.C) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: bad component", __func__);
  if (nstreams < 1 || nstreams > 16 || !streams) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: 1..16 streams", __func__);
  cudaStream_t st[16];
  for (int i = 0; i < nstreams; ++i) st[i] = as_stream(streams[i]);
  return h->impl.pipe_transpose(component, backward, st, nstreams, reinterpret_cast<const float2*>(nyquist_local));
}
int sopht_poisson_slab_pipe_yz(sopht_poisson_slab_t h, int component, void* work_buffer, void* nyquist_work,
                               void* stream) {
  if (!h) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  if (component < 0 || component >= h->impl.d.C) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: bad component", __func__);
  return h->impl.pipe_yz(component, reinterpret_cast<float2*>(work_buffer), reinterpret_cast<float2*>(nyquist_work),
                         as_stream(stream));
}
int sopht_poisson_slab_pipe_inverse_x(sopht_poisson_slab_t h, const sopht_field_t* solution_field, int component,
                                      void* nyquist_local, void* stream) {
  if (!h || !nyquist_local) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  if (component < 0 || component >= h->impl.d.C) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: bad component", __func__);
  int rc = h->impl.pipe_nyquist_slice(component, reinterpret_cast<float2*>(nyquist_local), as_stream(stream));
  if (rc) return rc;
  return h->impl.pipe_inverse_x(solution_field, component, reinterpret_cast<float2*>(nyquist_local), as_stream(stream));
}

int sopht_poisson_slab_destroy(sopht_poisson_slab_t h) {
  delete h;
  return SOPHT_OK;
}

}  // extern "C"
