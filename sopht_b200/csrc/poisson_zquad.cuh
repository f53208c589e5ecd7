// z pass of the pruned FFT pipeline, "four columns per warp quartet" form (2 nz = 1024, three components).
//
// Why a third form. The decomposition of zrow_kernel (profiles/r02_zpass_experiments.txt) shows that its bound is the
// LSU / shared-memory pipe (~1 wavefront per clock and SM), not the FP32 pipe: per 12-column unit the pipe carries
// 4224 wavefronts of consumer exchanges plus ~4200 cycles of producer traffic - LDGSTS.64 that land sector by sector
// (8 wavefronts + 8 tag cycles per instruction because a warp's 32 elements come from 8 rows 4 MB apart) and the drain of
// the results through shared memory (STS by the consumers, LDS + STG by the producers). Both disappear when the lanes of
// a warp cover WHOLE 32-byte sectors of the global arrays:
//   * the four warps of a component own the unit's four kx columns TOGETHER: lane (jq = lane / 4, c = lane % 4) of warp
//     w works on butterfly t = 8 w + jq of column c. Element z = 32 n + t of the four columns is one 32-byte row of the
//     x-major spectrum and one 32-byte sector of the tile-major output, so
//   * the rows are staged exactly as they lie in memory ([z][4 columns], 16-byte cp.async: half the copy instructions of
//     the 8-byte transposing copies, dense 256-byte reads by the consumers instead of a column-major stage), and
//   * the inverse last pass stores its results straight from registers: every warp-wide store writes 8 complete
//     sectors. No drain, no second trip through shared memory, and the stage buffer is free again after the FIRST pass
//     of a unit instead of after the last (the prefetch distance grows by two passes).
// Price: the passes of a column now exchange data between the four warps of the quartet, so the pass boundaries are
// 128-thread named barriers (one warp per scheduler takes part) instead of __syncwarp().
// Shared memory, conflict-free by construction: exchange columns with pitch PQ = 4 (mod 16) float2 (the half warp
// jq 0..3 x c 0..3 hits 16 distinct 8-byte slots in every pass), G_hat columns with pitch 2 (mod 8) float4.
// stage 2 x 3 x 16 KB + exchange 12 x 1060 x 8 B + G 3 x 4 x 552 x 4 B + barriers = 221.4 KB, one persistent CTA per SM.
#pragma once
#include "poisson_zrow.cuh"

namespace sopht {
namespace p2 {

#ifdef __CUDACC__
// element z of one column of a [z][4 columns] stage tile
struct QuadLoad {
  const float2* p;
  __device__ __forceinline__ float2 operator()(int e) const { return p[e * 4]; }
};
// element z of a column of the kx-tile(8)-major spectrum: planes z % 8 are 8 complex apart, blocks z / 8 bz8 apart
struct TileStore {
  float2* p;
  int64_t bz8;
  __device__ __forceinline__ void operator()(int z, float2 v) const { p[(z >> 3) * bz8 + (z & 7) * 8] = v; }
};
__device__ __forceinline__ void quartet_sync(int g) {
  asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}

template <int L>
struct ZQuad {
  using C = Cfg<L>;
  using R = ZRow<L>;
  static_assert(L == 1024 && C::T == 32 && C::NP == 2, "written for the two-pass radix-32 length");
  static constexpr int CWARPS = 12, CTHREADS = 32 * CWARPS, PWARPS = 4, PTHREADS = 32 * PWARPS;
  static constexpr int THREADS = CTHREADS + PTHREADS;
  static constexpr int TX = 4, NCOL = 3 * TX, H = L / 2;
  static constexpr int TILE = H * TX;                       // float2 per (component, buffer) stage tile
  static constexpr int PQ = 1060;                           // exchange column pitch
  static constexpr int GP = R::GP, GPQ = GP + 24;           // floats per G column in global memory / in shared memory
  static constexpr int NIT = H * TX / 2 / PTHREADS;         // 16-byte copies per producer thread, component and unit
  static constexpr int G_CHUNKS = TX * GP / 4;
  static constexpr size_t SMEM_BYTES = sizeof(float2) * (2 * 3 * TILE + NCOL * PQ) + sizeof(float) * 3 * TX * GPQ + 128;
  static_assert(PQ >= RowAcc<L>::PITCH && PQ % 16 == 4, "exchange pitch");
  static_assert((GPQ / 4) % 8 == 2 && GP % 4 == 0, "G pitch");
  static_assert(NIT == 8, "one batch of pinned copy addresses");
};

template <int L>
__global__ void __launch_bounds__(ZQuad<L>::THREADS, 1) zquad_kernel(const ZRowParams p, int nunits) {
  using K = ZQuad<L>;
  extern __shared__ __align__(16) unsigned char zquad_smem[];
  float2* stage = reinterpret_cast<float2*>(zquad_smem);              // [2][3][H][TX]
  float2* work = stage + 2 * 3 * K::TILE;                             // [NCOL][PQ]
  float* gbuf = reinterpret_cast<float*>(work + K::NCOL * K::PQ);     // [3][TX][GPQ]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbuf + 3 * K::TX * K::GPQ);
  // mbarriers: full[g][b] = bars[2 g + b] (rows of component g, buffer b, have landed), free[g][b] = bars[6 + 2 g + b]
  // (the quartet of component g has read buffer b), gfull[j] = bars[12 + j] (G tile buffer j has landed)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Units are handed out in PAIRS of neighbouring kx quads (pair P = units 2P, 2P + 1; CTA b owns pairs b, b + gridDim.x,
  // ...): the two 32-byte rows of a pair are the halves of one 64-byte L2 fetch, so the second unit's rows are L2 hits
  // instead of a second DRAM fetch by another CTA (DRAM reads of the launch 10.2 -> 8.8 GB at 512^3; same time). nunits is even
  // (the kx range is a multiple of 8).
  const int npairs = nunits / 2;
  if ((int)blockIdx.x >= npairs) return;
  const int cnt = 2 * ((npairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);  // units of this CTA
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(bars + i, K::PTHREADS);
    for (int i = 6; i < 12; ++i) mbar_init(bars + i, 4);
    for (int i = 12; i < 15; ++i) mbar_init(bars + i, K::PTHREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();  // the rows are the previous kernel's output

  if (warp >= K::CWARPS) {
    // ---- producer warps: rows and G tiles, global -> shared, two units ahead ----
    const int pt = tid - K::CTHREADS;
    const int half = pt & 1, z0 = pt >> 1;           // 16-byte half of a row, rows z0 + 64 i
    auto refill = [&](int g, int n, int b) {
      const int64_t s = 2 * ((int64_t)blockIdx.x + (int64_t)(n >> 1) * gridDim.x) + (n & 1);
      const int kxt = (int)(s % p.ntx), ky = (int)(s / p.ntx);
      const float2* src = p.in + g * p.d_c + ky * p.d_by + kxt * K::TX + 2 * half + (int64_t)z0 * p.rs;
      float2* dst = stage + (b * 3 + g) * K::TILE + z0 * K::TX + 2 * half;
      // the eight source addresses formed first and pinned in distinct registers (an LDGSTS keeps its address pair
      // scoreboarded until the LSU has taken it)
      const float2* a[K::NIT];
#pragma unroll
      for (int i = 0; i < K::NIT; ++i) {
        a[i] = src + (size_t)i * (size_t)(64 * p.rs);
        asm volatile("" : "+l"(a[i]));
      }
#pragma unroll
      for (int i = 0; i < K::NIT; ++i) async_copy16(dst + i * 64 * K::TX, a[i]);
      mbar_arrive_on_copies(bars + 2 * g + b);
    };
    auto load_green = [&](int n, int j) {  // G tile of the n-th unit -> G buffer j (column pitch GP -> GPQ)
      const int64_t s = 2 * ((int64_t)blockIdx.x + (int64_t)(n >> 1) * gridDim.x) + (n & 1);
      const int kxt = (int)(s % p.ntx), ky = (int)(s / p.ntx);
      const int fy = ky <= p.n2y / 2 ? ky : p.n2y - ky;
      const float* gsrc = p.gt + ((int64_t)fy * p.ntx + kxt) * (K::TX * K::GP);
      float* gd = gbuf + j * (K::TX * K::GPQ);
      for (int i = pt; i < K::G_CHUNKS; i += K::PTHREADS) {
        const int col = i / (K::GP / 4), w = i % (K::GP / 4);
        async_copy16(gd + col * K::GPQ + 4 * w, gsrc + 4 * i);
      }
      mbar_arrive_on_copies(bars + 12 + j);
    };
    for (int n = 0; n < 2 && n < cnt; ++n) {
      for (int g = 0; g < 3; ++g) refill(g, n, n);
      load_green(n, n);
    }
    int j2 = 2;  // (n + 2) % 3
    for (int n = 0; n < cnt; ++n) {
      const int b = n & 1;
      for (int g = 0; g < 3; ++g) {
        mbar_wait(bars + 6 + 2 * g + b, (n >> 1) & 1);  // quartet g is past the first pass of unit n
        if (n + 2 < cnt) {
          refill(g, n + 2, b);
          // a quartet past the first pass of unit n has finished the middle pass of unit n - 1: once all three are,
          // G buffer (n + 2) % 3 = (n - 1) % 3 is free
          if (g == 2) load_green(n + 2, j2);
        }
      }
      j2 = j2 == 2 ? 0 : j2 + 1;
    }
  } else {
    // ---- consumer warps: quartet g = warp / 4 owns component g; lane (jq, c) of its warp w works on butterfly / block
    // t = 8 w + jq of column c ----
    const int g = warp >> 2, w = warp & 3, c = lane & 3, t = 8 * w + (lane >> 2);
    RowAcc<L> sm{work + (g * K::TX + c) * K::PQ};
    const float2 wj = p.tw[t];  // exp(-2 pi i t / L): base of this thread's twiddles in the first and the last pass
    int j = 0, jpar = 0;        // n % 3, (n / 3) & 1
#pragma unroll 1
    for (int n = 0; n < cnt; ++n) {
      const int b = n & 1;
      if (n == cnt - 1) pdl_launch_dependents();  // last unit: the next kernel may start filling the tail
      // opaque copy: keeps the compiler from hoisting the loop-invariant twiddle powers of both passes out of the loop,
      // where they would live in local memory (16 reloads per pass through the very pipe that bounds this kernel)
      float2 wl = wj;
      asm volatile("" : "+f"(wl.x), "+f"(wl.y));
      mbar_wait(bars + 2 * g + b, (n >> 1) & 1);  // rows of this unit's component have landed
      zrow_fwd_first<L>(QuadLoad{stage + (b * 3 + g) * K::TILE + c}, sm, t, wl);
      quartet_sync(g);
      if (lane == 0) mbar_arrive(bars + 6 + 2 * g + b);  // the stage tile may be refilled
      mbar_wait(bars + 12 + j, jpar);                    // G tile
      zrow_mid<L>(sm, t, gbuf + j * (K::TX * K::GPQ) + c * K::GPQ);
      quartet_sync(g);
      const int64_t s = 2 * ((int64_t)blockIdx.x + (int64_t)(n >> 1) * gridDim.x) + (n & 1);
      const int kxt = (int)(s % p.ntx), ky = (int)(s / p.ntx);
      const int kx = kxt * K::TX + c;
      asm volatile("" : "+f"(wl.x), "+f"(wl.y));
      zrow_inv_last<L>(sm, t, wl, TileStore{p.out + g * p.o_c + (kx >> 3) * p.o_bx8 + ky * p.o_by + (kx & 7), p.o_bz8});
      // no barrier here: the next first pass of THIS thread overwrites exactly the positions (k S + t of column c) it has
      // just read, nobody else's
      if (++j == 3) j = 0, jpar ^= 1;
    }
  }
}
#endif  // __CUDACC__

}  // namespace p2
}  // namespace sopht
