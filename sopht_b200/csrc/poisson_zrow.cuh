// z pass of the pruned FFT pipeline, "one column per thread group" form (replaces p2::ZConv for the transform lengths
// that dominate the run time: 2 nz = 512, 1024, three components).
//
// Why a second form. p2::ZConv keeps a (L x 8 columns) tile in column mode: the threads of a warp own butterflies of
// eight different columns, every pass ends in a CTA-wide barrier, and all warps therefore move through LDS burst ->
// butterflies -> STS burst in lock step. ncu (profiles/r01_ncu_poisson_u512_v4.txt): issue slots 56 % busy, the
// shared-memory pipe and the FP32 pipe are each idle while the other works, 8 warps per SM. Here a column (one z
// sequence of one component at one (ky, kx)) belongs to ONE thread group of T = L / E <= 32 threads - a warp (L = 1024)
// or half a warp (L = 512) - in row mode (fft_tile.cuh passes over p2::RowAcc), so the only synchronisation between
// the passes of a transform is __syncwarp(). The 12 warps of a CTA (168 registers each: three warps per scheduler
// instead of two) drift apart and the shared-memory phases of one warp overlap the butterflies of the others.
//
// Unit of work = one (ky, kx tile of TX columns) x three components = 3 TX columns = one column per thread group.
// Global I/O is decoupled from the transforms:
//   * the rows of a unit (TX complex = 32 or 64 bytes each, 2ny * nx * 8 bytes apart) are fetched by cp.async
//     (LDGSTS, 8 bytes per lane, lanes = rows x TX columns so that every request covers whole 32-byte sectors) straight
//     into a column-major stage buffer (the transpose happens in the copy); an mbarrier per buffer (cp.async.mbarrier.
//     arrive.noinc) publishes it; two buffers: unit u + 1 is in flight while unit u is transformed;
//   * the inverse last pass writes its nz results back into the unit's own stage column; after a second mbarrier
//     ("all twelve warps are done with unit u") the threads stream the buffer out with the same lane mapping (whole
//     sectors) and refill the very addresses they just read with the rows of unit u + 2 - no barrier in between;
//   * a warp waits for "unit u - 1 done" only after its first pass of unit u, so nobody idles at a barrier: the
//     slowest warp of unit u - 1 has one whole pass of slack.
// The folded Green's function is stored tile-major for this kernel (gt: per (fold(ky), kx tile) a contiguous block
// [TX][GP] floats, z fastest), fetched with 16-byte cp.async into one of two buffers together with the unit's rows.
//
// Shared memory (L = 1024: TX = 4, 12 columns): stage 2 x 12 x 516 x 8 B = 96.8 KB, work 12 x 1057 x 8 B = 99.1 KB,
// twiddles 8 KB, G 2 x 4 x 516 x 4 B = 16.1 KB: 220 KB of the 227 KB a CTA may use; one persistent CTA per SM.
#pragma once
#include "poisson_pow2_phases.cuh"

namespace sopht {
namespace p2 {

struct ZRowParams {
  const float2* in;        // x-major spectrum B: element (c, z, ky, kx) at c*d_c + z*rs + ky*d_by + kx
  int64_t rs, d_c, d_by;
  float2* out;             // kx-tile(8)-major spectrum B2: (c, z, ky, kx) at c*o_c + (kx/8)*o_bx8 + ky*o_by + z*8 + kx%8
  int64_t o_c, o_bx8, o_by;
  int64_t o_bz8;           // warp-quartet kernel: distance of the z / 8 blocks (64 = the plain layout above; 2ny * 64: z-blocked)
  const float* gt;         // tile-major folded G_hat: block (fy * ntx + kxt) of TX * GP floats, [col][kq][r][4]:
                           // element (kq, r, i) = G_hat(fz = r + NBLK * (4 kq + i)), r <= NBLK (see ZRow::GP)
  int ntx;                 // kx tiles of TX columns (this rank's kx range)
  int n2y;                 // 2 ny
  const float2* tw;        // exp(-2 pi i j / L)
};

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// completion of this thread's earlier cp.async copies counts as one arrival (the barrier's count includes it)
__device__ __forceinline__ void mbar_arrive_on_copies(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ZROW_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra ZROW_DONE;\n"
      "bra ZROW_WAIT;\n"
      "ZROW_DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void async_copy16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}

template <int L>
struct ZRow {
  using C = Cfg<L>;
  static constexpr int T = C::T;                 // threads per column (<= 32)
  static constexpr int CWARPS = 12, CTHREADS = 32 * CWARPS;  // consumer warps: the transforms
  static constexpr int PWARPS = 4, PTHREADS = 32 * PWARPS;   // producer warps: all global memory traffic
  // 16 warps = 4 per SM sub-partition, whose 16384 registers allow 128 per thread (any count above 12 warps does)
  static constexpr int THREADS = CTHREADS + PTHREADS;
  static constexpr int NCOL = CTHREADS / T;      // columns per unit (three components)
  static constexpr int TX = NCOL / 3;            // kx columns per component: 4 (L = 1024) or 8 (L = 512)
  static constexpr int H = L / 2;                // rows read and written per column (the other half is zero / dropped)
  static constexpr int SP = H + 16 / TX;         // stage column pitch: (col * SP + dz) mod 16 distinct over a half warp
  static constexpr int PITCH = RowAcc<L>::PITCH;
  // G_hat of a column in "consumer order": the thread that owns block blk of the fused middle pass needs fz = blk +
  // NBLK k (k < K2) and, folded, fz = (NBLK - blk) + NBLK j (j < K2): rows blk and NBLK - blk of A[r][k] = G(r + NBLK k),
  // r = 0 .. NBLK. Stored [k / 4][r][4] so that both rows are K2 / 4 conflict-free 16-byte loads.
  static constexpr int NBLK = L / C::RLAST, K2 = C::RLAST / 2;
  static constexpr int GP = (K2 / 4) * (NBLK + 1) * 4;  // floats per G column
  static constexpr int RPI = 32 / TX;            // rows one warp moves per copy instruction
  static constexpr int NIT = H / (PWARPS * RPI); // copy instructions per producer thread, component and unit
  static constexpr int G_CHUNKS = TX * GP / 4;   // 16-byte chunks of a G tile
  static constexpr size_t SMEM_BYTES =
      sizeof(float2) * (2 * NCOL * SP + NCOL * PITCH) + sizeof(float) * 3 * TX * GP + 128;
  static_assert(T <= 32 && 32 % T == 0, "a column must fit a warp");
  static_assert(NCOL % 3 == 0 && H % (PWARPS * RPI) == 0 && NIT % 8 == 0, "unit shape");
};

// Powers w^k, k = 1 .. R - 1, of a thread's base twiddle a = w (its butterfly index is fixed for the whole kernel), fed
// to f(k, w^k) in increasing k. Replaces the R - 1 shared-memory table reads per pass (a quarter of this kernel's
// shared-memory wavefronts, and every one of them a latency the three warps of a scheduler cannot hide) by R - 2 complex
// products. Every power is the product of two values at most three products deep (blocks of four: w^(4m) from
// squarings and one product, w^(4m + i) = w^(4m) w^i), so the rounding error stays at a few ulp instead of growing
// with k as it would for w^(k+1) = w^k w.
template <int R, class F>
__device__ __forceinline__ void twiddle_powers(float2 a, F f) {
  static_assert(R == 32 || R == 16, "block scheme written for radix 16 / 32");
  const float2 a2 = fft::cmul(a, a), a3 = fft::cmul(a2, a);
  f(1, a), f(2, a2), f(3, a3);
  const float2 a4 = fft::cmul(a2, a2), a8 = fft::cmul(a4, a4), a12 = fft::cmul(a8, a4);
  auto block = [&](int m, float2 b) {
    f(4 * m, b), f(4 * m + 1, fft::cmul(b, a)), f(4 * m + 2, fft::cmul(b, a2)), f(4 * m + 3, fft::cmul(b, a3));
  };
  block(1, a4), block(2, a8), block(3, a12);
  if (R == 32) {
    const float2 a16 = fft::cmul(a8, a8);
    block(4, a16), block(5, fft::cmul(a16, a4)), block(6, fft::cmul(a16, a8)), block(7, fft::cmul(a16, a12));
  }
}
// forward first pass / inverse last pass of fft_tile.cuh with the twiddles from twiddle_powers (one butterfly per thread)
template <int L, class Load>
__device__ __forceinline__ void zrow_fwd_first(Load ld, RowAcc<L> sm, int j, float2 a) {
  using C = Cfg<L>;
  constexpr int R = C::R1, S = L / R;
  static_assert(C::E == R, "one first-pass butterfly per thread");
  float2 v[R];
#pragma unroll
  for (int n = 0; n < R; ++n) v[n] = n < R / 2 ? ld(n * S + j) : make_float2(0.f, 0.f);
  fft::Dft<R>::run_half(v);
  sm.at(0, j) = v[0];
  twiddle_powers<R>(a, [&](int k, float2 w) { sm.at(k * S, j) = fft::cmul(v[k], w); });
}
template <int L, class Store>
__device__ __forceinline__ void zrow_inv_last(RowAcc<L> sm, int j, float2 a, Store st) {
  using C = Cfg<L>;
  constexpr int R = C::R1, S = L / R;
  float2 v[R];
#pragma unroll
  for (int k = 0; k < R; ++k) v[k] = sm.at(k * S, j);
  twiddle_powers<R>(a, [&](int k, float2 w) { v[k] = fft::cmul_conj(v[k], w); });
  fft::dft<R, true>(v);
#pragma unroll
  for (int n = 0; n < R / 2; ++n) st(n * S + j, v[n]);
}

// fused middle pass of a column in row mode: forward last pass -> x G_hat -> inverse first pass, with the RLAST values of
// G_hat fetched up front as 16-byte loads (fft::fwd_last_mul_inv_first issues one 4-byte load right before each use)
template <int L>
__device__ __forceinline__ void zrow_mid(RowAcc<L> sm, int t, const float* gcol) {
  using C = Cfg<L>;
  using K = ZRow<L>;
  constexpr int R = C::RLAST, NB = C::E / R;
  static_assert(C::NP == 2, "two-pass lengths only");
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int blk = t + q * C::T;
    float g[R];
    const float4* own = reinterpret_cast<const float4*>(gcol) + blk;
    const float4* par = reinterpret_cast<const float4*>(gcol) + (K::NBLK - blk);
#pragma unroll
    for (int kq = 0; kq < K::K2 / 4; ++kq) {
      const float4 a = own[kq * (K::NBLK + 1)], b = par[kq * (K::NBLK + 1)];
      g[4 * kq + 0] = a.x, g[4 * kq + 1] = a.y, g[4 * kq + 2] = a.z, g[4 * kq + 3] = a.w;
      g[R - 1 - (4 * kq + 0)] = b.x, g[R - 1 - (4 * kq + 1)] = b.y;
      g[R - 1 - (4 * kq + 2)] = b.z, g[R - 1 - (4 * kq + 3)] = b.w;
    }
    // keep these loads up here, in front of the butterflies that hide their latency (the compiler otherwise sinks
    // them down to the multiplies and every FMUL waits for shared memory)
#pragma unroll
    for (int k = 0; k < R; ++k) asm volatile("" : "+f"(g[k]));
    float2 v[R];
#pragma unroll
    for (int n = 0; n < R; ++n) v[n] = sm.at(n, blk * R);
    fft::dft<R, false>(v);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      v[k] = fft::cscale(v[k], g[k]);
    }
    fft::dft<R, true>(v);
#pragma unroll
    for (int n = 0; n < R; ++n) sm.at(n, blk * R) = v[n];
  }
}

template <int L>
__global__ void __launch_bounds__(ZRow<L>::THREADS, 1) zrow_kernel(const ZRowParams p, int nunits) {
  using K = ZRow<L>;
  extern __shared__ __align__(16) unsigned char zrow_smem[];
  float2* stage = reinterpret_cast<float2*>(zrow_smem);            // [2][NCOL][SP]
  float2* work = stage + 2 * K::NCOL * K::SP;                      // [NCOL][PITCH]
  float* gbuf = reinterpret_cast<float*>(work + K::NCOL * K::PITCH);  // [3][TX][GP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbuf + 3 * K::TX * K::GP);
  // mbarriers: full[g][b] = bars[2 g + b] (rows of component g, buffer b, have landed), done[g][b] = bars[6 + 2 g + b]
  // (the four warps of component g are through with buffer b), gfull[j] = bars[12 + j] (G tile buffer j has landed)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((int)blockIdx.x >= nunits) return;
  const int cnt = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // units of this CTA
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(bars + i, K::PTHREADS);
    for (int i = 6; i < 12; ++i) mbar_init(bars + i, K::CWARPS / 3);
    for (int i = 12; i < 15; ++i) mbar_init(bars + i, K::PTHREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();  // the rows are the previous kernel's output

  if (warp >= K::CWARPS) {
    // ---- producer warps: all global memory traffic, one component (= one group of four consumer warps) at a time ----
    const int pt = tid - K::CTHREADS, pw = pt >> 5;
    const int ccol = lane % K::TX;
    const int z0 = pw * K::RPI + lane / K::TX;       // rows z0 + it * PWARPS * RPI
    const unsigned in_step = (unsigned)(K::PWARPS * K::RPI * p.rs);
    auto refill = [&](int g, int n, int b) {  // rows of component g of this CTA's n-th unit -> buffer b
      const int64_t s = (int64_t)blockIdx.x + (int64_t)n * gridDim.x;
      const int kxt = (int)(s % p.ntx), ky = (int)(s / p.ntx);
      const float2* src = p.in + g * p.d_c + ky * p.d_by + kxt * K::TX + ccol + (int64_t)z0 * p.rs;
      float2* dst = stage + b * (K::NCOL * K::SP) + (g * K::TX + ccol) * K::SP + z0;
      // Eight copies per batch with their eight source addresses formed first and pinned in distinct registers: an
      // LDGSTS keeps its address pair scoreboarded until the LSU has taken it, so re-using one pair per copy (what
      // the compiler does with a single induction pointer) costs a long-scoreboard stall per copy.
#pragma unroll 1
      for (int h = 0; h < K::NIT; h += 8) {
        const float2* a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] = src + (size_t)(h + i) * in_step;
          asm volatile("" : "+l"(a[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) fft::async_copy8(dst + (h + i) * (K::PWARPS * K::RPI), a[i]);
      }
      mbar_arrive_on_copies(bars + 2 * g + b);
    };
    auto load_green = [&](int n, int j) {  // G tile of the n-th unit -> G buffer j
      const int64_t s = (int64_t)blockIdx.x + (int64_t)n * gridDim.x;
      const int kxt = (int)(s % p.ntx), ky = (int)(s / p.ntx);
      const int fy = ky <= p.n2y / 2 ? ky : p.n2y - ky;
      const float* gsrc = p.gt + ((int64_t)fy * p.ntx + kxt) * (K::TX * K::GP);
      float* gd = gbuf + j * (K::TX * K::GP);
      for (int i = pt; i < K::G_CHUNKS; i += K::PTHREADS) async_copy16(gd + 4 * i, gsrc + 4 * i);
      mbar_arrive_on_copies(bars + 12 + j);
    };
    auto drain = [&](int g, int n, int b) {  // results of component g of the n-th unit -> B2, whole sectors per request
      const int64_t s = (int64_t)blockIdx.x + (int64_t)n * gridDim.x;
      const int kxt = (int)(s % p.ntx), ky = (int)(s / p.ntx);
      const int kx = kxt * K::TX + ccol;
      float2* dst = p.out + g * p.o_c + (kx >> 3) * p.o_bx8 + ky * p.o_by + (kx & 7) + (int64_t)z0 * 8;
      const float2* st = stage + b * (K::NCOL * K::SP) + (g * K::TX + ccol) * K::SP + z0;
#pragma unroll 1
      for (int h = 0; h < K::NIT; h += 8) {  // eight loads in flight, then eight stores
        float2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = st[(h + i) * (K::PWARPS * K::RPI)];
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[(int64_t)(h + i) * (K::PWARPS * K::RPI * 8)] = v[i];
      }
    };
    for (int n = 0; n < 2 && n < cnt; ++n) {
      for (int g = 0; g < 3; ++g) refill(g, n, n);
      load_green(n, n);
    }
    int j2 = 2;  // (n + 2) % 3
    for (int n = 0; n < cnt; ++n) {
      const int b = n & 1;
      for (int g = 0; g < 3; ++g) {
        mbar_wait(bars + 6 + 2 * g + b, (n >> 1) & 1);  // the four warps of component g are done with unit n
        drain(g, n, b);
        if (n + 2 < cnt) {
          refill(g, n + 2, b);
          // every group has passed unit n - 1 (this loop has seen their done barriers), so G buffer (n + 2) % 3 =
          // (n - 1) % 3 is free
          if (g == 0) load_green(n + 2, j2);
        }
      }
      j2 = j2 == 2 ? 0 : j2 + 1;
    }
  } else {
    // ---- consumer warps: one column per thread group; the four warps of a component (one per scheduler) share their
    // barriers, and the three components drift apart (the producer serves them one after the other), so the
    // shared-memory bursts of one group overlap the butterflies of the others.
    // Measured alternatives (profiles/r02_zpass_experiments.txt): all twelve warps on one pair of barriers (they move in
    // lock step again: 6.9 ms at 512^3 against 6.0 ms); the three warps of a scheduler kept in phase by a named barrier
    // to spare its instruction cache (6.7 ms); the four butterflies of a column rolled through one copy of the DFT code
    // (instruction-fetch stalls 14 % -> 3 %, but 6 % more instructions: 6.07 ms).
    const int slot = tid / K::T, t = tid % K::T;  // this group's column of the unit, butterfly index in it
    const int col = slot % K::TX, g = warp >> 2;  // kx column inside the tile, component
    RowAcc<L> sm{work + slot * K::PITCH};
    const float2 wj = p.tw[t];  // exp(-2 pi i t / L): base of this thread's twiddles in the first and the last pass
    int j = 0, jpar = 0;  // n % 3, (n / 3) & 1
    // not unrolled: the three passes are ~2400 instructions already
#pragma unroll 1
    for (int n = 0; n < cnt; ++n) {
      const int b = n & 1;
      float2* scol = stage + b * (K::NCOL * K::SP) + slot * K::SP;
      mbar_wait(bars + 2 * g + b, (n >> 1) & 1);  // rows of this unit's component have landed
      zrow_fwd_first<L>(RowLoad{scol}, sm, t, wj);
      __syncwarp();
      mbar_wait(bars + 12 + j, jpar);  // and its G tile
      zrow_mid<L>(sm, t, gbuf + j * (K::TX * K::GP) + col * K::GP);
      __syncwarp();
      zrow_inv_last<L>(sm, t, wj, RowStore{scol});
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 6 + 2 * g + b);
      if (++j == 3) j = 0, jpar ^= 1;
    }
  }
}
#endif  // __CUDACC__

}  // namespace p2

// warp-quartet form of the same pass for 2 nz = 1024 (poisson_zquad.cuh; its own translation unit: packed FP32 math)
int launch_zquad(const p2::ZRowParams& p, int nunits, cudaStream_t st);
}  // namespace sopht
