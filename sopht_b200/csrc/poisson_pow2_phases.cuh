// Phase functions of the pruned, fused FFT pipeline for the unbounded Poisson solve (fp32, power-of-two
// grids). Each kernel is a sequence of PHASES separated by __syncthreads(); a phase reads shared/global
// memory, works in registers and writes shared/global memory, so the same code is emulated thread by thread
// on the CPU (tests/host/fft_emul.cu) to pin the index arithmetic before it runs on the GPU.
//
// Pipeline for rhs (C, nz, ny, nx) real, doubled domain (2nz, 2ny, 2nx) never materialised:
//   XFwd : rows (c,z,y): real row of nx (+ nx implicit zeros) -> half-length complex FFT (length nx, upper
//          half of its input zero) + Hermitian post-processing -> A[c][z][y][kx<nx], nyqA[c][z][y] (kx = nx)
//   YFwd : (c,z,kx-tile): ny rows (+ ny zeros) -> FFT length 2ny -> B[c][z][ky<2ny][kx]
//   ZConv: (ky,kx-tile), loop c: nz rows (+ nz zeros) -> FFT 2nz -> x G_hat (real, even: folded storage) ->
//          inverse FFT 2nz -> first nz rows back in place
//   YInv : (c,z<nz,kx-tile): 2ny rows -> inverse FFT -> first ny rows -> A
//   XInv : rows: Hermitian pre-processing -> inverse half-length FFT -> first nx reals -> solution
// The kx = nx (Nyquist) plane travels as a separate small (C, nz, ny) array through the same Y/Z kernels
// with different strides. ref: UnboundedPoissonSolverPYFFTW3D.py:111-149 (what is computed).
//
// Latency hiding: kernels are persistent (one CTA per SM slot loops over tiles). While a tile is in its
// second..last phase, the FIRST-phase inputs of the CTA's next tile are already in flight as cp.async copies
// into a staging buffer (`stage`), issued by the very threads that will consume them, so no registers are
// held across the wait. prefetch() and phase<0>() must enumerate the same elements.
#pragma once
#include <stdint.h>

#include "fft_tile.cuh"

namespace sopht {
namespace p2 {

using fft::Cfg;

// ---- shared-memory accessors ------------------------------------------------------------------------------
// column mode: TX sequences side by side, column index fastest; a skew row every RLAST rows keeps the
// stride-RLAST accesses of the last pass off the same banks.
template <int L, int TX>
struct ColAcc {
  float2* s;
  int col;
  static constexpr int PADR = Cfg<L>::RLAST;
  static constexpr int ROWS = L + L / PADR;
  FFT_HD float2& operator()(int e) const { return s[(e + e / PADR) * TX + col]; }
  // position c + v where c is a compile-time constant after unrolling and (c % PADR) + (v % PADR) < PADR,
  // so the padding term splits and the constant part folds into the instruction's immediate offset
  FFT_HD float2& at(int c, int v) const { return s[(v + v / PADR) * TX + col + (c + c / PADR) * TX]; }
};
// row mode: each sequence contiguous, one pad element every RLAST elements
template <int L>
struct RowAcc {
  float2* s;
  static constexpr int PADR = Cfg<L>::RLAST;
  static constexpr int PITCH = L + L / PADR + 1;
  FFT_HD float2& operator()(int e) const { return s[e + e / PADR]; }
  FFT_HD float2& at(int c, int v) const { return s[(v + v / PADR) + (c + c / PADR)]; }
};

// ---- shared-memory tables --------------------------------------------------------------------------------
// The twiddle table (fft::TwTable: per-pass, bank-conflict-free order) and, in the z pass, the tile's slice of the
// folded Green's function live in shared memory right behind the data tile:
// smem = [data SMEM_ELEMS][tables EXTRA_ELEMS][staging]. Read from global memory they cost one LDG with a 64-bit
// address register pair per use, and those registers stay scoreboarded until the load has read them - measured
// as the top stall of the 255-register radix-32 kernels. init() fills the twiddle part once per (persistent) CTA.
using fft::TwTable;

template <int THREADS>
FFT_HD void copy_table(float2* dst, const float2* src, int n, int tid) {
  for (int i = tid; i < n; i += THREADS) dst[i] = src[i];
}

// ---- Y forward / inverse (column mode) --------------------------------------------------------------------
struct ColParams {
  const float2* in;
  float2* out;
  int64_t in_rs, in_cs, out_rs, out_cs;          // row / column strides (complex elements)
  // input tile origin = in + bx*in_bx + (by >> log2_bz)*in_bc + (by & (2^log2_bz - 1))*in_by, by = component*nz + z
  // (two terms because the tile-major spectrum the y inverse reads is not linear in by); output: out + bx*out_bx + by*out_by
  int64_t in_bx, in_by, in_bc, out_bx, out_by, out_bc;
  int log2_bz;
  // in_by8 != 0: the input is blocked in z (planes z / 8 are in_by8 apart, planes z % 8 in_by apart inside a block): the
  // z-blocked tile-major spectrum the warp-quartet z pass writes
  int64_t in_by8;
  const float2* tw;                              // forward twiddles, length L
  FFT_HD const float2* in_plane(int by) const {
    const int z = by & ((1 << log2_bz) - 1);
    const int64_t zoff = in_by8 ? (z >> 3) * in_by8 + (z & 7) * in_by : z * in_by;
    return in + (by >> log2_bz) * in_bc + zoff;
  }
  FFT_HD float2* out_plane(int by) const {
    return out + (by >> log2_bz) * out_bc + (by & ((1 << log2_bz) - 1)) * out_by;
  }
};

// Row strides travel as 32-bit unsigned: for every eligible grid (nz, ny <= 1024, nx <= 2048) the largest
// element offset inside a tile column, (L - 1) * rs, stays below 2^32, and an unsigned 32-bit product plus one
// IMAD.WIDE replaces a 64-bit multiply per access.
struct GlobalLoad {
  const float2* p;
  unsigned rs;
  FFT_HD float2 operator()(int e) const { return p[(unsigned)e * rs]; }
};
struct GlobalStoreIdx {  // sink(k, pos, v) -> out[k]
  float2* p;
  unsigned rs;
  FFT_HD void operator()(int k, int, float2 v) const { p[(unsigned)k * rs] = v; }
};
struct GlobalSrcIdx {  // src(k, pos) -> in[k]
  const float2* p;
  unsigned rs;
  FFT_HD float2 operator()(int k, int) const { return p[(unsigned)k * rs]; }
};
struct GlobalStore {
  float2* p;
  unsigned rs;
  FFT_HD void operator()(int e, float2 v) const { p[(unsigned)e * rs] = v; }
};

// staged variants: the element was copied to stage[index * TX + col] by this thread's own cp.async
template <int TX>
struct StageLoad {
  const float2* s;
  FFT_HD float2 operator()(int e) const { return s[e * TX]; }
};
template <int TX>
struct StageSrcIdx {
  const float2* s;
  FFT_HD float2 operator()(int k, int) const { return s[k * TX]; }
};
template <int TX>
struct StageCopy {
  float2* s;
  const float2* g;
  unsigned rs;
  FFT_HD void operator()(int e) const { fft::async_copy8(s + e * TX, g + (unsigned)e * rs); }
};

// L2 prefetch of the element a thread will load in its next tile's first phase: one request per 32-byte sector
// (columns are 8 bytes apart, so every fourth column's thread asks). The y kernels run as two unstaged
// 128-register CTAs per SM and have neither registers nor shared memory left to hold a prefetched tile; pulling it
// into L2 costs neither and hides the DRAM part of the first-phase latency: y forward 2.02 -> 1.90 ms, y inverse
// 2.19 -> 1.98 ms at 512^3, x inverse 0.049 -> 0.042 ms at 128x128x256.
struct L2Prefetch {
  const float2* g;
  unsigned rs;
  bool lead;
  FFT_HD void operator()(int e) const {
    if (lead) fft::prefetch_l2(g + (unsigned)e * rs);
  }
};
#ifndef SOPHT_P2_NO_L2_PREFETCH
#define SOPHT_P2_L2_PREFETCH 1
#else
#define SOPHT_P2_L2_PREFETCH 0
#endif

template <int L, int TX, bool FULL = false>
struct YFwd {
  static constexpr int SYNC_THREADS = Cfg<L>::T * TX;  // whole CTA (narrower groups: see profiles/r01_poisson_layout_experiments.txt)
  using Params = ColParams;
  static constexpr int THREADS = Cfg<L>::T * TX;
  static constexpr int NPHASE = Cfg<L>::NP;
  static constexpr int NITER = 1;
  static constexpr int SMEM_ELEMS = ColAcc<L, TX>::ROWS * TX;
  static constexpr int EXTRA_ELEMS = TwTable<L>::SIZE;
  static constexpr int STAGE_ELEMS = (FULL ? L : L / 2) * TX;
  static constexpr bool STAGE_SHARED = false;  // a thread reads back only what it copied itself
  static constexpr bool WANT_STAGE = Cfg<L>::E >= 32;
  static constexpr bool DEFAULT_TWO_CTAS = true, DEFAULT_STAGED = false;
  FFT_HD static int niter(const Params&) { return 1; }
  FFT_HD static void init(const Params& p, int tid, float2* smem) {
    TwTable<L>::template fill<THREADS>(smem + SMEM_ELEMS, p.tw, tid);
  }
  FFT_HD static void prefetch(const Params& p, int bx, int by, int, int tid, float2* stage) {
    const int col = tid % TX, t = tid / TX;
    StageCopy<TX> cp{stage + col, p.in_plane(by) + bx * p.in_bx + col * p.in_cs, (unsigned)p.in_rs};
    fft::fwd_first_elems_x<L, FULL>(t, cp);
  }
  // unstaged variants: the next tile's first-phase inputs are pulled into L2 while this tile is transformed
  static constexpr bool L2_PREFETCH = SOPHT_P2_L2_PREFETCH;
  FFT_HD static void l2_prefetch(const Params& p, int bx, int by, int, int tid) {
    const int col = tid % TX, t = tid / TX;
    L2Prefetch pf{p.in_plane(by) + bx * p.in_bx + col * p.in_cs, (unsigned)p.in_rs,
                  (col & 3) == 0 && p.in_cs == 1};
    fft::fwd_first_elems_x<L, FULL>(t, pf);
  }
  template <int P>
  FFT_HD static void phase(const Params& p, int bx, int by, int, int tid, float2* smem, const float2* stage) {
    const int col = tid % TX, t = tid / TX;
    ColAcc<L, TX> sm{smem, col};
    const float2* tw = smem + SMEM_ELEMS;
    if (P == 0) {
      if (stage) {
        fft::fwd_first_x<L, FULL>(StageLoad<TX>{stage + col}, sm, t, tw);
      } else {
        GlobalLoad ld{p.in_plane(by) + bx * p.in_bx + col * p.in_cs, (unsigned)p.in_rs};
        fft::fwd_first_x<L, FULL>(ld, sm, t, tw);
      }
    } else if (P == NPHASE - 1) {
      GlobalStoreIdx st{p.out_plane(by) + bx * p.out_bx + col * p.out_cs, (unsigned)p.out_rs};
      fft::fwd_last<L>(sm, t, st);
    } else {
      fft::fwd_mid<L>(sm, t, tw);
    }
  }
};

template <int L, int TX, bool FULL = false>
struct YInv {
  static constexpr int SYNC_THREADS = Cfg<L>::T * TX;  // whole CTA (narrower groups: see profiles/r01_poisson_layout_experiments.txt)
  using Params = ColParams;
  static constexpr int THREADS = Cfg<L>::T * TX;
  static constexpr int NPHASE = Cfg<L>::NP;
  static constexpr int NITER = 1;
  static constexpr int SMEM_ELEMS = ColAcc<L, TX>::ROWS * TX;
  static constexpr int EXTRA_ELEMS = TwTable<L>::SIZE;
  static constexpr int STAGE_ELEMS = L * TX;
  static constexpr bool STAGE_SHARED = false;
  static constexpr bool WANT_STAGE = Cfg<L>::E >= 32;
  static constexpr bool DEFAULT_TWO_CTAS = true, DEFAULT_STAGED = false;
  FFT_HD static int niter(const Params&) { return 1; }
  FFT_HD static void init(const Params& p, int tid, float2* smem) {
    TwTable<L>::template fill<THREADS>(smem + SMEM_ELEMS, p.tw, tid);
  }
  FFT_HD static void prefetch(const Params& p, int bx, int by, int, int tid, float2* stage) {
    const int col = tid % TX, t = tid / TX;
    StageCopy<TX> cp{stage + col, p.in_plane(by) + bx * p.in_bx + col * p.in_cs, (unsigned)p.in_rs};
    fft::inv_first_elems<L>(t, cp);
  }
  static constexpr bool L2_PREFETCH = SOPHT_P2_L2_PREFETCH;
  FFT_HD static void l2_prefetch(const Params& p, int bx, int by, int, int tid) {
    const int col = tid % TX, t = tid / TX;
    L2Prefetch pf{p.in_plane(by) + bx * p.in_bx + col * p.in_cs, (unsigned)p.in_rs,
                  (col & 3) == 0 && p.in_cs == 1};
    fft::inv_first_elems<L>(t, pf);
  }
  template <int P>
  FFT_HD static void phase(const Params& p, int bx, int by, int, int tid, float2* smem, const float2* stage) {
    const int col = tid % TX, t = tid / TX;
    ColAcc<L, TX> sm{smem, col};
    const float2* tw = smem + SMEM_ELEMS;
    if (P == 0) {
      if (stage) {
        fft::inv_first<L>(StageSrcIdx<TX>{stage + col}, sm, t);
      } else {
        GlobalSrcIdx src{p.in_plane(by) + bx * p.in_bx + col * p.in_cs, (unsigned)p.in_rs};
        fft::inv_first<L>(src, sm, t);
      }
    } else if (P == NPHASE - 1) {
      GlobalStore st{p.out_plane(by) + bx * p.out_bx + col * p.out_cs, (unsigned)p.out_rs};
      fft::inv_last_x<L, FULL>(sm, t, tw, st);
    } else {
      fft::inv_mid<L>(sm, t, tw);
    }
  }
};

// ---- Z: forward, Green's function multiply, inverse — in place ----------------------------------------------
struct ZParams {
  float2* data;            // input tile origin = data + bx*d_bx + by*d_by + c*d_c, rows rs apart
  int64_t rs, cs, d_bx, d_by, d_c;
  float2* out;             // output tile origin = out + bx*o_bx + by*o_by + c*o_c, rows o_rs apart (may be `data`)
  int64_t o_rs, o_bx, o_by, o_c;
  int ncomp;
  const float* g;          // folded G_hat: g[fold(kz)*g_zs + goff]
  int64_t g_zs;
  int64_t g_ky_stride;     // main: nx (goff = fold(ky)*nx + kx); nyquist plane: 1 (goff = fold(ky))
  int n2y;                 // 2*ny, for fold(ky)
  int nyq;                 // 0: columns are kx (ky = by); 1: columns are ky (ky = bx*TX + col)
  const float2* tw;
};

// G_hat(kz) for one column, kz given as (blk, klast) of the position the forward last pass leaves it at.
// Whether kz folds (kz > L/2 -> L - kz) depends on klast alone: the spectrum index is
// rev(blk) + (L/RLAST) * klast with rev(blk) < L/RLAST, and L/2 is a multiple of L/RLAST.
template <int L, int TX>
struct GreenTile {  // from the tile's shared-memory slice gs[f * TX + col], f = 0 .. L/2
  const float* gs;
  FFT_HD float operator()(int blk, int klast) const {
    const int kz = fft::spectrum_index<L>(blk, klast);
    const int f = klast < Cfg<L>::RLAST / 2 ? kz : L - kz;
    return gs[f * TX];
  }
};

template <int L, int TX>
struct ZConv {
  static constexpr int SYNC_THREADS = Cfg<L>::T * TX;  // whole CTA (narrower groups: see profiles/r01_poisson_layout_experiments.txt)
  using Params = ZParams;
  static constexpr int THREADS = Cfg<L>::T * TX;
  static constexpr int NP = Cfg<L>::NP;
  static constexpr int NPHASE = 2 * NP - 1;  // fwd_first [fwd_mid] fused [inv_mid] inv_last
  static constexpr int NITER = 0;            // runtime: ncomp
  static constexpr int SMEM_ELEMS = ColAcc<L, TX>::ROWS * TX;
  static constexpr int G_FLOATS = (L / 2 + 1) * TX;  // the tile's folded G_hat slice, shared by its components
  static constexpr int EXTRA_ELEMS = TwTable<L>::SIZE + (G_FLOATS + 1) / 2;
  static constexpr int STAGE_ELEMS = (L / 2) * TX;
  static constexpr bool STAGE_SHARED = false;
  static constexpr bool WANT_STAGE = Cfg<L>::E >= 32;
  // L = 1024 (256 threads): one staged 255-register CTA; L = 512 (128 threads): four unstaged CTAs
  static constexpr bool DEFAULT_TWO_CTAS = THREADS < 256, DEFAULT_STAGED = THREADS >= 256;
  FFT_HD static int niter(const Params& p) { return p.ncomp; }
  FFT_HD static void init(const Params& p, int tid, float2* smem) {
    TwTable<L>::template fill<THREADS>(smem + SMEM_ELEMS, p.tw, tid);
  }
  FFT_HD static void prefetch(const Params& p, int bx, int by, int c, int tid, float2* stage) {
    const int col = tid % TX, t = tid / TX;
    StageCopy<TX> cp{stage + col, p.data + bx * p.d_bx + by * p.d_by + c * p.d_c + col * p.cs, (unsigned)p.rs};
    fft::fwd_first_elems<L>(t, cp);
  }
  // measured (B200): the z pass loses with it (L = 512, four unstaged CTAs: 0.62 -> 0.77 ms at 256^3)
  static constexpr bool L2_PREFETCH = false;
  FFT_HD static void l2_prefetch(const Params& p, int bx, int by, int c, int tid) {
    const int col = tid % TX, t = tid / TX;
    L2Prefetch pf{p.data + bx * p.d_bx + by * p.d_by + c * p.d_c + col * p.cs, (unsigned)p.rs,
                  (col & 3) == 0 && p.cs == 1};
    fft::fwd_first_elems<L>(t, pf);
  }
  FFT_HD static int64_t green_offset(const Params& p, int bx, int by, int col) {
    const int ky = p.nyq ? bx * TX + col : by;
    const int fky = ky <= p.n2y / 2 ? ky : p.n2y - ky;
    return p.nyq ? (int64_t)fky : (int64_t)fky * p.g_ky_stride + bx * TX + col;
  }
  template <int P>
  FFT_HD static void phase(const Params& p, int bx, int by, int c, int tid, float2* smem, const float2* stage) {
    const int col = tid % TX, t = tid / TX;
    ColAcc<L, TX> sm{smem, col};
    const float2* base = p.data + bx * p.d_bx + by * p.d_by + c * p.d_c + col * p.cs;
    const float2* tw = smem + SMEM_ELEMS;
    float* gs = reinterpret_cast<float*>(smem + SMEM_ELEMS + TwTable<L>::SIZE);
    if (P == 0) {
      if (c == 0) {  // this tile's G_hat slice, asynchronously, behind the first butterflies
        const float* g = p.g + green_offset(p, bx, by, col);
        for (int f = t; f <= L / 2; f += Cfg<L>::T) fft::async_copy4(gs + f * TX + col, g + f * p.g_zs);
      }
      if (stage) {
        fft::fwd_first<L>(StageLoad<TX>{stage + col}, sm, t, tw);
      } else {
        GlobalLoad ld{base, (unsigned)p.rs};
        fft::fwd_first<L>(ld, sm, t, tw);
      }
      if (c == 0) fft::async_commit_wait_all();  // published by the barrier after this phase
    } else if (P == NP - 1) {
      fft::fwd_last_mul_inv_first<L>(sm, t, GreenTile<L, TX>{gs + col});
    } else if (P == NPHASE - 1) {
      GlobalStore st{p.out + bx * p.o_bx + by * p.o_by + c * p.o_c + col * p.cs, (unsigned)p.o_rs};
      fft::inv_last<L>(sm, t, tw, st);
    } else if (P < NP - 1) {
      fft::fwd_mid<L>(sm, t, tw);
    } else {
      fft::inv_mid<L>(sm, t, tw);
    }
  }
};

// ---- Z, periodic box: forward, x 1 / symbol, inverse - in place, nothing padded, nothing dropped ------------------------
// spectrum *= norm / (lx[kx] + ly[ky] + lz[kz]) (three 1-D tables: (2 pi m / L)^2 or the three-point symbol), mean mode -> 0
struct ZSymParams {
  float2* data;            // tile origin = data + bx*d_bx + by*d_by + c*d_c, rows rs apart, columns cs apart
  int64_t rs, cs, d_bx, d_by, d_c;
  int ncomp;
  const float *lz, *ly, *lx;
  float norm;
  int kx0;                 // global kx of column 0 of tile bx = 0 (a rank's kx slab)
  int nyq;                 // 0: columns are kx (ky = by); 1: columns are ky (ky = bx*TX + col), kx = kx_fixed
  int kx_fixed;
  const float2* tw;
};
template <int L>
struct SymCol {
  const float* lz;  // shared-memory copy, L entries
  float lyx, norm;
  FFT_HD float operator()(int blk, int klast) const {
    const int kz = fft::spectrum_index<L>(blk, klast);
    const float lam = lz[kz] + lyx;
    return lam > 0.f ? norm / lam : 0.f;
  }
};
template <int L, int TX>
struct ZSym {
  static constexpr int SYNC_THREADS = Cfg<L>::T * TX;
  using Params = ZSymParams;
  static constexpr int THREADS = Cfg<L>::T * TX;
  static constexpr int NP = Cfg<L>::NP;
  static constexpr int NPHASE = 2 * NP - 1;
  static constexpr int NITER = 0;
  static constexpr int SMEM_ELEMS = ColAcc<L, TX>::ROWS * TX;
  static constexpr int EXTRA_ELEMS = TwTable<L>::SIZE + L / 2;  // twiddles, lz (L floats)
  static constexpr int STAGE_ELEMS = L * TX;
  static constexpr bool STAGE_SHARED = false;
  static constexpr bool WANT_STAGE = Cfg<L>::E >= 32;
  static constexpr bool DEFAULT_TWO_CTAS = true, DEFAULT_STAGED = false;
  static constexpr bool L2_PREFETCH = SOPHT_P2_L2_PREFETCH;
  FFT_HD static int niter(const Params& p) { return p.ncomp; }
  FFT_HD static void init(const Params& p, int tid, float2* smem) {
    TwTable<L>::template fill<THREADS>(smem + SMEM_ELEMS, p.tw, tid);
    float* lz = reinterpret_cast<float*>(smem + SMEM_ELEMS + TwTable<L>::SIZE);
    for (int i = tid; i < L; i += THREADS) lz[i] = p.lz[i];
  }
  FFT_HD static const float2* col_ptr(const Params& p, int bx, int by, int c, int col) {
    return p.data + bx * p.d_bx + by * p.d_by + c * p.d_c + col * p.cs;
  }
  FFT_HD static void prefetch(const Params& p, int bx, int by, int c, int tid, float2* stage) {
    const int col = tid % TX, t = tid / TX;
    StageCopy<TX> cp{stage + col, col_ptr(p, bx, by, c, col), (unsigned)p.rs};
    fft::fwd_first_elems_x<L, true>(t, cp);
  }
  FFT_HD static void l2_prefetch(const Params& p, int bx, int by, int c, int tid) {
    const int col = tid % TX, t = tid / TX;
    L2Prefetch pf{col_ptr(p, bx, by, c, col), (unsigned)p.rs, (col & 3) == 0 && p.cs == 1};
    fft::fwd_first_elems_x<L, true>(t, pf);
  }
  template <int P>
  FFT_HD static void phase(const Params& p, int bx, int by, int c, int tid, float2* smem, const float2* stage) {
    const int col = tid % TX, t = tid / TX;
    ColAcc<L, TX> sm{smem, col};
    float2* base = const_cast<float2*>(col_ptr(p, bx, by, c, col));
    const float2* tw = smem + SMEM_ELEMS;
    const float* lz = reinterpret_cast<const float*>(smem + SMEM_ELEMS + TwTable<L>::SIZE);
    if (P == 0) {
      if (stage) {
        fft::fwd_first_x<L, true>(StageLoad<TX>{stage + col}, sm, t, tw);
      } else {
        GlobalLoad ld{base, (unsigned)p.rs};
        fft::fwd_first_x<L, true>(ld, sm, t, tw);
      }
    } else if (P == NP - 1) {
      const int ky = p.nyq ? bx * TX + col : by;
      const int kx = p.nyq ? p.kx_fixed : p.kx0 + bx * TX + col;
      fft::fwd_last_mul_inv_first<L>(sm, t, SymCol<L>{lz, p.ly[ky] + p.lx[kx], p.norm});
    } else if (P == NPHASE - 1) {
      GlobalStore st{base, (unsigned)p.rs};
      fft::inv_last_x<L, true>(sm, t, tw, st);
    } else if (P < NP - 1) {
      fft::fwd_mid<L>(sm, t, tw);
    } else {
      fft::inv_mid<L>(sm, t, tw);
    }
  }
};

// ---- X forward: real rows -> half spectrum (row mode) -----------------------------------------------------
struct XParams {
  const float* real_in;    // XFwd: rhs;  element (c,z,y,x) at c*sc + z*sz + y*sy + x
  float* real_out;         // XInv: solution
  int64_t sc, sz, sy;      // strides of the real field (floats)
  float2* spec;            // A: (rows, nx) complex
  float2* nyq;             // (rows) complex
  int nz, ny;              // row = (c*nz + z)*ny + y
  const float2* tw;        // forward twiddles, length L = nx
  const float2* tw2;       // exp(-2 pi i k / (2 nx)), k = 0..nx-1
  // Address of bin k of spectrum row `row` in `spec` (complex elements):
  //   (row >> rpc_shift) * comp_stride + (row & rpc_mask) * chunk_len + (k >> chunk_shift) * chunk_stride
  //   + (k & (chunk_len - 1)),   chunk_len = 1 << chunk_shift.
  // One GPU: chunk_len = L, comp_stride = rows_per_component * L (plain (rows, L) rows). Slab-decomposed
  // solve: spec is the all-to-all buffer (C, P, nz_local, ny, nx/P) - the kx range of rank r is chunk r -
  // so the transposes need no pack / unpack pass.
  // XFwd can write chunk q straight into rank q's exchange buffer over NVLink (peer memory): chunk[q] is then
  // that buffer's base and self_offset this rank's slot (r * chunk_stride) in it; without peers chunk[q] =
  // spec + q * chunk_stride and self_offset = 0.
  int rpc_shift, chunk_shift;
  int64_t comp_stride, chunk_stride, self_offset;
  float2* chunk[8];
  FFT_HD int64_t row_base(int64_t row) const {
    return (row >> rpc_shift) * comp_stride + ((row & (((int64_t)1 << rpc_shift) - 1)) << chunk_shift);
  }
  FFT_HD int64_t bin(int k) const {
    return (int64_t)(k >> chunk_shift) * chunk_stride + (k & ((1 << chunk_shift) - 1));
  }
  FFT_HD float2* out_bin(int64_t row_base_, int k) const {
    return chunk[k >> chunk_shift] + row_base_ + self_offset + (k & ((1 << chunk_shift) - 1));
  }
  // XInv reads through the same table: with peers, chunk[q] is the buffer rank q's y inverse left its kx-slab in
  // and self_offset this rank's z range in it, so the reverse transpose is a PULL of whole row chunks over NVLink
  // (nx / P contiguous bins per request; the y inverse's 64-byte segments would make poor NVLink stores).
  FFT_HD const float2* in_bin(int64_t row_base_, int k) const { return out_bin(row_base_, k); }
  int peer_read;  // chunk[] are peer mappings: no L2 prefetch (peer lines are not cached in the local L2)
};

template <int L>
struct InPlaceSink {
  RowAcc<L> sm;
  FFT_HD void operator()(int, int pos, float2 v) const { sm(pos) = v; }
};
template <int L>
struct InPlaceSrc {
  RowAcc<L> sm;
  FFT_HD float2 operator()(int, int pos) const { return sm(pos); }
};
struct RowLoad {
  const float2* p;
  FFT_HD float2 operator()(int e) const { return p[e]; }
};
struct RowStore {
  float2* p;
  FFT_HD void operator()(int e, float2 v) const { fft::store2(p + e, v); }
};

template <int L, int RX, bool FULL = false>
struct XFwd {
  using Params = XParams;
  static constexpr int T = Cfg<L>::T;
  static constexpr int THREADS = T * RX;
  // rows are independent transforms: the warps of a row synchronise among themselves only (-2 % at 512^3; the
  // same change made the inverse kernel 7 % slower, which therefore keeps the CTA-wide barrier)
  static constexpr int SYNC_THREADS = T >= 32 ? T : (THREADS >= 32 ? 32 : THREADS);
  static constexpr int NP = Cfg<L>::NP;
  static constexpr int NPHASE = NP + 1;
  static constexpr int NITER = 1;
  static constexpr int SMEM_ELEMS = RowAcc<L>::PITCH * RX;
  static constexpr int EXTRA_ELEMS = TwTable<L>::SIZE + L;  // twiddle table, tw2
  static constexpr int STAGE_ELEMS = (FULL ? L : L / 2) * RX;
  static constexpr bool STAGE_SHARED = false;
  static constexpr bool WANT_STAGE = false;
  static constexpr bool DEFAULT_TWO_CTAS = false, DEFAULT_STAGED = false;
  FFT_HD static int niter(const Params&) { return 1; }
  FFT_HD static void init(const Params& p, int tid, float2* smem) {
    TwTable<L>::template fill<THREADS>(smem + SMEM_ELEMS, p.tw, tid);
    copy_table<THREADS>(smem + SMEM_ELEMS + TwTable<L>::SIZE, p.tw2, L, tid);
  }
  FFT_HD static const float2* row_ptr(const Params& p, const float* base, int64_t row) {
    const int y = (int)(row % p.ny);
    const int64_t cz = row / p.ny;
    const int z = (int)(cz % p.nz);
    const int64_t c = cz / p.nz;
    return reinterpret_cast<const float2*>(base + c * p.sc + z * p.sz + y * p.sy);
  }
  FFT_HD static void prefetch(const Params& p, int bx, int, int, int tid, float2* stage) {
    const int t = tid % T, r = tid / T;
    StageCopy<1> cp{stage + r * (FULL ? L : L / 2), row_ptr(p, p.real_in, (int64_t)bx * RX + r), 1u};
    fft::fwd_first_elems_x<L, FULL>(t, cp);
  }
  // next tile's real rows (L / 2 float2 = L / 8 sectors per row, spread over the row's T threads);
  // measured: 2-5 % slower with it (contiguous rows: the hardware already streams them), so off
  static constexpr bool L2_PREFETCH = false;
  FFT_HD static void l2_prefetch(const Params& p, int bx, int, int, int tid) {
    const int t = tid % T, r = tid / T;
    const float2* row = row_ptr(p, p.real_in, (int64_t)bx * RX + r);
    for (int s = t; s < (FULL ? L / 4 : L / 8); s += T) fft::prefetch_l2(row + s * 4);
  }
  template <int P>
  FFT_HD static void phase(const Params& p, int bx, int, int, int tid, float2* smem, const float2* stage) {
    const int t = tid % T, r = tid / T;
    const int64_t row = (int64_t)bx * RX + r;
    RowAcc<L> sm{smem + r * RowAcc<L>::PITCH};
    const float2* tw = smem + SMEM_ELEMS;
    const float2* tw2 = smem + SMEM_ELEMS + TwTable<L>::SIZE;
    if (P == 0) {
      if (stage) {
        fft::fwd_first_x<L, FULL>(StageLoad<1>{stage + r * (FULL ? L : L / 2)}, sm, t, tw);
      } else {
        RowLoad ld{row_ptr(p, p.real_in, row)};
        fft::fwd_first_x<L, FULL>(ld, sm, t, tw);
      }
    } else if (P == NP - 1) {
      fft::fwd_last<L>(sm, t, InPlaceSink<L>{sm});
    } else if (P == NP) {
      // X_k = E_k + w^k O_k, E = (Z_k + conj Z_{L-k})/2, O = (Z_k - conj Z_{L-k})/(2i); X_L = E_0 - O_0.
      // Bins k and L - k come from the same pair: X_{L-k} = conj(E_k - w^k O_k) (E_{L-k} = conj E_k, O_{L-k} =
      // conj O_k, w^{L-k} = -conj w^k), so a thread owns k < L/2 and produces both - half the shared-memory reads and
      // twiddle products of the one-bin-per-read form; k = L/2 (its own partner: X = conj Z) rides with thread 0.
      const int64_t rb = p.row_base(row);
#pragma unroll
      for (int q = 0; q < Cfg<L>::E / 2; ++q) {
        const int k = t + q * T;
        const float2 a = sm(fft::spectrum_position<L>(k));
        const float2 b = sm(fft::spectrum_position<L>((L - k) & (L - 1)));
        const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
        const float2 o = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
        const float2 w = fft::cmul(o, tw2[k]);
        *p.out_bin(rb, k) = fft::cadd(e, w);
        if (k == 0)
          p.nyq[row] = make_float2(e.x - o.x, e.y - o.y);
        else
          *p.out_bin(rb, L - k) = make_float2(e.x - w.x, w.y - e.y);
      }
      if (t == 0) {
        const float2 a = sm(fft::spectrum_position<L>(L / 2));
        *p.out_bin(rb, L / 2) = make_float2(a.x, -a.y);
      }
    } else {
      fft::fwd_mid<L>(sm, t, tw);
    }
  }
};

template <int L, int RX, bool FULL = false>
struct XInv {
  using Params = XParams;
  static constexpr int T = Cfg<L>::T;
  static constexpr int THREADS = T * RX;
  static constexpr int SYNC_THREADS = THREADS;
  static constexpr int NP = Cfg<L>::NP;
  static constexpr int NPHASE = NP + 1;
  static constexpr int NITER = 1;
  static constexpr int SMEM_ELEMS = RowAcc<L>::PITCH * RX;
  static constexpr int EXTRA_ELEMS = TwTable<L>::SIZE + L;  // twiddle table, tw2
  static constexpr int STAGE_ELEMS = (L + 1) * RX;  // spectrum row + its Nyquist bin
  static constexpr bool STAGE_SHARED = true;        // bin k is combined with bin L-k, staged by another thread
  static constexpr bool WANT_STAGE = false;
  static constexpr bool DEFAULT_TWO_CTAS = false, DEFAULT_STAGED = false;
  FFT_HD static int niter(const Params&) { return 1; }
  FFT_HD static void init(const Params& p, int tid, float2* smem) { XFwd<L, RX, FULL>::init(p, tid, smem); }
  FFT_HD static void prefetch(const Params& p, int bx, int, int, int tid, float2* stage) {
    const int t = tid % T, r = tid / T;
    const int64_t row = (int64_t)bx * RX + r;
    const int64_t rb = p.row_base(row);
    float2* s = stage + r * (L + 1);
#pragma unroll
    for (int q = 0; q < Cfg<L>::E; ++q) fft::async_copy8(s + t + q * T, p.in_bin(rb, t + q * T));
    if (t == 0) fft::async_copy8(s + L, p.nyq + row);
  }
  // next tile's spectrum rows (L float2 = L / 4 sectors per row; chunks are multiples of a sector)
  static constexpr bool L2_PREFETCH = SOPHT_P2_L2_PREFETCH;
  FFT_HD static void l2_prefetch(const Params& p, int bx, int, int, int tid) {
    const int t = tid % T, r = tid / T;
    if (p.peer_read) return;
    const int64_t rb = p.row_base((int64_t)bx * RX + r);
    for (int s = t; s < L / 4; s += T) fft::prefetch_l2(p.in_bin(rb, s * 4));
  }
  template <int P>
  FFT_HD static void phase(const Params& p, int bx, int, int, int tid, float2* smem, const float2* stage) {
    const int t = tid % T, r = tid / T;
    const int64_t row = (int64_t)bx * RX + r;
    RowAcc<L> sm{smem + r * RowAcc<L>::PITCH};
    const float2* tw = smem + SMEM_ELEMS;
    const float2* tw2 = smem + SMEM_ELEMS + TwTable<L>::SIZE;
    if (P == 0) {
      // Z_k = E_k + i O_k, E = (X_k + conj X_{L-k})/2, O = conj(w^k) (X_k - conj X_{L-k})/2
      const float2* in = stage ? stage + r * (L + 1) : p.spec + p.row_base(row);
      const float2* nq = stage ? stage + r * (L + 1) + L : p.nyq + row;
      // one chunk (single GPU) or a staged row: bin k is element k, no chunk arithmetic per access
      const bool plain = stage != nullptr || (1 << p.chunk_shift) >= L;
      // Z_{L-k} = conj(E_k - i O_k) from the same pair (E_{L-k} = conj E_k, O_{L-k} = conj O_k): a thread owns
      // k < L/2 and writes both; k = L/2 (Z = conj X) rides with thread 0
      auto combine = [&](auto load) {
#pragma unroll
        for (int q = 0; q < Cfg<L>::E / 2; ++q) {
          const int k = t + q * T;
          const float2 a = load(k);
          const float2 b = k == 0 ? *nq : load(L - k);
          const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
          const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
          const float2 o = fft::cmul_conj(d, tw2[k]);
          sm(fft::spectrum_position<L>(k)) = make_float2(e.x - o.y, e.y + o.x);
          if (k != 0) sm(fft::spectrum_position<L>(L - k)) = make_float2(e.x + o.y, o.x - e.y);
        }
        if (t == 0) {
          const float2 a = load(L / 2);
          sm(fft::spectrum_position<L>(L / 2)) = make_float2(a.x, -a.y);
        }
      };
      if (plain) {
        combine([&](int k) { return in[k]; });
      } else {  // chunked (possibly peer) spectrum: through the chunk table
        const int64_t rb = p.row_base(row);
        combine([&](int k) { return *p.in_bin(rb, k); });
      }
    } else if (P == 1) {
      fft::inv_first<L>(InPlaceSrc<L>{sm}, sm, t);
    } else if (P == NPHASE - 1) {
      RowStore st{const_cast<float2*>(XFwd<L, RX, FULL>::row_ptr(p, p.real_out, row))};
      fft::inv_last_x<L, FULL>(sm, t, tw, st);
    } else {
      fft::inv_mid<L>(sm, t, tw);
    }
  }
};

// ---- parameter blocks of the pipeline, shared by the library (poisson_pow2.cu) and the host emulation ---------
// Whole grid on one GPU: P = 1. z-slab decomposition over P ranks (nz_local = nz / P planes per rank for the
// x passes, kx range of nx / P bins per rank for the y and z passes):
//   x forward (z-slab)  -> all-to-all per component ->  y forward, z convolution, y inverse (kx-slab)
//                       -> all-to-all back           ->  x inverse (z-slab)
// The kx = nx (Nyquist) plane is all-gathered and processed redundantly by every rank.
struct SlabDims {
  int C, nz, ny, nx, P, rank;
  FFT_HD int nzl() const { return nz / P; }
  FFT_HD int nxl() const { return nx / P; }
};
FFT_HD int ilog2(int64_t v) {
  int s = 0;
  while (((int64_t)1 << s) < v) ++s;
  return s;
}
// x passes on this rank's z-slab; `spec` is (C, P, nzl, ny, nxl) complex, `nyq` (C, nzl, ny)
inline XParams slab_x_params(const SlabDims& d, const float* real_in, float* real_out, int64_t sc, int64_t sz,
                             int64_t sy, float2* spec, float2* nyq, const float2* tw, const float2* tw2) {
  XParams xp{};
  xp.real_in = real_in;
  xp.real_out = real_out;
  xp.sc = sc, xp.sz = sz, xp.sy = sy;
  xp.spec = spec;
  xp.nyq = nyq;
  xp.nz = d.nzl();
  xp.ny = d.ny;
  xp.tw = tw;
  xp.tw2 = tw2;
  xp.rpc_shift = ilog2((int64_t)d.nzl() * d.ny);
  xp.chunk_shift = ilog2(d.nxl());
  xp.chunk_stride = (int64_t)d.nzl() * d.ny * d.nxl();
  xp.comp_stride = xp.chunk_stride * d.P;
  xp.self_offset = 0;
  for (int q = 0; q < 8; ++q) xp.chunk[q] = q < d.P ? spec + q * xp.chunk_stride : nullptr;
  return xp;
}
// the same with the spectrum chunks addressed in the peers' exchange buffers (peer[q]: base of rank q's buffer):
// x forward WRITES chunk q into rank q's buffer, x inverse READS chunk q from rank q's buffer
inline XParams slab_x_params_peer(XParams xp, const SlabDims& d, float2* const* peer) {
  xp.self_offset = (int64_t)d.rank * xp.chunk_stride;
  xp.peer_read = 1;
  for (int q = 0; q < 8; ++q) xp.chunk[q] = q < d.P ? peer[q] : nullptr;
  return xp;
}
// Two layouts of the doubled-in-y spectrum between the y and z passes:
//   B  (x-major, written by the y forward pass, read by the z pass):  (C, nz, 2ny, nxl)
//   B2 ("kx-tile-major", written by the z pass, read by the y inverse): element (c, z, ky, kx) at
//       ((((c * (nxl/TX) + kx/TX) * 2ny + ky) * nz + z) * TX + kx % TX
// In B2 the nz x TX tile a z-pass CTA produces is one contiguous block, written as a stream instead of nz rows
// 2ny * nxl * 8 bytes apart; the y inverse reads its 2ny rows nz * TX * 8 bytes apart with adjacent z planes next to
// each other (its L2 prefetch hides that stride). The y forward pass keeps writing the x-major B: its 64-byte
// stores lose more in the tile-major layout than the z pass gains (profiles/r01_poisson_layout_experiments.txt).
// y passes on this rank's kx-slab: a = (C, nz, ny, nxl); forward a -> B, inverse B2 -> a; grid (nxl / TX, C * nz)
// z8 (inverse only): the input is the z-blocked tile-major spectrum (C, nxl/8, nz/8, 2ny, 8 z, 8 kx) - a ky row of a
// tile is 64 bytes, ky rows 512 bytes apart, and the eight z planes of a block (eight neighbouring CTAs of the y inverse
// pass) read one contiguous 512 KB region together - instead of (C, nxl/8, 2ny, nz, 8 kx) with ky rows nz * 64 bytes apart
inline ColParams slab_y_params(const SlabDims& d, int TX, const float2* in, float2* out, bool forward,
                               const float2* tw, bool z8 = false) {
  const int64_t nxl = d.nxl(), LY = 2 * d.ny;
  ColParams yp{};
  yp.in = in;
  yp.out = out;
  yp.in_cs = 1, yp.out_cs = 1;
  yp.log2_bz = ilog2(d.nz);
  if (forward) {
    yp.in_rs = nxl, yp.in_bx = TX, yp.in_by = (int64_t)d.ny * nxl, yp.in_bc = (int64_t)d.nz * d.ny * nxl;
    yp.out_rs = nxl, yp.out_bx = TX, yp.out_by = LY * nxl, yp.out_bc = (int64_t)d.nz * LY * nxl;
  } else {
    yp.in_rs = (int64_t)d.nz * TX, yp.in_bx = LY * d.nz * TX, yp.in_by = TX, yp.in_bc = LY * d.nz * nxl;
    if (z8) yp.in_rs = 8 * TX, yp.in_by8 = LY * 8 * TX;  // tile and component strides are the same in both layouts
    yp.out_rs = nxl, yp.out_bx = TX, yp.out_by = (int64_t)d.ny * nxl, yp.out_bc = (int64_t)d.nz * d.ny * nxl;
  }
  yp.tw = tw;
  return yp;
}
// Nyquist plane (C, nz, ny) <-> (C, nz, 2ny): columns are the (c, z) index; grid (C * nz / TX, 1)
inline ColParams nyquist_y_params(const SlabDims& d, int TX, const float2* in, float2* out, bool forward,
                                  const float2* tw) {
  const int64_t LY = 2 * d.ny;
  ColParams yn{};
  yn.in = in;
  yn.out = out;
  yn.in_rs = 1, yn.out_rs = 1;
  yn.in_cs = forward ? d.ny : LY;
  yn.out_cs = forward ? LY : d.ny;
  yn.in_bx = TX * yn.in_cs, yn.out_bx = TX * yn.out_cs;  // by = 0: no plane offsets
  yn.tw = tw;
  return yn;
}
// z pass on the kx-slab: reads the x-major b = (C, nz, 2ny, nxl), writes the kx-tile-major b2 (C, nxl/TX, 2ny, nz, TX);
// gm is the folded G_hat stored as (nz+1, ny+1, g_row) whose column g_kx0 is this rank's first kx bin (whole
// spectrum: g_row = nx, g_kx0 = rank * nxl; a per-rank slice: g_row = nxl, g_kx0 = 0); grid (nxl / TX, 2ny)
inline ZParams slab_z_params(const SlabDims& d, int TX, float2* b, float2* b2, const float* gm, int g_row,
                             int g_kx0, const float2* tw) {
  const int64_t nxl = d.nxl(), LY = 2 * d.ny;
  ZParams zp{};
  zp.data = b;
  zp.rs = LY * nxl, zp.cs = 1, zp.d_bx = TX, zp.d_by = nxl, zp.d_c = (int64_t)d.nz * LY * nxl;
  zp.out = b2;
  zp.o_rs = TX, zp.o_by = (int64_t)d.nz * TX, zp.o_bx = LY * zp.o_by, zp.o_c = LY * d.nz * nxl;
  zp.ncomp = d.C;
  zp.g = gm + g_kx0;
  zp.g_zs = (int64_t)(d.ny + 1) * g_row;
  zp.g_ky_stride = g_row;
  zp.n2y = (int)LY;
  zp.nyq = 0;
  zp.tw = tw;
  return zp;
}
// z pass on the Nyquist plane (C, nz, 2ny); gn is (nz+1, ny+1); grid (2ny / TX, 1)
inline ZParams nyquist_z_params(const SlabDims& d, int TX, float2* b, const float* gn, const float2* tw) {
  const int64_t LY = 2 * d.ny;
  ZParams zn{};
  zn.data = b;
  zn.rs = LY, zn.cs = 1, zn.d_bx = TX, zn.d_by = 0, zn.d_c = (int64_t)d.nz * LY;
  zn.out = b, zn.o_rs = zn.rs, zn.o_bx = zn.d_bx, zn.o_by = 0, zn.o_c = zn.d_c;  // in place
  zn.ncomp = d.C;
  zn.g = gn;
  zn.g_zs = d.ny + 1;
  zn.g_ky_stride = 1;
  zn.n2y = (int)LY;
  zn.nyq = 1;
  zn.tw = tw;
  return zn;
}

}  // namespace p2
}  // namespace sopht
