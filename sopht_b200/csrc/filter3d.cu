// Fused "convolution" Laplacian filter (the dominant stencil cost of the rod case C3).
//
// The reference applies, per component and per direction d in (x, y, z):
//     buf = f;  repeat `order` times { flux = M_d buf  on the ring-1 interior (flux ring stays 0);  buf = flux };
//     f = f - flux
// with M_d g = 0.25 (-g[+1] - g[-1] + 2 g) along d: 5 + 4 order full-array passes per direction (25 at order 5,
// 225 for a vector field). Every grid line along d is an independent 1-D problem, so one kernel per direction keeps
// a line (x) or a line segment with an `order`-cell halo (y, z) in shared memory, iterates there and touches HBM
// once for the read and once for the write: 8 B / cell / direction, 24 B / cell / component = the compulsory traffic
// of SURVEY 8d ("+72 B" per vector field).
//   x : one warp per row, in place (a row is read completely before it is written)
//   y : f -> scratch, z : scratch -> f   (segments read their neighbours' cells, hence out of place)
// ref: sopht/numeric/eulerian_grid_ops/stencil_ops_3d/laplacian_filter_3d.py:58-80 (the three 1-D stencils),
//      :129-163 (convolution closure: order of copies, passes and the final saxpby)
#include "common.cuh"

namespace sopht {
namespace {

template <typename T>
__device__ __forceinline__ T filter_point(T minus, T centre, T plus) {
  return T(0.25) * (-plus - minus + T(2) * centre);
}

// Away from the two ends of a line the masks never act and `order` passes of the three-point stencil collapse into one
// symmetric (2 order + 1)-tap filter: coefficients of (-1/4, 1/2, -1/4) convolved with itself `order` times, c[j] for
// offsets +-j. Cells within `order` of a line end see the zeroed ring of the intermediate passes and keep the
// pass-by-pass evaluation.
constexpr int FIR_MAX_ORDER = 8;
constexpr int XPAD = 8;  // >= FIR_MAX_ORDER, a multiple of 4
struct FirTaps {
  double c[FIR_MAX_ORDER + 1];
};
inline FirTaps fir_taps(int order) {
  double cur[2 * FIR_MAX_ORDER + 1] = {0.0}, nxt[2 * FIR_MAX_ORDER + 1];
  cur[FIR_MAX_ORDER] = 1.0;
  for (int m = 0; m < order; ++m) {
    for (int j = 0; j <= 2 * FIR_MAX_ORDER; ++j) {
      const double lo = j > 0 ? cur[j - 1] : 0.0, hi = j < 2 * FIR_MAX_ORDER ? cur[j + 1] : 0.0;
      nxt[j] = 0.25 * (-hi - lo + 2.0 * cur[j]);
    }
    for (int j = 0; j <= 2 * FIR_MAX_ORDER; ++j) cur[j] = nxt[j];
  }
  FirTaps t;
  for (int j = 0; j <= FIR_MAX_ORDER; ++j) t.c[j] = cur[FIR_MAX_ORDER + j];
  return t;
}

// ---- x: whole rows in shared memory, one warp per batch of rows --------------------------------------------
// A warp keeps `rb` rows (about 512 cells) in flight so that short rows still put enough loads on the wire; rows are
// independent, the batch is just a longer index space with the ring mask applied per row. All components of a vector
// field go into one launch (in place: a batch is read completely before it is written).
template <typename T, int K>  // K = the order when the one-pass filter applies (1..FIR_MAX_ORDER), 0 = pass by pass only
__global__ void __launch_bounds__(256)
    filter_rows_x_kernel(T* f, int64_t sc, int64_t sz, int64_t sy, int ncomp, int nz, int ny, int nx, int order,
                         int rb, int vec4, FirTaps taps) {
  extern __shared__ __align__(16) unsigned char filter_smem_raw[];
  T* smem = reinterpret_cast<T*>(filter_smem_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int span = rb * nx;
  constexpr bool fir = K > 0;  // the host picks K > 0 only when nx >= 4 K
  // one-pass form: the originals of the batch + the K corrected cells at either end of every row; pass by pass: the
  // originals and two ping-pong copies
  // XPAD cells in front of and behind the originals: the 16-byte window reads of the first / last quads of the batch
  // stay inside the warp's block (what they fetch there is never used)
  const int per_warp = fir ? ((span + 2 * K * rb + 2 * XPAD + 3) & ~3) : 3 * span;
  T* orig = smem + (size_t)w * per_warp + (fir ? XPAD : 0);
  T* a = orig + span;   // fir: ends[(2 rr + right) K + t]
  T* b = a + span;      // pass by pass only
  T c[K + 1];
#pragma unroll
  for (int j = 0; j <= K; ++j) c[j] = T(taps.c[j]);
  const int64_t rows_per_comp = (int64_t)(nz - 2) * (ny - 2);  // rows on the y / z ring keep their values
  const int64_t rows = rows_per_comp * ncomp;
  const int64_t batches = (rows + rb - 1) / rb;
  for (int64_t bt = (int64_t)blockIdx.x * warps + w; bt < batches; bt += (int64_t)gridDim.x * warps) {
    const int64_t r0 = bt * rb;
    const int nrow = (int)(rows - r0 < rb ? rows - r0 : rb);
    auto row_ptr = [&](int64_t r) -> T* {
      const int64_t c = r / rows_per_comp, q = r - c * rows_per_comp;
      return f + c * sc + (1 + q / (ny - 2)) * sz + (1 + q % (ny - 2)) * sy;
    };
    // lane rr works out where row rr of the batch lives (64-bit divisions: once per row, not once per use)
    const unsigned long long my_row = lane < nrow ? (unsigned long long)row_ptr(r0 + lane) : 0ull;
    for (int rr = 0; rr < nrow; ++rr) {  // all loads of the batch are issued before anything waits on them
      const T* row = reinterpret_cast<const T*>(__shfl_sync(0xffffffffu, my_row, rr));
      if (fir && vec4) {  // float rows, nx a multiple of 4, 16-byte aligned: a lane moves quads
        for (int q = lane; q < nx / 4; q += 32)
          reinterpret_cast<float4*>(orig + rr * nx)[q] = reinterpret_cast<const float4*>(row)[q];
        continue;
      }
      for (int i = lane; i < nx; i += 32) {
        const T v = row[i];
        orig[rr * nx + i] = v;
        if (!fir) a[rr * nx + i] = v;
      }
    }
    __syncwarp();
    // pass by pass only where the line ends are felt: the first and last 2 K columns (what is computed there is
    // right for the outer K columns after K passes); everything else is one (2 K + 1)-tap filter of the originals
    if constexpr (fir) {
      // one lane per (row, line end): the 2 K cells next to the end live in registers, index 0 = the end cell (both
      // ends are the same problem mirrored, the stencil is symmetric); no shared-memory traffic, no barriers
      for (int job = lane; job < 2 * nrow; job += 32) {
        const int rr = job >> 1;
        const bool right = job & 1;
        const T* o = orig + rr * nx;
        T u[2 * K];
#pragma unroll
        for (int t = 0; t < 2 * K; ++t) u[t] = o[right ? nx - 1 - t : t];
#pragma unroll
        for (int m = 0; m < K; ++m) {
          T prev = u[0];
          u[0] = T(0);  // the end cell is on the ring: no flux, in every pass
#pragma unroll
          for (int t = 1; t < 2 * K; ++t) {
            const T cur = u[t];
            const T nxt = t + 1 < 2 * K ? u[t + 1] : T(0);  // beyond the window: only cells >= 2 K - m see it
            u[t] = filter_point(prev, cur, nxt);
            prev = cur;
          }
        }
        T* fl = a + job * K;
#pragma unroll
        for (int t = 0; t < K; ++t) fl[t] = u[t];
      }
      __syncwarp();
    } else {
      for (int m = 0; m < order; ++m) {
        for (int rr = 0; rr < nrow; ++rr)
          for (int i = lane; i < nx; i += 32) {
            const int e = rr * nx + i;
            b[e] = (i >= 1 && i <= nx - 2) ? filter_point(a[e - 1], a[e], a[e + 1]) : T(0);
          }
        __syncwarp();
        T* t = a;
        a = b;
        b = t;
      }
    }
    for (int rr = 0; rr < nrow; ++rr) {
      T* row = reinterpret_cast<T*>(__shfl_sync(0xffffffffu, my_row, rr));
      const T* o = orig + rr * nx;
      if (fir && vec4) {
        // a lane owns four consecutive cells: the 20-cell window [i0 - 8, i0 + 12) comes in as five 16-byte loads and
        // the (2 K + 1)-tap sums run on registers (12 instead of 22 instructions per cell), one 16-byte store
        for (int q = lane; q < nx / 4; q += 32) {
          const int i0 = 4 * q;
          float wv[20];
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            const float4 v = reinterpret_cast<const float4*>(o + i0 - 8)[t];
            wv[4 * t] = v.x, wv[4 * t + 1] = v.y, wv[4 * t + 2] = v.z, wv[4 * t + 3] = v.w;
          }
          float r[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            float flux = (float)c[0] * wv[8 + e];
#pragma unroll
            for (int j = 1; j <= K; ++j) flux += (float)c[j] * (wv[8 + e - j] + wv[8 + e + j]);
            if (i < K)
              flux = (float)a[(2 * rr) * K + i];
            else if (i > nx - 1 - K)
              flux = (float)a[(2 * rr + 1) * K + (nx - 1 - i)];
            r[e] = wv[8 + e] - flux;
          }
          reinterpret_cast<float4*>(row)[q] = make_float4(r[0], r[1], r[2], r[3]);
        }
        continue;
      }
      for (int i = lane; i < nx; i += 32) {
        T flux;
        if (fir && i >= K && i <= nx - 1 - K) {
          flux = c[0] * o[i];
#pragma unroll
          for (int j = 1; j <= K; ++j) flux += c[j] * (o[i - j] + o[i + j]);
        } else if (fir) {
          flux = i < K ? a[(2 * rr) * K + i] : a[(2 * rr + 1) * K + (nx - 1 - i)];
        } else {
          flux = a[rr * nx + i];
        }
        row[i] = o[i] - flux;
      }
    }
    __syncwarp();
  }
}

// ---- y / z, orders up to 8: register pipeline marching along the line -----------------------------------------
// One thread per (x, line, segment): it walks its segment (plus an `K`-cell run-in and run-out) once, keeping the last
// three values of every intermediate pass and the K + 1 pending originals in registers; pass m at position p - m is
// formed as soon as pass m - 1 reaches p - m + 1. Loads and stores are coalesced across x; no shared memory, no
// barriers. Out of place (a segment's run-in cells are another segment's outputs).
template <typename T, int K>
__global__ void __launch_bounds__(128)
    filter_march_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t src_ls, int64_t src_os, int64_t dst_ls,
                        int64_t dst_os, int line_len, int n_other, int nx, int seg_len, FirTaps taps) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
  if (x >= nx) return;
  const int p_begin = blockIdx.z * seg_len;
  const int p_end = p_begin + seg_len < line_len ? p_begin + seg_len : line_len;
  const bool line_active = x >= 1 && x <= nx - 2 && o >= 1 && o <= n_other - 2;
  const T* s = src + o * src_os + x;
  T* d = dst + o * dst_os + x;
  if (!line_active) {  // lines on the ring of the other two axes receive no flux
    for (int p = p_begin; p < p_end; ++p) d[p * dst_ls] = s[p * src_ls];
    return;
  }
  if (p_begin >= K && p_end <= line_len - K) {
    // segment away from both line ends (uniform per CTA): sliding window of 2 K + 1 originals, one symmetric filter
    T w[2 * K + 1];
    T c[K + 1];
#pragma unroll
    for (int j = 0; j <= K; ++j) c[j] = T(taps.c[j]);
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) w[j + 1] = s[(p_begin - K + j) * src_ls];
#pragma unroll(2 * K + 1)
    for (int q = p_begin; q < p_end; ++q) {
#pragma unroll
      for (int j = 0; j < 2 * K; ++j) w[j] = w[j + 1];
      w[2 * K] = s[(q + K) * src_ls];
      T flux = c[0] * w[K];
#pragma unroll
      for (int j = 1; j <= K; ++j) flux += c[j] * (w[K - j] + w[K + j]);
      d[q * dst_ls] = w[K] - flux;
    }
    return;
  }
  T g[K][3];       // g[m]: pass m at the three newest positions it has reached
  T pending[K + 1];  // originals of the positions whose result is not out yet
#pragma unroll
  for (int m = 0; m < K; ++m) g[m][0] = g[m][1] = g[m][2] = T(0);
#pragma unroll
  for (int m = 0; m <= K; ++m) pending[m] = T(0);
#pragma unroll 6
  for (int p = p_begin - K; p < p_end + K; ++p) {
    const T v = (p >= 0 && p < line_len) ? s[p * src_ls] : T(0);
    g[0][0] = g[0][1], g[0][1] = g[0][2], g[0][2] = v;
#pragma unroll
    for (int m = 0; m < K; ++m) pending[m] = pending[m + 1];
    pending[K] = v;
    T out = T(0);
#pragma unroll
    for (int m = 1; m <= K; ++m) {
      const int pos = p - m;
      const T val = (pos >= 1 && pos <= line_len - 2) ? filter_point(g[m - 1][0], g[m - 1][1], g[m - 1][2]) : T(0);
      if (m < K)
        g[m][0] = g[m][1], g[m][1] = g[m][2], g[m][2] = val;
      else
        out = val;
    }
    const int q = p - K;
    if (q >= p_begin && q < p_end) d[q * dst_ls] = pending[0] - out;
  }
}

// ---- y / z: 32 contiguous x by a segment of the line, halo of `order` cells on both sides -----------------------
constexpr int FILTER_SEG = 64;

template <typename T>
__global__ void __launch_bounds__(256)
    filter_lines_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t src_ls, int64_t src_os, int64_t dst_ls,
                        int64_t dst_os, int line_len, int n_other, int nx, int order) {
  extern __shared__ __align__(16) unsigned char filter_smem_raw[];
  const int H = FILTER_SEG + 2 * order;
  T* orig = reinterpret_cast<T*>(filter_smem_raw);
  T* a = orig + (size_t)H * 32;
  T* b = a + (size_t)H * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x = blockIdx.x * 32 + tx, o = blockIdx.z, p0 = blockIdx.y * FILTER_SEG - order;
  const bool xin = x < nx;
  for (int q = ty; q < H; q += 8) {
    const int p = p0 + q;
    const T v = (xin && p >= 0 && p < line_len) ? src[o * src_os + p * src_ls + x] : T(0);
    orig[q * 32 + tx] = v;
    a[q * 32 + tx] = v;
  }
  __syncthreads();
  for (int m = 0; m < order; ++m) {
    for (int q = ty; q < H; q += 8) {
      const int p = p0 + q;
      T v = T(0);
      // the tile's two edge rows have no neighbour in the tile: what they miss moves inwards one row per pass and
      // stays inside the halo
      if (q >= 1 && q <= H - 2 && p >= 1 && p <= line_len - 2)
        v = filter_point(a[(q - 1) * 32 + tx], a[q * 32 + tx], a[(q + 1) * 32 + tx]);
      b[q * 32 + tx] = v;
    }
    __syncthreads();
    T* t = a;
    a = b;
    b = t;
  }
  // lines on the ring of the other two axes receive no flux (the flux kernels skip the ring in every axis)
  const bool line_active = x >= 1 && x <= nx - 2 && o >= 1 && o <= n_other - 2;
  for (int q = order + ty; q < order + FILTER_SEG; q += 8) {
    const int p = p0 + q;
    if (xin && p < line_len)
      dst[o * dst_os + p * dst_ls + x] = orig[q * 32 + tx] - (line_active ? a[q * 32 + tx] : T(0));
  }
}

template <typename T>
int filter_rows_x(T* f, int64_t sc, int64_t sz, int64_t sy, int ncomp, int nz, int ny, int nx, int order,
                  cudaStream_t st) {
  int rb = 512 / nx;
  if (rb < 1) rb = 1;
  if (rb > 32) rb = 32;  // one lane per row of the batch holds its address
  const int k = (order <= FIR_MAX_ORDER && nx >= 4 * order) ? order : 0;
  // one-pass form: the originals + the corrected end cells (a third of the pass-by-pass footprint: three times the
  // resident warps, which is what this latency-bound kernel is short of - 43 % occupancy, 8.8 long-scoreboard stalls
  // per issue with the full footprint)
  const size_t per_warp =
      sizeof(T) * (k > 0 ? (((size_t)nx * rb + 2 * (size_t)k * rb + 2 * XPAD + 3) & ~(size_t)3) : 3 * (size_t)nx * rb);
  // quads: float rows whose length, base and strides keep every row 16-byte aligned
  const int vec4 = sizeof(T) == 4 && k > 0 && nx % 4 == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 && sc % 4 == 0 &&
                   sz % 4 == 0 && sy % 4 == 0;
  int warps = (int)((96 * 1024) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) SOPHT_FAIL(SOPHT_ERR_SHAPE, "laplacian filter (fused): rows of %d cells do not fit in shared memory", nx);
  const size_t smem = per_warp * warps;
  const int64_t batches = ((int64_t)(nz - 2) * (ny - 2) * ncomp + rb - 1) / rb;
  int64_t blocks = (batches + warps - 1) / warps;
  if (blocks > 148 * 16) blocks = 148 * 16;
  SOPHT_PROF("laplacian_filter.x", st);
#define ROWS_X(KK)                                                                                              \
  case KK: {                                                                                                    \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      SOPHT_CUDA(cudaFuncSetAttribute(filter_rows_x_kernel<T, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      96 * 1024));                                                              \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    filter_rows_x_kernel<T, KK><<<(int)blocks, 32 * warps, smem, st>>>(f, sc, sz, sy, ncomp, nz, ny, nx, order, rb, \
                                                                      vec4, fir_taps(KK));                            \
    break;                                                                                                      \
  }
  switch (k) {
    ROWS_X(0) ROWS_X(1) ROWS_X(2) ROWS_X(3) ROWS_X(4) ROWS_X(5) ROWS_X(6) ROWS_X(7) ROWS_X(8)
  }
#undef ROWS_X
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

template <typename T, int K>
void launch_march(const T* src, T* dst, int64_t src_ls, int64_t src_os, int64_t dst_ls, int64_t dst_os, int line_len,
                  int n_other, int nx, cudaStream_t st) {
  // segments short enough to give every SM a few thousand threads, long enough to amortise the 2 K run-in / run-out
  int seg = 64;
  while (seg > 4 * K && (int64_t)nx * n_other * ((line_len + seg - 1) / seg) < (int64_t)148 * 2048 * 2) seg /= 2;
  const dim3 grid((nx + 127) / 128, n_other, (line_len + seg - 1) / seg);
  filter_march_kernel<T, K><<<grid, 128, 0, st>>>(src, dst, src_ls, src_os, dst_ls, dst_os, line_len, n_other, nx, seg,
                                                  fir_taps(K));
}

// lines along one strided axis, src -> dst
template <typename T>
int filter_lines(const T* src, T* dst, int64_t src_ls, int64_t src_os, int64_t dst_ls, int64_t dst_os, int line_len,
                 int n_other, int nx, int order, cudaStream_t st) {
#define MARCH(KK)                                                                                   \
  case KK:                                                                                          \
    launch_march<T, KK>(src, dst, src_ls, src_os, dst_ls, dst_os, line_len, n_other, nx, st);       \
    break;
  switch (order) {
    MARCH(1) MARCH(2) MARCH(3) MARCH(4) MARCH(5) MARCH(6) MARCH(7) MARCH(8)
    default: {  // long filters: shared-memory segments
      const size_t smem = sizeof(T) * 3 * 32 * (size_t)(FILTER_SEG + 2 * order);
      static bool attr_set = false;
      if (!attr_set) {
        SOPHT_CUDA(cudaFuncSetAttribute(filter_lines_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        160 * 1024));
        attr_set = true;
      }
      const dim3 grid((nx + 31) / 32, (line_len + FILTER_SEG - 1) / FILTER_SEG, n_other);
      filter_lines_kernel<T><<<grid, dim3(32, 8), smem, st>>>(src, dst, src_ls, src_os, dst_ls, dst_os, line_len,
                                                             n_other, nx, order);
    }
  }
#undef MARCH
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

template <typename T>
int filter_field(T* f, int64_t sc, int64_t sz, int64_t sy, int ncomp, T* scratch, int64_t tz, int64_t ty_, int nz,
                 int ny, int nx, int order, cudaStream_t st) {
  int rc = filter_rows_x<T>(f, sc, sz, sy, ncomp, nz, ny, nx, order, st);  // x, in place, every component
  if (rc) return rc;
  for (int c = 0; c < ncomp; ++c) {
    T* fc = f + c * sc;
    {  // y: lines along y, the other axis is z; f -> scratch
      SOPHT_PROF("laplacian_filter.y", st);
      if ((rc = filter_lines<T>(fc, scratch, sy, sz, ty_, tz, ny, nz, nx, order, st))) return rc;
    }
    {  // z: lines along z, the other axis is y; scratch -> f
      SOPHT_PROF("laplacian_filter.z", st);
      if ((rc = filter_lines<T>(scratch, fc, tz, ty_, sz, sy, nz, ny, nx, order, st))) return rc;
    }
  }
  return SOPHT_OK;
}

}  // namespace
}  // namespace sopht

using namespace sopht;

extern "C" int sopht_laplacian_filter_convolution_3d(int dtype, const sopht_field_t* field,
                                                     const sopht_field_t* scratch, int filter_order, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!valid_field(field, 3, 4) || !valid_field(scratch, 3, 3))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field must be a 3-D scalar or vector field, scratch a 3-D scalar field", __func__);
  if (filter_order < 1 || filter_order > 64)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: filter_order must be in [1, 64] (got %d)", __func__, filter_order);
  const int o = field->ndim - 3;
  const int ncomp = o ? (int)field->shape[0] : 1;
  for (int d = 0; d < 3; ++d)
    if (field->shape[o + d] != scratch->shape[d])
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: scratch and field grids differ", __func__);
  if (field->shape[o] > 65535 || field->shape[o + 1] > 65535 || field->shape[o + 2] > 0x7fffffff)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid too large for this entry point", __func__);
  if (field->stride[o + 2] != 1 || scratch->stride[2] != 1)
    SOPHT_FAIL(SOPHT_ERR_STRIDE, "%s: unit x-stride required", __func__);
  const int nz = (int)field->shape[o], ny = (int)field->shape[o + 1], nx = (int)field->shape[o + 2];
  if (nz < 3 || ny < 3 || nx < 3) return SOPHT_OK;  // no interior: every flux is zero
  cudaStream_t st = as_stream(stream);
  const int64_t sc = o ? field->stride[0] : 0;
  if (dtype == SOPHT_F32)
    return filter_field<float>(reinterpret_cast<float*>(field->data), sc, field->stride[o], field->stride[o + 1],
                               ncomp, reinterpret_cast<float*>(scratch->data), scratch->stride[0], scratch->stride[1],
                               nz, ny, nx, filter_order, st);
  return filter_field<double>(reinterpret_cast<double*>(field->data), sc, field->stride[o], field->stride[o + 1],
                              ncomp, reinterpret_cast<double*>(scratch->data), scratch->stride[0], scratch->stride[1],
                              nz, ny, nx, filter_order, st);
}
