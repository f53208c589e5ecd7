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

// ---- x: whole rows, one warp each -----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    filter_rows_x_kernel(T* f, int64_t sz, int64_t sy, int nz, int ny, int nx, int order) {
  extern __shared__ unsigned char filter_smem_raw[];
  T* smem = reinterpret_cast<T*>(filter_smem_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, warps = blockDim.x >> 5;
  T* orig = smem + (size_t)w * 3 * nx;
  T* a = orig + nx;
  T* b = a + nx;
  const int64_t rows = (int64_t)(nz - 2) * (ny - 2);  // rows on the y / z ring keep their values
  for (int64_t r = (int64_t)blockIdx.x * warps + w; r < rows; r += (int64_t)gridDim.x * warps) {
    const int z = 1 + (int)(r / (ny - 2)), y = 1 + (int)(r % (ny - 2));
    T* row = f + z * sz + y * sy;
    for (int i = lane; i < nx; i += 32) {
      const T v = row[i];
      orig[i] = v;
      a[i] = v;
    }
    __syncwarp();
    for (int m = 0; m < order; ++m) {
      for (int i = lane; i < nx; i += 32)
        b[i] = (i >= 1 && i <= nx - 2) ? filter_point(a[i - 1], a[i], a[i + 1]) : T(0);
      __syncwarp();
      T* t = a;
      a = b;
      b = t;
    }
    for (int i = lane; i < nx; i += 32) row[i] = orig[i] - a[i];
    __syncwarp();
  }
}

// ---- y / z: 32 contiguous x by a segment of the line, halo of `order` cells on both sides -----------------------
constexpr int FILTER_SEG = 64;

template <typename T>
__global__ void __launch_bounds__(256)
    filter_lines_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t src_ls, int64_t src_os, int64_t dst_ls,
                        int64_t dst_os, int line_len, int n_other, int nx, int order) {
  extern __shared__ unsigned char filter_smem_raw[];
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
int filter_component(T* f, int64_t sz, int64_t sy, T* scratch, int64_t tz, int64_t ty_, int nz, int ny, int nx,
                     int order, cudaStream_t st) {
  // x, in place
  {
    const size_t per_warp = sizeof(T) * 3 * (size_t)nx;
    int warps = (int)((96 * 1024) / per_warp);
    if (warps > 8) warps = 8;
    if (warps < 1) SOPHT_FAIL(SOPHT_ERR_SHAPE, "laplacian filter (fused): rows of %d cells do not fit in shared memory", nx);
    const size_t smem = per_warp * warps;
    static bool attr_set = false;
    if (!attr_set) {
      SOPHT_CUDA(cudaFuncSetAttribute(filter_rows_x_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set = true;
    }
    const int64_t rows = (int64_t)(nz - 2) * (ny - 2);
    int64_t blocks = (rows + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    SOPHT_PROF("laplacian_filter.x", st);
    filter_rows_x_kernel<T><<<(int)blocks, 32 * warps, smem, st>>>(f, sz, sy, nz, ny, nx, order);
    SOPHT_CHECK_LAUNCH();
  }
  const size_t smem = sizeof(T) * 3 * 32 * (size_t)(FILTER_SEG + 2 * order);
  static bool attr_set = false;
  if (!attr_set) {
    SOPHT_CUDA(cudaFuncSetAttribute(filter_lines_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  const dim3 block(32, 8);
  {  // y: lines along y, one plane z per blockIdx.z; f -> scratch
    const dim3 grid((nx + 31) / 32, (ny + FILTER_SEG - 1) / FILTER_SEG, nz);
    SOPHT_PROF("laplacian_filter.y", st);
    filter_lines_kernel<T><<<grid, block, smem, st>>>(f, scratch, sy, sz, ty_, tz, ny, nz, nx, order);
    SOPHT_CHECK_LAUNCH();
  }
  {  // z: lines along z, one row y per blockIdx.z; scratch -> f
    const dim3 grid((nx + 31) / 32, (nz + FILTER_SEG - 1) / FILTER_SEG, ny);
    SOPHT_PROF("laplacian_filter.z", st);
    filter_lines_kernel<T><<<grid, block, smem, st>>>(scratch, f, tz, ty_, sz, sy, nz, ny, nx, order);
    SOPHT_CHECK_LAUNCH();
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
  for (int c = 0; c < ncomp; ++c) {
    int rc;
    if (dtype == SOPHT_F32)
      rc = filter_component<float>(reinterpret_cast<float*>(field->data) + (o ? c * field->stride[0] : 0),
                                   field->stride[o], field->stride[o + 1], reinterpret_cast<float*>(scratch->data),
                                   scratch->stride[0], scratch->stride[1], nz, ny, nx, filter_order, st);
    else
      rc = filter_component<double>(reinterpret_cast<double*>(field->data) + (o ? c * field->stride[0] : 0),
                                    field->stride[o], field->stride[o + 1], reinterpret_cast<double*>(scratch->data),
                                    scratch->stride[0], scratch->stride[1], nz, ny, nx, filter_order, st);
    if (rc) return rc;
  }
  return SOPHT_OK;
}
