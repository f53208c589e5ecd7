// Microbenchmark: issue rate of scalar FADD/FFMA vs packed FADD2/FFMA2 (add.f32x2 / fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/f32x2_rate tools/micro/f32x2_rate.cu && build/f32x2_rate
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(float2* out, float2 seed, int iters) {
  float2 a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
  const float2 c = make_float2(seed.y * 0.5f + 1.0f, seed.x * 0.25f + 1.0f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) {  // scalar FADD x2
        a[i].x += c.x;
        a[i].y += c.y;
      } else if (MODE == 1) {  // packed FADD2
        a[i] = __fadd2_rn(a[i], c);
      } else if (MODE == 2) {  // scalar FFMA x2
        a[i].x = fmaf(a[i].x, c.x, c.y);
        a[i].y = fmaf(a[i].y, c.y, c.x);
      } else {  // packed FFMA2
        a[i] = __ffma2_rn(a[i], c, c);
      }
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) s.x += a[i].x, s.y += a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(const char* name, int ctas_per_sm) {
  float2* out;
  cudaMalloc(&out, sizeof(float2) * 148 * 8 * 256);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE, ILP><<<148 * ctas_per_sm, 256>>>(out, make_float2(1.f, 2.f), 100);
  cudaEventRecord(e0);
  k<MODE, ILP><<<148 * ctas_per_sm, 256>>>(out, make_float2(1.f, 2.f), iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop_elems = (double)148 * ctas_per_sm * 256 * iters * ILP * 2;  // float results
  printf("%-14s ILP=%2d ctas/SM=%d: %.3f ms, %.1f G float-results/s per SM-clock-free, %.2f results/clk/SM @1.965GHz\n",
         name, ILP, ctas_per_sm, ms, flop_elems / ms / 1e6, flop_elems / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int c : {1, 2, 4}) {
    run<0, 8>("FADD scalar", c);
    run<1, 8>("FADD2 packed", c);
    run<2, 8>("FFMA scalar", c);
    run<3, 8>("FFMA2 packed", c);
  }
  return 0;
}
