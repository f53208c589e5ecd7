// Launcher of the warp-quartet z pass (poisson_zquad.cuh). A translation unit of its own because it builds the FFT
// building blocks of fft_tile.cuh in their packed-FP32 form (FADD2 / FMUL2 / FFMA2), which the HBM-bound kernels of
// poisson_pow2.cu do not want; -DSOPHT_ZQUAD_SCALAR builds it with the scalar form for comparison.
#ifndef SOPHT_ZQUAD_SCALAR
#define SOPHT_FFT_USE_PACKED 1
#endif
#include "common.cuh"
#include "poisson_zquad.cuh"

namespace sopht {

int launch_zquad(const p2::ZRowParams& p, int nunits, cudaStream_t st) {
  using K = p2::ZQuad<1024>;
  static int num_sm = 0;
  if (!num_sm) {
    SOPHT_CUDA(cudaFuncSetAttribute(p2::zquad_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)K::SMEM_BYTES));
    int dev = 0;
    SOPHT_CUDA(cudaGetDevice(&dev));
    SOPHT_CUDA(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const int g = nunits / 2 < num_sm ? nunits / 2 : num_sm;  // a CTA owns pairs of units
  SOPHT_PROF("poisson.z_conv", st);
  SOPHT_CUDA(launch_pdl(p2::zquad_kernel<1024>, dim3(g), dim3(K::THREADS), K::SMEM_BYTES, st, p, nunits));
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

}  // namespace sopht
