// CPU emulation of the pruned FFT Poisson pipeline (csrc/poisson_pow2_phases.cuh): every kernel is run
// block by block, phase by phase, thread by thread on the host and the result is compared with a
// double-precision doubled-domain convolution. Pins the index arithmetic of the CUDA kernels without a GPU.
//   nvcc -O2 -std=c++17 -Isopht_b200/csrc -Iinclude tests/host/fft_emul.cu -o build/fft_emul && build/fft_emul
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <complex>
#include <vector>

#include "poisson_pow2_phases.cuh"

using namespace sopht;
using cd = std::complex<double>;

template <class K, int P>
struct PhaseLoop {
  static void run(const typename K::Params& p, int bx, int by, int it, float2* smem, const float2* stage) {
    PhaseLoop<K, P - 1>::run(p, bx, by, it, smem, stage);
    for (int tid = 0; tid < K::THREADS; ++tid) K::template phase<P>(p, bx, by, it, tid, smem, stage);
  }
};
template <class K>
struct PhaseLoop<K, -1> {
  static void run(const typename K::Params&, int, int, int, float2*, const float2*) {}
};

// staged = true emulates the persistent kernel's cp.async prefetch into the staging buffer
template <class K>
void emulate(const typename K::Params& p, int gx, int gy, int niter, bool staged = true) {
  std::vector<float2> smem(K::SMEM_ELEMS + K::EXTRA_ELEMS), stage(K::STAGE_ELEMS);
  for (auto& v : smem) v = make_float2(NAN, NAN);
  for (int tid = 0; tid < K::THREADS; ++tid) K::init(p, tid, smem.data());  // once per (persistent) CTA
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx)
      for (int it = 0; it < niter; ++it) {
        // poison the data tile: catches reads of unwritten slots (the tables persist like on the device)
        for (int q = 0; q < K::SMEM_ELEMS; ++q) smem[q] = make_float2(NAN, NAN);
        for (auto& v : stage) v = make_float2(NAN, NAN);
        if (staged)
          for (int tid = 0; tid < K::THREADS; ++tid) K::prefetch(p, bx, by, it, tid, stage.data());
        PhaseLoop<K, K::NPHASE - 1>::run(p, bx, by, it, smem.data(), staged ? stage.data() : nullptr);
      }
}

static std::vector<float2> twiddles(int L, int denom) {
  std::vector<float2> t(L);
  for (int j = 0; j < L; ++j) {
    const double a = -2.0 * M_PI * j / denom;
    t[j] = make_float2((float)cos(a), (float)sin(a));
  }
  return t;
}

// reference radix-2 FFT (double), in place, sign = -1 forward / +1 inverse (unnormalised)
static void fft_ref(std::vector<cd>& a, int sign) {
  const int n = (int)a.size();
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int len = 2; len <= n; len <<= 1) {
    const double ang = sign * 2 * M_PI / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const cd w(cos(ang * k), sin(ang * k));
        const cd u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

static void fft3_ref(std::vector<cd>& f, int n2z, int n2y, int n2x, int sign) {
  std::vector<cd> line;
  line.resize(n2x);
  for (int k = 0; k < n2z; ++k)
    for (int j = 0; j < n2y; ++j) {
      for (int i = 0; i < n2x; ++i) line[i] = f[((size_t)k * n2y + j) * n2x + i];
      fft_ref(line, sign);
      for (int i = 0; i < n2x; ++i) f[((size_t)k * n2y + j) * n2x + i] = line[i];
    }
  line.resize(n2y);
  for (int k = 0; k < n2z; ++k)
    for (int i = 0; i < n2x; ++i) {
      for (int j = 0; j < n2y; ++j) line[j] = f[((size_t)k * n2y + j) * n2x + i];
      fft_ref(line, sign);
      for (int j = 0; j < n2y; ++j) f[((size_t)k * n2y + j) * n2x + i] = line[j];
    }
  line.resize(n2z);
  for (int j = 0; j < n2y; ++j)
    for (int i = 0; i < n2x; ++i) {
      for (int k = 0; k < n2z; ++k) line[k] = f[((size_t)k * n2y + j) * n2x + i];
      fft_ref(line, sign);
      for (int k = 0; k < n2z; ++k) f[((size_t)k * n2y + j) * n2x + i] = line[k];
    }
}

// P = 1: the whole grid on one GPU. P > 1: the z-slab decomposed solve, every rank emulated in turn and the
// two all-to-all transposes / the Nyquist all-gather done by plain copies (what parallel/slab_poisson.py does
// with torch.distributed on the device).
template <int NZ, int NY, int NX, int P = 1, bool PEER = false>
int run_case() {
  constexpr int C = 3, TX = 8, RX = 4;
  constexpr int LX = NX, LY = 2 * NY, LZ = 2 * NZ;
  constexpr int NZL = NZ / P, NXL = NX / P;
  static_assert(NZ % P == 0 && NX % P == 0 && NXL % TX == 0, "slab split");
  const size_t ncell = (size_t)NZ * NY * NX;
  std::vector<float> rhs(C * ncell), sol(C * ncell, NAN);
  srand(1234 + NX + 7 * NY + 13 * NZ);
  for (auto& v : rhs) v = (float)rand() / RAND_MAX - 0.5f;
  // folded real even spectrum, (NZ+1, NY+1, NX+1); main part [fz][fy][kx<NX], nyquist part [fz][fy]
  std::vector<float> gmain((size_t)(NZ + 1) * (NY + 1) * NX), gnyq((size_t)(NZ + 1) * (NY + 1));
  std::vector<double> gfull((size_t)(NZ + 1) * (NY + 1) * (NX + 1));
  for (int a = 0; a <= NZ; ++a)
    for (int b = 0; b <= NY; ++b)
      for (int c = 0; c <= NX; ++c) {
        const float v = 1.0f / (1.0f + a * 0.37f + b * 0.11f + c * 0.05f) + 0.01f * ((a * 7 + b * 3 + c) % 5);
        gfull[((size_t)a * (NY + 1) + b) * (NX + 1) + c] = v;
        if (c < NX)
          gmain[((size_t)a * (NY + 1) + b) * NX + c] = v;
        else
          gnyq[(size_t)a * (NY + 1) + b] = v;
      }
  auto twx = twiddles(LX, LX), twx2 = twiddles(LX, 2 * LX), twy = twiddles(LY, LY), twz = twiddles(LZ, LZ);
  const float2 poison = make_float2(NAN, NAN);
  // per-rank buffers
  const size_t slab = (size_t)C * NZL * NY * NX;  // = C * P * NZL * NY * NXL = C * NZ * NY * NXL
  std::vector<std::vector<float2>> send(P, std::vector<float2>(slab, poison)),
      recv(P, std::vector<float2>(slab, poison)), nyq_local(P, std::vector<float2>((size_t)C * NZL * NY, poison));
  std::vector<float2> work((size_t)C * NZ * LY * NXL), work2((size_t)C * NZ * LY * NXL), nyq_all((size_t)C * NZ * NY), nyq_work((size_t)C * NZ * LY);

  // 1. x forward on every rank's z-slab
  for (int r = 0; r < P; ++r) {
    p2::SlabDims d{C, NZ, NY, NX, P, r};
    auto xp = p2::slab_x_params(d, rhs.data() + (size_t)r * NZL * NY * NX, nullptr, (int64_t)ncell,
                                (int64_t)NY * NX, NX, send[r].data(), nyq_local[r].data(), twx.data(), twx2.data());
    if (PEER) {  // the kernel writes chunk q straight into rank q's recv buffer ("peer memory")
      float2* peers[8] = {};
      for (int q = 0; q < P; ++q) peers[q] = recv[q].data();
      xp = p2::slab_x_params_peer(xp, d, peers);
    }
    emulate<p2::XFwd<LX, RX>>(xp, (int)((size_t)C * NZL * NY / RX), 1, 1);
  }
  // 2. all-to-all per component: chunk q of rank r's send[c] -> chunk r of rank q's recv[c]; Nyquist all-gather
  const size_t chunk = (size_t)NZL * NY * NXL;
  for (int c = 0; c < C && !PEER; ++c)
    for (int r = 0; r < P; ++r)
      for (int q = 0; q < P; ++q)
        std::copy(send[r].begin() + ((size_t)c * P + q) * chunk, send[r].begin() + ((size_t)c * P + q + 1) * chunk,
                  recv[q].begin() + ((size_t)c * P + r) * chunk);
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < P; ++r)
      std::copy(nyq_local[r].begin() + (size_t)c * NZL * NY, nyq_local[r].begin() + (size_t)(c + 1) * NZL * NY,
                nyq_all.begin() + ((size_t)c * NZ + (size_t)r * NZL) * NY);
  // 3. y forward, z convolution, y inverse on every rank's kx-slab (recv is (C, NZ, NY, NXL))
  if (PEER)
    for (int r = 0; r < P; ++r)
      for (auto& v : send[r]) v = poison;
  for (int r = 0; r < P; ++r) {
    p2::SlabDims d{C, NZ, NY, NX, P, r};
    for (auto& v : work) v = poison;
    for (auto& v : work2) v = poison;
    emulate<p2::YFwd<LY, TX>>(p2::slab_y_params(d, TX, recv[r].data(), work.data(), true, twy.data()), NXL / TX,
                              C * NZ, 1);
    emulate<p2::ZConv<LZ, TX>>(p2::slab_z_params(d, TX, work.data(), work2.data(), gmain.data(), NX, r * NXL, twz.data()), NXL / TX, LY, C);
    // PEER: the kx-slab stays in this rank's send buffer, from where the x inverse of every rank pulls its chunk
    auto yi = p2::slab_y_params(d, TX, work2.data(), PEER ? send[r].data() : recv[r].data(), false, twy.data());
    emulate<p2::YInv<LY, TX>>(yi, NXL / TX, C * NZ, 1);
  }
  {  // Nyquist plane (every rank would do this redundantly)
    p2::SlabDims d{C, NZ, NY, NX, P, 0};
    for (auto& v : nyq_work) v = poison;
    emulate<p2::YFwd<LY, TX>>(p2::nyquist_y_params(d, TX, nyq_all.data(), nyq_work.data(), true, twy.data()),
                              C * NZ / TX, 1, 1);
    emulate<p2::ZConv<LZ, TX>>(p2::nyquist_z_params(d, TX, nyq_work.data(), gnyq.data(), twz.data()), LY / TX, 1, C);
    emulate<p2::YInv<LY, TX>>(p2::nyquist_y_params(d, TX, nyq_work.data(), nyq_all.data(), false, twy.data()),
                              C * NZ / TX, 1, 1);
  }
  // 4. all-to-all back: z range q of rank r's recv[c] -> chunk r of rank q's send[c]; Nyquist slices
  for (int r = 0; r < P && !PEER; ++r)
    for (auto& v : send[r]) v = poison;
  for (int c = 0; c < C && !PEER; ++c)
    for (int r = 0; r < P; ++r)
      for (int q = 0; q < P; ++q)
        std::copy(recv[r].begin() + ((size_t)c * P + q) * chunk, recv[r].begin() + ((size_t)c * P + q + 1) * chunk,
                  send[q].begin() + ((size_t)c * P + r) * chunk);
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < P; ++r)
      std::copy(nyq_all.begin() + ((size_t)c * NZ + (size_t)r * NZL) * NY,
                nyq_all.begin() + ((size_t)c * NZ + (size_t)(r + 1) * NZL) * NY,
                nyq_local[r].begin() + (size_t)c * NZL * NY);
  // 5. x inverse on every rank's z-slab
  for (int r = 0; r < P; ++r) {
    p2::SlabDims d{C, NZ, NY, NX, P, r};
    auto xp = p2::slab_x_params(d, nullptr, sol.data() + (size_t)r * NZL * NY * NX, (int64_t)ncell,
                                (int64_t)NY * NX, NX, send[r].data(), nyq_local[r].data(), twx.data(), twx2.data());
    if (PEER) {  // chunk q of a row is read from rank q's send buffer ("peer memory")
      float2* peers[8] = {};
      for (int q = 0; q < P; ++q) peers[q] = send[q].data();
      xp = p2::slab_x_params_peer(xp, d, peers);
    }
    emulate<p2::XInv<LX, RX>>(xp, (int)((size_t)C * NZL * NY / RX), 1, 1);
  }

  // reference: doubled-domain convolution in double
  double err2 = 0, ref2 = 0;
  const int n2z = 2 * NZ, n2y = 2 * NY, n2x = 2 * NX;
  const double scale = 1.0 / ((double)NX * n2y * n2z);  // pipeline's unnormalised round trip factor
  for (int c = 0; c < C; ++c) {
    std::vector<cd> f((size_t)n2z * n2y * n2x, cd(0, 0));
    for (int k = 0; k < NZ; ++k)
      for (int j = 0; j < NY; ++j)
        for (int i = 0; i < NX; ++i)
          f[((size_t)k * n2y + j) * n2x + i] = rhs[c * ncell + ((size_t)k * NY + j) * NX + i];
    fft3_ref(f, n2z, n2y, n2x, -1);
    for (int k = 0; k < n2z; ++k)
      for (int j = 0; j < n2y; ++j)
        for (int i = 0; i < n2x; ++i) {
          const int fk = k <= NZ ? k : n2z - k, fj = j <= NY ? j : n2y - j, fi = i <= NX ? i : n2x - i;
          f[((size_t)k * n2y + j) * n2x + i] *= gfull[((size_t)fk * (NY + 1) + fj) * (NX + 1) + fi];
        }
    fft3_ref(f, n2z, n2y, n2x, +1);
    for (int k = 0; k < NZ; ++k)
      for (int j = 0; j < NY; ++j)
        for (int i = 0; i < NX; ++i) {
          // reference inverse is unnormalised by 8N; the pipeline by NX*2NY*2NZ = 4N
          const double r = f[((size_t)k * n2y + j) * n2x + i].real() / ((double)n2x * n2y * n2z);
          const double s = sol[c * ncell + ((size_t)k * NY + j) * NX + i] * scale;
          err2 += (r - s) * (r - s);
          ref2 += r * r;
        }
  }
  const double rel = sqrt(err2 / ref2);
  printf("grid (%d,%d,%d) ranks %d%s: LX=%d LY=%d LZ=%d rel L2 err = %.3e %s\n", NZ, NY, NX, P, PEER ? " (peer push / pull)" : "", LX, LY, LZ, rel,
         rel < 2e-6 ? "ok" : "FAIL");
  return rel < 2e-6 ? 0 : 1;
}

int main(int argc, char** argv) {
  const bool full = argc > 1 && argv[1][0] == 'f';
  int bad = 0;
  if (argc > 1 && argv[1][0] == 'k') {  // only the length-1024 column transforms (y and z)
    bad += run_case<512, 8, 16>();
    bad += run_case<8, 512, 16>();
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad;
  }
  bad += run_case<8, 8, 16>();     // L = 16 everywhere
  bad += run_case<16, 32, 64>();   // 32, 64, 64
  bad += run_case<8, 64, 128>();   // 16, 128, 128
  bad += run_case<128, 8, 16>();   // 256 (z)
  bad += run_case<16, 8, 32, 2>();    // z-slab decomposition over 2 ranks
  bad += run_case<8, 16, 64, 4>();    // ... over 4 ranks
  bad += run_case<16, 8, 32, 2, true>();   // transposes fused into the x forward / y inverse kernels (peer push / pull)
  bad += run_case<8, 16, 64, 4, true>();
  if (full) {                      // minutes on one core: every remaining decomposition, incl. three-pass 2048
    bad += run_case<64, 128, 256>();
    bad += run_case<256, 8, 512>();
    bad += run_case<8, 512, 1024>();
    bad += run_case<1024, 8, 2048>();
  }
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
