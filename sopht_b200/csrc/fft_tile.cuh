// CTA-level power-of-two complex FFT building blocks for the pruned Poisson pipeline (poisson_pow2.cu).
//
// A sequence of length L = R1*R2*R3 lives in shared memory (one "column" per sequence); T = L/E threads
// per column each own E elements per pass. Forward transforms are decimation-in-frequency, inverse
// transforms decimation-in-time with the mirrored pass order, so the spectrum a forward last pass leaves in
// REGISTERS is exactly the input of the inverse first pass (used by the fused z-pass: forward -> x G_hat ->
// inverse with no reordering and only two shared-memory exchanges for L <= 1024).
//
// Position/ordering conventions (all passes are in place):
//   pass p works on blocks of length Sprev_p (Sprev_0 = L, Sprev_1 = L/R1, Sprev_2 = L/(R1 R2)); inside a
//   block, butterfly j (0 <= j < S = Sprev/R) touches positions blk*Sprev + n*S + j, n = 0..R-1.
//   After the forward last pass, position blk*Rlast + klast holds X[rev(blk) + (L/Rlast)*klast], where rev()
//   reverses the (k1[,k2]) digits of blk:  two passes: rev = blk;  three passes: blk = k1*R2 + k2 -> k1 + R1*k2.
//
// Everything is __host__ __device__ so that the exact index arithmetic is unit-tested on the CPU
// (tests/host/fft_emul.cu emulates the thread/phase structure) before it runs on the GPU.
#pragma once
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define FFT_HD __host__ __device__ __forceinline__
#else
#define FFT_HD inline
#endif

namespace sopht {
namespace fft {

// Two implementations of the complex arithmetic and the in-register DFTs below. A translation unit that defines
// SOPHT_FFT_USE_PACKED before including this file gets the packed-FP32 form on sm_100a (used where the kernel is bound by
// issue slots); everything else keeps the scalar form (measured: the HBM-bound x / y kernels are 2-19 % slower with the
// packed form, profiles/r02_zpass_experiments.txt).
#if defined(SOPHT_FFT_USE_PACKED) && defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define SOPHT_FFT_PACKED 1
#endif

#if defined(SOPHT_FFT_USE_PACKED)
// Complex arithmetic. On the device the (re, im) pair of a float2 is one 64-bit register pair and sm_100a has packed
// FP32 instructions on such pairs (FADD2 / FMUL2 / FFMA2: add.f32x2, mul.f32x2, fma.rn.f32x2) whose operands take a
// per-operand half swap (.LO_HI), a per-half sign (.NP) and a scalar broadcast (.F32) for free: a complex add is ONE
// instruction instead of two, a complex product TWO (FMUL2 + FFMA2) instead of four, multiplication by -i / +i folds
// into the operand modifiers of the next add. The FP32 pipe retires the same number of results per clock either way
// (tools/micro/f32x2_rate.cu: 120 results / clk / SM scalar and packed) - what the packed forms halve is the ISSUE
// slots, which is what bounds the z pass (profiles/r02_zpass_experiments.txt). ptxas recognises the forms below from
// the intrinsics with make_float2 operand shuffles; the host build (tests/host) keeps plain scalar arithmetic.

FFT_HD float2 cadd(float2 a, float2 b) {
#ifdef SOPHT_FFT_PACKED
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
FFT_HD float2 csub(float2 a, float2 b) {
#ifdef SOPHT_FFT_PACKED
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
FFT_HD float2 cmul(float2 a, float2 b) {
#ifdef SOPHT_FFT_PACKED
  const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x));  // (ay by, ay bx)
  return __ffma2_rn(make_float2(a.x, a.x), b, make_float2(-t.x, t.y));
#else
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
FFT_HD float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
#ifdef SOPHT_FFT_PACKED
  const float2 t = __fmul2_rn(make_float2(a.x, a.x), b);  // (ax bx, ax by)
  return __ffma2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x), make_float2(t.x, -t.y));
#else
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
#endif
}
FFT_HD float2 cscale(float2 a, float s) {  // a * s, s real
#ifdef SOPHT_FFT_PACKED
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}
// 8-byte store of a packed result: handing the store ONE 64-bit register (the pair the packed instruction wrote) keeps
// ptxas from copying the pair into a common scratch pair first (MOV, MOV, STS chains that serialise on that pair)
FFT_HD void store2(float2* p, float2 v) {
#ifdef SOPHT_FFT_PACKED
  unsigned long long u;
  asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(v.x), "f"(v.y));
  *reinterpret_cast<unsigned long long*>(p) = u;
#else
  *p = v;
#endif
}
FFT_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
FFT_HD float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)
// a * (-i) in a forward transform, a * (+i) in an inverse one
template <bool INV>
FFT_HD float2 mul_qi(float2 a) {
  return INV ? mul_pi(a) : mul_mi(a);
}

// w32(j) = exp(-2 pi i j / 32); j is a compile-time constant after unrolling.
FFT_HD float2 w32(int j) {
  constexpr float c[32] = {
      1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
      0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.f, -0.19509032201612819f,
      -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
      -0.92387953251128674f, -0.98078528040323043f, -1.f, -0.98078528040323043f, -0.92387953251128685f,
      -0.83146961230254546f, -0.70710678118654768f, -0.55557023301960218f, -0.38268343236509034f,
      -0.19509032201612866f, 0.f, 0.1950903220161283f, 0.38268343236509f, 0.55557023301960184f,
      0.70710678118654735f, 0.83146961230254524f, 0.92387953251128652f, 0.98078528040323032f};
  constexpr float s[32] = {
      0.f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f, -0.70710678118654746f,
      -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f, -1.f, -0.98078528040323043f,
      -0.92387953251128674f, -0.83146961230254546f, -0.70710678118654757f, -0.55557023301960218f,
      -0.38268343236508989f, -0.19509032201612861f, 0.f, 0.19509032201612836f, 0.38268343236508967f,
      0.55557023301960196f, 0.70710678118654746f, 0.83146961230254524f, 0.92387953251128652f,
      0.98078528040323032f, 1.f, 0.98078528040323043f, 0.92387953251128663f, 0.83146961230254546f,
      0.70710678118654768f, 0.55557023301960218f, 0.38268343236509039f, 0.19509032201612872f};
  return make_float2(c[j & 31], s[j & 31]);
}
#ifdef __CUDACC__
// The same 32 values as (re, im) pairs in constant memory: a packed product takes the pair as ONE 64-bit uniform-register
// operand (LDCU.64, hoisted out of the tile loops) where immediates would cost two MOVs per pair and use.
static __constant__ float2 kW32[32] = {
    {1.f, 0.f}, {0.98078528040323043f, -0.19509032201612825f}, {0.92387953251128674f, -0.38268343236508978f},
    {0.83146961230254524f, -0.55557023301960218f}, {0.70710678118654757f, -0.70710678118654746f},
    {0.55557023301960229f, -0.83146961230254524f}, {0.38268343236508984f, -0.92387953251128674f},
    {0.19509032201612833f, -0.98078528040323043f}, {0.f, -1.f}, {-0.19509032201612819f, -0.98078528040323043f},
    {-0.38268343236508973f, -0.92387953251128674f}, {-0.55557023301960196f, -0.83146961230254546f},
    {-0.70710678118654746f, -0.70710678118654757f}, {-0.83146961230254535f, -0.55557023301960218f},
    {-0.92387953251128674f, -0.38268343236508989f}, {-0.98078528040323043f, -0.19509032201612861f}, {-1.f, 0.f},
    {-0.98078528040323043f, 0.19509032201612836f}, {-0.92387953251128685f, 0.38268343236508967f},
    {-0.83146961230254546f, 0.55557023301960196f}, {-0.70710678118654768f, 0.70710678118654746f},
    {-0.55557023301960218f, 0.83146961230254524f}, {-0.38268343236509034f, 0.92387953251128652f},
    {-0.19509032201612866f, 0.98078528040323032f}, {0.f, 1.f}, {0.1950903220161283f, 0.98078528040323043f},
    {0.38268343236509f, 0.92387953251128663f}, {0.55557023301960184f, 0.83146961230254546f},
    {0.70710678118654735f, 0.70710678118654768f}, {0.83146961230254524f, 0.55557023301960218f},
    {0.92387953251128652f, 0.38268343236509039f}, {0.98078528040323032f, 0.19509032201612872f}};
#endif
FFT_HD float2 w32c(int j) {
#ifdef SOPHT_FFT_PACKED
  float2 w = kW32[j & 31];
  return w;
#else
  return w32(j);
#endif
}

// a * exp(-+ 2 pi i j / 32) (forward: minus; INV: plus). Only w^1, w^2, w^3 are ever multiplied with (three constant
// pairs that stay in uniform registers): w^4 = h (1 - i) is one add and one real scaling, w^(8 - r) = -i conj(w^r), and
// the quadrant factor (-i)^q is a half swap / sign that the consuming add takes as an operand modifier.
template <bool INV = false>
FFT_HD float2 rot32(float2 a, int j) {
  j &= 31;
  if (INV) j = (32 - j) & 31;  // conj(w^j) = w^(32 - j)
  constexpr float h = 0.70710678118654757f;
  const int q = j >> 3, r = j & 7;
  float2 b = a;
  if (r == 4)
    b = cscale(cadd(a, mul_mi(a)), h);
  else if (r >= 1 && r <= 3)
    b = cmul(a, w32c(r));
  else if (r >= 5)
    b = mul_mi(cmul_conj(a, w32c(8 - r)));
  if (q == 1) return mul_mi(b);
  if (q == 2) return make_float2(-b.x, -b.y);
  if (q == 3) return mul_pi(b);
  return b;
}

// ---- in-register DFTs, natural order in and out: v[k] <- sum_n v[n] exp(-+ 2 pi i n k / R) (INV: plus, unnormalised) --
template <int R, bool INV = false>
struct Dft;

// run_half: the same transform when the upper half of the input (v[R/2..R-1]) is known to be zero - the first
// radix-2 stage degenerates into copies. Written out because x + 0.0f is not foldable under IEEE rules
// (-0.0f + 0.0f = +0.0f), so the compiler keeps those additions when it is merely handed literal zeros.
template <bool INV>
struct Dft<1, INV> {
  FFT_HD static void run(float2*) {}
  FFT_HD static void run_half(float2*) {}
};
template <bool INV>
struct Dft<2, INV> {
  FFT_HD static void run(float2* v) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
  FFT_HD static void run_half(float2* v) { v[1] = v[0]; }
};
template <bool INV>
struct Dft<4, INV> {
  FFT_HD static void run(float2* v) {
    const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    const float2 c = cadd(v[1], v[3]), d = mul_qi<INV>(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
  }
  FFT_HD static void run_half(float2* v) {
    const float2 a = v[0], c = v[1], d = mul_qi<INV>(v[1]);
    v[0] = cadd(a, c);
    v[1] = cadd(a, d);
    v[2] = csub(a, c);
    v[3] = csub(a, d);
  }
};

// R = RA*RB: n = RB*a + b, k = ka + RA*kb. HALF: inputs n >= R/2 are zero, i.e. a >= RA/2 in every sub-transform.
template <int RA, int RB, bool INV, bool HALF = false>
FFT_HD void dft_composite(float2* v) {
  constexpr int R = RA * RB;
  float2 u[RB][RA];
#pragma unroll
  for (int b = 0; b < RB; ++b) {
#pragma unroll
    for (int a = 0; a < RA; ++a) u[b][a] = v[RB * a + b];
    if (HALF)
      Dft<RA, INV>::run_half(u[b]);
    else
      Dft<RA, INV>::run(u[b]);
#pragma unroll
    for (int ka = 1; ka < RA; ++ka)
      if (b > 0) u[b][ka] = rot32<INV>(u[b][ka], (32 / R) * b * ka);
  }
#pragma unroll
  for (int ka = 0; ka < RA; ++ka) {
    float2 w[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) w[b] = u[b][ka];
    Dft<RB, INV>::run(w);
#pragma unroll
    for (int kb = 0; kb < RB; ++kb) v[ka + RA * kb] = w[kb];
  }
}
template <bool INV>
struct Dft<8, INV> {
  FFT_HD static void run(float2* v) { dft_composite<4, 2, INV>(v); }
  FFT_HD static void run_half(float2* v) { dft_composite<4, 2, INV, true>(v); }
};
template <bool INV>
struct Dft<16, INV> {
  FFT_HD static void run(float2* v) { dft_composite<4, 4, INV>(v); }
  FFT_HD static void run_half(float2* v) { dft_composite<4, 4, INV, true>(v); }
};
template <bool INV>
struct Dft<32, INV> {
  FFT_HD static void run(float2* v) { dft_composite<8, 4, INV>(v); }
  FFT_HD static void run_half(float2* v) { dft_composite<8, 4, INV, true>(v); }
};

// forward (INV = false) or unnormalised inverse (conjugated twiddles) DFT of v[0 .. R-1]
template <int R, bool INV>
FFT_HD void dft(float2* v) {
  Dft<R, INV>::run(v);
}

#else  // scalar form

FFT_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }  // a * s, s real
FFT_HD void store2(float2* p, float2 v) { *p = v; }
FFT_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFT_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FFT_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
FFT_HD float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
FFT_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// w32(j) = exp(-2 pi i j / 32); j is a compile-time constant after unrolling.
FFT_HD float2 w32(int j) {
  constexpr float c[32] = {
      1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
      0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.f, -0.19509032201612819f,
      -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
      -0.92387953251128674f, -0.98078528040323043f, -1.f, -0.98078528040323043f, -0.92387953251128685f,
      -0.83146961230254546f, -0.70710678118654768f, -0.55557023301960218f, -0.38268343236509034f,
      -0.19509032201612866f, 0.f, 0.1950903220161283f, 0.38268343236509f, 0.55557023301960184f,
      0.70710678118654735f, 0.83146961230254524f, 0.92387953251128652f, 0.98078528040323032f};
  constexpr float s[32] = {
      0.f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f, -0.70710678118654746f,
      -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f, -1.f, -0.98078528040323043f,
      -0.92387953251128674f, -0.83146961230254546f, -0.70710678118654757f, -0.55557023301960218f,
      -0.38268343236508989f, -0.19509032201612861f, 0.f, 0.19509032201612836f, 0.38268343236508967f,
      0.55557023301960196f, 0.70710678118654746f, 0.83146961230254524f, 0.92387953251128652f,
      0.98078528040323032f, 1.f, 0.98078528040323043f, 0.92387953251128663f, 0.83146961230254546f,
      0.70710678118654768f, 0.55557023301960218f, 0.38268343236509039f, 0.19509032201612872f};
  return make_float2(c[j & 31], s[j & 31]);
}

// a * exp(-2 pi i j / 32) with the trivial rotations done without multiplies
FFT_HD float2 rot32(float2 a, int j) {
  j &= 31;
  if (j == 0) return a;
  if (j == 8) return make_float2(a.y, -a.x);
  if (j == 16) return make_float2(-a.x, -a.y);
  if (j == 24) return make_float2(-a.y, a.x);
  return cmul(a, w32(j));
}

// ---- in-register DFTs, natural order in and out: v[k] <- sum_n v[n] exp(-2 pi i n k / R) ----------------
template <int R>
struct Dft;

// run_half: the same transform when the upper half of the input (v[R/2..R-1]) is known to be zero - the first
// radix-2 stage degenerates into copies. Written out because x + 0.0f is not foldable under IEEE rules
// (-0.0f + 0.0f = +0.0f), so the compiler keeps those additions when it is merely handed literal zeros.
template <>
struct Dft<1> {
  FFT_HD static void run(float2*) {}
  FFT_HD static void run_half(float2*) {}
};
template <>
struct Dft<2> {
  FFT_HD static void run(float2* v) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
  FFT_HD static void run_half(float2* v) { v[1] = v[0]; }
};
template <>
struct Dft<4> {
  FFT_HD static void run(float2* v) {
    const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    const float2 c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
  }
  FFT_HD static void run_half(float2* v) {
    const float2 a = v[0], c = v[1], d = mul_mi(v[1]);
    v[0] = cadd(a, c);
    v[1] = cadd(a, d);
    v[2] = csub(a, c);
    v[3] = csub(a, d);
  }
};

// R = RA*RB: n = RB*a + b, k = ka + RA*kb. HALF: inputs n >= R/2 are zero, i.e. a >= RA/2 in every sub-transform.
template <int RA, int RB, bool HALF = false>
FFT_HD void dft_composite(float2* v) {
  constexpr int R = RA * RB;
  float2 u[RB][RA];
#pragma unroll
  for (int b = 0; b < RB; ++b) {
#pragma unroll
    for (int a = 0; a < RA; ++a) u[b][a] = v[RB * a + b];
    if (HALF)
      Dft<RA>::run_half(u[b]);
    else
      Dft<RA>::run(u[b]);
#pragma unroll
    for (int ka = 1; ka < RA; ++ka)
      if (b > 0) u[b][ka] = rot32(u[b][ka], (32 / R) * b * ka);
  }
#pragma unroll
  for (int ka = 0; ka < RA; ++ka) {
    float2 w[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) w[b] = u[b][ka];
    Dft<RB>::run(w);
#pragma unroll
    for (int kb = 0; kb < RB; ++kb) v[ka + RA * kb] = w[kb];
  }
}
template <>
struct Dft<8> {
  FFT_HD static void run(float2* v) { dft_composite<4, 2>(v); }
  FFT_HD static void run_half(float2* v) { dft_composite<4, 2, true>(v); }
};
template <>
struct Dft<16> {
  FFT_HD static void run(float2* v) { dft_composite<4, 4>(v); }
  FFT_HD static void run_half(float2* v) { dft_composite<4, 4, true>(v); }
};
template <>
struct Dft<32> {
  FFT_HD static void run(float2* v) { dft_composite<8, 4>(v); }
  FFT_HD static void run_half(float2* v) { dft_composite<8, 4, true>(v); }
};

template <int R>
FFT_HD void swap_xy(float2* v) {
#pragma unroll
  for (int i = 0; i < R; ++i) v[i] = make_float2(v[i].y, v[i].x);
}
// inverse DFT (unnormalised) through the swap identity: IDFT(x) = swap(DFT(swap(x)))
template <int R, bool INV>
FFT_HD void dft(float2* v) {
  if (INV) swap_xy<R>(v);
  Dft<R>::run(v);
  if (INV) swap_xy<R>(v);
}

#endif  // SOPHT_FFT_USE_PACKED

// ---- decomposition table --------------------------------------------------------------------------------
template <int L>
struct Cfg;
#define SOPHT_FFT_CFG(L_, R1_, R2_, R3_, E_)                      \
  template <>                                                     \
  struct Cfg<L_> {                                                \
    static constexpr int R1 = R1_, R2 = R2_, R3 = R3_, E = E_;    \
    static constexpr int T = L_ / E_;                             \
    static constexpr int NP = R3_ > 1 ? 3 : 2;                    \
    static constexpr int RLAST = R3_ > 1 ? R3_ : R2_;             \
  };
SOPHT_FFT_CFG(16, 4, 4, 1, 4)
SOPHT_FFT_CFG(32, 8, 4, 1, 8)
SOPHT_FFT_CFG(64, 8, 8, 1, 8)
SOPHT_FFT_CFG(128, 16, 8, 1, 16)
SOPHT_FFT_CFG(256, 16, 16, 1, 16)
SOPHT_FFT_CFG(512, 32, 16, 1, 32)
SOPHT_FFT_CFG(1024, 32, 32, 1, 32)  // (4, 16, 16; E = 16: 512 threads, 5 phases) measured 25-30% slower
SOPHT_FFT_CFG(2048, 16, 16, 8, 16)
#undef SOPHT_FFT_CFG

// spectrum index held at position blk*RLAST + klast after the forward last pass
template <int L>
FFT_HD int spectrum_index(int blk, int klast) {
  using C = Cfg<L>;
  if (C::NP == 2) return blk + C::R1 * klast;
  const int k1 = blk / C::R2, k2 = blk % C::R2;
  return k1 + C::R1 * k2 + C::R1 * C::R2 * klast;
}
// position of spectrum index k after the forward last pass (inverse of the above)
template <int L>
FFT_HD int spectrum_position(int k) {
  using C = Cfg<L>;
  if (C::NP == 2) return (k % C::R1) * C::R2 + k / C::R1;
  const int k1 = k % C::R1, k2 = (k / C::R1) % C::R2, k3 = k / (C::R1 * C::R2);
  return (k1 * C::R2 + k2) * C::R3 + k3;
}

// ---- twiddle table ---------------------------------------------------------------------------------------
// The passes multiply butterfly output k of butterfly j by w^(j k) (first / last pass) or w^(j k L/SP) (middle
// pass of the three-pass lengths), w = exp(-2 pi i / L). Read from the natural table w^n the lanes of a warp
// (consecutive j, one k) would hit addresses k * 8 bytes apart - up to 16-way bank conflicts for even k, 2.6x
// the shared-memory wavefronts on average. The table is therefore stored per pass, butterfly index fastest:
//   table[k * S1 + j]     = w^(j k),         k < R1, j < S1 = L / R1            (L entries)
//   table[L + k * S2 + j] = w^(j k L / SP),  k < R2, j < S2 = L / (R1 R2)       (L / R1 entries, NP = 3 only)
template <int L>
struct TwTable {
  using C = Cfg<L>;
  static constexpr int S1 = L / C::R1;
  static constexpr int S2 = L / (C::R1 * C::R2);
  static constexpr int SIZE = L + (C::NP == 3 ? L / C::R1 : 0);
  // natural: w^n, n < L (global memory); one call per thread of a THREADS-wide CTA
  template <int THREADS>
  FFT_HD static void fill(float2* table, const float2* natural, int tid) {
    for (int i = tid; i < L; i += THREADS) table[i] = natural[(i % S1) * (i / S1)];
    if (C::NP == 3)
      for (int i = tid; i < L / C::R1; i += THREADS) table[L + i] = natural[(i % S2) * (i / S2) * C::R1];
  }
};

// ---- passes. `Acc` maps a logical position (0..L-1) of THIS thread's column to a float2& in smem ---------
// tw: the TwTable<L> above

// forward first pass: global -> butterfly R1 -> twiddle -> smem. FULL = false: the input is zero-padded to twice its
// length (positions >= L/2 are zero and never read - the unbounded solver's doubled domain); FULL = true: all L
// positions are data (the periodic solver).
template <int L, bool FULL, class Load, class Acc>
FFT_HD void fwd_first_x(Load ld, Acc sm, int t, const float2* __restrict__ tw) {
  using C = Cfg<L>;
  constexpr int R = C::R1, S = L / R, NB = C::E / R;
  float2 v[NB][R];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = t + q * C::T;
#pragma unroll
    for (int n = 0; n < R; ++n) v[q][n] = (FULL || n < R / 2) ? ld(n * S + j) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = t + q * C::T;
    if (FULL)
      Dft<R>::run(v[q]);
    else
      Dft<R>::run_half(v[q]);  // pruned: the upper half of the padded input is zero
    sm.at(0, j) = v[q][0];
#pragma unroll
    for (int k = 1; k < R; ++k) sm.at(k * S, j) = cmul(v[q][k], tw[k * S + j]);
  }
}
template <int L, class Load, class Acc>
FFT_HD void fwd_first(Load ld, Acc sm, int t, const float2* __restrict__ tw) {
  fwd_first_x<L, false>(ld, sm, t, tw);
}

// forward middle pass (three-pass configurations only): smem -> butterfly R2 -> twiddle -> smem
template <int L, class Acc>
FFT_HD void fwd_mid(Acc sm, int t, const float2* __restrict__ tw) {
  using C = Cfg<L>;
  constexpr int R = C::R2, SP = L / C::R1, S = SP / R, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int b = t + q * C::T, blk = b / S, j = b % S, base = blk * SP + j;
    float2 v[R];
#pragma unroll
    for (int n = 0; n < R; ++n) v[n] = sm.at(n * S, base);
    dft<R, false>(v);
    sm.at(0, base) = v[0];
#pragma unroll
    for (int k = 1; k < R; ++k) sm.at(k * S, base) = cmul(v[k], tw[L + k * S + j]);
  }
}

// forward last pass: smem -> butterfly RLAST -> sink(spectrum index k, position, value)
template <int L, class Acc, class Sink>
FFT_HD void fwd_last(Acc sm, int t, Sink sink) {
  using C = Cfg<L>;
  constexpr int R = C::RLAST, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int blk = t + q * C::T;
    float2 v[R];
#pragma unroll
    for (int n = 0; n < R; ++n) v[n] = sm.at(n, blk * R);
    dft<R, false>(v);
#pragma unroll
    for (int k = 0; k < R; ++k) sink(spectrum_index<L>(blk, k), blk * R + k, v[k]);
  }
}

// inverse first pass: src(spectrum index k, position) -> inverse butterfly RLAST -> smem
template <int L, class Src, class Acc>
FFT_HD void inv_first(Src src, Acc sm, int t) {
  using C = Cfg<L>;
  constexpr int R = C::RLAST, NB = C::E / R;
  float2 v[NB][R];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int blk = t + q * C::T;
#pragma unroll
    for (int k = 0; k < R; ++k) v[q][k] = src(spectrum_index<L>(blk, k), blk * R + k);
  }
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int blk = t + q * C::T;
    dft<R, true>(v[q]);
#pragma unroll
    for (int n = 0; n < R; ++n) sm.at(n, blk * R) = v[q][n];
  }
}

// fused: forward last pass -> multiply by a real spectrum g(k) -> inverse first pass, all in registers
template <int L, class Acc, class G>
FFT_HD void fwd_last_mul_inv_first(Acc sm, int t, G g) {
  using C = Cfg<L>;
  constexpr int R = C::RLAST, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int blk = t + q * C::T;
    float2 v[R];
#pragma unroll
    for (int n = 0; n < R; ++n) v[n] = sm.at(n, blk * R);
    dft<R, false>(v);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float s = g(blk, k);
      v[k] = cscale(v[k], s);
    }
    dft<R, true>(v);
#pragma unroll
    for (int n = 0; n < R; ++n) sm.at(n, blk * R) = v[n];
  }
}

// inverse middle pass (three-pass configurations): smem -> conj twiddle -> inverse butterfly R2 -> smem
template <int L, class Acc>
FFT_HD void inv_mid(Acc sm, int t, const float2* __restrict__ tw) {
  using C = Cfg<L>;
  constexpr int R = C::R2, SP = L / C::R1, S = SP / R, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int b = t + q * C::T, blk = b / S, j = b % S, base = blk * SP + j;
    float2 v[R];
    v[0] = sm.at(0, base);
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cmul_conj(sm.at(k * S, base), tw[L + k * S + j]);
    dft<R, true>(v);
#pragma unroll
    for (int n = 0; n < R; ++n) sm.at(n * S, base) = v[n];
  }
}

// inverse last pass: smem -> conj twiddle -> inverse butterfly R1 -> st(position e, value) for e < L/2 only
// (FULL = true: every position)
template <int L, bool FULL, class Acc, class Store>
FFT_HD void inv_last_x(Acc sm, int t, const float2* __restrict__ tw, Store st) {
  using C = Cfg<L>;
  constexpr int R = C::R1, S = L / R, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = t + q * C::T;
    float2 v[R];
    v[0] = sm.at(0, j);
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cmul_conj(sm.at(k * S, j), tw[k * S + j]);
    dft<R, true>(v);
#pragma unroll
    for (int n = 0; n < (FULL ? R : R / 2); ++n) st(n * S + j, v[n]);
  }
}
template <int L, class Acc, class Store>
FFT_HD void inv_last(Acc sm, int t, const float2* __restrict__ tw, Store st) {
  inv_last_x<L, false>(sm, t, tw, st);
}


// ---- which elements a thread consumes in its first pass (used to prefetch exactly those) -----------------
template <int L, bool FULL, class F>
FFT_HD void fwd_first_elems_x(int t, F f) {  // f(position e), e < L/2 (FULL: e < L)
  using C = Cfg<L>;
  constexpr int R = C::R1, S = L / R, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = t + q * C::T;
#pragma unroll
    for (int n = 0; n < (FULL ? R : R / 2); ++n) f(n * S + j);
  }
}
template <int L, class F>
FFT_HD void fwd_first_elems(int t, F f) {
  fwd_first_elems_x<L, false>(t, f);
}
template <int L, class F>
FFT_HD void inv_first_elems(int t, F f) {  // f(spectrum index k)
  using C = Cfg<L>;
  constexpr int R = C::RLAST, NB = C::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int blk = t + q * C::T;
#pragma unroll
    for (int k = 0; k < R; ++k) f(spectrum_index<L>(blk, k));
  }
}

// 8-byte asynchronous global -> shared copy (cp.async / LDGSTS); a plain copy in the host emulation
FFT_HD void async_copy8(float2* smem_dst, const float2* gsrc) {
#ifdef __CUDA_ARCH__
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
#else
  *smem_dst = *gsrc;
#endif
}
FFT_HD void async_copy4(float* smem_dst, const float* gsrc) {
#ifdef __CUDA_ARCH__
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
#else
  *smem_dst = *gsrc;
#endif
}
// bring the L2 sector holding *p in from DRAM, no register or shared-memory destination
FFT_HD void prefetch_l2(const void* p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
#else
  (void)p;
#endif
}
FFT_HD void async_commit_wait_all() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
#endif
}

}  // namespace fft
}  // namespace sopht
