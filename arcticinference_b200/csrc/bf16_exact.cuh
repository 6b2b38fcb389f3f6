// bf16_exact.cuh - bit-exact emulation of PyTorch-eager bf16 elementwise arithmetic at a fraction
// of the instruction count (used by the fused LSTM cell kernel, lstm_cell.cu).
//
// PyTorch eager computes every bf16 elementwise op in fp32 and rounds its output to bf16
// (SURVEY.md 7-H1).  Three facts make that cheap to reproduce exactly:
//
//  1. For bf16 operands, fp32 add / mul followed by a bf16 rounding equals the correctly rounded
//     bf16 result (double rounding is innocuous when the wide format has >= 2p+2 = 18 significand
//     bits; fp32 has 24).  So R(a*b), R(a+b) on bf16 values ARE `mul.rn.bf16x2` / `add.rn.bf16x2`:
//     two elements per instruction and no conversion.  (`.rn` forbids contraction into an fma.)
//  2. sigmoid and GELU take a bf16 INPUT and produce a bf16 OUTPUT: functions of 65 536 values.
//     A fast approximation f~ with |f~ - f| <= K fp32-ulps rounds to the same bf16 as ATen's f
//     unless f~ lies within K ulps of a bf16 rounding tie; that case (probability 2K/65536) falls
//     back to ATen's exact formula.  tests/test_cell_gpu.py checks ALL 65 536 inputs of both
//     functions against torch on the device, so exactness is proven, not estimated.
//  3. Where GELU cancels (x <= -2: 1 + erf(x/sqrt2) loses bits, and ATen's own fp32 garbage bits
//     must be reproduced) only 256 bf16 inputs exist in (-8, -2]: a 1 KB shared-memory table
//     filled with ATen's formula at kernel start; below -8 the result is exactly -0.
#pragma once
#include "common.cuh"

namespace ab200 {

// MUFU.EX2 / MUFU.RCP without the denormal fix-up code (inputs here are never denormal-sensitive)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16x2 <-> 2 x fp32 (element 0 = low half)
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// packed fp32 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 work on aligned register pairs): two
// independent IEEE fp32 operations per issue slot, bit-identical to the scalar forms
struct f2 { float a, b; };
__device__ __forceinline__ f2 fma2(f2 x, f2 y, f2 z) {
  f2 d;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.a), "=f"(d.b)
      : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b), "f"(z.a), "f"(z.b));
  return d;
}
__device__ __forceinline__ f2 mul2f(f2 x, f2 y) {
  f2 d;
  asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; "
      "mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.a), "=f"(d.b)
      : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b));
  return d;
}
__device__ __forceinline__ f2 add2f(f2 x, f2 y) {
  f2 d;
  asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; "
      "add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.a), "=f"(d.b)
      : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b));
  return d;
}
__device__ __forceinline__ f2 bcast2(float v) { return f2{v, v}; }
__device__ __forceinline__ f2 unpack2(uint32_t u) { return f2{bf_lo(u), bf_hi(u)}; }

// packed bf16x2 comparisons: one HSETP2 yields the predicates of both halves (lo = element 0)
#define AB200_SETP2(NAME, OP)                                                                  \
  __device__ __forceinline__ void NAME(uint32_t a, uint32_t b, bool& lo, bool& hi) {           \
    uint32_t l, h;                                                                             \
    asm("{ .reg .pred p, q; setp." OP ".bf16x2 p|q, %2, %3; selp.u32 %0, 1, 0, p; "            \
        "selp.u32 %1, 1, 0, q; }"                                                              \
        : "=r"(l), "=r"(h)                                                                     \
        : "r"(a), "r"(b));                                                                     \
    lo = l != 0u;                                                                              \
    hi = h != 0u;                                                                              \
  }
AB200_SETP2(setp2_ge, "ge")
AB200_SETP2(setp2_gt, "gt")
AB200_SETP2(setp2_lt, "lt")
#undef AB200_SETP2

// true iff the low 16 bits of v are within K of the bf16 rounding tie 0x8000:
// ((bits + K - 0x8000) mod 2^16) <= 2K, evaluated in the top half of a 32-bit word (IMAD + ISETP)
template <uint32_t K>
__device__ __forceinline__ bool near_bf16_tie(float v) {
  return __float_as_uint(v) * 0x10000u + ((K - 0x8000u) << 16) <= ((2u * K) << 16);
}

// ATen CUDA functors, verbatim arithmetic (the slow, exact paths)
static __device__ __noinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }
static __device__ __noinline__ float gelu_erf_aten(float x) {
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}

#ifndef AB200_SIGMOID_GUARD
#define AB200_SIGMOID_GUARD 20
#endif
#ifndef AB200_GELU_GUARD
#define AB200_GELU_GUARD 28
#endif
// Guards in fp32 ulps.  Measured on all 65 536 inputs (tests/test_cell_gpu.py prints them):
// max |fast - ATen| = 12 (sigmoid, x >= -16) and 15 (GELU, y > -2).
constexpr uint32_t kSigmoidGuard = AB200_SIGMOID_GUARD;
constexpr uint32_t kGeluGuard = AB200_GELU_GUARD;
constexpr int kGeluLutSize = 256;                          // the bf16 values in (-8, -2]

// Fast sigmoid: 2 MUFU + 2 FP ops.  The error is dominated by the fp32 rounding of x*log2(e)
// (<= |x| * 1.44 * 2^-24 * ln2 relative), hence the range limit x >= -16.
__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f));
}
__device__ __forceinline__ bool sigmoid_needs_exact(float x, float s) {
  return near_bf16_tie<kSigmoidGuard>(s) || !(x >= -16.0f);
}
// x0, x1 hold bf16 values.  Returns bf16x2 {RN(sigmoid(x0)), RN(sigmoid(x1))}, bit-equal to
// torch.sigmoid on a bf16 CUDA tensor.
__device__ __forceinline__ uint32_t sigmoid2_bf16_exact(float x0, float x1) {
  float s0 = sigmoid_fast(x0), s1 = sigmoid_fast(x1);
  const bool b0 = sigmoid_needs_exact(x0, s0), b1 = sigmoid_needs_exact(x1, s1);
  if (b0 || b1) {   // ~0.1 % of pairs
    if (b0) s0 = sigmoid_aten(x0);
    if (b1) s1 = sigmoid_aten(x1);
  }
  return pack_rn(s0, s1);
}
// Three gates of one element pair with a single (rarely taken) branch.  GUARD = false skips the
// exactness guard ("faithful" mode: a 1-bf16-ulp difference from ATen in ~0.1 % of elements).
template <bool GUARD>
__device__ __forceinline__ void sigmoid6_bf16_exact(uint32_t af, uint32_t ai, uint32_t ao,
                                                    uint32_t& f, uint32_t& i, uint32_t& o) {
  constexpr uint32_t kM16 = 0xC180C180u;   // {-16, -16}
  const uint32_t a[3] = {af, ai, ao};
  float x[6], s[6];
  bool any = false, bad[6], in_range[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const f2 xx = unpack2(a[k]);
    const f2 t = mul2f(xx, bcast2(-1.4426950408889634f));
    const f2 d = add2f(f2{ex2_approx(t.a), ex2_approx(t.b)}, bcast2(1.0f));
    x[2 * k] = xx.a;
    x[2 * k + 1] = xx.b;
    s[2 * k] = rcp_approx(d.a);
    s[2 * k + 1] = rcp_approx(d.b);
    if constexpr (GUARD) setp2_ge(a[k], kM16, in_range[2 * k], in_range[2 * k + 1]);
  }
  if constexpr (GUARD) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      bad[k] = near_bf16_tie<kSigmoidGuard>(s[k]) || !in_range[k];
      any = any || bad[k];
    }
    if (any) {
#pragma unroll
      for (int k = 0; k < 6; ++k)
        if (bad[k]) s[k] = sigmoid_aten(x[k]);
    }
  }
  f = pack_rn(s[0], s[1]);
  i = pack_rn(s[2], s[3]);
  o = pack_rn(s[4], s[5]);
}

// ATen's GELU of the i-th bf16 value counted down from -2.0 (index 0) to just above -8.0
__device__ __forceinline__ float gelu_lut_entry(int i) {
  return gelu_erf_aten(__uint_as_float((0xC000u + static_cast<uint32_t>(i)) << 16));
}
// copies the per-device table (filled once by gelu_lut_init_kernel) into shared memory
__device__ __forceinline__ void gelu_lut_fill(float* lut, const float* __restrict__ table) {
#pragma unroll 1
  for (int i = threadIdx.x; i < kGeluLutSize; i += blockDim.x) lut[i] = __ldg(table + i);
}

// GELU(y) = y * Phi(y).  With u = |y|: erfc(u/sqrt2)/2 = 2^(u*Q(u) - 1), Q = degree-6 fit of
// log2(erfc(u/sqrt2))/u on [0, 4*sqrt2] (relative error <= 14 fp32 ulps for u <= 2, absolute
// <= 2e-8 above); Phi(y) = that for y < 0 (no cancellation) and 1 - that for y >= 0.
__device__ __forceinline__ float gelu_fast(float y, bool negative) {
  const float u = fminf(fabsf(y), 5.6568542494923806f);
  float q = 4.4679095481e-06f;
  q = fmaf(q, u, -1.4861870113e-05f);
  q = fmaf(q, u, -5.6600669632e-04f);
  q = fmaf(q, u, 7.6454123482e-03f);
  q = fmaf(q, u, -5.2912395447e-02f);
  q = fmaf(q, u, -4.5906913280e-01f);
  q = fmaf(q, u, -1.1511206627e+00f);
  const float h = ex2_approx(fmaf(u, q, -1.0f));
  return y * (negative ? h : 1.0f - h);
}
// y in (-8, -2]: ATen's 1 + erff(..) cancels and its fp32 noise decides the bf16 result -> the
// table (256 inputs).  `lut_s` = 32-bit shared-window address of the table (one LDS, no generic
// address arithmetic).
__device__ __forceinline__ float gelu_fast_lut(float y, bool negative, uint32_t lut_s) {
  float g = gelu_fast(y, negative);
  const uint32_t off = (__float_as_uint(y) >> 14) - (0xC000u << 2);   // byte offset into the table
  if (off < static_cast<uint32_t>(kGeluLutSize) * 4u)
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(lut_s + off));
  return g;
}
// erfc(|y|/sqrt2)/2 for a pair, the polynomial in packed fp32
__device__ __forceinline__ f2 half_erfc2(f2 y) {
  const f2 u = f2{fminf(fabsf(y.a), 5.6568542494923806f), fminf(fabsf(y.b), 5.6568542494923806f)};
  f2 q = bcast2(4.4679095481e-06f);
  q = fma2(q, u, bcast2(-1.4861870113e-05f));
  q = fma2(q, u, bcast2(-5.6600669632e-04f));
  q = fma2(q, u, bcast2(7.6454123482e-03f));
  q = fma2(q, u, bcast2(-5.2912395447e-02f));
  q = fma2(q, u, bcast2(-4.5906913280e-01f));
  q = fma2(q, u, bcast2(-1.1511206627e+00f));
  q = fma2(q, u, bcast2(-1.0f));
  return f2{ex2_approx(q.a), ex2_approx(q.b)};
}
__device__ __forceinline__ float gelu_lut_or(float g, float y, uint32_t lut_s) {
  const uint32_t off = (__float_as_uint(y) >> 14) - (0xC000u << 2);   // byte offset into the table
  if (off < static_cast<uint32_t>(kGeluLutSize) * 4u)
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(lut_s + off));
  return g;
}
// y2 = bf16x2.  Returns bf16x2 of ATen's erf-GELU, bit-equal to F.gelu on a bf16 CUDA tensor.
// At and below -8 (and for NaN) ATen's formula itself runs: it yields exactly -0 there.
template <bool GUARD>
__device__ __forceinline__ uint32_t gelu2_bf16_exact(uint32_t y2, uint32_t lut_s) {
  const f2 y = unpack2(y2);
  bool n0, n1;
  setp2_lt(y2, 0u, n0, n1);               // y < +0
  const f2 h = half_erfc2(y);
  f2 g = mul2f(y, f2{n0 ? h.a : 1.0f - h.a, n1 ? h.b : 1.0f - h.b});
  if constexpr (GUARD) {
    bool r0, r1;
    setp2_gt(y2, 0xC100C100u, r0, r1);    // y > -8
    g.a = gelu_lut_or(g.a, y.a, lut_s);
    g.b = gelu_lut_or(g.b, y.b, lut_s);
    const bool b0 = near_bf16_tie<kGeluGuard>(g.a) || !r0, b1 = near_bf16_tie<kGeluGuard>(g.b) || !r1;
    if (b0 || b1) {
      if (b0) g.a = gelu_erf_aten(y.a);
      if (b1) g.b = gelu_erf_aten(y.b);
    }
  }
  return pack_rn(g.a, g.b);
}

}  // namespace ab200
