// bf16_exact.cuh - bit-exact emulation of PyTorch-eager bf16 elementwise arithmetic at a fraction
// of the instruction count (used by the fused LSTM cell kernel, lstm_cell.cu).
//
// PyTorch eager computes every bf16 elementwise op in fp32 and rounds its output to bf16
// (SURVEY.md 7-H1).  Two facts make that cheap to reproduce exactly:
//
//  1. For bf16 operands, fp32 add / mul followed by a bf16 rounding equals the correctly rounded
//     bf16 result (double rounding is innocuous when the wide format has >= 2p+2 = 18 significand
//     bits; fp32 has 24).  So R(a*b), R(a+b) on bf16 values ARE `mul.rn.bf16x2` / `add.rn.bf16x2`:
//     two elements per instruction and no conversion.  (`.rn` forbids contraction into an fma.)
//  2. sigmoid and GELU take a bf16 INPUT and produce a bf16 OUTPUT: functions of 65 536 values, so
//     "is this shortcut exact" is a finite question that the device answers exhaustively:
//     * sigmoid: RN_bf16(rcp.approx(1 + ex2.approx(-x*log2 e))) equals RN_bf16 of ATen's
//       1/(1+expf(-x)) for EVERY bf16 x >= -87 - the approximation is within 12 fp32 ulps and no
//       input has a rounding tie inside that gap (tools/sigmoid_search.py; ab200_debug_sigmoid_search).
//       Only x < -87 (bf16-denormal results, flushed by .ftz) and NaN run ATen's formula, behind one
//       packed range test per 16-byte vector.
//     * GELU: an 8 KB shared-memory table of ATen's own results for the 2 x 2048 inputs with
//       2^-11 <= |y| < 32 (one LDS per element, no arithmetic); everything outside runs ATen's
//       formula.  (The polynomial erfc path kept for the "faithful" mode and the MLP cell is itself
//       exact for y > -3, tools/gelu_search.py; below that ATen's 1 + erf(..) cancels and its fp32
//       noise decides the bf16 result, which only a table of ATen's outputs can reproduce.)
//     tests/test_cell_gpu.py checks ALL 65 536 inputs of every variant against torch on the device
//     on every run, so exactness is proven for the hardware the tests run on, not estimated.
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

// ATen CUDA functors, verbatim arithmetic (the slow, exact paths)
static __device__ __noinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }
static __device__ __noinline__ float gelu_erf_aten(float x) {
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}

// Measured on all 65 536 inputs (tests/test_cell_gpu.py prints it): max |fast - ATen| = 12 fp32 ulps
// (sigmoid, x >= -16) and 15 (GELU, y > -2) - and, decisively, no input where that error straddles a
// bf16 rounding tie, so neither function needs a near-tie guard (see below).
constexpr int kGeluLutSize = 256;                          // the bf16 values in (-8, -2]

// Fast sigmoid: 2 MUFU + 2 FP ops (<= 12 fp32 ulps from ATen's value for x >= -16).
__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f));
}
__device__ __forceinline__ bool sigmoid_needs_exact(float x) { return !(x >= -87.0f); }
// x0, x1 hold bf16 values.  Returns bf16x2 {RN(sigmoid(x0)), RN(sigmoid(x1))}, bit-equal to
// torch.sigmoid on a bf16 CUDA tensor.
__device__ __forceinline__ uint32_t sigmoid2_bf16_exact(float x0, float x1) {
  float s0 = sigmoid_fast(x0), s1 = sigmoid_fast(x1);
  const bool b0 = sigmoid_needs_exact(x0), b1 = sigmoid_needs_exact(x1);
  if (b0 || b1) {
    if (b0) s0 = sigmoid_aten(x0);
    if (b1) s1 = sigmoid_aten(x1);
  }
  return pack_rn(s0, s1);
}
// Three gates of one element pair.  The fast path IS exact for every bf16 input x >= -87: an
// exhaustive search on the device (tools/sigmoid_search.py, ab200_debug_sigmoid_search) found that
// RN_bf16(rcp.approx(1 + ex2.approx(x * -log2e))) equals RN_bf16 of ATen's 1/(1+expf(-x)) on all of
// them - no input lands between the approximation and a rounding tie - and tests/test_cell_gpu.py
// re-proves it on all 65 536 inputs on every run.  Below -87 (results in the bf16 denormal range,
// where .ftz flushes) and for NaN, ATen's formula runs.  So the guard is one range test on the
// packed NaN-propagating minimum of the six inputs.  GUARD = false ("faithful" mode) skips it.
template <bool GUARD>
__device__ __forceinline__ void sigmoid6_bf16_exact(uint32_t af, uint32_t ai, uint32_t ao,
                                                    uint32_t& f, uint32_t& i, uint32_t& o) {
  constexpr uint32_t kLow = 0xC2AEC2AEu;   // {-87, -87}
  const uint32_t a[3] = {af, ai, ao};
  float s[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const f2 t = mul2f(unpack2(a[k]), bcast2(-1.4426950408889634f));
    const f2 d = add2f(f2{ex2_approx(t.a), ex2_approx(t.b)}, bcast2(1.0f));
    s[2 * k] = rcp_approx(d.a);
    s[2 * k + 1] = rcp_approx(d.b);
  }
  if constexpr (GUARD) {
    uint32_t m;
    asm("{ .reg .b32 t; min.NaN.bf16x2 t, %1, %2; min.NaN.bf16x2 %0, t, %3; }"
        : "=r"(m) : "r"(af), "r"(ai), "r"(ao));
    bool lo, hi;
    setp2_ge(m, kLow, lo, hi);
    if (!(lo && hi)) {   // never taken on real activations
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float x = (k & 1) ? bf_hi(a[k >> 1]) : bf_lo(a[k >> 1]);
        if (!(x >= -87.0f)) s[k] = sigmoid_aten(x);
      }
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
// y <= -3 or NaN: (-8, -3] from the table, the rest by ATen's formula
__device__ __forceinline__ float gelu_cold(float y, uint32_t lut_s) {
  const uint32_t off = (__float_as_uint(y) >> 14) - (0xC000u << 2);
  float g;
  if (off < static_cast<uint32_t>(kGeluLutSize) * 4u) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(lut_s + off));
  else g = gelu_erf_aten(y);
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
    // y > -3: the polynomial path is exact as it stands (exhaustive search, tools/gelu_search.py: no
    // bf16 input lands between it and a rounding tie; first failure at -3.14).  One packed compare
    // per pair; the rest (0.13 % of LayerNorm outputs) reads the table or runs ATen's formula.
    bool r0, r1;
    setp2_gt(y2, 0xC040C040u, r0, r1);    // y > -3
    if (!(r0 && r1)) {
      if (!r0) g.a = gelu_cold(y.a, lut_s);
      if (!r1) g.b = gelu_cold(y.b, lut_s);
    }
  }
  return pack_rn(g.a, g.b);
}

// ------------------------------------------------------------------------------------------
// GELU by table.  GELU maps a bf16 to a bf16, and all inputs with 2^-11 <= |y| < 32 (where neither
// "0.5*y" nor "y" nor "-0" is the trivially right answer) are 2 x 2048 bit patterns: an 8 KB
// table of ATen's own results, indexed by 2*(|bits| - lo) + sign, replaces the polynomial, the
// near-tie guard and the cancellation table above with one LDS per element.  Inputs outside the
// window (|y| < 2^-11; |y| >= 32, NaN) have closed forms, see
// gelu_bits_outside.
constexpr uint32_t kGeluTabLo = 0x3a00u;                      // bf16 bits of 2^-11
constexpr uint32_t kGeluTabHi = 0x4200u;                      // bf16 bits of 32.0
constexpr int kGeluTabEntries = 2 * (kGeluTabHi - kGeluTabLo);   // 4096 x u16
__device__ __forceinline__ uint16_t gelu_tab_entry(int i) {
  const uint32_t bits = (kGeluTabLo + (static_cast<uint32_t>(i) >> 1)) | ((static_cast<uint32_t>(i) & 1u) << 15);
  const float g = gelu_erf_aten(__uint_as_float(bits << 16));
  return static_cast<uint16_t>(pack_rn(g, 0.f) & 0xffffu);
}
// copies the per-device table into shared memory (16-byte vectors)
__device__ __forceinline__ void gelu_tab_fill(uint16_t* tab, const uint16_t* __restrict__ table) {
#pragma unroll 1
  for (int i = threadIdx.x; i < kGeluTabEntries / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(tab)[i] = __ldg(reinterpret_cast<const uint4*>(table) + i);
}
// Outside the table window ATen's formula has closed forms (checked, like the table, on all 65 536
// inputs): |y| < 2^-11: 1 + erf(..) is within 2^-11 of 1, so the product rounds to RN(y/2);
// y >= 32: erf = 1 exactly -> y; y <= -32: erf = -1 exactly -> y/2 * 0 = -0, except -inf * 0 = NaN;
// NaN stays NaN.  No transcendental, no call.  The window is placed so that both tails are rare:
// |y| < 2^-11 is one in ~2 500 LayerNorm outputs, |y| >= 32 needs a 30-sigma outlier.
__device__ __forceinline__ uint32_t gelu_bits_outside(uint32_t bits16) {
  const uint32_t a = bits16 & 0x7fffu;
  if (a < kGeluTabLo) return pack_rn(0.5f * __uint_as_float(bits16 << 16), 0.f) & 0xffffu;
  if (a > 0x7f80u || !(bits16 & 0x8000u)) return bits16;   // NaN, or y >= 32 (incl. +inf)
  return a == 0x7f80u ? 0x7fc0u : 0x8000u;                 // -inf -> NaN, finite y <= -32 -> -0
}
// y2 = bf16x2; tab_s = 32-bit shared-window address of the table.  Branch-free for in-window pairs:
// 3 ops rotate each half's sign down to bit 0, one packed subtraction rebases both halves, one
// LOP3 tests both windows, two LDS.U16 fetch the results.
__device__ __forceinline__ uint32_t gelu2_bf16_table(uint32_t y2, uint32_t tab_s) {
  uint32_t w;   // per 16-bit half: 2*|bits| + sign  ==  (y2 << 1 & 0xfffefffe) | (y2 >> 15 & 0x00010001)
  asm("lop3.b32 %0, %1, %2, 0xfffefffe, 0xE4;" : "=r"(w) : "r"(y2 << 1), "r"(y2 >> 15));
  // both halves minus 2*lo in one subtraction: a borrow out of the low half only happens when the
  // low half is outside the window, and then the whole pair takes the slow path anyway
  const uint32_t d = w - 2u * kGeluTabLo * 0x10001u;
  constexpr uint32_t kOut = ~(2u * (kGeluTabHi - kGeluTabLo) - 1u) & 0xffffu;   // 0xf000
  uint32_t r0, r1;
  if ((d & (kOut * 0x10001u)) == 0u) {
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r0) : "r"(tab_s + ((d + d) & 0x1ffeu)));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r1) : "r"(tab_s + (d >> 15)));   // bit 15 of d is clear here
  } else {
    const uint32_t f0 = (w & 0xffffu) - 2u * kGeluTabLo, f1 = (w >> 16) - 2u * kGeluTabLo;
    if (f0 < 2u * (kGeluTabHi - kGeluTabLo)) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r0) : "r"(tab_s + 2u * f0));
    else r0 = gelu_bits_outside(y2 & 0xffffu);
    if (f1 < 2u * (kGeluTabHi - kGeluTabLo)) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r1) : "r"(tab_s + 2u * f1));
    else r1 = gelu_bits_outside(y2 >> 16);
  }
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(r0), "r"(r1));
  return r;
}
// ------------------------------------------------------------------------------------------
// Eight elements (one 16-byte vector) at a time: the fast paths run unconditionally and ONE packed
// range test over the NaN-propagating minimum of all inputs decides whether the (out-of-line, cold)
// fix-up runs.  Per-pair guards cost five issue slots each (compare, predicate merge, BSSY / BRA /
// BSYNC); per-vector guards cost a quarter of that.
__device__ __forceinline__ uint32_t min2_nan(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("min.NaN.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// cold, out-of-line fix-up: a whole 16-byte vector in and out (register ABI, one CALL)
static __device__ __noinline__ uint4 sigmoid8_fix(uint4 a) {
  return make_uint4(sigmoid2_bf16_exact(bf_lo(a.x), bf_hi(a.x)), sigmoid2_bf16_exact(bf_lo(a.y), bf_hi(a.y)),
                    sigmoid2_bf16_exact(bf_lo(a.z), bf_hi(a.z)), sigmoid2_bf16_exact(bf_lo(a.w), bf_hi(a.w)));
}
// three gates of eight elements: f/i/o[p] = RN(sigmoid(af/ai/ao[p])), bit-equal to torch.sigmoid
template <bool GUARD>
__device__ __forceinline__ void sigmoid24_bf16_exact(const uint32_t (&af)[4], const uint32_t (&ai)[4],
                                                     const uint32_t (&ao)[4], uint32_t (&f)[4],
                                                     uint32_t (&i)[4], uint32_t (&o)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) sigmoid6_bf16_exact<false>(af[p], ai[p], ao[p], f[p], i[p], o[p]);
  if constexpr (GUARD) {
    uint32_t m = min2_nan(min2_nan(af[0], ai[0]), ao[0]);
#pragma unroll
    for (int p = 1; p < 4; ++p) m = min2_nan(min2_nan(min2_nan(m, af[p]), ai[p]), ao[p]);
    bool lo, hi;
    setp2_ge(m, 0xC2AEC2AEu, lo, hi);   // all 24 inputs >= -87 (and none NaN)
    if (!(lo && hi)) {                  // never taken on real activations
      const uint4 rf = sigmoid8_fix(make_uint4(af[0], af[1], af[2], af[3]));
      const uint4 ri = sigmoid8_fix(make_uint4(ai[0], ai[1], ai[2], ai[3]));
      const uint4 ro = sigmoid8_fix(make_uint4(ao[0], ao[1], ao[2], ao[3]));
      f[0] = rf.x, f[1] = rf.y, f[2] = rf.z, f[3] = rf.w;
      i[0] = ri.x, i[1] = ri.y, i[2] = ri.z, i[3] = ri.w;
      o[0] = ro.x, o[1] = ro.y, o[2] = ro.z, o[3] = ro.w;
    }
  }
}
// y[p] <- GELU(y[p]) by table: the index arithmetic of all four pairs first, so that the eight
// LDS are in flight together
__device__ __forceinline__ void gelu8_bf16_table(uint32_t (&y)[4], uint32_t tab_s) {
#pragma unroll
  for (int p = 0; p < 4; ++p) y[p] = gelu2_bf16_table(y[p], tab_s);
}

}  // namespace ab200
