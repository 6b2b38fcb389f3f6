// kv_common.cuh - device/host helpers shared by the paged-KV-cache writers (kv_write.cu, kv_rope_write.cu)
#pragma once

#include "common.cuh"

namespace ab200 {

template <int VEC> struct VecT;
template <> struct VecT<16> { using type = uint4; };
template <> struct VecT<8> { using type = uint2; };
template <> struct VecT<4> { using type = uint32_t; };
template <> struct VecT<2> { using type = uint16_t; };

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type kv_ld(const char* p) {
  if constexpr (VEC == 16) {
    return ld_stream_16(p);
  } else {
    return __ldg(reinterpret_cast<const typename VecT<VEC>::type*>(p));
  }
}
template <int VEC>
__device__ __forceinline__ void kv_st(char* p, const typename VecT<VEC>::type& v) {
  if constexpr (VEC == 16) {
    st_stream_16(p, v);
  } else {
    *reinterpret_cast<typename VecT<VEC>::type*>(p) = v;
  }
}

// slot -> (block, offset inside the block): a shift for power-of-two block sizes (block_size < 0
// encodes -(log2 + 1), see encode_block_size), else 32-bit or (for huge caches) 64-bit division -
// once per thread, but a 64-bit division costs as much as moving a whole chunk
__device__ __forceinline__ void slot_to_block(int64_t slot, int32_t block_size, int64_t& blk, int64_t& off) {
  if (block_size < 0) {
    const int sh = -block_size - 1;
    blk = slot >> sh;
    off = slot & ((int64_t{1} << sh) - 1);
  } else if (slot <= 0x7fffffffLL) {
    const uint32_t s32 = static_cast<uint32_t>(slot), b32 = static_cast<uint32_t>(block_size);
    blk = s32 / b32;
    off = s32 - static_cast<uint32_t>(blk) * b32;
  } else {
    blk = slot / block_size;
    off = slot - blk * block_size;
  }
}

// power-of-two block sizes travel as -(log2 + 1) so that the kernel shifts instead of dividing
inline int32_t encode_block_size(int32_t bs) {
  if (bs > 0 && (bs & (bs - 1)) == 0) {
    int sh = 0;
    while ((1 << sh) < bs) ++sh;
    return -(sh + 1);
  }
  return bs;
}

// Two correctly rounded quotients {x0/s, x1/s} in packed fp32 (FMUL2 / FFMA2).  |q0| is clamped to
// 1e6 first: beyond that the fp8 conversion saturates whatever the exact quotient, and the clamp
// keeps inf / overflowing operands on the right side (inf stays inf, NaN re-enters through the
// remainder) without a branch.  copysign: x = -0 has q0 = -0 but (+0 remainder)*r + (-0) = +0.
struct kvf2 { float a, b; };
__device__ __forceinline__ kvf2 div2_rn_by(float x0, float x1, float s, float r) {
  kvf2 q0, rem, q;
  asm("{ .reg .b64 x, rr, t; mov.b64 x, {%2, %3}; mov.b64 rr, {%4, %4}; mul.rn.f32x2 t, x, rr; "
      "mov.b64 {%0, %1}, t; }"
      : "=f"(q0.a), "=f"(q0.b)
      : "f"(x0), "f"(x1), "f"(r));
  const float c0 = fminf(fmaxf(q0.a, -1.0e6f), 1.0e6f), c1 = fminf(fmaxf(q0.b, -1.0e6f), 1.0e6f);
  asm("{ .reg .b64 x, ss, c, t; mov.b64 x, {%2, %3}; mov.b64 ss, {%4, %4}; mov.b64 c, {%5, %6}; "
      "fma.rn.f32x2 t, c, ss, x; mov.b64 {%0, %1}, t; }"
      : "=f"(rem.a), "=f"(rem.b)
      : "f"(x0), "f"(x1), "f"(-s), "f"(c0), "f"(c1));
  asm("{ .reg .b64 m, rr, c, t; mov.b64 m, {%2, %3}; mov.b64 rr, {%4, %4}; mov.b64 c, {%5, %6}; "
      "fma.rn.f32x2 t, m, rr, c; mov.b64 {%0, %1}, t; }"
      : "=f"(q.a), "=f"(q.b)
      : "f"(rem.a), "f"(rem.b), "f"(r), "f"(c0), "f"(c1));
  q.a = copysignf(q.a, q0.a);
  q.b = copysignf(q.b, q0.b);
  return q;
}

template <bool E5M2>
__device__ __forceinline__ uint32_t cvt2_fp8(float lo, float hi) {
  return static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE,
                                                        E5M2 ? __NV_E5M2 : __NV_E4M3));
}

// SRC: AB200_F32 / AB200_F16 / AB200_BF16.  Returns the packed fp8 bytes of one 16-byte chunk
// (8 bytes in .x/.y for 2-byte sources, 4 bytes in .x for f32).
template <int SRC, bool E5M2>
__device__ __forceinline__ uint2 quant_chunk(const uint4& in, float s, float r) {
  uint2 out = make_uint2(0u, 0u);
  if constexpr (SRC == AB200_F32) {
    const kvf2 a = div2_rn_by(__uint_as_float(in.x), __uint_as_float(in.y), s, r);
    const kvf2 b = div2_rn_by(__uint_as_float(in.z), __uint_as_float(in.w), s, r);
    out.x = cvt2_fp8<E5M2>(a.a, a.b) | (cvt2_fp8<E5M2>(b.a, b.b) << 16);
  } else {
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
    uint32_t q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      if constexpr (SRC == AB200_BF16) {
        lo = __uint_as_float(w[j] << 16);
        hi = __uint_as_float(w[j] & 0xffff0000u);
      } else {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
        lo = __low2float(h);
        hi = __high2float(h);
      }
      const kvf2 d = div2_rn_by(lo, hi, s, r);
      q[j] = cvt2_fp8<E5M2>(d.a, d.b);
    }
    out.x = q[0] | (q[1] << 16);
    out.y = q[2] | (q[3] << 16);
  }
  return out;
}

}  // namespace ab200
