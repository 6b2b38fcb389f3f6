// kv_rope_write.cu - SwiftKV: K-rotary + paged-cache scatter for ALL SwiftKV layers in one launch, straight
// from the per-layer kv_proj outputs.
//
// Replaces, for the tokens of one forward pass (arctic_inference/vllm/swiftkv/llama_swiftkv.py):
//   :263-274  for every SwiftKV layer: kv = kv_proj_swiftkv(h); k, v = kv.chunk(2, -1);
//             _, k = rotary_emb(positions, q, k)     (vLLM's in-place rotary kernel: read k, write k)
//             k_states = cat(k...), v_states = cat(v...)        (read k, v; write [T, L*n] twice)
//   :599-628  reshape_and_cache_flash_bulk(k_states, v_states, caches, slot_mapping)  (read [T, L*n]
//             twice, write the caches)
//   :669-685  index_select(k_states / v_states, logits_indices)  (the rows the decode runner consumes)
// The reference moves 10 * T*L*n*e bytes after the projection GEMM; this kernel reads every kv_proj output
// once and writes the cache once (4 * T*L*n*e) - the [T, L*n] staging tensors never exist.  The rows of the
// surviving tokens (one per request) are emitted compactly by extra CTAs of the same launch.
//
// Rotary arithmetic: x' = x * cos - y * sin,  y' = y * cos + x * sin, in one of two rounding contracts:
//   AB200_ROPE_FP32 (default) - what vLLM's rotary_embedding CUDA op (the op the reference reaches through
//     rotary_emb, csrc/pos_encoding_kernels.cu) computes: fp32 internally, ONE rounding to the storage dtype.
//     Measured on the GPU box against the installed wheel's op: bit-identical on every element, bf16 and f16,
//     NeoX and GPT-J (tools/rope_probe.py, tests/test_kv_rope_gpu.py).  The products of two bf16 / f16 values
//     are exact in fp32, so the result is R_storage(R_fp32(x*cos -+ y*sin)) whatever the contraction.
//   AB200_ROPE_STORAGE - vLLM's RotaryEmbedding.forward_native / apply_rotary_emb_torch in the storage dtype
//     (one rounding after every operation): R(R(x*cos) -+ R(y*sin)).  For bf16 / f16 operands an fp32 operation
//     followed by one rounding IS the correctly rounded narrow operation (fp32 has >= 2p+2 significand bits), so
//     this is mul.rn / add.rn on packed bf16x2 / f16x2 with contraction forbidden.  Differs from the CUDA op by the
//     rounding of the two products (on ~28 % of the elements, usually one storage ulp).
//   Either way the cache is BIT-IDENTICAL to the corresponding reference sequence.  NeoX style pairs
//   (i, i + rot/2) inside a head - the partner 16-byte unit lives in lane ^ (rot/16) of the same warp and is
//   fetched by shuffle; GPT-J style pairs (2i, 2i+1) are in-lane.
//
// Work split like kv_write.cu: tokens on grid.x, a token's (layer, 16-byte K unit) space split evenly over
// grid.y; a thread moves kRopeUnroll K units and the matching V units, all 2 x kRopeUnroll 16-byte loads in flight
// before its first store;
// per-layer pointers travel by value.  HBM-bound: algorithmic bytes = T * L * 2n * (e + e_cache) + 16 T.
#include "common.cuh"
#include "kv_common.cuh"

namespace ab200 {

constexpr int kRopeMaxLayers = 64;     // 5 pointers per layer = 2.5 KiB of kernel-parameter space
constexpr int kRopeThreads = 256;
constexpr int kRopeUnroll = 4;          // K units per thread; the same thread moves the matching V units

struct RopeTable {
  const char* kv[kRopeMaxLayers];   // kv_proj output of the layer: rows [k (n) | ... | v (n)]
  char* k[kRopeMaxLayers];          // key cache of the layer
  char* v[kRopeMaxLayers];          // value cache
  const float* ks[kRopeMaxLayers];  // fp8 cache: per-layer scales (device scalars)
  const float* vs[kRopeMaxLayers];
};

struct RopeParams {
  int64_t kv_row_bytes;        // token stride of every kv tensor
  uint32_t v_off_bytes;        // offset of the V half inside a row
  const int64_t* positions;    // [T]
  const char* cos_sin;         // [max_pos, rot_dim]: cos (rot/2) | sin (rot/2), storage dtype
  int64_t cos_sin_row_bytes;
  int64_t max_position;
  uint32_t rot_dim;            // elements rotated per head
  uint32_t head_units;         // head_size / 8
  uint32_t n_units;            // n / 8: 16-byte units of K (and of V) per layer
  int32_t u_shift;             // log2(n_units) when that is a power of two, else -1
  const int64_t* slot_mapping; // [T]
  int64_t num_tokens;
  int64_t block_stride_bytes;  // cache bytes between blocks
  int32_t block_size;          // encoded (kv_common.cuh)
  uint32_t cache_row_bytes;    // n * cache element size
  uint32_t total_units;        // layers * n_units: (layer, K unit) work items (each also moves its V unit)
  uint32_t per_cta;            // units per CTA of grid.y (multiple of kRopeThreads)
  // compact rows of the surviving tokens (optional)
  const int64_t* sel_indices;  // [num_sel] token indices (negative wraps like torch.index_select)
  char* k_states_out;          // [num_sel, L*n] storage dtype
  char* v_states_out;
  int64_t states_row_bytes;
};

template <typename T> struct Pk;   // packed pair arithmetic in the storage dtype, no contraction
template <> struct Pk<__nv_bfloat16> {
  using V2 = __nv_bfloat162;
  static constexpr int kSrc = AB200_BF16;
};
template <> struct Pk<__half> {
  using V2 = __half2;
  static constexpr int kSrc = AB200_F16;
};

template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename T> __device__ __forceinline__ uint32_t pack2_rn(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2_rn<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t pack2_rn<__half>(float lo, float hi) {
  const __half2 r = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&r);
}

// Two packed outputs: mine * c +- other * s, the sign per half from sign_mask (bit 15 / bit 31 set = minus).
// STORAGE = 0: fp32 internally, one rounding (vLLM's CUDA op); 1: one rounding after every operation
// (forward_native in the storage dtype).
template <typename T, int STORAGE>
__device__ __forceinline__ uint32_t rot_word(uint32_t mine, uint32_t other, uint32_t c, uint32_t s, uint32_t sign_mask) {
  if constexpr (STORAGE) {
    using V2 = typename Pk<T>::V2;
    const V2 t1 = __hmul2_rn(*reinterpret_cast<const V2*>(&mine), *reinterpret_cast<const V2*>(&c));
    V2 t2 = __hmul2_rn(*reinterpret_cast<const V2*>(&other), *reinterpret_cast<const V2*>(&s));
    uint32_t t2b = *reinterpret_cast<uint32_t*>(&t2) ^ sign_mask;     // a - b == a + (-b), exactly
    const V2 r = __hadd2_rn(t1, *reinterpret_cast<const V2*>(&t2b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const float2 m = unpack2<T>(mine), o = unpack2<T>(other), cc = unpack2<T>(c), ss = unpack2<T>(s);
    // other * s is exact in fp32 (<= 22 significand bits), so fma(m, c, +-(o*s)) = R_fp32(m*c +- o*s)
    const float tl = __uint_as_float(__float_as_uint(__fmul_rn(o.x, ss.x)) ^ ((sign_mask & 0x8000u) << 16));
    const float th = __uint_as_float(__float_as_uint(__fmul_rn(o.y, ss.y)) ^ (sign_mask & 0x80000000u));
    return pack2_rn<T>(__fmaf_rn(m.x, cc.x, tl), __fmaf_rn(m.y, cc.y, th));
  }
}

// NEOX: 1 = pairs (i, i + rot/2), partner unit by warp shuffle; 0 = GPT-J pairs (2i, 2i+1), in-lane.
// KVD: 0 = cache in the storage dtype, 1 = fp8 e4m3, 2 = fp8 e5m2.  STORAGE: rounding contract (rot_word).
template <typename T, int NEOX, int KVD, int STORAGE>
__global__ void __launch_bounds__(kRopeThreads, 4)
kv_rope_write_kernel(const __grid_constant__ RopeTable tbl, const __grid_constant__ RopeParams p) {
  // ---- which token, and where its rows go ----
  const bool to_cache = static_cast<int64_t>(blockIdx.x) < p.num_tokens;
  int64_t token, dst_row = 0, out_row = 0;
  if (to_cache) {
    token = blockIdx.x;
    const int64_t slot = __ldg(p.slot_mapping + token);
    if (slot < 0) return;                               // padded token (kernels.cu:32-35)
    int64_t blk, off;
    slot_to_block(slot, p.block_size, blk, off);
    dst_row = blk * p.block_stride_bytes + off * static_cast<int64_t>(p.cache_row_bytes);
  } else {
    out_row = static_cast<int64_t>(blockIdx.x) - p.num_tokens;
    token = __ldg(p.sel_indices + out_row);
    if (token < 0) token += p.num_tokens;
    if (token < 0 || token >= p.num_tokens) return;     // never read outside the tensors
  }
  int64_t pos = __ldg(p.positions + token);
  pos = pos < 0 ? 0 : (pos >= p.max_position ? p.max_position - 1 : pos);   // a bad position must not fault
  const char* cs_row = p.cos_sin + pos * p.cos_sin_row_bytes;
  const uint32_t half_units = p.rot_dim >> 4;           // 16-byte units in one half of the rotated span
  const uint32_t rot_units = p.rot_dim >> 3;

  // Work item q = (layer, 16-byte unit u of the layer's K row); the SAME thread also moves unit u of the V row,
  // so K work (exchange + rotation) and V work (copy) never share a branch and no lane shuffles for V.
  const uint32_t q0 = blockIdx.y * p.per_cta + threadIdx.x;
  const uint32_t q_end = min(p.total_units, (blockIdx.y + 1) * p.per_cta);
  const uint32_t U = p.n_units;

  // The thread's unit position inside a head is the same in every iteration (per_cta, kRopeThreads and U are
  // multiples of head_units, a power of two <= 32): its cos / sin vectors are loaded once.
  const uint32_t hu = q0 & (p.head_units - 1);
  uint4 cosv = make_uint4(0, 0, 0, 0), sinv = make_uint4(0, 0, 0, 0);
  uint32_t sign = 0;
  const bool rotates = hu < rot_units;
  if (rotates) {
    if constexpr (NEOX) {
      const bool is_x = hu < half_units;
      const uint32_t i0 = (is_x ? hu : hu - half_units) * 8;             // first pair index of this unit
      cosv = __ldg(reinterpret_cast<const uint4*>(cs_row + i0 * sizeof(T)));
      sinv = __ldg(reinterpret_cast<const uint4*>(cs_row + (p.rot_dim / 2 + i0) * sizeof(T)));
      sign = is_x ? 0x80008000u : 0u;                                    // x' = x c - y s ; y' = y c + x s
    } else {
      // pairs (2i, 2i+1): 4 pairs per unit -> cos / sin of 4 consecutive indices, each duplicated to both halves
      const uint32_t i0 = hu * 4;
      const uint2 c4 = __ldg(reinterpret_cast<const uint2*>(cs_row + i0 * sizeof(T)));
      const uint2 s4 = __ldg(reinterpret_cast<const uint2*>(cs_row + (p.rot_dim / 2 + i0) * sizeof(T)));
      cosv = make_uint4(__byte_perm(c4.x, 0, 0x1010), __byte_perm(c4.x, 0, 0x3232), __byte_perm(c4.y, 0, 0x1010),
                        __byte_perm(c4.y, 0, 0x3232));
      sinv = make_uint4(__byte_perm(s4.x, 0, 0x1010), __byte_perm(s4.x, 0, 0x3232), __byte_perm(s4.y, 0, 0x1010),
                        __byte_perm(s4.y, 0, 0x3232));
      sign = 0x00008000u;                                                // low half: x c - y s ; high: y c + x s
    }
  }

  uint4 kreg[kRopeUnroll], vreg[kRopeUnroll];
  const int64_t tok_off = token * p.kv_row_bytes;
#pragma unroll
  for (int j = 0; j < kRopeUnroll; ++j) {
    const uint32_t q = q0 + j * kRopeThreads;
    if (q < q_end) {
      const uint32_t layer = p.u_shift >= 0 ? q >> p.u_shift : q / U, u = q - layer * U;
      const char* row = tbl.kv[layer] + tok_off + u * 16u;
      kreg[j] = ld_stream_16(row);
      vreg[j] = ld_stream_16(row + p.v_off_bytes);
    } else {
      kreg[j] = make_uint4(0, 0, 0, 0);
      vreg[j] = make_uint4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int j = 0; j < kRopeUnroll; ++j) {
    const uint32_t q = q0 + j * kRopeThreads;
    const bool live = q < q_end;
    uint4 val = kreg[j];
    if constexpr (NEOX) {
      // every lane takes part in the exchange (warp-uniform trip count)
      uint4 other;
      other.x = __shfl_xor_sync(0xffffffffu, val.x, half_units);
      other.y = __shfl_xor_sync(0xffffffffu, val.y, half_units);
      other.z = __shfl_xor_sync(0xffffffffu, val.z, half_units);
      other.w = __shfl_xor_sync(0xffffffffu, val.w, half_units);
      if (rotates) {
        val.x = rot_word<T, STORAGE>(val.x, other.x, cosv.x, sinv.x, sign);
        val.y = rot_word<T, STORAGE>(val.y, other.y, cosv.y, sinv.y, sign);
        val.z = rot_word<T, STORAGE>(val.z, other.z, cosv.z, sinv.z, sign);
        val.w = rot_word<T, STORAGE>(val.w, other.w, cosv.w, sinv.w, sign);
      }
    } else {
      if (rotates) {
        val.x = rot_word<T, STORAGE>(val.x, __byte_perm(val.x, 0, 0x1032), cosv.x, sinv.x, sign);
        val.y = rot_word<T, STORAGE>(val.y, __byte_perm(val.y, 0, 0x1032), cosv.y, sinv.y, sign);
        val.z = rot_word<T, STORAGE>(val.z, __byte_perm(val.z, 0, 0x1032), cosv.z, sinv.z, sign);
        val.w = rot_word<T, STORAGE>(val.w, __byte_perm(val.w, 0, 0x1032), cosv.w, sinv.w, sign);
      }
    }
    if (!live) continue;
    const uint32_t layer = p.u_shift >= 0 ? q >> p.u_shift : q / U, u = q - layer * U;
    if (to_cache) {
      if constexpr (KVD == 0) {
        const int64_t d = dst_row + static_cast<int64_t>(u) * 16;
        st_stream_16(tbl.k[layer] + d, val);
        st_stream_16(tbl.v[layer] + d, vreg[j]);
      } else {
        const float sk = __ldg(tbl.ks[layer]), sv = __ldg(tbl.vs[layer]);
        const uint2 qk = quant_chunk<Pk<T>::kSrc, KVD == 2>(val, sk, __frcp_rn(sk));
        const uint2 qv = quant_chunk<Pk<T>::kSrc, KVD == 2>(vreg[j], sv, __frcp_rn(sv));
        const int64_t d = dst_row + static_cast<int64_t>(u) * 8;
        *reinterpret_cast<uint2*>(tbl.k[layer] + d) = qk;
        *reinterpret_cast<uint2*>(tbl.v[layer] + d) = qv;
      }
    } else {
      const int64_t d = out_row * p.states_row_bytes + (static_cast<int64_t>(layer) * p.n_units + u) * 16;
      st_stream_16(p.k_states_out + d, val);
      st_stream_16(p.v_states_out + d, vreg[j]);
    }
  }
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_kv_rope_write_bulk(const void* const* kv_ptrs, int64_t kv_row_stride, int64_t v_offset,
                                            const int64_t* positions, const void* cos_sin_cache,
                                            int64_t cos_sin_row_stride, int64_t max_position, int32_t rot_dim,
                                            int32_t is_neox, int32_t rope_rounding,
                                            void* const* key_cache_ptrs,
                                            void* const* value_cache_ptrs, const float* const* k_scale_ptrs,
                                            const float* const* v_scale_ptrs, int32_t num_layers,
                                            const int64_t* slot_mapping, int64_t num_tokens, int64_t block_stride,
                                            int32_t block_size, int32_t num_heads, int32_t head_size,
                                            int32_t src_dtype, int32_t kv_dtype, const int64_t* sel_indices,
                                            int64_t num_sel, void* k_states_out, void* v_states_out,
                                            int64_t states_row_stride, void* stream) {
  if (num_layers == 0 || num_tokens == 0) return AB200_OK;
  AB200_REQUIRE(num_layers > 0 && num_tokens > 0 && num_sel >= 0, AB200_E_INVALID,
                "kv_rope_write_bulk: negative num_layers / num_tokens / num_sel");
  AB200_REQUIRE(src_dtype == AB200_F16 || src_dtype == AB200_BF16, AB200_E_UNSUPPORTED,
                "kv_rope_write_bulk: src dtype must be bf16 or f16 (the rotary rounding contract is defined for "
                "16-bit storage)");
  AB200_REQUIRE(kv_dtype == AB200_KV_AUTO || kv_dtype == AB200_KV_FP8_E4M3 || kv_dtype == AB200_KV_FP8_E5M2,
                AB200_E_UNSUPPORTED, "kv_rope_write_bulk: unsupported kv cache dtype %d", kv_dtype);
  AB200_REQUIRE(rope_rounding == AB200_ROPE_FP32 || rope_rounding == AB200_ROPE_STORAGE, AB200_E_INVALID,
                "kv_rope_write_bulk: rope_rounding must be AB200_ROPE_FP32 or AB200_ROPE_STORAGE");
  AB200_REQUIRE(kv_ptrs && positions && cos_sin_cache && key_cache_ptrs && value_cache_ptrs && slot_mapping,
                AB200_E_INVALID, "kv_rope_write_bulk: NULL pointer argument");
  AB200_REQUIRE(kv_dtype == AB200_KV_AUTO || (k_scale_ptrs && v_scale_ptrs), AB200_E_INVALID,
                "kv_rope_write_bulk: fp8 kv cache needs k_scales / v_scales");
  AB200_REQUIRE(block_size > 0 && num_heads > 0 && head_size > 0 && max_position > 0, AB200_E_INVALID,
                "kv_rope_write_bulk: block_size / num_heads / head_size / max_position must be positive");
  AB200_REQUIRE(num_tokens + num_sel <= 0x7fffffffLL, AB200_E_INVALID, "kv_rope_write_bulk: too many tokens for grid.x");
  AB200_REQUIRE(num_sel == 0 || (sel_indices && k_states_out && v_states_out), AB200_E_INVALID,
                "kv_rope_write_bulk: selected rows need indices and both output buffers");
  const int e = 2;
  const int64_t n = static_cast<int64_t>(num_heads) * head_size;
  // geometry the shuffle scheme needs: heads made of a power-of-two number (<= 32) of 16-byte units, the
  // rotated span a power of two of at least one unit per half
  const int hu = head_size / 8;
  AB200_REQUIRE(head_size % 8 == 0 && hu >= 1 && hu <= 32 && (hu & (hu - 1)) == 0, AB200_E_UNSUPPORTED,
                "kv_rope_write_bulk: head_size must be 8, 16, 32, 64, 128 or 256 (got %d)", head_size);
  AB200_REQUIRE(rot_dim >= 16 && rot_dim <= head_size && (rot_dim & (rot_dim - 1)) == 0, AB200_E_UNSUPPORTED,
                "kv_rope_write_bulk: rot_dim must be a power of two in [16, head_size] (got %d)", rot_dim);
  AB200_REQUIRE(kv_row_stride >= v_offset + n && v_offset >= n && kv_row_stride % 8 == 0 && v_offset % 8 == 0,
                AB200_E_INVALID, "kv_rope_write_bulk: kv rows must hold k (n) then v (n) at 16-byte aligned offsets");
  AB200_REQUIRE(cos_sin_row_stride >= rot_dim && cos_sin_row_stride % 8 == 0 && aligned(cos_sin_cache, 16),
                AB200_E_INVALID, "kv_rope_write_bulk: cos_sin_cache rows must be 16-byte aligned and hold rot_dim elements");
  const int ce = kv_dtype == AB200_KV_AUTO ? e : 1;
  AB200_REQUIRE((block_stride * ce) % (kv_dtype == AB200_KV_AUTO ? 16 : 8) == 0, AB200_E_INVALID,
                "kv_rope_write_bulk: cache block stride must keep rows 16-byte (fp8: 8-byte) aligned");
  AB200_REQUIRE(num_sel == 0 || (states_row_stride >= n * num_layers && states_row_stride % 8 == 0 &&
                                 aligned(k_states_out, 16) && aligned(v_states_out, 16)),
                AB200_E_INVALID, "kv_rope_write_bulk: k_states / v_states rows must hold num_layers*n elements, 16-byte aligned");
  AB200_REQUIRE(n * 2 * kRopeMaxLayers / 8 < (1LL << 31), AB200_E_INVALID, "kv_rope_write_bulk: num_heads*head_size too large");

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int l0 = 0; l0 < num_layers; l0 += kRopeMaxLayers) {
    const int layers = num_layers - l0 < kRopeMaxLayers ? num_layers - l0 : kRopeMaxLayers;
    RopeTable tbl;
    for (int l = 0; l < kRopeMaxLayers; ++l) {
      const bool in = l < layers;
      tbl.kv[l] = in ? static_cast<const char*>(kv_ptrs[l0 + l]) : nullptr;
      tbl.k[l] = in ? static_cast<char*>(key_cache_ptrs[l0 + l]) : nullptr;
      tbl.v[l] = in ? static_cast<char*>(value_cache_ptrs[l0 + l]) : nullptr;
      tbl.ks[l] = in && kv_dtype != AB200_KV_AUTO ? k_scale_ptrs[l0 + l] : nullptr;
      tbl.vs[l] = in && kv_dtype != AB200_KV_AUTO ? v_scale_ptrs[l0 + l] : nullptr;
      if (in) {
        AB200_REQUIRE(tbl.kv[l] && tbl.k[l] && tbl.v[l] && aligned(tbl.kv[l], 16) &&
                          aligned(tbl.k[l], kv_dtype == AB200_KV_AUTO ? 16 : 8) &&
                          aligned(tbl.v[l], kv_dtype == AB200_KV_AUTO ? 16 : 8),
                      AB200_E_INVALID, "kv_rope_write_bulk: NULL or misaligned pointer for layer %d", l0 + l);
        AB200_REQUIRE(kv_dtype == AB200_KV_AUTO || (tbl.ks[l] && tbl.vs[l]), AB200_E_INVALID,
                      "kv_rope_write_bulk: NULL scale pointer for layer %d", l0 + l);
      }
    }
    RopeParams p;
    p.kv_row_bytes = kv_row_stride * e;
    p.v_off_bytes = static_cast<uint32_t>(v_offset * e);
    p.positions = positions;
    p.cos_sin = static_cast<const char*>(cos_sin_cache);
    p.cos_sin_row_bytes = cos_sin_row_stride * e;
    p.max_position = max_position;
    p.rot_dim = static_cast<uint32_t>(rot_dim);
    p.head_units = static_cast<uint32_t>(hu);
    p.n_units = static_cast<uint32_t>(n / 8);
    p.u_shift = -1;
    for (int sh = 0; sh < 31; ++sh)
      if ((1u << sh) == p.n_units) p.u_shift = sh;
    p.slot_mapping = slot_mapping;
    p.num_tokens = num_tokens;
    p.block_stride_bytes = block_stride * ce;
    p.block_size = encode_block_size(block_size);
    p.cache_row_bytes = static_cast<uint32_t>(n * ce);
    p.total_units = static_cast<uint32_t>(layers) * p.n_units;
    const uint32_t cap = kRopeThreads * kRopeUnroll;
    const uint32_t gy = (p.total_units + cap - 1) / cap;
    // even split over grid.y, rounded up to whole CTAs-worth of threads so that a thread's unit position
    // inside a head (and its lane's shuffle partner) is the same in every CTA and iteration
    p.per_cta = ((p.total_units + gy - 1) / gy + kRopeThreads - 1) / kRopeThreads * kRopeThreads;
    p.sel_indices = sel_indices;
    p.k_states_out = num_sel ? static_cast<char*>(k_states_out) + static_cast<int64_t>(l0) * n * e : nullptr;
    p.v_states_out = num_sel ? static_cast<char*>(v_states_out) + static_cast<int64_t>(l0) * n * e : nullptr;
    p.states_row_bytes = states_row_stride * e;
    dim3 grid(static_cast<unsigned>(num_tokens + num_sel), gy);
#define AB200_ROPE_LAUNCH(T, NEOX, KVD)                                                                  \
  do {                                                                                                  \
    if (rope_rounding == AB200_ROPE_STORAGE) kv_rope_write_kernel<T, NEOX, KVD, 1><<<grid, kRopeThreads, 0, s>>>(tbl, p); \
    else kv_rope_write_kernel<T, NEOX, KVD, 0><<<grid, kRopeThreads, 0, s>>>(tbl, p);                    \
  } while (0)
#define AB200_ROPE_KVD(T, NEOX)                                   \
  do {                                                            \
    if (kv_dtype == AB200_KV_AUTO) AB200_ROPE_LAUNCH(T, NEOX, 0); \
    else if (kv_dtype == AB200_KV_FP8_E4M3) AB200_ROPE_LAUNCH(T, NEOX, 1); \
    else AB200_ROPE_LAUNCH(T, NEOX, 2);                           \
  } while (0)
    if (src_dtype == AB200_BF16) {
      if (is_neox) AB200_ROPE_KVD(__nv_bfloat16, 1);
      else AB200_ROPE_KVD(__nv_bfloat16, 0);
    } else {
      if (is_neox) AB200_ROPE_KVD(__half, 1);
      else AB200_ROPE_KVD(__half, 0);
    }
#undef AB200_ROPE_KVD
#undef AB200_ROPE_LAUNCH
    AB200_CHECK_LAUNCH("kv_rope_write_kernel");
  }
  return AB200_OK;
}
