// kv_write.cu - SwiftKV multi-layer bulk reshape-and-cache for sm_100a (hot path B).
//
// Replaces csrc/custom_ops/kernels.cu:12-155 of the reference.  Differences in design:
//   * tokens on grid.x (no 65 535-token limit of the reference's grid(L, T)), a 16 KiB slab of the
//     token's contiguous [L*n] source row per CTA on grid.y;
//   * the source row of a token is contiguous across layers, so the read side is one perfectly
//     coalesced stream; every thread moves 4 x 16 B of K and 4 x 16 B of V with all eight loads
//     in flight before the first store (ld.global.nc.L1::no_allocate / st.global.L1::no_allocate);
//   * the per-layer cache pointers travel by value in a __grid_constant__ table: no per-call H2D
//     copies (the reference does four blocking ones), so the call is CUDA-graph capturable;
//   * the slot is read once per CTA instead of once per (layer, token) block, no per-element
//     div/mod by head_size.
// HBM-bound: algorithmic bytes = T_valid * L * n * 2 * (e + e) + 8 * T.
#include "common.cuh"
#include "kv_common.cuh"

namespace ab200 {

constexpr int kKvMaxLayers = 128;   // 2 KiB of kernel-parameter space
constexpr int kKvThreads = 256;
constexpr int kKvUnroll = 4;

struct KvPtrTable {
  char* k[kKvMaxLayers];
  char* v[kKvMaxLayers];
};

// One CTA = one token x one slab of kKvThreads*kKvUnroll vector chunks of the token's K row and
// the same slab of its V row.
template <int VEC>
__global__ void __launch_bounds__(kKvThreads)
kv_write_bulk_kernel(const char* __restrict__ keys, const char* __restrict__ values,
                     int64_t key_row_bytes, int64_t value_row_bytes,
                     const __grid_constant__ KvPtrTable tbl,
                     const int64_t* __restrict__ slot_mapping,
                     int64_t block_stride_bytes, int32_t block_size,
                     uint32_t row_bytes,        // n * e: one (token, layer) row
                     uint32_t chunks_per_row,   // row_bytes / VEC
                     int32_t cpr_shift,         // log2(chunks_per_row) or -1
                     uint32_t total_chunks,     // layers_in_this_launch * chunks_per_row
                     uint32_t per_cta) {        // chunks per CTA: total split evenly over grid.y
  using V = typename VecT<VEC>::type;
  const int64_t token = blockIdx.x;
  const int64_t slot = __ldg(slot_mapping + token);
  if (slot < 0) return;  // padded token (kernels.cu:32-35)
  // block / offset of the slot: a shift for power-of-two block sizes (block_size < 0 encodes
  // -(log2 + 1)), else 32-bit or (for huge caches) 64-bit division - once per thread, but a
  // 64-bit division costs as much as moving a whole chunk
  int64_t blk, off;
  slot_to_block(slot, block_size, blk, off);
  const int64_t dst_row = blk * block_stride_bytes + off * static_cast<int64_t>(row_bytes);
  const char* ksrc = keys + token * key_row_bytes;
  const char* vsrc = values + token * value_row_bytes;

  const uint32_t q0 = blockIdx.y * per_cta + threadIdx.x;
  const uint32_t q_end = min(total_chunks, (blockIdx.y + 1) * per_cta);
  V kreg[kKvUnroll], vreg[kKvUnroll];
#pragma unroll
  for (int j = 0; j < kKvUnroll; ++j) {
    const uint32_t q = q0 + j * kKvThreads;
    if (q < q_end) {
      kreg[j] = kv_ld<VEC>(ksrc + static_cast<int64_t>(q) * VEC);
      vreg[j] = kv_ld<VEC>(vsrc + static_cast<int64_t>(q) * VEC);
    }
  }
#pragma unroll
  for (int j = 0; j < kKvUnroll; ++j) {
    const uint32_t q = q0 + j * kKvThreads;
    if (q < q_end) {
      const uint32_t layer = cpr_shift >= 0 ? (q >> cpr_shift) : (q / chunks_per_row);
      const uint32_t c = q - layer * chunks_per_row;
      const int64_t d = dst_row + static_cast<int64_t>(c) * VEC;
      kv_st<VEC>(tbl.k[layer] + d, kreg[j]);
      kv_st<VEC>(tbl.v[layer] + d, vreg[j]);
    }
  }
}

template <int VEC>
static int launch_kv(const char* keys, const char* values, int64_t krb, int64_t vrb,
                     const KvPtrTable& tbl, int layers, const int64_t* slot_mapping,
                     int64_t num_tokens, int64_t bsb, int32_t block_size, uint32_t row_bytes,
                     cudaStream_t stream) {
  const uint32_t cpr = row_bytes / VEC;
  int shift = -1;
  if ((cpr & (cpr - 1)) == 0) {
    shift = 0;
    while ((1u << shift) < cpr) ++shift;
  }
  const uint32_t total = cpr * static_cast<uint32_t>(layers);
  // split a token's chunks EVENLY over the CTAs of grid.y (18 KiB rows: 2 x 576 chunks, not 1024 + 128)
  const uint32_t cap = kKvThreads * kKvUnroll;
  const uint32_t gy = (total + cap - 1) / cap;
  const uint32_t per_cta = (total + gy - 1) / gy;
  dim3 grid(static_cast<unsigned>(num_tokens), gy);
  kv_write_bulk_kernel<VEC><<<grid, kKvThreads, 0, stream>>>(
      keys, values, krb, vrb, tbl, slot_mapping, bsb, encode_block_size(block_size), row_bytes, cpr, shift, total,
      per_cta);
  AB200_CHECK_LAUNCH("kv_write_bulk_kernel");
  return AB200_OK;
}


// ------------------------------------------------------------------------------------------
// FP8 KV cache (kv_cache_dtype "fp8" / "fp8_e4m3" / "fp8_e5m2") - SURVEY 8f rank 2.
// Spec = the conversion the reference kernel was written for (dead there: ENABLE_FP8 is never
// defined) and that vLLM's per-layer op performs: q = cvt_satfinite_fp8(float(x) / scale_layer)
// (csrc/custom_ops/quant_utils.cuh:456-489, kernels.cu:57-66).  The correctly rounded quotient
// is obtained without the ~10-instruction IEEE division: with r = RN(1/s), q0 = RN(x*r),
// q = RN(q0 + fma(-q0, s, x) * r) is the correctly rounded x/s (Markstein), two at a time in
// packed fp32.  One 16-byte source chunk -> 8 (bf16/f16) or 4 (f32) cache bytes.
// ------------------------------------------------------------------------------------------
constexpr int kKvFp8MaxLayers = 64;   // 4 pointers per layer: 2 KiB of parameter space
struct KvFp8Table {
  char* k[kKvFp8MaxLayers];
  char* v[kKvFp8MaxLayers];
  const float* ks[kKvFp8MaxLayers];
  const float* vs[kKvFp8MaxLayers];
};

template <int SRC, bool E5M2>
__global__ void __launch_bounds__(kKvThreads)
kv_write_bulk_fp8_kernel(const char* __restrict__ keys, const char* __restrict__ values,
                         int64_t key_row_bytes, int64_t value_row_bytes,
                         const __grid_constant__ KvFp8Table tbl,
                         const int64_t* __restrict__ slot_mapping, int64_t block_stride_bytes,
                         int32_t block_size, uint32_t row_elems, uint32_t chunks_per_row,
                         int32_t cpr_shift, uint32_t total_chunks, uint32_t per_cta) {
  constexpr int EPC = SRC == AB200_F32 ? 4 : 8;   // elements (= cache bytes) per 16-byte chunk
  const int64_t token = blockIdx.x;
  const int64_t slot = __ldg(slot_mapping + token);
  if (slot < 0) return;
  // block / offset of the slot: a shift for power-of-two block sizes (block_size < 0 encodes
  // -(log2 + 1)), else 32-bit or (for huge caches) 64-bit division - once per thread, but a
  // 64-bit division costs as much as moving a whole chunk
  int64_t blk, off;
  slot_to_block(slot, block_size, blk, off);
  const int64_t dst_row = blk * block_stride_bytes + off * static_cast<int64_t>(row_elems);
  const char* ksrc = keys + token * key_row_bytes;
  const char* vsrc = values + token * value_row_bytes;
  const uint32_t q0 = blockIdx.y * per_cta + threadIdx.x;
  const uint32_t q_end = min(total_chunks, (blockIdx.y + 1) * per_cta);
  uint4 kreg[kKvUnroll], vreg[kKvUnroll];
#pragma unroll
  for (int j = 0; j < kKvUnroll; ++j) {
    const uint32_t q = q0 + j * kKvThreads;
    if (q < q_end) {
      kreg[j] = ld_stream_16(ksrc + static_cast<int64_t>(q) * 16);
      vreg[j] = ld_stream_16(vsrc + static_cast<int64_t>(q) * 16);
    }
  }
#pragma unroll
  for (int j = 0; j < kKvUnroll; ++j) {
    const uint32_t q = q0 + j * kKvThreads;
    if (q < q_end) {
      const uint32_t layer = cpr_shift >= 0 ? (q >> cpr_shift) : (q / chunks_per_row);
      const uint32_t c = q - layer * chunks_per_row;
      const int64_t d = dst_row + static_cast<int64_t>(c) * EPC;
      const float sk = __ldg(tbl.ks[layer]), sv = __ldg(tbl.vs[layer]);
      const uint2 qk = quant_chunk<SRC, E5M2>(kreg[j], sk, __frcp_rn(sk));
      const uint2 qv = quant_chunk<SRC, E5M2>(vreg[j], sv, __frcp_rn(sv));
      if constexpr (EPC == 8) {
        *reinterpret_cast<uint2*>(tbl.k[layer] + d) = qk;
        *reinterpret_cast<uint2*>(tbl.v[layer] + d) = qv;
      } else {
        *reinterpret_cast<uint32_t*>(tbl.k[layer] + d) = qk.x;
        *reinterpret_cast<uint32_t*>(tbl.v[layer] + d) = qv.x;
      }
    }
  }
}

template <int SRC, bool E5M2>
static int launch_kv_fp8(const char* keys, const char* values, int64_t krb, int64_t vrb,
                         const KvFp8Table& tbl, int layers, const int64_t* slot_mapping,
                         int64_t num_tokens, int64_t bsb, int32_t block_size, uint32_t row_elems,
                         cudaStream_t stream) {
  constexpr int EPC = SRC == AB200_F32 ? 4 : 8;
  const uint32_t cpr = row_elems / EPC;
  int shift = -1;
  if ((cpr & (cpr - 1)) == 0) {
    shift = 0;
    while ((1u << shift) < cpr) ++shift;
  }
  const uint32_t total = cpr * static_cast<uint32_t>(layers);
  const uint32_t cap = kKvThreads * kKvUnroll;
  const uint32_t gy = (total + cap - 1) / cap;
  const uint32_t per_cta = (total + gy - 1) / gy;
  dim3 grid(static_cast<unsigned>(num_tokens), gy);
  kv_write_bulk_fp8_kernel<SRC, E5M2><<<grid, kKvThreads, 0, stream>>>(
      keys, values, krb, vrb, tbl, slot_mapping, bsb, encode_block_size(block_size), row_elems, cpr, shift, total,
      per_cta);
  AB200_CHECK_LAUNCH("kv_write_bulk_fp8_kernel");
  return AB200_OK;
}

static int kv_write_fp8(const void* keys, const void* values, int64_t key_row_stride,
                        int64_t value_row_stride, void* const* key_cache_ptrs,
                        void* const* value_cache_ptrs, const float* const* k_scale_ptrs,
                        const float* const* v_scale_ptrs, int32_t num_layers,
                        const int64_t* slot_mapping, int64_t num_tokens, int64_t block_stride,
                        int32_t block_size, int64_t n, int32_t src_dtype, bool e5m2, cudaStream_t s) {
  AB200_REQUIRE(k_scale_ptrs && v_scale_ptrs, AB200_E_INVALID,
                "kv_write_bulk: fp8 kv cache needs k_scales / v_scales");
  const int e = dtype_bytes(src_dtype);
  const int epc = 16 / e;
  AB200_REQUIRE(n % epc == 0 && key_row_stride % epc == 0 && value_row_stride % epc == 0 &&
                    block_stride % epc == 0 && aligned(keys, 16) && aligned(values, 16),
                AB200_E_INVALID,
                "kv_write_bulk: fp8 kv cache needs num_heads*head_size, the row strides and the block "
                "stride to be multiples of %d and 16-byte aligned sources", epc);
  const int64_t krb = key_row_stride * e, vrb = value_row_stride * e;
  for (int l0 = 0; l0 < num_layers; l0 += kKvFp8MaxLayers) {
    const int layers = num_layers - l0 < kKvFp8MaxLayers ? num_layers - l0 : kKvFp8MaxLayers;
    KvFp8Table tbl;
    for (int l = 0; l < kKvFp8MaxLayers; ++l) {
      const bool in = l < layers;
      tbl.k[l] = in ? static_cast<char*>(key_cache_ptrs[l0 + l]) : nullptr;
      tbl.v[l] = in ? static_cast<char*>(value_cache_ptrs[l0 + l]) : nullptr;
      tbl.ks[l] = in ? k_scale_ptrs[l0 + l] : nullptr;
      tbl.vs[l] = in ? v_scale_ptrs[l0 + l] : nullptr;
      if (in)
        AB200_REQUIRE(tbl.k[l] && tbl.v[l] && tbl.ks[l] && tbl.vs[l] && aligned(tbl.k[l], epc) &&
                          aligned(tbl.v[l], epc),
                      AB200_E_INVALID, "kv_write_bulk: NULL or misaligned cache / scale pointer for layer %d",
                      l0 + l);
    }
    const char* kp = static_cast<const char*>(keys) + static_cast<int64_t>(l0) * n * e;
    const char* vp = static_cast<const char*>(values) + static_cast<int64_t>(l0) * n * e;
    int rc;
#define AB200_KV8(SRC)                                                                           \
  (e5m2 ? launch_kv_fp8<SRC, true>(kp, vp, krb, vrb, tbl, layers, slot_mapping, num_tokens,      \
                                   block_stride, block_size, static_cast<uint32_t>(n), s)        \
        : launch_kv_fp8<SRC, false>(kp, vp, krb, vrb, tbl, layers, slot_mapping, num_tokens,     \
                                    block_stride, block_size, static_cast<uint32_t>(n), s))
    switch (src_dtype) {
      case AB200_F32: rc = AB200_KV8(AB200_F32); break;
      case AB200_F16: rc = AB200_KV8(AB200_F16); break;
      default: rc = AB200_KV8(AB200_BF16); break;
    }
#undef AB200_KV8
    if (rc != AB200_OK) return rc;
  }
  return AB200_OK;
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_kv_write_bulk(const void* keys, const void* values,
                                       int64_t key_row_stride, int64_t value_row_stride,
                                       void* const* key_cache_ptrs, void* const* value_cache_ptrs,
                                       const float* const* k_scale_ptrs,
                                       const float* const* v_scale_ptrs, int32_t num_layers,
                                       const int64_t* slot_mapping, int64_t num_tokens,
                                       int64_t block_stride, int32_t block_size,
                                       int32_t num_heads, int32_t head_size, int32_t src_dtype,
                                       int32_t kv_dtype, void* stream) {
  if (num_layers == 0 || num_tokens == 0) return AB200_OK;  // kernels.cu:99-101
  AB200_REQUIRE(num_layers > 0 && num_tokens > 0, AB200_E_INVALID,
                "kv_write_bulk: negative num_layers/num_tokens");
  AB200_REQUIRE(src_dtype == AB200_F32 || src_dtype == AB200_F16 || src_dtype == AB200_BF16,
                AB200_E_INVALID, "kv_write_bulk: unsupported src dtype %d", src_dtype);
  AB200_REQUIRE(kv_dtype == AB200_KV_AUTO || kv_dtype == AB200_KV_FP8_E4M3 || kv_dtype == AB200_KV_FP8_E5M2,
                AB200_E_UNSUPPORTED, "kv_write_bulk: unsupported kv cache dtype %d", kv_dtype);
  AB200_REQUIRE(keys && values && key_cache_ptrs && value_cache_ptrs && slot_mapping,
                AB200_E_INVALID, "kv_write_bulk: NULL pointer argument");
  AB200_REQUIRE(block_size > 0 && num_heads > 0 && head_size > 0, AB200_E_INVALID,
                "kv_write_bulk: block_size/num_heads/head_size must be positive");
  AB200_REQUIRE(num_tokens <= 0x7fffffffLL, AB200_E_INVALID,
                "kv_write_bulk: num_tokens %lld exceeds grid.x", (long long)num_tokens);
  const int e = dtype_bytes(src_dtype);
  const int64_t n = static_cast<int64_t>(num_heads) * head_size;
  const int64_t row_bytes = n * e;
  AB200_REQUIRE(row_bytes * kKvMaxLayers < (1LL << 31), AB200_E_INVALID,
                "kv_write_bulk: num_heads*head_size too large");
  AB200_REQUIRE(key_row_stride >= n * num_layers && value_row_stride >= n * num_layers,
                AB200_E_INVALID, "kv_write_bulk: source row stride smaller than num_layers*n");
  if (kv_dtype != AB200_KV_AUTO)   // block_stride counts cache elements = bytes
    return kv_write_fp8(keys, values, key_row_stride, value_row_stride, key_cache_ptrs, value_cache_ptrs,
                        k_scale_ptrs, v_scale_ptrs, num_layers, slot_mapping, num_tokens, block_stride,
                        block_size, n, src_dtype, kv_dtype == AB200_KV_FP8_E5M2,
                        static_cast<cudaStream_t>(stream));
  const int64_t krb = key_row_stride * e, vrb = value_row_stride * e, bsb = block_stride * e;

  // widest vector every address involved is aligned to
  int vec = 16;
  auto fits = [&](int v) {
    if (row_bytes % v || krb % v || vrb % v || bsb % v) return false;
    if (!aligned(keys, v) || !aligned(values, v)) return false;
    for (int l = 0; l < num_layers; ++l)
      if (!aligned(key_cache_ptrs[l], v) || !aligned(value_cache_ptrs[l], v)) return false;
    return true;
  };
  while (vec > e && !fits(vec)) vec >>= 1;
  AB200_REQUIRE(fits(vec), AB200_E_INVALID, "kv_write_bulk: pointers not aligned to the element size");

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int l0 = 0; l0 < num_layers; l0 += kKvMaxLayers) {
    const int layers = num_layers - l0 < kKvMaxLayers ? num_layers - l0 : kKvMaxLayers;
    KvPtrTable tbl;
    for (int l = 0; l < layers; ++l) {
      AB200_REQUIRE(key_cache_ptrs[l0 + l] && value_cache_ptrs[l0 + l], AB200_E_INVALID,
                    "kv_write_bulk: NULL cache pointer for layer %d", l0 + l);
      tbl.k[l] = static_cast<char*>(key_cache_ptrs[l0 + l]);
      tbl.v[l] = static_cast<char*>(value_cache_ptrs[l0 + l]);
    }
    for (int l = layers; l < kKvMaxLayers; ++l) tbl.k[l] = tbl.v[l] = nullptr;
    const char* kp = static_cast<const char*>(keys) + static_cast<int64_t>(l0) * row_bytes;
    const char* vp = static_cast<const char*>(values) + static_cast<int64_t>(l0) * row_bytes;
    int rc;
    switch (vec) {
      case 16: rc = launch_kv<16>(kp, vp, krb, vrb, tbl, layers, slot_mapping, num_tokens, bsb, block_size, (uint32_t)row_bytes, s); break;
      case 8: rc = launch_kv<8>(kp, vp, krb, vrb, tbl, layers, slot_mapping, num_tokens, bsb, block_size, (uint32_t)row_bytes, s); break;
      case 4: rc = launch_kv<4>(kp, vp, krb, vrb, tbl, layers, slot_mapping, num_tokens, bsb, block_size, (uint32_t)row_bytes, s); break;
      default: rc = launch_kv<2>(kp, vp, krb, vrb, tbl, layers, slot_mapping, num_tokens, bsb, block_size, (uint32_t)row_bytes, s); break;
    }
    if (rc != AB200_OK) return rc;
  }
  return AB200_OK;
}

// ------------------------------------------------------------------------------------------
// SwiftKV select: gather the surviving ("logit") tokens' rows of several tensors in one launch.
// Replaces the five torch.index_select calls of LlamaSwiftKVModel.swiftkv_select
// (arctic_inference/vllm/swiftkv/llama_swiftkv.py:669-685: hidden_states, residual, positions,
// k_states, v_states by final_logits_indices).  Byte copy, HBM/latency-bound, graph-capturable.
// ------------------------------------------------------------------------------------------
namespace ab200 {
constexpr int kGatherMaxTensors = 8;
struct GatherTable {
  const char* src[kGatherMaxTensors];
  char* dst[kGatherMaxTensors];
  int64_t src_stride[kGatherMaxTensors];   // bytes between rows
  int64_t dst_stride[kGatherMaxTensors];
  uint32_t row_bytes[kGatherMaxTensors];
  uint32_t chunk_begin[kGatherMaxTensors + 1];   // prefix sums of 16-byte chunks (rounded up) per row
  int32_t vec16[kGatherMaxTensors];        // 1: 16-byte vector path is legal for this tensor
  int32_t count;
};

constexpr int kGatherUnroll = 4;

// One CTA = one output row x one slab of 256 * kGatherUnroll 16-byte chunks of the concatenated rows
// of all tensors; a thread has up to four independent 16-byte loads in flight before its first store.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __grid_constant__ GatherTable t, const int64_t* __restrict__ indices,
                   int64_t num_src_rows) {
  const int64_t out_row = blockIdx.x;
  int64_t src_row = __ldg(indices + out_row);
  if (src_row < 0) src_row += num_src_rows;                       // torch index_select accepts negatives
  // An out-of-range index is never dereferenced (torch.index_select raises for it; detecting that here would
  // need a device -> host round trip): the output row is ZERO-FILLED instead, so the result is deterministic.
  const bool in_range = src_row >= 0 && src_row < num_src_rows;
  if (!in_range) src_row = 0;
  const uint32_t total = t.chunk_begin[t.count];
  for (uint32_t base = blockIdx.y * (256u * kGatherUnroll) + threadIdx.x; base < total;
       base += gridDim.y * (256u * kGatherUnroll)) {
    uint4 reg[kGatherUnroll];
    char* dptr[kGatherUnroll];
#pragma unroll
    for (int j = 0; j < kGatherUnroll; ++j) {
      const uint32_t q = base + j * 256u;
      dptr[j] = nullptr;
      if (q < total) {
        int k = 0;
#pragma unroll
        for (int i = 1; i < kGatherMaxTensors; ++i)
          if (i < t.count && q >= t.chunk_begin[i]) k = i;
        const uint32_t c = q - t.chunk_begin[k];
        const char* s = t.src[k] + src_row * t.src_stride[k] + static_cast<int64_t>(c) * 16;
        char* d = t.dst[k] + out_row * t.dst_stride[k] + static_cast<int64_t>(c) * 16;
        const uint32_t left = t.row_bytes[k] - c * 16;
        if (t.vec16[k] && left >= 16) {
          reg[j] = in_range ? ld_stream_16(s) : make_uint4(0u, 0u, 0u, 0u);
          dptr[j] = d;
        } else {   // unaligned tensor or ragged row end: bytes
          const uint32_t nb = left < 16 ? left : 16;
          for (uint32_t b = 0; b < nb; ++b) d[b] = in_range ? s[b] : 0;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kGatherUnroll; ++j)
      if (dptr[j] != nullptr) st_stream_16(dptr[j], reg[j]);
  }
}
}  // namespace ab200

extern "C" int32_t ab200_gather_rows(const void* const* src_ptrs, void* const* dst_ptrs,
                                     const int64_t* src_row_strides_bytes,
                                     const int64_t* dst_row_strides_bytes, const int64_t* row_bytes,
                                     int32_t num_tensors, const int64_t* indices, int64_t num_indices,
                                     int64_t num_src_rows, void* stream) {
  if (num_tensors == 0 || num_indices == 0) return AB200_OK;
  AB200_REQUIRE(num_tensors > 0 && num_tensors <= kGatherMaxTensors, AB200_E_INVALID,
                "gather_rows: 1..%d tensors per call", kGatherMaxTensors);
  AB200_REQUIRE(src_ptrs && dst_ptrs && src_row_strides_bytes && dst_row_strides_bytes && row_bytes &&
                    indices && num_indices > 0 && num_src_rows > 0 && num_indices <= 0x7fffffffLL,
                AB200_E_INVALID, "gather_rows: bad arguments");
  GatherTable t;
  uint32_t acc = 0;
  for (int i = 0; i < kGatherMaxTensors; ++i) {
    const bool in = i < num_tensors;
    t.src[i] = in ? static_cast<const char*>(src_ptrs[i]) : nullptr;
    t.dst[i] = in ? static_cast<char*>(dst_ptrs[i]) : nullptr;
    t.src_stride[i] = in ? src_row_strides_bytes[i] : 0;
    t.dst_stride[i] = in ? dst_row_strides_bytes[i] : 0;
    t.row_bytes[i] = in ? static_cast<uint32_t>(row_bytes[i]) : 0;
    t.chunk_begin[i] = acc;
    t.vec16[i] = 0;
    if (in) {
      AB200_REQUIRE(t.src[i] && t.dst[i] && row_bytes[i] > 0 && row_bytes[i] < (1LL << 31), AB200_E_INVALID,
                    "gather_rows: NULL pointer or bad row size for tensor %d", i);
      t.vec16[i] = aligned(t.src[i], 16) && aligned(t.dst[i], 16) && t.src_stride[i] % 16 == 0 &&
                   t.dst_stride[i] % 16 == 0;
      acc += (t.row_bytes[i] + 15) / 16;
    }
  }
  t.chunk_begin[kGatherMaxTensors] = acc;
  for (int i = num_tensors; i <= kGatherMaxTensors; ++i) t.chunk_begin[i] = acc;
  t.count = num_tensors;
  const unsigned slab = 256u * kGatherUnroll;
  unsigned gy = (acc + slab - 1) / slab;
  if (gy > 64) gy = 64;
  dim3 grid(static_cast<unsigned>(num_indices), gy);
  gather_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, indices, num_src_rows);
  AB200_CHECK_LAUNCH("gather_rows_kernel");
  return AB200_OK;
}
