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

namespace ab200 {

constexpr int kKvMaxLayers = 128;   // 2 KiB of kernel-parameter space
constexpr int kKvThreads = 256;
constexpr int kKvUnroll = 4;

struct KvPtrTable {
  char* k[kKvMaxLayers];
  char* v[kKvMaxLayers];
};

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
                     uint32_t total_chunks) {   // layers_in_this_launch * chunks_per_row
  using V = typename VecT<VEC>::type;
  const int64_t token = blockIdx.x;
  const int64_t slot = __ldg(slot_mapping + token);
  if (slot < 0) return;  // padded token (kernels.cu:32-35)
  const int64_t blk = slot / block_size;
  const int64_t off = slot - blk * block_size;
  const int64_t dst_row = blk * block_stride_bytes + off * static_cast<int64_t>(row_bytes);
  const char* ksrc = keys + token * key_row_bytes;
  const char* vsrc = values + token * value_row_bytes;

  const uint32_t q0 = blockIdx.y * (kKvThreads * kKvUnroll) + threadIdx.x;
  V kreg[kKvUnroll], vreg[kKvUnroll];
#pragma unroll
  for (int j = 0; j < kKvUnroll; ++j) {
    const uint32_t q = q0 + j * kKvThreads;
    if (q < total_chunks) {
      kreg[j] = kv_ld<VEC>(ksrc + static_cast<int64_t>(q) * VEC);
      vreg[j] = kv_ld<VEC>(vsrc + static_cast<int64_t>(q) * VEC);
    }
  }
#pragma unroll
  for (int j = 0; j < kKvUnroll; ++j) {
    const uint32_t q = q0 + j * kKvThreads;
    if (q < total_chunks) {
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
  const uint32_t per_cta = kKvThreads * kKvUnroll;
  dim3 grid(static_cast<unsigned>(num_tokens), (total + per_cta - 1) / per_cta);
  kv_write_bulk_kernel<VEC><<<grid, kKvThreads, 0, stream>>>(
      keys, values, krb, vrb, tbl, slot_mapping, bsb, block_size, row_bytes, cpr, shift, total);
  AB200_CHECK_LAUNCH("kv_write_bulk_kernel");
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
  (void)k_scale_ptrs;
  (void)v_scale_ptrs;
  if (num_layers == 0 || num_tokens == 0) return AB200_OK;  // kernels.cu:99-101
  AB200_REQUIRE(num_layers > 0 && num_tokens > 0, AB200_E_INVALID,
                "kv_write_bulk: negative num_layers/num_tokens");
  AB200_REQUIRE(src_dtype == AB200_F32 || src_dtype == AB200_F16 || src_dtype == AB200_BF16,
                AB200_E_INVALID, "kv_write_bulk: unsupported src dtype %d", src_dtype);
  AB200_REQUIRE(kv_dtype == AB200_KV_AUTO, AB200_E_UNSUPPORTED,
                "kv_write_bulk: fp8 kv_cache_dtype is not built (dead code in the reference: "
                "ENABLE_FP8 undefined)");
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
