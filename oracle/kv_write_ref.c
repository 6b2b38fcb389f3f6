/* TEST INFRASTRUCTURE ONLY - plain-C restatement (oracle) of the reference's bulk
 * reshape-and-cache kernel, /root/reference/csrc/custom_ops/kernels.cu:12-69, with the host
 * launcher's argument derivation (kernels.cu:86-155) left to the caller.
 *
 * One (layer, token) pair of the reference's grid = one iteration of the two outer loops; the
 * inner loop is the reference's per-thread loop including its (algebraically redundant)
 * head_idx / head_offset split (kernels.cu:52-56).  Element type is opaque: the "auto"
 * kv_cache_dtype path is a bit copy (kernels.cu:60-62), so elements are moved as bytes.
 *
 * Parity pinned: against tests/golden/kv_write_*.npz, produced on a B200 by the reference's own
 * kernel compiled for sm_100a (oracle/build_ref.py, oracle/gen_golden_kv_gpu.py).
 * Never linked into the product library. */
#include <stdint.h>
#include <string.h>

int oracle_kv_write_bulk(const void* keys, const void* values,
                         int64_t key_stride, int64_t value_stride, /* elements, kernels.cu:111-112 */
                         void* const* key_cache_ptrs, void* const* value_cache_ptrs,
                         int32_t num_layers,
                         const int64_t* slot_mapping, int64_t num_tokens,
                         int64_t block_stride,                     /* elements, kernels.cu:113 */
                         int32_t block_size, int32_t num_heads, int32_t head_size,
                         int32_t elem_bytes) {
  if (num_layers == 0) return 0;                                   /* kernels.cu:99-101 */
  const int64_t n = (int64_t)num_heads * head_size;                /* kernels.cu:38 */
  const int64_t total = (int64_t)num_layers * num_tokens;
  const char* k = (const char*)keys;
  const char* v = (const char*)values;
#pragma omp parallel for schedule(static)
  for (int64_t lt = 0; lt < total; ++lt) {
    const int64_t token_idx = lt / num_layers;                     /* blockIdx.y */
    const int64_t layer_idx = lt % num_layers;                     /* blockIdx.x */
    const int64_t slot_idx = slot_mapping[token_idx];
    if (slot_idx < 0) continue;                                    /* kernels.cu:32-35 */
    const int64_t block_idx = slot_idx / block_size;
    const int64_t block_offset = slot_idx % block_size;
    char* key_cache = (char*)key_cache_ptrs[layer_idx];
    char* value_cache = (char*)value_cache_ptrs[layer_idx];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t src_key_idx = token_idx * key_stride + layer_idx * n + i;
      const int64_t src_value_idx = token_idx * value_stride + layer_idx * n + i;
      const int64_t head_idx = i / head_size;
      const int64_t head_offset = i % head_size;
      const int64_t tgt = block_idx * block_stride + block_offset * num_heads * head_size +
                          head_idx * head_size + head_offset;      /* kernels.cu:54-56 */
      memcpy(key_cache + tgt * elem_bytes, k + src_key_idx * elem_bytes, (size_t)elem_bytes);
      memcpy(value_cache + tgt * elem_bytes, v + src_value_idx * elem_bytes, (size_t)elem_bytes);
    }
  }
  return 0;
}
