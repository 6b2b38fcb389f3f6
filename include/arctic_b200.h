/* arctic_b200.h - C ABI of libarctic_b200.so: hand-written sm_100a kernels for ArcticInference's
 * data-parallel hot path (Arctic LSTM speculator draft head + SwiftKV bulk KV write).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host".
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All entry
 *     points only enqueue work on that stream: no allocation, no host<->device copy, no
 *     synchronisation, so they can be captured into a CUDA graph.
 *   - return value 0 = success, negative = error (AB200_E_*); the message of the last error on the
 *     calling thread is returned by ab200_last_error().
 *   - strides are in ELEMENTS of the tensor's dtype.
 *   - the library contains no CPU fallback: without an sm_100 device every compute call fails.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * ArcticInference tree).
 */
#ifndef ARCTIC_B200_H_
#define ARCTIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AB200_API __attribute__((visibility("default")))
#else
#define AB200_API
#endif

#define AB200_OK 0
#define AB200_E_INVALID (-1)   /* bad argument (shape, alignment, dtype, NULL)        */
#define AB200_E_UNSUPPORTED (-2) /* valid in the reference but not built here (e.g. fp8 KV) */
#define AB200_E_CUDA (-3)      /* a CUDA runtime/driver call failed                    */
#define AB200_E_WORKSPACE (-4) /* workspace too small                                  */

enum ab200_dtype { AB200_F32 = 0, AB200_F16 = 1, AB200_BF16 = 2 };
enum ab200_kv_cache_dtype { AB200_KV_AUTO = 0, AB200_KV_FP8_E4M3 = 1, AB200_KV_FP8_E5M2 = 2 };

/* Message of the last error raised on this thread ("" if none). Never NULL. */
AB200_API const char* ab200_last_error(void);
/* ABI version (bumped on any signature change). */
AB200_API int32_t ab200_abi_version(void);
/* Number of kernel launches this library has enqueued on this thread since load (bench claim). */
AB200_API int64_t ab200_launch_count(void);

/* One-time per-device initialisation for the CURRENT device: shared-memory attributes of the tcgen05
 * kernels, the constant tables of the cell kernels (filled on `stream`, which is then synchronised)
 * and the cached SM count.  Thread-safe and idempotent.  Every entry point also initialises lazily
 * on first use, but a first use inside a CUDA stream capture is refused (AB200_E_INVALID): call this
 * once per device before capturing graphs (the host classes do, at construction). */
AB200_API int32_t ab200_init(void* stream);

/* Device-side timeline for profiling a (graph-captured) draft loop.  device_slots: AB200_TRACE_WORDS (8)
 * uint64 per traced launch, initialised by the caller to ~0 in the "min" words {0, 1, 3, 5} and 0 elsewhere; each
 * speculator kernel launched while tracing is on takes the next slot and records %globaltimer ns: [0] = earliest
 * CTA entry, [1] = earliest time its dependency wait returned, [2] = latest CTA exit, and (GEMM kernels)
 * [3] = first operand bytes landed in shared memory, [4] = last operand bytes landed, [5] = earliest CTA exit.
 * NULL switches it off.
 * The trace state belongs to the calling host thread (only its own launches are traced): a
 * profiling aid, not part of the product path. */
#define AB200_TRACE_WORDS 8
AB200_API int32_t ab200_set_trace_buffer(uint64_t* device_slots, int32_t max_launches);

/* ------------------------------------------------------------------------------------------
 * Hot path B: SwiftKV multi-layer bulk reshape-and-cache.
 * Replaces: csrc/custom_ops/kernels.cu:86-155 (host launcher reshape_and_cache_flash_bulk) and
 *           kernels.cu:12-69 (kernel), reached through torch op
 *           arctic_inference::reshape_and_cache_flash_bulk (csrc/custom_ops/torch_bindings.cpp:5-19)
 *           from arctic_inference/vllm/swiftkv/llama_swiftkv.py:623-628.
 *
 * For every token t < num_tokens with slot = slot_mapping[t] >= 0 and every layer l:
 *   key_cache_l  [slot / block_size][slot % block_size][0 .. n) = keys  [t][l*n .. (l+1)*n)
 *   value_cache_l[slot / block_size][slot % block_size][0 .. n) = values[t][l*n .. (l+1)*n)
 * with n = num_heads * head_size; a cache block is `block_stride` elements apart (bs*n for vLLM's
 * FlashAttention layout, 2*bs*n for FlashInfer's), dense [block_size, n] inside.  Bit-exact copy
 * for kv_dtype AB200_KV_AUTO.  AB200_KV_FP8_E4M3 / _E5M2 (one-byte cache
 * elements): q = cvt_satfinite_fp8(float(x) / *k_scale_ptrs[l]) - the conversion the reference
 * kernel was written for (csrc/custom_ops/quant_utils.cuh:456-489, kernels.cu:57-66; dead code
 * there because ENABLE_FP8 is never defined) and that vLLM's per-layer op performs; pinned against
 * vLLM's op on a B200 (tests/test_kv_write_fp8_gpu.py).  Needs n and the strides to be multiples
 * of 16 / sizeof(src element).
 *
 * key_cache_ptrs / value_cache_ptrs: HOST arrays of num_layers device pointers.  They are passed
 * to the kernel by value (no per-call H2D copy, unlike kernels.cu:135-146).
 * k_scale_ptrs / v_scale_ptrs: HOST arrays of device float* (one scalar per layer, read on the
 * device; ignored for AUTO, where they may be NULL).
 * num_layers == 0 or num_tokens == 0 is a successful no-op (kernels.cu:99-101).
 */
AB200_API int32_t ab200_kv_write_bulk(const void* keys, const void* values,
                            int64_t key_row_stride, int64_t value_row_stride,
                            void* const* key_cache_ptrs, void* const* value_cache_ptrs,
                            const float* const* k_scale_ptrs, const float* const* v_scale_ptrs,
                            int32_t num_layers,
                            const int64_t* slot_mapping, int64_t num_tokens,
                            int64_t block_stride, int32_t block_size,
                            int32_t num_heads, int32_t head_size,
                            int32_t src_dtype, int32_t kv_dtype, void* stream);

/* SwiftKV: K-rotary + paged-cache scatter for all SwiftKV layers in ONE launch, straight from the per-layer
 * kv_proj outputs - no [T, L*n] k_states / v_states staging tensors.
 * Replaces, per forward pass: the loop arctic_inference/vllm/swiftkv/llama_swiftkv.py:263-274
 *   (kv = kv_proj_swiftkv(h); k, v = kv.chunk(2, -1); _, k = rotary_emb(positions, q, k); cat),
 *   the bulk cache write :599-628 (reshape_and_cache_flash_bulk) and, optionally, the
 *   index_select of k_states / v_states for the surviving tokens :669-685.
 *   kv_ptrs        HOST array of num_layers device pointers: layer l's kv_proj output [T, >= v_offset + n],
 *                  row stride kv_row_stride; K in columns [0, n), V in columns [v_offset, v_offset + n)
 *                  (kv.chunk(2, dim=-1): v_offset = n).  n = num_heads * head_size.
 *   positions      [T] int64;  cos_sin_cache [max_position, >= rot_dim] in the kv dtype, row = cos (rot_dim/2)
 *                  | sin (rot_dim/2) - vLLM's RotaryEmbedding.cos_sin_cache;  is_neox: 1 = pairs (i, i + rot_dim/2),
 *                  0 = GPT-J pairs (2i, 2i+1);  only the first rot_dim elements of every head are rotated.
 *   rope_rounding  AB200_ROPE_FP32 (default contract): x' = R(x*cos - y*sin), y' = R(y*cos + x*sin) with fp32
 *                  arithmetic and ONE rounding to the storage dtype - bit-identical to vLLM's rotary_embedding CUDA
 *                  op (the op the reference reaches through rotary_emb; pinned against the installed wheel's op on
 *                  a B200).  AB200_ROPE_STORAGE: one rounding after every operation, R(R(x*cos) - R(y*sin)) =
 *                  vLLM's forward_native / apply_rotary_emb_torch in the storage dtype (differs by the rounding
 *                  of the two products, usually one storage ulp).  Either way the cache is bit-identical to the corresponding reference sequence; V is a
 *                  bit copy (tests/test_kv_rope_gpu.py).
 *   caches, slot_mapping, block_stride, block_size, kv_dtype, scales: as ab200_kv_write_bulk.
 *   sel_indices    (optional, device [num_sel] int64, negative wraps): the rotated K rows and V rows of these
 *                  tokens are also written compactly to k_states_out / v_states_out [num_sel, num_layers*n]
 *                  (what swiftkv_select hands to the decode runner), by extra CTAs of the same launch.
 * Requires bf16 / f16, head_size a power of two in [8, 256], rot_dim a power of two in [16, head_size]. */
enum ab200_rope_rounding { AB200_ROPE_FP32 = 0, AB200_ROPE_STORAGE = 1 };
AB200_API int32_t ab200_kv_rope_write_bulk(const void* const* kv_ptrs, int64_t kv_row_stride, int64_t v_offset,
                                           const int64_t* positions, const void* cos_sin_cache,
                                           int64_t cos_sin_row_stride, int64_t max_position, int32_t rot_dim,
                                           int32_t is_neox, int32_t rope_rounding, void* const* key_cache_ptrs,
                                           void* const* value_cache_ptrs, const float* const* k_scale_ptrs,
                                           const float* const* v_scale_ptrs, int32_t num_layers,
                                           const int64_t* slot_mapping, int64_t num_tokens, int64_t block_stride,
                                           int32_t block_size, int32_t num_heads, int32_t head_size,
                                           int32_t src_dtype, int32_t kv_dtype, const int64_t* sel_indices,
                                           int64_t num_sel, void* k_states_out, void* v_states_out,
                                           int64_t states_row_stride, void* stream);

/* SwiftKV select: dst_i[j, :] = src_i[indices[j], :] for up to 8 tensors in
 * ONE launch.  Replaces the five torch.index_select calls of LlamaSwiftKVModel.swiftkv_select
 * (arctic_inference/vllm/swiftkv/llama_swiftkv.py:669-685) that compact hidden_states, residual,
 * positions, k_states and v_states to the surviving tokens.  Host arrays of num_tensors entries;
 * strides and row sizes in bytes; indices [num_indices] int64 on the device, each in
 * [-num_src_rows, num_src_rows); an out-of-range index is never dereferenced and its output row is ZERO-FILLED
 * (torch.index_select raises instead - detecting that would need a device -> host round trip). */
AB200_API int32_t ab200_gather_rows(const void* const* src_ptrs, void* const* dst_ptrs,
                                    const int64_t* src_row_strides_bytes,
                                    const int64_t* dst_row_strides_bytes, const int64_t* row_bytes,
                                    int32_t num_tensors, const int64_t* indices, int64_t num_indices,
                                    int64_t num_src_rows, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hot path A: Arctic LSTM speculator draft head (method "sum_lstm").
 * Replaces the eager-PyTorch op sequences inside
 *   arctic_inference/vllm/spec_dec/arctic_speculator.py  (no custom op exists in the reference).
 *
 * `emulate_eager_rounding` (bf16/f16 only): round every intermediate to the storage dtype where
 * PyTorch eager would have materialised a tensor (needed for >= 99.9 % draft-token agreement with
 * the reference's bf16 path); 0 = keep fp32 inside the kernel and round once on output.
 */

/* y = x * rsqrt(mean(x^2, -1) + eps) * post_scale, row-wise.
 * Replaces: MLPSpeculatorLayerNorm.forward without affine (arctic_speculator.py:88-91) followed
 * by "/ SQRT2" (arctic_speculator.py:656-657; post_scale = 1/sqrt(2) in fp32). */
AB200_API int32_t ab200_rms_scale(const void* x, void* y, int64_t rows, int64_t cols,
                        int64_t x_row_stride, int64_t y_row_stride,
                        float eps, float post_scale, int32_t dtype,
                        int32_t emulate_eager_rounding, void* stream);

/* The fused sum + RMS-LayerNorm + LSTM-gate + GELU cell.
 * Replaces: arctic_speculator.py:667-689 (embedding lookup, repeat(1,1,4), torch.add with alpha,
 * split, sigmoid, cell_ln + GELU, input/forget gating, state_ln + GELU, output gate), ~30 ATen
 * launches.
 *   gates      [B, 4*D] gate pre-activations x @ W^T, order f | i | o | c  (row stride given)
 *   emb_table  [vocab, D]
 *   tokens     [B] int64 (used when packed_argmax == NULL)
 *   packed_argmax [B] uint64 as written by ab200_head_argmax (token = 0xFFFFFFFF - low 32 bits);
 *              when non-NULL it is the token source and `tokens` is ignored.
 *   cell_prev  [B, D] or NULL (= zeros, the first step: arctic_speculator.py:781-785)
 *   ln vectors [D]; outputs state_out, cell_out [B, D] (cell_out may alias cell_prev).
 *   clear_packed [B] / clear_counter [1] (optional): zeroed by this kernel so that the head
 *   kernel that follows finds its arg-max slot and completion counter reset (no memset node).
 *   state_q_out [B, D] e4m3 / state_scale_out [B] f32 (optional, bf16 only): dynamic per-token
 *   FP8 copy of the new state for the quantised head (see ab200_head_argmax_fp8).
 *   emulate_eager_rounding (bf16 / f16 only): 0 = fp32 internally, one rounding at the end;
 *   1 = every value PyTorch eager materialises is rounded to the storage dtype at that point and
 *   sigmoid / erf-GELU round exactly as ATen's CUDA kernels do (bit-identical states);
 *   2 (bf16) = the same rounding points with unguarded fast sigmoid / GELU ("faithful": about
 *   0.1 % of the elements differ from mode 1 by one bf16 ulp) for the bandwidth-bound regime.
 *   OR-ing AB200_CELL_SIDE_INPUTS_READY into it promises that everything except `gates` (tokens /
 *   packed_argmax, cell_prev, the LayerNorm vectors, the embedding table) was complete before the
 *   kernel that PRECEDES this launch in the stream started; the kernel may then fetch them before
 *   its programmatic-dependent-launch wait (the draft loop sets it: the predecessor there is the
 *   gate GEMM).  Without the flag nothing is read before the wait.
 * D must be a multiple of 8 (16-byte vectors). */
#define AB200_CELL_SIDE_INPUTS_READY 0x100
AB200_API int32_t ab200_sum_lstm_cell(const void* gates, int64_t gates_row_stride,
                            const void* emb_table, const int64_t* tokens,
                            const uint64_t* packed_argmax,
                            const void* cell_prev,
                            const void* cell_ln_weight, const void* cell_ln_bias,
                            const void* state_ln_weight, const void* state_ln_bias,
                            void* state_out, void* cell_out,
                            int64_t B, int64_t D, int64_t vocab,
                            float alpha, float eps, int32_t dtype,
                            int32_t emulate_eager_rounding,
                            uint64_t* clear_packed, uint32_t* clear_counter,
                            void* state_q_out, float* state_scale_out, void* stream);

/* out[M, N] = x[M, K] @ w[N, K]^T  (nn.Linear without bias).
 * Replaces: self.projs[i](prev_state), arctic_speculator.py:555-561,668 (cuBLAS in the reference).
 * bf16: TMA + tcgen05.mma (fp32 accumulation in TMEM), output rounded once to bf16.
 * f32 : SIMT FFMA kernel (exact fp32, no TF32), the 1e-5 parity mode.
 * bf16 requirements: K % 8 == 0, x and w 16-byte aligned, ldx % 8 == 0, ldw % 8 == 0. */
AB200_API int32_t ab200_linear(const void* x, const void* w, void* out,
                     int64_t M, int64_t N, int64_t K,
                     int64_t ldx, int64_t ldw, int64_t ldo,
                     int32_t dtype, void* stream);

/* Fused vocabulary projection + greedy pick; logits are never written to memory (bf16).
 * Replaces: LogitsProcessorOpt._get_logits (logits_processor_opt.py:83-107) ->
 * UnquantizedEmbeddingMethod.apply (vocab_parallel_embedding.py:61-65) -> logits[..., :vocab] ->
 * torch.argmax (arctic_speculator.py:726-732).
 *   x [M, K], w [w_rows >= vocab, K] (vocab-padded head), only rows < vocab compete.
 *   packed_out [M] uint64, MUST be zero on entry; on exit
 *     packed = (order_preserving_bits(logit) << 32) | (0xFFFFFFFF - token)
 *   so that a 64-bit max picks the largest logit and, on ties, the LOWEST token id (torch.argmax).
 *   bf16: the logit is the fp32 accumulator rounded to bf16 (what the reference's GEMM returns).
 *   tokens_out (optional): tokens_out[r * tokens_stride] = token of row r as int64.  For bf16 and
 *     M <= 512 the last CTA to finish writes them (done_counter [1] must be zero on entry); in the
 *     other cases a separate unpack launch follows.
 *   Batch rows up to 512 share ONE pass over the weights (1, 2 or 4 accumulator tiles of 128 rows).
 *   f32 : needs workspace of M * vocab floats (logits are materialised), else AB200_E_WORKSPACE. */
AB200_API int32_t ab200_head_argmax(const void* x, const void* w,
                          int64_t M, int64_t w_rows, int64_t K, int64_t vocab,
                          int64_t ldx, int64_t ldw,
                          uint64_t* packed_out,
                          int64_t* tokens_out, int64_t tokens_stride, uint32_t* done_counter,
                          int32_t dtype,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * FP8 (e4m3) quantised head - SURVEY 8f "next" rank 1.
 * Replaces: the reference's `qhead` used when the padded batch is <= 32
 * (arctic_speculator.py:455-465, 726-727) = vLLM Fp8LinearMethod on an unquantised checkpoint
 * (arctic_inference/vllm/spec_dec/fp8.py:207-221 weight quantisation, :276-308 apply):
 *   weights     s_w = amax|W| / 448 (one scale),  q = sat_e4m3(w * (1/s_w))
 *   activations s_x[row] = max(amax|x_row| / 448, 1/(448*512)),  q = sat_e4m3(x / s_x)
 *   output      bf16( s_x[row] * (s_w * acc_fp32) )
 * The arithmetic is pinned against vLLM's own ops (scaled_fp8_quant, cutlass_scaled_mm) run on a
 * B200 (tests/test_fp8_gpu.py); vLLM 0.9.2 itself is absent, so parity with the reference's exact
 * binary is unpinned.
 */
AB200_API int32_t ab200_fp8_quantize_weight(const void* w /*bf16 [rows,K]*/, void* q_out /*e4m3 [rows,K]*/,
                                            float* scale_out /*[1]*/, int64_t rows, int64_t K,
                                            int64_t ldw, int64_t ldq, void* stream);
AB200_API int32_t ab200_fp8_quantize_rows(const void* x /*bf16 [M,K]*/, void* q_out /*e4m3 [M,K]*/,
                                          float* scales_out /*[M]*/, int64_t M, int64_t K,
                                          int64_t ldx, int64_t ldq, void* stream);
/* out[M,N] (bf16) = dequantised xq[M,K] @ wq[N,K]^T; tcgen05.mma kind::f8f6f4. K % 16 == 0. */
AB200_API int32_t ab200_linear_fp8(const void* xq, const float* x_scales, const void* wq,
                                   const float* w_scale, void* out, int64_t M, int64_t N, int64_t K,
                                   int64_t ldx, int64_t ldw, int64_t ldo, void* stream);
/* Fused FP8 head + arg-max; arguments as ab200_head_argmax. */
AB200_API int32_t ab200_head_argmax_fp8(const void* xq, const float* x_scales, const void* wq,
                                        const float* w_scale, int64_t M, int64_t w_rows, int64_t K,
                                        int64_t vocab, int64_t ldx, int64_t ldw, uint64_t* packed_out,
                                        int64_t* tokens_out, int64_t tokens_stride,
                                        uint32_t* done_counter, void* stream);

/* tokens_out[i * out_stride] = 0xFFFFFFFF - low32(packed[i])  for i < count (int64 output). */
AB200_API int32_t ab200_argmax_unpack(const uint64_t* packed, int64_t* tokens_out, int64_t count,
                            int64_t out_stride, void* stream);

/* Tensor-parallel head (vocab sharded over ranks, arctic_speculator.py:733-744): rewrites a shard-local
 * packed arg-max in place into a key with the GLOBAL token index (local + index_offset) whose signed
 * 64-bit order equals (logit, lowest index) order, so that ONE all_reduce(MAX) over int64 combines
 * the ranks; ab200_argmax_unpack and the cell kernels read such keys like ordinary packed values. */
AB200_API int32_t ab200_argmax_shard_key(uint64_t* packed, int64_t count, int64_t index_offset, void* stream);

/* Row-wise argmax over materialised logits [M, N] (lowest index on ties) into the packed format.
 * Used by the f32 path and by tests. packed_out need not be zeroed. */
AB200_API int32_t ab200_argmax_rows(const void* logits, int64_t M, int64_t N, int64_t ld, int32_t dtype,
                          uint64_t* packed_out, void* stream);

/* Test / reporting hooks (not on the product path).
 * ab200_linear_simt_bf16: bf16 inputs, fp32 FFMA accumulation, bf16 output - an independent
 *   comparator for the tcgen05 kernel.
 * ab200_gemm_block_n: tile width the tcgen05 GEMM picks for N output features on `sms` SMs
 *   (sms <= 0: the current device). */
AB200_API int32_t ab200_linear_simt_bf16(const void* x, const void* w, void* out,
                               int64_t M, int64_t N, int64_t K,
                               int64_t ldx, int64_t ldw, int64_t ldo, void* stream);
AB200_API int32_t ab200_gemm_block_n(int64_t N, int32_t sms);

/* ---- test hooks: exported ONLY by the debug flavour libarctic_b200_dbg.so (-DAB200_DEBUG), which
 * tests/ and tools/ load; the release library has no ab200_debug_* symbol ----------------------
 * ab200_debug_cell_generic(1): route bf16 + eager-rounding emulation through the generic
 *   (instruction-heavy) cell kernel instead of the bf16x2 kernel; returns the previous setting.
 *   The two must agree bit for bit (tests/test_cell_gpu.py).
 * ab200_debug_unary_bf16: op 0 = sigmoid, 1 = erf-GELU on n bf16 bit patterns; out_fast is the
 *   cell kernel's shortcut, out_exact ATen's formula (csrc/bf16_exact.cuh), both rounded to bf16;
 *   out_fast32 / out_exact32: the unguarded approximation and ATen's value in fp32.
 * ab200_debug_sigmoid_search: for each candidate (c1, k, addend) [n_candidates x 3 floats, device]
 *   counts the bf16 inputs x >= lo_bound where RN(rcp(fma(ex2(x*c1), k, addend))) differs from ATen's
 *   sigmoid (mismatches[c], zero-initialised by the caller) and the lowest such input (first_bad[c],
 *   initialised to 0xffffffff).  Used once, offline, to pick constants that need no rounding guard. */
#ifdef AB200_DEBUG
AB200_API int32_t ab200_debug_cell_generic(int32_t on);
AB200_API int32_t ab200_debug_sigmoid_search(const float* candidates, int32_t n_candidates,
                                             uint32_t* mismatches, uint32_t* first_bad, float lo_bound,
                                             void* stream);
AB200_API int32_t ab200_debug_unary_bf16(int32_t op, const uint16_t* in, uint16_t* out_fast,
                                         uint16_t* out_exact, float* out_fast32 /*optional*/,
                                         float* out_exact32 /*optional*/, int64_t n, void* stream);
/* test hook: 1 = generic MLP-cell kernel also for bf16 + emulation (must agree bit for bit) */
AB200_API int32_t ab200_debug_mlp_generic(int32_t on);
/* Launch-path tunables for parameter sweeps (tools/): "pdl", "umma_m64", "gemm_2cta", "rows_per_pass",
 * "late_trigger", "pre_stages", "l2_prefetch_kb", "gate_block_n", "head_block_n", "max_stages", ...
 * The release library compiles the defaults in and reads no environment variable. */
AB200_API int32_t ab200_debug_set_tunable(const char* name, int32_t value);
/* how many 4-CTA clusters of the pair GEMM the current device co-schedules (cudaOccupancyMaxActiveClusters) */
AB200_API int32_t ab200_debug_max_clusters_of_4(void);
#endif /* AB200_DEBUG */

/* Whole draft loop: ln0 -> S x (gate GEMM -> fused cell -> head GEMM + argmax) -> token unpack.
 * Replaces: ArcticLSTMSpeculator.generate_token_ids (arctic_speculator.py:706-751) as driven by
 * generate_proposals (arctic_speculator.py:753-866), TP = 1, unquantised head.
 * All launches go to `stream` (3*S + 1 kernels, chained with programmatic dependent launch),
 * graph-capturable. */
typedef struct ab200_lstm_speculator {
  int32_t dtype;            /* AB200_BF16 or AB200_F32 */
  int32_t emulate_eager_rounding;
  int64_t vocab;            /* org vocab size (logits[..., :vocab]) */
  int64_t head_rows;        /* rows of head weight (vocab padded to 64) */
  int64_t input_hidden_dim; /* H */
  int64_t inner_dim;        /* D = proj_dim = emb_dim = inner_dim */
  int32_t scale_input;      /* apply ln0(h)/sqrt(2) before step 0 (arctic_speculator.py:656) */
  float alpha;              /* emb_weight / state_weight (arctic_speculator.py:575-577,671) */
  float eps;                /* 1e-6 (arctic_speculator.py:78) */
  const void* forget_emb;   /* [vocab, D] */
  const void* proj0;        /* [4D, H]  projs.0.weight */
  const void* proj1;        /* [4D, D]  projs.1.weight */
  const void* cell_ln_w; const void* cell_ln_b;   /* [D] */
  const void* state_ln_w; const void* state_ln_b; /* [D] */
  const void* head;         /* [head_rows, D] head.0.weight (tied across steps) */
  /* optional FP8 head (reference: qhead, used iff the padded batch <= fp8_head_max_batch = 32) */
  const void* head_q;       /* [head_rows, D] e4m3, or NULL */
  const float* head_scale;  /* [1] */
  int32_t fp8_head_max_batch;
} ab200_lstm_speculator;

AB200_API size_t ab200_lstm_workspace_bytes(const ab200_lstm_speculator* spec, int64_t B, int32_t num_steps);

/* input_ids [B] int64, hidden [B, H] (row stride ld_hidden), tokens_out [B, num_steps] int64
 * row-major.  If states_out / cells_out are non-NULL they receive the per-step hidden and cell
 * states, [num_steps, B, D] (parity tests). */
AB200_API int32_t ab200_lstm_propose(const ab200_lstm_speculator* spec,
                           const int64_t* input_ids, const void* hidden, int64_t ld_hidden,
                           int64_t B, int32_t num_steps, int64_t* tokens_out,
                           void* states_out, void* cells_out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Drafter glue: ArcticProposer.prepare_hidden_states
 * (arctic_inference/vllm/spec_dec/arctic_proposer.py:113-147).  For request r:
 *   gen_len = #(sampled_token_ids[r, 0..max_gen_len) != -1)
 *   out[r, :] = sample_hidden[(gen_len - 1) + row_offsets[r], :]
 * row_offsets [num_requests] int64 (device) = exclusive cumsum of (num_draft_tokens + 1).
 * One launch instead of ne / sum / cumsum / sub / add / index. */
AB200_API int32_t ab200_select_hidden_rows(const void* sample_hidden, int64_t hidden_row_stride_bytes,
                                           int64_t num_hidden_rows, const int64_t* sampled_token_ids,
                                           int64_t sampled_row_stride, int32_t max_gen_len,
                                           const int64_t* row_offsets, int64_t num_requests, void* out,
                                           int64_t out_row_stride_bytes, int64_t row_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Arctic MLP speculator / "sum_rnn" step - SURVEY 8f "next" rank 3.
 * Replaces: ArcticMLPSpeculator.generate_states (arctic_speculator.py:265-285) and the sum_rnn
 * branch of ArcticLSTMSpeculator.generate_states (:693-704):
 *   state = GELU( ln( pre + alpha * emb[token] ) ),  pre = proj(x) from ab200_linear,
 *   ln = MLPSpeculatorLayerNorm with scale and shift (:88-95).
 * Arguments as ab200_sum_lstm_cell (pre [B, D] instead of the four gates, no cell state).
 */
AB200_API int32_t ab200_sum_mlp_cell(const void* pre, int64_t pre_row_stride, const void* emb_table,
                                     const int64_t* tokens, const uint64_t* packed_argmax,
                                     const void* ln_weight, const void* ln_bias, void* state_out,
                                     int64_t B, int64_t D, int64_t vocab, float alpha, float eps,
                                     int32_t dtype, int32_t emulate_eager_rounding,
                                     uint64_t* clear_packed, uint32_t* clear_counter,
                                     void* state_q_out, float* state_scale_out, void* stream);
/* Whole draft loop of the MLP speculator: ln0 -> S x (proj GEMM -> fused cell -> head GEMM +
 * arg-max).  Replaces ArcticMLPSpeculator.generate_token_ids (arctic_speculator.py:287-326) as
 * driven by generate_proposals (:328-389), TP = 1.  Per-step weight pointers: tied checkpoints
 * repeat the same pointer (emb/head/ln: one set; proj: first + tied), untied ones list every step. */
#define AB200_MLP_MAX_STAGES 16
typedef struct ab200_mlp_speculator {
  int32_t dtype;            /* AB200_BF16 or AB200_F32 */
  int32_t emulate_eager_rounding;
  int64_t vocab;
  int64_t head_rows;        /* rows of the head weights (vocab padded to 64) */
  int64_t input_hidden_dim; /* emb_dim of the reference config: width of the base model's state */
  int64_t inner_dim;        /* D */
  int32_t scale_input;
  float alpha;              /* emb_weight / state_weight (:213-215, :281) */
  float eps;
  const void* emb[AB200_MLP_MAX_STAGES];    /* [vocab, D] */
  const void* proj[AB200_MLP_MAX_STAGES];   /* [D, H] for step 0, [D, D] after */
  const void* ln_w[AB200_MLP_MAX_STAGES];   /* [D] */
  const void* ln_b[AB200_MLP_MAX_STAGES];   /* [D] */
  const void* head[AB200_MLP_MAX_STAGES];   /* [head_rows, D] */
  const void* head_q[AB200_MLP_MAX_STAGES];       /* optional e4m3 copy of head (qhead) */
  const float* head_scale[AB200_MLP_MAX_STAGES];  /* [1] each */
  int32_t fp8_head_max_batch;
} ab200_mlp_speculator;
AB200_API size_t ab200_mlp_workspace_bytes(const ab200_mlp_speculator* spec, int64_t B, int32_t num_steps);
/* states_out (optional): [num_steps, B, D] */
AB200_API int32_t ab200_mlp_propose(const ab200_mlp_speculator* spec, const int64_t* input_ids,
                                    const void* hidden, int64_t ld_hidden, int64_t B, int32_t num_steps,
                                    int64_t* tokens_out, void* states_out, void* workspace,
                                    size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARCTIC_B200_H_ */
