// speculator.cu - library plumbing (error string, launch counter) and the whole-loop entry point
// ab200_lstm_propose: ln0 -> S x (gate GEMM -> fused cell -> head GEMM + arg-max) -> token unpack.
// Replaces ArcticLSTMSpeculator.generate_token_ids (arctic_speculator.py:706-751): ~35 ATen
// launches per step in the reference become 3 (plus ln0 once per call), chained with programmatic
// dependent launch and enqueued on one stream so the caller can capture the loop into a single
// CUDA graph.  The arg-max slots are reset by the cell kernel and the int64 tokens are written by
// the last CTA of the head kernel, so there is no memset node and no unpack launch (bf16, B<=512).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace ab200 {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
// trace state: per host thread (a profiling thread traces its own launches, nobody else's)
static thread_local unsigned long long* g_trace_base = nullptr;
static thread_local int g_trace_cap = 0, g_trace_next = 0;
unsigned long long* next_trace_slot() {
  if (g_trace_base == nullptr || g_trace_next >= g_trace_cap) return nullptr;
  return g_trace_base + kTraceWords * static_cast<size_t>(g_trace_next++);
}
#ifdef AB200_DEBUG
Tune& tune_mut() {
  static Tune t;
  return t;
}
#endif
int device_sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    (void)cudaGetLastError();
    return 148;
  }
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    (void)cudaGetLastError();
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// tokens_out[b * S + s] = token of packed[s * B + b]
__global__ void unpack_steps_kernel(const unsigned long long* __restrict__ packed,
                                    int64_t* __restrict__ tokens_out, int64_t B, int32_t S) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < B * S) {
    const int64_t s = i / B, b = i - s * B;
    tokens_out[b * S + s] = static_cast<int64_t>(unpack_idx(packed[i]));
  }
}

// ArcticProposer.prepare_hidden_states (arctic_proposer.py:113-147): request r keeps the hidden
// state of its LAST accepted token: row (gen_len[r] - 1) + row_offsets[r] of sample_hidden_states,
// gen_len[r] = #(sampled_token_ids[r, :] != -1), row_offsets = exclusive cumsum of (num_draft + 1).
// One CTA per request; replaces ne / sum / cumsum / sub / add / index (8 ATen launches).
__global__ void __launch_bounds__(256)
select_hidden_rows_kernel(const char* __restrict__ hidden, int64_t row_stride_bytes, int64_t num_rows,
                          const int64_t* __restrict__ sampled, int64_t ld_sampled, int32_t max_gen_len,
                          const int64_t* __restrict__ row_offsets, char* __restrict__ out,
                          int64_t out_stride_bytes, uint32_t row_bytes, int32_t vec16) {
  const int64_t r = blockIdx.x;
  int gen = 0;
  for (int g = 0; g < max_gen_len; ++g) gen += __ldg(sampled + r * ld_sampled + g) != -1;
  int64_t src = (gen - 1) + __ldg(row_offsets + r);
  // gen == 0 indexes row offset-1 in the reference (negative wraps in torch); clamp into the tensor
  if (src < 0) src += num_rows;
  // an index outside the tensor (the reference's fancy indexing would raise) is never dereferenced: the output
  // row is zero-filled, so the result is deterministic
  const bool ok = src >= 0 && src < num_rows;
  const char* s = hidden + (ok ? src : 0) * row_stride_bytes;
  char* d = out + r * out_stride_bytes;
  if (vec16) {
    for (uint32_t q = threadIdx.x; q < row_bytes / 16; q += blockDim.x)
      st_stream_16(d + q * 16ull, ok ? ld_stream_16(s + q * 16ull) : make_uint4(0u, 0u, 0u, 0u));
    for (uint32_t b = (row_bytes / 16) * 16 + threadIdx.x; b < row_bytes; b += blockDim.x) d[b] = ok ? s[b] : 0;
  } else {
    for (uint32_t b = threadIdx.x; b < row_bytes; b += blockDim.x) d[b] = ok ? s[b] : 0;
  }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  size_t x0, gates, state[2], cell, packed, counters, state_q, state_scale, logits, total;
};

static Workspace plan(const ab200_lstm_speculator* sp, int64_t B, int32_t S) {
  const size_t e = dtype_bytes(sp->dtype);
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  w.x0 = take(static_cast<size_t>(B) * sp->input_hidden_dim * e);
  w.gates = take(static_cast<size_t>(B) * 4 * sp->inner_dim * e);
  w.state[0] = take(static_cast<size_t>(B) * sp->inner_dim * e);
  w.state[1] = take(static_cast<size_t>(B) * sp->inner_dim * e);
  w.cell = take(static_cast<size_t>(B) * sp->inner_dim * e);
  w.packed = take(static_cast<size_t>(B) * S * sizeof(uint64_t));
  w.counters = take(static_cast<size_t>(S) * sizeof(uint32_t));
  w.state_q = take(static_cast<size_t>(B) * sp->inner_dim);
  w.state_scale = take(static_cast<size_t>(B) * sizeof(float));
  w.logits = take(sp->dtype == AB200_F32 ? static_cast<size_t>(B) * sp->vocab * sizeof(float) : 0);
  w.total = off;
  return w;
}

}  // namespace ab200

using namespace ab200;

extern "C" const char* ab200_last_error(void) { return g_err; }
extern "C" int32_t ab200_abi_version(void) { return 5; }
extern "C" int32_t ab200_init(void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  (void)device_sm_count();
  int rc = init_gemm_kernels();
  if (rc != AB200_OK) return rc;
  rc = init_lstm_tables(s);
  if (rc != AB200_OK) return rc;
  return init_mlp_tables(s);
}
#ifdef AB200_DEBUG
extern "C" int32_t ab200_debug_set_tunable(const char* name, int32_t value) {
  AB200_REQUIRE(name != nullptr, AB200_E_INVALID, "debug_set_tunable: NULL name");
  Tune& t = tune_mut();
  struct { const char* n; int* p; } tab[] = {
      {"pdl", &t.pdl}, {"umma_m64", &t.umma_m64}, {"gemm_2cta", &t.gemm_2cta},
      {"rows_per_pass", &t.rows_per_pass}, {"late_trigger", &t.late_trigger},
      {"pre_stages", &t.pre_stages}, {"l2_prefetch_kb", &t.l2_prefetch_kb}, {"gate_l2_prefetch", &t.gate_l2_prefetch}, {"gemm_4cta", &t.gemm_4cta}, {"dbg_skip", &t.dbg_skip},
      {"gate_block_n", &t.gate_block_n}, {"head_block_n", &t.head_block_n},
      {"max_stages", &t.max_stages}, {"gate_mh2", &t.gate_mh2}, {"gate_kbs", &t.gate_kbs}, {"head_kbs", &t.head_kbs}, {"pair_gate_kbs", &t.pair_gate_kbs}, {"pair_head_kbs", &t.pair_head_kbs}};
  for (auto& e : tab)
    if (strcmp(e.n, name) == 0) {
      *e.p = value;
      return AB200_OK;
    }
  AB200_REQUIRE(false, AB200_E_INVALID, "debug_set_tunable: unknown tunable '%s'", name);
}
#endif
extern "C" int64_t ab200_launch_count(void) { return g_launches; }
extern "C" int32_t ab200_set_trace_buffer(uint64_t* device_slots, int32_t max_launches) {
  g_trace_base = reinterpret_cast<unsigned long long*>(device_slots);
  g_trace_cap = device_slots ? max_launches : 0;
  g_trace_next = 0;
  return AB200_OK;
}

extern "C" size_t ab200_lstm_workspace_bytes(const ab200_lstm_speculator* spec, int64_t B,
                                             int32_t num_steps) {
  if (!spec || B <= 0 || num_steps <= 0) return 0;
  return plan(spec, B, num_steps).total;
}

extern "C" int32_t ab200_lstm_propose(const ab200_lstm_speculator* sp, const int64_t* input_ids,
                                      const void* hidden, int64_t ld_hidden, int64_t B,
                                      int32_t num_steps, int64_t* tokens_out, void* states_out,
                                      void* cells_out, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  AB200_REQUIRE(sp != nullptr, AB200_E_INVALID, "lstm_propose: NULL speculator descriptor");
  if (B == 0 || num_steps == 0) return AB200_OK;
  AB200_REQUIRE(B > 0 && num_steps > 0 && input_ids && hidden && tokens_out, AB200_E_INVALID,
                "lstm_propose: bad arguments");
  AB200_REQUIRE(sp->dtype == AB200_BF16 || sp->dtype == AB200_F32, AB200_E_INVALID,
                "lstm_propose: dtype must be bf16 or f32");
  AB200_REQUIRE(sp->forget_emb && sp->proj0 && sp->proj1 && sp->cell_ln_w && sp->cell_ln_b &&
                    sp->state_ln_w && sp->state_ln_b && sp->head,
                AB200_E_INVALID, "lstm_propose: NULL weight pointer");
  AB200_REQUIRE(sp->vocab > 0 && sp->head_rows >= sp->vocab && sp->inner_dim > 0 &&
                    sp->input_hidden_dim > 0,
                AB200_E_INVALID, "lstm_propose: bad model dimensions");
  const Workspace w = plan(sp, B, num_steps);
  AB200_REQUIRE(workspace && workspace_bytes >= w.total && aligned(workspace, 256),
                AB200_E_WORKSPACE, "lstm_propose: workspace of %zu bytes (256-byte aligned) required",
                w.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const size_t e = dtype_bytes(sp->dtype);
  const int64_t D = sp->inner_dim, H = sp->input_hidden_dim;
  const size_t bd = static_cast<size_t>(B) * D * e;
  uint64_t* packed = reinterpret_cast<uint64_t*>(ws + w.packed);

  uint32_t* counters = reinterpret_cast<uint32_t*>(ws + w.counters);
  // bf16, B <= rows per GEMM pass (256): the cell kernel resets the arg-max slot / counter of its step and the head
  // kernel's last CTA writes the int64 tokens -> no memset node, no unpack launch.
  const bool fused_tokens = sp->dtype == AB200_BF16 && B <= gemm_rows_per_pass();
  // reference selection rule (arctic_speculator.py:726-727): quantised head iff batch <= 32
  const bool fp8_head = sp->dtype == AB200_BF16 && sp->head_q != nullptr && sp->head_scale != nullptr &&
                        B <= sp->fp8_head_max_batch && D % 16 == 0;
  void* state_q = fp8_head ? ws + w.state_q : nullptr;
  float* state_scale = fp8_head ? reinterpret_cast<float*>(ws + w.state_scale) : nullptr;
  if (!fused_tokens)
    AB200_CUDA_OK(cudaMemsetAsync(packed, 0, static_cast<size_t>(B) * num_steps * sizeof(uint64_t), s));

  const void* xin = hidden;
  int64_t ldx = ld_hidden;
  if (sp->scale_input) {
    // ln0(h) / sqrt(2): ATen CUDA multiplies by the fp32 reciprocal of the scalar divisor
    const float inv_sqrt2 = 1.0f / 1.41421356237309504880f;
    int rc = ab200_rms_scale(hidden, ws + w.x0, B, H, ld_hidden, H, sp->eps, inv_sqrt2, sp->dtype,
                             sp->emulate_eager_rounding, stream);
    if (rc != AB200_OK) return rc;
    xin = ws + w.x0;
    ldx = H;
  }
  const void* cell_prev = nullptr;
  for (int32_t step = 0; step < num_steps; ++step) {
    const void* W = step == 0 ? sp->proj0 : sp->proj1;   // tie_weights: projs[min(step, 1)]
    const int64_t K = step == 0 ? H : D;
    int rc = ab200_linear(xin, W, ws + w.gates, B, 4 * D, K, ldx, K, 4 * D, sp->dtype, stream);
    if (rc != AB200_OK) return rc;
    void* state = states_out ? static_cast<char*>(states_out) + step * bd : ws + w.state[step & 1];
    void* cell = cells_out ? static_cast<char*>(cells_out) + step * bd : ws + w.cell;
    rc = ab200_sum_lstm_cell(ws + w.gates, 4 * D, sp->forget_emb, step == 0 ? input_ids : nullptr,
                             step == 0 ? nullptr : packed + static_cast<size_t>(step - 1) * B,
                             cell_prev, sp->cell_ln_w, sp->cell_ln_b, sp->state_ln_w, sp->state_ln_b,
                             state, cell, B, D, sp->vocab, sp->alpha, sp->eps, sp->dtype,
                             // the gate GEMM just launched signals its dependents only after it has
                             // itself waited for everything before it: tokens, cell state and
                             // weights are complete by the time this kernel can start
                             sp->emulate_eager_rounding | (gemm_late_trigger() ? AB200_CELL_SIDE_INPUTS_READY : 0),
                             fused_tokens ? packed + static_cast<size_t>(step) * B : nullptr,
                             fused_tokens ? counters + step : nullptr, state_q, state_scale, stream);
    if (rc != AB200_OK) return rc;
    if (fp8_head)
      rc = ab200_head_argmax_fp8(state_q, state_scale, sp->head_q, sp->head_scale, B, sp->head_rows, D,
                                 sp->vocab, D, D, packed + static_cast<size_t>(step) * B,
                                 fused_tokens ? tokens_out + step : nullptr, num_steps,
                                 fused_tokens ? counters + step : nullptr, stream);
    else
      rc = ab200_head_argmax(state, sp->head, B, sp->head_rows, D, sp->vocab, D, D,
                             packed + static_cast<size_t>(step) * B,
                             fused_tokens ? tokens_out + step : nullptr, num_steps,
                             fused_tokens ? counters + step : nullptr, sp->dtype, ws + w.logits,
                             w.total - w.logits, stream);
    if (rc != AB200_OK) return rc;
    xin = state;
    ldx = D;
    cell_prev = cell;
  }
  if (!fused_tokens) {
    const int64_t n = B * num_steps;
    unpack_steps_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const unsigned long long*>(packed), tokens_out, B, num_steps);
    AB200_CHECK_LAUNCH("unpack_steps_kernel");
  }
  return AB200_OK;
}

extern "C" int32_t ab200_select_hidden_rows(const void* sample_hidden, int64_t hidden_row_stride_bytes,
                                            int64_t num_hidden_rows, const int64_t* sampled_token_ids,
                                            int64_t sampled_row_stride, int32_t max_gen_len,
                                            const int64_t* row_offsets, int64_t num_requests, void* out,
                                            int64_t out_row_stride_bytes, int64_t row_bytes, void* stream) {
  if (num_requests == 0) return AB200_OK;
  AB200_REQUIRE(sample_hidden && sampled_token_ids && row_offsets && out && num_requests > 0 &&
                    num_hidden_rows > 0 && max_gen_len > 0 && row_bytes > 0 && row_bytes < (1LL << 31),
                AB200_E_INVALID, "select_hidden_rows: bad arguments");
  const int vec = aligned(sample_hidden, 16) && aligned(out, 16) && hidden_row_stride_bytes % 16 == 0 &&
                  out_row_stride_bytes % 16 == 0;
  select_hidden_rows_kernel<<<static_cast<unsigned>(num_requests), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const char*>(sample_hidden), hidden_row_stride_bytes, num_hidden_rows, sampled_token_ids,
      sampled_row_stride, max_gen_len, row_offsets, static_cast<char*>(out), out_row_stride_bytes,
      static_cast<uint32_t>(row_bytes), vec);
  AB200_CHECK_LAUNCH("select_hidden_rows_kernel");
  return AB200_OK;
}
