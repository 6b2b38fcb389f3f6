// mlp_cell.cu - the Arctic MLP speculator ("sum_rnn") step for sm_100a:
//   state = GELU( ln( proj(x) + alpha * emb[tok] ) )
// Replaces ArcticMLPSpeculator.generate_states (arctic_speculator.py:265-285) and the sum_rnn
// branch of ArcticLSTMSpeculator.generate_states (arctic_speculator.py:693-704): F.embedding,
// add_(alpha), the 7 ATen kernels of MLPSpeculatorLayerNorm (:88-95) and GELU become one launch;
// the projection and the vocab head reuse the tcgen05 GEMMs of gemm_sm100.cu.
// Same structure and the same exact-rounding shortcuts as the LSTM cell (lstm_cell.cu,
// bf16_exact.cuh): one CTA per row, values in registers across the row reduction.
#include <type_traits>

#include "bf16_exact.cuh"
#include "common.cuh"

namespace ab200 {

__device__ float g_gelu_lut_mlp[kGeluLutSize];
__global__ void gelu_lut_init_mlp_kernel() { g_gelu_lut_mlp[threadIdx.x] = gelu_lut_entry(threadIdx.x); }

template <typename T, bool EMU>
__device__ __forceinline__ float RR(float v) {
  if constexpr (EMU) return ST<T>::round(v);
  else return v;
}

// generic storage type / rounding mode (fp32 parity mode, f16, bf16 without emulation)
template <typename T, bool EMU, int ITEMS>
__global__ void __launch_bounds__(1024)
sum_mlp_cell_kernel(const T* __restrict__ pre, int64_t ldp, const T* __restrict__ emb,
                    const int64_t* __restrict__ tokens, const unsigned long long* __restrict__ packed,
                    const T* __restrict__ lw, const T* __restrict__ lb, T* __restrict__ state_out,
                    int32_t D, int64_t vocab, float alpha, float eps, float inv_d,
                    unsigned long long* __restrict__ clear_packed, unsigned int* __restrict__ clear_counter,
                    unsigned long long* trace, uint8_t* __restrict__ state_q,
                    float* __restrict__ state_scale) {
  constexpr int VN = Vec16<T>::N;
  __shared__ float scratch[33];
  const int64_t row = blockIdx.x;
  const int nvec = D / VN;
  if (threadIdx.x == 0) trace_stamp(trace, 0);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    trace_stamp(trace, 1);
    if (clear_packed) clear_packed[row] = 0ull;
    if (clear_counter && row == 0) *clear_counter = 0u;
  }
  int64_t tok = packed ? static_cast<int64_t>(unpack_idx(packed[row])) : tokens[row];
  if (static_cast<uint64_t>(tok) >= static_cast<uint64_t>(vocab)) tok = tok < 0 ? 0 : vocab - 1;
  const T* g = pre + row * ldp;
  const T* z = emb + tok * static_cast<int64_t>(D);
  float a[ITEMS][VN];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      Vec16<T> gv, zv;
      gv.u = ld_stream_16(g + vi * VN);
      zv.u = *reinterpret_cast<const uint4*>(z + vi * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        // states.add_(z, alpha): one rounding of fma(alpha, z, s)
        a[it][j] = RR<T, EMU>(fmaf(alpha, ST<T>::to_f(zv.e[j]), ST<T>::to_f(gv.e[j])));
        ss += RR<T, EMU>(a[it][j] * a[it][j]);
      }
    }
  }
  ss = block_sum(ss, scratch);
  const float mean = RR<T, EMU>(ss * inv_d);
  const float inv = RR<T, EMU>(rsqrtf(RR<T, EMU>(mean + eps)));
  float amax = 0.f;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      Vec16<T> w, b, so;
      w.u = *reinterpret_cast<const uint4*>(lw + vi * VN);
      b.u = *reinterpret_cast<const uint4*>(lb + vi * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float y = RR<T, EMU>(a[it][j] * inv);
        y = RR<T, EMU>(ST<T>::to_f(w.e[j]) * y);
        y = RR<T, EMU>(y + ST<T>::to_f(b.e[j]));
        y = y * 0.5f * (1.0f + erff(y * 0.70710678118654752440f));
        so.e[j] = ST<T>::from_f(y);
        a[it][j] = ST<T>::to_f(so.e[j]);
        amax = fmaxf(amax, fabsf(a[it][j]));
      }
      *reinterpret_cast<uint4*>(state_out + row * D + vi * VN) = so.u;
    }
  }
  if (state_q != nullptr) {
    amax = block_max(amax, scratch);
    const float scale = fmaxf(amax / kFp8Max, kFp8MinScale);
    if (threadIdx.x == 0) state_scale[row] = scale;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        uint8_t qv[VN];
#pragma unroll
        for (int j = 0; j < VN; ++j) qv[j] = to_e4m3(a[it][j] / scale);
        if constexpr (VN == 8) {
          *reinterpret_cast<uint2*>(state_q + row * D + vi * VN) = *reinterpret_cast<uint2*>(qv);
        } else {
          *reinterpret_cast<uint32_t*>(state_q + row * D + vi * VN) = *reinterpret_cast<uint32_t*>(qv);
        }
      }
    }
  }
  if (threadIdx.x == 0) trace_stamp(trace, 2);
}

// bf16 + eager-rounding emulation in packed arithmetic (bit-identical to the generic kernel)
template <int ITEMS, bool GUARD>
__global__ void __launch_bounds__(1024)
sum_mlp_cell_bf16x2_kernel(const __nv_bfloat16* __restrict__ pre, int64_t ldp,
                           const __nv_bfloat16* __restrict__ emb, const int64_t* __restrict__ tokens,
                           const unsigned long long* __restrict__ packed,
                           const __nv_bfloat16* __restrict__ lw, const __nv_bfloat16* __restrict__ lb,
                           __nv_bfloat16* __restrict__ state_out, int32_t D, int64_t vocab, float alpha,
                           float eps, float inv_d, unsigned long long* __restrict__ clear_packed,
                           unsigned int* __restrict__ clear_counter, unsigned long long* trace,
                           uint8_t* __restrict__ state_q, float* __restrict__ state_scale) {
  using T = __nv_bfloat16;
  __shared__ float scratch[33];
  __shared__ float gelu_lut[kGeluLutSize];
  const int64_t row = blockIdx.x;
  const int nvec = D / 8;
  if (threadIdx.x == 0) trace_stamp(trace, 0);
  pdl_launch_dependents();
  if constexpr (GUARD) gelu_lut_fill(gelu_lut, g_gelu_lut_mlp);
  uint32_t lut_s = static_cast<uint32_t>(__cvta_generic_to_shared(gelu_lut));
  asm volatile("mov.u32 %0, %0;" : "+r"(lut_s));
  pdl_wait();
  if (threadIdx.x == 0) {
    trace_stamp(trace, 1);
    if (clear_packed) clear_packed[row] = 0ull;
    if (clear_counter && row == 0) *clear_counter = 0u;
  }
  int64_t tok = packed ? static_cast<int64_t>(unpack_idx(packed[row])) : tokens[row];
  if (static_cast<uint64_t>(tok) >= static_cast<uint64_t>(vocab)) tok = tok < 0 ? 0 : vocab - 1;
  const T* g = pre + row * ldp;
  const T* z = emb + tok * static_cast<int64_t>(D);
  uint32_t A[ITEMS][4];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      const uint4 gv = ld_stream_16(g + vi * 8);
      const uint4 zv = *reinterpret_cast<const uint4*>(z + vi * 8);
      const uint32_t ga[4] = {gv.x, gv.y, gv.z, gv.w}, za[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const f2 s = fma2(bcast2(alpha), unpack2(za[p]), unpack2(ga[p]));
        A[it][p] = pack_rn(s.a, s.b);
        const uint32_t sq = mul2(A[it][p], A[it][p]);
        ss += bf_lo(sq);
        ss += bf_hi(sq);
      }
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) A[it][p] = 0u;
    }
  }
  ss = block_sum(ss, scratch);   // also orders the table fill before its first read
  const float mean = ST<T>::round(ss * inv_d);
  const float inv = ST<T>::round(rsqrtf(ST<T>::round(mean + eps)));
  const uint32_t inv2 = pack_rn(inv, inv);
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      const uint4 w = *reinterpret_cast<const uint4*>(lw + vi * 8);
      const uint4 b = *reinterpret_cast<const uint4*>(lb + vi * 8);
      const uint32_t wa[4] = {w.x, w.y, w.z, w.w}, ba[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        uint32_t y = mul2(A[it][p], inv2);
        y = mul2(wa[p], y);
        y = add2(y, ba[p]);
        A[it][p] = gelu2_bf16_exact<GUARD>(y, lut_s);
      }
      *reinterpret_cast<uint4*>(state_out + row * D + vi * 8) =
          make_uint4(A[it][0], A[it][1], A[it][2], A[it][3]);
    }
  }
  if (state_q != nullptr) {
    float amax = 0.f;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
#pragma unroll
      for (int p = 0; p < 4; ++p)
        amax = fmaxf(amax, fmaxf(fabsf(bf_lo(A[it][p])), fabsf(bf_hi(A[it][p]))));
    amax = block_max(amax, scratch);
    const float scale = fmaxf(amax / kFp8Max, kFp8MinScale);
    if (threadIdx.x == 0) state_scale[row] = scale;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        uint32_t q[2] = {0u, 0u};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const uint32_t lo = to_e4m3(bf_lo(A[it][p]) / scale), hi = to_e4m3(bf_hi(A[it][p]) / scale);
          q[p >> 1] |= (lo | (hi << 8)) << (16 * (p & 1));
        }
        *reinterpret_cast<uint2*>(state_q + row * D + vi * 8) = make_uint2(q[0], q[1]);
      }
    }
  }
  if (threadIdx.x == 0) trace_stamp(trace, 2);
}

#ifdef AB200_DEBUG
static std::atomic<int> g_mlp_generic{0};
#else
static constexpr int g_mlp_generic = 0;
#endif
static DeviceOnce g_gelu_lut_mlp_once;
int init_mlp_tables(cudaStream_t s) {
  return device_once(g_gelu_lut_mlp_once, s, "GELU table (MLP)", [](cudaStream_t st) {
    gelu_lut_init_mlp_kernel<<<1, kGeluLutSize, 0, st>>>();
    AB200_CHECK_LAUNCH("gelu_lut_init_mlp_kernel");
    return AB200_OK;
  });
}

static bool mlp_row_shape(int64_t nvec, int* threads, int* items) {
  for (int it : {1, 2, 4}) {
    int64_t t = (nvec + it - 1) / it;
    t = (t + 31) / 32 * 32;
    if (t <= 1024) {
      *threads = static_cast<int>(t < 32 ? 32 : t);
      *items = it;
      return true;
    }
  }
  return false;
}

template <typename T, bool EMU>
static int launch_mlp(const void* pre, int64_t ldp, const void* emb, const int64_t* tokens,
                      const uint64_t* packed, const void* lw, const void* lb, void* state_out,
                      int64_t B, int64_t D, int64_t vocab, float alpha, float eps,
                      uint64_t* clear_packed, uint32_t* clear_counter, void* state_q,
                      float* state_scale, bool guard, cudaStream_t s) {
  int threads, items;
  AB200_REQUIRE(mlp_row_shape(D / Vec16<T>::N, &threads, &items), AB200_E_INVALID,
                "sum_mlp_cell: D %lld too large", (long long)D);
  LaunchCfg lc(dim3(static_cast<unsigned>(B)), dim3(threads), 0, s);
  const T* a_pre = static_cast<const T*>(pre);
  const T* a_emb = static_cast<const T*>(emb);
  const unsigned long long* a_packed = reinterpret_cast<const unsigned long long*>(packed);
  const T* a_lw = static_cast<const T*>(lw);
  const T* a_lb = static_cast<const T*>(lb);
  T* a_state = static_cast<T*>(state_out);
  const int32_t a_D = static_cast<int32_t>(D);
  const float inv_d = 1.0f / static_cast<float>(D);
  unsigned long long* a_clear = reinterpret_cast<unsigned long long*>(clear_packed);
  unsigned int* a_cnt = reinterpret_cast<unsigned int*>(clear_counter);
  unsigned long long* a_trace = next_trace_slot();
  uint8_t* a_sq = static_cast<uint8_t*>(state_q);
#define AB200_MLP_ARGS                                                                         \
  a_pre, ldp, a_emb, tokens, a_packed, a_lw, a_lb, a_state, a_D, vocab, alpha, eps, inv_d,     \
      a_clear, a_cnt, a_trace, a_sq, state_scale
  if constexpr (EMU && std::is_same<T, __nv_bfloat16>::value) {
    if (!g_mlp_generic) {
      int rc = init_mlp_tables(s);   // first use on this device: fill the GELU table (synchronous, once)
      if (rc != AB200_OK) return rc;
#define AB200_MLP_FAST(IT)                                                                     \
  do {                                                                                         \
    if (guard)                                                                                 \
      AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_mlp_cell_bf16x2_kernel<IT, true>, AB200_MLP_ARGS)); \
    else                                                                                       \
      AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_mlp_cell_bf16x2_kernel<IT, false>, AB200_MLP_ARGS)); \
  } while (0)
      if (items == 1) AB200_MLP_FAST(1);
      else if (items == 2) AB200_MLP_FAST(2);
      else AB200_MLP_FAST(4);
#undef AB200_MLP_FAST
      AB200_CHECK_LAUNCH("sum_mlp_cell_bf16x2_kernel");
      return AB200_OK;
    }
  }
  switch (items) {
    case 1: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_mlp_cell_kernel<T, EMU, 1>, AB200_MLP_ARGS)); break;
    case 2: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_mlp_cell_kernel<T, EMU, 2>, AB200_MLP_ARGS)); break;
    default: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_mlp_cell_kernel<T, EMU, 4>, AB200_MLP_ARGS)); break;
  }
#undef AB200_MLP_ARGS
  AB200_CHECK_LAUNCH("sum_mlp_cell_kernel");
  return AB200_OK;
}

struct MlpWorkspace {
  size_t x0, pre, state[2], packed, counters, state_q, state_scale, logits, total;
};
static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
static MlpWorkspace mlp_plan(const ab200_mlp_speculator* sp, int64_t B, int32_t S) {
  const size_t e = dtype_bytes(sp->dtype);
  MlpWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align256(off + bytes);
    return o;
  };
  w.x0 = take(static_cast<size_t>(B) * sp->input_hidden_dim * e);
  w.pre = take(static_cast<size_t>(B) * sp->inner_dim * e);
  w.state[0] = take(static_cast<size_t>(B) * sp->inner_dim * e);
  w.state[1] = take(static_cast<size_t>(B) * sp->inner_dim * e);
  w.packed = take(static_cast<size_t>(B) * S * sizeof(uint64_t));
  w.counters = take(static_cast<size_t>(S) * sizeof(uint32_t));
  w.state_q = take(static_cast<size_t>(B) * sp->inner_dim);
  w.state_scale = take(static_cast<size_t>(B) * sizeof(float));
  w.logits = take(sp->dtype == AB200_F32 ? static_cast<size_t>(B) * sp->vocab * sizeof(float) : 0);
  w.total = off;
  return w;
}

__global__ void mlp_unpack_steps_kernel(const unsigned long long* __restrict__ packed,
                                        int64_t* __restrict__ tokens_out, int64_t B, int32_t S) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < B * S) {
    const int64_t s = i / B, b = i - s * B;
    tokens_out[b * S + s] = static_cast<int64_t>(unpack_idx(packed[i]));
  }
}

}  // namespace ab200

using namespace ab200;

#ifdef AB200_DEBUG
extern "C" int32_t ab200_debug_mlp_generic(int32_t on) { return g_mlp_generic.exchange(on != 0); }
#endif

extern "C" int32_t ab200_sum_mlp_cell(const void* pre, int64_t pre_row_stride, const void* emb_table,
                                      const int64_t* tokens, const uint64_t* packed_argmax,
                                      const void* ln_weight, const void* ln_bias, void* state_out,
                                      int64_t B, int64_t D, int64_t vocab, float alpha, float eps,
                                      int32_t dtype, int32_t emulate_eager_rounding,
                                      uint64_t* clear_packed, uint32_t* clear_counter,
                                      void* state_q_out, float* state_scale_out, void* stream) {
  if (B == 0) return AB200_OK;
  AB200_REQUIRE(pre && emb_table && (tokens || packed_argmax) && ln_weight && ln_bias && state_out,
                AB200_E_INVALID, "sum_mlp_cell: NULL pointer argument");
  AB200_REQUIRE(B > 0 && D > 0 && vocab > 0, AB200_E_INVALID, "sum_mlp_cell: bad sizes");
  AB200_REQUIRE(state_q_out == nullptr ||
                    (state_scale_out != nullptr && dtype == AB200_BF16 && aligned(state_q_out, 8)),
                AB200_E_INVALID, "sum_mlp_cell: fp8 state output needs bf16 states and a scale buffer");
  const int vn = 16 / dtype_bytes(dtype);
  AB200_REQUIRE(D % vn == 0 && pre_row_stride % vn == 0, AB200_E_INVALID,
                "sum_mlp_cell: D and the row stride must be multiples of %d", vn);
  AB200_REQUIRE(aligned(pre, 16) && aligned(emb_table, 16) && aligned(ln_weight, 16) &&
                    aligned(ln_bias, 16) && aligned(state_out, 16),
                AB200_E_INVALID, "sum_mlp_cell: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool emu = emulate_eager_rounding != 0;
#define AB200_ARGS                                                                             \
  pre, pre_row_stride, emb_table, tokens, packed_argmax, ln_weight, ln_bias, state_out, B, D,  \
      vocab, alpha, eps, clear_packed, clear_counter, state_q_out, state_scale_out,            \
      emulate_eager_rounding != 2, s
  switch (dtype) {
    case AB200_F32: return launch_mlp<float, false>(AB200_ARGS);
    case AB200_BF16:
      return emu ? launch_mlp<__nv_bfloat16, true>(AB200_ARGS) : launch_mlp<__nv_bfloat16, false>(AB200_ARGS);
    case AB200_F16:
      return emu ? launch_mlp<__half, true>(AB200_ARGS) : launch_mlp<__half, false>(AB200_ARGS);
    default: AB200_REQUIRE(false, AB200_E_INVALID, "sum_mlp_cell: unsupported dtype %d", dtype);
  }
#undef AB200_ARGS
}

extern "C" size_t ab200_mlp_workspace_bytes(const ab200_mlp_speculator* spec, int64_t B, int32_t num_steps) {
  if (!spec || B <= 0 || num_steps <= 0) return 0;
  return mlp_plan(spec, B, num_steps).total;
}

extern "C" int32_t ab200_mlp_propose(const ab200_mlp_speculator* sp, const int64_t* input_ids,
                                     const void* hidden, int64_t ld_hidden, int64_t B, int32_t num_steps,
                                     int64_t* tokens_out, void* states_out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  AB200_REQUIRE(sp != nullptr, AB200_E_INVALID, "mlp_propose: NULL speculator descriptor");
  if (B == 0 || num_steps == 0) return AB200_OK;
  AB200_REQUIRE(B > 0 && num_steps > 0 && num_steps <= AB200_MLP_MAX_STAGES && input_ids && hidden && tokens_out,
                AB200_E_INVALID, "mlp_propose: bad arguments (at most %d steps)", AB200_MLP_MAX_STAGES);
  AB200_REQUIRE(sp->dtype == AB200_BF16 || sp->dtype == AB200_F32, AB200_E_INVALID,
                "mlp_propose: dtype must be bf16 or f32");
  AB200_REQUIRE(sp->vocab > 0 && sp->head_rows >= sp->vocab && sp->inner_dim > 0 && sp->input_hidden_dim > 0,
                AB200_E_INVALID, "mlp_propose: bad model dimensions");
  for (int i = 0; i < num_steps; ++i)
    AB200_REQUIRE(sp->emb[i] && sp->proj[i] && sp->ln_w[i] && sp->ln_b[i] && sp->head[i], AB200_E_INVALID,
                  "mlp_propose: NULL weight pointer for step %d", i);
  const MlpWorkspace w = mlp_plan(sp, B, num_steps);
  AB200_REQUIRE(workspace && workspace_bytes >= w.total && aligned(workspace, 256), AB200_E_WORKSPACE,
                "mlp_propose: workspace of %zu bytes (256-byte aligned) required", w.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const size_t e = dtype_bytes(sp->dtype);
  const int64_t D = sp->inner_dim, H = sp->input_hidden_dim;
  const size_t bd = static_cast<size_t>(B) * D * e;
  uint64_t* packed = reinterpret_cast<uint64_t*>(ws + w.packed);
  uint32_t* counters = reinterpret_cast<uint32_t*>(ws + w.counters);
  const bool fused_tokens = sp->dtype == AB200_BF16 && B <= gemm_rows_per_pass();
  // reference selection rule (arctic_speculator.py:298-299): quantised head iff batch <= 32
  const bool fp8_ok = sp->dtype == AB200_BF16 && B <= sp->fp8_head_max_batch && D % 16 == 0;
  if (!fused_tokens)
    AB200_CUDA_OK(cudaMemsetAsync(packed, 0, static_cast<size_t>(B) * num_steps * sizeof(uint64_t), s));
  const void* xin = hidden;
  int64_t ldx = ld_hidden;
  if (sp->scale_input) {
    const float inv_sqrt2 = 1.0f / 1.41421356237309504880f;
    int rc = ab200_rms_scale(hidden, ws + w.x0, B, H, ld_hidden, H, sp->eps, inv_sqrt2, sp->dtype,
                             sp->emulate_eager_rounding != 0, stream);
    if (rc != AB200_OK) return rc;
    xin = ws + w.x0;
    ldx = H;
  }
  for (int step = 0; step < num_steps; ++step) {
    const int64_t K = step == 0 ? H : D;
    void* pre = ws + w.pre;
    int rc = ab200_linear(xin, sp->proj[step], pre, B, D, K, ldx, K, D, sp->dtype, stream);
    if (rc != AB200_OK) return rc;
    void* state = states_out ? static_cast<char*>(states_out) + static_cast<size_t>(step) * bd
                             : ws + w.state[step & 1];
    const bool fp8_head = fp8_ok && sp->head_q[step] != nullptr && sp->head_scale[step] != nullptr;
    void* state_q = fp8_head ? ws + w.state_q : nullptr;
    float* state_scale = fp8_head ? reinterpret_cast<float*>(ws + w.state_scale) : nullptr;
    rc = ab200_sum_mlp_cell(pre, D, sp->emb[step], step == 0 ? input_ids : nullptr,
                            step == 0 ? nullptr : packed + static_cast<size_t>(step - 1) * B,
                            sp->ln_w[step], sp->ln_b[step], state, B, D, sp->vocab, sp->alpha, sp->eps,
                            sp->dtype, sp->emulate_eager_rounding,
                            fused_tokens ? packed + static_cast<size_t>(step) * B : nullptr,
                            fused_tokens ? counters + step : nullptr, state_q, state_scale, stream);
    if (rc != AB200_OK) return rc;
    if (fp8_head)
      rc = ab200_head_argmax_fp8(state_q, state_scale, sp->head_q[step], sp->head_scale[step], B,
                                 sp->head_rows, D, sp->vocab, D, D, packed + static_cast<size_t>(step) * B,
                                 fused_tokens ? tokens_out + step : nullptr, num_steps,
                                 fused_tokens ? counters + step : nullptr, stream);
    else
      rc = ab200_head_argmax(state, sp->head[step], B, sp->head_rows, D, sp->vocab, D, D,
                             packed + static_cast<size_t>(step) * B,
                             fused_tokens ? tokens_out + step : nullptr, num_steps,
                             fused_tokens ? counters + step : nullptr, sp->dtype, ws + w.logits,
                             w.total - w.logits, stream);
    if (rc != AB200_OK) return rc;
    xin = state;
    ldx = D;
  }
  if (!fused_tokens) {
    const int64_t n = B * num_steps;
    mlp_unpack_steps_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const unsigned long long*>(packed), tokens_out, B, num_steps);
    AB200_CHECK_LAUNCH("mlp_unpack_steps_kernel");
  }
  return AB200_OK;
}
