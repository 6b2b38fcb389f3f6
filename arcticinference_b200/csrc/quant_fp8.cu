// quant_fp8.cu - FP8 (e4m3) quantisation for the speculator's quantised head.
//
// Mirrors what the reference's `qhead` does through vLLM (arctic_inference/vllm/spec_dec/fp8.py:
// process_weights_after_loading :207-221 -> ops.scaled_fp8_quant(weight, scale=None); apply
// :276-308 -> Fp8LinearOp: dynamic per-token activation quantisation + cutlass_scaled_mm):
//   weights    : one scale s_w = amax(|W|) / 448,                q = sat_e4m3(w * (1 / s_w))
//   activations: per row s_x = max(amax(|x_row|) / 448, 1/(448*512)), q = sat_e4m3(x / s_x)
//   output     : bf16( s_x[row] * (s_w * acc_fp32) )              (epilogue of the fp8 GEMM)
// Conversions are round-to-nearest-even with saturation to +-448 (satfinite).
#include "common.cuh"

namespace ab200 {

// amax over a [rows, K] bf16 matrix -> *amax_bits (uint view of a non-negative float, atomicMax)
__global__ void __launch_bounds__(256)
absmax_kernel(const __nv_bfloat16* __restrict__ w, int64_t rows, int64_t K, int64_t ld,
              unsigned int* __restrict__ amax_bits) {
  __shared__ float scratch[33];
  float m = 0.f;
  const int64_t nvec = K / 8;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const __nv_bfloat16* row = w + r * ld;
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      Vec16<__nv_bfloat16> in;
      in.u = ld_stream_16(row + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(__bfloat162float(in.e[j])));
    }
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) atomicMax(amax_bits, __float_as_uint(m));
}

__global__ void __launch_bounds__(256)
quant_weight_kernel(const __nv_bfloat16* __restrict__ w, uint8_t* __restrict__ q, int64_t rows,
                    int64_t K, int64_t ld, int64_t ldq, const unsigned int* __restrict__ amax_bits) {
  const float scale = __uint_as_float(*amax_bits) / kFp8Max;
  const float inv = 1.0f / scale;   // vLLM's per-tensor kernel multiplies by the inverted scale
  const int64_t nvec = K / 8;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const __nv_bfloat16* row = w + r * ld;
    uint8_t* qrow = q + r * ldq;
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      Vec16<__nv_bfloat16> in;
      in.u = ld_stream_16(row + v * 8);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lo |= static_cast<uint32_t>(to_e4m3(__bfloat162float(in.e[j]) * inv)) << (8 * j);
        hi |= static_cast<uint32_t>(to_e4m3(__bfloat162float(in.e[j + 4]) * inv)) << (8 * j);
      }
      *reinterpret_cast<uint2*>(qrow + v * 8) = make_uint2(lo, hi);
    }
  }
}

// amax bits -> scale, in place.  A separate launch: every CTA of quant_weight_kernel reads the amax
// word, so it may only be overwritten once that grid has finished.
__global__ void amax_to_scale_kernel(float* __restrict__ scale_io) {
  *scale_io = __uint_as_float(*reinterpret_cast<unsigned int*>(scale_io)) / kFp8Max;
}

// one CTA per row: dynamic per-token quantisation
__global__ void __launch_bounds__(1024)
quant_rows_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                  float* __restrict__ scales, int32_t K, int64_t ldx, int64_t ldq) {
  __shared__ float scratch[33];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * ldx;
  const int nvec = K / 8;
  float m = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    Vec16<__nv_bfloat16> in;
    in.u = *reinterpret_cast<const uint4*>(xr + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(__bfloat162float(in.e[j])));
  }
  m = block_max(m, scratch);
  const float scale = fmaxf(m / kFp8Max, kFp8MinScale);
  if (threadIdx.x == 0) scales[row] = scale;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    Vec16<__nv_bfloat16> in;
    in.u = *reinterpret_cast<const uint4*>(xr + v * 8);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lo |= static_cast<uint32_t>(to_e4m3(__bfloat162float(in.e[j]) / scale)) << (8 * j);
      hi |= static_cast<uint32_t>(to_e4m3(__bfloat162float(in.e[j + 4]) / scale)) << (8 * j);
    }
    *reinterpret_cast<uint2*>(q + row * ldq + v * 8) = make_uint2(lo, hi);
  }
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_fp8_quantize_weight(const void* w, void* q_out, float* scale_out,
                                             int64_t rows, int64_t K, int64_t ldw, int64_t ldq,
                                             void* stream) {
  AB200_REQUIRE(w && q_out && scale_out && rows > 0 && K > 0, AB200_E_INVALID,
                "fp8_quantize_weight: bad arguments");
  AB200_REQUIRE(K % 8 == 0 && ldw % 8 == 0 && ldq % 8 == 0 && aligned(w, 16) && aligned(q_out, 8),
                AB200_E_INVALID, "fp8_quantize_weight: K, ldw, ldq must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // scale_out doubles as the amax accumulator (uint view) between the two kernels
  AB200_CUDA_OK(cudaMemsetAsync(scale_out, 0, sizeof(float), s));
  const int grid = rows < 1184 ? static_cast<int>(rows) : 1184;
  unsigned int* amax = reinterpret_cast<unsigned int*>(scale_out);
  absmax_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(w), rows, K, ldw, amax);
  AB200_CHECK_LAUNCH("absmax_kernel");
  quant_weight_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(w),
                                           static_cast<uint8_t*>(q_out), rows, K, ldw, ldq, amax);
  AB200_CHECK_LAUNCH("quant_weight_kernel");
  amax_to_scale_kernel<<<1, 1, 0, s>>>(scale_out);
  AB200_CHECK_LAUNCH("amax_to_scale_kernel");
  return AB200_OK;
}

extern "C" int32_t ab200_fp8_quantize_rows(const void* x, void* q_out, float* scales_out, int64_t M,
                                           int64_t K, int64_t ldx, int64_t ldq, void* stream) {
  if (M == 0) return AB200_OK;
  AB200_REQUIRE(x && q_out && scales_out && M > 0 && K > 0, AB200_E_INVALID,
                "fp8_quantize_rows: bad arguments");
  AB200_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldq % 8 == 0 && aligned(x, 16) && aligned(q_out, 8),
                AB200_E_INVALID, "fp8_quantize_rows: K, ldx, ldq must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int threads = static_cast<int>((K / 8 + 31) / 32 * 32);
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  LaunchCfg lc(dim3(static_cast<unsigned>(M)), dim3(threads), 0, s);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  uint8_t* qp = static_cast<uint8_t*>(q_out);
  const int32_t k32 = static_cast<int32_t>(K);
  AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, quant_rows_kernel, xp, qp, scales_out, k32, ldx, ldq));
  AB200_CHECK_LAUNCH("quant_rows_kernel");
  return AB200_OK;
}
