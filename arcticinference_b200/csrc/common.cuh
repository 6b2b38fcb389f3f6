// common.cuh - shared helpers for the sm_100a kernels of libarctic_b200.so
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/arctic_b200.h"

namespace ab200 {

// ---- error plumbing (thread-local last-error string, launch counter) ------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define AB200_REQUIRE(cond, code, ...)          \
  do {                                          \
    if (!(cond)) {                              \
      ::ab200::set_error(__VA_ARGS__);          \
      return (code);                            \
    }                                           \
  } while (0)

#define AB200_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::ab200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                      \
      return AB200_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// Checks the launch that was just enqueued (cudaPeekAtLastError does not clear sticky errors and
// is legal during stream capture).
#define AB200_CHECK_LAUNCH(name)                                                  \
  do {                                                                            \
    cudaError_t _e = cudaPeekAtLastError();                                       \
    if (_e != cudaSuccess) {                                                      \
      ::ab200::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
      (void)cudaGetLastError();                                                   \
      return AB200_E_CUDA;                                                        \
    }                                                                             \
    ::ab200::count_launch();                                                      \
  } while (0)

// Launch-path tunables.  The release library uses the compiled-in defaults below (nothing on a launch
// path reads the environment); the debug flavour (-DAB200_DEBUG, libarctic_b200_dbg.so, loaded by
// tests/ and tools/ only) can change them through ab200_debug_set_tunable for parameter sweeps.
struct Tune {
  int pdl = 1;             // programmatic dependent launch between the kernels of a draft step
  int umma_m64 = 1;        // <= 64 batch rows: UMMA M = 64 (8 KiB activation slot -> more stages)
  int gemm_2cta = 1;       // 129..256 rows: CTA-pair kernel (tcgen05 cta_group::2)
  int rows_per_pass = 256; // batch rows per launch of the arg-max head / with fused token output (128 / 256)
  int gate_mh2 = 1;        // store-epilogue GEMM: 257..512 rows in ONE CTA-pair launch (two m-blocks per tile)
  int late_trigger = 1;    // GEMMs signal their dependents only after their own dependency wait
  int pre_stages = -1;     // smem stages filled with W before the dependency wait (-1 = all)
  int l2_prefetch_kb = 0;  // arg-max GEMM: extra W pipeline steps prefetched into L2 before the dependency wait
  int gate_l2_prefetch = 0; // store-epilogue GEMM: the same
  int gemm_4cta = 0;        // > 128 rows: clusters of two CTA pairs sharing the activation tile by TMA multicast
  int dbg_skip = 0;         // debug flavour only: CTA-pair kernel leaves operand loads out (timing experiments)
  int gate_block_n = 0;    // debug: force the tile width of the store-epilogue GEMM (0 = heuristic)
  int head_block_n = 0;    // debug: force the tile width of the arg-max GEMM
  int max_stages = 0;      // debug: cap the smem ring depth (0 = as many as fit)
  int gate_kbs = 2;        // k-blocks per TMA request / pipeline stage of the store-epilogue GEMM (1 or 2)
  int head_kbs = 2;        // ... of the arg-max (vocab head) GEMM (applies to <= 64 rows only)
  int pair_gate_kbs = 0;   // CTA-pair kernel (129..512 rows): k-blocks per TMA request, store-epilogue GEMM (0 = auto, 1..4)
  int pair_head_kbs = 0;   // the same for the arg-max head
};
#ifdef AB200_DEBUG
Tune& tune_mut();
inline const Tune& tune() { return tune_mut(); }
#else
inline const Tune& tune() {
  static constexpr Tune kTune{};
  return kTune;
}
#endif

// Programmatic dependent launch (PDL): kernels of one draft step are chained so that kernel N+1's
// prologue (barrier init, TMEM allocation, weight prefetch) overlaps kernel N's tail.  Every
// kernel calls pdl_launch_dependents() early and pdl_wait() before touching anything the previous
// kernel produced; both are no-ops for launches without the attribute.
inline bool pdl_enabled() { return tune().pdl != 0; }

// Optional device-side timeline (ab200_set_trace_buffer): every traced launch owns kTraceWords u64 words
// of %globaltimer nanoseconds (see trace_stamp).
// next_trace_slot() returns nullptr when tracing is off.  The trace state is per host thread.
unsigned long long* next_trace_slot();

// rows of a batch that share one pass over the weights in the tcgen05 GEMMs (gemm_sm100.cu)
inline int gemm_rows_per_pass() {
  const int e = tune().rows_per_pass;
  return e >= 512 ? 512 : (e >= 256 ? 256 : 128);
}
// 1 (default): the GEMM kernels signal their dependents only after their own dependency wait
inline int gemm_late_trigger() { return tune().late_trigger; }

// ---- per-device one-time state (thread-safe) --------------------------------------------------
constexpr int kMaxDevices = 64;
// SM count of the CURRENT device (cached per device ordinal)
int device_sm_count();
// One-time per-device work that must have COMPLETED before any later launch on any stream may rely
// on it (constant-table fills).  `fill(stream)` enqueues the work; device_once runs it once per
// device under a mutex and synchronises the stream, so other threads / streams can never observe a
// partly filled table.  First use inside a stream capture is refused (call ab200_init() - or any
// eager launch - before capturing).
struct DeviceOnce {
  std::atomic<int> done[kMaxDevices];
  std::mutex mu;
};
template <typename F>
inline int device_once(DeviceOnce& o, cudaStream_t s, const char* what, F&& fill) {
  int dev = 0;
  AB200_CUDA_OK(cudaGetDevice(&dev));
  AB200_REQUIRE(dev >= 0 && dev < kMaxDevices, AB200_E_INVALID, "device ordinal %d out of range", dev);
  if (o.done[dev].load(std::memory_order_acquire)) return AB200_OK;
  std::lock_guard<std::mutex> lock(o.mu);
  if (o.done[dev].load(std::memory_order_acquire)) return AB200_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  AB200_CUDA_OK(cudaStreamIsCapturing(s, &cap));
  AB200_REQUIRE(cap == cudaStreamCaptureStatusNone, AB200_E_INVALID,
                "%s: first use on device %d inside a stream capture - call ab200_init() before capturing",
                what, dev);
  int rc = fill(s);
  if (rc != AB200_OK) return rc;
  AB200_CUDA_OK(cudaStreamSynchronize(s));
  o.done[dev].store(1, std::memory_order_release);
  return AB200_OK;
}
// cudaFuncSetAttribute(max dynamic smem) once per (kernel, device); racing callers are harmless
template <typename K>
inline int ensure_max_smem(K kernel, std::atomic<bool> (&flags)[kMaxDevices], int bytes) {
  int dev = 0;
  AB200_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || !flags[dev].load(std::memory_order_acquire)) {
    AB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < kMaxDevices) flags[dev].store(true, std::memory_order_release);
  }
  return AB200_OK;
}
// per-translation-unit initialisers called by ab200_init()
int init_lstm_tables(cudaStream_t s);
int init_mlp_tables(cudaStream_t s);
int init_gemm_kernels();

struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
  }
};

inline int dtype_bytes(int32_t dt) { return dt == AB200_F32 ? 4 : 2; }

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- device helpers -------------------------------------------------------------------------
// Storage-type traits: load/store/round through the tensor dtype.
template <typename T> struct ST;
template <> struct ST<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <> struct ST<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float round(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
  }
};
template <> struct ST<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
};

// 16-byte vector of storage elements
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 u;
    T e[N];
  };
};

__device__ __forceinline__ uint4 ld_stream_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// plain (coherent) 16-byte load as volatile asm: keeps its place among the other volatile loads
__device__ __forceinline__ uint4 ld_ordered_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Trace words per launch: 0 = first CTA entry (min), 1 = dependency resolved (min), 2 = last CTA exit (max),
// 3 = first operand bytes landed (min), 4 = last operand bytes landed (max), 5 = first CTA exit (min).
// Words 3..5 are written by the GEMM kernels only.
constexpr int kTraceWords = 8;
__device__ __forceinline__ void trace_stamp(unsigned long long* slot, int which) {
  if (slot != nullptr) {
    const unsigned long long t = globaltimer_ns();
    if (which == 2 || which == 4) atomicMax(slot + which, t);
    else atomicMin(slot + which, t);
  }
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, result broadcast to every thread. `scratch` holds >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarps ? scratch[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  float r = scratch[32];
  __syncthreads();  // scratch may be reused immediately
  return r;
}

// Block-wide sum with ONE barrier: every warp reduces the per-warp partials itself (same xor tree,
// same operands as block_sum's warp 0 -> bit-identical result).  `scratch` (>= 32 floats) must not
// be reused before the next barrier of the caller; use a different buffer for back-to-back sums.
__device__ __forceinline__ float block_sum_1sync(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  return warp_sum(lane < nwarps ? scratch[lane] : 0.f);
}

// Block-wide max of non-negative values, broadcast. `scratch` holds >= 33 floats.
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarps ? scratch[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  const float r = scratch[32];
  __syncthreads();
  return r;
}

// FP8 e4m3 (vLLM semantics: round-to-nearest-even, saturate to +-448)
constexpr float kFp8Max = 448.0f;
constexpr float kFp8MinScale = 1.0f / (448.0f * 512.0f);
__device__ __forceinline__ uint8_t to_e4m3(float v) {
  return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
}

// Packed (value, index) for a 64-bit max: larger logit wins; equal logits -> lower index wins.
__device__ __forceinline__ uint32_t order_bits(float v) {
  v = v + 0.0f;  // -0.0 -> +0.0 so that the two zeros tie (torch.argmax compares with >)
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_arg(float v, uint32_t idx) {
  return (static_cast<unsigned long long>(order_bits(v)) << 32) |
         static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ uint32_t unpack_idx(unsigned long long p) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(p & 0xFFFFFFFFull);
}

}  // namespace ab200
