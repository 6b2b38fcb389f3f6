// gemm_sm100.cu - projection GEMMs of the Arctic LSTM speculator for sm_100a (hot path A).
//
//   out[M, N] = x[M, K] . w[N, K]^T        gate projection (arctic_speculator.py:555-561, 668)
//   argmax_n ( x[M, K] . w[N, K]^T )       vocab head + greedy pick, logits never materialised
//                                          (logits_processor_opt.py:83-107, arctic_speculator.py:726-732)
//
// Regime: M = draft batch (<= 128 per single-CTA launch, <= 512 per CTA-pair launch), N = 4D or the padded vocab,
// K = H or D.  The weight matrix is streamed from HBM exactly once per call and dominates traffic (bytes = N*K*2), so
// the kernel is a weight-streaming pipeline whose job is to keep HBM saturated:
//   * warp 0  : TMA producer - cp.async.bulk.tensor boxes (SWIZZLE_128B) of x [box_m x 64] and w [block_n x 64] into a
//               multi-stage smem ring, mbarrier complete_tx signalling.  It owns the stage barriers and starts
//               streaming W before TMEM is allocated and long before the dependency wait (the weights never depend
//               on the previous kernel).  Where it pays, ONE 3-D request brings TWO k-blocks per operand (kbs = 2):
//               the per-SM rate of TMA requests, not HBM, bounded the narrow-tile GEMMs (profiles/r2_gemm_tune.md);
//   * warp 1  : MMA issuer   - one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//               (M = 128 or 64, N = block_n, K = 16; bf16 x bf16 -> fp32) with the accumulator in TMEM;
//               tcgen05.commit releases smem stages / publishes the accumulator;
//   * warps 2-5: epilogue    - tcgen05.ld 32x32b (lane = batch row, column = output feature), so the vocab arg-max
//               is a thread-local scan; the load of the next 16 columns is in flight while the current ones are
//               converted / compared / stored; two TMEM accumulators let the epilogue of tile i overlap the main
//               loop of tile i+1.
//   * batch rows are the UMMA M dimension (lanes), features the N dimension: rows >= M are never
//     fetched (TMA box covers box_m rows only) and cost no HBM traffic.
//   * block_n is chosen per problem so that the tile count is a near multiple of the SM count
//     (e.g. N = 16384 -> 147 tiles of 112 on 148 SMs).
//   * bf16 results are rounded once from the fp32 accumulator, like the reference's cuBLAS GEMM;
//     the arg-max compares the bf16-rounded logits and keeps the lowest index on ties
//     (torch.argmax), across CTAs through a packed 64-bit atomicMax.
//   * every launch can record a device-side trace (entry, dependency resolved, first / last operand bytes
//     landed, first / last CTA exit) - ab200_set_trace_buffer - which is how ramp / stream / tail are told apart.
//
// The f32 parity mode (1e-5) uses a plain SIMT FFMA GEMM (exact fp32, no TF32) + row arg-max.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace ab200 {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU - after ~2 s the kernel traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int32_t c0,
                                            int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// 3-D box {128 bytes of K, rows, k-blocks}: ONE request brings several k-blocks, each landing as its own
// [rows x 128 B] SWIZZLE_128B slab (the k-block index is the outermost box dimension)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int32_t c0,
                                            int32_t c1, int32_t c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* desc, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const void* desc, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, single CTA, fp32 accumulate.
// KIND 0: bf16 x bf16 (kind::f16, K = 16 per instruction); KIND 1: e4m3 x e4m3 (kind::f8f6f4, K = 32)
enum { KIND_BF16 = 0, KIND_FP8 = 1 };
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                     uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == KIND_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// tcgen05 weight-streaming GEMM
// ------------------------------------------------------------------------------------------
constexpr int kRowBytes = 128;  // one SWIZZLE_128B row = one k-block: 64 bf16 or 128 e4m3 elements
constexpr int kBlockK = 64;     // bf16 elements per k-block (fp8: 128)
constexpr int kUmmaK = 16;      // bf16 elements per MMA (fp8: 32); both are 32 bytes of the row
constexpr int kTileM = 128;   // UMMA M = TMEM lanes = batch rows per m-tile
constexpr int kMaxBlockN = 256;
constexpr int kGemmThreads = 192;
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KiB slot per m-tile (only box_m rows are written)
constexpr int kTmemCols = 512;                 // 2 accumulator buffers x MT m-tiles x (256 / MT) columns
constexpr int kMaxStages = 12;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxRowsPerPassLimit = 512;      // MT <= 4 m-tiles can share one weight stream
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;  // weights: streamed once
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;   // activations: re-read by every tile

enum { EPI_STORE = 0, EPI_ARGMAX = 1 };

struct GemmParams {
  int32_t m_rows;       // valid rows in this pass (<= MT * 128)
  int32_t n_valid;      // columns that take part in the epilogue (N, or vocab for arg-max)
  int32_t block_n;
  int32_t num_n_tiles;
  int32_t num_k_blocks;
  int32_t stages;
  int32_t l2_prefetch_kb;  // extra W k-blocks prefetched into L2 before the dependency wait
#ifdef AB200_DEBUG
  int32_t dbg_skip;        // 2cta kernel, timing experiments only: bit 0 = no x loads, bit 1 = no W loads after the ring fill
#endif
  int32_t pre_stages;      // smem stages filled with W before the dependency wait
  int32_t late_trigger;    // 1: griddepcontrol.launch_dependents after the last TMA load
  int32_t m64;             // 1: UMMA M = 64 (MT == 1, <= 64 rows): row r lives in TMEM lane (r/16)*32 + r%16
  int32_t mh;              // CTA-pair kernel: 256-row m-blocks that share one launch (1 or 2)
  int32_t kbs;             // k-blocks per pipeline stage = per TMA request (1: 2-D maps; 2: 3-D maps, MT == 1)
  uint32_t a_pitch;        // kbs == 2: bytes between the k-block slabs of the activation box (box rows * 128)
  uint32_t tx_bytes;    // bytes landing per stage
  uint32_t idesc;
  __nv_bfloat16* out;   // EPI_STORE
  int64_t ldo;
  unsigned long long* packed;  // EPI_ARGMAX, [m_rows]
  int64_t* tokens_out;         // EPI_ARGMAX, optional: written by the last CTA to finish
  int64_t tokens_stride;
  unsigned int* done_counter;  // EPI_ARGMAX, zero on entry (needed iff tokens_out)
  unsigned long long* trace;   // optional timeline slot (profiling)
  const float* x_scales;       // KIND_FP8: per-row activation scales [m_rows]
  const float* w_scale;        // KIND_FP8: per-tensor weight scale [1]
  uint32_t a_slot;             // bytes of a stage reserved for the activation tile(s)
};

// K-major SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);   // start address        bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                   // LBO (unused here)    bits [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;           // SBO = 1024 B         bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                   // descriptor version 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                   // SWIZZLE_128B
  return d;
}

// MT = number of 128-row m-tiles (batch rows up to MT*128) that share ONE pass over the weights.
template <int EPI, int MT, int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_sm100(const __grid_constant__ CUtensorMap tmap_x,
                   const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  constexpr int kAccSlot = kMaxBlockN / MT;   // TMEM columns per (buffer, m-tile); block_n <= kAccSlot
  // UMMA M = 64 (<= 64 batch rows) needs an 8 KiB activation slot instead of 16: more stages, i.e.
  // more weight bytes in flight per SM
  const uint32_t kASlot = MT == 1 ? p.a_slot : MT * kABytes;
  constexpr int kKElems = KIND == KIND_BF16 ? 64 : 128;   // elements per 128-byte k-block
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * 128u * (MT == 1 ? static_cast<uint32_t>(p.kbs) : 1u);
  const uint32_t stage_bytes = kASlot + b_bytes;
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4);
  const uint32_t flag_slot = tmem_slot + 8u;

  const int num_kb = p.num_k_blocks;   // pipeline steps per tile (each of kbs k-blocks)
  const int my_tiles = static_cast<int>(blockIdx.x) < p.num_n_tiles
                           ? (p.num_n_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1
                           : 0;
  // producer's view: linearised (tile, k-block) sequence of this CTA
  const int total = my_tiles * num_kb;
  // kbs == 2: a stage holds TWO k-blocks fetched by ONE request per operand (3-D tensor map).  The per-SM rate
  // of TMA requests, not HBM, bounded the narrow-tile GEMM (time per k-block ~0.3 us whatever the box size,
  // profiles/r2_gemm_tune.md): half the requests and half the barrier round trips for the same bytes in flight.
  // A k-block index beyond K/64 (odd block counts) is out of bounds for the map and arrives as zeros.
  const int kbs = MT == 1 ? p.kbs : 1;
  const uint32_t b_pitch = b_bytes / kbs;     // bytes of one k-block slab of W
  auto n0_of = [&](int g) { return (static_cast<int>(blockIdx.x) + (g / num_kb) * static_cast<int>(gridDim.x)) * p.block_n; };
  // (kb = pipeline step inside the tile, n0 = first weight row of the tile: the steady-state loop keeps them as
  //  running counters - no division by run-time values on the producer's per-step path)
  auto load_x_at = [&](int kb, int stage) {
    const uint32_t sa = smem_base + stage * stage_bytes;
    if (kbs == 2) {
      tma_load_3d(sa, &tmap_x, full_bar(stage), 0, 0, 2 * kb, kEvictLast);
    } else {
#pragma unroll
      for (int t = 0; t < MT; ++t)
        tma_load_2d(sa + t * kABytes, &tmap_x, full_bar(stage), kb * kKElems, t * kTileM, kEvictLast);
    }
  };
  auto load_w_at = [&](int kb, int n0, int stage) {
    const uint32_t sa = smem_base + stage * stage_bytes;
    if (kbs == 2) tma_load_3d(sa + kASlot, &tmap_w, full_bar(stage), 0, n0, 2 * kb, kEvictFirst);
    else tma_load_2d(sa + kASlot, &tmap_w, full_bar(stage), kb * kKElems, n0, kEvictFirst);
  };
  auto load_x = [&](int g, int stage) { load_x_at(g % num_kb, stage); };
  auto load_w = [&](int g, int stage) { load_w_at(g % num_kb, n0_of(g), stage); };
  const int pre = total < p.pre_stages ? total : p.pre_stages;

  if (warp == 0 && lane == 0) {
    trace_stamp(p.trace, 0);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    // The producer owns the stage barriers and starts streaming right away: the weights do not depend
    // on the previous kernel, so the ring is filled with W (and L2 warmed with the k-blocks after
    // that) while warp 1 is still allocating TMEM and long BEFORE the dependency wait.
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    fence_mbar_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int g = 0; g < pre; ++g) {
      mbar_arrive_expect_tx(full_bar(g), p.tx_bytes);
      load_w(g, g);
    }
    const int l2_end = total < pre + p.l2_prefetch_kb ? total : pre + p.l2_prefetch_kb;
    for (int g = pre; g < l2_end; ++g) {
      if (kbs == 2) tma_prefetch_l2_3d(&tmap_w, 0, n0_of(g), 2 * (g % num_kb));
      else tma_prefetch_l2_2d(&tmap_w, (g % num_kb) * kKElems, n0_of(g));
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4);  // one arrival per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // Programmatic dependent launch: the next kernel in the stream may start its prologue once every
  // CTA of this grid has signalled (it still blocks in griddepcontrol.wait until this grid has
  // completed and flushed).  late_trigger: signal only after this CTA has issued its last TMA
  // load, so that the dependent's launch overlaps just the tail of this kernel.
  if (!p.late_trigger) pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      trace_stamp(p.trace, 1);
      for (int g = 0; g < pre; ++g) load_x(g, g);
      int stage = pre == stages ? 0 : pre;
      uint32_t phase = pre == stages ? 1u : 0u;
      int kb = pre % num_kb;
      int n0 = n0_of(pre);
      const int n0_step = static_cast<int>(gridDim.x) * p.block_n;
      for (int g = pre; g < total; ++g) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), p.tx_bytes);
        load_x_at(kb, stage);
        load_w_at(kb, n0, stage);
        if (++kb == num_kb) { kb = 0; n0 += n0_step; }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (p.late_trigger) pdl_launch_dependents();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // (no griddepcontrol.wait here: this warp touches no global memory and is gated by the
    //  full barriers, which complete only after the producer's post-wait x loads)
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);  // TMA bytes have landed
        tc_fence_after();
        if (p.trace != nullptr && lane == 0) {
          if (it == 0 && kb == 0) trace_stamp(p.trace, 3);
          if (it == my_tiles - 1 && kb == num_kb - 1) trace_stamp(p.trace, 4);
        }
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (MT == 1 && p.kbs == 2) {
            const uint32_t tmem_d = tmem_base + acc * kAccSlot;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t da = smem_desc_sw128(sa + kk * p.a_pitch);
              const uint64_t db = smem_desc_sw128(sa + kASlot + kk * (b_bytes >> 1));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma<KIND>(tmem_d, da + 2u * k, db + 2u * k, p.idesc, (kb | kk | k) != 0 ? 1u : 0u);
            }
          } else {
            const uint64_t db = smem_desc_sw128(sa + kASlot);
#pragma unroll
            for (int t = 0; t < MT; ++t) {
              const uint64_t da = smem_desc_sw128(sa + t * kABytes);
              const uint32_t tmem_d = tmem_base + (acc * MT + t) * kAccSlot;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // one MMA consumes 32 B of the 128 B swizzle row: +2 in the (>>4) address field
                umma<KIND>(tmem_d, da + 2u * k, db + 2u * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(empty_bar(stage));                      // smem stage reusable once MMAs retire
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));  // accumulators complete
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // Gated by the accumulator-full barrier, i.e. transitively by the producer's dependency wait.
    // Everything it touches from the previous kernel (packed slots, done counter) is accessed
    // with L2 atomics only, so these warps sleep in mbarrier.try_wait instead of sitting in
    // griddepcontrol.wait next to the (issue-bound) cell kernel that shares the SM.
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    // batch row of this thread inside m-tile 0 (M = 128: lane = row; M = 64: 16 rows per quadrant)
    const bool m64 = MT == 1 && p.m64 != 0;
    const int row_in_tile = m64 ? quad * 16 + lane : quad * 32 + lane;
    const bool lane_used = !m64 || lane < 16;
    const int warp_row0 = m64 ? quad * 16 : quad * 32;
    float best[MT];
    uint32_t best_idx[MT];
#pragma unroll
    for (int t = 0; t < MT; ++t) { best[t] = -INFINITY; best_idx[t] = 0xFFFFFFFFu; }
    float sa[MT], sb = 1.0f;   // fp8: out = s_x[row] * (s_w * acc), the order of vLLM's ScaledEpilogue
    bool scales_loaded = false;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int n0 = (static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x)) * p.block_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if constexpr (KIND == KIND_FP8) {
        if (!scales_loaded) {   // produced by the previous kernel: read (L2) only after the barrier
          scales_loaded = true;
          sb = __ldcg(p.w_scale);
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const int r = t * kTileM + row_in_tile;
            sa[t] = (lane_used && r < p.m_rows) ? __ldcg(p.x_scales + r) : 0.f;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int row = t * kTileM + row_in_tile;   // batch row inside the pass
        const bool row_ok = lane_used && row < p.m_rows;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (acc * MT + t) * kAccSlot;
        if (t * kTileM + warp_row0 >= p.m_rows) continue;  // whole warp out of range (warp-uniform)
        // One 16-column chunk; the TMEM load of the NEXT chunk is already in flight while this runs
        // (double-buffered registers: the epilogue was a serial ld -> wait -> convert -> store chain).
        auto process = [&](uint32_t (&r)[16], int c0) {
          if constexpr (KIND == KIND_FP8) {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(sa[t] * (sb * __uint_as_float(r[j])));
          }
          const int col0 = n0 + c0;
          if constexpr (EPI == EPI_STORE) {
            if (row_ok && col0 < p.n_valid) {
              __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col0;
              if (col0 + 16 <= p.n_valid) {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * j]),
                                                           __uint_as_float(r[2 * j + 1]));
                  w[j] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              } else {
                for (int j = 0; j < 16 && col0 + j < p.n_valid; ++j)
                  dst[j] = __float2bfloat16_rn(__uint_as_float(r[j]));
              }
            }
          } else {
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                // the reference's GEMM returns bf16 logits: compare after rounding
                const float v = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j])));
                const int col = col0 + j;
                if (col < p.n_valid && v > best[t]) {  // strict >: lowest index wins ties
                  best[t] = v;
                  best_idx[t] = static_cast<uint32_t>(col);
                }
              }
            }
          }
        };
        uint32_t ra[16], rb[16];
        tmem_ld16(taddr, ra);
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          tmem_ld_wait();
          const bool has_b = c0 + 16 < p.block_n;          // warp-uniform
          if (has_b) tmem_ld16(taddr + c0 + 16, rb);
          process(ra, c0);
          if (has_b) {
            tmem_ld_wait();
            if (c0 + 32 < p.block_n) tmem_ld16(taddr + c0 + 32, ra);
            process(rb, c0 + 16);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if constexpr (EPI == EPI_ARGMAX) {
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int row = t * kTileM + row_in_tile;
        if (lane_used && row < p.m_rows && best_idx[t] != 0xFFFFFFFFu)
          atomicMax(p.packed + row, pack_arg(best[t], best_idx[t]));
      }
      if (p.tokens_out != nullptr) {
        // The last CTA to finish turns the packed maxima into int64 token ids (folds the
        // separate unpack launch).  128 epilogue threads, named barrier 1.
        const int et = threadIdx.x - 64;
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const unsigned int prev = atomicAdd(p.done_counter, 1u);
          const uint32_t last = prev == gridDim.x - 1 ? 1u : 0u;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag_slot), "r"(last) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t last;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(flag_slot) : "memory");
        if (last) {
          __threadfence();
          for (int r = et; r < p.m_rows; r += 128) {
            const unsigned long long pk = *reinterpret_cast<volatile unsigned long long*>(p.packed + r);
            p.tokens_out[static_cast<int64_t>(r) * p.tokens_stride] = static_cast<int64_t>(unpack_idx(pk));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) {
    trace_stamp(p.trace, 2);
    trace_stamp(p.trace, 5);
  }
}


// ------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2) for 129..256 batch rows per pass.
//
// Two CTAs of one cluster (the two SMs of a TPC) execute ONE UMMA of M = 256 x N = block_n:
// CTA r holds batch rows [128 r, 128 r + 128) of x and the accumulators of those rows in its own
// TMEM, and HALF of the weight tile (block_n / 2 rows) in its shared memory; the tensor cores of
// both SMs read both halves.  Per output element an SM therefore ingests 1.0 byte from L2 instead
// of the 1.5 bytes of the single-CTA MT = 2 shape - that ingest (13.9 TB/s aggregate at B = 256),
// not HBM or the tensor pipe, is what bounds the single-CTA kernel above 128 rows.
//
// Protocol (after CUTLASS' 2-SM pipelines): the stage-full barrier lives in the leader CTA (rank 0)
// and counts the TMA bytes of BOTH CTAs (cp.async.bulk.tensor.cta_group::2 with the peer bit of the
// barrier address cleared); only the leader issues tcgen05.mma.cta_group::2; tcgen05.commit with
// .multicast::cluster releases the shared-memory stage / publishes the accumulators in both CTAs;
// the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in CTA rank 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on rank 0's barrier at this offset
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* desc, uint32_t bar, int32_t c0,
                                                int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const void* desc, uint32_t bar, int32_t c0,
                                                int32_t c1, int32_t c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
// the same load, delivered to the same shared-memory offset of every CTA in `mask` (cluster ranks); the bytes are
// counted on the stage barrier of EACH destination's pair leader (peer bit cleared in every destination)
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const void* desc, uint32_t bar, int32_t c0,
                                                   int32_t c1, uint16_t mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "h"(mask), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_2sm_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// CL = 2: one CTA pair per cluster.  CL = 4: TWO pairs per cluster work on ADJACENT weight tiles and share the
// activation k-blocks: every CTA fetches half of the 128 rows it needs and multicasts them to its counterpart in the
// other pair, so the x bytes crossing L2 -> SM halve.  (Measured, profiles/r2_gemm_tune.md section 6: at 256 rows
// the pair kernel is bound by the L2 -> SM fabric - 2.05 GB per head launch at ~8.8 TB/s, half of it x re-reads -
// neither by HBM (167 us) nor by the tensor pipe (145 us).)  A stage is released only when BOTH pairs have consumed
// it (their commits arrive on the stage-empty barrier of all four CTAs), since both pairs' producers write into it.
template <int EPI, int CL>
__device__ __forceinline__ void gemm_pair_body(const CUtensorMap& tmap_x, const CUtensorMap& tmap_w,
                                               const GemmParams& p) {
  constexpr int kAcc = 256;   // TMEM columns per accumulator buffer (block_n <= 256), two buffers
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;                           // rank inside the MMA pair
  const int sub = CL == 4 ? static_cast<int>(crank >> 1) : 0; // which pair of the cluster
  // `pair` / `npairs`: index and number of the units that own weight tiles round-robin (CL = 4: clusters)
  const int pair = static_cast<int>(blockIdx.x) / CL, npairs = static_cast<int>(gridDim.x) / CL;
  auto tile_of = [&](int round) { return CL == 4 ? 2 * (pair + round * npairs) + sub : pair + round * npairs; };
  const int stages = p.stages;
  const int half_n = p.block_n >> 1;
  // kbs > 1: a stage holds kbs k-blocks per operand, fetched by ONE 3-D request each ([A kb0 | A kb1 ..][B kb0 | B kb1 ..]):
  // half the TMA requests and barrier round trips per byte, as in the single-CTA kernel (profiles/r2_gemm_tune.md s.2, s.8)
  const int kbs = CL == 4 ? 1 : p.kbs;
  const uint32_t b_pitch = static_cast<uint32_t>(half_n) * 128u;   // one k-block slab of this CTA's half of the W tile
  const uint32_t a_bytes = kABytes * static_cast<uint32_t>(kbs);
  const uint32_t b_bytes = b_pitch * static_cast<uint32_t>(kbs);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                      // used in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };          // one per CTA
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };      // one per CTA
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); }; // used in the leader only
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4);
  const uint32_t flag_slot = tmem_slot + 8u;

  if (warp == 0 && lane == 0) {
    trace_stamp(p.trace, 0);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), CL / 2);   // one commit per pair that reads (and whose producers refill) the stage
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 8);   // 4 epilogue warps x 2 CTAs
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (!p.late_trigger) pdl_launch_dependents();

  const int num_kb = p.num_k_blocks;
  // CL = 4: the unit owns tile PAIRS (an odd tile count leaves the last partner out of bounds: zero-filled by TMA,
  // masked by n_valid in the epilogue) - both pairs of a cluster must run the same number of pipeline steps
  const int units_of_work = CL == 4 ? (p.num_n_tiles + 1) / 2 : p.num_n_tiles;
  const int my_tiles = pair < units_of_work ? (units_of_work - 1 - pair) / npairs + 1 : 0;
  // Batches of 257..512 rows: every weight tile is used for MH = 2 m-blocks of 256 rows back to back (work item =
  // (tile, m-block), m-block fastest), each into its own TMEM accumulator buffer - ONE launch and one pass over
  // the tile list instead of two launches that each stream all the weights (the second read of a tile comes from
  // L2 where it still is; at these batches the kernel is tensor-bound either way).
  const int MH = p.mh;
  const int my_items = my_tiles * MH;
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const int total = my_items * num_kb;
      // Position in this CTA's (tile, m-block, k-step) sequence as running counters: the producer thread's per-step
      // instruction path is on the critical path of the pair kernel (a handful of extra instructions per step cost
      // the debug flavour 15 %), so no division / modulo by run-time values in here.
      struct Cursor { int kb, mhi, round; };
      auto advance = [&](Cursor& c) {
        if (++c.kb == num_kb) {
          c.kb = 0;
          if (++c.mhi == MH) { c.mhi = 0; ++c.round; }
        }
      };
      auto load_x = [&](const Cursor& c, int stage) {
        const int row0 = c.mhi * 2 * kTileM + static_cast<int>(rank) * kTileM;
        const uint32_t dst = smem_base + stage * stage_bytes;
        if constexpr (CL == 4) {
          // rows [sub*64, sub*64 + 64) of this CTA's 128, to both CTAs of the cluster with this pair rank
          tma_load_2d_2sm_mc(dst + sub * (kABytes / 2), &tmap_x, full_bar(stage), c.kb * kBlockK,
                             row0 + sub * (kTileM / 2), static_cast<uint16_t>((1u << rank) | (4u << rank)), kEvictLast);
        } else if (kbs > 1) {
          tma_load_3d_2sm(dst, &tmap_x, full_bar(stage), 0, row0, kbs * c.kb, kEvictLast);
        } else {
          tma_load_2d_2sm(dst, &tmap_x, full_bar(stage), c.kb * kBlockK, row0, kEvictLast);
        }
      };
      auto load_w = [&](const Cursor& c, int stage) {
        const int n0 = tile_of(c.round) * p.block_n + static_cast<int>(rank) * half_n;
        const uint32_t dst = smem_base + stage * stage_bytes + a_bytes;
        if (kbs > 1) tma_load_3d_2sm(dst, &tmap_w, full_bar(stage), 0, n0, kbs * c.kb, kEvictFirst);
        else tma_load_2d_2sm(dst, &tmap_w, full_bar(stage), c.kb * kBlockK, n0, kEvictFirst);
      };
      const int pre = total < p.pre_stages ? total : p.pre_stages;
      Cursor cw{0, 0, 0}, cx{0, 0, 0};
      for (int g = 0; g < pre; ++g) {
        if (rank == 0) mbar_arrive_expect_tx(full_bar(g), p.tx_bytes);
        load_w(cw, g);
        advance(cw);
      }
      pdl_wait();
      trace_stamp(p.trace, 1);
      for (int g = 0; g < pre; ++g) {
        load_x(cx, g);
        advance(cx);
      }
      int stage = pre == stages ? 0 : pre;
      uint32_t phase = pre == stages ? 1u : 0u;
      for (int g = pre; g < total; ++g) {       // cw == cx from here on
        mbar_wait(empty_bar(stage), phase ^ 1u);
#ifdef AB200_DEBUG
        // bound-finding experiment (WRONG results, timing only): leave operand loads out after the first ring fill
        const uint32_t skip = static_cast<uint32_t>(p.dbg_skip);
        const uint32_t tx = p.tx_bytes - ((skip & 1u) ? 2u * a_bytes : 0u) - ((skip & 2u) ? 2u * b_bytes : 0u);
        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), tx);
        if (!(skip & 1u)) load_x(cw, stage);
        if (!(skip & 2u)) load_w(cw, stage);
#else
        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), p.tx_bytes);
        load_x(cw, stage);
        load_w(cw, stage);
#endif
        advance(cw);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (p.late_trigger) pdl_launch_dependents();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_items; ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);          // the bytes of both CTAs have landed
          tc_fence_after();
          if (p.trace != nullptr && lane == 0) {
            if (it == 0 && kb == 0) trace_stamp(p.trace, 3);
            if (it == my_items - 1 && kb == num_kb - 1) trace_stamp(p.trace, 4);
          }
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t tmem_d = tmem_base + acc * kAcc;
            for (int kk = 0; kk < kbs; ++kk) {
              const uint64_t da = smem_desc_sw128(sa + kk * kABytes);
              const uint64_t db = smem_desc_sw128(sa + a_bytes + kk * b_pitch);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_2sm_bf16(tmem_d, da + 2u * k, db + 2u * k, p.idesc, (kb | kk | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar(stage), CL == 4 ? 0xF : 0x3);
            if (kb == num_kb - 1) umma_commit_2sm(tfull_bar(acc), static_cast<uint16_t>(0x3u << (2 * sub)));
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int quad = warp & 3;
    const int row_in_block = static_cast<int>(rank) * kTileM + quad * 32 + lane;   // row inside a 256-row m-block
    float bests[2] = {-INFINITY, -INFINITY};
    uint32_t best_idxs[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    for (int it = 0; it < my_items; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int mhi = it % MH;
      const int n0 = tile_of(it / MH) * p.block_n;
      const int row = mhi * 2 * kTileM + row_in_block;                     // batch row of this thread
      const bool row_ok = row < p.m_rows;
      const bool warp_has_rows = mhi * 2 * kTileM + static_cast<int>(rank) * kTileM + quad * 32 < p.m_rows;
      float best = mhi ? bests[1] : bests[0];
      uint32_t best_idx = mhi ? best_idxs[1] : best_idxs[0];
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAcc;
      if (warp_has_rows) {
        auto process = [&](uint32_t (&r)[16], int c0) {
          const int col0 = n0 + c0;
          if constexpr (EPI == EPI_STORE) {
            if (row_ok && col0 < p.n_valid) {
              __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col0;
              if (col0 + 16 <= p.n_valid) {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                  w[j] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              } else {
                for (int j = 0; j < 16 && col0 + j < p.n_valid; ++j) dst[j] = __float2bfloat16_rn(__uint_as_float(r[j]));
              }
            }
          } else {
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float v = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j])));
                const int col = col0 + j;
                if (col < p.n_valid && v > best) {
                  best = v;
                  best_idx = static_cast<uint32_t>(col);
                }
              }
            }
          }
        };
        uint32_t ra[16], rb[16];
        tmem_ld16(taddr, ra);
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          tmem_ld_wait();
          const bool has_b = c0 + 16 < p.block_n;          // warp-uniform
          if (has_b) tmem_ld16(taddr + c0 + 16, rb);
          process(ra, c0);
          if (has_b) {
            tmem_ld_wait();
            if (c0 + 32 < p.block_n) tmem_ld16(taddr + c0 + 32, ra);
            process(rb, c0 + 16);
          }
        }
      }
      if (mhi) { bests[1] = best; best_idxs[1] = best_idx; }
      else { bests[0] = best; best_idxs[0] = best_idx; }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
    }
    if constexpr (EPI == EPI_ARGMAX) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int row = m * 2 * kTileM + row_in_block;
        if (m < MH && row < p.m_rows && best_idxs[m] != 0xFFFFFFFFu)
          atomicMax(p.packed + row, pack_arg(bests[m], best_idxs[m]));
      }
      if (p.tokens_out != nullptr) {
        const int et = threadIdx.x - 64;
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const unsigned int prev = atomicAdd(p.done_counter, 1u);
          const uint32_t last = prev == gridDim.x - 1 ? 1u : 0u;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag_slot), "r"(last) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t last;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(flag_slot) : "memory");
        if (last) {
          __threadfence();
          for (int r = et; r < p.m_rows; r += 128) {
            const unsigned long long pk = *reinterpret_cast<volatile unsigned long long*>(p.packed + r);
            p.tokens_out[static_cast<int64_t>(r) * p.tokens_stride] = static_cast<int64_t>(unpack_idx(pk));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves (and frees its barriers / TMEM) while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) {
    trace_stamp(p.trace, 2);
    trace_stamp(p.trace, 5);
  }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tn_sm100_2cta(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const GemmParams p) {
  gemm_pair_body<EPI, 2>(tmap_x, tmap_w, p);
}

#ifdef AB200_DEBUG
// Two CTA pairs per cluster sharing the activation tile by TMA multicast (tmap_x box = 64 rows).  DEBUG flavour only:
// bit-identical to the pair kernel and ~12 % faster per SM, but a B200 co-schedules only 33 clusters of 4 (132 of
// its 148 SMs - GPC sizes are not multiples of 4), which cancels the gain (tools/cl4_probe.py,
// profiles/r2_gemm_tune.md section 6).
template <int EPI>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tn_sm100_4cta(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const GemmParams p) {
  gemm_pair_body<EPI, 4>(tmap_x, tmap_w, p);
}
#endif

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // magic static: initialised exactly once, thread-safe
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    EncodeTiledFn f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      f = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
    return f;
  }();
  return fn;
}

// 2-D tensor map over a row-major [rows, K] matrix (row stride ld elements of esize bytes),
// box = [box_rows x 128 bytes]
static int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld,
                     int box_rows, int esize) {
  EncodeTiledFn enc = get_encode_fn();
  AB200_REQUIRE(enc != nullptr, AB200_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kRowBytes / esize), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                   const_cast<void*>(base), gdim, gstr, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AB200_REQUIRE(r == CUDA_SUCCESS, AB200_E_CUDA,
                "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld box_rows=%d", (int)r,
                (long long)rows, (long long)K, (long long)ld, box_rows);
  return AB200_OK;
}

// 3-D view of the same matrix for multi-k-block requests: dims {128 bytes of K, rows, K / k_elems}, box
// {k_elems, box_rows, kbs}; needs K % k_elems == 0 (a partial k-block would run into the next row)
static int make_tmap3(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows,
                      int esize, int kbs) {
  EncodeTiledFn enc = get_encode_fn();
  AB200_REQUIRE(enc != nullptr, AB200_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const int k_elems = kRowBytes / esize;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(k_elems), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(K / k_elems)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * esize, static_cast<cuuint64_t>(kRowBytes)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(k_elems), static_cast<cuuint32_t>(box_rows),
                       static_cast<cuuint32_t>(kbs)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AB200_REQUIRE(r == CUDA_SUCCESS, AB200_E_CUDA,
                "cuTensorMapEncodeTiled(3d) failed (%d) rows=%lld K=%lld ld=%lld box_rows=%d", (int)r,
                (long long)rows, (long long)K, (long long)ld, box_rows);
  return AB200_OK;
}

static int sm_count() { return device_sm_count(); }

// Tensor maps depend only on (base, rows, K, ld, box rows, element size), never on the contents, so a
// cached map stays valid for as long as those match: a small per-thread cache removes the two driver
// encode calls from every eager launch (graph replays never reach this code).
struct TmapKey {
  const void* base;
  int64_t rows, K, ld;
  int32_t box_rows, esize, kbs;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && K == o.K && ld == o.ld && box_rows == o.box_rows &&
           esize == o.esize && kbs == o.kbs;
  }
};
struct TmapCache {
  static constexpr int kEntries = 32;
  TmapKey keys[kEntries];
  alignas(64) CUtensorMap maps[kEntries];
  int used = 0, next = 0;
};
static int get_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows,
                    int esize, int kbs = 1) {
  static thread_local TmapCache cache;
  const TmapKey key{base, rows, K, ld, box_rows, esize, kbs};
  for (int i = 0; i < cache.used; ++i)
    if (cache.keys[i] == key) {
      *map = cache.maps[i];
      return AB200_OK;
    }
  int rc = kbs == 1 ? make_tmap(map, base, rows, K, ld, box_rows, esize)
                    : make_tmap3(map, base, rows, K, ld, box_rows, esize, kbs);
  if (rc != AB200_OK) return rc;
  const int slot = cache.used < TmapCache::kEntries ? cache.used++ : (cache.next++ % TmapCache::kEntries);
  cache.keys[slot] = key;
  cache.maps[slot] = *map;
  return AB200_OK;
}

// Tile width: minimise waves * (block_n + fixed per-tile cost) so the tile count lands just under
// a multiple of the SM count; ties go to the wider tile.
int pick_block_n(int64_t N, int sms, int max_bn) {
  int best_bn = 16;
  double best_cost = 1e30;
  for (int bn = 16; bn <= max_bn; bn += 16) {
    const int64_t tiles = (N + bn - 1) / bn;
    const int64_t waves = (tiles + sms - 1) / sms;
    const double cost = static_cast<double>(waves) * (bn + 8.0);
    if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && bn > best_bn)) {
      best_cost = cost;
      best_bn = bn;
    }
  }
  return best_bn;
}

struct ArgmaxExtra {
  unsigned long long* packed = nullptr;
  int64_t* tokens_out = nullptr;
  int64_t tokens_stride = 0;
  unsigned int* done_counter = nullptr;
  const float* x_scales = nullptr;   // KIND_FP8
  const float* w_scale = nullptr;    // KIND_FP8
};

template <int EPI, int MT, int KIND>
static int launch_gemm_pass(const void* x, const void* w, int64_t m_rows, int64_t w_rows, int64_t K,
                            int64_t ldx, int64_t ldw, int64_t n_valid, __nv_bfloat16* out,
                            int64_t ldo, const ArgmaxExtra& ax, cudaStream_t s) {
  // function attributes are per device: one flag per device ordinal (one process may drive several)
  static std::atomic<bool> attr_set[kMaxDevices];
  int rc = ensure_max_smem(gemm_tn_sm100<EPI, MT, KIND>, attr_set, kSmemLimit);
  if (rc != AB200_OK) return rc;
  const int sms = sm_count();
  int bn = pick_block_n(n_valid, sms, kMaxBlockN / MT);
  const int forced_bn = EPI == EPI_STORE ? tune().gate_block_n : tune().head_block_n;
  // debug sweeps only (UMMA N must be a multiple of 16 for M = 128, of 8 for M = 64)
  if (forced_bn >= 16 && forced_bn <= kMaxBlockN / MT &&
      forced_bn % ((MT == 1 && m_rows <= 64 && tune().umma_m64) ? 8 : 16) == 0)
    bn = forced_bn;
  const int num_n_tiles = static_cast<int>((n_valid + bn - 1) / bn);
  const int box_m = MT == 1 ? static_cast<int>((m_rows + 7) / 8 * 8) : kTileM;
  // weight bytes in flight per SM = stages x tile: with a 7-stage ring the 8B gate GEMM was bound by
  // (bytes in flight) / (HBM latency) = 56-61 GB/s per SM, not by HBM - the activation slot is
  // therefore sized to the rows the batch really has
  // (the MMA reads all M rows of the slot, valid or not: M = 64 -> 8 KiB, M = 128 -> 16 KiB)
  const bool m64 = MT == 1 && m_rows <= 64 && tune().umma_m64;
  constexpr int esize = KIND == KIND_BF16 ? 2 : 1;
  constexpr int k_elems = kRowBytes / esize;
  // Two k-blocks per TMA request (3-D maps): the per-SM rate of TMA requests, not HBM, bounded the GEMMs with
  // narrow tiles - time per k-block was ~0.3 us whatever the box size (profiles/r2_gemm_tune.md).  Always for
  // the store-epilogue (gate / proj) GEMM: 8B gate 5.9 -> 6.1 TB/s while streaming, GPT-OSS gate (K = 2880) 4.6 ->
  // 6.3 TB/s; for the vocab head only with the 8 KiB activation slab (<= 64 rows: 3 stages of 60 KB beat 7 of
  // 30 KB, 6.17 -> 6.30 TB/s) - with 16 KiB slabs only 2 stages would fit and it loses.
  int kbs = 1;
  const int want = EPI == EPI_STORE ? tune().gate_kbs : (m64 ? tune().head_kbs : 1);
  if (MT == 1 && want == 2 && K % k_elems == 0 && K >= 2 * k_elems) kbs = 2;
  const uint32_t a_slab = m64 ? 64u * 128u : MT * kABytes;      // one k-block of activations (rows the MMA reads)
  const uint32_t a_slot = a_slab * kbs;
  const uint32_t stage_bytes = a_slot + bn * 128 * kbs;
  int stages = (kSmemLimit - 1024 /*align slack*/ - 256 /*barriers*/) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (tune().max_stages >= 2 && stages > tune().max_stages) stages = tune().max_stages;
  AB200_REQUIRE(stages >= 2, AB200_E_INVALID, "bf16 GEMM: not enough shared memory");

  CUtensorMap tmx, tmw;
  rc = get_tmap(&tmx, x, m_rows, K, ldx, box_m, esize, kbs);
  if (rc != AB200_OK) return rc;
  rc = get_tmap(&tmw, w, w_rows, K, ldw, bn, esize, kbs);
  if (rc != AB200_OK) return rc;

  GemmParams p;
  p.m_rows = static_cast<int32_t>(m_rows);
  p.n_valid = static_cast<int32_t>(n_valid);
  p.block_n = bn;
  p.num_n_tiles = num_n_tiles;
  p.num_k_blocks = static_cast<int32_t>(((K + k_elems - 1) / k_elems + kbs - 1) / kbs);   // pipeline steps per tile
  p.kbs = kbs;
  p.a_pitch = static_cast<uint32_t>(box_m) * 128u;
  p.stages = stages;
  p.l2_prefetch_kb = pdl_enabled() ? (EPI == EPI_STORE ? tune().gate_l2_prefetch : tune().l2_prefetch_kb) : 0;
  p.late_trigger = gemm_late_trigger();
  p.pre_stages = tune().pre_stages < 0 ? stages : tune().pre_stages;
  if (p.pre_stages > stages) p.pre_stages = stages;
  if (p.pre_stages < 0) p.pre_stages = 0;
  p.tx_bytes = (static_cast<uint32_t>(MT * box_m) * 128u + static_cast<uint32_t>(bn) * 128u) * kbs;
  // instruction descriptor: D=f32 [4,6), A/B format [7,10)/[10,13) (kind::f16: 1 = bf16;
  // kind::f8f6f4: 0 = e4m3), K-major A/B, N>>3 [17,23), M>>4 [24,29)
  const uint32_t fmt = KIND == KIND_BF16 ? 1u : 0u;
  // <= 64 rows: UMMA M = 64 halves the tensor-pipe work (and power) of this HBM-bound kernel
  p.m64 = m64 ? 1 : 0;
  p.mh = 1;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(bn >> 3) << 17) |
            (static_cast<uint32_t>((p.m64 ? 64 : kTileM) >> 4) << 24);
  p.x_scales = ax.x_scales;
  p.w_scale = ax.w_scale;
  p.a_slot = a_slot;
  p.out = out;
  p.ldo = ldo;
  p.packed = ax.packed;
  p.tokens_out = ax.tokens_out;
  p.tokens_stride = ax.tokens_stride;
  p.done_counter = ax.done_counter;
  p.trace = next_trace_slot();
  const int grid = num_n_tiles < sms ? num_n_tiles : sms;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024 + 256;
  LaunchCfg cfg(dim3(grid), dim3(kGemmThreads), smem, s);
  AB200_CUDA_OK(cudaLaunchKernelEx(&cfg.cfg, gemm_tn_sm100<EPI, MT, KIND>, tmx, tmw, p));
  AB200_CHECK_LAUNCH("gemm_tn_sm100");
  return AB200_OK;
}


#ifdef AB200_DEBUG
// How many 4-CTA clusters of the pair kernel the device can run at once (one CTA per SM): asked once per device.
template <int EPI>
static int max_clusters_of_4() {
  static std::atomic<int> cached[kMaxDevices];   // 0 = not asked yet, -1 = none
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  int v = cached[dev].load(std::memory_order_acquire);
  if (v != 0) return v < 0 ? 0 : v;
  static std::atomic<bool> attr_set[kMaxDevices];
  if (ensure_max_smem(gemm_tn_sm100_4cta<EPI>, attr_set, kSmemLimit) != AB200_OK) return 0;
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(4 * 64);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kSmemLimit;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_tn_sm100_4cta<EPI>, &cfg) != cudaSuccess) {
    (void)cudaGetLastError();
    n = 0;
  }
  cached[dev].store(n > 0 ? n : -1, std::memory_order_release);
  return n;
}
#endif

// 129..256 rows: the CTA-pair kernel (see gemm_tn_sm100_2cta)
template <int EPI>
static int launch_gemm_pass_2cta(const void* x, const void* w, int64_t m_rows, int64_t w_rows, int64_t K,
                                 int64_t ldx, int64_t ldw, int64_t n_valid, __nv_bfloat16* out, int64_t ldo,
                                 const ArgmaxExtra& ax, cudaStream_t s) {
  const int mh = m_rows > 2 * kTileM ? 2 : 1;     // 257..512 rows: two 256-row m-blocks per weight tile
  static std::atomic<bool> attr_set[kMaxDevices];
  int rc = ensure_max_smem(gemm_tn_sm100_2cta<EPI>, attr_set, kSmemLimit);
  if (rc != AB200_OK) return rc;
#ifdef AB200_DEBUG
  // experiment: 4-CTA clusters (two pairs sharing the activation tile by multicast), as many as the device
  // co-schedules
  const int clusters4 = tune().gemm_4cta ? max_clusters_of_4<EPI>() : 0;
  const bool cl4 = clusters4 * 4 >= sm_count() * 7 / 8;
#else
  constexpr int clusters4 = 0;
  constexpr bool cl4 = false;
#endif
  const int pairs = cl4 ? clusters4 : sm_count() / 2;   // units that own weight tiles (cl4: tile PAIRS)
  // tile width: multiple of 32 (each CTA holds half of it), balanced over the CTA pairs
  int bn = 32;
  double best_cost = 1e30;
  for (int c = 32; c <= 256; c += 32) {
    int64_t tiles = (n_valid + c - 1) / c;
    if (cl4) tiles = (tiles + 1) / 2;
    const int64_t waves = (tiles + pairs - 1) / pairs;
    const double cost = static_cast<double>(waves) * (c + 16.0);
    if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && c > bn)) {
      best_cost = cost;
      bn = c;
    }
  }
  const int forced_bn = EPI == EPI_STORE ? tune().gate_block_n : tune().head_block_n;   // debug sweeps only
  if (forced_bn >= 32 && forced_bn <= 256 && forced_bn % 32 == 0) bn = forced_bn;
  const int num_n_tiles = static_cast<int>((n_valid + bn - 1) / bn);
  // two k-blocks per TMA request where it pays (tools/gemm_tune.py pairk, profiles/r2_gemm_tune.md s.8)
  int kbs = 1;
  int want_kbs = EPI == EPI_STORE ? tune().pair_gate_kbs : tune().pair_head_kbs;
  // 0 = auto: two k-blocks per request only for narrow tiles (each CTA's W slab of a k-block < 12 KB), where the fixed
  // cost per request still shows (GPT-OSS gate at 256 / 512 rows: -10 %); wide tiles are level or 3-4 % better with one
  if (want_kbs == 0) want_kbs = bn < 192 ? 2 : 1;
  if (!cl4 && want_kbs >= 2 && want_kbs <= 4 && K % kBlockK == 0 && K >= want_kbs * kBlockK) kbs = want_kbs;
  while (kbs > 1 && (kSmemLimit - 1024 - 256) / ((kABytes + (bn / 2) * 128) * kbs) < 2) --kbs;   // keep >= 2 stages
  const uint32_t stage_bytes = (kABytes + (bn / 2) * 128) * kbs;
  int stages = (kSmemLimit - 1024 - 256) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (tune().max_stages >= 2 && stages > tune().max_stages) stages = tune().max_stages;
  AB200_REQUIRE(stages >= 2, AB200_E_INVALID, "bf16 GEMM: not enough shared memory");
  CUtensorMap tmx, tmw;
  rc = get_tmap(&tmx, x, m_rows, K, ldx, cl4 ? kTileM / 2 : kTileM, 2, kbs);
  if (rc != AB200_OK) return rc;
  rc = get_tmap(&tmw, w, w_rows, K, ldw, bn / 2, 2, kbs);
  if (rc != AB200_OK) return rc;
  GemmParams p;
  p.m_rows = static_cast<int32_t>(m_rows);
  p.n_valid = static_cast<int32_t>(n_valid);
  p.block_n = bn;
  p.num_n_tiles = num_n_tiles;
  p.num_k_blocks = static_cast<int32_t>(((K + kBlockK - 1) / kBlockK + kbs - 1) / kbs);   // pipeline steps per tile
  p.stages = stages;
  p.l2_prefetch_kb = 0;
#ifdef AB200_DEBUG
  p.dbg_skip = tune().dbg_skip;
#endif
  p.late_trigger = gemm_late_trigger();
  p.pre_stages = tune().pre_stages < 0 ? stages : tune().pre_stages;
  if (p.pre_stages > stages) p.pre_stages = stages;
  p.m64 = 0;
  p.kbs = kbs;
  p.a_pitch = 0;
  p.mh = mh;
  p.tx_bytes = 2u * (static_cast<uint32_t>(kABytes) + static_cast<uint32_t>(bn / 2) * 128u) * kbs;   // both CTAs
  // instruction descriptor: D = f32, A/B = bf16, K-major, N >> 3, M = 256 >> 4
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | (16u << 24);
  p.x_scales = nullptr;
  p.w_scale = nullptr;
  p.out = out;
  p.ldo = ldo;
  p.packed = ax.packed;
  p.tokens_out = ax.tokens_out;
  p.tokens_stride = ax.tokens_stride;
  p.done_counter = ax.done_counter;
  p.trace = next_trace_slot();
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024 + 256;
#ifdef AB200_DEBUG
  if (cl4) {
    const int work = (num_n_tiles + 1) / 2;
    const int use = work < pairs ? work : pairs;
    LaunchCfg cfg(dim3(4 * use), dim3(kGemmThreads), smem, s);
    AB200_CUDA_OK(cudaLaunchKernelEx(&cfg.cfg, gemm_tn_sm100_4cta<EPI>, tmx, tmw, p));
    AB200_CHECK_LAUNCH("gemm_tn_sm100_4cta");
    return AB200_OK;
  }
#endif
  const int use_pairs = num_n_tiles < pairs ? num_n_tiles : pairs;
  LaunchCfg cfg(dim3(2 * use_pairs), dim3(kGemmThreads), smem, s);
  AB200_CUDA_OK(cudaLaunchKernelEx(&cfg.cfg, gemm_tn_sm100_2cta<EPI>, tmx, tmw, p));
  AB200_CHECK_LAUNCH("gemm_tn_sm100_2cta");
  return AB200_OK;
}

// Rows that share ONE launch.  Store-epilogue bf16 GEMM (gate / proj): 512 - the CTA-pair kernel runs two 256-row
// m-blocks per weight tile (8B gate at B = 512: 82 -> 68 us).  Arg-max head: 256 per launch - it is tensor-bound
// there (1.3 PFLOP/s = 0.92 of the sustained cuBLAS peak) and two launches were 2 % faster than the merged one.
// fp8: 256 (the MT = 4 single-CTA shape is ingest-bound).
static int rows_per_launch(int kind, int epi) {
  const int r = gemm_rows_per_pass();
  if (kind == KIND_BF16 && epi == EPI_STORE && r >= 256 && tune().gemm_2cta && tune().gate_mh2) return 512;
  return r < 256 ? r : 256;
}

template <int EPI, int KIND = KIND_BF16>
static int launch_gemm_bf16(const void* x, const void* w, int64_t M, int64_t w_rows, int64_t K,
                            int64_t ldx, int64_t ldw, int64_t n_valid, __nv_bfloat16* out,
                            int64_t ldo, const ArgmaxExtra& ax, cudaStream_t s) {
  constexpr int kAlign = KIND == KIND_BF16 ? 8 : 16;   // 16-byte rows / strides
  constexpr int esize = KIND == KIND_BF16 ? 2 : 1;
  AB200_REQUIRE(K % kAlign == 0 && ldx % kAlign == 0 && ldw % kAlign == 0 && aligned(x, 16) && aligned(w, 16),
                AB200_E_INVALID,
                "tcgen05 GEMM: K, ldx, ldw must be multiples of %d and x/w 16-byte aligned", kAlign);
  AB200_REQUIRE(KIND == KIND_BF16 || (ax.x_scales != nullptr && ax.w_scale != nullptr), AB200_E_INVALID,
                "fp8 GEMM: scales required");
  AB200_REQUIRE(M > 0 && w_rows > 0 && K > 0 && n_valid > 0 && n_valid <= w_rows &&
                    w_rows < (1LL << 31) && K < (1LL << 31),
                AB200_E_INVALID, "bf16 GEMM: bad sizes");
  const int per_launch = rows_per_launch(KIND, EPI);
  AB200_REQUIRE(ax.tokens_out == nullptr || (M <= per_launch && ax.done_counter != nullptr),
                AB200_E_INVALID, "head_argmax: fused token output needs M <= %d and a done counter", per_launch);
  for (int64_t m0 = 0; m0 < M; m0 += per_launch) {
    const int64_t rows = M - m0 < per_launch ? M - m0 : per_launch;
    const char* xp = static_cast<const char*>(x) + m0 * ldx * esize;
    __nv_bfloat16* op = out ? out + m0 * ldo : nullptr;
    ArgmaxExtra a = ax;
    if (a.packed) a.packed += m0;
    if (a.x_scales) a.x_scales += m0;
    int rc;
    if (rows <= kTileM)
      rc = launch_gemm_pass<EPI, 1, KIND>(xp, w, rows, w_rows, K, ldx, ldw, n_valid, op, ldo, a, s);
    else if (rows <= 4 * kTileM && KIND == KIND_BF16 && tune().gemm_2cta)
      rc = launch_gemm_pass_2cta<EPI>(xp, w, rows, w_rows, K, ldx, ldw, n_valid, op, ldo, a, s);
    else if (rows <= 2 * kTileM)
      rc = launch_gemm_pass<EPI, 2, KIND>(xp, w, rows, w_rows, K, ldx, ldw, n_valid, op, ldo, a, s);
    else
      rc = launch_gemm_pass<EPI, 4, KIND>(xp, w, rows, w_rows, K, ldx, ldw, n_valid, op, ldo, a, s);
    if (rc != AB200_OK) return rc;
  }
  return AB200_OK;
}

// ab200_init(): set the shared-memory attribute of every tcgen05 instantiation on the current device
// (so that no first-use attribute call can land inside a stream capture) and resolve the driver's
// tensor-map encoder.
int init_gemm_kernels() {
  AB200_REQUIRE(get_encode_fn() != nullptr, AB200_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
#define AB200_INIT_ATTR(...)                                                                              \
  AB200_CUDA_OK(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit))
  AB200_INIT_ATTR(gemm_tn_sm100<EPI_STORE, 1, KIND_BF16>);
  AB200_INIT_ATTR(gemm_tn_sm100<EPI_ARGMAX, 1, KIND_BF16>);
  AB200_INIT_ATTR(gemm_tn_sm100<EPI_STORE, 1, KIND_FP8>);
  AB200_INIT_ATTR(gemm_tn_sm100<EPI_ARGMAX, 1, KIND_FP8>);
  AB200_INIT_ATTR(gemm_tn_sm100_2cta<EPI_STORE>);
  AB200_INIT_ATTR(gemm_tn_sm100_2cta<EPI_ARGMAX>);
#ifdef AB200_DEBUG
  AB200_INIT_ATTR(gemm_tn_sm100_4cta<EPI_STORE>);
  AB200_INIT_ATTR(gemm_tn_sm100_4cta<EPI_ARGMAX>);
  (void)max_clusters_of_4<EPI_STORE>();    // occupancy asked here, outside any capture
  (void)max_clusters_of_4<EPI_ARGMAX>();
#endif
#undef AB200_INIT_ATTR
  return AB200_OK;
}

#ifdef AB200_DEBUG
extern "C" AB200_API int32_t ab200_debug_max_clusters_of_4(void) { return max_clusters_of_4<EPI_ARGMAX>(); }
#endif

// ------------------------------------------------------------------------------------------
// exact-fp32 SIMT GEMM (parity mode) : out[M,N] = x[M,K] . w[N,K]^T
// ------------------------------------------------------------------------------------------
constexpr int kSBM = 64, kSBN = 64, kSBK = 16;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
gemm_tn_simt(const TIn* __restrict__ x, const TIn* __restrict__ w, TOut* __restrict__ out, int M,
             int N, int K, int64_t ldx, int64_t ldw, int64_t ldo) {
  __shared__ float xs[kSBK][kSBM + 4];
  __shared__ float wsm[kSBK][kSBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kSBM, n0 = blockIdx.x * kSBN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kSBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;       // 0..1023
      const int r = e >> 4, kk = e & 15; // row in tile, k in chunk
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      xs[kk][r] = (gm < M && gk < K) ? ST<TIn>::to_f(x[gm * ldx + gk]) : 0.f;
      wsm[kk][r] = (gn < N && gk < K) ? ST<TIn>::to_f(w[gn * ldw + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = wsm[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) out[gm * ldo + gn] = ST<TOut>::from_f(acc[i][j]);
    }
  }
}

template <typename TIn, typename TOut>
static int launch_simt(const void* x, const void* w, void* out, int64_t M, int64_t N, int64_t K,
                       int64_t ldx, int64_t ldw, int64_t ldo, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((N + kSBN - 1) / kSBN), static_cast<unsigned>((M + kSBM - 1) / kSBM));
  gemm_tn_simt<TIn, TOut><<<grid, 256, 0, s>>>(static_cast<const TIn*>(x), static_cast<const TIn*>(w),
                                               static_cast<TOut*>(out), (int)M, (int)N, (int)K, ldx,
                                               ldw, ldo);
  AB200_CHECK_LAUNCH("gemm_tn_simt");
  return AB200_OK;
}

// ------------------------------------------------------------------------------------------
// arg-max helpers
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const T* __restrict__ logits, int64_t N, int64_t ld,
                   unsigned long long* __restrict__ packed) {
  __shared__ unsigned long long sm[8];
  const T* rowp = logits + blockIdx.x * ld;
  float best = -INFINITY;
  uint32_t bi = 0xFFFFFFFFu;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float v = ST<T>::to_f(rowp[i]);
    if (v > best) { best = v; bi = static_cast<uint32_t>(i); }
  }
  unsigned long long pk = bi == 0xFFFFFFFFu ? 0ull : pack_arg(best, bi);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, pk, o);
    pk = other > pk ? other : pk;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = pk;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) pk = sm[i] > pk ? sm[i] : pk;
    packed[blockIdx.x] = pk;
  }
}

__global__ void argmax_unpack_kernel(const unsigned long long* __restrict__ packed,
                                     int64_t* __restrict__ tokens, int64_t count, int64_t stride) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < count) tokens[i * stride] = static_cast<int64_t>(unpack_idx(packed[i]));
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_linear(const void* x, const void* w, void* out, int64_t M, int64_t N,
                                int64_t K, int64_t ldx, int64_t ldw, int64_t ldo, int32_t dtype,
                                void* stream) {
  if (M == 0 || N == 0) return AB200_OK;
  AB200_REQUIRE(x && w && out && M > 0 && N > 0 && K > 0, AB200_E_INVALID, "linear: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == AB200_F32) return launch_simt<float, float>(x, w, out, M, N, K, ldx, ldw, ldo, s);
  AB200_REQUIRE(dtype == AB200_BF16, AB200_E_INVALID, "linear: dtype must be bf16 or f32");
  AB200_REQUIRE(ldo % 8 == 0 && aligned(out, 16), AB200_E_INVALID,
                "linear: ldo must be a multiple of 8 and out 16-byte aligned");
  return launch_gemm_bf16<EPI_STORE>(x, w, M, N, K, ldx, ldw, N, static_cast<__nv_bfloat16*>(out), ldo,
                                     ArgmaxExtra(), s);
}

extern "C" int32_t ab200_argmax_rows(const void* logits, int64_t M, int64_t N, int64_t ld,
                                     int32_t dtype, uint64_t* packed_out, void* stream) {
  if (M == 0) return AB200_OK;
  AB200_REQUIRE(logits && packed_out && M > 0 && N > 0, AB200_E_INVALID, "argmax_rows: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto* pk = reinterpret_cast<unsigned long long*>(packed_out);
  dim3 grid(static_cast<unsigned>(M));
  if (dtype == AB200_F32)
    argmax_rows_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(logits), N, ld, pk);
  else if (dtype == AB200_BF16)
    argmax_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), N, ld, pk);
  else
    AB200_REQUIRE(false, AB200_E_INVALID, "argmax_rows: dtype must be bf16 or f32");
  AB200_CHECK_LAUNCH("argmax_rows_kernel");
  return AB200_OK;
}

extern "C" int32_t ab200_head_argmax(const void* x, const void* w, int64_t M, int64_t w_rows,
                                     int64_t K, int64_t vocab, int64_t ldx, int64_t ldw,
                                     uint64_t* packed_out, int64_t* tokens_out,
                                     int64_t tokens_stride, uint32_t* done_counter, int32_t dtype,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (M == 0) return AB200_OK;
  AB200_REQUIRE(x && w && packed_out && M > 0 && K > 0 && vocab > 0 && vocab <= w_rows,
                AB200_E_INVALID, "head_argmax: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == AB200_F32) {
    const size_t need = static_cast<size_t>(M) * static_cast<size_t>(vocab) * sizeof(float);
    AB200_REQUIRE(workspace && workspace_bytes >= need, AB200_E_WORKSPACE,
                  "head_argmax(f32): workspace of %zu bytes required", need);
    int rc = launch_simt<float, float>(x, w, workspace, M, vocab, K, ldx, ldw, vocab, s);
    if (rc != AB200_OK) return rc;
    rc = ab200_argmax_rows(workspace, M, vocab, vocab, AB200_F32, packed_out, stream);
    if (rc != AB200_OK || tokens_out == nullptr) return rc;
    return ab200_argmax_unpack(packed_out, tokens_out, M, tokens_stride, stream);
  }
  AB200_REQUIRE(dtype == AB200_BF16, AB200_E_INVALID, "head_argmax: dtype must be bf16 or f32");
  ArgmaxExtra ax;
  ax.packed = reinterpret_cast<unsigned long long*>(packed_out);
  if (tokens_out != nullptr && M <= gemm_rows_per_pass() && done_counter != nullptr) {
    ax.tokens_out = tokens_out;
    ax.tokens_stride = tokens_stride;
    ax.done_counter = done_counter;
    return launch_gemm_bf16<EPI_ARGMAX>(x, w, M, w_rows, K, ldx, ldw, vocab, nullptr, 0, ax, s);
  }
  int rc = launch_gemm_bf16<EPI_ARGMAX>(x, w, M, w_rows, K, ldx, ldw, vocab, nullptr, 0, ax, s);
  if (rc != AB200_OK || tokens_out == nullptr) return rc;
  return ab200_argmax_unpack(packed_out, tokens_out, M, tokens_stride, stream);
}

// FP8 (e4m3 x e4m3 -> fp32) projections: out = bf16(x_scales[row] * (w_scale * acc)).
extern "C" int32_t ab200_linear_fp8(const void* xq, const float* x_scales, const void* wq,
                                    const float* w_scale, void* out, int64_t M, int64_t N, int64_t K,
                                    int64_t ldx, int64_t ldw, int64_t ldo, void* stream) {
  if (M == 0 || N == 0) return AB200_OK;
  AB200_REQUIRE(xq && x_scales && wq && w_scale && out && M > 0 && N > 0 && K > 0, AB200_E_INVALID,
                "linear_fp8: bad arguments");
  AB200_REQUIRE(ldo % 8 == 0 && aligned(out, 16), AB200_E_INVALID,
                "linear_fp8: ldo must be a multiple of 8 and out 16-byte aligned");
  ArgmaxExtra ax;
  ax.x_scales = x_scales;
  ax.w_scale = w_scale;
  return launch_gemm_bf16<EPI_STORE, KIND_FP8>(xq, wq, M, N, K, ldx, ldw, N,
                                               static_cast<__nv_bfloat16*>(out), ldo, ax,
                                               static_cast<cudaStream_t>(stream));
}

extern "C" int32_t ab200_head_argmax_fp8(const void* xq, const float* x_scales, const void* wq,
                                         const float* w_scale, int64_t M, int64_t w_rows, int64_t K,
                                         int64_t vocab, int64_t ldx, int64_t ldw, uint64_t* packed_out,
                                         int64_t* tokens_out, int64_t tokens_stride,
                                         uint32_t* done_counter, void* stream) {
  if (M == 0) return AB200_OK;
  AB200_REQUIRE(xq && x_scales && wq && w_scale && packed_out && M > 0 && K > 0 && vocab > 0 &&
                    vocab <= w_rows,
                AB200_E_INVALID, "head_argmax_fp8: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ArgmaxExtra ax;
  ax.packed = reinterpret_cast<unsigned long long*>(packed_out);
  ax.x_scales = x_scales;
  ax.w_scale = w_scale;
  if (tokens_out != nullptr && M <= rows_per_launch(KIND_FP8, EPI_ARGMAX) && done_counter != nullptr) {
    ax.tokens_out = tokens_out;
    ax.tokens_stride = tokens_stride;
    ax.done_counter = done_counter;
    return launch_gemm_bf16<EPI_ARGMAX, KIND_FP8>(xq, wq, M, w_rows, K, ldx, ldw, vocab, nullptr, 0, ax, s);
  }
  int rc = launch_gemm_bf16<EPI_ARGMAX, KIND_FP8>(xq, wq, M, w_rows, K, ldx, ldw, vocab, nullptr, 0, ax, s);
  if (rc != AB200_OK || tokens_out == nullptr) return rc;
  return ab200_argmax_unpack(packed_out, tokens_out, M, tokens_stride, stream);
}

// Tensor-parallel (vocab-sharded) head: turn a shard-local packed arg-max into a key that a plain
// signed 64-bit MAX all-reduce combines correctly - global token index, order-preserving for int64.
// Replaces the topk / pack-as-f64 / all_gather / argmax / gather sequence of
// arctic_speculator.py:733-744 by one kernel + one all_reduce(MAX).
__global__ void argmax_shard_key_kernel(unsigned long long* __restrict__ packed, int64_t count,
                                        uint32_t index_offset) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  unsigned long long p = packed[i];
  if (p != 0ull) {   // 0 = this shard found nothing: stays the smallest key
    const uint32_t gidx = unpack_idx(p) + index_offset;
    p = (p & 0xFFFFFFFF00000000ull) | static_cast<unsigned long long>(0xFFFFFFFFu - gidx);
  }
  packed[i] = p ^ 0x8000000000000000ull;   // unsigned order -> signed order
}

extern "C" int32_t ab200_argmax_shard_key(uint64_t* packed, int64_t count, int64_t index_offset, void* stream) {
  if (count == 0) return AB200_OK;
  AB200_REQUIRE(packed && count > 0 && index_offset >= 0 && index_offset < (1LL << 31), AB200_E_INVALID,
                "argmax_shard_key: bad arguments");
  argmax_shard_key_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<unsigned long long*>(packed), count, static_cast<uint32_t>(index_offset));
  AB200_CHECK_LAUNCH("argmax_shard_key_kernel");
  return AB200_OK;
}

extern "C" int32_t ab200_argmax_unpack(const uint64_t* packed, int64_t* tokens_out, int64_t count,
                                       int64_t out_stride, void* stream) {
  if (count == 0) return AB200_OK;
  AB200_REQUIRE(packed && tokens_out && count > 0, AB200_E_INVALID, "argmax_unpack: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  argmax_unpack_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const unsigned long long*>(packed), tokens_out, count, out_stride);
  AB200_CHECK_LAUNCH("argmax_unpack_kernel");
  return AB200_OK;
}

// Tile width the tcgen05 GEMM would choose for N output features on `sms` SMs (bench reporting).
extern "C" int32_t ab200_gemm_block_n(int64_t N, int32_t sms) {
  return pick_block_n(N, sms > 0 ? sms : sm_count(), kMaxBlockN);
}

// bf16-input SIMT GEMM with fp32 accumulation, bf16 output: test/debug comparator for the tcgen05
// kernel (not on the product path; exported for tests only).
extern "C" int32_t ab200_linear_simt_bf16(const void* x, const void* w, void* out, int64_t M,
                                          int64_t N, int64_t K, int64_t ldx, int64_t ldw,
                                          int64_t ldo, void* stream) {
  if (M == 0 || N == 0) return AB200_OK;
  return launch_simt<__nv_bfloat16, __nv_bfloat16>(x, w, out, M, N, K, ldx, ldw, ldo,
                                                    static_cast<cudaStream_t>(stream));
}
