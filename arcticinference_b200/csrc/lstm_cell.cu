// lstm_cell.cu - the Arctic LSTM speculator's fused elementwise work for sm_100a (hot path A):
//   * rms_scale_kernel      : ln0(h) / sqrt(2)                 (arctic_speculator.py:88-91, 656-657)
//   * sum_lstm_cell_kernel  : embedding gather + alpha-sum + 3 sigmoids + 2 RMS-LayerNorms with
//                             affine + 2 erf-GELUs + LSTM gating  (arctic_speculator.py:667-689)
// One CTA per batch row, 16-byte vector loads, values held in registers across the two row
// reductions (warp-shuffle + one smem hop).  ~30 ATen launches of the reference become one.
//
// Rounding: with EMU (bf16/f16 storage) every value PyTorch eager would have materialised as a
// tensor is rounded to the storage dtype at that point (the list is restated and
// pinned in oracle/speculator_ref.py::emulated_cell).  The arithmetic between roundings follows
// ATen's CUDA functors: fma for add-with-alpha, 1/(1+exp(-x)), x*0.5*(1+erf(x*M_SQRT1_2)),
// mean = sum * (1/N), rsqrtf.
//
// Roofline: HBM bytes per row = e * D * (4 gates + 1 emb row + 1 c_prev + 2 outputs); at the
// BASELINE batch sizes (64..512 rows) the kernel is launch-latency bound.
#include <cstdlib>
#include <type_traits>

#include "bf16_exact.cuh"
#include "common.cuh"

namespace ab200 {

template <typename T, bool EMU>
__device__ __forceinline__ float R(float v) {
  if constexpr (EMU) return ST<T>::round(v);
  else return v;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) {
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}

template <typename T, bool EMU, int ITEMS>
__global__ void __launch_bounds__(1024)
rms_scale_kernel(const T* __restrict__ x, T* __restrict__ y, int32_t cols, int64_t ldx,
                 int64_t ldy, float eps, float post_scale, unsigned long long* trace) {
  constexpr int VN = Vec16<T>::N;
  __shared__ float scratch[33];
  const int64_t row = blockIdx.x;
  const int nvec = cols / VN;
  const T* xr = x + row * ldx;
  T* yr = y + row * ldy;
  float v[ITEMS][VN];
  float ss = 0.f;
  if (threadIdx.x == 0) trace_stamp(trace, 0);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) trace_stamp(trace, 1);
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      Vec16<T> in;
      in.u = *reinterpret_cast<const uint4*>(xr + vi * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        v[it][j] = ST<T>::to_f(in.e[j]);
        ss += R<T, EMU>(v[it][j] * v[it][j]);
      }
    }
  }
  ss = block_sum(ss, scratch);
  const float mean = R<T, EMU>(ss * (1.0f / static_cast<float>(cols)));
  const float inv = R<T, EMU>(rsqrtf(R<T, EMU>(mean + eps)));
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      Vec16<T> out;
#pragma unroll
      for (int j = 0; j < VN; ++j)
        out.e[j] = ST<T>::from_f(R<T, EMU>(v[it][j] * inv) * post_scale);
      *reinterpret_cast<uint4*>(yr + vi * VN) = out.u;
    }
  }
  if (threadIdx.x == 0) trace_stamp(trace, 2);
}

template <typename T, bool EMU, int ITEMS>
__global__ void __launch_bounds__(1024)
sum_lstm_cell_kernel(const T* __restrict__ gates, int64_t ldg, const T* __restrict__ emb,
                     const int64_t* __restrict__ tokens,
                     const unsigned long long* __restrict__ packed,
                     const T* cell_prev,  // may alias cell_out
                     const T* __restrict__ wc, const T* __restrict__ bc,
                     const T* __restrict__ ws, const T* __restrict__ bs,
                     T* __restrict__ state_out, T* cell_out, int32_t D, int64_t vocab,
                     float alpha, float eps, unsigned long long* __restrict__ clear_packed,
                     unsigned int* __restrict__ clear_counter, unsigned long long* trace,
                     uint8_t* __restrict__ state_q, float* __restrict__ state_scale) {
  constexpr int VN = Vec16<T>::N;
  __shared__ float scratch[33];
  const int64_t row = blockIdx.x;
  const int nvec = D / VN;

  if (threadIdx.x == 0) trace_stamp(trace, 0);
  pdl_launch_dependents();   // the head GEMM may start prefetching its weights now
  pdl_wait();                // gate pre-activations / previous arg-max are complete and visible
  if (threadIdx.x == 0) {
    trace_stamp(trace, 1);
    // prepare this step's arg-max slot and completion counter for the head kernel that follows
    if (clear_packed) clear_packed[row] = 0ull;
    if (clear_counter && row == 0) *clear_counter = 0u;
  }
  int64_t tok = packed ? static_cast<int64_t>(unpack_idx(packed[row])) : tokens[row];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);  // never read outside the table

  const T* g = gates + row * ldg;
  const T* z = emb + tok * static_cast<int64_t>(D);
  const float inv_d = 1.0f / static_cast<float>(D);

  float f[ITEMS][VN], ig[ITEMS][VN], og[ITEMS][VN], cc[ITEMS][VN], c[ITEMS][VN];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      Vec16<T> gf, gi, go, gc, ze, cp;
      gf.u = ld_stream_16(g + 0 * D + vi * VN);
      gi.u = ld_stream_16(g + 1 * D + vi * VN);
      go.u = ld_stream_16(g + 2 * D + vi * VN);
      gc.u = ld_stream_16(g + 3 * D + vi * VN);
      ze.u = *reinterpret_cast<const uint4*>(z + vi * VN);
      if (cell_prev) cp.u = *reinterpret_cast<const uint4*>(cell_prev + row * D + vi * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const float zz = ST<T>::to_f(ze.e[j]);
        // torch.add(states, z.repeat(1,1,4), alpha): one rounding of fma(alpha, z, s)
        f[it][j] = R<T, EMU>(sigmoid_f(R<T, EMU>(fmaf(alpha, zz, ST<T>::to_f(gf.e[j])))));
        ig[it][j] = R<T, EMU>(sigmoid_f(R<T, EMU>(fmaf(alpha, zz, ST<T>::to_f(gi.e[j])))));
        og[it][j] = R<T, EMU>(sigmoid_f(R<T, EMU>(fmaf(alpha, zz, ST<T>::to_f(go.e[j])))));
        cc[it][j] = R<T, EMU>(fmaf(alpha, zz, ST<T>::to_f(gc.e[j])));
        c[it][j] = cell_prev ? ST<T>::to_f(cp.e[j]) : 0.f;
        ss += R<T, EMU>(cc[it][j] * cc[it][j]);
      }
    }
  }
  // cell_ln (RMS, affine) -> GELU -> * input gate ; c = c_prev * f + candidate
  ss = block_sum(ss, scratch);
  {
    const float mean = R<T, EMU>(ss * inv_d);
    const float inv = R<T, EMU>(rsqrtf(R<T, EMU>(mean + eps)));
    ss = 0.f;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        Vec16<T> w, b;
        w.u = *reinterpret_cast<const uint4*>(wc + vi * VN);
        b.u = *reinterpret_cast<const uint4*>(bc + vi * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          float y = R<T, EMU>(cc[it][j] * inv);
          y = R<T, EMU>(ST<T>::to_f(w.e[j]) * y);
          y = R<T, EMU>(y + ST<T>::to_f(b.e[j]));
          y = R<T, EMU>(gelu_erf_f(y));
          y = R<T, EMU>(y * ig[it][j]);
          float cn = R<T, EMU>(c[it][j] * f[it][j]);
          cn = R<T, EMU>(cn + y);
          c[it][j] = cn;
          ss += R<T, EMU>(cn * cn);
        }
      }
    }
  }
  // state_ln (RMS, affine) -> GELU -> * output gate
  ss = block_sum(ss, scratch);
  float amax = 0.f;
  {
    const float mean = R<T, EMU>(ss * inv_d);
    const float inv = R<T, EMU>(rsqrtf(R<T, EMU>(mean + eps)));
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        Vec16<T> w, b, so, co;
        w.u = *reinterpret_cast<const uint4*>(ws + vi * VN);
        b.u = *reinterpret_cast<const uint4*>(bs + vi * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          float y = R<T, EMU>(c[it][j] * inv);
          y = R<T, EMU>(ST<T>::to_f(w.e[j]) * y);
          y = R<T, EMU>(y + ST<T>::to_f(b.e[j]));
          y = R<T, EMU>(gelu_erf_f(y));
          y = R<T, EMU>(y * og[it][j]);
          so.e[j] = ST<T>::from_f(y);
          co.e[j] = ST<T>::from_f(c[it][j]);
          f[it][j] = ST<T>::to_f(so.e[j]);   // the stored (rounded) state, for the fp8 copy below
          amax = fmaxf(amax, fabsf(f[it][j]));
        }
        *reinterpret_cast<uint4*>(state_out + row * D + vi * VN) = so.u;
        *reinterpret_cast<uint4*>(cell_out + row * D + vi * VN) = co.u;
      }
    }
  }
  if (state_q != nullptr) {
    // Dynamic per-token FP8 copy of the state for the quantised head (what the reference's qhead
    // does inside Fp8LinearOp.apply): s = max(amax/448, 1/(448*512)), q = sat_e4m3(x / s).
    amax = block_max(amax, scratch);
    const float scale = fmaxf(amax / kFp8Max, kFp8MinScale);
    if (threadIdx.x == 0) state_scale[row] = scale;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        uint8_t qv[VN];
#pragma unroll
        for (int j = 0; j < VN; ++j) qv[j] = to_e4m3(f[it][j] / scale);
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


// ------------------------------------------------------------------------------------------
// bf16 + eager-rounding emulation, the production configuration: same results as
// sum_lstm_cell_kernel<__nv_bfloat16, true, ITEMS> bit for bit (tests/test_cell_gpu.py), at about
// a third of the instructions - see bf16_exact.cuh for why each shortcut is exact.  Values that
// live across the two row reductions are kept as packed bf16x2 (they ARE bf16 in eager mode).
// ------------------------------------------------------------------------------------------
// per-device copy of the GELU table of bf16_exact.cuh (8 KB, L2-resident)
__device__ __align__(16) uint16_t g_gelu_tab[kGeluTabEntries];
__global__ void gelu_lut_init_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kGeluTabEntries) g_gelu_tab[i] = gelu_tab_entry(i);
}

// FULL: D == 8 * ITEMS * blockDim.x (no ragged tail: bounds checks and zero-fills compile away)
// EXACT: bit-identical to eager PyTorch (guard-free fast sigmoid with a range test per 16-byte
// vector, GELU by table); !EXACT: "faithful" - the same rounding points with the unguarded fast
// sigmoid and polynomial GELU (differs from ATen by one bf16 ulp on the 0.13 % of GELU inputs <= -3).
template <int ITEMS, int MAXT, int MINB, bool FULL, bool EXACT>
__global__ void __launch_bounds__(MAXT, MINB)
sum_lstm_cell_bf16x2_kernel(const __nv_bfloat16* __restrict__ gates, int64_t ldg,
                            const __nv_bfloat16* __restrict__ emb,
                            const int64_t* __restrict__ tokens,
                            const unsigned long long* __restrict__ packed,
                            const __nv_bfloat16* cell_prev,  // may alias cell_out
                            const __nv_bfloat16* __restrict__ wc, const __nv_bfloat16* __restrict__ bc,
                            const __nv_bfloat16* __restrict__ ws, const __nv_bfloat16* __restrict__ bs,
                            __nv_bfloat16* __restrict__ state_out, __nv_bfloat16* cell_out, int32_t D,
                            int64_t vocab, float alpha, float eps, float inv_d,
                            unsigned long long* __restrict__ clear_packed,
                            unsigned int* __restrict__ clear_counter, unsigned long long* trace,
                            uint8_t* __restrict__ state_q, float* __restrict__ state_scale, int early) {
  using T = __nv_bfloat16;
  __shared__ float scratch[33], scratch2[32];
  __shared__ __align__(16) uint16_t gelu_tab[EXACT ? kGeluTabEntries : 8];
  const int64_t row = blockIdx.x;
  const int nvec = D / 8;

  if (threadIdx.x == 0) trace_stamp(trace, 0);
  pdl_launch_dependents();
  // the table fill is independent of the previous kernel: it runs before the dependency wait
  if constexpr (EXACT) gelu_tab_fill(gelu_tab, g_gelu_tab);
  uint32_t lut_s = static_cast<uint32_t>(__cvta_generic_to_shared(gelu_tab));
  asm volatile("mov.u32 %0, %0;" : "+r"(lut_s));   // opaque: keep it in a register, do not rematerialise
  // Latency variant (batches of at most ~one CTA per SM): the LayerNorm vectors are fetched up front
  // instead of after each row reduction, so that their L2 latency is off the critical path.
  constexpr bool PRE = MAXT == 512 && MINB == 1;
  const T* g = gates + row * ldg;
  int64_t tok = 0;
  uint4 lnw1[PRE ? ITEMS : 1], lnb1[PRE ? ITEMS : 1], lnw2[PRE ? ITEMS : 1], lnb2[PRE ? ITEMS : 1];
  uint4 zeP[PRE ? ITEMS : 1], cpP[PRE ? ITEMS : 1];
  // Everything but the gate pre-activations: token -> embedding row, previous cell state, LayerNorm
  // vectors.  With `early` (set by the draft loop, where the kernel before this one is the gate GEMM
  // and all of these were complete before THAT kernel started) it runs before the dependency wait,
  // so that only the L2-hot gate rows are still to be fetched once the GEMM has finished.
  auto fetch_side_inputs = [&]() {
    tok = packed ? static_cast<int64_t>(unpack_idx(packed[row])) : tokens[row];
    if constexpr (PRE) {
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int vi = threadIdx.x + it * blockDim.x;
        if (FULL || vi < nvec) {
          cpP[it] = cell_prev ? ld_ordered_16(cell_prev + row * D + vi * 8) : make_uint4(0u, 0u, 0u, 0u);
          lnw1[it] = ld_ordered_16(wc + vi * 8);
          lnb1[it] = ld_ordered_16(bc + vi * 8);
          lnw2[it] = ld_ordered_16(ws + vi * 8);
          lnb2[it] = ld_ordered_16(bs + vi * 8);
        }
      }
      int64_t tk = tok;
      if (static_cast<uint64_t>(tk) >= static_cast<uint64_t>(vocab)) tk = tk < 0 ? 0 : vocab - 1;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int vi = threadIdx.x + it * blockDim.x;
        if (FULL || vi < nvec) zeP[it] = ld_stream_16(emb + tk * static_cast<int64_t>(D) + vi * 8);
      }
    }
  };
  const bool early_side = PRE && early;
  if (early_side) fetch_side_inputs();
  pdl_wait();
  if (threadIdx.x == 0) {
    trace_stamp(trace, 1);
    if (clear_packed) clear_packed[row] = 0ull;
    if (clear_counter && row == 0) *clear_counter = 0u;
  }

  uint32_t F[ITEMS][4], I[ITEMS][4], O[ITEMS][4], CC[ITEMS][4], C[ITEMS][4];
  f2 ss2 = bcast2(0.f);   // sum of squares, even / odd elements (one FADD2 per pair)
  uint4 gfv[PRE ? ITEMS : 1], giv[PRE ? ITEMS : 1], gov[PRE ? ITEMS : 1], gcv[PRE ? ITEMS : 1];
  if constexpr (PRE) {
    // gate rows first (they are what the wait was for), then - unless already in flight - the rest
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (FULL || vi < nvec) {
        gfv[it] = ld_stream_16(g + 0 * D + vi * 8);
        giv[it] = ld_stream_16(g + 1 * D + vi * 8);
        gov[it] = ld_stream_16(g + 2 * D + vi * 8);
        gcv[it] = ld_stream_16(g + 3 * D + vi * 8);
      }
    }
    if (!early_side) fetch_side_inputs();
  } else {
    fetch_side_inputs();
  }
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (FULL || vi < nvec) {
      uint4 gf, gi, go, gc, cp, ze;
      if constexpr (PRE) {
        gf = gfv[it], gi = giv[it], go = gov[it], gc = gcv[it], cp = cpP[it], ze = zeP[it];
      } else {
        gf = ld_stream_16(g + 0 * D + vi * 8);
        gi = ld_stream_16(g + 1 * D + vi * 8);
        go = ld_stream_16(g + 2 * D + vi * 8);
        gc = ld_stream_16(g + 3 * D + vi * 8);
        cp = make_uint4(0u, 0u, 0u, 0u);
        if (cell_prev) cp = *reinterpret_cast<const uint4*>(cell_prev + row * D + vi * 8);
        // everything that does not depend on the token is in flight before the token is consumed
        // (the embedding row is the only load behind the token's L2 round trip)
        asm volatile("" : "+l"(tok));
        int64_t tk = tok;
        if (static_cast<uint64_t>(tk) >= static_cast<uint64_t>(vocab)) tk = tk < 0 ? 0 : vocab - 1;
        ze = ld_stream_16(emb + tk * static_cast<int64_t>(D) + vi * 8);
      }
      const uint32_t gfa[4] = {gf.x, gf.y, gf.z, gf.w}, gia[4] = {gi.x, gi.y, gi.z, gi.w};
      const uint32_t goa[4] = {go.x, go.y, go.z, go.w}, gca[4] = {gc.x, gc.y, gc.z, gc.w};
      const uint32_t zea[4] = {ze.x, ze.y, ze.z, ze.w}, cpa[4] = {cp.x, cp.y, cp.z, cp.w};
      uint32_t af[4], ai[4], ao[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const f2 zz = unpack2(zea[p]), al = bcast2(alpha);
        // torch.add(states, z.repeat(1,1,4), alpha): one rounding of fma(alpha, z, s)
        const f2 sf = fma2(al, zz, unpack2(gfa[p])), si = fma2(al, zz, unpack2(gia[p]));
        const f2 so = fma2(al, zz, unpack2(goa[p])), sc = fma2(al, zz, unpack2(gca[p]));
        af[p] = pack_rn(sf.a, sf.b);
        ai[p] = pack_rn(si.a, si.b);
        ao[p] = pack_rn(so.a, so.b);
        CC[it][p] = pack_rn(sc.a, sc.b);
        C[it][p] = cpa[p];
        ss2 = add2f(ss2, unpack2(mul2(CC[it][p], CC[it][p])));
        if constexpr (!EXACT) sigmoid6_bf16_exact<false>(af[p], ai[p], ao[p], F[it][p], I[it][p], O[it][p]);
      }
      if constexpr (EXACT) sigmoid24_bf16_exact<true>(af, ai, ao, F[it], I[it], O[it]);
    } else if constexpr (!FULL) {
#pragma unroll
      for (int p = 0; p < 4; ++p) F[it][p] = I[it][p] = O[it][p] = CC[it][p] = C[it][p] = 0u;
    }
  }
  // cell_ln (RMS, affine) -> GELU -> * input gate ; c = c_prev * f + candidate
  float ss = block_sum_1sync(ss2.a + ss2.b, scratch);   // also orders the table fill before its first read
  {
    const float mean = ST<T>::round(ss * inv_d);
    const float inv = ST<T>::round(rsqrtf(ST<T>::round(mean + eps)));
    const uint32_t inv2 = pack_rn(inv, inv);
    ss2 = bcast2(0.f);
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (FULL || vi < nvec) {
        uint4 w, b;
        if constexpr (PRE) {
          w = lnw1[it];
          b = lnb1[it];
        } else {
          w = *reinterpret_cast<const uint4*>(wc + vi * 8);
          b = *reinterpret_cast<const uint4*>(bc + vi * 8);
        }
        const uint32_t wa[4] = {w.x, w.y, w.z, w.w}, ba[4] = {b.x, b.y, b.z, b.w};
        uint32_t y[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          y[p] = add2(mul2(wa[p], mul2(CC[it][p], inv2)), ba[p]);
          if constexpr (!EXACT) y[p] = gelu2_bf16_exact<false>(y[p], lut_s);
        }
        if constexpr (EXACT) gelu8_bf16_table(y, lut_s);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const uint32_t cn = add2(mul2(C[it][p], F[it][p]), mul2(y[p], I[it][p]));
          C[it][p] = cn;
          ss2 = add2f(ss2, unpack2(mul2(cn, cn)));
        }
      }
    }
  }
  // state_ln (RMS, affine) -> GELU -> * output gate
  ss = block_sum_1sync(ss2.a + ss2.b, scratch2);
  {
    const float mean = ST<T>::round(ss * inv_d);
    const float inv = ST<T>::round(rsqrtf(ST<T>::round(mean + eps)));
    const uint32_t inv2 = pack_rn(inv, inv);
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (FULL || vi < nvec) {
        uint4 w, b;
        if constexpr (PRE) {
          w = lnw2[it];
          b = lnb2[it];
        } else {
          w = *reinterpret_cast<const uint4*>(ws + vi * 8);
          b = *reinterpret_cast<const uint4*>(bs + vi * 8);
        }
        const uint32_t wa[4] = {w.x, w.y, w.z, w.w}, ba[4] = {b.x, b.y, b.z, b.w};
        uint32_t y[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          y[p] = add2(mul2(wa[p], mul2(C[it][p], inv2)), ba[p]);
          if constexpr (!EXACT) y[p] = gelu2_bf16_exact<false>(y[p], lut_s);
        }
        if constexpr (EXACT) gelu8_bf16_table(y, lut_s);
#pragma unroll
        for (int p = 0; p < 4; ++p) F[it][p] = mul2(y[p], O[it][p]);   // the new state (F is dead: reuse its registers)
        *reinterpret_cast<uint4*>(state_out + row * D + vi * 8) =
            make_uint4(F[it][0], F[it][1], F[it][2], F[it][3]);
        *reinterpret_cast<uint4*>(cell_out + row * D + vi * 8) =
            make_uint4(C[it][0], C[it][1], C[it][2], C[it][3]);
      }
    }
  }
  if (state_q != nullptr) {
    // dynamic per-token FP8 copy of the state for the quantised head (see the generic kernel)
    float amax = 0.f;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
#pragma unroll
      for (int p = 0; p < 4; ++p)
        amax = fmaxf(amax, fmaxf(fabsf(bf_lo(F[it][p])), fabsf(bf_hi(F[it][p]))));
    __syncthreads();   // scratch is free again (block_sum_1sync leaves no trailing barrier)
    amax = block_max(amax, scratch);
    const float scale = fmaxf(amax / kFp8Max, kFp8MinScale);
    if (threadIdx.x == 0) state_scale[row] = scale;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (FULL || vi < nvec) {
        uint32_t q[2] = {0u, 0u};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const uint32_t lo = to_e4m3(bf_lo(F[it][p]) / scale), hi = to_e4m3(bf_hi(F[it][p]) / scale);
          q[p >> 1] |= (lo | (hi << 8)) << (16 * (p & 1));
        }
        *reinterpret_cast<uint2*>(state_q + row * D + vi * 8) = make_uint2(q[0], q[1]);
      }
    }
  }
  if (threadIdx.x == 0) trace_stamp(trace, 2);
}

#ifdef AB200_DEBUG
// bf16 unary functions of the bit-exact emulation on arbitrary inputs (test hook: all 65 536
// values).  fast/exact: bf16 results of the shortcut / of ATen's formula; fast32/exact32
// (optional): the unguarded fast approximation and ATen's value in fp32, to measure the guard.
__global__ void unary_bf16_kernel(int op, const uint16_t* __restrict__ in, uint16_t* __restrict__ fast,
                                  uint16_t* __restrict__ exact, float* __restrict__ fast32,
                                  float* __restrict__ exact32, int64_t n) {
  __shared__ float gelu_lut[kGeluLutSize];
  __shared__ __align__(16) uint16_t gelu_tab[kGeluTabEntries];
#pragma unroll 1
  for (int i = threadIdx.x; i < kGeluLutSize; i += blockDim.x) gelu_lut[i] = gelu_lut_entry(i);
#pragma unroll 1
  for (int i = threadIdx.x; i < kGeluTabEntries; i += blockDim.x) gelu_tab[i] = gelu_tab_entry(i);
  uint32_t lut_s = static_cast<uint32_t>(__cvta_generic_to_shared(gelu_lut));
  asm volatile("mov.u32 %0, %0;" : "+r"(lut_s));
  uint32_t tab_s = static_cast<uint32_t>(__cvta_generic_to_shared(gelu_tab));
  asm volatile("mov.u32 %0, %0;" : "+r"(tab_s));
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // ops 2 / 3: table GELU on the pair (in[i], in[n-1-i]); 2 reports the low half, 3 the high half
    const uint16_t xin = op == 3 ? in[n - 1 - i] : in[i];
    const float x = __uint_as_float(static_cast<uint32_t>(xin) << 16);
    const uint32_t x2 = op >= 2 ? static_cast<uint32_t>(in[i]) | (static_cast<uint32_t>(in[n - 1 - i]) << 16)
                                : static_cast<uint32_t>(in[i]) * 0x10001u;
    uint32_t f, f_i, f_o;
    if (op == 0) sigmoid6_bf16_exact<true>(x2, x2, x2, f, f_i, f_o);
    else if (op == 1) f = gelu2_bf16_exact<true>(x2, lut_s);
    else f = gelu2_bf16_table(x2, tab_s);
    if (op == 3) f >>= 16;
    const float e = op == 0 ? sigmoid_aten(x) : gelu_erf_aten(x);
    fast[i] = static_cast<uint16_t>(f & 0xFFFFu);
    exact[i] = static_cast<uint16_t>(pack_rn(e, 0.f) & 0xFFFFu);
    if (fast32) {
      const f2 h = half_erfc2(f2{x, x});
      fast32[i] = op == 0 ? sigmoid_fast(x) : gelu_lut_or(x * (x < 0.0f ? h.a : 1.0f - h.a), x, lut_s);
    }
    if (exact32) exact32[i] = e;
  }
}

// Search hook for the guard-free sigmoid: candidate c = (c1, k, addend) evaluates
// rcp(fma(ex2(x * c1), k, addend)) on all 65 536 bf16 inputs x >= lo_bound and counts the inputs whose
// bf16 rounding differs from ATen's formula (tools/sigmoid_search.py).  One block per candidate.
__global__ void sigmoid_search_kernel(const float* __restrict__ cand, uint32_t* __restrict__ mismatches,
                                      uint32_t* __restrict__ first_bad, float lo_bound) {
  const float c1 = cand[3 * blockIdx.x], k = cand[3 * blockIdx.x + 1], addend = cand[3 * blockIdx.x + 2];
  uint32_t count = 0u, bad = 0xffffffffu;
  for (uint32_t bits = threadIdx.x; bits < 65536u; bits += blockDim.x) {
    const float x = __uint_as_float(bits << 16);
    if (!(x >= lo_bound)) continue;
    const float s = rcp_approx(__fmaf_rn(ex2_approx(__fmul_rn(x, c1)), k, addend));
    const uint32_t fast = pack_rn(s, 0.f) & 0xffffu, exact = pack_rn(sigmoid_aten(x), 0.f) & 0xffffu;
    if (fast != exact) {
      ++count;
      bad = min(bad, bits);
    }
  }
  if (count) {
    atomicAdd(mismatches + blockIdx.x, count);
    atomicMin(first_bad + blockIdx.x, bad);
  }
}

// Same search for the polynomial GELU on y > lo_bound: candidate = (the last two coefficients and the
// -1 term); arithmetic as in half_erfc2 / gelu2_bf16_exact.
__global__ void gelu_search_kernel(const float* __restrict__ cand, uint32_t* __restrict__ mismatches,
                                   uint32_t* __restrict__ first_bad, float lo_bound) {
  const float c5 = cand[3 * blockIdx.x], c6 = cand[3 * blockIdx.x + 1], cm1 = cand[3 * blockIdx.x + 2];
  uint32_t count = 0u, bad = 0xffffffffu;
  for (uint32_t bits = threadIdx.x; bits < 65536u; bits += blockDim.x) {
    const float y = __uint_as_float(bits << 16);
    if (!(y > lo_bound)) continue;
    const float u = fminf(fabsf(y), 5.6568542494923806f);
    float q = 4.4679095481e-06f;
    q = __fmaf_rn(q, u, -1.4861870113e-05f);
    q = __fmaf_rn(q, u, -5.6600669632e-04f);
    q = __fmaf_rn(q, u, 7.6454123482e-03f);
    q = __fmaf_rn(q, u, -5.2912395447e-02f);
    q = __fmaf_rn(q, u, c5);
    q = __fmaf_rn(q, u, c6);
    const float h = ex2_approx(__fmaf_rn(q, u, cm1));
    const float g = __fmul_rn(y, (bits & 0x8000u) ? h : __fsub_rn(1.0f, h));
    const uint32_t fast = pack_rn(g, 0.f) & 0xffffu, exact = pack_rn(gelu_erf_aten(y), 0.f) & 0xffffu;
    if (fast != exact) {
      ++count;
      bad = min(bad, bits);
    }
  }
  if (count) {
    atomicAdd(mismatches + blockIdx.x, count);
    atomicMin(first_bad + blockIdx.x, bad);
  }
}

#endif  // AB200_DEBUG

// debug switch (ab200_debug_cell_generic, debug flavour only): 1 = run the generic kernel also for
// bf16 + emulation
#ifdef AB200_DEBUG
static std::atomic<int> g_cell_generic{0};
#else
static constexpr int g_cell_generic = 0;
#endif
static DeviceOnce g_gelu_tab_once;
int init_lstm_tables(cudaStream_t s) {
  return device_once(g_gelu_tab_once, s, "GELU table", [](cudaStream_t st) {
    gelu_lut_init_kernel<<<kGeluTabEntries / 256, 256, 0, st>>>();
    AB200_CHECK_LAUNCH("gelu_lut_init_kernel");
    return AB200_OK;
  });
}
// block size / items-per-thread for a row of `nvec` 16-byte vectors
static bool row_launch_shape(int64_t nvec, int* threads, int* items) {
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
static int launch_rms(const void* x, void* y, int64_t rows, int64_t cols, int64_t ldx, int64_t ldy,
                      float eps, float post_scale, cudaStream_t s) {
  int threads, items;
  AB200_REQUIRE(row_launch_shape(cols / Vec16<T>::N, &threads, &items), AB200_E_INVALID,
                "rms_scale: cols %lld too large", (long long)cols);
  const T* xp = static_cast<const T*>(x);
  T* yp = static_cast<T*>(y);
  LaunchCfg lc(dim3(static_cast<unsigned>(rows)), dim3(threads), 0, s);
  const int icols = static_cast<int>(cols);
  unsigned long long* tr = next_trace_slot();
  switch (items) {
    case 1: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, rms_scale_kernel<T, EMU, 1>, xp, yp, icols, ldx, ldy, eps, post_scale, tr)); break;
    case 2: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, rms_scale_kernel<T, EMU, 2>, xp, yp, icols, ldx, ldy, eps, post_scale, tr)); break;
    default: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, rms_scale_kernel<T, EMU, 4>, xp, yp, icols, ldx, ldy, eps, post_scale, tr)); break;
  }
  AB200_CHECK_LAUNCH("rms_scale_kernel");
  return AB200_OK;
}

template <typename T, bool EMU>
static int launch_cell(const void* gates, int64_t ldg, const void* emb, const int64_t* tokens,
                       const uint64_t* packed, const void* cell_prev, const void* wc,
                       const void* bc, const void* ws, const void* bs, void* state_out,
                       void* cell_out, int64_t B, int64_t D, int64_t vocab, float alpha, float eps,
                       uint64_t* clear_packed, uint32_t* clear_counter, void* state_q,
                       float* state_scale, bool exact, int early, cudaStream_t s) {
  int threads, items;
  AB200_REQUIRE(row_launch_shape(D / Vec16<T>::N, &threads, &items), AB200_E_INVALID,
                "sum_lstm_cell: D %lld too large", (long long)D);
  LaunchCfg lc(dim3(static_cast<unsigned>(B)), dim3(threads), 0, s);
  const T* a_gates = static_cast<const T*>(gates);
  const T* a_emb = static_cast<const T*>(emb);
  const unsigned long long* a_packed = reinterpret_cast<const unsigned long long*>(packed);
  const T* a_cprev = static_cast<const T*>(cell_prev);
  const T* a_wc = static_cast<const T*>(wc);
  const T* a_bc = static_cast<const T*>(bc);
  const T* a_ws = static_cast<const T*>(ws);
  const T* a_bs = static_cast<const T*>(bs);
  T* a_state = static_cast<T*>(state_out);
  T* a_cell = static_cast<T*>(cell_out);
  const int32_t a_D = static_cast<int32_t>(D);
  unsigned long long* a_clear = reinterpret_cast<unsigned long long*>(clear_packed);
  unsigned int* a_cnt = reinterpret_cast<unsigned int*>(clear_counter);
  unsigned long long* a_trace = next_trace_slot();
  uint8_t* a_sq = static_cast<uint8_t*>(state_q);
#define AB200_CELL_ARGS                                                                          \
  a_gates, ldg, a_emb, tokens, a_packed, a_cprev, a_wc, a_bc, a_ws, a_bs, a_state, a_cell, a_D, \
      vocab, alpha, eps, a_clear, a_cnt, a_trace, a_sq, state_scale
  if constexpr (EMU && std::is_same<T, __nv_bfloat16>::value) {
    if (!g_cell_generic) {
      // first use on this device: fill the GELU table (once, synchronously, under a mutex - the
      // kernel reads the table before its dependency wait, so it must be complete before ANY launch)
      int rc = init_lstm_tables(s);
      if (rc != AB200_OK) return rc;
      // <= 512 threads per row (D <= 4096): 40 registers so that three CTAs share an SM
      const float inv_d = 1.0f / static_cast<float>(D);
      const bool full = D == 8LL * items * threads;
#define AB200_FAST_ARGS                                                                          \
  a_gates, ldg, a_emb, tokens, a_packed, a_cprev, a_wc, a_bc, a_ws, a_bs, a_state, a_cell, a_D, \
      vocab, alpha, eps, inv_d, a_clear, a_cnt, a_trace, a_sq, state_scale, early
#define AB200_FAST_LAUNCH2(IT, MAXT, MINB, FULLV)                                                 \
  do {                                                                                           \
    if (exact)                                                                                   \
      AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_lstm_cell_bf16x2_kernel<IT, MAXT, MINB, FULLV, true>, \
                                       AB200_FAST_ARGS));                                        \
    else                                                                                         \
      AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_lstm_cell_bf16x2_kernel<IT, MAXT, MINB, FULLV, false>, \
                                       AB200_FAST_ARGS));                                        \
  } while (0)
#define AB200_FAST_LAUNCH(IT, MAXT, MINB)                                                        \
  do {                                                                                           \
    if (full) AB200_FAST_LAUNCH2(IT, MAXT, MINB, true);                                          \
    else AB200_FAST_LAUNCH2(IT, MAXT, MINB, false);                                              \
  } while (0)
      // 16 elements per thread amortise the per-thread prologue once the grid fills the GPU many
      // times over; below that the shorter per-thread chain of 8 elements wins on latency
      if (B >= 4096 && items == 1 && threads % 64 == 0 && threads >= 128 && threads <= 512) {
        lc.cfg.blockDim = dim3(threads / 2);
        AB200_FAST_LAUNCH(2, 256, 4);
      } else if (items == 1 && threads <= 512 && B <= 160) AB200_FAST_LAUNCH(1, 512, 1);   // latency variant: <= 1 CTA per SM anyway
      else if (items == 1 && threads <= 512) AB200_FAST_LAUNCH(1, 512, 3);
      else if (items == 1) AB200_FAST_LAUNCH(1, 1024, 1);
      else if (items == 2) AB200_FAST_LAUNCH(2, 1024, 1);
      else AB200_FAST_LAUNCH(4, 1024, 1);
#undef AB200_FAST_LAUNCH2
#undef AB200_FAST_LAUNCH
#undef AB200_FAST_ARGS
      AB200_CHECK_LAUNCH("sum_lstm_cell_bf16x2_kernel");
      return AB200_OK;
    }
  }
  switch (items) {
    case 1: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_lstm_cell_kernel<T, EMU, 1>, AB200_CELL_ARGS)); break;
    case 2: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_lstm_cell_kernel<T, EMU, 2>, AB200_CELL_ARGS)); break;
    default: AB200_CUDA_OK(cudaLaunchKernelEx(&lc.cfg, sum_lstm_cell_kernel<T, EMU, 4>, AB200_CELL_ARGS)); break;
  }
#undef AB200_CELL_ARGS
  AB200_CHECK_LAUNCH("sum_lstm_cell_kernel");
  return AB200_OK;
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_rms_scale(const void* x, void* y, int64_t rows, int64_t cols,
                                   int64_t x_row_stride, int64_t y_row_stride, float eps,
                                   float post_scale, int32_t dtype,
                                   int32_t emulate_eager_rounding, void* stream) {
  if (rows == 0) return AB200_OK;
  AB200_REQUIRE(x && y && rows > 0 && cols > 0, AB200_E_INVALID, "rms_scale: bad arguments");
  const int vn = 16 / dtype_bytes(dtype);
  AB200_REQUIRE(cols % vn == 0 && x_row_stride % vn == 0 && y_row_stride % vn == 0 &&
                    aligned(x, 16) && aligned(y, 16),
                AB200_E_INVALID, "rms_scale: cols/strides must be multiples of %d elements and "
                "pointers 16-byte aligned", vn);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool emu = emulate_eager_rounding != 0;
  switch (dtype) {
    case AB200_F32: return launch_rms<float, false>(x, y, rows, cols, x_row_stride, y_row_stride, eps, post_scale, s);
    case AB200_BF16:
      return emu ? launch_rms<__nv_bfloat16, true>(x, y, rows, cols, x_row_stride, y_row_stride, eps, post_scale, s)
                 : launch_rms<__nv_bfloat16, false>(x, y, rows, cols, x_row_stride, y_row_stride, eps, post_scale, s);
    case AB200_F16:
      return emu ? launch_rms<__half, true>(x, y, rows, cols, x_row_stride, y_row_stride, eps, post_scale, s)
                 : launch_rms<__half, false>(x, y, rows, cols, x_row_stride, y_row_stride, eps, post_scale, s);
    default: AB200_REQUIRE(false, AB200_E_INVALID, "rms_scale: unsupported dtype %d", dtype);
  }
}

extern "C" int32_t ab200_sum_lstm_cell(const void* gates, int64_t gates_row_stride,
                                       const void* emb_table, const int64_t* tokens,
                                       const uint64_t* packed_argmax, const void* cell_prev,
                                       const void* cell_ln_weight, const void* cell_ln_bias,
                                       const void* state_ln_weight, const void* state_ln_bias,
                                       void* state_out, void* cell_out, int64_t B, int64_t D,
                                       int64_t vocab, float alpha, float eps, int32_t dtype,
                                       int32_t emulate_eager_rounding, uint64_t* clear_packed,
                                       uint32_t* clear_counter, void* state_q_out,
                                       float* state_scale_out, void* stream) {
  if (B == 0) return AB200_OK;
  AB200_REQUIRE(gates && emb_table && (tokens || packed_argmax) && cell_ln_weight &&
                    cell_ln_bias && state_ln_weight && state_ln_bias && state_out && cell_out,
                AB200_E_INVALID, "sum_lstm_cell: NULL pointer argument");
  AB200_REQUIRE(B > 0 && D > 0 && vocab > 0, AB200_E_INVALID, "sum_lstm_cell: bad sizes");
  AB200_REQUIRE(state_q_out == nullptr || (state_scale_out != nullptr && dtype == AB200_BF16 &&
                                            aligned(state_q_out, 8)),
                AB200_E_INVALID, "sum_lstm_cell: fp8 state output needs bf16 states and a scale buffer");
  const int vn = 16 / dtype_bytes(dtype);
  AB200_REQUIRE(D % vn == 0 && gates_row_stride % vn == 0, AB200_E_INVALID,
                "sum_lstm_cell: D and the gate row stride must be multiples of %d", vn);
  AB200_REQUIRE(aligned(gates, 16) && aligned(emb_table, 16) && aligned(cell_ln_weight, 16) &&
                    aligned(cell_ln_bias, 16) && aligned(state_ln_weight, 16) &&
                    aligned(state_ln_bias, 16) && aligned(state_out, 16) && aligned(cell_out, 16) &&
                    (!cell_prev || aligned(cell_prev, 16)),
                AB200_E_INVALID, "sum_lstm_cell: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int early = (emulate_eager_rounding & AB200_CELL_SIDE_INPUTS_READY) ? 1 : 0;
  emulate_eager_rounding &= ~AB200_CELL_SIDE_INPUTS_READY;
  const bool emu = emulate_eager_rounding != 0;
#define AB200_ARGS                                                                               \
  gates, gates_row_stride, emb_table, tokens, packed_argmax, cell_prev, cell_ln_weight,          \
      cell_ln_bias, state_ln_weight, state_ln_bias, state_out, cell_out, B, D, vocab, alpha, eps, \
      clear_packed, clear_counter, state_q_out, state_scale_out, emulate_eager_rounding != 2,    \
      early, s
  switch (dtype) {
    case AB200_F32: return launch_cell<float, false>(AB200_ARGS);
    case AB200_BF16:
      return emu ? launch_cell<__nv_bfloat16, true>(AB200_ARGS)
                 : launch_cell<__nv_bfloat16, false>(AB200_ARGS);
    case AB200_F16:
      return emu ? launch_cell<__half, true>(AB200_ARGS) : launch_cell<__half, false>(AB200_ARGS);
    default: AB200_REQUIRE(false, AB200_E_INVALID, "sum_lstm_cell: unsupported dtype %d", dtype);
  }
#undef AB200_ARGS
}

#ifdef AB200_DEBUG
extern "C" int32_t ab200_debug_cell_generic(int32_t on) {
  return g_cell_generic.exchange(on != 0);
}

extern "C" int32_t ab200_debug_sigmoid_search(const float* candidates, int32_t n_candidates,
                                              uint32_t* mismatches, uint32_t* first_bad,
                                              float lo_bound, void* stream) {
  AB200_REQUIRE(candidates && mismatches && first_bad && n_candidates != 0, AB200_E_INVALID,
                "debug_sigmoid_search: bad arguments");
  if (n_candidates < 0)   // negative count: the GELU search
    gelu_search_kernel<<<-n_candidates, 256, 0, static_cast<cudaStream_t>(stream)>>>(candidates, mismatches,
                                                                                     first_bad, lo_bound);
  else
    sigmoid_search_kernel<<<n_candidates, 256, 0, static_cast<cudaStream_t>(stream)>>>(candidates, mismatches,
                                                                                       first_bad, lo_bound);
  AB200_CHECK_LAUNCH("sigmoid_search_kernel");
  return AB200_OK;
}

extern "C" int32_t ab200_debug_unary_bf16(int32_t op, const uint16_t* in, uint16_t* out_fast,
                                          uint16_t* out_exact, float* out_fast32,
                                          float* out_exact32, int64_t n, void* stream) {
  AB200_REQUIRE(op >= 0 && op <= 3 && in && out_fast && out_exact && n > 0, AB200_E_INVALID,
                "debug_unary_bf16: bad arguments");
  unary_bf16_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(op, in, out_fast, out_exact,
                                                                         out_fast32, out_exact32, n);
  AB200_CHECK_LAUNCH("unary_bf16_kernel");
  return AB200_OK;
}
#endif  // AB200_DEBUG
