// K2s — skinny GEMM for the batch-1 streaming step (<= 16 output rows in total).
//
// With 8 rows per codec token the per-token step is a sequence of GEMVs: every weight byte is used once, so the
// step is bound by streaming the 429 MB of fp16 weights from HBM (>= 66 us at 6.5 TB/s), not by tensor throughput.
// A 128-row tcgen05 tile would leave all but N/256 SMs idle; here each CTA owns 8 output columns so a 1024-column
// layer spreads over 128 SMs, the 8 warps of the CTA split K, every lane streams 16-byte weight vectors (all of a
// warp's vectors in flight at once), the (im2col'd, causal) activation rows sit in shared memory, and the products run
// on mma.sync.m16n8k16 (m = the <= 16 rows, n = the 8 columns) with fp32 accumulation.  Same contract / epilogues as
// gemm_tc.
#include <cstdlib>

#include "common.cuh"
#include "gemm_skinny_body.cuh"

namespace frt2 {

namespace {

template <int MR, int NT>
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(GemmDesc g, int mtot) {
  extern __shared__ __align__(16) uint8_t sk_smem_dyn[];
  gemm_skinny_body<MR, true, NT>(g, mtot, blockIdx.x, threadIdx.x, sk_smem_dyn, [] { __syncthreads(); });
}

// ---- round-1 FMA-loop kernel (one output column per warp, fp32 FMAs), kept for A/B measurements ----
__device__ __forceinline__ float dot8(const uint4& w, const uint4& a, float acc) {
  const __half2* w2 = reinterpret_cast<const __half2*>(&w);
  const __half2* a2 = reinterpret_cast<const __half2*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 wf = __half22float2(w2[i]);
    const float2 af = __half22float2(a2[i]);
    acc = fmaf(wf.x, af.x, acc);
    acc = fmaf(wf.y, af.y, acc);
  }
  return acc;
}

template <int MR>
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_fma_kernel(GemmDesc g, int mtot) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  __half* sA = reinterpret_cast<__half*>(sk_smem);                     // [MR][kchunk]
  const int Ktot = g.ntaps * g.Kc;
  const int kchunk = min(Ktot, SK_KCHUNK);
  float* sOut = reinterpret_cast<float*>(sk_smem + static_cast<size_t>(MR) * kchunk * 2);  // [MR][SK_COLS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * SK_COLS + warp;
  const bool col_ok = n < g.N;
  const uint4* wrow = reinterpret_cast<const uint4*>(g.W + static_cast<long long>(col_ok ? n : 0) * Ktot);

  // ---- predecessor-independent prologue: the first weight vectors of this warp's row are already in flight while
  //      the previous kernel of the step is still running (PDL), and while the activations are staged below
  constexpr int PRE = 4;   // (kept as the A/B baseline: FRT2_SKINNY_FMA=1)
  uint4 wpre[PRE];
  {
    const int kc8_first = min(kchunk, Ktot) >> 3;
#pragma unroll
    for (int u = 0; u < PRE; ++u) {
      const int k8 = lane + 32 * u;
      wpre[u] = (col_ok && k8 < kc8_first) ? __ldg(wrow + k8) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // the epilogue's bias values (thread t -> column t % 8, and its polar partner) are also predecessor-independent
  float bias_c = 0.f, bias_p = 0.f;
  if (g.bias != nullptr && threadIdx.x < MR * SK_COLS) {
    const int c = threadIdx.x % SK_COLS;
    const int nn = blockIdx.x * SK_COLS + c;
    if (nn < g.N) bias_c = __ldg(g.bias + nn);
    if ((nn ^ 1) < g.N) bias_p = __ldg(g.bias + (nn ^ 1));
  }
  pdl_wait();      // everything below may read what the previous kernel wrote
  pdl_trigger();   // the next kernel may start its own weight prefetch now

  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;

  for (int kc0 = 0; kc0 < Ktot; kc0 += kchunk) {
    const int kc = min(kchunk, Ktot - kc0);
    const int kc8 = kc >> 3;
    if (g.ln_gamma != nullptr) {
      // ---- fused LayerNorm(+SiLU): one warp per row, fp32 statistics, result straight into the fp16 A tile
      //      (reference nn.LayerNorm eps 1e-5 / 1e-6: whisper.py:134,140, decoder.py:246)
      for (int m = warp; m < MR; m += SK_WARPS) {
        __half* arow = sA + static_cast<size_t>(m) * kchunk;
        if (m >= mtot) {
          for (int c = lane * 8; c < kc; c += 256) *reinterpret_cast<uint4*>(arow + c) = make_uint4(0u, 0u, 0u, 0u);
          continue;
        }
        const float4* xr = reinterpret_cast<const float4*>(g.ln_x + static_cast<long long>(m) * g.ln_ldx);
        const int C4 = g.Kc >> 2;
        float s = 0.f;
        for (int c = lane; c < C4; c += 32) {
          const float4 v = xr[c];
          s += (v.x + v.y) + (v.z + v.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / static_cast<float>(g.Kc);
        float q = 0.f;
        for (int c = lane; c < C4; c += 32) {
          const float4 v = xr[c];
          const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
          q += (a * a + b * b) + (cc * cc + d * d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / static_cast<float>(g.Kc) + g.ln_eps);
        const float4* g4 = reinterpret_cast<const float4*>(g.ln_gamma);
        const float4* b4 = reinterpret_cast<const float4*>(g.ln_beta);
        for (int c = lane; c < C4; c += 32) {
          const float4 v = xr[c], gg = __ldg(g4 + c), bb = __ldg(b4 + c);
          float y0 = (v.x - mean) * rstd * gg.x + bb.x, y1 = (v.y - mean) * rstd * gg.y + bb.y;
          float y2 = (v.z - mean) * rstd * gg.z + bb.z, y3 = (v.w - mean) * rstd * gg.w + bb.w;
          if (g.ln_silu) { y0 = silu(y0); y1 = silu(y1); y2 = silu(y2); y3 = silu(y3); }
          uint2 h;
          h.x = pack_half2(y0, y1);
          h.y = pack_half2(y2, y3);
          *reinterpret_cast<uint2*>(arow + 4 * c) = h;
        }
      }
    } else {
      // ---- activation chunk -> smem (causal taps gathered here: row m, tap j reads input row r + j + row_shift)
      for (int e = threadIdx.x; e < MR * kc8; e += blockDim.x) {
        const int m = e / kc8, k8 = e - m * kc8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (m < mtot) {
          const int b = m / g.rows_out, r = m - b * g.rows_out;
          const int kk = kc0 + k8 * 8;
          const int tap = kk / g.Kc, c = kk - tap * g.Kc;
          const int src = r + tap + g.row_shift;
          if (src >= 0 && src < g.rows_a)
            v = *reinterpret_cast<const uint4*>(g.A + static_cast<long long>(b) * g.a_batch_pitch +
                                                static_cast<long long>(src) * g.a_row_pitch + c);
        }
        *reinterpret_cast<uint4*>(sA + static_cast<size_t>(m) * kchunk + k8 * 8) = v;
      }
    }
    __syncthreads();
    if (col_ok) {
      const uint4* wp = wrow + (kc0 >> 3);
#pragma unroll
      for (int u = 0; u < PRE; ++u) {      // the prefetched vectors (first chunk) or fresh loads (later chunks)
        const int k8 = lane + 32 * u;
        if (k8 < kc8) {
          const uint4 w = (kc0 == 0) ? wpre[u] : __ldg(wp + k8);
#pragma unroll
          for (int m = 0; m < MR; ++m) {
            const uint4 a = *reinterpret_cast<const uint4*>(sA + static_cast<size_t>(m) * kchunk + k8 * 8);
            acc[m] = dot8(w, a, acc[m]);
          }
        }
      }
#pragma unroll 4
      for (int k8 = lane + 32 * PRE; k8 < kc8; k8 += 32) {
        const uint4 w = __ldg(wp + k8);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
          const uint4 a = *reinterpret_cast<const uint4*>(sA + static_cast<size_t>(m) * kchunk + k8 * 8);
          acc[m] = dot8(w, a, acc[m]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) {
    float v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sOut[m * SK_COLS + warp] = v;
  }
  __syncthreads();
  // ---- epilogue: thread t -> (row m, column c); polar pairs read the neighbouring column from smem
  const int t = threadIdx.x;
  if (t < MR * SK_COLS) {
    const int m = t / SK_COLS, c = t - m * SK_COLS;
    const int nn = blockIdx.x * SK_COLS + c;
    if (m < mtot && nn < g.N) {
      auto pre = [&](int cc) {   // cc == c (own column) or c ^ 1 (polar partner): biases were prefetched
        return fmaf(sOut[m * SK_COLS + cc], g.alpha, cc == c ? bias_c : bias_p);
      };
      float v = pre(c);
      if (g.act == ACT_GELU) {
        v = gelu_erf(v);
      } else if (g.act == ACT_POLAR) {   // (log-magnitude, phase) pairs: reference decoder.py:505-518
        const float lm = (c & 1) ? pre(c ^ 1) : v;
        const float ph = (c & 1) ? v : pre(c ^ 1);
        const float mag = fminf(expf(lm), 100.0f);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        v = (c & 1) ? mag * sn : mag * cs;
      }
      const int b = m / g.rows_out, r = m - b * g.rows_out;
      const int roff = (g.out_row_off != nullptr) ? __ldg(g.out_row_off + b * g.row_off_stride) : 0;
      const long long o32 = static_cast<long long>(b) * g.pitch32 + static_cast<long long>(r + roff) * g.ld32 + nn;
      const long long o16 = static_cast<long long>(b) * g.pitch16 + static_cast<long long>(r + roff) * g.ld16 + nn;
      if (g.split_col > 0 && nn >= g.split_col) {
        const int roff_b = (g.row_off_b != nullptr) ? __ldg(g.row_off_b + b * g.row_off_stride) : 0;
        g.out16_b[static_cast<long long>(b) * g.pitch16_b + static_cast<long long>(r + roff_b) * g.ld16_b +
                  (nn - g.split_col)] = to_half_sat(v);
      } else {
        if (g.resid != nullptr) v += g.resid[o32];
        if (g.out32 != nullptr) g.out32[o32] = v;
        if (g.out16 != nullptr) g.out16[o16] = to_half_sat(v);
      }
    }
  }
}

}  // namespace

int gemm_skinny_init() {
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_fma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    8 * SK_KCHUNK * 2 + 8 * SK_COLS * 4));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_fma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    16 * SK_KCHUNK * 2 + 16 * SK_COLS * 4));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 1))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(16, SK_KCHUNK, SK_KCHUNK, 1))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 2))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(16, SK_KCHUNK, SK_KCHUNK, 2))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 3))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 4))));
  return FRT2_OK;
}

bool gemm_skinny_applicable(const GemmDesc& g) {
  return static_cast<long long>(g.batches) * g.rows_out <= SK_MAXROWS && g.Kc % 8 == 0 &&
         (g.act != ACT_POLAR || g.N % 2 == 0);
}

int gemm_skinny(const GemmDesc& g, cudaStream_t stream) {
  FRT2_REQUIRE(gemm_skinny_applicable(g), FRT2_ERR_BAD_ARG, "gemm_skinny: more than 16 rows");
  FRT2_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0 &&
                   g.a_row_pitch % 8 == 0 && g.a_batch_pitch % 8 == 0,
               FRT2_ERR_BAD_ARG, "gemm_skinny: operands must be 16-byte aligned");
  const int mtot = g.batches * g.rows_out;
  const int Ktot = g.ntaps * g.Kc;
  const int kchunk = Ktot < SK_KCHUNK ? Ktot : SK_KCHUNK;
  const int grid = (g.N + SK_COLS - 1) / SK_COLS;
  FRT2_REQUIRE(g.ln_gamma == nullptr || (g.ntaps == 1 && g.Kc <= SK_KCHUNK && g.Kc % 4 == 0 && g.ln_x != nullptr),
               FRT2_ERR_BAD_ARG, "gemm_skinny: fused LayerNorm needs ntaps == 1 and Kc <= 4096");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_WARPS * 32);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap the weight prefetch with the
  attr[0].val.programmaticStreamSerializationAllowed = 1;            // tail of the previous kernel of the step
  // Programmatic dependent launch: the next skinny kernel's prologue (all of its weight vectors, gamma/beta) is in flight
  // while the current kernel drains.  With the round-1 kernels this cost ~100 us per step inside the captured graph
  // (early-resident dependents competed for SM slots); with the current ones it saves ~20 (557 -> 538 us per token, full
  // GPU suite green with it), so it is on by default; FRT2_NO_PDL=1 restores plain stream order (A/B).
  static const bool use_pdl = (getenv("FRT2_NO_PDL") == nullptr);
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  static const bool use_fma = (getenv("FRT2_SKINNY_FMA") != nullptr);   // A/B: the round-1 FMA-loop kernel
  if (use_fma) {
    if (mtot <= 8) {
      cfg.dynamicSmemBytes = static_cast<size_t>(8) * kchunk * 2 + 8 * SK_COLS * 4;
      FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_fma_kernel<8>, g, mtot));
    } else {
      cfg.dynamicSmemBytes = static_cast<size_t>(16) * kchunk * 2 + 16 * SK_COLS * 4;
      FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_fma_kernel<16>, g, mtot));
    }
    FRT2_CUDA_OK(cudaGetLastError());
    return FRT2_OK;
  }
  // wide layers (QKV, fc1): 16 columns per CTA — one wave of CTAs, activation rows staged / normalised half as often
  static const bool no_nt2 = (getenv("FRT2_SKINNY_NT1") != nullptr);   // A/B switch for measurements
  static const int nt2_min_n = [] {
    const char* e = getenv("FRT2_SKINNY_NT2_MIN_N");
    return e != nullptr ? atoi(e) : 2048;
  }();
  static const int nt_max = [] {   // up to 3 / 4 n-tiles per CTA: one CTA per SM for N = 3072 / 4096 (A/B: =2 -> 562 vs 554 us)
    const char* e = getenv("FRT2_SKINNY_NTMAX");
    return e != nullptr ? atoi(e) : 4;
  }();
  int nt = (g.N >= nt2_min_n && !no_nt2) ? 2 : 1;
  if (nt == 2 && nt_max > 2 && mtot <= 8) nt = std::max(2, std::min(nt_max, (g.N / 8 + 147) / 148));
  cfg.gridDim = dim3((g.N + SK_COLS * nt - 1) / (SK_COLS * nt));
  const int ln_c = g.ln_gamma != nullptr ? g.Kc : 0;
  if (mtot <= 8) {
    cfg.dynamicSmemBytes = sk_smem_bytes(8, kchunk, ln_c, nt);
    switch (nt) {
      case 4: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 4>, g, mtot)); break;
      case 3: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 3>, g, mtot)); break;
      case 2: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 2>, g, mtot)); break;
      default: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 1>, g, mtot)); break;
    }
  } else {
    cfg.dynamicSmemBytes = sk_smem_bytes(16, kchunk, ln_c, nt);
    if (nt == 2) FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<16, 2>, g, mtot));
    else FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<16, 1>, g, mtot));
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
