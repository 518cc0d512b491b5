// K2s — skinny GEMM for the batch-1 streaming step (<= 16 output rows in total).
//
// With 8 rows per codec token the per-token step is a sequence of GEMVs: every weight byte is used once, so the
// step is bound by streaming the 429 MB of fp16 weights from HBM (>= 66 us at 6.5 TB/s), not by tensor throughput.
// A 128-row tcgen05 tile would leave all but N/256 SMs idle; here each CTA owns 8 output columns so a 1024-column
// layer spreads over 128 SMs, each warp streams one weight row with coalesced 16-byte loads, the (im2col'd, causal)
// activation rows sit in shared memory, and accumulation is fp32.  Same contract / epilogues as gemm_tc.
#include "common.cuh"

namespace frt2 {

namespace {

constexpr int SK_WARPS = 8;        // one output column per warp
constexpr int SK_COLS = SK_WARPS;
constexpr int SK_MAXROWS = 16;
constexpr int SK_KCHUNK = 4096;    // activation chunk held in smem: 16 rows x 4096 halves = 128 KB

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
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(GemmDesc g, int mtot) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  __half* sA = reinterpret_cast<__half*>(sk_smem);                     // [MR][kchunk]
  const int Ktot = g.ntaps * g.Kc;
  const int kchunk = min(Ktot, SK_KCHUNK);
  float* sOut = reinterpret_cast<float*>(sk_smem + static_cast<size_t>(MR) * kchunk * 2);  // [MR][SK_COLS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * SK_COLS + warp;
  const bool col_ok = n < g.N;
  const uint4* wrow = reinterpret_cast<const uint4*>(g.W + static_cast<long long>(col_ok ? n : 0) * Ktot);

  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;

  for (int kc0 = 0; kc0 < Ktot; kc0 += kchunk) {
    const int kc = min(kchunk, Ktot - kc0);
    const int kc8 = kc >> 3;
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
    __syncthreads();
    if (col_ok) {
      const uint4* wp = wrow + (kc0 >> 3);
#pragma unroll 4
      for (int k8 = lane; k8 < kc8; k8 += 32) {
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
      auto pre = [&](int cc) {
        const int n2 = blockIdx.x * SK_COLS + cc;
        float v = sOut[m * SK_COLS + cc] * g.alpha;
        if (g.bias != nullptr && n2 < g.N) v += __ldg(g.bias + n2);
        return v;
      };
      float v = pre(c);
      if (g.act == ACT_GELU) {
        v = gelu_erf(v);
      } else if (g.act == ACT_POLAR) {   // (log-magnitude, phase) pairs: reference decoder.py:505-518
        const float lm = (c & 1) ? pre(c - 1) : v;
        const float ph = (c & 1) ? v : pre(c + 1);
        const float mag = fminf(expf(lm), 100.0f);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        v = (c & 1) ? mag * sn : mag * cs;
      }
      const int b = m / g.rows_out, r = m - b * g.rows_out;
      const int roff = (g.out_row_off != nullptr) ? __ldg(g.out_row_off) : 0;
      const long long o32 = static_cast<long long>(b) * g.pitch32 + static_cast<long long>(r + roff) * g.ld32 + nn;
      const long long o16 = static_cast<long long>(b) * g.pitch16 + static_cast<long long>(r + roff) * g.ld16 + nn;
      if (g.resid != nullptr) v += g.resid[o32];
      if (g.out32 != nullptr) g.out32[o32] = v;
      if (g.out16 != nullptr) g.out16[o16] = to_half_sat(v);
    }
  }
}

}  // namespace

int gemm_skinny_init() {
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    8 * SK_KCHUNK * 2 + 8 * SK_COLS * 4));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    16 * SK_KCHUNK * 2 + 16 * SK_COLS * 4));
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
  if (mtot <= 8) {
    const size_t smem = static_cast<size_t>(8) * kchunk * 2 + 8 * SK_COLS * 4;
    gemm_skinny_kernel<8><<<grid, SK_WARPS * 32, smem, stream>>>(g, mtot);
  } else {
    const size_t smem = static_cast<size_t>(16) * kchunk * 2 + 16 * SK_COLS * 4;
    gemm_skinny_kernel<16><<<grid, SK_WARPS * 32, smem, stream>>>(g, mtot);
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
