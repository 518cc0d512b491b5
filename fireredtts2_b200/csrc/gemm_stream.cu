// K2w — persistent weight-streaming GEMM for <= 8 activation rows (the frame tail of the speech LM, frame_decoder.cu).
//
// out[m, n] = act(sum_k A[m, k] W[n, k] + bias[n]) (+ resid), m < B <= 8.  Every weight byte is used once, so the kernel is
// an HBM stream with a little arithmetic attached.  Against gemm_skinny (one CTA per 8..32 columns, load -> wait -> MMA ->
// reduce -> store per CTA; 164 registers, one CTA per SM and 3.8 waves on the 17920-column gate|up layer: 1.8 TB/s):
//   * ONE CTA of 16 warps per SM for the whole launch.  The activation rows are staged (and RMS-normalised) once per CTA,
//     not once per column tile.
//   * Weights are repacked at load into the order the warps consume them: tile-blocked [N/8][K/32][8 columns][32 k], so
//     the 16-byte vector of lane (g, t) = W[8j+g][32b+8t..] sits at lane*16 bytes of a 512-byte block: every warp request
//     is four full 128-byte lines.
//   * A CTA owns column tiles round-robin (tile j -> CTA j mod grid) and its 16 warps split the tile's K, so at any moment
//     the whole GPU reads one compact window of the weight matrix; partial sums meet in shared memory (fixed order:
//     deterministic), double-buffered so that one barrier per tile suffices.  (A layer with more tiles than warps in the
//     grid lets every warp stream whole tiles instead: no reduction.)
//   * Every lane keeps a RING of 8 16-byte vectors in flight in registers and refills a slot the moment it is consumed:
//     the stream never drains at tile boundaries, reductions or epilogues (16 warps x 8 x 512 B = 64 KB in flight per SM).
// Measured alternatives: 8 warps x ring 16 / 24 / 32, 16 x 16, a cp.async ring in shared memory (4..12 deep,
// 64 registers, two co-resident CTAs), warp groups of 1..16 per tile with the epilogue operands prefetched — none faster.
// Products on mma.sync.m16n8k16 (rows 8..15 unused) with the k-permutation of gemm_skinny (a dot product does not care
// about the order of k inside a 32-block).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace frt2 {

namespace {

#ifndef GS_WARPS_N
#define GS_WARPS_N 16
#endif
#ifndef GS_RING_N
#define GS_RING_N 8
#endif
#ifndef GS_CTAS_PER_SM
#define GS_CTAS_PER_SM 1     // CTAs of ONE launch per SM (grid = this x SMs)
#endif
constexpr int GS_WARPS = GS_WARPS_N;
constexpr int GS_RING = GS_RING_N;
#ifndef GS_MIN_CTAS
#define GS_MIN_CTAS 1
#endif
constexpr int GS_PAD = 32;     // halves between activation rows beyond K: rows g, g+1 land on disjoint banks

__device__ __forceinline__ void gs_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned int d = static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gs_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// a0 / a2: rows 0..7 of the m16 tile, a1 / a3: rows 8..15 (zero when the launch has <= 8 rows)
__device__ __forceinline__ void gs_mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                            uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// streaming load: the weights are read once per launch, keep them out of L1
__device__ __forceinline__ uint4 gs_ldw(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// columns n, n+1 of row m (n even): bias, activation, residual, stores
__device__ __forceinline__ void gs_emit(const StreamGemm& d, int m, int n, float v0, float v1) {
  if (m >= d.B || n >= d.N) return;
  const bool two = n + 1 < d.N;
  if (d.bias != nullptr) {
    v0 += __ldg(d.bias + n);
    if (two) v1 += __ldg(d.bias + n + 1);
  }
  if (d.act == ACT_SWIGLU) {   // (gate, up) pair -> silu(gate) * up at column n / 2
    if (two) d.out16[static_cast<long long>(m) * d.ld16 + (n >> 1)] = to_half_sat(v0 / (1.0f + expf(-v0)) * v1);
    return;
  }
  if (d.act == ACT_GELU) {
    v0 = gelu_erf(v0);
    v1 = gelu_erf(v1);
  }
  const long long o32 = static_cast<long long>(m) * d.ld32 + n;
  if (d.resid != nullptr) {
    v0 += d.resid[o32];
    if (two) v1 += d.resid[o32 + 1];
  }
  if (d.out32 != nullptr) {
    d.out32[o32] = v0;
    if (two) d.out32[o32 + 1] = v1;
  }
  if (d.out16 != nullptr) {
    const long long o16 = static_cast<long long>(m) * d.ld16 + n;
    d.out16[o16] = to_half_sat(v0);
    if (two) d.out16[o16 + 1] = to_half_sat(v1);
  }
}

// MR = 8: up to 8 activation rows (rows 8..15 of the MMA tile idle); MR = 16: up to 16
template <bool SPLIT, int MR>
__global__ void __launch_bounds__(GS_WARPS * 32, GS_MIN_CTAS) gemm_stream_kernel(StreamGemm d) {
  extern __shared__ __align__(16) uint8_t gs_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int nb = d.K >> 5;                 // k32 blocks per column tile
  const int T = (d.N + 7) >> 3;            // column tiles
  const int pitch = d.K + GS_PAD;
  __half* sA = reinterpret_cast<__half*>(gs_smem);                                            // [B][pitch]
  float* sRed = reinterpret_cast<float*>(gs_smem + static_cast<size_t>(d.B) * pitch * 2);     // SPLIT: [2][GS_WARPS][MR * 8]

  // ---- this warp's segments: (tile, blocks [kb0, kb1)); consecutive segments are `tstep` tiles apart
  int tile0, ntiles, tstep, kb0, kb1;
  if (SPLIT) {
    tile0 = blockIdx.x;
    tstep = gridDim.x;
    ntiles = tile0 < T ? (T - tile0 + tstep - 1) / tstep : 0;
    kb0 = static_cast<int>(static_cast<long long>(warp) * nb / GS_WARPS);
    kb1 = static_cast<int>(static_cast<long long>(warp + 1) * nb / GS_WARPS);
  } else {
    const long long gw = static_cast<long long>(blockIdx.x) * GS_WARPS + warp, nw = static_cast<long long>(gridDim.x) * GS_WARPS;
    tile0 = static_cast<int>(gw * T / nw);
    ntiles = static_cast<int>((gw + 1) * T / nw) - tile0;
    tstep = 1;
    kb0 = 0;
    kb1 = nb;
  }
  const int seg_len = kb1 - kb0;
  const int total = ntiles * seg_len;                                    // blocks this warp consumes
  // load cursor: a pointer that walks the warp's blocks (32 vectors apart) and jumps from the end of a segment to the
  // start of the next one
  const uint4* lp = reinterpret_cast<const uint4*>(d.Wt) + lane + (static_cast<long long>(tile0) * nb + kb0) * 32;
  const int seg_jump = (tstep * nb - seg_len) * 32;
  int l_kb = seg_len, l_left = total;
  auto next_load = [&]() -> uint4 {
    if (l_left <= 0) return make_uint4(0u, 0u, 0u, 0u);
    const uint4 v = gs_ldw(lp);
    --l_left;
    lp += 32;
    if (--l_kb == 0) {
      l_kb = seg_len;
      lp += seg_jump;
    }
    return v;
  };
  // ---- predecessor-independent prologue: fill the ring (programmatic dependent launch: the previous kernel of the
  //      frame is still draining)
  uint4 ring[GS_RING];
#pragma unroll
  for (int u = 0; u < GS_RING; ++u) ring[u] = next_load();
  pdl_wait();
  pdl_trigger();

  // ---- activation rows -> shared memory (fp16), RMS-normalised on the way when gamma is given
  if (d.gamma != nullptr) {
    // torchtune RMSNorm: x * rsqrt(mean(x^2) + eps) * scale in fp32; one warp per row, two passes (sum of squares, emit)
    const int C4 = d.K >> 2;
    for (int m = warp; m < d.B; m += GS_WARPS) {
      const float4* xr = reinterpret_cast<const float4*>(d.x + static_cast<long long>(m) * d.ldx);
      float q = 0.f;
#pragma unroll 4
      for (int c = lane; c < C4; c += 32) {
        const float4 v = xr[c];
        q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / static_cast<float>(d.K) + d.eps);
      const float4* g4 = reinterpret_cast<const float4*>(d.gamma);
      __half* arow = sA + static_cast<size_t>(m) * pitch;
#pragma unroll 4
      for (int c = lane; c < C4; c += 32) {
        const float4 v = xr[c], gg = __ldg(g4 + c);
        uint2 h;
        h.x = pack_half2(v.x * rstd * gg.x, v.y * rstd * gg.y);
        h.y = pack_half2(v.z * rstd * gg.z, v.w * rstd * gg.w);
        *reinterpret_cast<uint2*>(arow + 4 * c) = h;
      }
    }
  } else {
    const int K8 = d.K >> 3;
    for (int e = tid; e < d.B * K8; e += GS_WARPS * 32) {
      const int m = e / K8, k8 = e - m * K8;
      gs_cp_async16(sA + static_cast<size_t>(m) * pitch + k8 * 8, d.A + static_cast<long long>(m) * d.lda + k8 * 8);
    }
    gs_cp_async_wait_all();
  }
  __syncthreads();

  // ---- the stream
  const __half* arow = sA + static_cast<size_t>(gq < d.B ? gq : 0) * pitch + tq * 8;   // rows >= B: results never stored
  const __half* arow_hi = sA + static_cast<size_t>(gq + 8 < d.B ? gq + 8 : 0) * pitch + tq * 8;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // end of segment c_seg: acc[0], acc[1] = row gq, columns 2tq, 2tq+1 of the tile, acc[2], acc[3] = row gq + 8 (this
  // warp's share of K when SPLIT)
  auto finish = [&](int c_seg) {
    const int tile = tile0 + c_seg * tstep;
    if (SPLIT) {
      float* r = sRed + ((c_seg & 1) * GS_WARPS + warp) * (MR * 8);
      *reinterpret_cast<float2*>(r + gq * 8 + 2 * tq) = make_float2(acc[0], acc[1]);
      if (MR == 16) *reinterpret_cast<float2*>(r + (gq + 8) * 8 + 2 * tq) = make_float2(acc[2], acc[3]);
      __syncthreads();   // one barrier per tile: the buffer of tile i is rewritten for tile i+2, behind barrier i+1
      if (tid < MR * 4) {    // thread -> (row, column pair); fixed summation order over the warps
        const float* rr = sRed + (c_seg & 1) * GS_WARPS * (MR * 8) + tid * 2;
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int ww = 0; ww < GS_WARPS; ++ww) {
          const float2 pv = *reinterpret_cast<const float2*>(rr + ww * (MR * 8));
          v0 += pv.x;
          v1 += pv.y;
        }
        gs_emit(d, tid >> 2, tile * 8 + 2 * (tid & 3), v0, v1);
      }
    } else {
      gs_emit(d, gq, tile * 8 + 2 * tq, acc[0], acc[1]);
      if (MR == 16) gs_emit(d, gq + 8, tile * 8 + 2 * tq, acc[2], acc[3]);
    }
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
  };
  if (SPLIT && seg_len == 0) {   // fewer k32 blocks than warps: this warp only takes part in the barriers
    for (int sg = 0; sg < ntiles; ++sg) finish(sg);
    return;
  }
  int c_seg = 0, c_kb = kb0;
  int done = 0;
  while (done < total) {
#pragma unroll
    for (int u = 0; u < GS_RING; ++u) {
      if (done < total) {                                  // warp-uniform
        const uint4 w = ring[u];
        ring[u] = next_load();
        const uint4 x = *reinterpret_cast<const uint4*>(arow + c_kb * 32);
        uint4 y = make_uint4(0u, 0u, 0u, 0u);
        if (MR == 16) y = *reinterpret_cast<const uint4*>(arow_hi + c_kb * 32);
        gs_mma16816(acc, x.x, y.x, x.y, y.y, w.x, w.y);
        gs_mma16816(acc, x.z, y.z, x.w, y.w, w.z, w.w);
        ++done;
        if (++c_kb == kb1) {
          finish(c_seg);
          c_kb = kb0;
          ++c_seg;
        }
      }
    }
  }
}

}  // namespace

// W (N, K) fp32 row-major -> tile-blocked fp16 [ceil(N/8)][K/32][8][32] (rows >= N are zero)
void gemm_stream_pack_host(const float* W, int64_t N, int64_t K, __half* out) {
  const int64_t T = (N + 7) / 8, nb = K / 32;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < T; ++j)
    for (int64_t kb = 0; kb < nb; ++kb)
      for (int g = 0; g < 8; ++g) {
        __half* dst = out + ((j * nb + kb) * 8 + g) * 32;
        const int64_t n = 8 * j + g;
        for (int kk = 0; kk < 32; ++kk) {
          const float v = n < N ? W[n * K + kb * 32 + kk] : 0.f;
          dst[kk] = __float2half_rn(std::min(65504.0f, std::max(-65504.0f, v)));
        }
      }
}

size_t gemm_stream_packed_elems(int64_t N, int64_t K) { return static_cast<size_t>((N + 7) / 8) * 8 * K; }

// rows of width K the activation tile can hold next to the reduction buffers (<= 16)
int gemm_stream_max_rows(int K) {
  const long long room = 200 * 1024 - 2 * GS_WARPS * 128 * 4;
  return static_cast<int>(std::min<long long>(16, room / (static_cast<long long>(K + GS_PAD) * 2)));
}

bool gemm_stream_applicable(int N, int K, int B) {
  return B >= 1 && K % 32 == 0 && K >= 32 && B <= gemm_stream_max_rows(K);
}

int gemm_stream_init() {
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_stream_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_stream_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_stream_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_stream_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return FRT2_OK;
}

int gemm_stream(const StreamGemm& d, cudaStream_t stream) {
  FRT2_REQUIRE(gemm_stream_applicable(d.N, d.K, d.B), FRT2_ERR_BAD_ARG, "gemm_stream: K a multiple of 32 and at most gemm_stream_max_rows(K) <= 16 rows required");
  FRT2_REQUIRE(d.gamma != nullptr ? (d.x != nullptr && d.ldx % 4 == 0) : (d.A != nullptr && d.lda % 8 == 0), FRT2_ERR_BAD_ARG,
               "gemm_stream: activation operand missing or misaligned");
  static const int sms = num_sms();
  const int T = (d.N + 7) / 8;
  const int MR = d.B > 8 ? 16 : 8;
  const size_t smem = static_cast<size_t>(d.B) * (d.K + GS_PAD) * 2 + 2 * GS_WARPS * (MR * 8) * 4;
  // GS_CTAS_PER_SM CTAs per SM while their shared memory fits next to the same number of CTAs of the NEXT launch
  const int per_sm = (GS_CTAS_PER_SM > 1 && smem * 2 * GS_CTAS_PER_SM <= 200 * 1024) ? GS_CTAS_PER_SM : 1;
  const int ctas = sms * per_sm;
  const bool split = T < ctas * GS_WARPS;         // fewer column tiles than warps in the grid: the CTA's warps split K
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(split ? std::min(ctas, T) : ctas);
  cfg.blockDim = dim3(GS_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool use_pdl = (getenv("FRT2_NO_PDL") == nullptr);
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  if (MR == 8) {
    if (split) FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_stream_kernel<true, 8>, d));
    else FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_stream_kernel<false, 8>, d));
  } else {
    if (split) FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_stream_kernel<true, 16>, d));
    else FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_stream_kernel<false, 16>, d));
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
