// K3 — attention of the causal Vocos backbone (reference whisper.py:49-118, mask utils.py:19-38).
//
//   offline : block-causal, key j visible to query i  iff  j <= (i | 7)   (8-frame blocks == one codec token);
//             the dense (b,t,t) bool mask the reference builds on the host is replaced by this predicate.
//   chunked : queries are the new frames at absolute positions q_pos0.., keys = KV state ++ chunk, no mask
//             inside the chunk (reference forward_chunk passes attn_mask=None).
//
// attention_tc   : tcgen05 flash attention, hd = 64 (attention_t3_kernel below).  64-key tiles, S = Q K^T and P V on
//                  the tensor cores with fp32 accumulators in TMEM; softmax in fp32 with one thread per query row
//                  (tcgen05.ld gives a thread its whole row: no shuffles); P stays in tensor memory (TS-form P V);
//                  V is consumed straight from its natural (key, d) layout as an MN-major operand.
// attention_warp : CUDA-core kernel, one warp per 8-query block (all 8 rows of a block see the same keys).
//                  Used for the short-query streaming step and as the on-device check for attention_tc.
#include <math_constants.h>

#include <cstdlib>

#include <mutex>

#include "attention_warp_body.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace frt2 {

// =====================================================================================================
// attention_warp
// =====================================================================================================
namespace {

template <int HD, int NSPLIT, int KS = 1>
__global__ void __launch_bounds__(NSPLIT == 1 ? 128 : NSPLIT * 32) attention_warp_kernel(AttnDesc a) {
  pdl_trigger();   // a following PDL-launched kernel (streaming skinny GEMM) may start its weight prefetch
  attention_warp_body<HD, NSPLIT, false, KS>(a, blockIdx.x);
}

}  // namespace

int attention_warp(const AttnDesc& a, cudaStream_t stream) {
  FRT2_REQUIRE(a.Tq % 8 == 0 && a.q_pos0 % 8 == 0, FRT2_ERR_BAD_ARG, "attention: Tq and q_pos0 must be multiples of 8");
  FRT2_REQUIRE(a.kv_row_pitch % 8 == 0, FRT2_ERR_BAD_ARG, "attention: K/V row pitch must be a multiple of 8");
  const long long warps = static_cast<long long>(a.B) * a.H * (a.Tq / 8);
  if (warps == 0) return FRT2_OK;
  constexpr int NS = 16;
  // few query blocks (the streaming step): 16 warps per block split the KV state; otherwise one warp per block
  const bool split = warps <= 1024 && a.hd == 64;
  const unsigned grid = static_cast<unsigned>((warps + 3) / 4);
  if (split && a.part != nullptr && a.part_count != nullptr && a.Tq == 8) {
    // a long K/V state behind 8 queries: ATTN_KSPLIT CTAs per (item, head), merged by the last one to finish
    attention_warp_kernel<64, NS, ATTN_KSPLIT><<<static_cast<unsigned>(warps * ATTN_KSPLIT), NS * 32, 0, stream>>>(a);
  } else if (split) {
    attention_warp_kernel<64, NS><<<static_cast<unsigned>(warps), NS * 32, 0, stream>>>(a);
  } else {
    switch (a.hd) {
      case 32: attention_warp_kernel<32, 1><<<grid, 128, 0, stream>>>(a); break;
      case 64: attention_warp_kernel<64, 1><<<grid, 128, 0, stream>>>(a); break;
      case 128: attention_warp_kernel<128, 1><<<grid, 128, 0, stream>>>(a); break;
      default:
        set_error("attention: head_dim must be 32, 64 or 128");
        return FRT2_ERR_BAD_ARG;
    }
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// =====================================================================================================
// attention_tc  (hd = 64)
// =====================================================================================================
namespace {

constexpr int AT_BQ = 128;       // query rows per query tile == TMEM lanes
constexpr int AT_BK = 64;        // keys per tile
constexpr int AT_HD = 64;
constexpr int AT_Q_BYTES = AT_BQ * AT_HD * 2;   // 16 KB per query tile
constexpr int AT_KV_BYTES = AT_BK * AT_HD * 2;  // 8 KB

__device__ __forceinline__ float fast_exp2(float x) {  // MUFU.EX2; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (x0, x1) fp32 -> packed halves (lo = 2^x0, hi = 2^x1); the conversion to half happens BEFORE the exponential so
// the result is directly the fp16 P operand
__device__ __forceinline__ uint32_t exp2_f16x2(float x0, float x1) {
#ifdef FRT2_EXP_F16X2
  uint32_t h, y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(h));
  return y;
#else
  uint32_t y;
  const float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(e1), "f"(e0));
  return y;
#endif
}

struct AttnKParams {
  int Tq, Tk, q_pos0, block_causal, H;
  float scale_log2;
  __half* out;
  long long o_row_pitch, o_batch_pitch;
};

// =====================================================================================================
// attention_t3: one CTA per SM, THREE query tiles that share every K/V tile, S / P / O of each tile in separate
// tensor-memory columns (S 64 | P 32 | O 64 = 160 columns per tile, 480 of the SM's 512).
// Why: with P aliased over S or a single smem P tile (the round-1 kernels, removed) a query tile's softmax cannot start
// on key tile j+1 before the MMA thread has seen its P_j, issued P_j V_j and Q K_{j+1}^T, and those have executed — a
// ~1400 clk round trip per key tile that four resident tiles only partly hide (profiles/r01_attention_ab.txt).  Here
// Q K_{j+1}^T is issued as soon as the softmax has READ S_j (s_empty), so S_{j+1} is waiting when the softmax of key
// tile j ends; the exponentials are computed into registers first and only the tcgen05.st of P_j waits for P_{j-1} V_{j-1}
// (issued a whole softmax period earlier).  The softmax warps never wait on the tensor pipe in steady state, and three
// tiles (12 softmax warps, 3 per SM sub-partition) keep the MUFU busy.
// =====================================================================================================
constexpr int A3_QT = 3;
constexpr int A3_STAGES = 4;
constexpr int A3_TILE_COLS = 160;                 // S [0,64) | P [64,96) | O [96,160)
constexpr int A3_P_COL = 64, A3_O_COL = 96;
constexpr int A3_TMEM_COLS = 512;
constexpr int A3_SOFTMAX_WARPS = 4 * A3_QT;
constexpr int A3_THREADS = (A3_SOFTMAX_WARPS + 2) * 32;
constexpr int A3_SMEM_BYTES = A3_QT * AT_Q_BYTES + 2 * A3_STAGES * AT_KV_BYTES + 1024 + 512;

__global__ void __launch_bounds__(A3_THREADS, 1)
attention_t3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // A3_QT tiles
  uint8_t* sK = sQ + A3_QT * AT_Q_BYTES;
  uint8_t* sV = sK + A3_STAGES * AT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A3_STAGES * AT_KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + A3_STAGES;
  uint64_t* v_full = k_empty + A3_STAGES;
  uint64_t* v_empty = v_full + A3_STAGES;
  uint64_t* s_full = v_empty + A3_STAGES;   // [A3_QT]
  uint64_t* s_empty = s_full + A3_QT;
  uint64_t* p_full = s_empty + A3_QT;
  uint64_t* pv_done = p_full + A3_QT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + A3_QT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qgrp = gridDim.x - 1 - blockIdx.x;   // heaviest (latest) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qgrp * (AT_BQ * A3_QT);

  int nt[A3_QT];   // number of 64-key tiles each query tile needs (0 = tile lies beyond the sequence); non-decreasing in t
  int ntiles = 0;
#pragma unroll
  for (int t = 0; t < A3_QT; ++t) {
    const int qs = q0 + t * AT_BQ;
    if (qs >= p.Tq) {
      nt[t] = 0;
    } else {
      int kmax = p.Tk - 1;
      if (p.block_causal) kmax = min(kmax, (p.q_pos0 + min(qs + AT_BQ, p.Tq) - 1) | 7);
      nt[t] = kmax / AT_BK + 1;
    }
    ntiles = max(ntiles, nt[t]);
  }

  if (warp == A3_SOFTMAX_WARPS && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if (warp == A3_SOFTMAX_WARPS + 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < A3_STAGES; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < A3_QT; ++t) {
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&s_empty[t], 4);
      ptx::mbar_init(&p_full[t], 4);
      ptx::mbar_init(&pv_done[t], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == A3_SOFTMAX_WARPS + 1) {
    ptx::tmem_alloc(tmem_slot, A3_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == A3_SOFTMAX_WARPS) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, A3_QT * AT_Q_BYTES);
#pragma unroll
      for (int t = 0; t < A3_QT; ++t) ptx::tma_load_3d(sQ + t * AT_Q_BYTES, &tmQ, q_full, h * AT_HD, q0 + t * AT_BQ, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        ptx::mbar_wait(&k_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&k_full[stage], AT_KV_BYTES);
        ptx::tma_load_3d(sK + stage * AT_KV_BYTES, &tmK, &k_full[stage], h * AT_HD, j * AT_BK, b);
        ptx::mbar_wait(&v_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&v_full[stage], AT_KV_BYTES);
        ptx::tma_load_3d(sV + stage * AT_KV_BYTES, &tmV, &v_full[stage], h * AT_HD, j * AT_BK, b);
        if (++stage == A3_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == A3_SOFTMAX_WARPS + 1) {
    // ------------------------------------------------------------ MMA issuer (all query tiles)
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_f16(AT_BQ, AT_BK, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(AT_BQ, AT_HD, 0, 1);  // P (TMEM)    x V (MN-major)
      ptx::mbar_wait(q_full, 0);
      for (int j = 0; j <= ntiles; ++j) {        // round j: Q K_j^T for every tile, then P_{j-1} V_{j-1}
        if (j < ntiles) {
          const int st = j % A3_STAGES;
          ptx::mbar_wait(&k_full[st], (j / A3_STAGES) & 1);
          const uint64_t dk = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sK + st * AT_KV_BYTES));
#pragma unroll
          for (int t = 0; t < A3_QT; ++t) {
            if (j < nt[t]) {
              if (j > 0) ptx::mbar_wait(&s_empty[t], (j - 1) & 1);   // the softmax has read S_{j-1} (early in its period)
              ptx::tc_fence_after();
              const uint64_t dq = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sQ + t * AT_Q_BYTES));
#pragma unroll
              for (int k = 0; k < AT_HD / 16; ++k)
                ptx::mma_f16_ss(tmem_base + t * A3_TILE_COLS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
              ptx::mma_commit(&s_full[t]);
            }
          }
          ptx::mma_commit(&k_empty[st]);
        }
        if (j > 0) {
          const int jp = j - 1, st = jp % A3_STAGES;
          ptx::mbar_wait(&v_full[st], (jp / A3_STAGES) & 1);
          const uint32_t vaddr = ptx::smem_u32(sV + st * AT_KV_BYTES);
          uint32_t pending = 0;
#pragma unroll
          for (int t = 0; t < A3_QT; ++t)
            if (jp < nt[t]) pending |= 1u << t;
          uint32_t spins = 0;
          while (pending) {                      // serve whichever tile's softmax finishes first
#pragma unroll
            for (int t = 0; t < A3_QT; ++t) {
              if (!(pending & (1u << t)) || !ptx::mbar_test(&p_full[t], jp & 1)) continue;
              ptx::tc_fence_after();
#pragma unroll
              for (int k = 0; k < AT_BK / 16; ++k) {
                const uint64_t dv = ptx::make_desc_mnmajor_sw128(vaddr + k * 2048, 1024, 1024);
                ptx::mma_f16_ts(tmem_base + t * A3_TILE_COLS + A3_O_COL, tmem_base + t * A3_TILE_COLS + A3_P_COL + k * 8,
                                dv, idesc_pv, (jp | k) != 0 ? 1u : 0u);
              }
              ptx::mma_commit(&pv_done[t]);
              pending &= ~(1u << t);
            }
            if (++spins > (1u << 26)) __trap();
          }
          ptx::mma_commit(&v_empty[st]);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax / output (thread == query row)
    const int t = warp >> 2;                       // query tile of this warpgroup
    const int r = threadIdx.x & (AT_BQ - 1);       // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + t * A3_TILE_COLS;
    const uint32_t tmem_p = tmem_s + A3_P_COL;
    const uint32_t tmem_o = tmem_s + A3_O_COL;
    const int my_tiles = nt[t];
    const int qi = q0 + t * AT_BQ + r;             // row inside this item's query block
    const int qabs = p.q_pos0 + qi;                // absolute position
    int limit = p.Tk - 1;
    if (p.block_causal) limit = min(limit, qabs | 7);
    constexpr float RESCALE_LOG2 = 8.0f;           // lazy rescaling: O is rescaled only when a row max grows by more than 2^8
    float m_ref = 0.f, l = 0.f;
    const uint32_t a_s_full = ptx::smem_u32(&s_full[t]), a_s_empty = ptx::smem_u32(&s_empty[t]);
    const uint32_t a_p_full = ptx::smem_u32(&p_full[t]), a_pv_done = ptx::smem_u32(&pv_done[t]);

    for (int j = 0; j < my_tiles; ++j) {
      ptx::mbar_wait_likely_ready(a_s_full, j & 1);
      ptx::tc_fence_after();
      uint32_t sa[32], sb[32];
      ptx::tmem_ld32(tmem_s + lane_off, sa);
      ptx::tmem_ld32(tmem_s + lane_off + 32, sb);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_s_empty);   // Q K_{j+1}^T may overwrite S now

      const int lim = limit - j * AT_BK;  // columns c <= lim are visible
      if (lim < AT_BK - 1) {              // diagonal / last tile: mask (interior tiles skip this entirely)
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c > lim) sa[c] = 0xff800000u;        // -inf
          if (c + 32 > lim) sb[c] = 0xff800000u;
        }
      }
      float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};  // 4 independent chains
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(sa[c + u]), __uint_as_float(sb[c + u])));
      }
      const float m_tile = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;  // scale > 0

      if (j == 0) {
        m_ref = (m_tile == -CUDART_INF_F) ? 0.f : m_tile;
      } else {
        const bool need = m_tile > m_ref + RESCALE_LOG2;
        if (__any_sync(0xffffffffu, need)) {
          // O must be quiescent: P_{j-1} V_{j-1} (the last one issued for this tile) has completed
          ptx::mbar_wait(a_pv_done, (j - 1) & 1);
          ptx::tc_fence_after();
          const float alpha = need ? fast_exp2(m_ref - m_tile) : 1.0f;
          if (need) m_ref = m_tile;
          l *= alpha;
#pragma unroll 1
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t tt[16];
            ptx::tmem_ld16(tmem_o + lane_off + q4 * 16, tt);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) tt[c] = __float_as_uint(__uint_as_float(tt[c]) * alpha);
            ptx::tmem_st16(tmem_o + lane_off + q4 * 16, tt);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
        }
      }
      // p = 2^(s*scale - m_ref): all 64 exponentials into registers (packed halves) BEFORE touching the P columns
      const float neg_m = -m_ref;
      unsigned long long scale2, negm2;
      asm("mov.b64 %0, {%1, %1};" : "=l"(scale2) : "f"(p.scale_log2));
      asm("mov.b64 %0, {%1, %1};" : "=l"(negm2) : "f"(neg_m));
      uint32_t w[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {       // pair e = keys (2e, 2e+1)
        const uint32_t s0 = (e < 16) ? sa[2 * e] : sb[2 * e - 32];
        const uint32_t s1 = (e < 16) ? sa[2 * e + 1] : sb[2 * e - 31];
        unsigned long long xx;
        asm("{\n\t.reg .b64 a;\n\tmov.b64 a, {%1, %2};\n\tfma.rn.f32x2 %0, a, %3, %4;\n\t}"
            : "=l"(xx) : "r"(s0), "r"(s1), "l"(scale2), "l"(negm2));
        float x0, x1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
        w[e] = exp2_f16x2(x0, x1);
      }
      __half2 acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = *reinterpret_cast<const __half2*>(&w[u]);
#pragma unroll
      for (int e = 4; e < 32; ++e) acc[e & 3] = __hadd2(acc[e & 3], *reinterpret_cast<const __half2*>(&w[e]));
      {
        const float2 f0 = __half22float2(acc[0]), f1 = __half22float2(acc[1]);
        const float2 f2 = __half22float2(acc[2]), f3 = __half22float2(acc[3]);
        l += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
      }
      // the P columns are free once P_{j-1} V_{j-1} has completed (issued a whole softmax period ago)
      if (j > 0) {
        ptx::mbar_wait_likely_ready(a_pv_done, (j - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::tmem_st32(tmem_p + lane_off, w);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_p_full);
    }
    if (my_tiles > 0) {
      ptx::mbar_wait(a_pv_done, (my_tiles - 1) & 1);
      ptx::tc_fence_after();
      const float inv = 1.0f / l;
      __half* op = p.out + b * p.o_batch_pitch + static_cast<long long>(qi) * p.o_row_pitch + h * AT_HD;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t tt[32];
        ptx::tmem_ld32(tmem_o + lane_off + half * 32, tt);
        ptx::tmem_ld_wait();
        if (qi < p.Tq) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 q;
            q.x = pack_half2(__uint_as_float(tt[c + 0]) * inv, __uint_as_float(tt[c + 1]) * inv);
            q.y = pack_half2(__uint_as_float(tt[c + 2]) * inv, __uint_as_float(tt[c + 3]) * inv);
            q.z = pack_half2(__uint_as_float(tt[c + 4]) * inv, __uint_as_float(tt[c + 5]) * inv);
            q.w = pack_half2(__uint_as_float(tt[c + 6]) * inv, __uint_as_float(tt[c + 7]) * inv);
            *reinterpret_cast<uint4*>(op + half * 32 + c) = q;
          }
        }
      }
      ptx::tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == A3_SOFTMAX_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, A3_TMEM_COLS);
  }
}

std::once_flag g_attn_once;
int g_attn_status = FRT2_OK;

}  // namespace

int attention_tc_init() {
  std::call_once(g_attn_once, [] {
    g_attn_status = gemm_tc_init();
    if (g_attn_status != FRT2_OK) return;
    cudaError_t e = cudaFuncSetAttribute(attention_t3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error(std::string("cudaFuncSetAttribute(attention_t3_kernel): ") + cudaGetErrorString(e));
      g_attn_status = FRT2_ERR_CUDA;
    }
  });
  return g_attn_status;
}

int attention_tc(const AttnDesc& a, cudaStream_t stream) {
  FRT2_TRY(attention_tc_init());
  FRT2_REQUIRE(a.hd == AT_HD, FRT2_ERR_BAD_ARG, "attention_tc: head_dim must be 64");
  FRT2_REQUIRE(a.Tq >= 1 && a.Tk >= 1, FRT2_ERR_BAD_ARG, "attention_tc: empty sequence");
  FRT2_REQUIRE(a.q_row_pitch % 8 == 0 && a.kv_row_pitch % 8 == 0 && a.q_batch_pitch % 8 == 0 &&
                   a.kv_batch_pitch % 8 == 0 && a.o_row_pitch % 8 == 0 && a.o_batch_pitch % 8 == 0,
               FRT2_ERR_BAD_ARG, "attention_tc: pitches must be multiples of 8 elements");
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t cols = static_cast<uint64_t>(a.H) * a.hd;
  {
    uint64_t dims[3] = {cols, static_cast<uint64_t>(a.Tq), static_cast<uint64_t>(a.B)};
    uint64_t strides[2] = {static_cast<uint64_t>(a.q_row_pitch) * 2,
                           static_cast<uint64_t>(a.B > 1 ? a.q_batch_pitch : a.q_row_pitch * a.Tq) * 2};
    uint32_t box[3] = {AT_HD, AT_BQ, 1};
    FRT2_TRY(tma_encode_fp16(&tmQ, a.q, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {cols, static_cast<uint64_t>(a.Tk), static_cast<uint64_t>(a.B)};
    uint64_t strides[2] = {static_cast<uint64_t>(a.kv_row_pitch) * 2,
                           static_cast<uint64_t>(a.B > 1 ? a.kv_batch_pitch : a.kv_row_pitch * a.Tk) * 2};
    uint32_t box[3] = {AT_HD, AT_BK, 1};
    FRT2_TRY(tma_encode_fp16(&tmK, a.k, 3, dims, strides, box));
    FRT2_TRY(tma_encode_fp16(&tmV, a.v, 3, dims, strides, box));
  }
  AttnKParams p;
  p.Tq = a.Tq;
  p.Tk = a.Tk;
  p.q_pos0 = a.q_pos0;
  p.block_causal = a.block_causal;
  p.H = a.H;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.out = a.out;
  p.o_row_pitch = a.o_row_pitch;
  p.o_batch_pitch = a.o_batch_pitch;
  dim3 grid3((a.Tq + AT_BQ * A3_QT - 1) / (AT_BQ * A3_QT), a.H, a.B);
  attention_t3_kernel<<<grid3, A3_THREADS, A3_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
