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

// 2^x for a packed pair on the FMA pipe (Cody-Waite: x = n + f, n = round(x), |f| <= 0.5; 2^f by a degree-3 minimax
// polynomial, rel. err 7.5e-5 — below the fp16 rounding of P; 2^n by adding n to the exponent field).  x >= -125 after
// the clamp, so masked scores (-inf) come out as 2^-125 -> 0 in fp16.
__device__ __forceinline__ uint32_t exp2_poly_f16x2(unsigned long long xx) {
  float x0, x1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  unsigned long long x, t, n, f, pl, magic2, nmagic2, c3, c2, c1, c0, neg1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(magic2) : "f"(12582912.0f));   // 1.5 * 2^23: rounds to the nearest integer
  asm("mov.b64 %0, {%1, %1};" : "=l"(nmagic2) : "f"(-12582912.0f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c3) : "f"(0.055179595f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(0.242611851f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c1) : "f"(0.693259533f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c0) : "f"(0.999927984f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(neg1) : "f"(-1.0f));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x), "l"(magic2));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(n) : "l"(t), "l"(nmagic2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f) : "l"(n), "l"(neg1), "l"(x));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pl) : "l"(c3), "l"(f), "l"(c2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pl) : "l"(pl), "l"(f), "l"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pl) : "l"(pl), "l"(f), "l"(c0));
  uint32_t t0, t1, p0, p1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(pl));
  const float r0 = __uint_as_float(p0 + (t0 << 23));   // exponent += n  (the MAGIC bits shift out)
  const float r1 = __uint_as_float(p1 + (t1 << 23));
  uint32_t y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(r1), "f"(r0));
  return y;
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


// =====================================================================================================
// attention_t4: PERSISTENT form of attention_t3 (one CTA per SM for the whole launch), head dim 64 (3 query tiles per
// CTA) or 128 (2 query tiles; S 64 | P 32 | O 128 columns per tile).
//   * work items (query group, head, item) come from a device counter in (item, head)-major order with the heaviest
//     (latest) query group of a head first: the CTAs that run concurrently work on neighbouring groups of the same
//     head, so its K/V is read from HBM once; the tail of the launch is one item long at most.
//   * the K/V ring, the barrier phases and the tensor-memory columns live across items: the TMA thread keeps loading
//     key tiles of the next item while the slowest query tile finishes the current one, the MMA thread runs one
//     independent state machine per query tile (a tile whose causal range ends earlier starts the next item earlier),
//     the per-CTA prologue (barrier init, tensor-memory allocation, first Q/K round trip) is paid once per SM instead
//     of once per 384 query rows (it was 9-13 % of the t3 kernel, profiles/r01_attention_ab.txt).
//   * one MMA-issuing warp per query tile (a single issuing thread for three tiles was the hidden bottleneck of t3 and
//     of the first persistent version: ~300 clocks per 4-MMA action, 6 actions per ~1700-clock round).
//   * softmax: the two barrier probes of a key tile sit inside the exponential stream, and EMU of every 8 score pairs
//     take their 2^x from a degree-3 polynomial on the FMA pipe (packed fp32x2) instead of the MUFU, which at head dim
//     64 is the binding unit (64 exponentials per row and key tile at 16 per clock per SM = twice the tensor time).
//     Measured (B=64, H=16, T=3000, isolated): t3 1.868 ms -> 1.72 ms (persistent, per-tile issuers, probes) -> 1.54-1.58 ms
//     with a quarter of the exponentials on the FMA pipe (profiles/r02_attention_ab.txt).
// =====================================================================================================
template <int HD, int QT>
struct A4 {
  static constexpr int NSUB = HD / 64;                 // 64-column (128 B, SWIZZLE_128B) sub-tiles per row
  static constexpr int STAGES = (HD == 64) ? 6 : 4;
  static constexpr int Q_BYTES = AT_BQ * HD * 2;
  static constexpr int KV_BYTES = AT_BK * HD * 2;
  static constexpr int P_COL = 64, O_COL = 96;
  static constexpr int TILE_COLS = 96 + HD;            // S [0,64) | P [64,96) | O [96, 96+HD)
  static constexpr int SOFTMAX_WARPS = 4 * QT;
  static constexpr int THREADS = (SOFTMAX_WARPS + 1 + QT) * 32;   // softmax | TMA | one MMA issuer per query tile
  static constexpr int NI = 4;                         // item ring slots
  static constexpr int SMEM_BYTES = QT * Q_BYTES + 2 * STAGES * KV_BYTES + 1024 + 1024;
  static_assert(QT * TILE_COLS <= 512, "tensor memory");
};

struct Attn4Params {
  int Tq, Tk, q_pos0, block_causal, H, groups, total;
  float scale_log2;
  __half* out;
  long long o_row_pitch, o_batch_pitch;
  unsigned int* sched;   // [0] next item, [1] CTAs finished (the last one to finish re-arms both)
  int tma_sleep_ns;                 // back-off of the polling TMA thread when a sweep found nothing to do
  int one_item_per_cta;             // A/B: grid == items, no device counter (the t3 launch shape on this kernel)
  uint32_t* trace;                  // TRACE builds: clock64 stamps of CTA 0, [warp][key-tile index < 128][8 points]
};

// number of 64-key tiles query tile t of the group starting at row q0 needs (0: the tile lies beyond the sequence)
__device__ __forceinline__ int a4_tile_keys(const Attn4Params& p, int q0, int t) {
  const int qs = q0 + t * AT_BQ;
  if (qs >= p.Tq) return 0;
  int kmax = p.Tk - 1;
  if (p.block_causal) kmax = min(kmax, (p.q_pos0 + min(qs + AT_BQ, p.Tq) - 1) | 7);
  return kmax / AT_BK + 1;
}

template <int HD, int QT, bool PROBE, bool TRACE, int EMU>
__global__ void __launch_bounds__(A4<HD, QT>::THREADS, 1)
attention_t4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const Attn4Params p) {
  using C = A4<HD, QT>;
  constexpr int STAGES = C::STAGES, NI = C::NI, NSUB = C::NSUB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + QT * C::Q_BYTES;
  uint8_t* sV = sK + STAGES * C::KV_BYTES;
  int4* items = reinterpret_cast<int4*>(sV + STAGES * C::KV_BYTES);   // [NI] (16-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(items + NI);
  uint64_t* q_full = bars;                  // [QT]
  uint64_t* q_empty = q_full + QT;          // [QT]
  uint64_t* s_full = q_empty + QT;          // [QT]
  uint64_t* s_empty = s_full + QT;
  uint64_t* p_full = s_empty + QT;
  uint64_t* pv_done = p_full + QT;
  uint64_t* o_free = pv_done + QT;
  uint64_t* k_full = o_free + QT;           // [STAGES]
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* item_full = v_empty + STAGES;   // [NI]
  uint64_t* item_empty = item_full + NI;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(item_empty + NI);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = C::SOFTMAX_WARPS, W_MMA = C::SOFTMAX_WARPS + 1;

  if (warp == W_TMA && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if (warp == W_MMA && lane == 0) {
    for (int t = 0; t < QT; ++t) {
      ptx::mbar_init(&q_full[t], 1);
      ptx::mbar_init(&q_empty[t], 1);
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&s_empty[t], 4);
      ptx::mbar_init(&p_full[t], 4);
      ptx::mbar_init(&pv_done[t], 1);
      ptx::mbar_init(&o_free[t], 4);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], QT);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], QT);
    }
    for (int s = 0; s < NI; ++s) {
      ptx::mbar_init(&item_full[s], 1);
      ptx::mbar_init(&item_empty[s], QT + C::SOFTMAX_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == W_MMA) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto stamp = [&](int w, uint32_t idx, int point) {
    if (TRACE && blockIdx.x == 0 && idx < 128u) p.trace[(w * 128 + idx) * 8 + point] = static_cast<uint32_t>(clock64());
  };

  if (warp == W_TMA) {
    // ------------------------------------------------------------ scheduler + TMA producer
    if (ptx::elect_one()) {
      uint32_t gpos = 0;   // key tiles loaded so far (ring position)
      for (uint32_t seq = 0;; ++seq) {
        const int slot = seq % NI;
        ptx::mbar_wait(&item_empty[slot], ((seq / NI) & 1) ^ 1);
        const unsigned int w = p.one_item_per_cta ? (seq == 0 ? blockIdx.x : static_cast<unsigned int>(p.total))
                                                  : atomicAdd(p.sched, 1u);
        int4 it;
        if (w >= static_cast<unsigned int>(p.total)) {
          it = make_int4(-1, 0, 0, 0);
        } else {
          const int bh = static_cast<int>(w) / p.groups, g = p.groups - 1 - static_cast<int>(w) % p.groups;
          it = make_int4(g * (AT_BQ * QT), bh % p.H, bh / p.H, 0);
        }
        ptx::st_shared_v4(ptx::smem_u32(&items[slot]), it.x, it.y, it.z, it.w);
        ptx::mbar_arrive(&item_full[slot]);
        if (it.x < 0) break;
        int ntiles = 0;
#pragma unroll
        for (int t = 0; t < QT; ++t) ntiles = max(ntiles, a4_tile_keys(p, it.x, t));
        uint32_t qpend = (1u << QT) - 1;
        int j = 0;
        uint32_t spins = 0;
        while (qpend != 0 || j < ntiles) {
          bool progress = false;
#pragma unroll
          for (int t = 0; t < QT; ++t) {
            if (((qpend >> t) & 1u) && ptx::mbar_test(&q_empty[t], (seq & 1) ^ 1)) {
              ptx::mbar_expect_tx(&q_full[t], C::Q_BYTES);
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)
                ptx::tma_load_3d(sQ + t * C::Q_BYTES + sub * (AT_BQ * 128), &tmQ, &q_full[t], it.y * HD + sub * 64,
                                 it.x + t * AT_BQ, it.z);
              qpend &= ~(1u << t);
              progress = true;
            }
          }
          if (j < ntiles) {
            const uint32_t st = gpos % STAGES, ph = (gpos / STAGES) & 1;
            if (ptx::mbar_test(&k_empty[st], ph ^ 1) && ptx::mbar_test(&v_empty[st], ph ^ 1)) {
              ptx::mbar_expect_tx(&k_full[st], C::KV_BYTES);
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)
                ptx::tma_load_3d(sK + st * C::KV_BYTES + sub * (AT_BK * 128), &tmK, &k_full[st], it.y * HD + sub * 64,
                                 j * AT_BK, it.z);
              ptx::mbar_expect_tx(&v_full[st], C::KV_BYTES);
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)
                ptx::tma_load_3d(sV + st * C::KV_BYTES + sub * (AT_BK * 128), &tmV, &v_full[st], it.y * HD + sub * 64,
                                 j * AT_BK, it.z);
              ++j;
              ++gpos;
              progress = true;
            }
          }
          if (progress) {
            spins = 0;
          } else {
            if (p.tma_sleep_ns > 0) __nanosleep(p.tma_sleep_ns);
            if (++spins > (1u << 24)) __trap();
          }
        }
      }
      // every CTA fetches exactly one item beyond the end; the last CTA to get here re-arms the counters
      if (!p.one_item_per_cta && atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) {
        p.sched[0] = 0;
        p.sched[1] = 0;
      }
    }
    __syncwarp();
  } else if (warp >= W_MMA) {
    // ------------------------------------------------------------ MMA issuers: one warp per query tile.
    // (One thread issuing for all tiles was the bottleneck of the first persistent version: ~300 clocks per
    // 4-MMA action incl. the descriptor moves to uniform registers, 6 actions per round of ~1700 clocks.)
    // Ordered stream of a tile: QK_0, QK_1, PV_0, QK_2, PV_1, ...; QK_{j+1} needs S_j read (early in the softmax of key
    // tile j), PV_j needs P_j (its end).  Every K/V ring position of an item is released by every tile, used or not.
    const int t = warp - W_MMA;
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_f16(AT_BQ, AT_BK, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(AT_BQ, 64, 0, 1);     // P (TMEM)    x V (MN-major), 64 columns
      const uint32_t tmem_s = tmem_base + t * C::TILE_COLS;
      uint32_t cum = 0, kb = 0;
      for (uint32_t n = 0;; ++n) {
        const int slot = n % NI;
        ptx::mbar_wait(&item_full[slot], (n / NI) & 1);
        const int4 it = ptx::ld_shared_v4(ptx::smem_u32(&items[slot]));
        if (it.x < 0) break;
        const int nt = a4_tile_keys(p, it.x, t);
        int ntl = 0;
#pragma unroll
        for (int u = 0; u < QT; ++u) ntl = max(ntl, a4_tile_keys(p, it.x, u));
        ptx::mbar_wait(&q_full[t], n & 1);
        if (nt == 0) ptx::mma_commit(&q_empty[t]);
        for (int j = 0; j <= ntl; ++j) {
          if (j < ntl) {                               // K side of ring position kb + j
            const uint32_t pos = kb + j, st = pos % STAGES;
            ptx::mbar_wait(&k_full[st], (pos / STAGES) & 1);
            if (j < nt) {
              if (cum + j > 0) ptx::mbar_wait(&s_empty[t], (cum + j - 1) & 1);
              ptx::tc_fence_after();
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub) {
                const uint64_t dq = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sQ + t * C::Q_BYTES + sub * (AT_BQ * 128)));
                const uint64_t dk = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sK + st * C::KV_BYTES + sub * (AT_BK * 128)));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, (sub | k) != 0 ? 1u : 0u);
              }
              ptx::mma_commit(&s_full[t]);
              stamp(W_TMA + t, cum + j, 0);
              if (j == nt - 1) ptx::mma_commit(&q_empty[t]);   // Q of this tile may be replaced by the next item's
            }
            ptx::mma_commit(&k_empty[st]);
          }
          if (j > 0) {                                 // V side of ring position kb + j - 1
            const int jp = j - 1;
            const uint32_t pos = kb + jp, st = pos % STAGES;
            ptx::mbar_wait(&v_full[st], (pos / STAGES) & 1);
            if (jp < nt) {
              ptx::mbar_wait(&p_full[t], (cum + jp) & 1);
              if (jp == 0 && n > 0) ptx::mbar_wait(&o_free[t], (n - 1) & 1);   // the previous item's O has been read
              ptx::tc_fence_after();
              const uint32_t vaddr = ptx::smem_u32(sV + st * C::KV_BYTES);
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
                for (int k = 0; k < AT_BK / 16; ++k) {
                  const uint64_t dv = ptx::make_desc_mnmajor_sw128(vaddr + sub * (AT_BK * 128) + k * 2048, 1024, 1024);
                  ptx::mma_f16_ts(tmem_s + C::O_COL + sub * 64, tmem_s + C::P_COL + k * 8, dv, idesc_pv,
                                  (jp | k) != 0 ? 1u : 0u);
                }
              }
              ptx::mma_commit(&pv_done[t]);
              stamp(W_TMA + t, cum + jp, 1);
            }
            ptx::mma_commit(&v_empty[st]);
          }
        }
        cum += nt;
        kb += ntl;
        ptx::mbar_arrive(&item_empty[slot]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax / output (thread == query row)
    const int t = warp >> 2;                       // query tile of this warpgroup
    const int r = threadIdx.x & (AT_BQ - 1);       // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + t * C::TILE_COLS;
    const uint32_t tmem_p = tmem_s + C::P_COL;
    const uint32_t tmem_o = tmem_s + C::O_COL;
    constexpr float RESCALE_LOG2 = 8.0f;           // lazy rescaling: O is rescaled only when a row max grows by more than 2^8
    const uint32_t a_s_full = ptx::smem_u32(&s_full[t]), a_s_empty = ptx::smem_u32(&s_empty[t]);
    const uint32_t a_p_full = ptx::smem_u32(&p_full[t]), a_pv_done = ptx::smem_u32(&pv_done[t]);
    unsigned long long scale2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(scale2) : "f"(p.scale_log2));
    uint32_t cum = 0;                              // key tiles this query tile has processed in earlier items
    bool s_ok = false;                             // the probe inside the last exponential stream saw the next S complete

    for (uint32_t n = 0;; ++n) {
      const int slot = n % NI;
      ptx::mbar_wait(&item_full[slot], (n / NI) & 1);
      const int4 it = ptx::ld_shared_v4(ptx::smem_u32(&items[slot]));
      if (it.x < 0) break;
      const int my_tiles = a4_tile_keys(p, it.x, t);
      const int qi = it.x + t * AT_BQ + r;           // row inside this item's query block
      int limit = p.Tk - 1;
      if (p.block_causal) limit = min(limit, (p.q_pos0 + qi) | 7);
      float m_ref = 0.f, l = 0.f;

      // Per key tile: S -> registers (S released at once: Q K_{j+1}^T runs under this tile's softmax), row max, 64
      // exponentials into registers, P -> tensor memory once P_{j-1} V_{j-1} has completed.  The two barrier probes
      // (pv_done for this tile, s_full for the next) are issued INSIDE the exponential stream — unconditionally, a branch
      // would cut the stream into separately scheduled blocks — so their shared-memory round trips (~180 clocks each
      // when taken in line) hide under the MUFU-bound work; the blocking waits remain as the fallback.
      uint32_t sa[32], sb[32], w[32];
      for (int j = 0; j < my_tiles; ++j) {
        const uint32_t cj = cum + j;
        if (lane == 0) stamp(warp, cj, 0);
        if (!s_ok) ptx::mbar_wait(a_s_full, cj & 1);
        ptx::tc_fence_after();
        if (lane == 0) stamp(warp, cj, 1);
        ptx::tmem_ld32(tmem_s + lane_off, sa);
        ptx::tmem_ld32(tmem_s + lane_off + 32, sb);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_s_empty);   // Q K_{j+1}^T may overwrite S now
        if (lane == 0) stamp(warp, cj, 2);

        const int lim = limit - j * AT_BK;  // columns c <= lim are visible
        if (lim < AT_BK - 1) {              // diagonal / last tile: mask (interior tiles skip this entirely)
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            if (cc > lim) sa[cc] = 0xff800000u;        // -inf
            if (cc + 32 > lim) sb[cc] = 0xff800000u;
          }
        }
        float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};  // 4 independent chains
#pragma unroll
        for (int cc = 0; cc < 32; cc += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(sa[cc + u]), __uint_as_float(sb[cc + u])));
        }
        const float m_tile = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;  // scale > 0
        if (j == 0) {
          m_ref = (m_tile == -CUDART_INF_F) ? 0.f : m_tile;
        } else {
          const bool need = m_tile > m_ref + RESCALE_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            // rare: a row maximum grew by more than 2^8.  O must be quiescent: P_{j-1} V_{j-1} has completed
            ptx::mbar_wait(a_pv_done, (cj - 1) & 1);
            ptx::tc_fence_after();
            const float alpha = need ? fast_exp2(m_ref - m_tile) : 1.0f;
            if (need) m_ref = m_tile;
            l *= alpha;
#pragma unroll 1
            for (int q4 = 0; q4 < HD / 16; ++q4) {
              uint32_t tt[16];
              ptx::tmem_ld16(tmem_o + lane_off + q4 * 16, tt);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int cc = 0; cc < 16; ++cc) tt[cc] = __float_as_uint(__uint_as_float(tt[cc]) * alpha);
              ptx::tmem_st16(tmem_o + lane_off + q4 * 16, tt);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
          }
        }
        if (lane == 0) stamp(warp, cj, 3);
        // p = 2^(s*scale - m_ref): all 64 exponentials into registers (packed halves)
        unsigned long long negm2;
        {
          const float neg_m = -m_ref;
          asm("mov.b64 %0, {%1, %1};" : "=l"(negm2) : "f"(neg_m));
        }
        __half2 acc[4];
        bool pv_ok = false;
#pragma unroll
        for (int e = 0; e < 32; ++e) {       // pair e = keys (2e, 2e+1)
          const uint32_t s0 = (e < 16) ? sa[2 * e] : sb[2 * e - 32];
          const uint32_t s1 = (e < 16) ? sa[2 * e + 1] : sb[2 * e - 31];
          unsigned long long xx;
          asm("{\n\t.reg .b64 a;\n\tmov.b64 a, {%1, %2};\n\tfma.rn.f32x2 %0, a, %3, %4;\n\t}"
              : "=l"(xx) : "r"(s0), "r"(s1), "l"(scale2), "l"(negm2));
          // EMU of every 8 pairs on the FMA pipe instead of the MUFU, spread evenly through the stream
          if (((e & 7) * EMU) / 8 != (((e & 7) + 1) * EMU) / 8) {
            w[e] = exp2_poly_f16x2(xx);
          } else {
            float x0, x1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
            w[e] = exp2_f16x2(x0, x1);
          }
          if (e < 4) acc[e] = *reinterpret_cast<const __half2*>(&w[e]);
          else acc[e & 3] = __hadd2(acc[e & 3], *reinterpret_cast<const __half2*>(&w[e]));
          if (PROBE && e == 16) pv_ok = ptx::mbar_test(a_pv_done, (cj - 1) & 1);   // fresh barrier: parity 1 is "complete"
          if (PROBE && e == 28) s_ok = ptx::mbar_test(a_s_full, (cj + 1) & 1);     // S of the next key tile (or item)
        }
        {
          const float2 f0 = __half22float2(acc[0]), f1 = __half22float2(acc[1]);
          const float2 f2 = __half22float2(acc[2]), f3 = __half22float2(acc[3]);
          l += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
        }
        if (lane == 0) stamp(warp, cj, 4);
        // the P columns are free once the previous P V of this query tile has completed (issued a whole softmax period ago)
        if (!pv_ok && cj > 0) ptx::mbar_wait(a_pv_done, (cj - 1) & 1);
        ptx::tc_fence_after();
        if (lane == 0) stamp(warp, cj, 5);
        ptx::tmem_st32(tmem_p + lane_off, w);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_p_full);
        if (lane == 0) stamp(warp, cj, 6);
      }
      if (my_tiles > 0) {
        ptx::mbar_wait(a_pv_done, (cum + my_tiles - 1) & 1);
        ptx::tc_fence_after();
        const float inv = 1.0f / l;
        __half* op = p.out + it.z * p.o_batch_pitch + static_cast<long long>(qi) * p.o_row_pitch + it.y * HD;
#pragma unroll
        for (int part = 0; part < HD / 32; ++part) {
          uint32_t tt[32];
          ptx::tmem_ld32(tmem_o + lane_off + part * 32, tt);
          ptx::tmem_ld_wait();
          if (qi < p.Tq) {
#pragma unroll
            for (int cc = 0; cc < 32; cc += 8) {
              uint4 q;
              q.x = pack_half2(__uint_as_float(tt[cc + 0]) * inv, __uint_as_float(tt[cc + 1]) * inv);
              q.y = pack_half2(__uint_as_float(tt[cc + 2]) * inv, __uint_as_float(tt[cc + 3]) * inv);
              q.z = pack_half2(__uint_as_float(tt[cc + 4]) * inv, __uint_as_float(tt[cc + 5]) * inv);
              q.w = pack_half2(__uint_as_float(tt[cc + 6]) * inv, __uint_as_float(tt[cc + 7]) * inv);
              *reinterpret_cast<uint4*>(op + part * 32 + cc) = q;
            }
          }
        }
        ptx::tc_fence_before();
        cum += my_tiles;
      }
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&o_free[t]);          // the next item's first P V may overwrite O
        ptx::mbar_arrive(&item_empty[slot]);
      }
    }
  }
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

std::once_flag g_attn_once;
int g_attn_status = FRT2_OK;


// Device counters of the item scheduler when the caller does not bring its own (AttnDesc::sched): a ring of pairs,
// one per launch in flight (each launch re-arms its pair when its last CTA finishes).
constexpr int A4_SCHED_RING = 256;
std::mutex g_sched_mu;
unsigned int* g_sched_ring[16] = {};
unsigned int g_sched_next[16] = {};

int a4_default_sched(unsigned int** out) {
  int dev = 0;
  FRT2_CUDA_OK(cudaGetDevice(&dev));
  FRT2_REQUIRE(dev >= 0 && dev < 16, FRT2_ERR_BAD_ARG, "attention_tc: device index out of range");
  std::lock_guard<std::mutex> lk(g_sched_mu);
  if (g_sched_ring[dev] == nullptr) {
    FRT2_CUDA_OK(cudaMalloc(&g_sched_ring[dev], A4_SCHED_RING * 2 * sizeof(unsigned int)));
    FRT2_CUDA_OK(cudaMemset(g_sched_ring[dev], 0, A4_SCHED_RING * 2 * sizeof(unsigned int)));
  }
  *out = g_sched_ring[dev] + 2 * (g_sched_next[dev]++ % A4_SCHED_RING);
  return FRT2_OK;
}

int g_attn_sms = 0;

uint32_t* g_a4_trace = nullptr;

template <int HD, int QT, bool PROBE, bool TRACE, int EMU>
int launch_t4(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, Attn4Params p, const AttnDesc& a,
              cudaStream_t stream) {
  using C = A4<HD, QT>;
  p.groups = (a.Tq + AT_BQ * QT - 1) / (AT_BQ * QT);
  p.total = p.groups * a.H * a.B;
  static const int tma_sleep = getenv("FRT2_A4_TMA_SLEEP") ? atoi(getenv("FRT2_A4_TMA_SLEEP")) : 100;
  static const int one_item = getenv("FRT2_A4_ONE_ITEM") ? atoi(getenv("FRT2_A4_ONE_ITEM")) : 0;
  p.tma_sleep_ns = tma_sleep;
  p.one_item_per_cta = one_item;
  const int grid = one_item ? p.total : std::min(p.total, g_attn_sms);
  p.trace = g_a4_trace;
  attention_t4_kernel<HD, QT, PROBE, TRACE, EMU><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace

int attention_tc_init() {
  std::call_once(g_attn_once, [] {
    g_attn_status = gemm_tc_init();
    if (g_attn_status != FRT2_OK) return;
    auto set = [](const void* fn, int bytes, const char* what) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e != cudaSuccess) {
        set_error(std::string("cudaFuncSetAttribute(") + what + "): " + cudaGetErrorString(e));
        g_attn_status = FRT2_ERR_CUDA;
      }
    };
    set(reinterpret_cast<const void*>(attention_t3_kernel), A3_SMEM_BYTES, "attention_t3_kernel");
    set(reinterpret_cast<const void*>(attention_t4_kernel<64, 3, true, false, 2>), A4<64, 3>::SMEM_BYTES, "attention_t4_kernel<64,3>");
    set(reinterpret_cast<const void*>(attention_t4_kernel<64, 3, true, false, 0>), A4<64, 3>::SMEM_BYTES, "attention_t4_kernel<64,3>");
    set(reinterpret_cast<const void*>(attention_t4_kernel<64, 3, false, false, 2>), A4<64, 3>::SMEM_BYTES, "attention_t4_kernel<64,3>");
    set(reinterpret_cast<const void*>(attention_t4_kernel<64, 3, true, true, 2>), A4<64, 3>::SMEM_BYTES, "attention_t4_kernel<64,3>");
    set(reinterpret_cast<const void*>(attention_t4_kernel<128, 2, true, false, 2>), A4<128, 2>::SMEM_BYTES, "attention_t4_kernel<128,2>");
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      set_error("attention_tc: cannot query the device");
      g_attn_status = FRT2_ERR_CUDA;
      return;
    }
    g_attn_sms = prop.multiProcessorCount;
  });
  return g_attn_status;
}

// debug: clock64 stamps of CTA 0 of the persistent kernel go to `dev_buf` (16 * 128 * 8 uint32) while it is set
void attention_tc_set_trace(void* dev_buf) { g_a4_trace = static_cast<uint32_t*>(dev_buf); }

int attention_tc(const AttnDesc& a, cudaStream_t stream) {
  FRT2_TRY(attention_tc_init());
  static const int ver = getenv("FRT2_ATTN_VER") ? atoi(getenv("FRT2_ATTN_VER")) : 4;   // 3: per-group CTAs (A/B)
  FRT2_REQUIRE(a.hd == 64 || (a.hd == 128 && ver != 3), FRT2_ERR_BAD_ARG, "attention_tc: head_dim must be 64 or 128");
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
    uint32_t box[3] = {64, AT_BQ, 1};
    FRT2_TRY(tma_encode_fp16(&tmQ, a.q, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {cols, static_cast<uint64_t>(a.Tk), static_cast<uint64_t>(a.B)};
    uint64_t strides[2] = {static_cast<uint64_t>(a.kv_row_pitch) * 2,
                           static_cast<uint64_t>(a.B > 1 ? a.kv_batch_pitch : a.kv_row_pitch * a.Tk) * 2};
    uint32_t box[3] = {64, AT_BK, 1};
    FRT2_TRY(tma_encode_fp16(&tmK, a.k, 3, dims, strides, box));
    FRT2_TRY(tma_encode_fp16(&tmV, a.v, 3, dims, strides, box));
  }
  if (ver == 3) {
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
  Attn4Params p;
  p.Tq = a.Tq;
  p.Tk = a.Tk;
  p.q_pos0 = a.q_pos0;
  p.block_causal = a.block_causal;
  p.H = a.H;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.out = a.out;
  p.o_row_pitch = a.o_row_pitch;
  p.o_batch_pitch = a.o_batch_pitch;
  p.sched = a.sched;
  if (p.sched == nullptr) FRT2_TRY(a4_default_sched(&p.sched));
  // A/B switches (read once): FRT2_A4_EMU=0 every exponential on the MUFU; FRT2_A4_PROBE=0 barrier probes in line
  static const int probe = getenv("FRT2_A4_PROBE") ? atoi(getenv("FRT2_A4_PROBE")) : 1;
  static const int emu = getenv("FRT2_A4_EMU") ? atoi(getenv("FRT2_A4_EMU")) : 2;
  if (a.hd == 128) return launch_t4<128, 2, true, false, 2>(tmQ, tmK, tmV, p, a, stream);
  if (g_a4_trace != nullptr) return launch_t4<64, 3, true, true, 2>(tmQ, tmK, tmV, p, a, stream);
  if (emu == 0) return launch_t4<64, 3, true, false, 0>(tmQ, tmK, tmV, p, a, stream);
  if (!probe) return launch_t4<64, 3, false, false, 2>(tmQ, tmK, tmV, p, a, stream);
  return launch_t4<64, 3, true, false, 2>(tmQ, tmK, tmV, p, a, stream);
}

}  // namespace frt2
