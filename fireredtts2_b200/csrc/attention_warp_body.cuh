// K3' body — warp attention (8-query blocks, optional split-KV over the warps of a CTA) as a device function, instantiated by
// attention_warp_kernel (attention.cu).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace frt2 {
namespace {

// NSPLIT == 1: four independent 8-query blocks per CTA (one warp each).  NSPLIT > 1 (streaming step, few queries and
// a long KV state): the NSPLIT warps of a CTA share ONE query block, take interleaved 32-key chunks and merge their
// partial (max, sum, accumulator) through shared memory — split-KV without a second kernel.
// COHERENT: the K rows may have been written earlier in the SAME kernel (persistent step kernel) — no ld.global.nc.
// KS > 1 (with NSPLIT > 1): KS CTAs share one (item, head): CTA k of the group takes the 32-key chunks
// (k*NSPLIT + warp) + i*KS*NSPLIT, so up to KS*NSPLIT chunks run in one round and the step's attention time stays flat
// as the K/V state grows.  Only ceil(chunks / NSPLIT) CTAs (at most KS) have work; the others exit at once.  With more
// than one working CTA each leaves its merged (max, sum, accumulator) in a.part, and the LAST one to arrive (atomic
// counter per (item, head), reset by that CTA) merges them in CTA order — deterministic — and writes the rows.
template <int HD, int NSPLIT, bool COHERENT, int KS = 1>
__device__ __forceinline__ void attention_warp_body(AttnDesc a, long long vblock) {
  const int kcta = (KS > 1) ? static_cast<int>(vblock % KS) : 0;
  if (KS > 1) vblock /= KS;
  constexpr int DPL = HD / 32;  // output dims per lane
  constexpr int QW = (NSPLIT == 1) ? 4 : 1;   // query blocks per CTA
  __shared__ float sq[QW][8][HD];
  __shared__ float s_part[NSPLIT == 1 ? 1 : NSPLIT][8][HD + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = a.Tq >> 3;
  const int qslot = (NSPLIT == 1) ? warp : 0;
  const int split = (NSPLIT == 1) ? 0 : warp;
  const long long wid = (NSPLIT == 1) ? vblock * 4 + warp : vblock;
  if (wid >= static_cast<long long>(a.B) * a.H * nblk) return;
  const int qb = static_cast<int>(wid % nblk);
  const int h = static_cast<int>((wid / nblk) % a.H);
  const int b = static_cast<int>(wid / (static_cast<long long>(nblk) * a.H));
  const float scale_log2 = a.scale * 1.4426950408889634f;
  if (a.ctrl != nullptr) {  // streaming inside a captured graph: the item's position comes from its control block
    const int* cb = a.ctrl + b * CTRL_INTS;
    if (cb[CTRL_ACTIVE] == 0) return;   // idle pool slot (uniform per CTA: a CTA never spans two items' control blocks
                                        // in the NSPLIT > 1 form; in the NSPLIT == 1 form the exit is per warp)
    a.q_pos0 = cb[CTRL_POS];
    a.Tk = a.q_pos0 + a.Tq;
  }

  const __half* qp = a.q + b * a.q_batch_pitch + static_cast<long long>(qb * 8) * a.q_row_pitch + h * HD;
  if (NSPLIT == 1) {
    for (int e = lane; e < 8 * HD; e += 32) {
      const int r = e / HD, d = e - r * HD;
      sq[qslot][r][d] = __half2float(qp[r * a.q_row_pitch + d]) * scale_log2;
    }
    __syncwarp();
  } else {
    for (int e = threadIdx.x; e < 8 * HD; e += NSPLIT * 32) {
      const int r = e / HD, d = e - r * HD;
      sq[0][r][d] = __half2float(qp[r * a.q_row_pitch + d]) * scale_log2;
    }
    __syncthreads();
  }

  int kend = a.Tk - 1;
  if (a.block_causal) kend = min(kend, (a.q_pos0 + qb * 8) | 7);
  const __half* kp = a.k + b * a.kv_batch_pitch + h * HD;
  const __half* vp = a.v + b * a.kv_batch_pitch + h * HD;

  float m[8], l[8], acc[8][DPL];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    m[r] = -CUDART_INF_F;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[r][i] = 0.f;
  }
  int nact = 1;   // CTAs of this (item, head) that have work
  if (KS > 1) {
    nact = min(KS, (kend / 32 + NSPLIT) / NSPLIT);
    if (kcta >= nact) return;   // uniform per CTA
  }
  for (int j0 = (kcta * NSPLIT + split) * 32; j0 <= kend; j0 += KS * NSPLIT * 32) {
    const int j = j0 + lane;
    const bool valid = j <= kend;
    float s[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] = 0.f;
    const int nk = min(32, kend - j0 + 1);
    // the chunk's first 16 V rows are requested together with its K rows (they do not depend on the scores): the
    // chunk costs one memory round trip instead of one for K plus one per batch of V rows
    __half2 vpre[16];
    if (DPL == 2) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int jj = min(u, nk - 1);
        vpre[u] = *reinterpret_cast<const __half2*>(vp + static_cast<long long>(j0 + jj) * a.kv_row_pitch + lane * DPL);
      }
    }
    {
      const uint4* krow = reinterpret_cast<const uint4*>(kp + static_cast<long long>(valid ? j : kend) * a.kv_row_pitch);
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 kv = COHERENT ? __ldcg(krow + c) : __ldg(krow + c);
        const __half2* k2 = reinterpret_cast<const __half2*>(&kv);
        float kf[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(k2[i]);
          kf[2 * i] = f.x;
          kf[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 q0 = *reinterpret_cast<const float4*>(&sq[qslot][r][c * 8]);
          const float4 q1 = *reinterpret_cast<const float4*>(&sq[qslot][r][c * 8 + 4]);
          s[r] += q0.x * kf[0] + q0.y * kf[1] + q0.z * kf[2] + q0.w * kf[3] + q1.x * kf[4] + q1.y * kf[5] +
                  q1.z * kf[6] + q1.w * kf[7];
        }
      }
    }
    float p[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float sv = valid ? s[r] : -CUDART_INF_F;
      float mx = sv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m[r], mx);
      const float alpha = exp2f(m[r] - m_new);
      p[r] = valid ? exp2f(sv - m_new) : 0.f;
      float ps = p[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l[r] = l[r] * alpha + ps;
      m[r] = m_new;
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc[r][i] *= alpha;
    }
    if (DPL == 2) {
      __half2 vnext[16];
      if (nk > 16) {   // second half of the chunk: requested now, consumed after the first half's FMAs
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int jj = min(16 + u, nk - 1);
          vnext[u] = *reinterpret_cast<const __half2*>(vp + static_cast<long long>(j0 + jj) * a.kv_row_pitch + lane * DPL);
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (u < nk) {
          const float2 f = __half22float2(vpre[u]);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float pj = __shfl_sync(0xffffffffu, p[r], u);
            acc[r][0] = fmaf(pj, f.x, acc[r][0]);
            acc[r][DPL - 1] = fmaf(pj, f.y, acc[r][DPL - 1]);
          }
        }
      }
      if (nk > 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (16 + u < nk) {
            const float2 f = __half22float2(vnext[u]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float pj = __shfl_sync(0xffffffffu, p[r], 16 + u);
              acc[r][0] = fmaf(pj, f.x, acc[r][0]);
              acc[r][DPL - 1] = fmaf(pj, f.y, acc[r][DPL - 1]);
            }
          }
        }
      }
    } else {
      for (int jj = 0; jj < nk; ++jj) {
        const __half* vrow = vp + static_cast<long long>(j0 + jj) * a.kv_row_pitch + lane * DPL;
        float vf[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) vf[i] = __half2float(vrow[i]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float pj = __shfl_sync(0xffffffffu, p[r], jj);
#pragma unroll
          for (int i = 0; i < DPL; ++i) acc[r][i] = fmaf(pj, vf[i], acc[r][i]);
        }
      }
    }
  }
  __half* op = a.out + b * a.o_batch_pitch + static_cast<long long>(qb * 8) * a.o_row_pitch + h * HD + lane * DPL;
  if (NSPLIT == 1) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float inv = 1.0f / l[r];
#pragma unroll
      for (int i = 0; i < DPL; ++i) op[r * a.o_row_pitch + i] = to_half_sat(acc[r][i] * inv);
    }
  } else {
    // merge the NSPLIT partial softmax states: out = sum_w acc_w 2^(m_w - M) / sum_w l_w 2^(m_w - M)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < DPL; ++i) s_part[split][r][lane * DPL + i] = acc[r][i];
      if (lane == 0) {
        s_part[split][r][HD] = m[r];
        s_part[split][r][HD + 1] = l[r];
      }
    }
    __syncthreads();
    __shared__ int s_last;
    float M = -CUDART_INF_F, L = 0.f, o[DPL];
    if (warp < 8) {
      const int r = warp;
      for (int w = 0; w < NSPLIT; ++w) M = fmaxf(M, s_part[w][r][HD]);
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[i] = 0.f;
      for (int w = 0; w < NSPLIT; ++w) {
        const float sc = exp2f(s_part[w][r][HD] - M);   // 0 for warps that saw no key (m = -inf)
        L += s_part[w][r][HD + 1] * sc;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[i] += s_part[w][r][lane * DPL + i] * sc;
      }
      if (KS == 1 || nact == 1) {
        const float inv = 1.0f / L;
#pragma unroll
        for (int i = 0; i < DPL; ++i) op[r * a.o_row_pitch + i] = to_half_sat(o[i] * inv);
      }
    }
    if (KS > 1 && nact > 1) {
      // ---- cross-CTA merge: this CTA's state -> a.part[(item, head)][kcta][row][HD + 2]
      const long long bh = static_cast<long long>(b) * a.H + h;
      float* mine = a.part + ((bh * KS + kcta) * 8) * (HD + 2);
      if (warp < 8) {
        const int r = warp;
#pragma unroll
        for (int i = 0; i < DPL; ++i) mine[r * (HD + 2) + lane * DPL + i] = o[i];
        if (lane == 0) {
          mine[r * (HD + 2) + HD] = M;
          mine[r * (HD + 2) + HD + 1] = L;
        }
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(a.part_count + bh, 1) == nact - 1) ? 1 : 0;
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (warp < 8) {
          const int r = warp;
          const float* base = a.part + (bh * KS * 8 + r) * (HD + 2);
          float M2 = -CUDART_INF_F;
          for (int k = 0; k < nact; ++k) M2 = fmaxf(M2, __ldcg(base + static_cast<long long>(k) * 8 * (HD + 2) + HD));
          float L2 = 0.f, o2[DPL];
#pragma unroll
          for (int i = 0; i < DPL; ++i) o2[i] = 0.f;
          for (int k = 0; k < nact; ++k) {
            const float* pk = base + static_cast<long long>(k) * 8 * (HD + 2);
            const float sc = exp2f(__ldcg(pk + HD) - M2);
            L2 += __ldcg(pk + HD + 1) * sc;
#pragma unroll
            for (int i = 0; i < DPL; ++i) o2[i] += __ldcg(pk + lane * DPL + i) * sc;
          }
          const float inv = 1.0f / L2;
#pragma unroll
          for (int i = 0; i < DPL; ++i) op[r * a.o_row_pitch + i] = to_half_sat(o2[i] * inv);
        }
        if (threadIdx.x == 0) a.part_count[bh] = 0;   // ready for the next step
      }
    }
  }
}

}  // namespace
}  // namespace frt2
