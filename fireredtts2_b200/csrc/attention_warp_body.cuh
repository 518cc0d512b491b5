// K3' body — warp attention (8-query blocks, optional split-KV over the warps of a CTA) as a device function, shared by
// attention_warp_kernel (attention.cu) and the persistent per-token step kernel (stream_mega.cu).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace frt2 {
namespace {

// NSPLIT == 1: four independent 8-query blocks per CTA (one warp each).  NSPLIT > 1 (streaming step, few queries and
// a long KV state): the NSPLIT warps of a CTA share ONE query block, take interleaved 32-key chunks and merge their
// partial (max, sum, accumulator) through shared memory — split-KV without a second kernel.
// COHERENT: the K rows may have been written earlier in the SAME kernel (persistent step kernel) — no ld.global.nc.
template <int HD, int NSPLIT, bool COHERENT>
__device__ __forceinline__ void attention_warp_body(AttnDesc a, long long vblock) {
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
  for (int j0 = split * 32; j0 <= kend; j0 += NSPLIT * 32) {
    const int j = j0 + lane;
    const bool valid = j <= kend;
    float s[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] = 0.f;
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
    const int nk = min(32, kend - j0 + 1);
    if (DPL == 2) {
      // eight V rows in flight per round trip (a load-use chain per key would cost one L2 latency per key)
      for (int jb = 0; jb < nk; jb += 8) {
        __half2 vh[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int jj = min(jb + u, nk - 1);
          vh[u] = *reinterpret_cast<const __half2*>(vp + static_cast<long long>(j0 + jj) * a.kv_row_pitch + lane * DPL);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (jb + u < nk) {
            const float2 f = __half22float2(vh[u]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float pj = __shfl_sync(0xffffffffu, p[r], jb + u);
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
    if (warp < 8) {
      const int r = warp;
      float M = -CUDART_INF_F;
      for (int w = 0; w < NSPLIT; ++w) M = fmaxf(M, s_part[w][r][HD]);
      float L = 0.f, o[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[i] = 0.f;
      for (int w = 0; w < NSPLIT; ++w) {
        const float sc = exp2f(s_part[w][r][HD] - M);   // 0 for warps that saw no key (m = -inf)
        L += s_part[w][r][HD + 1] * sc;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[i] += s_part[w][r][lane * DPL + i] * sc;
      }
      const float inv = 1.0f / L;
#pragma unroll
      for (int i = 0; i < DPL; ++i) op[r * a.o_row_pitch + i] = to_half_sat(o[i] * inv);
    }
  }
}

}  // namespace
}  // namespace frt2
