// Device-function bodies of the small bandwidth-bound kernels (LayerNorm rows, overlap-add sample) of kernels_misc.cu.
#pragma once
#include "common.cuh"

namespace frt2 {
namespace {

constexpr int LN_MAX_V4 = 8;  // register cache: up to 32 lanes * 8 float4 = 1024 channels; wider rows re-read

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two rows per warp: 2 x (C/128) independent 16-byte loads in flight per lane before the first reduction.
// row0 = first of the RPW rows this WARP normalises (the caller maps warps to rows)
template <int RPW>
__device__ __forceinline__ void layer_norm_body(const float* x, long long ldx, long long rows, int rows_per_batch, int C,
                                                const float* gamma, const float* beta, float eps, int apply_silu,
                                                __half* out16, long long ld16, long long out_batch_pitch,
                                                long long row0, float* mean_out = nullptr) {
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  const int C4 = C >> 2;
  float4 cache[RPW][LN_MAX_V4];
  float s[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    s[r] = 0.f;
    const long long row = min(row0 + r, rows - 1);
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int c = lane + 32 * i;
      if (c < C4) cache[r][i] = __ldcs(xr + c);
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long long row = min(row0 + r, rows - 1);
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int c = lane + 32 * i;
      if (c < C4) s[r] += (cache[r][i].x + cache[r][i].y) + (cache[r][i].z + cache[r][i].w);
    }
    for (int c = lane + 32 * LN_MAX_V4; c < C4; c += 32) {
      const float4 v = xr[c];
      s[r] += (v.x + v.y) + (v.z + v.w);
    }
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  // gamma / beta of the register-cached columns: in flight together with the row, not one round trip per use
  // (one row per warp only — the latency-bound streaming step; the two-row throughput kernel has no registers to spare
  // and hides the latency with occupancy)
  constexpr int NPRE = (RPW == 1) ? LN_MAX_V4 : 1;
  float4 gc[NPRE], bc[NPRE];
  if (RPW == 1) {
#pragma unroll
    for (int i = 0; i < NPRE; ++i) {
      const int c = lane + 32 * i;
      if (c < C4) {
        gc[i] = __ldg(g4 + c);
        bc[i] = __ldg(b4 + c);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long long row = row0 + r;
    const float4* xr = reinterpret_cast<const float4*>(x + min(row, rows - 1) * ldx);
    const float mean = warp_sum(s[r]) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int c = lane + 32 * i;
      if (c < C4) {
        const float a = cache[r][i].x - mean, b = cache[r][i].y - mean, cc = cache[r][i].z - mean,
                    d = cache[r][i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    for (int c = lane + 32 * LN_MAX_V4; c < C4; c += 32) {
      const float4 v = xr[c];
      const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
    if (row >= rows) continue;   // (warp-uniform) duplicate of the last row
    if (mean_out != nullptr && lane == 0) mean_out[row] = mean;
    const long long bidx = row / rows_per_batch;
    const long long t = row - bidx * rows_per_batch;
    uint2* orow = reinterpret_cast<uint2*>(out16 + bidx * out_batch_pitch + t * ld16);
    auto emit = [&](int c, const float4& v, const float4& g, const float4& bb) {
      float y0 = (v.x - mean) * rstd * g.x + bb.x;
      float y1 = (v.y - mean) * rstd * g.y + bb.y;
      float y2 = (v.z - mean) * rstd * g.z + bb.z;
      float y3 = (v.w - mean) * rstd * g.w + bb.w;
      if (apply_silu) { y0 = silu(y0); y1 = silu(y1); y2 = silu(y2); y3 = silu(y3); }
      uint2 h;
      h.x = pack_half2(y0, y1);
      h.y = pack_half2(y2, y3);
      orow[c] = h;
    };
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int c = lane + 32 * i;
      if (c < C4) {
        if (RPW == 1) emit(c, cache[r][i], gc[i % NPRE], bc[i % NPRE]);
        else emit(c, cache[r][i], __ldg(g4 + c), __ldg(b4 + c));
      }
    }
    for (int c = lane + 32 * LN_MAX_V4; c < C4; c += 32) emit(c, xr[c], __ldg(g4 + c), __ldg(b4 + c));
  }
}

// one output sample n of item b
__device__ __forceinline__ void overlap_add_sample(const OlaDesc& d, int ntail, int start, int n_out_full, int b, int n) {
  if (d.ctrl != nullptr) {  // streaming inside a captured graph: the item's first / last / active come from HBM
    const int* cb = d.ctrl + b * CTRL_INTS;
    const int first = (cb[CTRL_POS] == 0), last = cb[CTRL_LAST];
    const int pad = (d.n_fft - d.hop) / 2;
    const int n_grid = n_out_full;   // upper bound the grid was sized for
    ntail = first ? 0 : (d.n_fft / d.hop - 1);
    start = first ? pad : (d.n_fft - d.hop);
    n_out_full = (ntail + d.T - 1) * d.hop + d.n_fft - start - (last ? pad : (d.n_fft - d.hop));
    if (cb[CTRL_ACTIVE] == 0) n_out_full = 0;
    if (n >= n_out_full) {   // past this item's sample count: defined zeros up to the common upper bound
      if (n < n_grid) {
        const long long o = static_cast<long long>(b) * d.audio_pitch + n;
        if (d.pcm16 != nullptr) d.pcm16[o] = 0;
        else d.audio[o] = 0.f;
      }
      return;
    }
  }
  if (n >= n_out_full) return;
  int TF = ntail + d.T;
  int n_out = n_out_full;
  if (d.lengths != nullptr) {  // offline var-len: item b has lengths[b] tokens == len_mul*lengths[b] frames
    TF = min(TF, d.lengths[b] * d.len_mul);
    n_out = TF * d.hop;
  }
  const long long oidx = (d.out_off != nullptr ? d.out_off[b] : static_cast<long long>(b) * d.audio_pitch) + n;
  if (n >= n_out) {
    if (d.out_off != nullptr) return;   // scatter form: the neighbouring unit owns those samples
    if (d.pcm16 != nullptr) d.pcm16[oidx] = 0;
    else d.audio[oidx] = 0.f;
    return;
  }
  const int m = n + start;
  int t_hi = m / d.hop;
  if (t_hi > TF - 1) t_hi = TF - 1;
  int t_lo = (m - d.n_fft + d.hop) / d.hop;  // ceil((m - n_fft + 1) / hop) for m - n_fft + 1 > 0
  if (m - d.n_fft + 1 <= 0) t_lo = 0;
  float y = 0.f, env = 0.f;
  if (t_hi - t_lo < 4) {   // n_fft == 4*hop: at most 4 covering frames, all loads issued before the first use
    float fv[4], wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = min(t_lo + u, t_hi);
      const int off = m - t * d.hop;
      const float* fr = (t < ntail)
                            ? d.tail + (static_cast<long long>(b) * 3 + t) * d.n_fft
                            : d.frames + static_cast<long long>(b) * d.frames_batch_pitch +
                                  static_cast<long long>(t - ntail) * d.n_fft;
      wv[u] = __ldg(d.window + off);
      fv[u] = fr[off];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t_lo + u <= t_hi) {
        y += fv[u];
        env += wv[u] * wv[u];
      }
    }
  } else {
    for (int t = t_lo; t <= t_hi; ++t) {
      const int off = m - t * d.hop;
      const float* fr = (t < ntail)
                            ? d.tail + (static_cast<long long>(b) * 3 + t) * d.n_fft
                            : d.frames + static_cast<long long>(b) * d.frames_batch_pitch +
                                  static_cast<long long>(t - ntail) * d.n_fft;
      const float w = __ldg(d.window + off);
      y += fr[off];
      env += w * w;
    }
  }
  const float smp = y / env;
  if (d.pcm16 != nullptr) {
    // the reference's wire format: (audio * 32767).astype(np.int16) (enhanced_fireredtts2.py:603,655) — truncation
    // toward zero; out-of-range samples saturate here instead of wrapping
    d.pcm16[oidx] = static_cast<int16_t>(__float2int_rz(fminf(fmaxf(smp * 32767.0f, -32768.0f), 32767.0f)));
  } else {
    d.audio[oidx] = smp;
  }
}

}  // namespace
}  // namespace frt2
