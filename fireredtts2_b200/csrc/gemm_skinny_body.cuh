// K2s body — the <= 16-row weight-streaming GEMM of the token step as a device function, instantiated by
// gemm_skinny_kernel (gemm_skinny.cu).  See gemm_skinny.cu for the design.
#pragma once
#include "common.cuh"

namespace frt2 {
namespace {

constexpr int SK_WARPS = 8;        // the 8 warps of a CTA split K
constexpr int SK_COLS = 8;         // output columns per n-tile == the n of mma.m16n8k16 (a CTA owns NT of them)
constexpr int SK_MAXROWS = 16;     // == the m of mma.m16n8k16
constexpr int SK_KCHUNK = 4096;    // activation chunk held in smem
constexpr int SK_PAD = 32;         // halves of padding per activation row: rows g and g+1 of a quarter-warp's
                                   // 16-byte loads land on disjoint banks
constexpr int SK_NB = 16;          // k32 blocks (16-byte weight vectors per lane) in flight per warp

__host__ __device__ inline int sk_pitch(int kchunk) { return ((kchunk + 31) & ~31) + SK_PAD; }
// ln_c = channels of a fused LayerNorm (its gamma / beta are staged in shared memory), 0 without one
// nt = n-tiles (8 columns each) per CTA
inline size_t sk_smem_bytes(int mr, int kchunk, int ln_c = 0, int nt = 1) {
  return static_cast<size_t>(mr) * sk_pitch(kchunk) * 2 + static_cast<size_t>(SK_WARPS + 1) * mr * SK_COLS * nt * 4 +
         static_cast<size_t>(ln_c) * 8;
}

// 16-byte global -> shared copy that does not pass through registers (LDGSTS); src_ok = false writes zeros
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool src_ok) {
  const unsigned int d = static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst));
  const int bytes = src_ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// D(16x8) += A(16x16, row) * B(16x8, col); f16 operands, f32 accumulate (legacy tensor path: the step is bound by
// streaming weights, the MMA only removes the ~100 CUDA-core instructions per 16 bytes of weights of an FMA loop)
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Work split: a CTA owns 8 output columns; warp w owns the k32 blocks w, w+8, w+16, ... of the reduction and lane
// (g = lane/4, t = lane%4) streams 16 bytes W[n0+g][k0+8t .. k0+8t+7] per block — its B fragments for TWO MMAs (halves
// 0-3, then 4-7).  The tensor core's k slots are a permutation of the true k inside the block; the activation fragment
// is read with the SAME permutation (one 16-byte LDS of X[g][k0+8t .. +7]), and a dot product does not care.  Rows
// 8..15 of the m16 tile are the second item of a 2-stream pool (MR = 16) or zero.
// `vblock` = which group of 8 output columns, `tid` = thread index inside the 256-thread group that runs the tile,
// `sync` = barrier over exactly those 256 threads (__syncthreads in the stand-alone kernel, a named barrier when two
// groups of the persistent step kernel work on different tiles), PDL = programmatic-dependent-launch hooks.
// NT = 8-column n-tiles per CTA: 2 for the wide layers (N >= 2048: QKV, fc1) so that their CTAs fit ONE wave of the
// 148 SMs and the activation rows are staged / normalised half as often; a column's arithmetic does not depend on NT.
template <int MR, bool PDL, int NT = 1, typename Sync>
__device__ __forceinline__ void gemm_skinny_body(const GemmDesc& g, int mtot, int vblock, int tid, uint8_t* sk_smem,
                                                 Sync sync) {
  constexpr int COLS = SK_COLS * NT;
  constexpr int NB = NT == 1 ? SK_NB : (NT == 2 ? SK_NB / 2 : SK_NB / 4);   // weight vectors in flight per lane and n-tile
  __half* sA = reinterpret_cast<__half*>(sk_smem);                     // [MR][pitch]
  const int Ktot = g.ntaps * g.Kc;
  const int kchunk = min(Ktot, SK_KCHUNK);
  const int pitch = sk_pitch(kchunk);
  float* sRed = reinterpret_cast<float*>(sk_smem + static_cast<size_t>(MR) * pitch * 2);   // [SK_WARPS][MR][COLS]
  float* sOut = sRed + SK_WARPS * MR * COLS;                                             // [MR][COLS]
  float* sGam = sOut + MR * COLS;                                                        // [Kc] (fused LayerNorm)
  float* sBet = sGam + g.Kc;                                                                // [Kc]
  const int warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  bool col_ok[NT];
  const uint4* wrow[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int n = vblock * COLS + t * SK_COLS + gq;
    col_ok[t] = n < g.N;
    wrow[t] = reinterpret_cast<const uint4*>(g.W + static_cast<long long>(col_ok[t] ? n : 0) * Ktot);
  }

  uint4 wv[NT][NB];
  // weight vectors of this warp's blocks ib0 .. ib0+NB-1 of the chunk starting at kc0 (kc halves long)
  auto load_weights = [&](int kc0, int kc, int ib0) {
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int k = (warp + SK_WARPS * (ib0 + u)) * 32 + tq * 8;
#pragma unroll
      for (int t = 0; t < NT; ++t)
        wv[t][u] = (col_ok[t] && k < kc) ? __ldg(wrow[t] + ((kc0 + k) >> 3)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  // ---- predecessor-independent prologue: the first weight vectors are in flight while the previous kernel of the
  //      step is still running (PDL) and while the activations are staged below
  load_weights(0, kchunk, 0);
  float bias_c = 0.f, bias_p = 0.f;
  if (g.bias != nullptr && tid < MR * COLS) {
    const int c = tid % COLS;
    const int nn = vblock * COLS + c;
    if (nn < g.N) bias_c = __ldg(g.bias + nn);
    if ((nn ^ 1) < g.N) bias_p = __ldg(g.bias + (nn ^ 1));
  }
  if (g.ln_gamma != nullptr) {   // gamma / beta -> shared memory, asynchronously: no round trip inside the LayerNorm
    for (int c = tid * 4; c < g.Kc; c += SK_WARPS * 32 * 4) {
      cp_async16(sGam + c, g.ln_gamma + c, true);
      cp_async16(sBet + c, g.ln_beta != nullptr ? g.ln_beta + c : g.ln_gamma, g.ln_beta != nullptr);   // RMSNorm: zeros
    }
  }
  if (PDL) {
    pdl_wait();      // everything below may read what the previous kernel wrote
    pdl_trigger();   // the next kernel may start its own weight prefetch now
  }
  // the epilogue's run-time row offsets and residual values are known now: fetch them under the main loop
  int ep_roff = 0, ep_roff_b = 0;
  float ep_resid = 0.f;
  if (tid < MR * COLS) {
    const int m = tid / COLS, c = tid - m * COLS;
    const int nn = vblock * COLS + c;
    if (m < mtot && nn < g.N) {
      const int b = m / g.rows_out, r = m - b * g.rows_out;
      if (g.out_row_off != nullptr) ep_roff = __ldg(g.out_row_off + b * g.row_off_stride);
      if (g.split_col > 0 && nn >= g.split_col) {
        if (g.row_off_b != nullptr) ep_roff_b = __ldg(g.row_off_b + b * g.row_off_stride);
      } else if (g.resid != nullptr && g.out_row_off == nullptr) {
        ep_resid = g.resid[static_cast<long long>(b) * g.pitch32 + static_cast<long long>(r) * g.ld32 + nn];
      }
    }
  }

  float acc[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  for (int kc0 = 0; kc0 < Ktot; kc0 += kchunk) {
    const int kc = min(kchunk, Ktot - kc0);
    const int kc8 = kc >> 3;
    const int kc_r8 = ((kc + 31) & ~31) >> 3;    // staged width incl. the zero tail up to a whole k32 block
    if (g.ln_gamma != nullptr) {
      // ---- fused LayerNorm(+SiLU): one warp per row, fp32 statistics, result straight into the fp16 A tile
      //      (reference nn.LayerNorm eps 1e-5 / 1e-6: whisper.py:134,140, decoder.py:246)
      const int C4 = g.Kc >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(sGam);
      const float4* b4 = reinterpret_cast<const float4*>(sBet);
      cp_async_wait_all();
      sync();
      for (int m = warp; m < MR; m += SK_WARPS) {
        __half* arow = sA + static_cast<size_t>(m) * pitch;
        if (m >= mtot) {
          for (int c = lane * 8; c < kc_r8 * 8; c += 256) *reinterpret_cast<uint4*>(arow + c) = make_uint4(0u, 0u, 0u, 0u);
          continue;
        }
        for (int c = kc + lane * 8; c < kc_r8 * 8; c += 256) *reinterpret_cast<uint4*>(arow + c) = make_uint4(0u, 0u, 0u, 0u);
        const float4* xr = reinterpret_cast<const float4*>(g.ln_x + static_cast<long long>(m) * g.ln_ldx);
        auto emit = [&](int c, const float4& v, float mean, float rstd) {
          const float4 gg = g4[c], bb = b4[c];
          float y0 = (v.x - mean) * rstd * gg.x + bb.x, y1 = (v.y - mean) * rstd * gg.y + bb.y;
          float y2 = (v.z - mean) * rstd * gg.z + bb.z, y3 = (v.w - mean) * rstd * gg.w + bb.w;
          if (g.ln_silu) { y0 = silu(y0); y1 = silu(y1); y2 = silu(y2); y3 = silu(y3); }
          uint2 h;
          h.x = pack_half2(y0, y1);
          h.y = pack_half2(y2, y3);
          *reinterpret_cast<uint2*>(arow + 4 * c) = h;
        };
        if (C4 <= 256) {
          // the row lives in registers: ONE round trip to L2 instead of three
          float4 xv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            xv[i] = (c < C4) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (lane + 32 * i < C4) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = g.ln_rms ? 0.f : s / static_cast<float>(g.Kc);
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (lane + 32 * i < C4) {
              const float a = xv[i].x - mean, b = xv[i].y - mean, cc = xv[i].z - mean, d = xv[i].w - mean;
              q += (a * a + b * b) + (cc * cc + d * d);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          const float rstd = rsqrtf(q / static_cast<float>(g.Kc) + g.ln_eps);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (lane + 32 * i < C4) emit(lane + 32 * i, xv[i], mean, rstd);
        } else {
          float s = 0.f;
          for (int c = lane; c < C4; c += 32) {
            const float4 v = xr[c];
            s += (v.x + v.y) + (v.z + v.w);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = g.ln_rms ? 0.f : s / static_cast<float>(g.Kc);
          float q = 0.f;
          for (int c = lane; c < C4; c += 32) {
            const float4 v = xr[c];
            const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          const float rstd = rsqrtf(q / static_cast<float>(g.Kc) + g.ln_eps);
          for (int c = lane; c < C4; c += 32) emit(c, xr[c], mean, rstd);
        }
      }
    } else {
      // ---- activation chunk -> smem (causal taps gathered here: row m, tap j reads input row r + j + row_shift)
      //      all copies of a thread are in flight at once (LDGSTS): one round trip for the whole tile, not one per row
      for (int e = tid; e < MR * kc_r8; e += SK_WARPS * 32) {
        const int m = e / kc_r8, k8 = e - m * kc_r8;
        const __half* src_p = g.A;
        bool ok = false;
        if (m < mtot && k8 < kc8) {
          const int b = m / g.rows_out, r = m - b * g.rows_out;
          const int kk = kc0 + k8 * 8;
          const int tap = kk / g.Kc, c = kk - tap * g.Kc;
          const int src = r + tap + g.row_shift;
          if (src >= 0 && src < g.rows_a) {
            ok = true;
            src_p = g.A + static_cast<long long>(b) * g.a_batch_pitch + static_cast<long long>(src) * g.a_row_pitch + c;
          }
        }
        cp_async16(sA + static_cast<size_t>(m) * pitch + k8 * 8, src_p, ok);
      }
      cp_async_wait_all();
    }
    sync();
    {
      const int nblk = (kc + 31) >> 5;
      const int niter = nblk > warp ? (nblk - warp + SK_WARPS - 1) / SK_WARPS : 0;   // blocks of this warp
      const __half* abase = sA + static_cast<size_t>(gq) * pitch + tq * 8;
      for (int ib0 = 0; ib0 < niter; ib0 += NB) {
        if (kc0 != 0 || ib0 != 0) load_weights(kc0, kc, ib0);
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          if (ib0 + u < niter) {                                   // warp-uniform
            const int j = warp + SK_WARPS * (ib0 + u);
            const uint4 x = *reinterpret_cast<const uint4*>(abase + j * 32);
            uint4 y = make_uint4(0u, 0u, 0u, 0u);
            if (MR == 16) y = *reinterpret_cast<const uint4*>(abase + static_cast<size_t>(8) * pitch + j * 32);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              mma16816(acc[t], x.x, y.x, x.y, y.y, wv[t][u].x, wv[t][u].y);
              mma16816(acc[t], x.z, y.z, x.w, y.w, wv[t][u].z, wv[t][u].w);
            }
          }
        }
      }
    }
    sync();
  }
  // ---- split-K reduction over the 8 warps (fixed order: deterministic).  acc: rows g (c0,c1) and g+8 (c2,c3),
  //      columns 2t, 2t+1
  {
    float* r = sRed + static_cast<size_t>(warp) * MR * COLS;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      r[gq * COLS + t * SK_COLS + 2 * tq] = acc[t][0];
      r[gq * COLS + t * SK_COLS + 2 * tq + 1] = acc[t][1];
      if (MR == 16) {
        r[(gq + 8) * COLS + t * SK_COLS + 2 * tq] = acc[t][2];
        r[(gq + 8) * COLS + t * SK_COLS + 2 * tq + 1] = acc[t][3];
      }
    }
  }
  sync();
  if (tid < MR * COLS) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += sRed[w * MR * COLS + tid];
    sOut[tid] = v;
  }
  sync();
  // ---- epilogue: thread t -> (row m, column c); polar pairs read the neighbouring column from smem
  const int t = tid;
  if (t < MR * COLS) {
    const int m = t / COLS, c = t - m * COLS;
    const int nn = vblock * COLS + c;
    if (m < mtot && nn < g.N) {
      auto pre = [&](int cc) {   // cc == c (own column) or c ^ 1 (polar partner): biases were prefetched
        return fmaf(sOut[m * COLS + cc], g.alpha, cc == c ? bias_c : bias_p);
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
      } else if (g.act == ACT_SWIGLU) {   // (gate, up) pairs -> silu(gate) * up at column nn / 2 (exact expf: a handful per CTA)
        if ((c & 1) == 0 && (nn ^ 1) < g.N) {
          const float up = pre(c ^ 1);
          g.out16[static_cast<long long>(m / g.rows_out) * g.pitch16 +
                  static_cast<long long>(m % g.rows_out) * g.ld16 + (nn >> 1)] = to_half_sat(v / (1.0f + expf(-v)) * up);
        }
        return;
      }
      const int b = m / g.rows_out, r = m - b * g.rows_out;
      const int roff = ep_roff;
      const long long o32 = static_cast<long long>(b) * g.pitch32 + static_cast<long long>(r + roff) * g.ld32 + nn;
      const long long o16 = static_cast<long long>(b) * g.pitch16 + static_cast<long long>(r + roff) * g.ld16 + nn;
      if (g.split_col > 0 && nn >= g.split_col) {
        const int roff_b = ep_roff_b;
        g.out16_b[static_cast<long long>(b) * g.pitch16_b + static_cast<long long>(r + roff_b) * g.ld16_b +
                  (nn - g.split_col)] = to_half_sat(v);
      } else {
        if (g.resid != nullptr) v += (g.out_row_off == nullptr) ? ep_resid : g.resid[o32];
        if (g.out32 != nullptr) g.out32[o32] = v;
        if (g.out16 != nullptr) g.out16[o16] = to_half_sat(v);
      }
    }
  }
}

}  // namespace
}  // namespace frt2
