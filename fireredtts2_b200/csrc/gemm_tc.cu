// K2 — tcgen05 / TMEM / TMA GEMM with a multi-tap causal A operand and fused epilogues.
//
//   C[b, m, n] = act(alpha * sum_{tap, c} A[b, m + tap + row_shift, c] * W[n, tap*Kc + c] + bias[n]) (+ resid)
//
// One kernel covers every dense contraction of the decode path: the Linear layers (ntaps = 1), the causal
// Conv1d k=3/7 and both ConvTranspose1d of the upsampler (implicit GEMM: each tap is the same TMA box shifted
// by one time step on a (C, T, B) tensor map; rows before t = 0 are out of bounds and arrive zero-filled,
// which IS the causal left padding — reference decoder.py:88-91), the RVQ output projection, the iSTFT head
// and the windowed inverse DFT.
//
// Structure (persistent, one CTA per SM, 384 threads):
//   warp 0     TMA producer   : cp.async.bulk.tensor (SWIZZLE_128B) A box 64x128, W box 64xBN per k-block
//   warp 1     MMA issuer     : one elected thread, tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16
//   warp 2     TMEM allocator : 2*BN columns = two fp32 accumulators (epilogue of tile i overlaps MMA of i+1)
//   warps 4-11 epilogue       : two warps per TMEM lane quarter, each owning half of the tile's columns.
//                               tcgen05.ld 32x32b.x32 (thread == output row) -> bias / GELU / polar in registers
//                               -> 128-byte swizzled rows in a per-warp shared-memory staging buffer -> one
//                               cp.async.bulk.tensor store per 32x(128 B) box (full-line, coalesced writes; rows and
//                               columns past the tensor edge are clipped by the TMA unit).  The fp32 residual is
//                               fetched with coalesced 16-byte loads one chunk ahead, turned to row order through the
//                               same staging buffer and added before the store.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace frt2 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int STAGING_BYTES = 32 * 128;  // per epilogue warp: 32 rows x 128 B

enum StoreMode : int { STORE_DIRECT = 0, STORE_TMA16 = 1, STORE_TMA32 = 2 };

struct GemmKParams {
  int rows_out;
  long long pitch32;
  long long pitch16;
  int N;
  int kblocks_per_tap;
  int num_kblocks;
  int row_shift;
  int tiles_m;
  int tiles_n;
  int num_tiles;
  float alpha;
  const float* bias;
  int act;
  const float* resid;
  float* out32;
  long long ld32;
  __half* out16;
  long long ld16;
  int vec32;  // 16-byte vector access legal on out32 / resid rows (direct mode)
  int vec16;  // ... on out16 rows
  int store_mode;
  const int* row_off_ptr;
  int row_off_stride;
  // packed-item tiles (short per-item row counts, e.g. the 8 frames per token of a batch of streams): one 128-row
  // tile holds pack_items items x pack_rpb rows, loaded by ONE 3-D TMA box {64, pack_rpb, pack_items}; 0 = off
  int pack_rpb;
  int pack_items;
  int batches;
  int item_mul;        // first item of tile group g is g * item_mul (pack_items when packed, else 1)
  int w_batch_k;       // k offset of batch item b in W (split of the reduction over the batch dimension), else 0
  int stage_tx_bytes;  // bytes one pipeline stage receives (A box + W box)
  // folded LayerNorm (see GemmDesc): producer side ...
  __half* x16_out;
  long long ld_x16;
  const float* x16_shift;   // per-row offset subtracted before the fp16 rounding of the copy, or null
  // ... and consumer side
  const float2* stats_in;   // (mean, rstd) per A row
  const float* colsum;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + EPI_WARPS * STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2): two fp32 operations per issued instruction.  The epilogue of
// the GELU GEMM is issue-bound (27 instructions per element measured), so halving the FMA-pipe instruction count of
// the per-element math is what buys tensor-pipe time back.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// exact-erf GELU on a pair of values, erf by Abramowitz-Stegun 7.1.28:
//   erf(z) = 1 - (1 + a1 z + .. + a6 z^6)^-16,  |err| <= 3e-7  (far below the fp16 rounding of the stored activation)
// 7 FFMA2 + 5 FMUL2 + 2 MUFU.RCP + sign handling per PAIR instead of erff's branchy ~25 instructions per value.
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const f32x2 x = pk2(x0, x1);
  const f32x2 z = pk2(fabsf(x0) * 0.70710678118654752440f, fabsf(x1) * 0.70710678118654752440f);
  f32x2 p = fma2(z, pk2(0.0000430638f, 0.0000430638f), pk2(0.0002765672f, 0.0002765672f));
  p = fma2(p, z, pk2(0.0001520143f, 0.0001520143f));
  p = fma2(p, z, pk2(0.0092705272f, 0.0092705272f));
  p = fma2(p, z, pk2(0.0422820123f, 0.0422820123f));
  p = fma2(p, z, pk2(0.0705230784f, 0.0705230784f));
  p = fma2(p, z, pk2(1.0f, 1.0f));
  float p0, p1;
  unpk2(p, p0, p1);
  f32x2 r = pk2(rcp_approx(p0), rcp_approx(p1));
  r = mul2(r, r); r = mul2(r, r); r = mul2(r, r); r = mul2(r, r);          // p^-16
  float r0, r1;
  unpk2(r, r0, r1);
  // gelu = 0.5 x (1 + sign(x)(1 - r)) = 0.5 x + 0.5 |x| (1 - r)
  const f32x2 one_m_r = pk2(1.0f - r0, 1.0f - r1);
  const f32x2 hx = mul2(x, pk2(0.5f, 0.5f));
  const f32x2 habs = pk2(fabsf(x0) * 0.5f, fabsf(x1) * 0.5f);
  unpk2(fma2(habs, one_m_r, hx), x0, x1);
}

// alpha, bias and activation on one 32-column chunk held by a thread (one output row)
__device__ __forceinline__ void apply_chunk(const GemmKParams& p, const uint32_t (&r)[32], int n0, float (&v)[32],
                                            float ln_mean = 0.f, float ln_rstd = 1.f) {
  const bool full_chunk = n0 + 32 <= p.N;
  if (p.colsum != nullptr) {
    // folded LayerNorm: v = rstd * (acc*alpha - mean*colsum[n]) + bias[n] = acc*k1 + (k2*colsum[n] + bias[n]) with the
    // per-row k1 = alpha*rstd, k2 = -rstd*mean: two packed FFMA2 per PAIR of columns, i.e. the instruction count of
    // the plain acc*alpha + bias epilogue
    const float k1 = p.alpha * ln_rstd, k2 = -ln_rstd * ln_mean;
    if (full_chunk) {
      const f32x2 k1p = pk2(k1, k1), k2p = pk2(k2, k2);
      const float4* c4 = reinterpret_cast<const float4*>(p.colsum + n0);
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(c4 + j), b = __ldg(b4 + j);
        const f32x2 t0 = fma2(k2p, pk2(c.x, c.y), pk2(b.x, b.y));
        const f32x2 t1 = fma2(k2p, pk2(c.z, c.w), pk2(b.z, b.w));
        unpk2(fma2(pk2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), k1p, t0), v[4 * j + 0], v[4 * j + 1]);
        unpk2(fma2(pk2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), k1p, t1), v[4 * j + 2], v[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool ok = n0 + j < p.N;
        v[j] = fmaf(__uint_as_float(r[j]), k1,
                    fmaf(k2, ok ? __ldg(p.colsum + n0 + j) : 0.f, ok ? __ldg(p.bias + n0 + j) : 0.f));
      }
    }
  } else if (p.bias != nullptr) {
    if (full_chunk && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // one FFMA per element: acc * alpha + bias
        const float4 b = __ldg(b4 + j);
        v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), p.alpha, b.x);
        v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), p.alpha, b.y);
        v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), p.alpha, b.z);
        v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), p.alpha, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = fmaf(__uint_as_float(r[j]), p.alpha, (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  }
  if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_fast2(v[j], v[j + 1]);
  } else if (p.act == ACT_POLAR) {
    // columns come in (log-magnitude, phase) pairs: reference decoder.py:505-518
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float mag = fminf(expf(v[j]), 100.0f);
      float s, c;
      sincosf(v[j + 1], &s, &c);
      v[j] = mag * c;
      v[j + 1] = mag * s;
    }
  }
}

// direct (non-TMA) store of one chunk, used when the output pitch is not 16-byte aligned or both an fp32 and
// an fp16 copy are requested
__device__ __forceinline__ void store_chunk_direct(const GemmKParams& p, float (&v)[32], long long off32,
                                                   long long off16, int n0) {
  const bool full_chunk = n0 + 32 <= p.N;
  if (p.resid != nullptr) {
    const float* rp = p.resid + off32 + n0;
    if (full_chunk && p.vec32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(rp + 4 * j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) v[j] += rp[j];
    }
  }
  if (p.out32 != nullptr) {
    float* op = p.out32 + off32 + n0;
    if (full_chunk && p.vec32) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) op[j] = v[j];
    }
  }
  if (p.out16 != nullptr) {
    __half* op = p.out16 + off16 + n0;
    if (full_chunk && p.vec16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q;
        q.x = pack_half2(v[8 * j + 0], v[8 * j + 1]);
        q.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
        q.z = pack_half2(v[8 * j + 4], v[8 * j + 5]);
        q.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(op + 8 * j) = q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) op[j] = to_half_sat(v[j]);
    }
  }
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Epilogue of one 128 x BN accumulator tile held in this CTA's TMEM (shared by the 1-CTA and the 2-CTA kernels).
// e = epilogue warp index 0..7, row_slab0 = first output row of this CTA's 128-row slab, tmem_acc = TMEM address of
// the accumulator stage (tmem_stage).  Ends with tcgen05.fence::before_thread_sync + __syncwarp; the caller signals "TMEM free".
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, const CUtensorMap* tmC, uint8_t* sStage, int e,
                                              int lane, uint32_t tmem_stage, int n_idx, int b, int row_slab0,
                                              uint64_t* tfull, uint32_t aphase) {
    const int wq = e & 3;        // TMEM lane quarter == warp_id % 4
    const int hsel = e >> 2;     // which half of the tile's columns this warp owns
    constexpr int CHUNKS_PER_WARP = BN / 64;  // 32-column chunks per warp
    uint8_t* stg = sStage + e * STAGING_BYTES;
    uint8_t* stg_row = stg + lane * 128;      // row-order access: thread == row
    const int sw = lane & 7;
    // coalesced residual access: lane covers row (i*4 + lane/8), 16-byte chunk (lane % 8)
    const int crow = lane >> 3, cchunk = lane & 7;
    const uint32_t tmem_lane = static_cast<uint32_t>(wq * 32) << 16;
    {
      const int row0 = row_slab0 + wq * 32;   // first row of this warp's 32-row slab
      int row = row0 + lane;
      int item = b;
      bool valid_row = row < p.rows_out;
      if (p.pack_rpb > 0) {   // packed-item tile (direct-store mode only): tile row -> (item, row of the item)
        const int gi = row / p.pack_rpb;
        item = b + gi;
        row -= gi * p.pack_rpb;
        valid_row = gi < p.pack_items && item < p.batches;
      }
      const int roff = (p.row_off_ptr != nullptr && valid_row) ? __ldg(p.row_off_ptr + item * p.row_off_stride) : 0;
      const long long off32 = static_cast<long long>(item) * p.pitch32 + static_cast<long long>(row + roff) * p.ld32;
      const long long off16 = static_cast<long long>(item) * p.pitch16 + static_cast<long long>(row + roff) * p.ld16;
      const int ncol0 = n_idx * BN + hsel * (BN / 2);     // first column owned by this warp
      const uint32_t tmem_acc = tmem_stage + static_cast<uint32_t>(hsel * (BN / 2)) + tmem_lane;

      // folded LayerNorm, consumer side: this thread's row (mean, rstd), computed by row_stats_kernel from the fp16 copy
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (p.stats_in != nullptr && valid_row) {
        const float2 t = __ldg(p.stats_in + static_cast<long long>(item) * p.rows_out + row);
        ln_mean = t.x;
        ln_rstd = t.y;
      }

      // residual prefetch of the first chunk (coalesced; overlaps the wait for the accumulator)
      float4 rpre[8];
      const bool use_resid = (p.store_mode == STORE_TMA32) && (p.resid != nullptr);
      auto prefetch_resid = [&](int n0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = row0 + i * 4 + crow;
          rpre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr < p.rows_out && n0 + cchunk * 4 < p.N)
            rpre[i] = *reinterpret_cast<const float4*>(p.resid + static_cast<long long>(b) * p.pitch32 +
                                                       static_cast<long long>(rr) * p.ld32 + n0 + cchunk * 4);
        }
      };
      if (use_resid && ncol0 < p.N) prefetch_resid(ncol0);
      // folded LayerNorm, producer side: offsets of the 8 rows this lane covers in the coalesced order
      float xsh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = row0 + i * 4 + crow;
        xsh[i] = (p.x16_shift != nullptr && rr < p.rows_out)
                     ? __ldg(p.x16_shift + static_cast<long long>(b) * p.rows_out + rr) : 0.f;
      }

      ptx::mbar_wait(tfull, aphase);
      ptx::tc_fence_after();

      if (p.store_mode == STORE_TMA16) {
        // ---- fp16 output: 64 columns (128 B per row) per TMA store
#pragma unroll 1
        for (int g = 0; g < CHUNKS_PER_WARP / 2; ++g) {
          const int n0 = ncol0 + g * 64;
          if (n0 >= p.N) break;
          uint32_t packed[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            float v[32];
            ptx::tmem_ld32(tmem_acc + g * 64 + hh * 32, r);
            ptx::tmem_ld_wait();
            apply_chunk(p, r, n0 + hh * 32, v, ln_mean, ln_rstd);
#pragma unroll
            for (int j = 0; j < 16; ++j) packed[hh * 16 + j] = pack_half2(v[2 * j], v[2 * j + 1]);
          }
          if (lane == 0) tma_store_wait_read();  // previous store has finished reading the staging buffer
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 q;
            q.x = packed[4 * c + 0]; q.y = packed[4 * c + 1]; q.z = packed[4 * c + 2]; q.w = packed[4 * c + 3];
            *reinterpret_cast<uint4*>(stg_row + ((c ^ sw) << 4)) = q;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_3d(tmC, stg, n0, row0, b);
        }
      } else if (p.store_mode == STORE_TMA32) {
        // ---- fp32 output (optionally + residual): 32 columns (128 B per row) per TMA store
#pragma unroll 1
        for (int c = 0; c < CHUNKS_PER_WARP; ++c) {
          const int n0 = ncol0 + c * 32;
          if (n0 >= p.N) break;
          uint32_t r[32];
          float v[32];
          ptx::tmem_ld32(tmem_acc + c * 32, r);
          ptx::tmem_ld_wait();
          apply_chunk(p, r, n0, v, ln_mean, ln_rstd);
          if (use_resid) {
            // accumulator rows -> staging (row order) -> back in coalesced order, + prefetched residual, then plain
            // coalesced 16-byte global stores (4 full 128-B lines per warp instruction): half the shared-memory
            // traffic of staging both the residual and the result
            __syncwarp();
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
              *reinterpret_cast<float4*>(stg_row + ((cc ^ sw) << 4)) =
                  make_float4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
            __syncwarp();
            const bool col_ok = n0 + cchunk * 4 < p.N;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = i * 4 + crow;
              float4 q = *reinterpret_cast<const float4*>(stg + rr * 128 + ((cchunk ^ (rr & 7)) << 4));
              q.x += rpre[i].x; q.y += rpre[i].y; q.z += rpre[i].z; q.w += rpre[i].w;
              if (row0 + rr < p.rows_out && col_ok) {
                *reinterpret_cast<float4*>(p.out32 + static_cast<long long>(b) * p.pitch32 +
                                           static_cast<long long>(row0 + rr) * p.ld32 + n0 + cchunk * 4) = q;
                if (p.x16_out != nullptr) {   // folded LayerNorm, producer side: fp16 copy of the residual stream
                  uint2 h;
                  h.x = pack_half2(q.x - xsh[i], q.y - xsh[i]);
                  h.y = pack_half2(q.z - xsh[i], q.w - xsh[i]);
                  *reinterpret_cast<uint2*>(p.x16_out + (static_cast<long long>(b) * p.rows_out + row0 + rr) * p.ld_x16 +
                                            n0 + cchunk * 4) = h;
                }
              }
            }
            if (c + 1 < CHUNKS_PER_WARP && n0 + 32 < p.N) prefetch_resid(n0 + 32);  // in flight during the next chunk
            continue;
          }
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < 8; ++cc)
            *reinterpret_cast<float4*>(stg_row + ((cc ^ sw) << 4)) =
                make_float4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_3d(tmC, stg, n0, row0, b);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < CHUNKS_PER_WARP; ++c) {
          const int n0 = ncol0 + c * 32;
          if (n0 >= p.N) break;
          uint32_t r[32];
          float v[32];
          ptx::tmem_ld32(tmem_acc + c * 32, r);
          ptx::tmem_ld_wait();
          if (valid_row) {
            apply_chunk(p, r, n0, v, ln_mean, ln_rstd);
            store_chunk_direct(p, v, off32, off16, n0);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sStage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + EPI_WARPS * STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (p.store_mode != STORE_DIRECT) ptx::prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_idx = tile % p.tiles_n;
        const int mb = tile / p.tiles_n;
        const int b = (mb / p.tiles_m) * p.item_mul;
        const int m0 = (mb % p.tiles_m) * BM;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], p.stage_tx_bytes);
          const int tap = kb / p.kblocks_per_tap;
          const int c0 = (kb - tap * p.kblocks_per_tap) * BK;
          ptx::tma_load_3d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], c0, m0 + tap + p.row_shift, b);
          ptx::tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK + b * p.w_batch_k, n_idx * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t db = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sB + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
            ptx::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull_bar[as]);       // accumulator ready for the epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int e = warp - 4;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_idx = tile % p.tiles_n;
      const int mb = tile / p.tiles_n;
      const int b = (mb / p.tiles_m) * p.item_mul;
      epilogue_tile<BN>(p, &tmC, sStage, e, lane, tmem_base + static_cast<uint32_t>(as * BN), n_idx, b,
                        (mb % p.tiles_m) * BM, &tfull_bar[as], aphase);
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0 && p.store_mode != STORE_DIRECT) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =====================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs on the two SMs of a TPC computes one 256 x 256 tile.  Each
// CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256 N-rows), so the shared-memory operand
// traffic per SM drops from 12 KB to 8 KB per K=16 step and 6 pipeline stages fit; one thread of the leader CTA
// issues tcgen05.mma.cta_group::2 for both SMs, each CTA's TMEM holds its own 128 x 256 fp32 accumulator (x2 stages)
// and runs the unchanged epilogue on it.  Barriers: both CTAs' TMA loads report their bytes to the leader's "full"
// barrier; MMA completion is multicast to the "empty"/"TMEM full" barriers of both CTAs; the peer's epilogue warps
// arrive remotely on the leader's "TMEM empty" barrier.
// =====================================================================================================
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int STAGES = 6;
  static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;           // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // 32 KB
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STAGING_BYTES + 1024 + 256;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmKParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sStage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + EPI_WARPS * STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (p.store_mode != STORE_DIRECT) ptx::prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);            // leader's: its own arrive.expect_tx (bytes of BOTH CTAs)
      ptx::mbar_init(&empty_bar[s], 1);           // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);           // multicast commit
      ptx::mbar_init(&tempty_bar[s], 2 * EPI_WARPS);  // leader's: epilogue warps of both CTAs
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();   // barrier inits and TMEM allocation of both CTAs visible before any cross-CTA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs) {
        const int n_idx = tile % p.tiles_n;
        const int mb = tile / p.tiles_n;
        const int b = mb / p.tiles_m;
        const int m0 = (mb % p.tiles_m) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n0 = n_idx * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const int tap = kb / p.kblocks_per_tap;
          const int c0 = (kb - tap * p.kblocks_per_tap) * BK;
          ptx::tma_load_3d_2cta(sA + stage * A_STAGE_BYTES, &tmA, full_leader, c0, m0 + tap + p.row_shift, b);
          ptx::tma_load_2d_2cta(sB + stage * Cfg::B_STAGE_BYTES, &tmB, full_leader, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, single thread)
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs) {
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t db = ptx::make_desc_kmajor_sw128(ptx::smem_u32(sB + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::mma_f16_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::mma_commit_2cta(&empty_bar[stage], 0x3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_2cta(&tfull_bar[as], 0x3);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (each CTA on its own 128 rows)
    const int e = warp - 4;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs) {
      const int n_idx = tile % p.tiles_n;
      const int mb = tile / p.tiles_n;
      const int b = mb / p.tiles_m;
      epilogue_tile<BN>(p, &tmC, sStage, e, lane, tmem_base + static_cast<uint32_t>(as * BN), n_idx, b,
                        (mb % p.tiles_m) * (2 * BM) + static_cast<int>(rank) * BM, &tfull_bar[as], aphase);
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0 && p.store_mode != STORE_DIRECT) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();   // the peer's smem / TMEM must stay alive until the leader's last MMA has retired
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::once_flag g_once;
int g_init_status = FRT2_OK;

void do_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    g_init_status = FRT2_ERR_CUDA;
    return;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::SMEM_BYTES);
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(gemm_tc_kernel): ") + cudaGetErrorString(e));
    g_init_status = FRT2_ERR_CUDA;
  }
}

int tma_encode(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box) {
  FRT2_TRY(gemm_tc_init());
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (rank " +
              std::to_string(rank) + ", dims " + std::to_string(dims[0]) + "x" + std::to_string(dims[1]) +
              ", stride0 " + std::to_string(strides_bytes[0]) + ")");
    return FRT2_ERR_CUDA;
  }
  return FRT2_OK;
}

}  // namespace

int gemm_tc_init() {
  std::call_once(g_once, do_init);
  return g_init_status;
}

int tma_encode_fp16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return tma_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box);
}

int num_sms() {
  gemm_tc_init();
  return g_num_sms;
}

int gemm_tc(const GemmDesc& g, cudaStream_t stream) {
  FRT2_TRY(gemm_tc_init());
  FRT2_REQUIRE(g.Kc % BK == 0 && g.Kc > 0, FRT2_ERR_BAD_ARG, "gemm_tc: Kc must be a positive multiple of 64");
  FRT2_REQUIRE(g.ntaps >= 1 && g.N >= 1 && g.rows_out >= 1 && g.batches >= 1, FRT2_ERR_BAD_ARG, "gemm_tc: bad shape");
  FRT2_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0,
               FRT2_ERR_BAD_ARG, "gemm_tc: operands must be 16-byte aligned");
  FRT2_REQUIRE(g.a_row_pitch % 8 == 0 && g.a_batch_pitch % 8 == 0, FRT2_ERR_BAD_ARG,
               "gemm_tc: A pitches must be multiples of 8 elements");
  FRT2_REQUIRE(g.act != ACT_POLAR || (g.N % 2 == 0), FRT2_ERR_BAD_ARG, "gemm_tc: polar epilogue needs even N");
  FRT2_REQUIRE(g.out32 != nullptr || g.out16 != nullptr, FRT2_ERR_BAD_ARG, "gemm_tc: no output");

  // Packed-item tiles: a batch of items with few rows each (the 8 frames per token of a batch of streams / pool
  // slots) would otherwise get one mostly empty 128-row tile per item, each re-streaming the whole weight panel.
  static const bool no_pack = (getenv("FRT2_GEMM_NOPACK") != nullptr);   // A/B switch for measurements
  const bool packed = !no_pack && g.batches > 1 && g.rows_out <= BM / 2 && g.w_batch_k == 0;
  const int pack_items = packed ? std::min(BM / g.rows_out, g.batches) : 0;
  const int groups = packed ? (g.batches + pack_items - 1) / pack_items : g.batches;
  // few row tiles (latency-bound steps): narrower N tiles put more SMs on the weight stream
  const long long row_tiles = packed ? groups : static_cast<long long>(g.batches) * ((g.rows_out + BM - 1) / BM);
  const long long tiles256 = row_tiles * ((g.N + 255) / 256);
  const int BN = (g.N >= 512 && tiles256 >= (g.narrow_tiles ? num_sms() : 64)) ? 256 : 128;
  // CTA pairs (256 x 256 tiles) for the large GEMMs; FRT2_GEMM_1CTA=1 forces the single-CTA kernel (A/B testing)
  static const bool force_1cta = (getenv("FRT2_GEMM_1CTA") != nullptr);
  const bool pair = !force_1cta && BN == 256 && g.rows_out >= 256 && g.out_row_off == nullptr && g.w_batch_k == 0;
  const int BMT = pair ? 2 * BM : BM;           // rows per tile
  const int BNB = pair ? BN / 2 : BN;           // W rows per TMA box
  const uint64_t Ktot = static_cast<uint64_t>(g.ntaps) * g.Kc;

  CUtensorMap tmA, tmB, tmC;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(g.Kc), static_cast<uint64_t>(g.rows_a), static_cast<uint64_t>(g.batches)};
    uint64_t strides[2] = {static_cast<uint64_t>(g.a_row_pitch) * 2,
                           static_cast<uint64_t>(g.batches > 1 ? g.a_batch_pitch : g.a_row_pitch * g.rows_a) * 2};
    uint32_t box[3] = {BK, BM, 1};
    if (packed) {
      box[1] = static_cast<uint32_t>(g.rows_out);
      box[2] = static_cast<uint32_t>(pack_items);
    }
    FRT2_TRY(tma_encode_fp16(&tmA, g.A, 3, dims, strides, box));
  }
  {
    const uint64_t Kw = Ktot + static_cast<uint64_t>(g.batches - 1) * g.w_batch_k;    // W row length (all reduction splits)
    uint64_t dims[2] = {Kw, static_cast<uint64_t>(g.N)};
    uint64_t strides[1] = {Kw * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(BNB)};
    FRT2_TRY(tma_encode_fp16(&tmB, g.W, 2, dims, strides, box));
  }

  GemmKParams p;
  p.rows_out = g.rows_out;
  p.pitch32 = g.pitch32;
  p.pitch16 = g.pitch16;
  p.N = g.N;
  p.kblocks_per_tap = g.Kc / BK;
  p.num_kblocks = g.ntaps * p.kblocks_per_tap;
  p.row_shift = g.row_shift;
  p.tiles_m = packed ? 1 : (g.rows_out + BMT - 1) / BMT;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.num_tiles = groups * p.tiles_m * p.tiles_n;
  p.pack_rpb = packed ? g.rows_out : 0;
  p.pack_items = pack_items;
  p.batches = g.batches;
  p.item_mul = packed ? pack_items : 1;
  p.w_batch_k = g.w_batch_k;
  p.stage_tx_bytes = (packed ? pack_items * g.rows_out * BK * 2 : A_STAGE_BYTES) + BN * BK * 2;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.act = g.act;
  p.resid = g.resid;
  p.out32 = g.out32;
  p.ld32 = g.ld32;
  p.out16 = g.out16;
  p.ld16 = g.ld16;
  p.vec32 = (g.ld32 % 4 == 0 && g.pitch32 % 4 == 0 && (reinterpret_cast<uintptr_t>(g.out32) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(g.resid) & 15) == 0);
  p.vec16 = (g.ld16 % 8 == 0 && g.pitch16 % 8 == 0 && (reinterpret_cast<uintptr_t>(g.out16) & 15) == 0);

  // output through TMA stores whenever the layout allows it (16-byte aligned base / pitches, single output)
  p.store_mode = STORE_DIRECT;
  p.row_off_ptr = g.out_row_off;
  p.row_off_stride = g.row_off_stride;
  p.x16_out = g.x16_out;
  p.ld_x16 = g.ld_x16;
  p.x16_shift = g.x16_out != nullptr ? g.x16_shift : nullptr;
  p.stats_in = g.stats_in;
  p.colsum = g.stats_in != nullptr ? g.colsum : nullptr;
  if (g.stats_in != nullptr) {
    FRT2_REQUIRE(g.colsum != nullptr && g.bias != nullptr && !packed &&
                     (reinterpret_cast<uintptr_t>(g.colsum) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0,
                 FRT2_ERR_BAD_ARG, "gemm_tc: folded LayerNorm consumer needs colsum and bias (16-byte aligned)");
  }
  tmC = tmA;
  const uint64_t batch_rows = static_cast<uint64_t>(g.batches);
  if (g.out_row_off != nullptr || packed) {
    // run-time row offset (streaming K/V append) or packed-item tile: plain predicated stores, one row per thread
  } else if (g.out16 != nullptr && g.out32 == nullptr && g.resid == nullptr && p.vec16) {
    uint64_t dims[3] = {static_cast<uint64_t>(g.N), static_cast<uint64_t>(g.rows_out), batch_rows};
    uint64_t strides[2] = {static_cast<uint64_t>(g.ld16) * 2,
                           static_cast<uint64_t>(g.batches > 1 ? g.pitch16 : g.ld16 * g.rows_out) * 2};
    uint32_t box[3] = {64, 32, 1};
    FRT2_TRY(tma_encode(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, g.out16, 3, dims, strides, box));
    p.store_mode = STORE_TMA16;
  } else if (g.out32 != nullptr && g.out16 == nullptr && p.vec32 && (g.resid == nullptr || g.N % 4 == 0)) {
    uint64_t dims[3] = {static_cast<uint64_t>(g.N), static_cast<uint64_t>(g.rows_out), batch_rows};
    uint64_t strides[2] = {static_cast<uint64_t>(g.ld32) * 4,
                           static_cast<uint64_t>(g.batches > 1 ? g.pitch32 : g.ld32 * g.rows_out) * 4};
    uint32_t box[3] = {32, 32, 1};
    FRT2_TRY(tma_encode(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, g.out32, 3, dims, strides, box));
    p.store_mode = STORE_TMA32;
  }

  if (g.x16_out != nullptr) {
    FRT2_REQUIRE(p.store_mode == STORE_TMA32 && g.resid != nullptr && g.N % 4 == 0 && g.ld_x16 % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(g.x16_out) & 7) == 0,
                 FRT2_ERR_BAD_ARG, "gemm_tc: the fp16 copy of the residual stream needs the fp32 residual epilogue");
    // the copy is written by the epilogue of early tiles while later tiles still load A: the two must not overlap
    const char* a_lo = reinterpret_cast<const char*>(g.A);
    const char* a_hi = a_lo + (static_cast<int64_t>(g.batches - 1) * g.a_batch_pitch +
                               static_cast<int64_t>(g.rows_a) * g.a_row_pitch) * 2;
    const char* x_lo = reinterpret_cast<const char*>(g.x16_out);
    const char* x_hi = x_lo + static_cast<int64_t>(g.batches) * g.rows_out * g.ld_x16 * 2;
    FRT2_REQUIRE(a_hi <= x_lo || x_hi <= a_lo, FRT2_ERR_BAD_ARG,
                 "gemm_tc: the fp16 copy of the output rows must not alias the A operand");
  }
  if (pair) {
    const int grid = 2 * std::min(p.num_tiles, g_num_sms / 2);
    gemm_tc2_kernel<<<grid, GEMM_THREADS, Gemm2Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
  } else {
    const int grid = std::min(p.num_tiles, g_num_sms);
    if (BN == 256) {
      gemm_tc_kernel<256><<<grid, GEMM_THREADS, GemmCfg<256>::SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
    } else {
      gemm_tc_kernel<128><<<grid, GEMM_THREADS, GemmCfg<128>::SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
    }
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
