// Frame tail of the speech LM: what `Model.generate_frame` (reference fireredtts2/llm/llm.py:274-330)
// runs after the backbone —
//
//   c0_logits = codebook0_head(last_h); c0 = sample_topk(c0_logits, topk, temperature)          llm.py:305-306
//   decoder K/V state reset; positions 0 (projection(last_h)) and 1 (projection(embed(0, c0)))   llm.py:307-322
//   for i = 1 .. ncb-1:  h = decoder(projection(curr))[:, -1]; logits = h @ audio_head[i-1];
//                        c_i = sample_topk(logits, 10, 0.75); curr = embed(i, c_i)                llm.py:318-328
//
// `decoder` is torchtune's qwen2 (modules.py:5-82): RMSNorm (eps 1e-6) -> q|k|v projections with bias -> rotary positions
// on the halves of every head -> grouped-query attention over the frame's <= ncb positions -> output projection ->
// RMSNorm -> SwiGLU -> final RMSNorm.  At batch 1 the whole tail is a chain of ncb dependent matrix-VECTOR passes: every
// weight byte is used once per position, so the tail is bound by streaming the decoder's fp16 weights ncb times from HBM
// (5.6 GB per frame for the 200M decoder: they do not fit the 126 MB L2).  Everything therefore runs on the weight-
// streaming kernel of the codec's token step (gemm_skinny: 8 warps split K, 16-byte weight vectors, mma.sync m16n8k16,
// programmatic dependent launch so the next kernel's weights are in flight while the current one drains) with
//   * RMSNorm computed on the fly from the fp32 residual row in the GEMM's prologue (no norm kernels),
//   * q|k|v as ONE launch, gate|up as ONE launch whose epilogue applies silu(gate)*up on interleaved weight rows,
//   * residual adds in the output-projection / down-projection epilogues,
// plus two small kernels per position: rotary + K/V append + attention (fd_attn_kernel) and top-k / softmax / sampling /
// next-input embedding lookup (fd_sample_kernel).  The (1 + 5 layers) * ncb + ... launches of a frame are ONE CUDA graph;
// codes stay on the device (they are the token operand of frt2_decode_chunk).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace frt2 {
namespace {

constexpr int FD_MAX_BATCH = 16;      // rows the weight-streaming path takes (the m16 of its MMA tile)
constexpr int FD_MAX_POS = 64;        // audio_num_codebooks limit (lane t of the attention warp owns positions t and t + 32)
constexpr int FD_SAMPLE_THREADS = 256;

// run-time parameters of a frame, written by fd_begin_kernel, read by the captured kernels
struct FdParams {
  const float* noise;      // (B, ncb, V) Exp(1) draws or null -> counter-based generator
  unsigned long long seed;
  unsigned long long frame;
  int topk;                // codebook 0 (llm.py:306); codebooks >= 1 use 10 / 0.75 (llm.py:324)
  float temperature;
  int has_c0;              // codebook-0 code given by the caller
  int has_forced;          // teacher forcing: every code given
};

__global__ void fd_begin_kernel(const float* __restrict__ last_h, int B, int Db, __half* __restrict__ in16,
                                const int* __restrict__ c0, const int* __restrict__ forced, int ncb, int* __restrict__ given,
                                FdParams p, FdParams* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *dst = p;
  if (i < B * ncb) given[i] = forced != nullptr ? forced[i] : ((c0 != nullptr && i % ncb == 0) ? c0[i / ncb] : 0);
  for (int e = i; e < B * Db; e += gridDim.x * blockDim.x) in16[e] = to_half_sat(last_h[e]);
}

// One position of one layer (or the first TWO, np = 2: rows 2b and 2b+1 of qkv / out16 are positions 0 and 1 of item b):
// rotary embedding of q and the new k, K/V append, attention over positions 0..pos.
// grid (Hk, B); one warp per query head of the kv group.  qkv: (B * np, (H + 2 Hk) hd) fp32 = [q | k | v] with bias.
// kc / vc: this layer's (B, npos, Hk, hd) fp32 state.  Scores are 1/sqrt(hd)-scaled dot products, softmax in fp32.
// Latency is all that matters here (a few KB of state between two weight streams): lane t owns position t — every lane
// reads ITS key row in one batch of independent 16-byte loads (one round trip for all positions), the softmax is a pair
// of warp reductions, and the value rows are read coalesced (lane = channel) with the weights broadcast by shuffles.
__global__ void __launch_bounds__(256) fd_attn_kernel(const float* __restrict__ qkv, float* __restrict__ kc,
                                                      float* __restrict__ vc, __half* __restrict__ out16,
                                                      const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                                                      int H, int Hk, int hd, int npos, int pos0, int np, float scale) {
  __shared__ __align__(16) float s_q[8][128];          // rotated query of each warp's head
  __shared__ __align__(16) float s_k[8][2][128];       // rotated new keys / new values (one copy per warp: no CTA barrier)
  __shared__ __align__(16) float s_v[8][2][128];
  const int kvh = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rep = H / Hk, half = hd >> 1;
  const int h = kvh * rep + warp;
  const int qkv_ld = (H + 2 * Hk) * hd;
  const long long row = static_cast<long long>(Hk) * hd;                      // one position of one item
  float* kcb = kc + (static_cast<long long>(b) * npos) * row + kvh * hd;
  float* vcb = vc + (static_cast<long long>(b) * npos) * row + kvh * hd;
  const int nv4 = hd >> 2;
  for (int j = 0; j < np; ++j) {
    const int pos = pos0 + j;
    const float* kr = qkv + static_cast<long long>(b * np + j) * qkv_ld + (H + kvh) * hd;
    const float* vr = qkv + static_cast<long long>(b * np + j) * qkv_ld + (H + Hk + kvh) * hd;
    for (int i = lane; i < half; i += 32) {
      const float c = rope_cos[pos * half + i], s = rope_sin[pos * half + i];
      const float ka = kr[i], kb = kr[i + half];
      const float k1 = ka * c - kb * s, k2 = kb * c + ka * s;
      const float v1 = vr[i], v2 = vr[i + half];
      s_k[warp][j][i] = k1;
      s_k[warp][j][i + half] = k2;
      s_v[warp][j][i] = v1;
      s_v[warp][j][i + half] = v2;
      if (warp == 0) {
        kcb[pos * row + i] = k1;
        kcb[pos * row + i + half] = k2;
        vcb[pos * row + i] = v1;
        vcb[pos * row + i + half] = v2;
      }
    }
  }
  for (int j = 0; j < np; ++j) {
    const int pos = pos0 + j;
    const float* qr = qkv + static_cast<long long>(b * np + j) * qkv_ld + h * hd;
    __syncwarp();
    for (int i = lane; i < half; i += 32) {
      const float c = rope_cos[pos * half + i], s = rope_sin[pos * half + i];
      const float a = qr[i], bq = qr[i + half];
      s_q[warp][i] = a * c - bq * s;
      s_q[warp][i + half] = bq * c + a * s;
    }
    __syncwarp();
    auto score = [&](int t) -> float {
      if (t > pos) return -INFINITY;
      const float4* kp = t >= pos0 ? reinterpret_cast<const float4*>(s_k[warp][t - pos0])
                                   : reinterpret_cast<const float4*>(kcb + t * row);
      const float4* qp = reinterpret_cast<const float4*>(s_q[warp]);
      float acc = 0.f;
#pragma unroll 8
      for (int i = 0; i < nv4; ++i) {
        const float4 kv = kp[i], qv = qp[i];
        acc += (kv.x * qv.x + kv.y * qv.y) + (kv.z * qv.z + kv.w * qv.w);
      }
      return acc * scale;
    };
    const int t0 = lane, t1 = lane + 32;
    float sc0 = score(t0), sc1 = pos >= 32 ? score(t1) : -INFINITY;
    float mx = fmaxf(sc0, sc1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sc0 = t0 <= pos ? expf(sc0 - mx) : 0.f;
    sc1 = t1 <= pos ? expf(sc1 - mx) : 0.f;
    float den = sc0 + sc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    const float inv = 1.0f / den;
    sc0 *= inv;
    sc1 *= inv;
    float o_acc[4] = {0.f, 0.f, 0.f, 0.f};              // channels lane, lane + 32, lane + 64, lane + 96
#pragma unroll 4
    for (int t = 0; t <= pos; ++t) {
      const float a = __shfl_sync(0xffffffffu, t < 32 ? sc0 : sc1, t & 31);
      const float* vp = t >= pos0 ? s_v[warp][t - pos0] : vcb + t * row;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int dch = lane + 32 * jj;
        if (dch < hd) o_acc[jj] += a * vp[dch];
      }
    }
    __half* orow = out16 + static_cast<long long>(b * np + j) * H * hd + h * hd;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int dch = lane + 32 * jj;
      if (dch < hd) orow[dch] = to_half_sat(o_acc[jj]);
    }
  }
}

// Large-batch path (> 16 items per frame, tcgen05 GEMMs): RMSNorm rows fp32 -> fp16 (one warp per row, the row kept in
// registers for widths <= 4096) and silu(gate) * up on the interleaved (gate_j, up_j) columns of the merged GEMM's output
__global__ void __launch_bounds__(256) fd_rms_rows_kernel(const float* __restrict__ x, long long ldx, int rows, int C,
                                                          const float* __restrict__ gamma, float eps,
                                                          __half* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int C4 = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
  float q = 0.f;
  for (int c = lane; c < C4; c += 32) {
    const float4 v = xr[c];
    q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  __half* orow = out + static_cast<long long>(row) * C;
  for (int c = lane; c < C4; c += 32) {
    const float4 v = xr[c], gg = __ldg(g4 + c);
    uint2 h;
    h.x = pack_half2(v.x * rstd * gg.x, v.y * rstd * gg.y);
    h.y = pack_half2(v.z * rstd * gg.z, v.w * rstd * gg.w);
    *reinterpret_cast<uint2*>(orow + 4 * c) = h;
  }
}
__global__ void __launch_bounds__(256) fd_swiglu_rows_kernel(const __half2* __restrict__ gu, long long n,
                                                             __half* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p = __half22float2(gu[i]);      // (gate_j, up_j)
  out[i] = to_half_sat(p.x / (1.0f + expf(-p.x)) * p.y);
}

// dst[r][n] = (accumulate ? dst[r][n] : bias[n] or 0) + sum_s part[s][r][n] in the order s = 0, 1, ... (deterministic); four
// elements per thread; part rows are N wide, dst rows ld_dst
__global__ void __launch_bounds__(256) fd_splitk_reduce_kernel(const float4* __restrict__ part, int rows, int N4, int S,
                                                               const float4* __restrict__ bias, int accumulate,
                                                               float* __restrict__ dst, long long ld_dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n4 = static_cast<long long>(rows) * N4;
  if (i >= n4) return;
  const int r = static_cast<int>(i / N4), c = static_cast<int>(i - static_cast<long long>(r) * N4);
  float4* dp = reinterpret_cast<float4*>(dst + r * ld_dst) + c;
  float4 a = accumulate ? *dp : (bias != nullptr ? __ldg(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f));
  for (int s = 0; s < S; ++s) {
    const float4 v = part[s * n4 + i];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *dp = a;
}

// Philox-4x32-10 (counter-based: the draw of (frame, item, codebook, entry) does not depend on launch geometry)
__device__ __forceinline__ float fd_exp1_draw(unsigned long long seed, unsigned long long frame, int b, int s, int v) {
  uint32_t c0 = static_cast<uint32_t>(v), c1 = static_cast<uint32_t>(s) | (static_cast<uint32_t>(b) << 16);
  uint32_t c2 = static_cast<uint32_t>(frame), c3 = static_cast<uint32_t>(frame >> 32);
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u = (static_cast<float>(c0 >> 8) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
  return fmaxf(-logf(u), 1e-30f);
}

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {   // larger value, then lower index (first maximum)
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax x, ArgMax* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                 // red[] of the previous round has been consumed
  if (lane == 0) red[warp] = x;
  __syncthreads();
  ArgMax r = red[0];
#pragma unroll
  for (int w = 1; w < FD_SAMPLE_THREADS / 32; ++w) r = better(r, red[w]);
  return r;                        // the same in every thread
}
__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < FD_SAMPLE_THREADS / 32; ++w) r += red[w];
  return r;
}

// sample_topk + _multinomial_sample_one_no_sync (llm.py:34-49) for codebook s of item blockIdx.x, then the embedding row
// of the next decoder input, already projected (llm.py:307,321,325-326: projection(audio_embeddings[code + s * V]), a
// table composed at load), as the next position's residual row.
//   logits / temperature; k-th largest value by k rounds of "remove the first maximum" (multiplicity counted like
//   torch.topk); entries < k-th are dropped (ties at the k-th value stay, llm.py:43); log_softmax, softmax, p / q, first
//   arg-max.  Everything in fp32 in the reference's order of operations.
__global__ void __launch_bounds__(FD_SAMPLE_THREADS) fd_sample_kernel(
    const float* __restrict__ logits, int s, int V, int ncb, const FdParams* __restrict__ pp, const int* __restrict__ given,
    int* __restrict__ codes, const float* __restrict__ table, int D, float* __restrict__ x_next, int x_ld,
    unsigned int* __restrict__ err_word) {
  extern __shared__ float fd_smem[];
  float* val = fd_smem;          // scaled logits
  float* work = fd_smem + V;     // copy consumed by the selection rounds
  __shared__ ArgMax red_a[FD_SAMPLE_THREADS / 32];
  __shared__ float red_f[FD_SAMPLE_THREADS / 32];
  __shared__ int s_code;
  const int b = blockIdx.x, tid = threadIdx.x;
  const FdParams p = *pp;
  const bool use_given = p.has_forced || (s == 0 && p.has_c0);
  int code;
  if (use_given) {
    code = given[b * ncb + s];
    if (code < 0 || code >= V) {   // nn.Embedding would raise IndexError: reported through the error word, row 0 used
      if (tid == 0) atomicOr(err_word, DEV_ERR_INDEX_OOR);
      code = 0;
    }
  } else {
    const float temperature = s == 0 ? p.temperature : 0.75f;
    const int topk = min(s == 0 ? p.topk : 10, V);
    const float* lrow = logits + (static_cast<long long>(b) * ncb + s) * V;
    for (int v = tid; v < V; v += FD_SAMPLE_THREADS) {
      const float x = lrow[v] / temperature;
      val[v] = x;
      work[v] = x;
    }
    __syncthreads();
    float kth = 0.f, top = 0.f;
    for (int r = 0; r < topk; ++r) {
      ArgMax m;
      m.v = -INFINITY;
      m.i = 0x7fffffff;
      for (int v = tid; v < V; v += FD_SAMPLE_THREADS) {
        ArgMax c;
        c.v = work[v];
        c.i = v;
        m = better(m, c);
      }
      m = block_argmax(m, red_a);
      if (r == 0) top = m.v;
      kth = m.v;
      if (tid == 0 && m.i < V) work[m.i] = -INFINITY;
      __syncthreads();
    }
    // log_softmax over the kept entries, then softmax of that (llm.py:45-46)
    float part = 0.f;
    for (int v = tid; v < V; v += FD_SAMPLE_THREADS)
      if (val[v] >= kth) part += expf(val[v] - top);
    const float lse = logf(block_sum(part, red_f));
    part = 0.f;
    for (int v = tid; v < V; v += FD_SAMPLE_THREADS)
      if (val[v] >= kth) part += expf(((val[v] - top) - lse) + lse);      // ls - max(ls), max(ls) = -lse
    const float den = block_sum(part, red_f);
    ArgMax m;
    m.v = -INFINITY;
    m.i = 0x7fffffff;
    for (int v = tid; v < V; v += FD_SAMPLE_THREADS) {
      if (val[v] >= kth) {
        const float q = p.noise != nullptr ? p.noise[(static_cast<long long>(b) * ncb + s) * V + v]
                                           : fd_exp1_draw(p.seed, p.frame, b, s, v);
        ArgMax c;
        c.v = (expf(((val[v] - top) - lse) + lse) / den) / q;
        c.i = v;
        m = better(m, c);
      }
    }
    m = block_argmax(m, red_a);
    code = m.i < V ? m.i : 0;
  }
  if (tid == 0) {
    codes[b * ncb + s] = code;
    s_code = code;
  }
  __syncthreads();
  code = s_code;
  if (s + 1 < ncb) {
    // projection(audio_embeddings[code + s V]) from the table composed at load: the residual row of the next position
    const float4* src = reinterpret_cast<const float4*>(table + (static_cast<long long>(s) * V + code) * D);
    float4* dst = reinterpret_cast<float4*>(x_next + static_cast<long long>(b) * x_ld);
    for (int i = tid; i < D / 4; i += FD_SAMPLE_THREADS) dst[i] = __ldg(src + i);
  }
}

struct HostT {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct FdLayer {
  __half* w_qkv = nullptr;  float* b_qkv = nullptr;   // ((H + 2 Hk) hd, D), bias
  __half* w_o = nullptr;                               // (D, H hd)
  __half* w_gu = nullptr;                              // (2 I, D): rows (gate_0, up_0, gate_1, up_1, ...)
  __half* w_down = nullptr;                            // (D, I)
  __half *w_qkv_rm = nullptr, *w_o_rm = nullptr, *w_gu_rm = nullptr, *w_down_rm = nullptr;   // row-major copies (large batches)
  float *g_sa = nullptr, *g_mlp = nullptr;
  float *kc = nullptr, *vc = nullptr;                  // (MAX_BATCH, ncb, Hk, hd) fp32
};

}  // namespace

struct FrameDecoder {
  int device = 0;
  std::mutex mu;
  std::map<std::string, HostT> raw;
  bool finalized = false;
  bool use_stream = true;        // gemm_stream (tile-blocked weights) for every GEMM; false: gemm_skinny (A/B, odd widths)
  frt2_fd_config cfg{};
  int hd = 0, qkv = 0;
  std::vector<void*> owned;
  std::vector<FdLayer> layers;
  __half* w_proj = nullptr;      // (D, Db)
  __half* w_head0 = nullptr;     // (V, Db)
  __half* w_heads = nullptr;     // (ncb - 1, V, D): audio_head[i] transposed (K contiguous)
  size_t head_stride = 0;        // elements between two heads
  bool big = false;              // max_batch > 16: row-major weight copies + tcgen05 GEMMs for frames of 17 .. max_batch items
  int mb = FD_MAX_BATCH;         // rows the activation buffers hold
  __half *w_proj_rm = nullptr, *w_head0_rm = nullptr, *w_heads_rm = nullptr;
  __half *n16 = nullptr, *gu16 = nullptr;
  float* part32 = nullptr;       // partial sums of a GEMM whose reduction is split over CTAs (large batches, narrow N)
  // Large-batch path: a layer with N / 128 column tiles per 128 rows leaves most SMs idle (12 tiles for N = 1536); its
  // reduction is then split over S batch items of ONE launch: S = the largest divisor of the k-blocks that keeps >= 4
  // blocks per split and tiles * S <= SMs (1 = no split)
  int pick_split(int rows, int N, int K) const {
    static const bool off = getenv("FRT2_FD_NO_SPLITK") != nullptr;
    static const int min_k = getenv("FRT2_FD_SPLITK_MIN_K") != nullptr ? atoi(getenv("FRT2_FD_SPLITK_MIN_K")) : 1024;
    if (off || K < min_k || N % 4) return 1;
    const int kblocks = K / 64, tiles = ((rows + 127) / 128) * ((N + 127) / 128);
    int best = 1;
    for (int sdiv = 1; sdiv <= kblocks; ++sdiv)
      if (kblocks % sdiv == 0 && kblocks / sdiv >= 4 && tiles * sdiv <= num_sms()) best = sdiv;
    return best;
  }
  __half* emb16 = nullptr;       // (ncb * V, Db): only until the table below is built
  float* proj_table = nullptr;   // (ncb * V, D) fp32 = projection(audio_embeddings): what a sampled code contributes to the
                                 // next position (composed at load with the frame's own GEMM kernel: the same bits as
                                 // embedding lookup + projection GEMM per position, 15 launches fewer per frame)
  float* g_final = nullptr;
  float *rope_cos = nullptr, *rope_sin = nullptr;     // (ncb, hd / 2)
  // activations of a frame (MAX_BATCH rows)
  __half *in16 = nullptr, *attn16 = nullptr, *h16 = nullptr;
  float *x32 = nullptr, *qkv32 = nullptr, *logits = nullptr;
  int *codes = nullptr, *given = nullptr;
  FdParams* params = nullptr;
  unsigned int* err_word = nullptr;
  unsigned long long frame = 0;
  std::map<int, cudaGraphExec_t> graphs;              // by batch size
  std::map<int, long long> graph_kernels;
  cudaStream_t cap_stream = nullptr;
  cudaEvent_t ws_event = nullptr;
  cudaStream_t ws_last = nullptr;
  bool ws_used = false;
  long long launches = 0;

  ~FrameDecoder() {
    cudaSetDevice(device);
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    for (void* p : owned) cudaFree(p);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (ws_event) cudaEventDestroy(ws_event);
  }
  template <typename T>
  int dev_alloc(T** p, size_t count) {
    void* q = nullptr;
    FRT2_CUDA_OK(cudaMalloc(&q, std::max<size_t>(count * sizeof(T), 16)));
    FRT2_CUDA_OK(cudaMemset(q, 0, std::max<size_t>(count * sizeof(T), 16)));
    owned.push_back(q);
    *p = static_cast<T*>(q);
    return FRT2_OK;
  }
  int up32(const float* v, size_t n, float** out) {
    FRT2_TRY(dev_alloc(out, n));
    FRT2_CUDA_OK(cudaMemcpy(*out, v, n * 4, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }
  int up16(const std::vector<float>& v, __half** out) {
    std::vector<__half> hb(v.size());
    const long long n = static_cast<long long>(v.size());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) hb[i] = __float2half_rn(std::min(65504.0f, std::max(-65504.0f, v[i])));
    FRT2_TRY(dev_alloc(out, hb.size()));
    FRT2_CUDA_OK(cudaMemcpy(*out, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }
  // GEMM weight (N, K) fp32 -> fp16 in the layout of the kernel that will stream it
  int up_w(const float* w, int64_t N, int64_t K, __half** out, __half** rm = nullptr) {
    if (rm != nullptr && big) FRT2_TRY(up16(std::vector<float>(w, w + N * K), rm));
    if (!use_stream) return up16(std::vector<float>(w, w + N * K), out);
    std::vector<__half> packed(gemm_stream_packed_elems(N, K));
    gemm_stream_pack_host(w, N, K, packed.data());
    FRT2_TRY(dev_alloc(out, packed.size()));
    FRT2_CUDA_OK(cudaMemcpy(*out, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }
  int need(const std::string& key, const HostT** out, std::initializer_list<int64_t> shape) {
    auto it = raw.find(key);
    if (it == raw.end()) {
      set_error("missing tensor: " + key);
      return FRT2_ERR_MISSING_TENSOR;
    }
    if (it->second.shape != std::vector<int64_t>(shape)) {
      std::string got, exp;
      for (auto d : it->second.shape) got += std::to_string(d) + ",";
      for (auto d : shape) exp += std::to_string(d) + ",";
      set_error("tensor " + key + " has shape (" + got + ") expected (" + exp + ")");
      return FRT2_ERR_BAD_ARG;
    }
    *out = &it->second;
    return FRT2_OK;
  }
  // positions 0 and 1 as one two-row pass: while 2 B rows fit the weight-streaming kernel's 16, and on the tcgen05 path
  bool pair_layout(int B) const {
    static const bool no_pair = getenv("FRT2_FD_NO_PAIR") != nullptr;
    return !no_pair && (2 * B <= FD_MAX_BATCH || B > FD_MAX_BATCH);
  }
  int finalize();
  int skinny(const __half* A, int K, const __half* W, int N, const float* bias, int act, const float* resid, float* out32,
             int64_t ld32, __half* out16, int64_t ld16, const float* ln_x, const float* ln_gamma, int B, cudaStream_t st,
             int64_t lda = 0, int64_t ldx = 0);
  int enqueue_frame(int B, cudaStream_t st);
  int enqueue_frame_big(int B, cudaStream_t st);
  int tc(const __half* A, int K, const __half* W, int N, const float* bias, const float* resid, float* out32, int64_t ld32,
         __half* out16, int64_t ld16, int rows, cudaStream_t st, int64_t lda = 0);
  int tc32(const __half* A, int K, const __half* W, int N, const float* bias, bool accumulate, float* dst, int64_t ld_dst,
           int rows, cudaStream_t st, int64_t lda = 0);
  int graph_for(int B, cudaGraphExec_t* out);
};

int FrameDecoder::finalize() {
  FRT2_CUDA_OK(cudaSetDevice(device));
  FRT2_TRY(gemm_skinny_init());
  FRT2_TRY(gemm_stream_init());
  use_stream = getenv("FRT2_FD_SKINNY") == nullptr && gemm_stream_applicable(8, cfg.dim, 8) &&
               gemm_stream_applicable(8, cfg.backbone_dim, 8) && gemm_stream_applicable(8, cfg.intermediate_dim, 8) &&
               gemm_stream_max_rows(cfg.dim) >= 16 && gemm_stream_max_rows(cfg.backbone_dim) >= 16;
  const int D = cfg.dim, Db = cfg.backbone_dim, I = cfg.intermediate_dim, V = cfg.audio_vocab_size, n = cfg.audio_num_codebooks;
  const int H = cfg.num_heads, Hk = cfg.num_kv_heads;
  big = cfg.max_batch > FD_MAX_BATCH;
  mb = big ? cfg.max_batch : FD_MAX_BATCH;
  const size_t act_rows = big ? 2 * static_cast<size_t>(mb) : mb;   // the two-row first pass of a large batch
  if (big) FRT2_TRY(gemm_tc_init());
  const HostT* t = nullptr;
  FRT2_TRY(need("projection.weight", &t, {D, Db}));
  FRT2_TRY(up_w(t->data.data(), D, Db, &w_proj, &w_proj_rm));
  FRT2_TRY(need("codebook0_head.weight", &t, {V, Db}));
  FRT2_TRY(up_w(t->data.data(), V, Db, &w_head0, &w_head0_rm));
  FRT2_TRY(need("audio_embeddings.weight", &t, {static_cast<int64_t>(V) * n, Db}));
  FRT2_TRY(up16(t->data, &emb16));
  FRT2_TRY(need("audio_head", &t, {n - 1, D, V}));
  {
    std::vector<float> tr(t->data.size());
    for (int i = 0; i < n - 1; ++i) {
      const float* src = t->data.data() + static_cast<size_t>(i) * D * V;
      float* dst = tr.data() + static_cast<size_t>(i) * D * V;
#pragma omp parallel for schedule(static)
      for (int v = 0; v < V; ++v)
        for (int d = 0; d < D; ++d) dst[static_cast<size_t>(v) * D + d] = src[static_cast<size_t>(d) * V + v];
    }
    if (big) FRT2_TRY(up16(tr, &w_heads_rm));        // (ncb - 1, V, D) row-major
    head_stride = use_stream ? gemm_stream_packed_elems(V, D) : static_cast<size_t>(V) * D;
    std::vector<__half> all(head_stride * (n - 1));
    FRT2_TRY(dev_alloc(&w_heads, all.size()));
    for (int i = 0; i < n - 1; ++i) {
      __half* one = nullptr;   // packed / converted one head at a time, copied into the common buffer
      const size_t mark = owned.size();
      FRT2_TRY(up_w(tr.data() + static_cast<size_t>(i) * D * V, V, D, &one));
      FRT2_CUDA_OK(cudaMemcpy(w_heads + head_stride * i, one, head_stride * 2, cudaMemcpyDeviceToDevice));
      cudaFree(one);
      owned.resize(mark);
    }
  }
  FRT2_TRY(need("decoder.norm.scale", &t, {D}));
  FRT2_TRY(up32(t->data.data(), D, &g_final));
  layers.resize(cfg.num_layers);
  for (int l = 0; l < cfg.num_layers; ++l) {
    FdLayer& L = layers[l];
    const std::string p = "decoder.layers." + std::to_string(l) + ".";
    const HostT *wq, *bq, *wk, *bk, *wv, *bv, *w1, *w3;
    FRT2_TRY(need(p + "attn.q_proj.weight", &wq, {H * hd, D}));
    FRT2_TRY(need(p + "attn.q_proj.bias", &bq, {H * hd}));
    FRT2_TRY(need(p + "attn.k_proj.weight", &wk, {Hk * hd, D}));
    FRT2_TRY(need(p + "attn.k_proj.bias", &bk, {Hk * hd}));
    FRT2_TRY(need(p + "attn.v_proj.weight", &wv, {Hk * hd, D}));
    FRT2_TRY(need(p + "attn.v_proj.bias", &bv, {Hk * hd}));
    std::vector<float> w(static_cast<size_t>(qkv) * D), bias(qkv);
    std::copy(wq->data.begin(), wq->data.end(), w.begin());
    std::copy(wk->data.begin(), wk->data.end(), w.begin() + static_cast<size_t>(H) * hd * D);
    std::copy(wv->data.begin(), wv->data.end(), w.begin() + static_cast<size_t>(H + Hk) * hd * D);
    std::copy(bq->data.begin(), bq->data.end(), bias.begin());
    std::copy(bk->data.begin(), bk->data.end(), bias.begin() + H * hd);
    std::copy(bv->data.begin(), bv->data.end(), bias.begin() + (H + Hk) * hd);
    FRT2_TRY(up_w(w.data(), qkv, D, &L.w_qkv, &L.w_qkv_rm));
    FRT2_TRY(up32(bias.data(), bias.size(), &L.b_qkv));
    FRT2_TRY(need(p + "attn.output_proj.weight", &t, {D, H * hd}));
    FRT2_TRY(up_w(t->data.data(), D, H * hd, &L.w_o, &L.w_o_rm));
    FRT2_TRY(need(p + "mlp.w1.weight", &w1, {I, D}));
    FRT2_TRY(need(p + "mlp.w3.weight", &w3, {I, D}));
    std::vector<float> gu(static_cast<size_t>(2) * I * D);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < I; ++j) {
      std::memcpy(gu.data() + static_cast<size_t>(2 * j) * D, w1->data.data() + static_cast<size_t>(j) * D, D * 4);
      std::memcpy(gu.data() + static_cast<size_t>(2 * j + 1) * D, w3->data.data() + static_cast<size_t>(j) * D, D * 4);
    }
    FRT2_TRY(up_w(gu.data(), 2 * I, D, &L.w_gu, &L.w_gu_rm));
    FRT2_TRY(need(p + "mlp.w2.weight", &t, {D, I}));
    FRT2_TRY(up_w(t->data.data(), D, I, &L.w_down, &L.w_down_rm));
    FRT2_TRY(need(p + "sa_norm.scale", &t, {D}));
    FRT2_TRY(up32(t->data.data(), D, &L.g_sa));
    FRT2_TRY(need(p + "mlp_norm.scale", &t, {D}));
    FRT2_TRY(up32(t->data.data(), D, &L.g_mlp));
    FRT2_TRY(dev_alloc(&L.kc, static_cast<size_t>(mb) * n * Hk * hd));
    FRT2_TRY(dev_alloc(&L.vc, static_cast<size_t>(mb) * n * Hk * hd));
  }
  {  // rotary tables in double (Qwen2RotaryPositionalEmbeddings: theta_i = base^(-2i/hd), angle = pos * theta_i)
    const int half = hd / 2;
    std::vector<float> c(static_cast<size_t>(n) * half), s(c.size());
    for (int pos = 0; pos < n; ++pos)
      for (int i = 0; i < half; ++i) {
        const double ang = pos * std::pow(static_cast<double>(cfg.rope_base), -2.0 * i / hd);
        c[static_cast<size_t>(pos) * half + i] = static_cast<float>(std::cos(ang));
        s[static_cast<size_t>(pos) * half + i] = static_cast<float>(std::sin(ang));
      }
    FRT2_TRY(up32(c.data(), c.size(), &rope_cos));
    FRT2_TRY(up32(s.data(), s.size(), &rope_sin));
  }
  FRT2_TRY(dev_alloc(&in16, static_cast<size_t>(act_rows) * Db));
  FRT2_TRY(dev_alloc(&attn16, static_cast<size_t>(act_rows) * H * hd));
  FRT2_TRY(dev_alloc(&h16, static_cast<size_t>(act_rows) * I));
  FRT2_TRY(dev_alloc(&x32, static_cast<size_t>(act_rows) * D));
  FRT2_TRY(dev_alloc(&qkv32, static_cast<size_t>(act_rows) * qkv));
  FRT2_TRY(dev_alloc(&logits, static_cast<size_t>(mb) * n * V));
  FRT2_TRY(dev_alloc(&codes, static_cast<size_t>(mb) * n));
  FRT2_TRY(dev_alloc(&given, static_cast<size_t>(mb) * n));
  if (big) {
    FRT2_TRY(dev_alloc(&n16, static_cast<size_t>(act_rows) * std::max(D, Db)));
    FRT2_TRY(dev_alloc(&gu16, static_cast<size_t>(act_rows) * 2 * I));
    // partial sums: the largest (split x columns) any of the layers uses at any batch size
    size_t part = 0;
    const int nk[6][2] = {{D, I}, {D, H * hd}, {qkv, D}, {D, Db}, {V, D}, {V, Db}};
    for (const auto& e : nk) part = std::max(part, static_cast<size_t>(pick_split(1, e[0], e[1])) * e[0]);
    FRT2_TRY(dev_alloc(&part32, part * act_rows));
  }
  FRT2_TRY(dev_alloc(&params, 1));
  FRT2_TRY(dev_alloc(&err_word, 1));
  FRT2_CUDA_OK(cudaFuncSetAttribute(fd_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * V * 4));
  int lo = 0, hi = 0;
  FRT2_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  FRT2_CUDA_OK(cudaStreamCreateWithPriority(&cap_stream, cudaStreamNonBlocking, hi));
  FRT2_CUDA_OK(cudaEventCreateWithFlags(&ws_event, cudaEventDisableTiming));
  {  // projection(audio_embeddings) row by row through the kernel the frame itself would have used (16 rows per launch)
    const long long total = static_cast<long long>(n) * V;
    FRT2_TRY(dev_alloc(&proj_table, static_cast<size_t>(total) * D));
    const long long before = launches;
    for (long long r0 = 0; r0 < total; r0 += FD_MAX_BATCH) {
      const int rows = static_cast<int>(std::min<long long>(FD_MAX_BATCH, total - r0));
      FRT2_TRY(skinny(emb16 + r0 * Db, Db, w_proj, D, nullptr, ACT_NONE, nullptr, proj_table + r0 * D, D, nullptr, 0, nullptr,
                      nullptr, rows, nullptr));
    }
    launches = before;
    FRT2_CUDA_OK(cudaDeviceSynchronize());
    owned.erase(std::find(owned.begin(), owned.end(), static_cast<void*>(emb16)));
    cudaFree(emb16);
    emb16 = nullptr;
  }
  raw.clear();
  finalized = true;
  return FRT2_OK;
}

int FrameDecoder::skinny(const __half* A, int K, const __half* W, int N, const float* bias, int act, const float* resid,
                         float* out32, int64_t ld32, __half* out16, int64_t ld16, const float* ln_x,
                         const float* ln_gamma, int B, cudaStream_t st, int64_t lda, int64_t ldx) {
  if (lda == 0) lda = K;     // row pitches of the fp16 / fp32 activation rows: every other row of a two-row-per-item buffer
  if (ldx == 0) ldx = K;
  ++launches;
  if (use_stream) {
    // the activation tile of a launch lives in shared memory: K = 8960 leaves room for 10 rows, so a wider batch runs the
    // layer in two row groups (each streams the weights once)
    const int fit = gemm_stream_max_rows(K);
    --launches;
    for (int r0 = 0; r0 < B; r0 += fit) {
      StreamGemm d{};
      d.Wt = W; d.N = N; d.K = K; d.B = std::min(fit, B - r0); d.lda = lda; d.ldx = ldx; d.gamma = ln_gamma; d.eps = cfg.norm_eps;
      d.A = A != nullptr ? A + r0 * lda : nullptr;
      d.x = ln_x != nullptr ? ln_x + r0 * ldx : nullptr;
      d.bias = bias; d.act = act; d.ld32 = ld32; d.ld16 = ld16;
      d.resid = resid != nullptr ? resid + r0 * ld32 : nullptr;
      d.out32 = out32 != nullptr ? out32 + r0 * ld32 : nullptr;
      d.out16 = out16 != nullptr ? out16 + r0 * ld16 : nullptr;
      ++launches;
      FRT2_TRY(gemm_stream(d, st));
    }
    return FRT2_OK;
  }
  GemmDesc g{};
  g.A = A; g.a_row_pitch = lda; g.a_batch_pitch = 0; g.rows_a = B; g.batches = 1; g.Kc = K; g.ntaps = 1; g.row_shift = 0;
  g.W = W; g.N = N; g.rows_out = B; g.alpha = 1.0f; g.bias = bias; g.act = act; g.resid = resid; g.out32 = out32;
  g.ld32 = ld32; g.out16 = out16; g.ld16 = ld16;
  if (ln_gamma != nullptr) {
    g.ln_x = ln_x; g.ln_ldx = ldx; g.ln_gamma = ln_gamma; g.ln_beta = nullptr; g.ln_eps = cfg.norm_eps; g.ln_rms = 1;
  }
  return gemm_skinny(g, st);
}

// every kernel of one frame for B items on `st` (captured into a graph by graph_for)
int FrameDecoder::enqueue_frame(int B, cudaStream_t st) {
  const int D = cfg.dim, Db = cfg.backbone_dim, I = cfg.intermediate_dim, V = cfg.audio_vocab_size, n = cfg.audio_num_codebooks;
  const int H = cfg.num_heads, Hk = cfg.num_kv_heads;
  const float scale = 1.0f / std::sqrt(static_cast<float>(hd));
  const int64_t ldl = static_cast<int64_t>(n) * V;
  // timing experiments only (results are wrong with any bit set): marginal cost of a kernel class inside the graph
  static const int skip = getenv("FRT2_FD_SKIP") != nullptr ? atoi(getenv("FRT2_FD_SKIP")) : 0;
  // positions 0 and 1 as ONE two-row pass while 2 B rows fit the GEMM's 8 (the reference's first decoder call has these
  // two positions too, llm.py:308-322): one weight stream less per frame.  Rows 2b / 2b+1 = position 0 / 1 of item b.
  const bool pair = pair_layout(B);
  // sampler of codebook s; the projected embedding of its code becomes the residual row of the next position
  auto sample = [&](int s, float* x_next, int x_ld) {
    fd_sample_kernel<<<B, FD_SAMPLE_THREADS, 2 * V * 4, st>>>(logits, s, V, n, params, given, codes, proj_table, D, x_next, x_ld,
                                                                err_word);
    ++launches;
    return cudaGetLastError();
  };
  // the decoder's layers over `rows` residual rows holding `np` new positions per item, the first of them at pos0
  auto layers_pass = [&](int rows, int pos0, int np) -> int {
    for (FdLayer& L : layers) {
      if (!(skip & 8))
        FRT2_TRY(skinny(nullptr, D, L.w_qkv, qkv, L.b_qkv, ACT_NONE, nullptr, qkv32, qkv, nullptr, 0, x32, L.g_sa, rows, st));
      if (!(skip & 1)) {
        fd_attn_kernel<<<dim3(Hk, B), 32 * (H / Hk), 0, st>>>(qkv32, L.kc, L.vc, attn16, rope_cos, rope_sin, H, Hk, hd, n,
                                                               pos0, np, scale);
        FRT2_CUDA_OK(cudaGetLastError());
        ++launches;
      }
      if (!(skip & 16))
        FRT2_TRY(skinny(attn16, H * hd, L.w_o, D, nullptr, ACT_NONE, x32, x32, D, nullptr, 0, nullptr, nullptr, rows, st));
      if (!(skip & 2))
        FRT2_TRY(skinny(nullptr, D, L.w_gu, 2 * I, nullptr, ACT_SWIGLU, nullptr, nullptr, 0, h16, I, x32, L.g_mlp, rows, st));
      if (!(skip & 4))
        FRT2_TRY(skinny(h16, I, L.w_down, D, nullptr, ACT_NONE, x32, x32, D, nullptr, 0, nullptr, nullptr, rows, st));
    }
    return FRT2_OK;
  };
  // llm.py:323-326 for position `pos` (its residual row: x_rows + b * x_ld): final norm (inside the head GEMM),
  // audio_head[pos-1], sampler, the next position's residual row b
  auto head_and_sample = [&](int pos, const float* x_rows, int64_t x_ld) -> int {
    if (skip & 32) return FRT2_OK;
    FRT2_TRY(skinny(nullptr, D, w_heads + head_stride * (pos - 1), V, nullptr, ACT_NONE, nullptr,
                    logits + static_cast<size_t>(pos) * V, ldl, nullptr, 0, x_rows, g_final, B, st, 0, x_ld));
    FRT2_CUDA_OK(sample(pos, x32, D));
    return FRT2_OK;
  };
  // codebook 0 and the projection of last_h (position 0; in16 holds last_h): llm.py:305-306,321
  FRT2_TRY(skinny(in16, Db, w_head0, V, nullptr, ACT_NONE, nullptr, logits, ldl, nullptr, 0, nullptr, nullptr, B, st));
  int pos;
  if (pair) {
    FRT2_TRY(skinny(in16, Db, w_proj, D, nullptr, ACT_NONE, nullptr, x32, 2 * D, nullptr, 0, nullptr, nullptr, B, st));   // rows 2b
    FRT2_CUDA_OK(sample(0, x32 + D, 2 * D));         // c0 -> rows 2b+1
    FRT2_TRY(layers_pass(2 * B, 0, 2));
    FRT2_TRY(head_and_sample(1, x32 + D, 2 * D));
    pos = 2;
  } else {
    FRT2_TRY(skinny(in16, Db, w_proj, D, nullptr, ACT_NONE, nullptr, x32, D, nullptr, 0, nullptr, nullptr, B, st));
    FRT2_TRY(layers_pass(B, 0, 1));
    FRT2_CUDA_OK(sample(0, x32, D));                 // c0 -> position 1's rows, free once position 0 has run
    pos = 1;
  }
  for (; pos < n; ++pos) {
    FRT2_TRY(layers_pass(B, pos, 1));
    FRT2_TRY(head_and_sample(pos, x32, D));
  }
  return FRT2_OK;
}

// ---- frames of 17 .. max_batch items: the same sequence on the tcgen05 GEMM (128-row tiles; every weight is still
//      streamed once per decoder pass, now shared by up to 128 rows per tile) with RMSNorm / SwiGLU as row kernels
int FrameDecoder::tc(const __half* A, int K, const __half* W, int N, const float* bias, const float* resid, float* out32,
                     int64_t ld32, __half* out16, int64_t ld16, int rows, cudaStream_t st, int64_t lda) {
  GemmDesc g{};
  g.A = A; g.a_row_pitch = lda ? lda : K; g.a_batch_pitch = 0; g.rows_a = rows; g.batches = 1; g.Kc = K; g.ntaps = 1; g.row_shift = 0;
  g.W = W; g.N = N; g.rows_out = rows; g.alpha = 1.0f; g.bias = bias; g.act = ACT_NONE; g.resid = resid; g.out32 = out32;
  g.ld32 = ld32; g.out16 = out16; g.ld16 = ld16; g.narrow_tiles = 1;
  ++launches;
  return gemm_tc(g, st);
}

// fp32 result dst = (accumulate ? dst : bias) + A W^T, with the reduction split over CTAs when the layer has too few tiles
int FrameDecoder::tc32(const __half* A, int K, const __half* W, int N, const float* bias, bool accumulate, float* dst,
                       int64_t ld_dst, int rows, cudaStream_t st, int64_t lda) {
  const int S = pick_split(rows, N, K);
  if (S == 1) return tc(A, K, W, N, bias, accumulate ? dst : nullptr, dst, ld_dst, nullptr, 0, rows, st, lda);
  GemmDesc g{};
  const int Kc = K / S;
  g.A = A; g.a_row_pitch = lda ? lda : K; g.a_batch_pitch = Kc; g.rows_a = rows; g.batches = S; g.Kc = Kc; g.ntaps = 1;
  g.W = W; g.w_batch_k = Kc; g.N = N; g.rows_out = rows; g.alpha = 1.0f; g.act = ACT_NONE;
  g.out32 = part32; g.ld32 = N; g.pitch32 = static_cast<int64_t>(rows) * N; g.narrow_tiles = 1;
  ++launches;
  FRT2_TRY(gemm_tc(g, st));
  const long long n4 = static_cast<long long>(rows) * N / 4;
  fd_splitk_reduce_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const float4*>(part32), rows, N / 4, S, reinterpret_cast<const float4*>(bias), accumulate ? 1 : 0, dst,
      ld_dst);
  FRT2_CUDA_OK(cudaGetLastError());
  ++launches;
  return FRT2_OK;
}

int FrameDecoder::enqueue_frame_big(int B, cudaStream_t st) {
  const int D = cfg.dim, Db = cfg.backbone_dim, I = cfg.intermediate_dim, V = cfg.audio_vocab_size, n = cfg.audio_num_codebooks;
  const int H = cfg.num_heads, Hk = cfg.num_kv_heads;
  const float scale = 1.0f / std::sqrt(static_cast<float>(hd));
  const int64_t ldl = static_cast<int64_t>(n) * V;
  const bool pair = pair_layout(B);       // rows 2b / 2b+1 = positions 0 / 1 of item b in the first pass
  auto sample = [&](int s, float* x_next, int x_ld) {
    fd_sample_kernel<<<B, FD_SAMPLE_THREADS, 2 * V * 4, st>>>(logits, s, V, n, params, given, codes, proj_table, D, x_next, x_ld,
                                                                err_word);
    ++launches;
    return cudaGetLastError();
  };
  auto rms = [&](const float* xrows, int64_t ldx, int rows, const float* g) {
    fd_rms_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(xrows, ldx, rows, D, g, cfg.norm_eps, n16);
    ++launches;
    return cudaGetLastError();
  };
  auto layers_pass = [&](int rows, int pos0, int np) -> int {
    for (FdLayer& L : layers) {
      FRT2_CUDA_OK(rms(x32, D, rows, L.g_sa));
      FRT2_TRY(tc32(n16, D, L.w_qkv_rm, qkv, L.b_qkv, false, qkv32, qkv, rows, st));
      fd_attn_kernel<<<dim3(Hk, B), 32 * (H / Hk), 0, st>>>(qkv32, L.kc, L.vc, attn16, rope_cos, rope_sin, H, Hk, hd, n, pos0, np,
                                                             scale);
      FRT2_CUDA_OK(cudaGetLastError());
      ++launches;
      FRT2_TRY(tc32(attn16, H * hd, L.w_o_rm, D, nullptr, true, x32, D, rows, st));
      FRT2_CUDA_OK(rms(x32, D, rows, L.g_mlp));
      FRT2_TRY(tc(n16, D, L.w_gu_rm, 2 * I, nullptr, nullptr, nullptr, 0, gu16, 2 * I, rows, st));
      const long long ne = static_cast<long long>(rows) * I;
      fd_swiglu_rows_kernel<<<static_cast<unsigned>((ne + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half2*>(gu16), ne, h16);
      FRT2_CUDA_OK(cudaGetLastError());
      ++launches;
      FRT2_TRY(tc32(h16, I, L.w_down_rm, D, nullptr, true, x32, D, rows, st));
    }
    return FRT2_OK;
  };
  auto head_and_sample = [&](int pos, const float* xrows, int64_t ldx) -> int {                      // llm.py:323-326
    FRT2_CUDA_OK(rms(xrows, ldx, B, g_final));
    FRT2_TRY(tc32(n16, D, w_heads_rm + static_cast<size_t>(pos - 1) * V * D, V, nullptr, false,
                  logits + static_cast<size_t>(pos) * V, ldl, B, st));
    FRT2_CUDA_OK(sample(pos, x32, D));
    return FRT2_OK;
  };
  FRT2_TRY(tc32(in16, Db, w_head0_rm, V, nullptr, false, logits, ldl, B, st));                        // llm.py:305
  int pos;
  if (pair) {
    FRT2_TRY(tc32(in16, Db, w_proj_rm, D, nullptr, false, x32, 2 * D, B, st));                         // llm.py:321, rows 2b
    FRT2_CUDA_OK(sample(0, x32 + D, 2 * D));                                                           // c0 -> rows 2b+1
    FRT2_TRY(layers_pass(2 * B, 0, 2));
    FRT2_TRY(head_and_sample(1, x32 + D, 2 * D));
    pos = 2;
  } else {
    FRT2_TRY(tc32(in16, Db, w_proj_rm, D, nullptr, false, x32, D, B, st));
    FRT2_TRY(layers_pass(B, 0, 1));
    FRT2_CUDA_OK(sample(0, x32, D));
    pos = 1;
  }
  for (; pos < n; ++pos) {
    FRT2_TRY(layers_pass(B, pos, 1));
    FRT2_TRY(head_and_sample(pos, x32, D));
  }
  return FRT2_OK;
}

int FrameDecoder::graph_for(int B, cudaGraphExec_t* out) {
  auto it = graphs.find(B);
  if (it != graphs.end()) {
    *out = it->second;
    return FRT2_OK;
  }
  const long long before = launches;
  FRT2_CUDA_OK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = B > FD_MAX_BATCH ? enqueue_frame_big(B, cap_stream) : enqueue_frame(B, cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
  if (rc != FRT2_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  FRT2_CUDA_OK(ce);
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  FRT2_CUDA_OK(ie);
  graph_kernels[B] = launches - before;
  launches = before;
  graphs[B] = exec;
  *out = exec;
  return FRT2_OK;
}

}  // namespace frt2

using namespace frt2;

struct frt2_frame_decoder { FrameDecoder d; };

int frt2_fd_create(const frt2_fd_config* cfg, int device, frt2_frame_decoder** out) {
  FRT2_REQUIRE(cfg != nullptr && out != nullptr, FRT2_ERR_BAD_ARG, "frt2_fd_create: null argument");
  FRT2_REQUIRE(cfg->dim > 0 && cfg->dim % 8 == 0 && cfg->dim <= 4096 && cfg->backbone_dim > 0 && cfg->backbone_dim % 8 == 0 &&
                   cfg->intermediate_dim > 0 && cfg->intermediate_dim % 8 == 0,
               FRT2_ERR_BAD_ARG, "frt2_fd_create: widths must be positive multiples of 8, dim <= 4096");
  FRT2_REQUIRE(cfg->num_layers > 0 && cfg->num_heads > 0 && cfg->num_kv_heads > 0 && cfg->dim % cfg->num_heads == 0 &&
                   cfg->num_heads % cfg->num_kv_heads == 0 && cfg->num_heads / cfg->num_kv_heads <= 8,
               FRT2_ERR_BAD_ARG, "frt2_fd_create: heads must divide dim, kv heads must divide heads (<= 8 per group)");
  const int hd = cfg->dim / cfg->num_heads;
  FRT2_REQUIRE(hd % 4 == 0 && hd <= 128, FRT2_ERR_BAD_ARG, "frt2_fd_create: head_dim must be a multiple of 4 and <= 128");
  FRT2_REQUIRE(cfg->audio_vocab_size > 0 && cfg->audio_vocab_size <= 24000 && cfg->audio_num_codebooks >= 2 &&
                   cfg->audio_num_codebooks <= FD_MAX_POS,
               FRT2_ERR_BAD_ARG, "frt2_fd_create: audio_vocab_size in [1, 24000], audio_num_codebooks in [2, 64]");
  FRT2_REQUIRE(cfg->max_batch >= 0 && cfg->max_batch <= 1024, FRT2_ERR_BAD_ARG, "frt2_fd_create: max_batch must be in [0, 1024]");
  FRT2_REQUIRE(cfg->max_batch <= FD_MAX_BATCH ||
                   (cfg->dim % 64 == 0 && cfg->backbone_dim % 64 == 0 && cfg->intermediate_dim % 64 == 0 && hd % 8 == 0),
               FRT2_ERR_BAD_ARG, "frt2_fd_create: max_batch > 16 needs widths that are multiples of 64");
  int ndev = 0;
  FRT2_CUDA_OK(cudaGetDeviceCount(&ndev));
  FRT2_REQUIRE(device >= 0 && device < ndev, FRT2_ERR_BAD_ARG, "frt2_fd_create: bad device index");
  cudaDeviceProp prop;
  FRT2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  FRT2_REQUIRE(prop.major == 10, FRT2_ERR_BAD_ARG, "frt2_fd_create: this library is sm_100a only (no fallback path)");
  auto* f = new frt2_frame_decoder();
  f->d.device = device;
  f->d.cfg = *cfg;
  f->d.hd = hd;
  f->d.qkv = (cfg->num_heads + 2 * cfg->num_kv_heads) * hd;
  *out = f;
  return FRT2_OK;
}

int frt2_fd_load_tensor(frt2_frame_decoder* f, const char* key, const float* data, int ndim, const int64_t* shape,
                        int on_device) {
  FRT2_REQUIRE(f && key && data && shape && ndim >= 0 && ndim <= 4, FRT2_ERR_BAD_ARG, "frt2_fd_load_tensor: bad argument");
  FrameDecoder& d = f->d;
  FRT2_REQUIRE(!d.finalized, FRT2_ERR_BAD_ARG, "frt2_fd_load_tensor: already finalized");
  const std::string k(key);
  const bool mine = k == "projection.weight" || k == "audio_embeddings.weight" || k == "codebook0_head.weight" ||
                    k == "audio_head" || k.rfind("decoder.", 0) == 0;
  if (!mine) return FRT2_OK;   // backbone / text tensors of Model.state_dict(): not part of the frame tail
  HostT t;
  t.shape.assign(shape, shape + ndim);
  int64_t n = 1;
  for (auto v : t.shape) n *= v;
  FRT2_REQUIRE(n >= 0, FRT2_ERR_BAD_ARG, "frt2_fd_load_tensor: negative dimension");
  t.data.resize(n);
  if (on_device) {
    FRT2_CUDA_OK(cudaSetDevice(d.device));
    FRT2_CUDA_OK(cudaMemcpy(t.data.data(), data, n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::memcpy(t.data.data(), data, n * 4);
  }
  std::lock_guard<std::mutex> lk(d.mu);
  d.raw[k] = std::move(t);
  return FRT2_OK;
}

int frt2_fd_finalize(frt2_frame_decoder* f) {
  FRT2_REQUIRE(f, FRT2_ERR_BAD_ARG, "null frame decoder");
  std::lock_guard<std::mutex> lk(f->d.mu);
  FRT2_REQUIRE(!f->d.finalized, FRT2_ERR_BAD_ARG, "already finalized");
  return f->d.finalize();
}

void frt2_fd_destroy(frt2_frame_decoder* f) { delete f; }

int frt2_fd_generate(frt2_frame_decoder* f, const float* last_h, int B, const int32_t* c0, const float* noise,
                     uint64_t seed, int topk, float temperature, const int32_t* forced, int32_t* codes, float* logits,
                     int64_t* launches, void* cuda_stream) {
  FRT2_REQUIRE(f && last_h && codes, FRT2_ERR_BAD_ARG, "frt2_fd_generate: null argument");
  FrameDecoder& d = f->d;
  FRT2_REQUIRE(d.finalized, FRT2_ERR_NOT_FINALIZED, "frt2_fd_generate: call frt2_fd_finalize first");
  FRT2_REQUIRE(B >= 1 && B <= d.mb, FRT2_ERR_BAD_ARG, "frt2_fd_generate: batch must be in [1, max(16, max_batch)]");
  FRT2_REQUIRE(topk >= 1 && temperature > 0.f, FRT2_ERR_BAD_ARG, "frt2_fd_generate: topk >= 1 and temperature > 0 required");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::lock_guard<std::mutex> lk(d.mu);
  FRT2_CUDA_OK(cudaSetDevice(d.device));
  cudaGraphExec_t exec = nullptr;
  FRT2_TRY(d.graph_for(B, &exec));
  // calls are asynchronous and share the activations: a call on another CUDA stream waits for the previous one on the device
  if (d.ws_used && st != d.ws_last) FRT2_CUDA_OK(cudaStreamWaitEvent(st, d.ws_event, 0));
  const int n = d.cfg.audio_num_codebooks, V = d.cfg.audio_vocab_size, Db = d.cfg.backbone_dim;
  FdParams p{};
  p.noise = noise; p.seed = seed; p.frame = d.frame++; p.topk = topk; p.temperature = temperature;
  p.has_c0 = c0 != nullptr; p.has_forced = forced != nullptr;
  fd_begin_kernel<<<std::max(1, (B * Db + 255) / 256), 256, 0, st>>>(last_h, B, Db, d.in16, c0, forced, n, d.given, p, d.params);
  FRT2_CUDA_OK(cudaGetLastError());
  FRT2_CUDA_OK(cudaGraphLaunch(exec, st));
  d.launches += 1 + d.graph_kernels[B];
  FRT2_CUDA_OK(cudaMemcpyAsync(codes, d.codes, static_cast<size_t>(B) * n * 4, cudaMemcpyDeviceToDevice, st));
  if (logits != nullptr)
    FRT2_CUDA_OK(cudaMemcpyAsync(logits, d.logits, static_cast<size_t>(B) * n * V * 4, cudaMemcpyDeviceToDevice, st));
  FRT2_CUDA_OK(cudaEventRecord(d.ws_event, st));
  d.ws_last = st;
  d.ws_used = true;
  if (launches != nullptr) *launches = 1 + d.graph_kernels[B];
  return FRT2_OK;
}

int frt2_fd_check_error(frt2_frame_decoder* f, void* cuda_stream) {
  FRT2_REQUIRE(f, FRT2_ERR_BAD_ARG, "null frame decoder");
  FrameDecoder& d = f->d;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::lock_guard<std::mutex> lk(d.mu);
  FRT2_CUDA_OK(cudaSetDevice(d.device));
  unsigned int word = 0;
  FRT2_CUDA_OK(cudaMemcpyAsync(&word, d.err_word, 4, cudaMemcpyDeviceToHost, st));
  FRT2_CUDA_OK(cudaStreamSynchronize(st));
  if (word != 0) {
    FRT2_CUDA_OK(cudaMemsetAsync(d.err_word, 0, 4, st));
    if (word & DEV_ERR_INDEX_OOR) {
      set_error("index out of range in self (a given codebook code is outside [0, audio_vocab_size))");
      return FRT2_ERR_INDEX_OUT_OF_RANGE;
    }
  }
  return FRT2_OK;
}

// single-operator parity hook: sample_topk + _multinomial_sample_one_no_sync (llm.py:34-49) on given logits
int frt2_op_sample_topk(const float* logits, int B, int V, int topk, float temperature, const float* noise, uint64_t seed,
                        int32_t* codes, void* cuda_stream) {
  FRT2_REQUIRE(logits && codes && B >= 1 && V >= 1 && V <= 24000 && topk >= 1 && temperature > 0.f, FRT2_ERR_BAD_ARG,
               "frt2_op_sample_topk: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  FdParams* dp = nullptr;
  FRT2_CUDA_OK(cudaMalloc(&dp, sizeof(FdParams)));
  FdParams p{};
  p.noise = noise; p.seed = seed; p.frame = 0; p.topk = topk; p.temperature = temperature;
  cudaError_t e = cudaMemcpyAsync(dp, &p, sizeof(p), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * V * 4);
  if (e == cudaSuccess) {
    // ncb = 1, s = 0: item b reads logits[b * V ..], noise[b * V ..] and writes codes[b]; no embedding lookup
    fd_sample_kernel<<<B, FD_SAMPLE_THREADS, 2 * V * 4, st>>>(logits, 0, V, 1, dp, nullptr, codes, nullptr, 0, nullptr, 0, nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(dp);
  FRT2_CUDA_OK(e);
  return FRT2_OK;
}
