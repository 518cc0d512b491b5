// Bandwidth-bound kernels of the decode path: K1 RVQ gather-and-sum, K4 LayerNorm(+SiLU) rows,
// K5c overlap-add / envelope / trim, plus the SIMT check GEMM used by the unit tests.
#include <algorithm>

#include "common.cuh"
#include "misc_bodies.cuh"

namespace frt2 {

// =====================================================================================================
// K1 — RVQ dequantisation: fused gather-and-sum over all codebooks (reference rvq.py:56-60,145-164).
// One CTA handles TOK tokens; the nq indices of each token are staged in shared memory after an
// always-on range check; every thread owns float4 columns of the D-wide row, so each codebook row is read
// with fully coalesced 16-byte loads.  The sum runs in index order from +0.0f (bit-exact with the reference
// for Identity projections).  `tables` is either the raw codebooks (D = codebook_dim) or the tables
// pre-folded with the weight-normed out_project (D = rvq_dim), see engine.cu.
// =====================================================================================================
constexpr int RVQ_TOK = 4;
constexpr int RVQ_MAX_NQ = 64;

template <typename IdxT>
__global__ void __launch_bounds__(128) rvq_gather_sum_kernel(const IdxT* __restrict__ tokens, long long sB, long long sQ,
                                                             long long sL, int B, int nq, int L,
                                                             const float* __restrict__ tables, int K, int D,
                                                             float* __restrict__ sum32, __half* __restrict__ sum16,
                                                             float* __restrict__ rows, unsigned int* err_word) {
  __shared__ int s_idx[RVQ_TOK][RVQ_MAX_NQ];
  pdl_trigger();   // a following PDL-launched kernel (streaming skinny GEMM) may start its weight prefetch
  const long long R = static_cast<long long>(B) * L;
  const long long r0 = static_cast<long long>(blockIdx.x) * RVQ_TOK;
  for (int e = threadIdx.x; e < RVQ_TOK * nq; e += blockDim.x) {
    const int t = e / nq, i = e - t * nq;
    const long long r = r0 + t;
    int idx = 0;
    if (r < R) {
      const long long b = r / L, l = r - b * L;
      const long long raw = static_cast<long long>(tokens[b * sB + i * sQ + l * sL]);
      if (raw < 0 || raw >= K) {
        atomicOr(err_word, DEV_ERR_INDEX_OOR);  // surfaced as FRT2_ERR_INDEX_OUT_OF_RANGE / IndexError
      } else {
        idx = static_cast<int>(raw);
      }
    }
    s_idx[t][i] = idx;
  }
  __syncthreads();
  const int D4 = D >> 2;
  const float4* tab4 = reinterpret_cast<const float4*>(tables);
  for (int t = 0; t < RVQ_TOK; ++t) {
    const long long r = r0 + t;
    if (r >= R) break;
    for (int c = threadIdx.x; c < D4; c += blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < nq; ++i) {
        const float4 v = __ldg(tab4 + (static_cast<long long>(i) * K + s_idx[t][i]) * D4 + c);
        if (rows != nullptr) reinterpret_cast<float4*>(rows)[(r * nq + i) * D4 + c] = v;
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (sum32 != nullptr) reinterpret_cast<float4*>(sum32)[r * D4 + c] = acc;
      if (sum16 != nullptr) {
        uint2 h;
        h.x = pack_half2(acc.x, acc.y);
        h.y = pack_half2(acc.z, acc.w);
        reinterpret_cast<uint2*>(sum16)[r * D4 + c] = h;
      }
    }
  }
}

int rvq_gather_sum(const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B, int nq, int L,
                   const float* tables, int K, int D, float* sum32, __half* sum16, float* rows,
                   unsigned int* err_word, cudaStream_t stream) {
  FRT2_REQUIRE(idx_bytes == 4 || idx_bytes == 8, FRT2_ERR_BAD_DTYPE, "tokens must be int32 or int64");
  FRT2_REQUIRE(nq >= 1 && nq <= RVQ_MAX_NQ, FRT2_ERR_BAD_ARG, "nq out of range");
  FRT2_REQUIRE(D % 4 == 0, FRT2_ERR_BAD_ARG, "codebook row width must be a multiple of 4");
  const long long R = static_cast<long long>(B) * L;
  if (R == 0) return FRT2_OK;
  const unsigned grid = static_cast<unsigned>((R + RVQ_TOK - 1) / RVQ_TOK);
  if (idx_bytes == 4) {
    rvq_gather_sum_kernel<int><<<grid, 128, 0, stream>>>(static_cast<const int*>(tokens), sB, sQ, sL, B, nq, L, tables,
                                                         K, D, sum32, sum16, rows, err_word);
  } else {
    rvq_gather_sum_kernel<long long><<<grid, 128, 0, stream>>>(static_cast<const long long*>(tokens), sB, sQ, sL, B,
                                                               nq, L, tables, K, D, sum32, sum16, rows, err_word);
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// =====================================================================================================
// K4 — LayerNorm over channels (+ optional SiLU), one warp per row, warp-shuffle reductions, fp32 in,
// fp16 out (the next op is always a tensor-core GEMM / conv).  Reference: nn.LayerNorm(eps 1e-5 / 1e-6)
// decoder.py:119,127,246; whisper.py:134,140; followed by nn.SiLU in the resnet blocks (decoder.py:121,129).
// Rows are addressed as (batch, t): in = x + (b*in_rows_per_batch + t)*ldx, out = out16 + b*out_batch_pitch + t*ld16
// so the streaming path can write straight behind the conv history rows.
// =====================================================================================================
template <int RPW>
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, long long ldx, long long rows,
                                                         int rows_per_batch, int C, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, int apply_silu,
                                                         __half* __restrict__ out16, long long ld16,
                                                         long long out_batch_pitch, float* __restrict__ mean_out) {
  pdl_trigger();
  layer_norm_body<RPW>(x, ldx, rows, rows_per_batch, C, gamma, beta, eps, apply_silu, out16, ld16, out_batch_pitch,
                       (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW, mean_out);
}

int layer_norm_rows_batched(const float* x, int64_t ldx, int64_t rows, int rows_per_batch, int C, const float* gamma,
                            const float* beta, float eps, int apply_silu, __half* out16, int64_t ld16,
                            int64_t out_batch_pitch, cudaStream_t stream, float* mean_out) {
  FRT2_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ld16 % 4 == 0 && out_batch_pitch % 4 == 0, FRT2_ERR_BAD_ARG,
               "layer_norm: C and pitches must be multiples of 4");
  if (rows == 0) return FRT2_OK;
  const int warps = 8;
  if (rows >= 4096) {
    const unsigned grid = static_cast<unsigned>((rows + 2 * warps - 1) / (2 * warps));
    layer_norm_kernel<2><<<grid, warps * 32, 0, stream>>>(x, ldx, rows, rows_per_batch, C, gamma, beta, eps, apply_silu,
                                                          out16, ld16, out_batch_pitch, mean_out);
  } else {
    const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
    layer_norm_kernel<1><<<grid, warps * 32, 0, stream>>>(x, ldx, rows, rows_per_batch, C, gamma, beta, eps, apply_silu,
                                                          out16, ld16, out_batch_pitch, mean_out);
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// =====================================================================================================
// K4b — row statistics of the fp16 residual copy (folded LayerNorm): one warp per row, the row lives in registers
// (C <= 2048: 8 x 16 B per lane), mean first, then the centred sum of squares.  2 B per element in, 8 B per row out.
// =====================================================================================================
__global__ void __launch_bounds__(256) row_stats_kernel(const __half* __restrict__ x, long long ld, long long rows, int C,
                                                        float eps, float2* __restrict__ stats,
                                                        float* __restrict__ shift_io) {
  // two rows per warp, all loads of both rows in flight before the first use; ONE shuffle round per row pair:
  // sums of (x - shift) and (x - shift)^2 with shift = the row's first element (no cancellation for rows with a
  // large common offset), mean = shift + s1/C, var = s2/C - (s1/C)^2
  const long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  const int C8 = C >> 3;
  uint4 cache[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + min(row0 + r, rows - 1) * ld);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < C8) cache[r][i] = __ldcs(xr + c);
    }
  }
  float s1[2], s2[2], shift[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float first = __half2float(*reinterpret_cast<const __half*>(&cache[r][0]));   // lane 0 holds element 0
    shift[r] = __shfl_sync(0xffffffffu, first, 0);
    s1[r] = s2[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane + 32 * i < C8) {
        const __half2* h = reinterpret_cast<const __half2*>(&cache[r][i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          const float a = f.x - shift[r], b = f.y - shift[r];
          s1[r] += a + b;
          s2[r] = fmaf(a, a, fmaf(b, b, s2[r]));
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
      s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
    }
  }
  if (lane < 2 && row0 + lane < rows) {
    const float inv_c = 1.0f / static_cast<float>(C);
    const float m = (lane == 0 ? s1[0] : s1[1]) * inv_c;
    const float var = fmaxf((lane == 0 ? s2[0] : s2[1]) * inv_c - m * m, 0.f);
    const float mean = (lane == 0 ? shift[0] : shift[1]) + m;
    stats[row0 + lane] = make_float2(mean, rsqrtf(var + eps));
    // the rows are the residual stream minus shift_io[row] (GemmDesc::x16_shift): keep the offsets on the row means
    if (shift_io != nullptr) shift_io[row0 + lane] += mean;
  }
}

int row_stats(const __half* x16, int64_t ld, int64_t rows, int C, float eps, float2* stats, cudaStream_t stream,
              float* shift_io) {
  FRT2_REQUIRE(C % 8 == 0 && C <= 2048 && ld % 8 == 0, FRT2_ERR_BAD_ARG, "row_stats: C must be a multiple of 8, <= 2048");
  if (rows == 0) return FRT2_OK;
  row_stats_kernel<<<static_cast<unsigned>((rows + 15) / 16), 256, 0, stream>>>(x16, ld, rows, C, eps, stats, shift_io);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

int layer_norm_rows(const float* x, int64_t ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                    int apply_silu, __half* out16, int64_t ld16, cudaStream_t stream) {
  return layer_norm_rows_batched(x, ldx, rows, rows > 0 ? rows : 1, C, gamma, beta, eps, apply_silu, out16, ld16, 0,
                                 stream);
}

// =====================================================================================================
// K5c — overlap-add of windowed frames, window-square envelope normalisation and "same" trimming
// (reference ISTFT.forward decoder.py:384-405 and ISTFT.forward_chunk decoder.py:431-467) in gather form:
// every output sample sums the <= n_fft/hop frames that cover it, so there is no scatter / atomics, the
// envelope is accumulated in the same pass (the reference rebuilds it with a second F.fold every call) and
// the trim is index arithmetic.  Frame list = [carried tail frames (streaming, not first) | this call's frames].
// =====================================================================================================
__global__ void __launch_bounds__(256) overlap_add_kernel(OlaDesc d, int ntail, int start, int n_out_full) {
  overlap_add_sample(d, ntail, start, n_out_full, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

// Four consecutive samples per thread (hop, start, pitches and offsets multiples of 4: always true for the codec's
// hop 240 / n_fft 960): the <= n_fft/hop frames covering them are read with 16-byte loads and the samples leave as ONE
// 16-byte (fp32) or 8-byte (int16) store — a quarter of the memory instructions of the scalar kernel, and full 16-byte
// packets when the destination is a peer GPU's buffer (frt2_decode_scatter over NVLink).
__global__ void __launch_bounds__(256) overlap_add_vec4_kernel(OlaDesc d, int ntail, int start, int n_out_full) {
  const int b = blockIdx.y;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (n >= n_out_full) return;
  int TF = ntail + d.T;
  int n_out = n_out_full;
  if (d.lengths != nullptr) {
    TF = min(TF, d.lengths[b] * d.len_mul);
    n_out = TF * d.hop;
  }
  const long long obase = (d.out_off != nullptr ? d.out_off[b] : static_cast<long long>(b) * d.audio_pitch);
  const long long oidx = obase + n;
  float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n >= n_out) {
    if (d.out_off != nullptr) return;
  } else {
    const int m = n + start;
    int t_hi = m / d.hop;
    if (t_hi > TF - 1) t_hi = TF - 1;
    int t_lo = (m - d.n_fft + d.hop) / d.hop;
    if (m - d.n_fft + 1 <= 0) t_lo = 0;
    float4 env = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = t_lo; t <= t_hi; ++t) {
      const int off = m - t * d.hop;
      const float* fr = (t < ntail)
                            ? d.tail + (static_cast<long long>(b) * 3 + t) * d.n_fft
                            : d.frames + static_cast<long long>(b) * d.frames_batch_pitch +
                                  static_cast<long long>(t - ntail) * d.n_fft;
      const float4 w = __ldg(reinterpret_cast<const float4*>(d.window + off));
      const float4 f = __ldcs(reinterpret_cast<const float4*>(fr + off));
      y.x += f.x; y.y += f.y; y.z += f.z; y.w += f.w;
      env.x += w.x * w.x; env.y += w.y * w.y; env.z += w.z * w.z; env.w += w.w * w.w;
    }
    y.x /= env.x; y.y /= env.y; y.z /= env.z; y.w /= env.w;
  }
  if (d.pcm16 != nullptr) {
    auto q = [](float v) {
      return static_cast<uint32_t>(static_cast<uint16_t>(static_cast<int16_t>(
          __float2int_rz(fminf(fmaxf(v * 32767.0f, -32768.0f), 32767.0f)))));
    };
    uint2 o;
    o.x = q(y.x) | (q(y.y) << 16);
    o.y = q(y.z) | (q(y.w) << 16);
    if ((oidx & 3) == 0) {
      *reinterpret_cast<uint2*>(d.pcm16 + oidx) = o;
    } else {   // scatter offset not a multiple of 4 samples
      int16_t* op = d.pcm16 + oidx;
      op[0] = static_cast<int16_t>(o.x & 0xffff); op[1] = static_cast<int16_t>(o.x >> 16);
      op[2] = static_cast<int16_t>(o.y & 0xffff); op[3] = static_cast<int16_t>(o.y >> 16);
    }
  } else if ((oidx & 3) == 0) {
    *reinterpret_cast<float4*>(d.audio + oidx) = y;
  } else {
    float* op = d.audio + oidx;
    op[0] = y.x; op[1] = y.y; op[2] = y.z; op[3] = y.w;
  }
}

int istft_overlap_add(const OlaDesc& d, cudaStream_t stream) {
  FRT2_REQUIRE(d.n_fft % d.hop == 0 && d.T >= 1 && d.B >= 1, FRT2_ERR_BAD_ARG, "overlap_add: bad shape");
  const int pad = (d.n_fft - d.hop) / 2;
  const int ntail = (d.tail != nullptr && !d.first) ? (d.n_fft / d.hop - 1) : 0;
  FRT2_REQUIRE(ntail == 0 || ntail == 3, FRT2_ERR_BAD_ARG, "overlap_add: streaming needs n_fft == 4*hop");
  const int TF = ntail + d.T;
  const int full = (TF - 1) * d.hop + d.n_fft;
  const int start = d.first ? pad : (d.n_fft - d.hop);
  int n_out = full - start - (d.last ? pad : (d.n_fft - d.hop));
  if (d.ctrl != nullptr) n_out = d.T * d.hop + pad;  // upper bound over first/last combinations
  if (n_out <= 0) return FRT2_OK;
  const bool vec4 = d.ctrl == nullptr && d.hop % 4 == 0 && start % 4 == 0 && pad % 4 == 0 && d.audio_pitch % 4 == 0 &&
                    d.frames_batch_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(d.frames) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d.window) & 15) == 0 &&
                    (d.tail == nullptr || (reinterpret_cast<uintptr_t>(d.tail) & 15) == 0) &&
                    (d.pcm16 != nullptr ? (reinterpret_cast<uintptr_t>(d.pcm16) & 7) == 0
                                        : (reinterpret_cast<uintptr_t>(d.audio) & 15) == 0);
  if (vec4) {
    dim3 grid((n_out / 4 + 255) / 256, d.B);
    overlap_add_vec4_kernel<<<grid, 256, 0, stream>>>(d, ntail, start, n_out);
    FRT2_CUDA_OK(cudaGetLastError());
    return FRT2_OK;
  }
  dim3 grid((n_out + 255) / 256, d.B);
  overlap_add_kernel<<<grid, 256, 0, stream>>>(d, ntail, start, n_out);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// =====================================================================================================
// K6 — rational-ratio sinc resampler (torchaudio.functional.resample as the reference calls it: fireredtts2.py:65 for
// the prompt, fireredtts2.py:389-391 for every generated turn, 24 kHz -> 16 kHz): polyphase FIR
//   y[b, new*m + p] = sum_k xpad[b, orig*m + k] * taps[p][k],   xpad = x zero-padded by (width, width + orig)
// A block stages the input span of `fr` output frames in shared memory (coalesced, zero fill outside the item's
// samples = the padding), every thread then produces outputs with K sequential fp32 FMAs.  Bandwidth-bound:
// 4 B in + 4*new/orig B out per input sample.
// =====================================================================================================
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long x_pitch, long long n_in,
                                                       const int* __restrict__ lengths, const float* __restrict__ taps,
                                                       int K, int width, int orig, int nnew, int fr,
                                                       float* __restrict__ y, long long y_pitch, long long n_out_max) {
  extern __shared__ float s_in[];
  const int b = blockIdx.y;
  const long long n = (lengths != nullptr) ? min(static_cast<long long>(lengths[b]), n_in) : n_in;
  const long long n_out = (n * nnew + orig - 1) / orig;              // ceil(new * n / orig)
  const long long m0 = static_cast<long long>(blockIdx.x) * fr;      // first output frame of this block
  const long long in0 = m0 * orig - width;                           // x index of s_in[0]
  const int span = (fr - 1) * orig + K;
  const float* xb = x + b * x_pitch;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const long long g = in0 + i;
    s_in[i] = (g >= 0 && g < n) ? __ldg(xb + g) : 0.f;
  }
  __syncthreads();
  float* yb = y + b * y_pitch;
  const int outs = fr * nnew;
  for (int o = threadIdx.x; o < outs; o += blockDim.x) {
    const int m = o / nnew, p = o - m * nnew;
    const long long j = (m0 + m) * nnew + p;
    if (j >= n_out_max) continue;
    float acc = 0.f;
    if (j < n_out) {
      const float* xi = s_in + m * orig;
      const float* tp = taps + static_cast<long long>(p) * K;
      for (int k = 0; k < K; ++k) acc = fmaf(xi[k], __ldg(tp + k), acc);
    }
    yb[j] = acc;   // zeros past the item's own length (ragged batches)
  }
}

// few output phases per frame (24 kHz -> 16 kHz: 2; 16 -> 24: 3): one thread produces ALL phases of a frame, so every
// staged input sample is read from shared memory once per frame instead of once per output, and the taps come from a
// [k][phase] shared-memory copy as one broadcast read per k.  Same sequential-k fp32 accumulation as the generic kernel.
template <int NP>
__global__ void __launch_bounds__(256) resample_np_kernel(const float* __restrict__ x, long long x_pitch, long long n_in,
                                                          const int* __restrict__ lengths,
                                                          const float* __restrict__ taps, int K, int width, int orig,
                                                          int fr, float* __restrict__ y, long long y_pitch,
                                                          long long n_out_max) {
  extern __shared__ float s_in[];
  const int span = (fr - 1) * orig + K;
  float* s_taps = s_in + ((span + 3) & ~3);      // [K][NP]
  const int b = blockIdx.y;
  const long long n = (lengths != nullptr) ? min(static_cast<long long>(lengths[b]), n_in) : n_in;
  const long long n_out = (n * NP + orig - 1) / orig;
  const long long m0 = static_cast<long long>(blockIdx.x) * fr;
  const long long in0 = m0 * orig - width;
  const float* xb = x + b * x_pitch;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const long long g = in0 + i;
    s_in[i] = (g >= 0 && g < n) ? __ldg(xb + g) : 0.f;
  }
  for (int i = threadIdx.x; i < K * NP; i += blockDim.x) {
    const int k = i / NP, p = i - k * NP;
    s_taps[i] = __ldg(taps + static_cast<long long>(p) * K + k);
  }
  __syncthreads();
  float* yb = y + b * y_pitch;
  for (int m = threadIdx.x; m < fr; m += blockDim.x) {
    const long long j0 = (m0 + m) * NP;
    if (j0 >= n_out_max) break;
    float acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = 0.f;
    const float* xi = s_in + m * orig;
    for (int k = 0; k < K; ++k) {
      const float xv = xi[k];
#pragma unroll
      for (int p = 0; p < NP; ++p) acc[p] = fmaf(xv, s_taps[k * NP + p], acc[p]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
      if (j0 + p < n_out_max) yb[j0 + p] = (j0 + p < n_out) ? acc[p] : 0.f;
  }
}

int resample_rows(const float* x, int64_t x_pitch, int B, int64_t n_in, const int* lengths, const float* taps, int K,
                  int width, int orig, int nnew, float* y, int64_t y_pitch, cudaStream_t stream) {
  if (B <= 0 || n_in <= 0) return FRT2_OK;
  const int64_t n_out_max = (n_in * nnew + orig - 1) / orig;
  int fr = (8192 - K) / orig;                    // <= 32 KB of staged input per block
  fr = std::max(1, std::min(fr, 256));
  const int span = (fr - 1) * orig + K;
  const int64_t frames = (n_out_max + nnew - 1) / nnew;
  dim3 grid(static_cast<unsigned>((frames + fr - 1) / fr), B);
  const size_t smem_np = (static_cast<size_t>((span + 3) & ~3) + static_cast<size_t>(K) * nnew) * 4;
  if (nnew == 2) {
    resample_np_kernel<2><<<grid, 256, smem_np, stream>>>(x, x_pitch, n_in, lengths, taps, K, width, orig, fr, y,
                                                          y_pitch, n_out_max);
  } else if (nnew == 3) {
    resample_np_kernel<3><<<grid, 256, smem_np, stream>>>(x, x_pitch, n_in, lengths, taps, K, width, orig, fr, y,
                                                          y_pitch, n_out_max);
  } else if (nnew == 1) {
    resample_np_kernel<1><<<grid, 256, smem_np, stream>>>(x, x_pitch, n_in, lengths, taps, K, width, orig, fr, y,
                                                          y_pitch, n_out_max);
  } else {
    resample_kernel<<<grid, 256, static_cast<size_t>(span) * 4, stream>>>(x, x_pitch, n_in, lengths, taps, K, width,
                                                                          orig, nnew, fr, y, y_pitch, n_out_max);
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// K5c + K6 in one pass ("resample fused onto the decoder output"): the block that stages the input span of
// its output frames does not READ the 24 kHz waveform — it computes those samples from the windowed iSTFT frames (the
// gather-form overlap-add of overlap_add_vec4_kernel, same order of additions: bit-identical samples), optionally stores
// its own part of them as the 24 kHz output, and resamples from shared memory with the FMA sequence of
// resample_np_kernel (bit-identical to decode -> frt2_resample).  The waveform never makes an HBM round trip between the
// two steps; every frame element is read once (plus the tile halo of 2 * width samples per fr * orig).
// Offline decode only (first = last = 1, no carried tail).
template <int NP>
__global__ void __launch_bounds__(256) ola_resample_np_kernel(OlaDesc d, int start, const float* __restrict__ taps, int K,
                                                              int width, int orig, int fr, float* __restrict__ y,
                                                              long long y_pitch, long long n_out_max) {
  extern __shared__ float s_in[];
  const int span = (fr - 1) * orig + K;
  float* s_taps = s_in + ((span + 3) & ~3);      // [K][NP]
  const int b = blockIdx.y;
  int TF = d.T;
  if (d.lengths != nullptr) TF = min(TF, d.lengths[b] * d.len_mul);
  const long long n_full = static_cast<long long>(d.T) * d.hop;     // samples of the longest item (row width)
  const long long n = static_cast<long long>(TF) * d.hop;           // this item's samples
  const long long n_out = (n * NP + orig - 1) / orig;
  const long long m0 = static_cast<long long>(blockIdx.x) * fr;
  const long long in0 = m0 * orig - width;
  const float* fb = d.frames + static_cast<long long>(b) * d.frames_batch_pitch;
  // four consecutive samples per thread, aligned to a multiple of 4 of the sample index (hop, pad and n_fft are
  // multiples of 4: the frames / window are read with 16-byte loads exactly as in overlap_add_vec4_kernel)
  const long long g_first = in0 >= 0 ? (in0 & ~3LL) : -((-in0 + 3) & ~3LL);
  const int quads = static_cast<int>((in0 + span - g_first + 3) >> 2);
  for (int qd = threadIdx.x; qd < quads; qd += blockDim.x) {
    const long long g = g_first + 4LL * qd;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g >= 0 && g < n) {       // n is a multiple of 4: the quad is inside the item or outside it as a whole
      const int m = static_cast<int>(g) + start;
      int t_hi = m / d.hop;
      if (t_hi > TF - 1) t_hi = TF - 1;
      int t_lo = (m - d.n_fft + d.hop) / d.hop;
      if (m - d.n_fft + 1 <= 0) t_lo = 0;
      float4 env = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = t_lo; t <= t_hi; ++t) {
        const int off = m - t * d.hop;
        const float4 w = __ldg(reinterpret_cast<const float4*>(d.window + off));
        const float4 f = __ldcs(reinterpret_cast<const float4*>(fb + static_cast<long long>(t) * d.n_fft + off));
        v.x += f.x; v.y += f.y; v.z += f.z; v.w += f.w;
        env.x += w.x * w.x; env.y += w.y * w.y; env.z += w.z * w.z; env.w += w.w * w.w;
      }
      v.x /= env.x; v.y /= env.y; v.z /= env.z; v.w /= env.w;
    }
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = g + e - in0;
      if (i >= 0 && i < span) s_in[i] = vv[e];
    }
    // this block owns the 24 kHz samples [m0 * orig, (m0 + fr) * orig): zeros behind the item's own length
    if (d.audio != nullptr) {
      const long long own0 = m0 * orig, own1 = own0 + static_cast<long long>(fr) * orig;
      float* ap = d.audio + static_cast<long long>(b) * d.audio_pitch;
      if (g >= own0 && g + 3 < own1 && g + 3 < n_full) {
        *reinterpret_cast<float4*>(ap + g) = v;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (g + e >= own0 && g + e < own1 && g + e >= 0 && g + e < n_full) ap[g + e] = vv[e];
      }
    }
  }
  for (int i = threadIdx.x; i < K * NP; i += blockDim.x) {
    const int k = i / NP, p = i - k * NP;
    s_taps[i] = __ldg(taps + static_cast<long long>(p) * K + k);
  }
  __syncthreads();
  float* yb = y + b * y_pitch;
  for (int m = threadIdx.x; m < fr; m += blockDim.x) {
    const long long j0 = (m0 + m) * NP;
    if (j0 >= n_out_max) break;
    float acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = 0.f;
    const float* xi = s_in + m * orig;
    for (int k = 0; k < K; ++k) {
      const float xv = xi[k];
#pragma unroll
      for (int p = 0; p < NP; ++p) acc[p] = fmaf(xv, s_taps[k * NP + p], acc[p]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
      if (j0 + p < n_out_max) yb[j0 + p] = (j0 + p < n_out) ? acc[p] : 0.f;
  }
}

int istft_overlap_add_resample(const OlaDesc& d, const float* taps, int K, int width, int orig, int nnew, float* y,
                               int64_t y_pitch, cudaStream_t stream) {
  FRT2_REQUIRE(d.n_fft % d.hop == 0 && d.T >= 1 && d.B >= 1 && d.first && d.last && d.tail == nullptr && d.ctrl == nullptr &&
                   d.out_off == nullptr && d.pcm16 == nullptr,
               FRT2_ERR_BAD_ARG, "overlap_add_resample: offline fp32 decode only");
  FRT2_REQUIRE(nnew >= 1 && nnew <= 3, FRT2_ERR_BAD_ARG,
               "overlap_add_resample: new_freq / gcd must be 1, 2 or 3 (24 kHz -> 16 / 8 / 12 kHz)");
  FRT2_REQUIRE(d.hop % 4 == 0 && d.n_fft % 4 == 0 && ((d.n_fft - d.hop) / 2) % 4 == 0 && d.frames_batch_pitch % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(d.frames) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.window) & 15) == 0 &&
                   (d.audio == nullptr || ((reinterpret_cast<uintptr_t>(d.audio) & 15) == 0 && d.audio_pitch % 4 == 0)),
               FRT2_ERR_BAD_ARG, "overlap_add_resample: hop / n_fft / pitches must be multiples of 4, buffers 16-byte aligned");
  const int pad = (d.n_fft - d.hop) / 2;
  const int64_t n_in = static_cast<int64_t>(d.T) * d.hop;
  const int64_t n_out_max = (n_in * nnew + orig - 1) / orig;
  int fr = (8192 - K) / orig;
  fr = std::max(1, std::min(fr, 256));
  const int span = (fr - 1) * orig + K;
  const int64_t frames = (n_out_max + nnew - 1) / nnew;
  dim3 grid(static_cast<unsigned>((frames + fr - 1) / fr), d.B);
  const size_t smem = (static_cast<size_t>((span + 3) & ~3) + static_cast<size_t>(K) * nnew) * 4;
  if (nnew == 2) ola_resample_np_kernel<2><<<grid, 256, smem, stream>>>(d, pad, taps, K, width, orig, fr, y, y_pitch, n_out_max);
  else if (nnew == 3) ola_resample_np_kernel<3><<<grid, 256, smem, stream>>>(d, pad, taps, K, width, orig, fr, y, y_pitch, n_out_max);
  else ola_resample_np_kernel<1><<<grid, 256, smem, stream>>>(d, pad, taps, K, width, orig, fr, y, y_pitch, n_out_max);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// =====================================================================================================
// SIMT check GEMM (tests only): same contract as gemm_tc, one thread per output column pair.
// =====================================================================================================
__global__ void __launch_bounds__(128) gemm_ref_kernel(GemmDesc g) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int m = blockIdx.y;
  const int b = blockIdx.z;
  if (n >= g.N) return;
  const bool has2 = (n + 1) < g.N;
  const long long Ktot = static_cast<long long>(g.ntaps) * g.Kc;
  float acc0 = 0.f, acc1 = 0.f;
  for (int tap = 0; tap < g.ntaps; ++tap) {
    const int ra = m + tap + g.row_shift;
    if (ra < 0 || ra >= g.rows_a) continue;
    const __half* arow = g.A + static_cast<long long>(b) * g.a_batch_pitch + static_cast<long long>(ra) * g.a_row_pitch;
    const __half* w0 = g.W + static_cast<long long>(n) * Ktot + static_cast<long long>(tap) * g.Kc;
    const __half* w1 = w0 + Ktot;
    for (int c = 0; c < g.Kc; ++c) {
      const float a = __half2float(arow[c]);
      acc0 = fmaf(a, __half2float(w0[c]), acc0);
      if (has2) acc1 = fmaf(a, __half2float(w1[c]), acc1);
    }
  }
  float v0 = acc0 * g.alpha, v1 = acc1 * g.alpha;
  if (g.bias != nullptr) {
    v0 += g.bias[n];
    if (has2) v1 += g.bias[n + 1];
  }
  if (g.act == ACT_GELU) {
    v0 = gelu_erf(v0);
    v1 = gelu_erf(v1);
  } else if (g.act == ACT_POLAR) {
    const float mag = fminf(expf(v0), 100.0f);
    float s, c;
    sincosf(v1, &s, &c);
    v0 = mag * c;
    v1 = mag * s;
  }
  const int roff = (g.out_row_off != nullptr) ? g.out_row_off[b * g.row_off_stride] : 0;
  const long long o32 = static_cast<long long>(b) * g.pitch32 + static_cast<long long>(m + roff) * g.ld32 + n;
  const long long o16 = static_cast<long long>(b) * g.pitch16 + static_cast<long long>(m + roff) * g.ld16 + n;
  if (g.resid != nullptr) {
    v0 += g.resid[o32];
    if (has2) v1 += g.resid[o32 + 1];
  }
  if (g.out32 != nullptr) {
    g.out32[o32] = v0;
    if (has2) g.out32[o32 + 1] = v1;
  }
  if (g.out16 != nullptr) {
    g.out16[o16] = to_half_sat(v0);
    if (has2) g.out16[o16 + 1] = to_half_sat(v1);
  }
}

int gemm_ref(const GemmDesc& g, cudaStream_t stream) {
  dim3 grid((g.N / 2 + 1 + 127) / 128, g.rows_out, g.batches);
  gemm_ref_kernel<<<grid, 128, 0, stream>>>(g);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
