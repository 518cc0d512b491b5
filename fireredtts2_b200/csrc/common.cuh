// Shared declarations for the libfrt2_b200 kernels and host engine.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/frt2.h"

namespace frt2 {

// ---- thread-local last-error string (frt2_last_error) ----
void set_error(const std::string& msg);
const char* get_error();

#define FRT2_CUDA_OK(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::frt2::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")");                                 \
      return FRT2_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define FRT2_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != FRT2_OK) return _s; \
  } while (0)

#define FRT2_REQUIRE(cond, code, msg)                  \
  do {                                                 \
    if (!(cond)) {                                     \
      ::frt2::set_error(std::string(msg));             \
      return (code);                                   \
    }                                                  \
  } while (0)

// Device-side error word bits (always-on checks; read back by frt2_check_error / at sync points)
enum : unsigned int { DEV_ERR_INDEX_OOR = 1u };

// ---- device math helpers ----
__device__ __forceinline__ float gelu_erf(float x) {  // exact erf GELU (reference F.gelu / nn.GELU default)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// x * sigmoid(x) with the fast exponential and division (2 MUFU + 3 FP32 instead of ~20 instructions: the LN + SiLU
// kernels were XU/issue-limited at 3.0 TB/s); the result is rounded to fp16 right after, 2 ulp of fp32 do not show.
// x -> -inf: __expf overflows to +inf and the quotient is -0, as for the exact form.
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// fp32 -> fp16 conversions saturate to +-65504 instead of overflowing to inf (one F2FP.SATFINITE instruction):
// an out-of-range activation degrades gracefully instead of poisoning the utterance with NaNs.
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {  // lo = a, hi = b
  uint32_t y;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(b), "f"(a));
  return y;
}
__device__ __forceinline__ __half to_half_sat(float a) {
  unsigned short y;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(y) : "f"(a));
  return __ushort_as_half(y);
}

// ---- per-item streaming control block (HBM) ----
// One block of CTRL_INTS ints per batch item / pool slot.  The kernels of a captured per-token step read every
// position-dependent quantity from here, so the same CUDA graph serves every token of every stream; with one block
// per item, the items of a batch may sit at different positions of different streams (continuous batching).
constexpr int CTRL_INTS = 4;
constexpr int CTRL_POS = 0;     // 100 Hz frames consumed so far (K/V append row, attention length, "first chunk")
constexpr int CTRL_LAST = 1;    // this step is the item's last chunk (iSTFT keeps the trailing pad)
constexpr int CTRL_ACTIVE = 2;  // 0: the slot is idle this step — no state update, no output

// ---- GEMM (tcgen05) ----
enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_POLAR = 2,
                     ACT_SWIGLU = 3 };   // gemm_skinny only: weight rows interleaved (gate_j, up_j); column pair (2j, 2j+1)
                                         // -> silu(gate) * up stored at column j of out16 (the frame decoder's SwiGLU)

struct GemmDesc {
  // A operand: fp16, logical (batches, rows_a, Kc) with element pitches; K of the GEMM = ntaps*Kc.
  const __half* A;
  int64_t a_row_pitch;    // elements between consecutive rows (time steps)
  int64_t a_batch_pitch;  // elements between batch items
  int rows_a;             // addressable rows per batch item (TMA zero-fills outside [0, rows_a))
  int batches;
  int Kc;                 // channels per tap (multiple of 64)
  int ntaps;              // causal taps (1 = plain GEMM); weight K index = tap*Kc + c
  int row_shift;          // A row for output row m, tap j is  m + j + row_shift
  // B operand: fp16 weights (N, ntaps*Kc) row-major (K contiguous)
  const __half* W;
  int N;
  // output: rows_out rows per batch item; element address = b*pitchX + m*ldX + n
  int rows_out;
  int64_t pitch32;        // batch pitch (elements) of out32 / resid
  int64_t pitch16;        // batch pitch (elements) of out16
  float alpha;            // acc *= alpha before bias
  const float* bias;      // (N) or null
  int act;                // GemmAct
  const float* resid;     // fp32 (.., ld32) added after activation, may alias out32
  float* out32;           // fp32 output or null
  int64_t ld32;
  __half* out16;          // fp16 output or null
  int64_t ld16;
  const int* out_row_off; // optional device ints: out_row_off[b * row_off_stride] is added to the output row index of
  int row_off_stride;     // batch item b at run time (streaming KV append inside a captured CUDA graph: the write
                          // position lives in HBM, not in a kernel argument; stride 0 = one offset for all items)
  // optional second fp16 destination (gemm_skinny only): columns >= split_col go to out16_b at column (n - split_col)
  // with their own pitches / run-time row offset — Q and K|V of the streaming step come out of ONE launch, Q into the
  // chunk buffer and K|V appended to the HBM state
  int split_col;
  __half* out16_b;
  int64_t ld16_b, pitch16_b;
  const int* row_off_b;   // indexed with row_off_stride like out_row_off
  // LayerNorm folded across two tcgen05 GEMMs (offline path, gemm_tc only).  LN(x) W^T = rstd (x (gamma.W)^T - mean
  // colsum) + (bias + beta W^T): the PRODUCER of the fp32 residual stream x (a GEMM with `resid`) also writes an fp16
  // copy of x; row_stats_kernel reads that copy (2 B per element instead of LayerNorm's 4 in + 2 out) and leaves
  // (mean, rstd) per row; the CONSUMER runs on the raw fp16 rows with gamma-folded weights and finishes the
  // normalisation in its epilogue.
  __half* x16_out;        // producer: fp16 copy of the output rows, flat (batches*rows_out, ld_x16)
  int64_t ld_x16;
  const float* x16_shift; // producer, optional: (batches*rows_out) per-row offsets subtracted BEFORE the fp16 rounding of the
                          // copy (LayerNorm is shift-invariant).  The engine keeps the last known row mean of the residual
                          // stream there, so a row whose mean is many times its spread still gets all 11 mantissa bits for
                          // the part LayerNorm keeps (measured: 39 -> 55 dB on tests' adversarial weights)
  const float2* stats_in; // consumer: (mean, rstd) of its A rows
  const float* colsum;    // consumer: (N) sum_k of the folded fp16 weights of column n
  // optional fused LayerNorm(+SiLU) prologue (gemm_skinny only, ntaps == 1): the A rows are LN(ln_x) computed on the
  // fly from the fp32 residual stream (rows at ln_x + m*ln_ldx) instead of being read from A
  const float* ln_x;
  int64_t ln_ldx;
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  int ln_silu;
  int w_batch_k;          // gemm_tc (single-CTA tiles, no packing): batch item b multiplies with W[:, b * w_batch_k ...] — with
                          // a_batch_pitch = Kc this turns `batches` into a split of the reduction (partial sums per batch
                          // item at pitch32; the frame tail's large-batch down projection)
  int narrow_tiles;       // gemm_tc hint: 128-column tiles unless 256-column tiles would already fill every SM (few-row,
                          // weight-streaming-bound GEMMs of the frame tail's large-batch path)
  int ln_rms;             // 1: RMSNorm (no mean subtraction, no beta: ln_beta may be null) — the frame decoder's norms
};

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor is still running; it must execute pdl_wait() before touching anything the predecessor
// wrote.  pdl_trigger() lets the NEXT kernel of the stream begin its predecessor-independent prologue early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int gemm_tc(const GemmDesc& g, cudaStream_t stream);    // tcgen05 / TMEM / TMA path (product)
int gemm_ref(const GemmDesc& g, cudaStream_t stream);   // plain SIMT fp32-accumulate check kernel (tests only)
int gemm_tc_init();                                      // resolves cuTensorMapEncodeTiled, sets smem attrs
int gemm_skinny(const GemmDesc& g, cudaStream_t stream); // <= 16 rows in total: weight-streaming CUDA-core kernel
bool gemm_skinny_applicable(const GemmDesc& g);
int gemm_skinny_init();

// ---- persistent weight-streaming GEMM for <= 8 rows (gemm_stream.cu; the frame tail of the speech LM) ----
struct StreamGemm {
  const __half* Wt;       // weights in tile-blocked order [ceil(N/8)][K/32][8][32] (gemm_stream_pack_host)
  int N, K, B;            // K a multiple of 32; B <= gemm_stream_max_rows(K) <= 16 rows
  const __half* A;        // fp16 rows (B, K), pitch lda — or, when gamma != null, the rows are RMSNorm(x) * gamma:
  int64_t lda;
  const float* x;         // fp32 rows (B, K), pitch ldx
  int64_t ldx;
  const float* gamma;     // (K)
  float eps;
  const float* bias;      // (N) or null
  int act;                // ACT_NONE / ACT_GELU / ACT_SWIGLU (interleaved gate/up rows -> out16 column n/2)
  const float* resid;     // fp32 (B, ld32), may alias out32
  float* out32;
  int64_t ld32;
  __half* out16;
  int64_t ld16;
};
int gemm_stream(const StreamGemm& d, cudaStream_t stream);
int gemm_stream_init();
bool gemm_stream_applicable(int N, int K, int B);
int gemm_stream_max_rows(int K);       // activation rows of width K one launch can take (<= 16)
size_t gemm_stream_packed_elems(int64_t N, int64_t K);
void gemm_stream_pack_host(const float* W, int64_t N, int64_t K, __half* out);

// ---- attention ----
struct AttnDesc {
  const __half* q;   // (B, Tq, *) rows, head h at column h*hd
  int64_t q_row_pitch, q_batch_pitch;
  const __half* k;   // (B, Tk, *)
  const __half* v;
  int64_t kv_row_pitch, kv_batch_pitch;
  __half* out;       // (B, Tq, H*hd)
  int64_t o_row_pitch, o_batch_pitch;
  int B, H, hd, Tq, Tk;
  int q_pos0;        // absolute position of query row 0 (== Tk - Tq in streaming)
  int block_causal;  // 1: key j visible iff j <= ((q_pos0+i) | 7); 0: all Tk keys visible
  float scale;
  const int* ctrl;   // optional per-item control blocks (attention_warp): item b has q_pos0 = ctrl[b*CTRL_INTS + CTRL_POS],
                     // Tk = q_pos0 + Tq at run time; idle items (CTRL_ACTIVE == 0) are skipped
  // optional (attention_warp, split form): workspace for splitting one (item, head) over ATTN_KSPLIT CTAs —
  // part: (B*H, ATTN_KSPLIT, 8, hd + 2) floats, part_count: (B*H) ints, zero-initialised (each use leaves them zero)
  float* part;
  int* part_count;
  // optional (attention_tc): two zero-initialised device words for the persistent kernel's item scheduler (each launch
  // leaves them zero); launches sharing a pair must be stream-ordered.  nullptr: a library-owned ring is used.
  unsigned int* sched;
};
constexpr int ATTN_KSPLIT = 8;
int attention_warp(const AttnDesc& a, cudaStream_t stream);  // CUDA-core, one warp per 8-query block
int attention_tc(const AttnDesc& a, cudaStream_t stream);    // tcgen05 flash attention (hd == 64 or 128)
int attention_tc_init();
void attention_tc_set_trace(void* dev_buf);

// ---- misc kernels ----
int rvq_gather_sum(const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B, int nq, int L,
                   const float* tables /*(nq,K,D)*/, int K, int D, float* sum32 /*(B*L,D) or null*/,
                   __half* sum16 /*(B*L,D) or null*/, float* rows /*(B,L,nq,D) or null*/, unsigned int* err_word,
                   cudaStream_t stream);
// RVQ encode (rvq_encode.cu): everything fp32, [k][column] weight layouts built at load
struct RvqEncDesc {
  const float* z;            // (B, input_dim, T) with element strides sB, sD, sT
  int64_t sB, sD, sT;
  int B, T, nq;
  int input_dim, rd, cd, K;
  const float* WinpT;        // (input_dim, rd) input_proj^T or null (Identity: input_dim == rd)
  const float* binp;         // (rd)
  const float* WinT;         // (nq, rd, cd) in_project^T or null (Identity: rd == cd)
  const float* bin;          // (nq, cd)
  const float* CT;           // (nq, cd, K) transposed codebooks
  const float* c2;           // (nq, K) |c|^2
  const float* C;            // (nq, K, cd) codebooks
  const float* WoutT;        // (nq, cd, rd) out_project^T or null
  const float* bout;         // (nq, rd)
  long long* codes;          // (nq, B, T) int64
  // tensor-core variant (rvq_encode_tc.cu): split-fp16 weights, rows [w_hi | w_lo*S | w_hi/S], or null
  const __half* s_inp;       // (rd, 3*input_dim)
  const __half* s_in;        // (nq, cd, 3*rd)
  const __half* s_C;         // (nq, K, 3*cd)
  const __half* s_out;       // (nq, rd, 3*cd)
  const float* nbout;        // (nq, rd) = -bout
};
int rvq_encode(const RvqEncDesc& d, cudaStream_t stream);   // CUDA-core fp32 kernel (checker; any widths)
// every contraction of the chain as a tcgen05 GEMM on split-fp16 operands (widths that are multiples of 64)
bool rvq_encode_tc_applicable(const RvqEncDesc& d);
size_t rvq_encode_tc_ws_bytes(const RvqEncDesc& d, long long R);
int rvq_encode_tc(const RvqEncDesc& d, uint8_t* ws, cudaStream_t stream, long long* launches);
void rvq_split_weight_host(const float* W, int64_t N, int64_t Kk, __half* out);
int layer_norm_rows(const float* x, int64_t ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                    int apply_silu, __half* out16, int64_t ld16, cudaStream_t stream);
// Overlap-add of windowed frames + window-square envelope normalisation + "same" trimming.
struct OlaDesc {
  const float* frames;   // (B, T, n_fft) windowed frames of this call
  int64_t frames_batch_pitch;
  const float* tail;     // (B, 3, n_fft) carried windowed frames (streaming) or null
  const float* window;   // (n_fft)
  const int* lengths;    // per-item length (offline var-len) or null; valid frames = lengths[b]*len_mul
  int len_mul;
  float* audio;          // (B, *)
  int64_t audio_pitch;
  int16_t* pcm16;        // optional: write int16 PCM = trunc(sample * 32767) (saturated) instead of fp32, same pitch
  int B, T, n_fft, hop;
  int first, last;       // streaming flags; offline == first && last
  const int* ctrl;       // optional per-item control blocks: first = (pos == 0), last and active come from HBM
  const long long* out_off;  // optional (offline): item b's samples go to audio/pcm16 + out_off[b] instead of b*audio_pitch,
                             // and nothing is written past the item's own sample count (scatter into a shared buffer)
};
int istft_overlap_add(const OlaDesc& d, cudaStream_t stream);
// overlap-add + rational resampler in one kernel (offline decode): d.audio (24 kHz, optional) and y (resampled)
int istft_overlap_add_resample(const OlaDesc& d, const float* taps, int K, int width, int orig, int nnew, float* y,
                               int64_t y_pitch, cudaStream_t stream);
// mean_out: optional (rows) — the row means, left behind for the folded LayerNorm's shifted fp16 copy
int layer_norm_rows_batched(const float* x, int64_t ldx, int64_t rows, int rows_per_batch, int C, const float* gamma,
                            const float* beta, float eps, int apply_silu, __half* out16, int64_t ld16,
                            int64_t out_batch_pitch, cudaStream_t stream, float* mean_out = nullptr);
int resample_rows(const float* x, int64_t x_pitch, int B, int64_t n_in, const int* lengths, const float* taps, int K,
                  int width, int orig, int nnew, float* y, int64_t y_pitch, cudaStream_t stream);
// (mean, rstd) per row of an fp16 matrix (the statistics of a folded LayerNorm).  shift_io: optional (rows) — the offsets
// the producer subtracted from these rows; the kernel adds the measured mean so that they track the stream's row means
int row_stats(const __half* x16, int64_t ld, int64_t rows, int C, float eps, float2* stats, cudaStream_t stream,
              float* shift_io = nullptr);
// ---- streaming state kernels (stream_state.cu) ----
struct ShiftEntry {
  __half* p;
  int hist;
  int rows;  // chunk rows written by this call
  long long batch_pitch;
};
struct ShiftTable {
  ShiftEntry e[16];
  int n;
};
struct StateRoll {          // new iSTFT tail + conv histories to the head of their buffers + position advance
  const float* frames;
  long long frames_batch_pitch;
  float* tail;
  int T, n_fft, E, B;
  ShiftTable tb;
  int* ctrl;
  int advance_frames;
  int all_items;            // 1: no control blocks in use (eager streaming) — every item is active
};
int stream_state_roll(const StateRoll& ro, cudaStream_t stream);
// back to "no token consumed": zero history rows, control blocks and the n_err error words, stream-ordered
int stream_state_reset(const ShiftTable& tb, int E, int B, int* ctrl, unsigned int* err_words, int n_err,
                       cudaStream_t stream);
int tma_encode_fp16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
int num_sms();

}  // namespace frt2
