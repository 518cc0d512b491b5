// Host engine of libfrt2_b200: weight import / one-time repack, workspace, the offline decode pipeline
// (reference RedCodecInfer.decode, codec/model.py:307-324) and the in-HBM streaming state
// (reference RedCodecInfer.decode_one_token + cache_dict, codec/model.py:326-376).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace frt2 {

// ------------------------------------------------------------------ error string
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

namespace {

constexpr int DBG_TAPS = 1;       // record intermediates (frt2_get_tap)
constexpr int DBG_GEMM_REF = 2;   // route every GEMM through the SIMT check kernel (tests only)
constexpr int DBG_ATTN_WARP = 4;  // route attention through the warp kernel (tests only)
constexpr int DBG_NO_SKINNY = 16; // streaming: keep the tcgen05 GEMM even for <= 16 rows
constexpr int DBG_NO_GRAPH = 8;   // streaming: launch kernel by kernel instead of replaying the captured CUDA graph
constexpr int DBG_NO_LNFOLD = 32; // offline: separate LayerNorm kernels instead of LN folded across the GEMMs (A/B, tests)

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

struct ResW {
  float *ln1_g, *ln1_b, *b1, *ln2_g, *ln2_b, *b2;
  __half *w1, *w2;  // (E, 3E) tap-major
};
struct LayerW {
  float *ln1_g, *ln1_b, *b_qkv, *b_o, *ln2_g, *ln2_b, *b_fc1, *b_fc2;
  __half *w_qkv, *w_o, *w_fc1, *w_fc2;
  // LayerNorm folded into the consuming GEMM: W' = W diag(gamma) (fp16), colsum[n] = sum_k W'[n,k] (of the ROUNDED
  // fp16 values the tensor cores see), bias' = bias + W beta
  __half *w_qkv_f, *w_fc1_f;
  float *s_qkv, *c_qkv, *s_fc1, *c_fc1;
};

// small elementwise helpers -------------------------------------------------------------------------
__global__ void half_to_float_kernel(const __half* __restrict__ src, long long ld, long long rows, int cols,
                                     float* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = static_cast<int>(i - r * cols);
  dst[i] = __half2float(src[r * ld + c]);
}

// strided int32/int64 tokens -> contiguous int32 staging (B,nq,L) + this call's flags into the per-item control
// blocks; out-of-range values are flagged here (an int64 could alias after narrowing) and decoded as code 0.
// slot_flags == nullptr (frt2_decode_chunk): every item is active and shares `last`.  Otherwise (frt2_pool_step) item b
// takes FRT2_SLOT_* bits from slot_flags.f[b]: idle slots get token 0 and no range check, RESET rewinds the slot.
struct SlotFlags { unsigned char f[FRT2_POOL_MAX_SLOTS]; };
template <typename IdxT, bool POOL>
__global__ void stage_tokens_kernel(const IdxT* __restrict__ tokens, long long sB, long long sQ, long long sL, int B,
                                    int nq, int L, int K, int* __restrict__ stage, int* ctrl, int last,
                                    unsigned int* err_words, const SlotFlags flags) {
  const int n = B * nq * L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    int* cb = ctrl + i * CTRL_INTS;
    if (POOL) {
      const int f = flags.f[i];
      if (f & FRT2_SLOT_RESET) cb[CTRL_POS] = 0;
      cb[CTRL_LAST] = (f & FRT2_SLOT_LAST) ? 1 : 0;
      cb[CTRL_ACTIVE] = (f & FRT2_SLOT_ACTIVE) ? 1 : 0;
    } else {
      cb[CTRL_LAST] = last;
      cb[CTRL_ACTIVE] = 1;
    }
  }
  if (i >= n) return;
  const int l = i % L, q = (i / L) % nq, b = i / (L * nq);
  if (POOL && !(flags.f[b] & FRT2_SLOT_ACTIVE)) {
    stage[i] = 0;
    return;
  }
  const long long raw = static_cast<long long>(tokens[b * sB + q * sQ + l * sL]);
  int v = static_cast<int>(raw);
  if (raw < 0 || raw >= K) {
    atomicOr(err_words, DEV_ERR_INDEX_OOR);           // the stream / pool as a whole
    atomicOr(err_words + 1 + b, DEV_ERR_INDEX_OOR);   // and the item (pool slot) that sent the bad code
    v = 0;   // decoded as code 0; the caller learns about it through the error word
  }
  stage[i] = v;
}
// staged fp32 chunk -> the caller's int16 PCM buffer, same rounding as the overlap-add kernel's direct PCM output
__global__ void emit_pcm16_kernel(const float* __restrict__ stage, long long stage_pitch, int16_t* __restrict__ pcm,
                                  long long pcm_pitch, int n, int B) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || b >= B) return;
  const float smp = stage[b * stage_pitch + i];
  pcm[b * pcm_pitch + i] = static_cast<int16_t>(__float2int_rz(fminf(fmaxf(smp * 32767.0f, -32768.0f), 32767.0f)));
}

// pool: a slot that starts a new stream this step gets its causal left padding back (zero conv history); the K/V
// state and the iSTFT tail need no clearing — nothing before position 0 is ever read
__global__ void reset_history_kernel(ShiftTable t, int E, int B, const SlotFlags flags) {
  const int ei = blockIdx.y;
  const ShiftEntry en = t.e[ei];
  const long long total = static_cast<long long>(B) * en.hist * E;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / (static_cast<long long>(E) * en.hist));
    if (!(flags.f[b] & FRT2_SLOT_RESET)) continue;
    const long long rc = i - static_cast<long long>(b) * en.hist * E;
    en.p[b * en.batch_pitch + rc] = __float2half_rn(0.f);
  }
}

// state export / import in the reference's cache layouts ----------------------------------------------
// src (B, rows, C) fp16 time-major with pitches  ->  dst (B, C_total, rows) fp32 channel-major at channel offset
__global__ void export_tm_to_cm_kernel(const __half* src, long long batch_pitch, int rows, int C, float* dst,
                                       int C_total, int c_off, int t_total, int t_off, int B) {
  const long long n = static_cast<long long>(B) * rows * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int r = static_cast<int>((i / C) % rows);
    const int b = static_cast<int>(i / (static_cast<long long>(C) * rows));
    dst[(static_cast<long long>(b) * C_total + c_off + c) * t_total + t_off + r] =
        __half2float(src[b * batch_pitch + static_cast<long long>(r) * C + c]);
  }
}
__global__ void import_cm_to_tm_kernel(const float* src, int C_total, int c_off, int t_total, int t_off, __half* dst,
                                       long long batch_pitch, int rows, int C, int B) {
  const long long n = static_cast<long long>(B) * rows * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int r = static_cast<int>((i / C) % rows);
    const int b = static_cast<int>(i / (static_cast<long long>(C) * rows));
    dst[b * batch_pitch + static_cast<long long>(r) * C + c] =
        to_half_sat(src[(static_cast<long long>(b) * C_total + c_off + c) * t_total + t_off + r]);
  }
}
// kv state (B, Tmax, 2E) fp16 [k | v]  <->  reference (B, nl, H, T, 2*hd) fp32 slice of layer `layer`
__global__ void export_kv_kernel(const __half* kv, long long batch_pitch, int T, int E, int H, int hd, float* dst,
                                 int nl, int layer, int B) {
  const long long n = static_cast<long long>(B) * T * 2 * E;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (2 * E));
    const int t = static_cast<int>((i / (2 * E)) % T);
    const int b = static_cast<int>(i / (static_cast<long long>(2 * E) * T));
    const int isv = c >= E;
    const int ch = isv ? c - E : c;
    const int h = ch / hd, d = ch - h * hd;
    dst[((((static_cast<long long>(b) * nl + layer) * H + h) * T + t) * 2 + isv) * hd + d] =
        __half2float(kv[b * batch_pitch + static_cast<long long>(t) * 2 * E + c]);
  }
}
__global__ void import_kv_kernel(const float* src, int nl, int layer, __half* kv, long long batch_pitch, int T, int E,
                                 int H, int hd, int B) {
  const long long n = static_cast<long long>(B) * T * 2 * E;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (2 * E));
    const int t = static_cast<int>((i / (2 * E)) % T);
    const int b = static_cast<int>(i / (static_cast<long long>(2 * E) * T));
    const int isv = c >= E;
    const int ch = isv ? c - E : c;
    const int h = ch / hd, d = ch - h * hd;
    kv[b * batch_pitch + static_cast<long long>(t) * 2 * E + c] =
        to_half_sat(src[((((static_cast<long long>(b) * nl + layer) * H + h) * T + t) * 2 + isv) * hd + d]);
  }
}
// is_cache (B, n_fft, 3) fp32 channel-major <-> tail (B, 3, n_fft) fp32
__global__ void transpose_tail_kernel(const float* src, float* dst, int B, int n_fft, int to_reference) {
  const int n = B * 3 * n_fft;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = i % n_fft, t = (i / n_fft) % 3, b = i / (3 * n_fft);
  const long long tm = (static_cast<long long>(b) * 3 + t) * n_fft + k;
  const long long cm = (static_cast<long long>(b) * n_fft + k) * 3 + t;
  if (to_reference) dst[cm] = src[tm];
  else dst[tm] = src[cm];
}

inline unsigned grid_for(long long n, int block = 256, unsigned cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  return static_cast<unsigned>(std::max<long long>(1, std::min<long long>(g, cap)));
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

// =====================================================================================================
struct Stream;

struct Handle {
  frt2_config cfg{};
  int device = 0;
  bool finalized = false;
  int debug = 0;
  std::map<std::string, HostTensor> raw;
  std::vector<void*> owned;  // device allocations of packed weights
  std::mutex mu;

  // derived
  int E = 0, H = 0, hd = 0, nl = 0, nq = 0, K = 0, cd = 0, rd = 0, hop = 0, n_fft = 0, n_bins = 0, spec_ld = 0;
  bool has_out_project = false, has_output_proj = false;

  // packed weights
  float* codebooks = nullptr;  // (nq,K,cd) raw
  float* tables = nullptr;     // (nq,K,rd) folded with out_project (== codebooks when Identity)
  // RVQ encode side (optional: built when the checkpoint's in_project / input_proj tensors were handed over)
  bool enc_ready = false;
  int enc_input_dim = 0;
  float *enc_WinpT = nullptr, *enc_binp = nullptr, *enc_WinT = nullptr, *enc_bin = nullptr, *enc_CT = nullptr,
        *enc_c2 = nullptr, *enc_WoutT = nullptr, *enc_bout = nullptr, *enc_nbout = nullptr;
  __half *enc_s_inp = nullptr, *enc_s_in = nullptr, *enc_s_C = nullptr, *enc_s_out = nullptr;   // split-fp16 (rvq_encode_tc)
  int upload_split(const std::vector<float>& W, int64_t N, int64_t Kk, __half** out) {
    std::vector<__half> hb(static_cast<size_t>(N) * 3 * Kk);
    rvq_split_weight_host(W.data(), N, Kk, hb.data());
    FRT2_TRY(dev_alloc(reinterpret_cast<void**>(out), hb.size() * 2));
    FRT2_CUDA_OK(cudaMemcpy(*out, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }
  int build_rvq_encoder();
  __half* w_outproj = nullptr; float* b_outproj = nullptr;
  __half* w_up_in = nullptr;   float* b_up_in = nullptr;
  __half* w_up_conv = nullptr;
  __half* w_up_comp = nullptr; float* b_up_comp = nullptr;   // output_proj . in_proj . up_conv composed at load: (4E, rd)
  __half* w_us0 = nullptr;     float* b_us0 = nullptr;
  __half* w_us2 = nullptr;     float* b_us2 = nullptr;
  __half* w_inproj = nullptr;  float* b_inproj = nullptr;
  ResW res[4]{};
  std::vector<LayerW> layers;
  float *fn_g = nullptr, *fn_b = nullptr;
  __half* w_head = nullptr;    float* b_head = nullptr;
  __half* w_head_f = nullptr;  float* s_head = nullptr;  float* c_head = nullptr;   // final_norm folded into the head
  __half* w_idft = nullptr;
  float* window = nullptr;
  unsigned int* err_word = nullptr;

  // optional per-kernel-class CUDA-event timing (bench.py roofline numbers) and launch counter
  struct ProfRec { int cls; cudaEvent_t a, b; double flops, bytes; };
  bool profile = false;
  std::vector<ProfRec> prof;
  size_t prof_n = 0;
  long long launches = 0;
  int prof_begin(int cls, double flops, double bytes, cudaStream_t st) {
    ++launches;
    if (!profile) return -1;
    if (prof_n == prof.size()) {
      ProfRec r{};
      if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
      prof.push_back(r);
    }
    ProfRec& r = prof[prof_n];
    r.cls = cls; r.flops = flops; r.bytes = bytes;
    cudaEventRecord(r.a, st);
    return static_cast<int>(prof_n++);
  }
  void prof_end(int id, cudaStream_t st) {
    if (id >= 0) cudaEventRecord(prof[id].b, st);
  }

  int16_t* pcm16_out = nullptr;  // set (under the mutex) by frt2_decode_pcm16 for the next pipeline run
  const long long* scatter_off = nullptr;  // set (under the mutex) by frt2_decode_scatter: per-item output offsets
  // set (under the mutex) by frt2_decode_resampled: the overlap-add kernel also resamples (taps of the rate pair)
  float* rs_out = nullptr;
  int64_t rs_pitch = 0;
  const float* rs_taps = nullptr;
  int rs_K = 0, rs_width = 0, rs_orig = 0, rs_new = 0;

  // workspace arena (grow-only) of the offline decode and of streaming chunks that do not run as a captured step.
  // The arena is shared by every call on this handle, and calls are asynchronous: a call that arrives on another CUDA
  // stream than the previous user first waits (on the device) for that user's event.
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  cudaEvent_t ws_event = nullptr;
  cudaStream_t ws_last_stream = nullptr;
  bool ws_in_use = false;
  int ws_acquire(cudaStream_t st) {
    if (ws_in_use && st != ws_last_stream) FRT2_CUDA_OK(cudaStreamWaitEvent(st, ws_event, 0));
    return FRT2_OK;
  }
  int ws_release(cudaStream_t st) {
    if (ws_event == nullptr) FRT2_CUDA_OK(cudaEventCreateWithFlags(&ws_event, cudaEventDisableTiming));
    FRT2_CUDA_OK(cudaEventRecord(ws_event, st));
    ws_last_stream = st;
    ws_in_use = true;
    return FRT2_OK;
  }
  // stream states handed back by frt2_stream_destroy, kept (with their captured step graphs) for the next
  // frt2_stream_create of the same shape: decode_one_token(tok, {}, ...) — the reference's call pattern — then costs no
  // allocation and no capture
  std::vector<frt2_stream*> free_streams;
  size_t free_stream_bytes = 0;
  std::map<std::string, std::pair<float*, int64_t>> taps;
  int tap_B = 0, tap_L = 0;

  ~Handle();
  void release_base() {
    cudaSetDevice(device);
    for (void* p : owned) cudaFree(p);
    if (ws) cudaFree(ws);
    if (ws_event) cudaEventDestroy(ws_event);
    for (auto& kv : taps) cudaFree(kv.second.first);
    for (auto& r : prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  }

  int dev_alloc(void** p, size_t bytes) {
    FRT2_CUDA_OK(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    owned.push_back(*p);
    return FRT2_OK;
  }
  int upload_f32(const std::vector<float>& v, float** out) {
    FRT2_TRY(dev_alloc(reinterpret_cast<void**>(out), v.size() * 4));
    FRT2_CUDA_OK(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }
  int upload_f16(const std::vector<float>& v, __half** out) {
    std::vector<__half> hbuf(v.size());
    const long long n = static_cast<long long>(v.size());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) hbuf[i] = __float2half_rn(std::min(65504.0f, std::max(-65504.0f, v[i])));
    FRT2_TRY(dev_alloc(reinterpret_cast<void**>(out), hbuf.size() * 2));
    FRT2_CUDA_OK(cudaMemcpy(*out, hbuf.data(), hbuf.size() * 2, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }

  // fold LayerNorm(gamma, beta) into the Linear (W (N,K) row-major, bias (N) or null) that consumes it
  int fold_ln(const float* W, const float* bias, const float* gamma, const float* beta, int64_t Nn, int64_t Kk,
              __half** w_f, float** colsum, float** bias_f) {
    std::vector<float> Wf(static_cast<size_t>(Nn) * Kk), cs(Nn), bf(Nn);
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < Nn; ++n) {
      double sacc = 0.0, bacc = bias ? bias[n] : 0.0;
      for (int64_t k = 0; k < Kk; ++k) {
        const float w = W[n * Kk + k];
        const float wf = w * gamma[k];
        Wf[n * Kk + k] = wf;
        sacc += static_cast<double>(__half2float(__float2half_rn(std::min(65504.0f, std::max(-65504.0f, wf)))));
        bacc += static_cast<double>(beta[k]) * w;
      }
      cs[n] = static_cast<float>(sacc);
      bf[n] = static_cast<float>(bacc);
    }
    FRT2_TRY(upload_f16(Wf, w_f));
    FRT2_TRY(upload_f32(cs, colsum));
    FRT2_TRY(upload_f32(bf, bias_f));
    return FRT2_OK;
  }

  const HostTensor* find(const std::string& key) const {
    auto it = raw.find(key);
    return it == raw.end() ? nullptr : &it->second;
  }
  int need(const std::string& key, const HostTensor** out, std::initializer_list<int64_t> shape) {
    const HostTensor* t = find(key);
    if (t == nullptr) {
      set_error("missing tensor: " + key);
      return FRT2_ERR_MISSING_TENSOR;
    }
    std::vector<int64_t> want(shape);
    if (t->shape != want) {
      std::string got, exp;
      for (auto d : t->shape) got += std::to_string(d) + ",";
      for (auto d : want) exp += std::to_string(d) + ",";
      set_error("tensor " + key + " has shape (" + got + ") expected (" + exp + ")");
      return FRT2_ERR_BAD_ARG;
    }
    *out = t;
    return FRT2_OK;
  }

  int finalize();
  int ensure_ws(size_t bytes);
  int run_gemm(const GemmDesc& g, cudaStream_t st) {
    const double flops = 2.0 * g.batches * g.rows_out * static_cast<double>(g.N) * g.ntaps * g.Kc;
    const double bytes = 2.0 * (static_cast<double>(g.batches) * g.rows_a * g.Kc + static_cast<double>(g.N) * g.ntaps * g.Kc) +
                         static_cast<double>(g.batches) * g.rows_out * g.N *
                             ((g.out32 ? 4 : 0) + (g.out16 ? 2 : 0) + (g.resid ? 4 : 0) + (g.x16_out ? 2 : 0));
    const bool skinny = !(debug & (DBG_GEMM_REF | DBG_NO_SKINNY)) && gemm_skinny_applicable(g);
    const int id = prof_begin(skinny ? FRT2_PROF_GEMM_SKINNY : FRT2_PROF_GEMM, flops, bytes, st);
    const int rc = (debug & DBG_GEMM_REF) ? gemm_ref(g, st) : (skinny ? gemm_skinny(g, st) : gemm_tc(g, st));
    prof_end(id, st);
    return rc;
  }
  int run_attn(const AttnDesc& a, cudaStream_t st) {
    // visible (query, key) pairs: block-causal sum_q ((q_pos0+q)|7)+1, else Tq*Tk
    double pairs = static_cast<double>(a.Tq) * a.Tk;
    if (a.block_causal) {
      pairs = 0;
      for (int q = 0; q < a.Tq; q += 8) pairs += 8.0 * std::min(a.Tk, ((a.q_pos0 + q) | 7) + 1);
    }
    const double flops = 4.0 * a.hd * pairs * a.B * a.H;
    const double bytes = 2.0 * a.B * a.H * a.hd * (2.0 * a.Tq + 2.0 * a.Tk);
    const bool warp = (debug & DBG_ATTN_WARP) || (a.hd != 64 && a.hd != 128) || a.Tq < 32;
    const int id = prof_begin(warp ? FRT2_PROF_ATTN_WARP : FRT2_PROF_ATTN_TC, flops, bytes, st);
    const int rc = warp ? attention_warp(a, st) : attention_tc(a, st);
    prof_end(id, st);
    return rc;
  }
  int run_ln(const float* x, int64_t rows, int rows_per_batch, const float* g, const float* b, float eps, int silu_,
             __half* out, int64_t pitch, cudaStream_t st, float* mean_out = nullptr) {
    const int id = prof_begin(FRT2_PROF_LAYER_NORM, 0.0, static_cast<double>(rows) * E * 6.0, st);
    const int rc = layer_norm_rows_batched(x, E, rows, rows_per_batch, E, g, b, eps, silu_, out, E, pitch, st, mean_out);
    prof_end(id, st);
    return rc;
  }
  int tap_f32(const char* name, const float* src, int64_t n, cudaStream_t st);
  int tap_f16(const char* name, const __half* src, int64_t ld, int64_t rows, int cols, cudaStream_t st);

  int pipeline(const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B, int nq_in, int L,
               const int32_t* lengths, float* audio, int64_t audio_pitch, Stream* s, int last, cudaStream_t st,
               bool graph_mode = false);
  size_t ws_bytes_for(int B, int L) const;
};

struct Stream {
  Handle* h = nullptr;
  int B = 0, max_tokens = 0, n_tokens = 0, chunk_cap = 0;
  // conv input buffers [hist | chunk], fp16: x50 (hist 1, 4 rows/token), a (hist 2), in_proj in (hist 6), 8 x res (hist 2)
  __half* conv[11]{};
  int conv_hist[11] = {1, 2, 6, 2, 2, 2, 2, 2, 2, 2, 2};
  int conv_rpt[11] = {4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8};  // rows per token
  std::vector<__half*> kv;  // per layer (B, Tmax, 2E)
  float* tail = nullptr;    // (B, 3, n_fft)
  float* attn_part = nullptr;   // (B*H, ATTN_KSPLIT, 8, hd + 2): cross-CTA split of the step's attention (hd == 64 only)
  int* attn_count = nullptr;    // (B*H) arrival counters, zero between launches
  int* ctrl = nullptr;      // device (B, CTRL_INTS) per-item control blocks: read by the kernels of the captured step
  // device error words of THIS stream / pool (1 + B): word 0 = some item sent an out-of-range code, word 1 + b = item b
  // did.  A handle-global word would let one request consume another request's error.
  unsigned int* err_words = nullptr;
  bool pooled = false;      // slot pool (frt2_pool_*): items are independent streams at their own positions
  std::vector<int> slot_tokens;  // pool: host mirror of the tokens each slot has consumed
  std::vector<char> slot_done;   // pool: the slot's stream received its LAST token
  int* tok_stage = nullptr; // (B, nq, chunk_cap) int32 contiguous
  float* audio_stage = nullptr;  // (B, audio_stage_pitch)
  int64_t audio_stage_pitch = 0;
  // step workspace of the captured per-token step (chunks of <= 3 tokens): private to the stream, so streams that
  // decode concurrently on different CUDA streams never share activations, and the captured graph never goes stale
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy default stream)
  struct GraphRec { cudaGraphExec_t exec; long long kernels; };
  std::map<std::pair<int, int>, GraphRec> graphs;  // (Lc, nq) -> captured step
  // stream-ordered bookkeeping: the state is reset by a kernel on the stream of the NEXT decode call, and a call that
  // arrives on another CUDA stream than the previous one waits (on the device) for the previous call's event
  bool pending_reset = true;
  cudaEvent_t last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool used = false;

  int64_t conv_pitch(int i) const { return static_cast<int64_t>(conv_hist[i] + conv_rpt[i] * chunk_cap) * h->E; }
  // one token of slack: an idle pool slot still "appends" (and later overwrites) one chunk at its current position
  int64_t kv_pitch() const { return static_cast<int64_t>(max_tokens + 1) * 8 * 2 * h->E; }
  size_t footprint() const {   // device bytes this state holds (bounds the handle's pool of spare states)
    size_t n = static_cast<size_t>(h->nl) * B * kv_pitch() * 2 + ws_bytes;
    for (int i = 0; i < 11; ++i) n += static_cast<size_t>(B) * conv_pitch(i) * 2;
    return n;
  }
  ShiftTable shift_table(int rows_mul) const {
    ShiftTable tb{};
    tb.n = 11;
    for (int i = 0; i < 11; ++i) tb.e[i] = {conv[i], conv_hist[i], conv_rpt[i] * rows_mul, conv_pitch(i)};
    return tb;
  }
  void free_conv() {
    for (auto& p : conv) {
      if (p) cudaFree(p);
      p = nullptr;
    }
  }
  ~Stream() {
    cudaSetDevice(h->device);
    free_conv();
    for (auto p : kv) cudaFree(p);
    if (tail) cudaFree(tail);
    if (ctrl) cudaFree(ctrl);
    if (err_words) cudaFree(err_words);
    if (attn_part) cudaFree(attn_part);
    if (attn_count) cudaFree(attn_count);
    if (tok_stage) cudaFree(tok_stage);
    if (audio_stage) cudaFree(audio_stage);
    if (ws) cudaFree(ws);
    drop_graphs();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (last_use) cudaEventDestroy(last_use);
  }
  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
    graphs.clear();
  }
  int ensure_chunk_cap(int Lc, cudaStream_t st);
  int ensure_ws(size_t bytes);
  // order this call after the stream's previous user and apply a pending reset, both on `st`
  int begin_use(cudaStream_t st);
  int end_use(cudaStream_t st);
  void mark_reset();
};

// ------------------------------------------------------------------ weight repack (host, once)
static std::vector<float> weight_norm_host(const HostTensor& g, const HostTensor& v) {
  // W[o] = g[o] * v[o] / ||v[o]||  (torch weight_norm dim=0, reference rvq.py:8-13)
  const int64_t out = v.shape[0];
  const int64_t inner = v.numel() / out;
  std::vector<float> W(v.data.size());
#pragma omp parallel for schedule(static)
  for (long long o = 0; o < out; ++o) {
    double ss = 0.0;
    for (int64_t i = 0; i < inner; ++i) ss += static_cast<double>(v.data[o * inner + i]) * v.data[o * inner + i];
    const float scale = g.data[o] / static_cast<float>(std::sqrt(ss));
    for (int64_t i = 0; i < inner; ++i) W[o * inner + i] = v.data[o * inner + i] * scale;
  }
  return W;
}

// Encode-side tables of ResidualVQ (rvq.py:62-89,128-143), fp32, [k][column] layouts for coalesced reads.
int Handle::build_rvq_encoder() {
  const std::string RVQ = "rvq.";
  if (has_out_project && find(RVQ + "quantizers.0.in_project.bias") == nullptr) return FRT2_OK;   // decode-only weights
  const HostTensor* t = nullptr;
  // input_proj: WNConv1d(input_dim, rvq_dim, 1) or Identity (rvq.py:110-114)
  enc_input_dim = rd;
  if (const HostTensor* v = find(RVQ + "input_proj.parametrizations.weight.original1")) {
    FRT2_REQUIRE(v->shape.size() == 3 && v->shape[0] == rd && v->shape[2] == 1, FRT2_ERR_BAD_ARG,
                 "rvq.input_proj weight must be (rvq_dim, input_dim, 1)");
    enc_input_dim = static_cast<int>(v->shape[1]);
    const HostTensor *g, *b;
    FRT2_TRY(need(RVQ + "input_proj.parametrizations.weight.original0", &g, {rd, 1, 1}));
    FRT2_TRY(need(RVQ + "input_proj.bias", &b, {rd}));
    const std::vector<float> W = weight_norm_host(*g, *v);   // (rd, input_dim)
    std::vector<float> WT(W.size());
    for (int o = 0; o < rd; ++o)
      for (int k = 0; k < enc_input_dim; ++k) WT[static_cast<size_t>(k) * rd + o] = W[static_cast<size_t>(o) * enc_input_dim + k];
    FRT2_TRY(upload_f32(WT, &enc_WinpT));
    FRT2_TRY(upload_f32(b->data, &enc_binp));
    FRT2_TRY(upload_split(W, rd, enc_input_dim, &enc_s_inp));
  }
  std::vector<float> CT(static_cast<size_t>(nq) * cd * K), c2(static_cast<size_t>(nq) * K);
  std::vector<float> WinT, bin, WoutT, bout;
  std::vector<float> Win_all, Wout_all, C_all(static_cast<size_t>(nq) * K * cd);   // (out, in) row-major, for the split copies
  if (has_out_project) {
    Win_all.resize(static_cast<size_t>(nq) * cd * rd);
    Wout_all.resize(static_cast<size_t>(nq) * rd * cd);
    WinT.resize(static_cast<size_t>(nq) * rd * cd);
    bin.resize(static_cast<size_t>(nq) * cd);
    WoutT.resize(static_cast<size_t>(nq) * cd * rd);
    bout.resize(static_cast<size_t>(nq) * rd);
  }
  for (int i = 0; i < nq; ++i) {
    const std::string q = RVQ + "quantizers." + std::to_string(i);
    FRT2_TRY(need(q + ".codebook", &t, {K, cd}));
    std::memcpy(&C_all[static_cast<size_t>(i) * K * cd], t->data.data(), static_cast<size_t>(K) * cd * 4);
    for (int64_t k = 0; k < K; ++k) {
      double ss = 0.0;
      for (int c = 0; c < cd; ++c) {
        const float v = t->data[k * cd + c];
        CT[(static_cast<size_t>(i) * cd + c) * K + k] = v;
        ss += static_cast<double>(v) * v;
      }
      c2[static_cast<size_t>(i) * K + k] = static_cast<float>(ss);
    }
    if (has_out_project) {
      const HostTensor *g, *v, *b;
      FRT2_TRY(need(q + ".in_project.parametrizations.weight.original0", &g, {cd, 1, 1}));
      FRT2_TRY(need(q + ".in_project.parametrizations.weight.original1", &v, {cd, rd, 1}));
      FRT2_TRY(need(q + ".in_project.bias", &b, {cd}));
      const std::vector<float> Wi = weight_norm_host(*g, *v);   // (cd, rd)
      for (int o = 0; o < cd; ++o)
        for (int k = 0; k < rd; ++k) WinT[(static_cast<size_t>(i) * rd + k) * cd + o] = Wi[static_cast<size_t>(o) * rd + k];
      std::memcpy(&bin[static_cast<size_t>(i) * cd], b->data.data(), static_cast<size_t>(cd) * 4);
      std::memcpy(&Win_all[static_cast<size_t>(i) * cd * rd], Wi.data(), Wi.size() * 4);
      FRT2_TRY(need(q + ".out_project.parametrizations.weight.original0", &g, {rd, 1, 1}));
      FRT2_TRY(need(q + ".out_project.parametrizations.weight.original1", &v, {rd, cd, 1}));
      FRT2_TRY(need(q + ".out_project.bias", &b, {rd}));
      const std::vector<float> Wo = weight_norm_host(*g, *v);   // (rd, cd)
      for (int o = 0; o < rd; ++o)
        for (int k = 0; k < cd; ++k) WoutT[(static_cast<size_t>(i) * cd + k) * rd + o] = Wo[static_cast<size_t>(o) * cd + k];
      std::memcpy(&bout[static_cast<size_t>(i) * rd], b->data.data(), static_cast<size_t>(rd) * 4);
      std::memcpy(&Wout_all[static_cast<size_t>(i) * rd * cd], Wo.data(), Wo.size() * 4);
    }
  }
  FRT2_TRY(upload_f32(CT, &enc_CT));
  FRT2_TRY(upload_f32(c2, &enc_c2));
  if (has_out_project) {
    FRT2_TRY(upload_f32(WinT, &enc_WinT));
    FRT2_TRY(upload_f32(bin, &enc_bin));
    FRT2_TRY(upload_f32(WoutT, &enc_WoutT));
    FRT2_TRY(upload_f32(bout, &enc_bout));
  }
  // split-fp16 copies for the tensor-core chain (widths that are multiples of 64; the CUDA-core kernel serves the rest)
  if (rd % 64 == 0 && cd % 64 == 0 && enc_input_dim % 64 == 0) {
    FRT2_TRY(upload_split(C_all, static_cast<int64_t>(nq) * K, cd, &enc_s_C));
    if (has_out_project) {
      FRT2_TRY(upload_split(Win_all, static_cast<int64_t>(nq) * cd, rd, &enc_s_in));
      FRT2_TRY(upload_split(Wout_all, static_cast<int64_t>(nq) * rd, cd, &enc_s_out));
      std::vector<float> nb(bout.size());
      for (size_t j = 0; j < nb.size(); ++j) nb[j] = -bout[j];
      FRT2_TRY(upload_f32(nb, &enc_nbout));
    }
  }
  enc_ready = true;
  return FRT2_OK;
}

int Handle::finalize() {
  FRT2_CUDA_OK(cudaSetDevice(device));
  FRT2_TRY(gemm_tc_init());
  FRT2_TRY(gemm_skinny_init());
  FRT2_TRY(attention_tc_init());
  const HostTensor* t = nullptr;
  const std::string RVQ = "rvq.", UP = "upsample.", AD = "acoustic_decoder.", BB = "acoustic_decoder.backbone.";

  // ---- RVQ: raw codebooks + tables folded with the weight-normed out_project (+ bias) ----
  {
    std::vector<float> cbs(static_cast<size_t>(nq) * K * cd);
    std::vector<float> tab;
    if (has_out_project) tab.resize(static_cast<size_t>(nq) * K * rd);
    for (int i = 0; i < nq; ++i) {
      const std::string q = RVQ + "quantizers." + std::to_string(i);
      FRT2_TRY(need(q + ".codebook", &t, {K, cd}));
      std::memcpy(&cbs[static_cast<size_t>(i) * K * cd], t->data.data(), static_cast<size_t>(K) * cd * 4);
      if (has_out_project) {
        const HostTensor *g, *v, *bq;
        FRT2_TRY(need(q + ".out_project.parametrizations.weight.original0", &g, {rd, 1, 1}));
        FRT2_TRY(need(q + ".out_project.parametrizations.weight.original1", &v, {rd, cd, 1}));
        FRT2_TRY(need(q + ".out_project.bias", &bq, {rd}));
        const std::vector<float> W = weight_norm_host(*g, *v);  // (rd, cd)
        const float* cb = t->data.data();
        float* dst = &tab[static_cast<size_t>(i) * K * rd];
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < K; ++k) {
          for (int o = 0; o < rd; ++o) {
            float acc = 0.f;
            const float* wr = &W[static_cast<size_t>(o) * cd];
            const float* cr = &cb[static_cast<size_t>(k) * cd];
            for (int c = 0; c < cd; ++c) acc += cr[c] * wr[c];
            dst[static_cast<size_t>(k) * rd + o] = acc + bq->data[o];
          }
        }
      }
    }
    FRT2_TRY(upload_f32(cbs, &codebooks));
    if (has_out_project) FRT2_TRY(upload_f32(tab, &tables));
    else tables = codebooks;
  }
  FRT2_TRY(build_rvq_encoder());
  if (has_output_proj) {
    const HostTensor *g, *v, *bo;
    FRT2_TRY(need(RVQ + "output_proj.parametrizations.weight.original0", &g, {E, 1, 1}));
    FRT2_TRY(need(RVQ + "output_proj.parametrizations.weight.original1", &v, {E, rd, 1}));
    FRT2_TRY(need(RVQ + "output_proj.bias", &bo, {E}));
    FRT2_TRY(upload_f16(weight_norm_host(*g, *v), &w_outproj));
    FRT2_TRY(upload_f32(bo->data, &b_outproj));
  }
  // ---- UpConv ----
  {
    const HostTensor *w, *b, *wc;
    FRT2_TRY(need(UP + "in_proj.weight", &w, {4 * E, E}));
    FRT2_TRY(need(UP + "in_proj.bias", &b, {4 * E}));
    FRT2_TRY(need(UP + "up_conv.weight", &wc, {4 * E, E, 4}));
    FRT2_TRY(upload_f16(w->data, &w_up_in));
    FRT2_TRY(upload_f32(b->data, &b_up_in));
    // ConvTranspose1d k = s = 4 (in=4E, out=E, k): x50[4t+k, o] = sum_c h[c] Wup[c,o,k]  ->  W[(k*E+o), c]
    std::vector<float> W(static_cast<size_t>(4) * E * 4 * E);
    const int64_t Cin = 4 * E;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < 4LL * E; ++n) {
      const int k = static_cast<int>(n / E), o = static_cast<int>(n % E);
      for (int64_t c = 0; c < Cin; ++c) W[n * Cin + c] = wc->data[(c * E + o) * 4 + k];
    }
    FRT2_TRY(upload_f16(W, &w_up_conv));
    // The RVQ output projection, UpConv.in_proj and the k = s = 4 transposed convolution are three linear maps in a row
    // (rvq.py:163, model.py:142-148: no non-linearity between them), so they compose at load into ONE (4E x rd) matrix:
    //   x50 = Wup (Win (Wo emb + bo) + bin) = (Wup Win Wo) emb + Wup (Win bo + bin)
    // a tenth of the multiply-adds and one launch instead of three; one fp16 rounding of the composed weights instead of
    // three weight roundings and two activation roundings.  fp32 on the host cores (rows in parallel, contiguous inner
    // loops).  FRT2_NO_UPCOMP=1 / DBG_TAPS keep the three separate GEMMs (A/B; the "z" tap only exists there).
    static const bool no_comp = (getenv("FRT2_NO_UPCOMP") != nullptr);
    if (!no_comp) {
      const int64_t E4 = 4 * E, Rd = has_output_proj ? rd : E;
      std::vector<float> Wo_m;                       // (E, rd) or identity
      const float* bo_p = nullptr;
      if (has_output_proj) {
        const HostTensor *g, *v, *bo;
        FRT2_TRY(need(RVQ + "output_proj.parametrizations.weight.original0", &g, {E, 1, 1}));
        FRT2_TRY(need(RVQ + "output_proj.parametrizations.weight.original1", &v, {E, rd, 1}));
        FRT2_TRY(need(RVQ + "output_proj.bias", &bo, {E}));
        Wo_m = weight_norm_host(*g, *v);
        bo_p = bo->data.data();
      }
      // T1 = Win Wo : (4E, Rd);  hb = Win bo + bin : (4E)
      std::vector<float> T1(static_cast<size_t>(E4) * Rd), hb(E4);
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < E4; ++i) {
        const float* wi = &w->data[i * E];
        float* t = &T1[i * Rd];
        double acc_b = b->data[i];
        if (has_output_proj) {
          for (int64_t j = 0; j < Rd; ++j) t[j] = 0.f;
          for (int64_t k = 0; k < E; ++k) {
            const float a = wi[k];
            const float* wo = &Wo_m[k * Rd];
            for (int64_t j = 0; j < Rd; ++j) t[j] += a * wo[j];
            acc_b += static_cast<double>(a) * bo_p[k];
          }
        } else {
          for (int64_t j = 0; j < Rd; ++j) t[j] = wi[j];
        }
        hb[i] = static_cast<float>(acc_b);
      }
      // Wc = Wup_flat T1 : (4E, Rd);  bc = Wup_flat hb
      std::vector<float> Wc(static_cast<size_t>(E4) * Rd), bc(E4);
#pragma omp parallel for schedule(static)
      for (long long n = 0; n < E4; ++n) {
        const float* wu = &W[n * Cin];
        float* o = &Wc[n * Rd];
        for (int64_t j = 0; j < Rd; ++j) o[j] = 0.f;
        double acc_b = 0.0;
        for (int64_t c = 0; c < Cin; ++c) {
          const float a = wu[c];
          const float* t = &T1[c * Rd];
          for (int64_t j = 0; j < Rd; ++j) o[j] += a * t[j];
          acc_b += static_cast<double>(a) * hb[c];
        }
        bc[n] = static_cast<float>(acc_b);
      }
      FRT2_TRY(upload_f16(Wc, &w_up_comp));
      FRT2_TRY(upload_f32(bc, &b_up_comp));
    }
  }
  // ---- upsample_conv (two ConvTranspose1d k=3) ----
  {
    const HostTensor *w0, *b0, *w2, *b2;
    FRT2_TRY(need(AD + "upsample_conv.0.weight", &w0, {E, E, 3}));
    FRT2_TRY(need(AD + "upsample_conv.0.bias", &b0, {E}));
    FRT2_TRY(need(AD + "upsample_conv.2.weight", &w2, {E, E, 3}));
    FRT2_TRY(need(AD + "upsample_conv.2.bias", &b2, {E}));
    // stride-2 polyphase (reference decoder.py:572-579): even y[2t] = W[:,:,0]^T x[t] + W[:,:,2]^T x[t-1],
    // odd y[2t+1] = W[:,:,1]^T x[t].  GEMM over taps (x[t-1], x[t]) with N = 2E: cols [0,E) even, [E,2E) odd.
    std::vector<float> W(static_cast<size_t>(2) * E * 2 * E, 0.f);
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < 2LL * E; ++n) {
      const int o = static_cast<int>(n % E);
      const bool odd = n >= E;
      float* row = &W[static_cast<size_t>(n) * 2 * E];
      for (int c = 0; c < E; ++c) {
        row[c] = odd ? 0.f : w0->data[(static_cast<size_t>(c) * E + o) * 3 + 2];          // tap x[t-1]
        row[E + c] = w0->data[(static_cast<size_t>(c) * E + o) * 3 + (odd ? 1 : 0)];      // tap x[t]
      }
    }
    FRT2_TRY(upload_f16(W, &w_us0));
    std::vector<float> bb(2 * E);
    for (int i = 0; i < 2 * E; ++i) bb[i] = b0->data[i % E];
    FRT2_TRY(upload_f32(bb, &b_us0));
    // stride-1 ConvTranspose == causal conv with flipped taps: v[j] = sum_k W[:,:,k]^T a[j-k]; tap tau <-> a[j-2+tau]
    std::vector<float> W2(static_cast<size_t>(E) * 3 * E);
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < E; ++n)
      for (int tau = 0; tau < 3; ++tau)
        for (int c = 0; c < E; ++c)
          W2[(static_cast<size_t>(n) * 3 + tau) * E + c] = w2->data[(static_cast<size_t>(c) * E + n) * 3 + (2 - tau)];
    FRT2_TRY(upload_f16(W2, &w_us2));
    FRT2_TRY(upload_f32(b2->data, &b_us2));
  }
  // causal Conv1d (out,in,k) -> (out, k*in) tap-major: tap tau <-> x[t-(k-1)+tau]  (reference decoder.py:88-91)
  auto pack_conv = [&](const std::string& key, int k, __half** w, float** b) -> int {
    const HostTensor *wt, *bt;
    FRT2_TRY(need(key + ".weight", &wt, {E, E, k}));
    FRT2_TRY(need(key + ".bias", &bt, {E}));
    std::vector<float> W(static_cast<size_t>(E) * k * E);
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < E; ++n)
      for (int tau = 0; tau < k; ++tau)
        for (int c = 0; c < E; ++c)
          W[(static_cast<size_t>(n) * k + tau) * E + c] = wt->data[(static_cast<size_t>(n) * E + c) * k + tau];
    FRT2_TRY(upload_f16(W, w));
    FRT2_TRY(upload_f32(bt->data, b));
    return FRT2_OK;
  };
  auto up_vec = [&](const std::string& key, int n, float** out) -> int {
    const HostTensor* v;
    FRT2_TRY(need(key, &v, {n}));
    return upload_f32(v->data, out);
  };
  FRT2_TRY(pack_conv(BB + "in_proj", 7, &w_inproj, &b_inproj));
  for (int r = 0; r < 4; ++r) {
    const std::string p = BB + (r < 2 ? "prior_net." : "post_net.") + std::to_string(r % 2) + ".";
    FRT2_TRY(up_vec(p + "block1.1.weight", E, &res[r].ln1_g));
    FRT2_TRY(up_vec(p + "block1.1.bias", E, &res[r].ln1_b));
    FRT2_TRY(pack_conv(p + "block1.4", 3, &res[r].w1, &res[r].b1));
    FRT2_TRY(up_vec(p + "block2.1.weight", E, &res[r].ln2_g));
    FRT2_TRY(up_vec(p + "block2.1.bias", E, &res[r].ln2_b));
    FRT2_TRY(pack_conv(p + "block2.5", 3, &res[r].w2, &res[r].b2));
  }
  layers.resize(nl);
  for (int i = 0; i < nl; ++i) {
    const std::string p = BB + "transformers." + std::to_string(i) + ".";
    LayerW& lw = layers[i];
    const HostTensor *wq, *bq, *wk, *wv, *bv, *wo, *w1, *w2;
    FRT2_TRY(need(p + "self_attn.q_proj.weight", &wq, {E, E}));
    FRT2_TRY(need(p + "self_attn.q_proj.bias", &bq, {E}));
    FRT2_TRY(need(p + "self_attn.k_proj.weight", &wk, {E, E}));
    FRT2_TRY(need(p + "self_attn.v_proj.weight", &wv, {E, E}));
    FRT2_TRY(need(p + "self_attn.v_proj.bias", &bv, {E}));
    FRT2_TRY(need(p + "self_attn.out_proj.weight", &wo, {E, E}));
    FRT2_TRY(need(p + "fc1.weight", &w1, {4 * E, E}));
    FRT2_TRY(need(p + "fc2.weight", &w2, {E, 4 * E}));
    std::vector<float> wqkv(static_cast<size_t>(3) * E * E), bqkv(3 * E, 0.f);  // k_proj has no bias (whisper.py:37)
    std::memcpy(&wqkv[0], wq->data.data(), static_cast<size_t>(E) * E * 4);
    std::memcpy(&wqkv[static_cast<size_t>(E) * E], wk->data.data(), static_cast<size_t>(E) * E * 4);
    std::memcpy(&wqkv[static_cast<size_t>(2) * E * E], wv->data.data(), static_cast<size_t>(E) * E * 4);
    std::memcpy(&bqkv[0], bq->data.data(), E * 4);
    std::memcpy(&bqkv[2 * E], bv->data.data(), E * 4);
    FRT2_TRY(upload_f16(wqkv, &lw.w_qkv));
    FRT2_TRY(upload_f32(bqkv, &lw.b_qkv));
    FRT2_TRY(upload_f16(wo->data, &lw.w_o));
    FRT2_TRY(up_vec(p + "self_attn.out_proj.bias", E, &lw.b_o));
    FRT2_TRY(upload_f16(w1->data, &lw.w_fc1));
    FRT2_TRY(up_vec(p + "fc1.bias", 4 * E, &lw.b_fc1));
    FRT2_TRY(upload_f16(w2->data, &lw.w_fc2));
    FRT2_TRY(up_vec(p + "fc2.bias", E, &lw.b_fc2));
    FRT2_TRY(up_vec(p + "self_attn_layer_norm.weight", E, &lw.ln1_g));
    FRT2_TRY(up_vec(p + "self_attn_layer_norm.bias", E, &lw.ln1_b));
    FRT2_TRY(up_vec(p + "final_layer_norm.weight", E, &lw.ln2_g));
    FRT2_TRY(up_vec(p + "final_layer_norm.bias", E, &lw.ln2_b));
    {
      const HostTensor *g1, *b1, *g2, *b2, *bf1;
      FRT2_TRY(need(p + "self_attn_layer_norm.weight", &g1, {E}));
      FRT2_TRY(need(p + "self_attn_layer_norm.bias", &b1, {E}));
      FRT2_TRY(need(p + "final_layer_norm.weight", &g2, {E}));
      FRT2_TRY(need(p + "final_layer_norm.bias", &b2, {E}));
      FRT2_TRY(need(p + "fc1.bias", &bf1, {4 * E}));
      FRT2_TRY(fold_ln(wqkv.data(), bqkv.data(), g1->data.data(), b1->data.data(), 3 * E, E, &lw.w_qkv_f, &lw.s_qkv,
                       &lw.c_qkv));
      FRT2_TRY(fold_ln(w1->data.data(), bf1->data.data(), g2->data.data(), b2->data.data(), 4 * E, E, &lw.w_fc1_f,
                       &lw.s_fc1, &lw.c_fc1));
    }
  }
  FRT2_TRY(up_vec(BB + "final_norm.weight", E, &fn_g));
  FRT2_TRY(up_vec(BB + "final_norm.bias", E, &fn_b));
  // ---- iSTFT head: Linear(E -> 2*n_bins) with rows interleaved (log-mag f, phase f) so the GEMM epilogue
  //      sees both halves of a bin in one thread (reference decoder.py:503-518) ----
  {
    const HostTensor *w, *b, *win;
    FRT2_TRY(need(AD + "isift.out.weight", &w, {2 * n_bins, E}));
    FRT2_TRY(need(AD + "isift.out.bias", &b, {2 * n_bins}));
    FRT2_TRY(need(AD + "isift.istft.window", &win, {n_fft}));
    std::vector<float> W(static_cast<size_t>(2) * n_bins * E), bb(2 * n_bins);
    for (int f = 0; f < n_bins; ++f) {
      std::memcpy(&W[static_cast<size_t>(2 * f) * E], &w->data[static_cast<size_t>(f) * E], E * 4);
      std::memcpy(&W[static_cast<size_t>(2 * f + 1) * E], &w->data[static_cast<size_t>(n_bins + f) * E], E * 4);
      bb[2 * f] = b->data[f];
      bb[2 * f + 1] = b->data[n_bins + f];
    }
    FRT2_TRY(upload_f16(W, &w_head));
    FRT2_TRY(upload_f32(bb, &b_head));
    {
      const HostTensor *fg, *fb;
      FRT2_TRY(need(BB + "final_norm.weight", &fg, {E}));
      FRT2_TRY(need(BB + "final_norm.bias", &fb, {E}));
      FRT2_TRY(fold_ln(W.data(), bb.data(), fg->data.data(), fb->data.data(), 2 * n_bins, E, &w_head_f, &s_head,
                       &c_head));
    }
    FRT2_TRY(upload_f32(win->data, &window));
    // windowed inverse real DFT as a (n_fft x spec_ld) matrix over interleaved (Re, Im) columns:
    // fr[n] = (1/N) [Re S0 + (-1)^n Re S_{N/2} + 2 sum_f (Re S_f cos(2 pi f n/N) - Im S_f sin(2 pi f n/N))]
    // (irfft, norm="backward"; Im of DC / Nyquist ignored) times window[n] (reference decoder.py:380-381).
    // The 1/N is applied as the GEMM's fp32 alpha so the fp16 basis entries stay in [-2, 2].
    std::vector<float> Bm(static_cast<size_t>(n_fft) * spec_ld, 0.f);
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < n_fft; ++n) {
      const double wn = win->data[n];
      for (int f = 0; f < n_bins; ++f) {
        const bool edge = (f == 0) || (f == n_fft / 2);
        const double cf = edge ? 1.0 : 2.0;
        const double ang = 2.0 * M_PI * static_cast<double>((static_cast<long long>(f) * n) % n_fft) / n_fft;
        Bm[n * spec_ld + 2 * f] = static_cast<float>(wn * cf * std::cos(ang));
        Bm[n * spec_ld + 2 * f + 1] = edge ? 0.f : static_cast<float>(-wn * cf * std::sin(ang));
      }
    }
    FRT2_TRY(upload_f16(Bm, &w_idft));
  }
  // err_word[0]: device error word; err_word[2..3]: item scheduler of the persistent attention kernel (offline arena
  // users are stream-ordered by ws_acquire, so one pair per handle is enough)
  FRT2_TRY(dev_alloc(reinterpret_cast<void**>(&err_word), 16));
  FRT2_CUDA_OK(cudaMemset(err_word, 0, 16));
  raw.clear();
  finalized = true;
  return FRT2_OK;
}

int Handle::ensure_ws(size_t bytes) {
  if (bytes <= ws_bytes) return FRT2_OK;
  if (ws) {
    FRT2_CUDA_OK(cudaDeviceSynchronize());
    FRT2_CUDA_OK(cudaFree(ws));
    ws = nullptr;
    ws_bytes = 0;
  }
  FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ws), bytes));
  ws_bytes = bytes;
  return FRT2_OK;
}

int Handle::tap_f32(const char* name, const float* src, int64_t n, cudaStream_t st) {
  if (!(debug & DBG_TAPS)) return FRT2_OK;
  auto& e = taps[name];
  if (e.second < n) {
    if (e.first) cudaFree(e.first);
    FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&e.first), n * 4));
  }
  e.second = n;
  FRT2_CUDA_OK(cudaMemcpyAsync(e.first, src, n * 4, cudaMemcpyDeviceToDevice, st));
  return FRT2_OK;
}
int Handle::tap_f16(const char* name, const __half* src, int64_t ld, int64_t rows, int cols, cudaStream_t st) {
  if (!(debug & DBG_TAPS)) return FRT2_OK;
  const int64_t n = rows * cols;
  auto& e = taps[name];
  if (e.second < n) {
    if (e.first) cudaFree(e.first);
    FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&e.first), n * 4));
  }
  e.second = n;
  half_to_float_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, ld, rows, cols, e.first);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

// ------------------------------------------------------------------ the decode pipeline
// One function serves both entry points.  Offline (s == nullptr): conv inputs have no history rows and the
// causal left padding comes from TMA out-of-bounds zero fill; attention is block-causal over the whole item.
// Streaming: every conv input buffer is [history | chunk] owned by the stream, K/V are appended to the HBM
// state and attention runs unmasked over state ++ chunk.
size_t Handle::ws_bytes_for(int B, int L) const {
  const size_t R = static_cast<size_t>(B) * L, M = 8 * R;
  const size_t sizes[] = {R * rd * 4, R * rd * 2, R * E * 2, R * 4 * E * 2, R * 4 * E * 2, M * E * 2, M * E * 2,
                          M * E * 4, M * E * 4, M * E * 2, M * 3 * E * 2, M * E * 2, M * 4 * E * 2,
                          M * spec_ld * 2, M * n_fft * 4, M * 8, M * 4};
  size_t off = 0;
  for (size_t b : sizes) off = align_up(off + b, 1024);
  return off;
}

// graph_mode (streaming only): tokens / audio are the stream's staging buffers and every position-dependent quantity
// (K/V append row, attention length, first/last of the iSTFT) is read from the stream's control block in HBM, so the
// recorded launch sequence is identical for every token and can be replayed as one CUDA graph.
int Handle::pipeline(const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B, int nq_in, int L,
                     const int32_t* lengths, float* audio, int64_t audio_pitch, Stream* s, int last, cudaStream_t st,
                     bool graph_mode) {
  const int64_t R = static_cast<int64_t>(B) * L;  // tokens
  const int T50 = 4 * L, T = 8 * L;
  const int64_t M = static_cast<int64_t>(B) * T;  // 100 Hz frames
  const bool streaming = s != nullptr;
  const int pos = streaming ? 8 * s->n_tokens : 0;

  // ---- workspace carve-up ----
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  const size_t o_emb32 = carve(R * rd * 4), o_emb16 = carve(R * rd * 2), o_z16 = carve(R * E * 2);
  const size_t o_h16 = carve(R * 4 * E * 2);
  const size_t o_x50 = carve(R * 4 * E * 2), o_a16 = carve(M * E * 2), o_u16 = carve(M * E * 2);
  const size_t o_x32 = carve(M * E * 4), o_y32 = carve(M * E * 4), o_n16 = carve(M * E * 2);
  const size_t o_qkv = carve(M * 3 * E * 2), o_o16 = carve(M * E * 2), o_g16 = carve(M * 4 * E * 2);
  const size_t o_spec = carve(M * spec_ld * 2), o_frames = carve(M * n_fft * 4);
  const size_t o_stats = carve(M * 8), o_shift = carve(M * 4);
  // The captured per-token step works in the stream's private workspace; everything else in the handle's arena, which
  // the call first acquires against users on other CUDA streams.
  uint8_t* ws = nullptr;
  if (graph_mode) {
    FRT2_TRY(s->ensure_ws(off));
    ws = s->ws;
  } else {
    FRT2_TRY(ensure_ws(off));
    FRT2_TRY(ws_acquire(st));
    ws = this->ws;
  }
  unsigned int* errw = streaming ? s->err_words : err_word;
  float2* stats = reinterpret_cast<float2*>(ws + o_stats);
  float* rowshift = reinterpret_cast<float*>(ws + o_shift);
  float* emb32 = reinterpret_cast<float*>(ws + o_emb32);
  __half* emb16 = reinterpret_cast<__half*>(ws + o_emb16);
  __half* z16 = reinterpret_cast<__half*>(ws + o_z16);
  __half* h16 = reinterpret_cast<__half*>(ws + o_h16);
  float* x32 = reinterpret_cast<float*>(ws + o_x32);
  float* y32 = reinterpret_cast<float*>(ws + o_y32);
  __half* n16 = reinterpret_cast<__half*>(ws + o_n16);
  __half* qkv16 = reinterpret_cast<__half*>(ws + o_qkv);
  __half* o16 = reinterpret_cast<__half*>(ws + o_o16);
  __half* g16 = reinterpret_cast<__half*>(ws + o_g16);
  __half* spec16 = reinterpret_cast<__half*>(ws + o_spec);
  float* frames32 = reinterpret_cast<float*>(ws + o_frames);

  // conv input buffers: pointer to row 0 (first history row), history rows present, batch pitch (elements)
  struct CB { __half* p; int hist; int64_t pitch; };
  CB cb[11];
  if (streaming) {
    for (int i = 0; i < 11; ++i) cb[i] = {s->conv[i], s->conv_hist[i], s->conv_pitch(i)};
  } else {
    cb[0] = {reinterpret_cast<__half*>(ws + o_x50), 0, static_cast<int64_t>(T50) * E};
    cb[1] = {reinterpret_cast<__half*>(ws + o_a16), 0, static_cast<int64_t>(T) * E};
    cb[2] = {reinterpret_cast<__half*>(ws + o_u16), 0, static_cast<int64_t>(T) * E};
    // the resnet-block convolutions read their LN+SiLU'd input from a16 (dead once upsample_conv has run), NOT from
    // n16: with the folded LayerNorm the last conv of a block writes the fp16 copy of its output rows to n16 from its
    // epilogue while later tiles of the same launch are still loading their A rows
    for (int i = 3; i < 11; ++i) cb[i] = {reinterpret_cast<__half*>(ws + o_a16), 0, static_cast<int64_t>(T) * E};
  }

  struct LnFuse { const float* x; const float* g; const float* b; float eps; };
  // With <= 16 rows (the per-token streaming step) the GEMMs run on the skinny weight-streaming kernel, which can
  // compute LayerNorm on the fly from the fp32 residual stream: the separate LN launch (and its round trip) goes.
  static const bool no_lnfuse = (getenv("FRT2_NO_LNFUSE") != nullptr);   // A/B switch for measurements
  const bool fuse_ln = (M <= 16) && !no_lnfuse && !(debug & (DBG_GEMM_REF | DBG_NO_SKINNY | DBG_TAPS));
  // Offline path: LayerNorm folded across the GEMMs (the producer of x32 also emits x16 + row partials into n16 /
  // stats, the consumer runs on x16 with gamma-folded weights) — 25 of the 33 LayerNorm launches disappear; the
  // LN + SiLU in front of the convolutions cannot be folded (non-linearity between LN and the contraction).
  static const bool no_lnfold_env = (getenv("FRT2_NO_LNFOLD") != nullptr);   // A/B switch for measurements
  const bool fold = !streaming && M > 16 && (B == 1 || T > 64) /* producers must not use packed-item tiles */ &&
                    E % 8 == 0 && E <= 2048 && nl > 0 && !no_lnfold_env &&
                    !(debug & (DBG_GEMM_REF | DBG_NO_LNFOLD)) &&
                    (!(debug & DBG_TAPS) || getenv("FRT2_FOLD_WITH_TAPS") != nullptr /* debugging: "final" is then raw */);
  // The fp16 copy is rounded AFTER subtracting the last known mean of its row (LayerNorm is shift-invariant): a
  // checkpoint whose residual rows carry a mean of many times their spread keeps all 11 mantissa bits for the part the
  // normalisation keeps.  rowshift is written by the LN + SiLU kernel that reads x32 in front of every producer chain
  // and advanced by row_stats (true mean = shift + mean of the copy).  FRT2_NO_LNSHIFT=1: plain fp16(x) copy (A/B).
  static const bool no_lnshift_env = (getenv("FRT2_NO_LNSHIFT") != nullptr);
  float* shift = (fold && !no_lnshift_env) ? rowshift : nullptr;
  struct Fold { const __half* W; const float* colsum; const float* bias; float eps; };
  auto run_stats = [&](float eps) -> int {   // (mean, rstd) of the rows of n16 (2 B per element in)
    const int id = prof_begin(FRT2_PROF_LAYER_NORM, 0.0, static_cast<double>(M) * E * 2.0, st);
    const int rc = row_stats(n16, E, M, E, eps, stats, st, shift);
    prof_end(id, st);
    return rc;
  };
  auto flat_gemm = [&](const __half* A, int64_t rows, int Kdim, const __half* W, int N, const float* bias, int act,
                       const float* resid, float* out32, __half* out16, int64_t ld16, float alpha = 1.0f,
                       const LnFuse* ln = nullptr, const Fold* fc = nullptr, bool emit_stats = false) {
    GemmDesc g{};
    g.A = A; g.a_row_pitch = Kdim; g.a_batch_pitch = 0; g.rows_a = static_cast<int>(rows); g.batches = 1;
    g.Kc = Kdim; g.ntaps = 1; g.row_shift = 0; g.W = W; g.N = N; g.rows_out = static_cast<int>(rows);
    g.pitch32 = 0; g.pitch16 = 0; g.alpha = alpha; g.bias = bias; g.act = act; g.resid = resid;
    g.out32 = out32; g.ld32 = N; g.out16 = out16; g.ld16 = ld16;
    if (ln != nullptr) {
      g.ln_x = ln->x; g.ln_ldx = Kdim; g.ln_gamma = ln->g; g.ln_beta = ln->b; g.ln_eps = ln->eps; g.ln_silu = 0;
    }
    if (fc != nullptr) {      // consumer of a folded LayerNorm: A is the raw fp16 residual stream
      FRT2_TRY(run_stats(fc->eps));
      g.W = fc->W; g.bias = fc->bias; g.colsum = fc->colsum; g.stats_in = stats;
    }
    if (emit_stats) {         // producer: fp16 copy of the new residual stream for the next folded LayerNorm
      g.x16_out = n16; g.ld_x16 = E; g.x16_shift = shift;
    }
    return run_gemm(g, st);
  };
  // causal conv over cb[i] with `taps` taps producing `rows` rows per item
  auto conv_gemm = [&](const CB& in, int rows, int taps, const __half* W, int N, const float* bias, int act,
                       const float* resid, float* out32, int64_t pitch32, __half* out16, int64_t ld16,
                       int64_t pitch16, bool emit_stats = false) {
    GemmDesc g{};
    g.A = in.p; g.a_row_pitch = E; g.a_batch_pitch = in.pitch; g.rows_a = in.hist + rows; g.batches = B;
    g.Kc = E; g.ntaps = taps; g.row_shift = in.hist - (taps - 1); g.W = W; g.N = N; g.rows_out = rows;
    g.pitch32 = pitch32; g.pitch16 = pitch16; g.alpha = 1.0f; g.bias = bias; g.act = act; g.resid = resid;
    g.out32 = out32; g.ld32 = E; g.out16 = out16; g.ld16 = ld16;
    if (emit_stats) {
      g.x16_out = n16; g.ld_x16 = E; g.x16_shift = shift;
    }
    return run_gemm(g, st);
  };
  auto chunk_ptr = [&](const CB& c) { return c.p + static_cast<int64_t>(c.hist) * E; };

  // ---- K1: RVQ gather-and-sum (+ output projection) ----
  {
    // algorithmic bytes per token: nq*(idx + D*4) in, D*2 out
    {
      const int id = prof_begin(FRT2_PROF_RVQ, 0.0, static_cast<double>(R) * (nq_in * (idx_bytes + rd * 4.0) + rd * 2.0), st);
      FRT2_TRY(rvq_gather_sum(tokens, idx_bytes, sB, sQ, sL, B, nq_in, L, tables, K, rd,
                              (debug & DBG_TAPS) ? emb32 : nullptr, emb16, nullptr, errw, st));
      prof_end(id, st);
    }
  }
  FRT2_TRY(tap_f32("emb", emb32, R * rd, st));
  if (w_up_comp != nullptr && !(debug & DBG_TAPS)) {
    // ---- output projection + UpConv (Linear E->4E, ConvTranspose k=s=4) as ONE GEMM with the weights composed at load:
    //      its (R, 4E) output is the (4R, E) 50 Hz sequence ----
    GemmDesc g{};
    g.A = emb16; g.a_row_pitch = rd; g.a_batch_pitch = static_cast<int64_t>(L) * rd; g.rows_a = L; g.batches = B;
    g.Kc = rd; g.ntaps = 1; g.row_shift = 0; g.W = w_up_comp; g.N = 4 * E; g.rows_out = L;
    g.alpha = 1.0f; g.bias = b_up_comp; g.act = ACT_NONE; g.resid = nullptr; g.out32 = nullptr; g.ld32 = 0; g.pitch32 = 0;
    g.out16 = chunk_ptr(cb[0]); g.ld16 = 4 * E; g.pitch16 = cb[0].pitch;
    FRT2_TRY(run_gemm(g, st));
  } else {
  const __half* z = emb16;
  if (has_output_proj) {
    FRT2_TRY(flat_gemm(emb16, R, rd, w_outproj, E, b_outproj, ACT_NONE, nullptr, nullptr, z16, E));
    z = z16;
  }
  FRT2_TRY(tap_f16("z", z, E, R, E, st));
  // ---- UpConv: Linear E->4E, ConvTranspose k=s=4 as a GEMM whose (R,4E) output is the (4R,E) 50 Hz sequence ----
  FRT2_TRY(flat_gemm(z, R, E, w_up_in, 4 * E, b_up_in, ACT_NONE, nullptr, nullptr, h16, 4 * E));
  {
    GemmDesc g{};
    g.A = h16; g.a_row_pitch = 4 * E; g.a_batch_pitch = static_cast<int64_t>(L) * 4 * E; g.rows_a = L; g.batches = B;
    g.Kc = 4 * E; g.ntaps = 1; g.row_shift = 0; g.W = w_up_conv; g.N = 4 * E; g.rows_out = L;
    g.alpha = 1.0f; g.bias = nullptr; g.act = ACT_NONE; g.resid = nullptr; g.out32 = nullptr; g.ld32 = 0; g.pitch32 = 0;
    g.out16 = chunk_ptr(cb[0]); g.ld16 = 4 * E; g.pitch16 = cb[0].pitch;
    FRT2_TRY(run_gemm(g, st));
  }
  }
  if (!streaming) FRT2_TRY(tap_f16("x50", cb[0].p, E, R * 4, E, st));
  // ---- upsample_conv: ConvT(k3,s2)+GELU as a 2-tap conv with N=2E (even|odd phases), ConvT(k3,s1)+GELU 3-tap ----
  FRT2_TRY(conv_gemm(cb[0], T50, 2, w_us0, 2 * E, b_us0, ACT_GELU, nullptr, nullptr, 0, chunk_ptr(cb[1]), 2 * E,
                     cb[1].pitch));
  FRT2_TRY(conv_gemm(cb[1], T, 3, w_us2, E, b_us2, ACT_GELU, nullptr, nullptr, 0, chunk_ptr(cb[2]), E, cb[2].pitch));
  if (!streaming) FRT2_TRY(tap_f16("up", cb[2].p, E, M, E, st));
  // ---- backbone: in_proj k7, 2 resnet blocks ----
  const int64_t xp = static_cast<int64_t>(T) * E;
  FRT2_TRY(conv_gemm(cb[2], T, 7, w_inproj, E, b_inproj, ACT_NONE, nullptr, x32, xp, nullptr, 0, 0));
  auto resblock = [&](int r, bool emit_stats = false) -> int {
    const ResW& w = res[r];
    const CB& c1 = cb[3 + 2 * r];
    const CB& c2 = cb[4 + 2 * r];
    FRT2_TRY(run_ln(x32, M, T, w.ln1_g, w.ln1_b, 1e-5f, 1, chunk_ptr(c1), c1.pitch, st, emit_stats ? shift : nullptr));
    FRT2_TRY(conv_gemm(c1, T, 3, w.w1, E, w.b1, ACT_NONE, nullptr, y32, xp, nullptr, 0, 0));
    FRT2_TRY(run_ln(y32, M, T, w.ln2_g, w.ln2_b, 1e-5f, 1, chunk_ptr(c2), c2.pitch, st));
    FRT2_TRY(conv_gemm(c2, T, 3, w.w2, E, w.b2, ACT_NONE, x32, x32, xp, nullptr, 0, 0, emit_stats));
    return FRT2_OK;
  };
  FRT2_TRY(resblock(0));
  FRT2_TRY(resblock(1, fold));           // feeds layer 0's self_attn_layer_norm
  FRT2_TRY(tap_f32("prior", x32, M * E, st));
  // ---- 12 pre-LN transformer layers ----
  for (int i = 0; i < nl; ++i) {
    const LayerW& w = layers[i];
    const LnFuse ln1{x32, w.ln1_g, w.ln1_b, 1e-5f};
    const LnFuse ln2{x32, w.ln2_g, w.ln2_b, 1e-5f};
    const Fold f_qkv{w.w_qkv_f, w.s_qkv, w.c_qkv, 1e-5f};
    const Fold f_fc1{w.w_fc1_f, w.s_fc1, w.c_fc1, 1e-5f};
    if (!fuse_ln && !fold) FRT2_TRY(run_ln(x32, M, static_cast<int>(M), w.ln1_g, w.ln1_b, 1e-5f, 0, n16, 0, st));
    AttnDesc a{};
    a.B = B; a.H = H; a.hd = hd; a.Tq = T; a.out = o16; a.o_row_pitch = E; a.o_batch_pitch = xp;
    a.scale = 1.0f / std::sqrt(static_cast<float>(hd));
    if (!streaming) {
      FRT2_TRY(flat_gemm(n16, M, E, w.w_qkv, 3 * E, w.b_qkv, ACT_NONE, nullptr, nullptr, qkv16, 3 * E, 1.0f,
                         fuse_ln ? &ln1 : nullptr, fold ? &f_qkv : nullptr));
      a.q = qkv16; a.q_row_pitch = 3 * E; a.q_batch_pitch = static_cast<int64_t>(T) * 3 * E;
      a.k = qkv16 + E; a.v = qkv16 + 2 * E; a.kv_row_pitch = 3 * E; a.kv_batch_pitch = a.q_batch_pitch;
      a.Tk = T; a.q_pos0 = 0; a.block_causal = 1;
      a.sched = err_word + 2;
    } else {
      // Q for the chunk; K|V appended in place to the HBM state of this layer (no re-copy: reference
      // whisper.py:100-104 + decoder.py:306 re-concatenate the whole cache every step)
      GemmDesc g{};
      g.A = n16; g.a_row_pitch = E; g.a_batch_pitch = xp; g.rows_a = T; g.batches = B; g.Kc = E; g.ntaps = 1;
      g.row_shift = 0; g.rows_out = T; g.alpha = 1.0f; g.act = ACT_NONE; g.resid = nullptr; g.out32 = nullptr;
      g.ld32 = 0; g.pitch32 = 0;
      if (fuse_ln) {   // global row m = b*T + r is also the row of x32: LayerNorm computed inside the projection
        g.ln_x = x32; g.ln_ldx = E; g.ln_gamma = w.ln1_g; g.ln_beta = w.ln1_b; g.ln_eps = 1e-5f; g.ln_silu = 0;
      }
      __half* kv_dst = graph_mode ? s->kv[i] : s->kv[i] + static_cast<int64_t>(pos) * 2 * E;
      const int* kv_off = graph_mode ? s->ctrl + CTRL_POS : nullptr;
      g.row_off_stride = CTRL_INTS;
      if (fuse_ln) {
        // skinny path: ONE launch for q|k|v, columns >= E routed to the K|V state
        g.W = w.w_qkv; g.N = 3 * E; g.bias = w.b_qkv; g.out16 = qkv16; g.ld16 = E; g.pitch16 = xp;
        g.split_col = E; g.out16_b = kv_dst; g.ld16_b = 2 * E; g.pitch16_b = s->kv_pitch(); g.row_off_b = kv_off;
        FRT2_TRY(run_gemm(g, st));
      } else {
        FRT2_TRY(flat_gemm(n16, M, E, w.w_qkv, E, w.b_qkv, ACT_NONE, nullptr, nullptr, qkv16, E));
        g.W = w.w_qkv + static_cast<int64_t>(E) * E; g.N = 2 * E; g.bias = w.b_qkv + E;
        g.out16 = kv_dst; g.ld16 = 2 * E; g.pitch16 = s->kv_pitch(); g.out_row_off = kv_off;
        FRT2_TRY(run_gemm(g, st));
      }
      a.q = qkv16; a.q_row_pitch = E; a.q_batch_pitch = xp;
      a.k = s->kv[i]; a.v = s->kv[i] + E; a.kv_row_pitch = 2 * E; a.kv_batch_pitch = s->kv_pitch();
      a.Tk = pos + T; a.q_pos0 = pos; a.block_causal = 0;
      if (graph_mode) a.ctrl = s->ctrl;
      a.part = s->attn_part; a.part_count = s->attn_count;   // long K/V state: several CTAs per (item, head)
    }
    FRT2_TRY(run_attn(a, st));
    FRT2_TRY(flat_gemm(o16, M, E, w.w_o, E, w.b_o, ACT_NONE, x32, x32, nullptr, 0, 1.0f, nullptr, nullptr, fold));
    if (!fuse_ln && !fold) FRT2_TRY(run_ln(x32, M, static_cast<int>(M), w.ln2_g, w.ln2_b, 1e-5f, 0, n16, 0, st));
    FRT2_TRY(flat_gemm(n16, M, E, w.w_fc1, 4 * E, w.b_fc1, ACT_GELU, nullptr, nullptr, g16, 4 * E, 1.0f,
                       fuse_ln ? &ln2 : nullptr, fold ? &f_fc1 : nullptr));
    // fc2 feeds the next layer's self_attn_layer_norm (the last layer is followed by an LN + SiLU: not folded)
    FRT2_TRY(flat_gemm(g16, M, 4 * E, w.w_fc2, E, w.b_fc2, ACT_NONE, x32, x32, nullptr, 0, 1.0f, nullptr, nullptr,
                       fold && i + 1 < nl));
    if (i == 0) FRT2_TRY(tap_f32("layer0", x32, M * E, st));
  }
  FRT2_TRY(tap_f32("layers", x32, M * E, st));
  FRT2_TRY(resblock(2));
  FRT2_TRY(resblock(3, fold));             // feeds final_norm
  const LnFuse lnf{x32, fn_g, fn_b, 1e-6f};
  const Fold f_head{w_head_f, s_head, c_head, 1e-6f};
  if (!fuse_ln && !fold) FRT2_TRY(run_ln(x32, M, static_cast<int>(M), fn_g, fn_b, 1e-6f, 0, n16, 0, st));
  FRT2_TRY(tap_f16("final", n16, E, M, E, st));
  // ---- K5: head GEMM with polar epilogue -> windowed inverse DFT GEMM -> overlap-add ----
  if (spec_ld > 2 * n_bins) {
    {
      FRT2_CUDA_OK(cudaMemset2DAsync(spec16 + 2 * n_bins, static_cast<size_t>(spec_ld) * 2, 0,
                                     static_cast<size_t>(spec_ld - 2 * n_bins) * 2, M, st));
    }
  }
  FRT2_TRY(flat_gemm(n16, M, E, w_head, 2 * n_bins, b_head, ACT_POLAR, nullptr, nullptr, spec16, spec_ld, 1.0f,
                     fuse_ln ? &lnf : nullptr, fold ? &f_head : nullptr));
  FRT2_TRY(tap_f16("spec", spec16, spec_ld, M, 2 * n_bins, st));
  FRT2_TRY(flat_gemm(spec16, M, spec_ld, w_idft, n_fft, nullptr, ACT_NONE, nullptr, frames32, nullptr, 0,
                     1.0f / static_cast<float>(n_fft)));
  FRT2_TRY(tap_f32("frames", frames32, M * n_fft, st));
  OlaDesc od{};
  od.frames = frames32; od.frames_batch_pitch = static_cast<int64_t>(T) * n_fft; od.window = window;
  od.lengths = lengths; od.len_mul = 8; od.audio = audio; od.audio_pitch = audio_pitch; od.B = B; od.T = T;
  od.pcm16 = pcm16_out;
  od.out_off = streaming ? nullptr : scatter_off;
  od.n_fft = n_fft; od.hop = hop;
  if (streaming) {
    od.tail = s->tail; od.first = (s->n_tokens == 0); od.last = last;
    if (graph_mode) od.ctrl = s->ctrl;
  } else {
    od.tail = nullptr; od.first = 1; od.last = 1;
  }
  {
    const int id = prof_begin(FRT2_PROF_OLA, 0.0, static_cast<double>(M) * (n_fft + hop) * 4.0, st);
    if (rs_out != nullptr && !streaming)
      FRT2_TRY(istft_overlap_add_resample(od, rs_taps, rs_K, rs_width, rs_orig, rs_new, rs_out, rs_pitch, st));
    else
      FRT2_TRY(istft_overlap_add(od, st));
    prof_end(id, st);
  }
  if (streaming) {
    // end-of-step state roll (new iSTFT tail, conv histories to the head, position advance) as one kernel
    StateRoll ro;
    ro.frames = frames32; ro.frames_batch_pitch = od.frames_batch_pitch; ro.tail = s->tail; ro.T = T; ro.n_fft = n_fft;
    ro.E = E; ro.B = B; ro.ctrl = s->ctrl; ro.advance_frames = T; ro.all_items = graph_mode ? 0 : 1;
    ro.tb = s->shift_table(L);
    FRT2_TRY(stream_state_roll(ro, st));
    launches += 1;
  }
  if (!graph_mode) FRT2_TRY(ws_release(st));
  tap_B = B;
  tap_L = L;
  return FRT2_OK;
}

int Stream::ensure_chunk_cap(int Lc, cudaStream_t st) {
  if (Lc <= chunk_cap) return FRT2_OK;
  // grow the [history | chunk] conv buffers, preserving the history rows
  const int old_cap = chunk_cap;
  __half* old[11];
  int64_t old_pitch[11];
  for (int i = 0; i < 11; ++i) {
    old[i] = conv[i];
    old_pitch[i] = conv_pitch(i);
  }
  chunk_cap = Lc;
  drop_graphs();
  if (old_cap > 0) FRT2_CUDA_OK(cudaStreamSynchronize(st));   // the old staging buffers may still be in use
  if (tok_stage) cudaFree(tok_stage);
  if (audio_stage) cudaFree(audio_stage);
  tok_stage = nullptr;
  audio_stage = nullptr;
  audio_stage_pitch = static_cast<int64_t>(8) * h->hop * Lc + (h->n_fft - h->hop) / 2;
  audio_stage_pitch = (audio_stage_pitch + 3) / 4 * 4;
  FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&tok_stage), static_cast<size_t>(B) * h->nq * Lc * 4));
  FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&audio_stage), static_cast<size_t>(B) * audio_stage_pitch * 4));
  for (int i = 0; i < 11; ++i) {
    const size_t bytes = static_cast<size_t>(B) * conv_pitch(i) * 2;
    FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&conv[i]), bytes));
    FRT2_CUDA_OK(cudaMemsetAsync(conv[i], 0, bytes, st));
    if (old_cap > 0) {
      FRT2_CUDA_OK(cudaMemcpy2DAsync(conv[i], conv_pitch(i) * 2, old[i], old_pitch[i] * 2,
                                     static_cast<size_t>(conv_hist[i]) * h->E * 2, B, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (old_cap > 0) {
    FRT2_CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < 11; ++i) cudaFree(old[i]);
  }
  return FRT2_OK;
}

int Stream::ensure_ws(size_t bytes) {
  if (bytes <= ws_bytes) return FRT2_OK;
  if (ws) {
    FRT2_CUDA_OK(cudaDeviceSynchronize());
    FRT2_CUDA_OK(cudaFree(ws));
    ws = nullptr;
    ws_bytes = 0;
    drop_graphs();   // captured steps point into the old workspace
  }
  FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ws), bytes));
  ws_bytes = bytes;
  return FRT2_OK;
}

void Stream::mark_reset() {
  n_tokens = 0;
  std::fill(slot_tokens.begin(), slot_tokens.end(), 0);
  std::fill(slot_done.begin(), slot_done.end(), 0);
  pending_reset = true;
}

int Stream::begin_use(cudaStream_t st) {
  if (used && st != last_stream) FRT2_CUDA_OK(cudaStreamWaitEvent(st, last_use, 0));
  if (pending_reset) {
    FRT2_TRY(stream_state_reset(shift_table(0), h->E, B, ctrl, err_words, 1 + B, st));
    pending_reset = false;
  }
  return FRT2_OK;
}

int Stream::end_use(cudaStream_t st) {
  if (last_use == nullptr) FRT2_CUDA_OK(cudaEventCreateWithFlags(&last_use, cudaEventDisableTiming));
  FRT2_CUDA_OK(cudaEventRecord(last_use, st));
  last_stream = st;
  used = true;
  return FRT2_OK;
}

}  // namespace frt2

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace frt2;

// The per-token step is captured on a stream of the HIGHEST priority: its kernel nodes inherit that priority, so a
// replay that shares the GPU with a bulk producer (the LLM's next frame on the caller's stream, an offline decode of
// another request) gets its ~90 small CTA waves scheduled ahead of the producer's pending ones instead of queueing
// behind every one of its kernels.  FRT2_STEP_PRIO=0 restores default-priority nodes (A/B).
static cudaError_t create_step_stream(cudaStream_t* out) {
  static const bool prio = !(getenv("FRT2_STEP_PRIO") != nullptr && atoi(getenv("FRT2_STEP_PRIO")) == 0);
  int least = 0, greatest = 0;
  if (prio && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess)
    return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, greatest);
  return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}

struct frt2_handle { Handle h; };
struct frt2_stream { Stream s; };

frt2::Handle::~Handle() {
  cudaSetDevice(device);
  for (frt2_stream* fs : free_streams) delete fs;
  release_base();
}

extern "C" {

const char* frt2_last_error(void) { return get_error(); }
const char* frt2_version(void) { return "frt2_b200 0.1 (sm_100a)"; }

int frt2_create(const frt2_config* cfg, int device, frt2_handle** out) {
  FRT2_REQUIRE(cfg != nullptr && out != nullptr, FRT2_ERR_BAD_ARG, "frt2_create: null argument");
  FRT2_REQUIRE(cfg->output_dim == cfg->embed_dim, FRT2_ERR_BAD_ARG, "rvq.output_dim must equal embed_dim");
  FRT2_REQUIRE(cfg->upconv_stride == 4, FRT2_ERR_BAD_ARG, "only UpConv stride 4 is supported");
  FRT2_REQUIRE(cfg->embed_dim % 64 == 0 && cfg->rvq_dim % 64 == 0, FRT2_ERR_BAD_ARG,
               "embed_dim and rvq_dim must be multiples of 64 (tensor-core K blocks)");
  FRT2_REQUIRE(cfg->codebook_dim % 4 == 0, FRT2_ERR_BAD_ARG, "codebook_dim must be a multiple of 4");
  FRT2_REQUIRE(cfg->embed_dim % cfg->num_heads == 0, FRT2_ERR_BAD_ARG, "embed_dim must be divisible by num_heads");
  const int hd = cfg->embed_dim / cfg->num_heads;
  FRT2_REQUIRE(hd == 32 || hd == 64 || hd == 128, FRT2_ERR_BAD_ARG, "head_dim must be 32, 64 or 128");
  FRT2_REQUIRE(cfg->num_quantizers >= 1 && cfg->num_quantizers <= 64, FRT2_ERR_BAD_ARG, "num_quantizers out of range");
  FRT2_REQUIRE(cfg->hop_length >= 1 && cfg->num_layers >= 0 && cfg->codebook_size >= 1, FRT2_ERR_BAD_ARG,
               "bad config");
  int ndev = 0;
  FRT2_CUDA_OK(cudaGetDeviceCount(&ndev));
  FRT2_REQUIRE(device >= 0 && device < ndev, FRT2_ERR_BAD_ARG, "frt2_create: no such CUDA device");
  FRT2_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  FRT2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  FRT2_REQUIRE(prop.major == 10, FRT2_ERR_CUDA,
               "libfrt2_b200 needs a Blackwell sm_100 device (tcgen05/TMEM/TMA); there is no fallback path");
  auto* w = new frt2_handle();
  Handle& h = w->h;
  h.cfg = *cfg;
  h.device = device;
  h.E = cfg->embed_dim; h.H = cfg->num_heads; h.hd = hd; h.nl = cfg->num_layers; h.nq = cfg->num_quantizers;
  h.K = cfg->codebook_size; h.cd = cfg->codebook_dim; h.rd = cfg->rvq_dim; h.hop = cfg->hop_length;
  h.n_fft = 4 * cfg->hop_length; h.n_bins = h.n_fft / 2 + 1;
  h.spec_ld = (2 * h.n_bins + 63) / 64 * 64;
  h.has_out_project = cfg->codebook_dim != cfg->rvq_dim;
  h.has_output_proj = cfg->rvq_dim != cfg->output_dim;
  *out = w;
  return FRT2_OK;
}

int frt2_load_tensor(frt2_handle* hh, const char* key, const float* data, int ndim, const int64_t* shape,
                     int on_device) {
  FRT2_REQUIRE(hh && key && data && shape && ndim >= 1 && ndim <= 4, FRT2_ERR_BAD_ARG, "frt2_load_tensor: bad argument");
  Handle& h = hh->h;
  FRT2_REQUIRE(!h.finalized, FRT2_ERR_BAD_ARG, "frt2_load_tensor: handle already finalized");
  const std::string k(key);
  // training-only buffers are not part of the path (in_project / input_proj feed frt2_rvq_encode)
  for (const char* skip : {"inited", "cluster_size", "embed_avg"})
    if (k.find(skip) != std::string::npos) return FRT2_OK;
  if (k.rfind("rvq.", 0) != 0 && k.rfind("upsample.", 0) != 0 && k.rfind("acoustic_decoder.", 0) != 0) return FRT2_OK;
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  FRT2_REQUIRE(n >= 0, FRT2_ERR_BAD_ARG, "frt2_load_tensor: negative dimension");
  t.data.resize(n);
  if (on_device) {
    FRT2_CUDA_OK(cudaSetDevice(h.device));
    FRT2_CUDA_OK(cudaMemcpy(t.data.data(), data, n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::memcpy(t.data.data(), data, n * 4);
  }
  std::lock_guard<std::mutex> lk(h.mu);
  h.raw[k] = std::move(t);
  return FRT2_OK;
}

int frt2_finalize(frt2_handle* hh) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  std::lock_guard<std::mutex> lk(hh->h.mu);
  FRT2_REQUIRE(!hh->h.finalized, FRT2_ERR_BAD_ARG, "already finalized");
  return hh->h.finalize();
}

void frt2_destroy(frt2_handle* hh) { delete hh; }

static int check_decode_args(Handle& h, const void* tokens, int idx_bytes, int B, int nq, int L, const float* audio) {
  FRT2_REQUIRE(h.finalized, FRT2_ERR_NOT_FINALIZED, "handle not finalized");
  FRT2_REQUIRE(tokens != nullptr && audio != nullptr, FRT2_ERR_BAD_ARG, "null tokens/audio pointer");
  FRT2_REQUIRE(idx_bytes == 4 || idx_bytes == 8, FRT2_ERR_BAD_DTYPE, "tokens must be int32 or int64");
  FRT2_REQUIRE(B >= 1 && L >= 1, FRT2_ERR_BAD_ARG, "B and L must be >= 1");
  FRT2_REQUIRE(nq >= 1 && nq <= h.nq, FRT2_ERR_BAD_ARG, "nq must be in [1, num_quantizers] (quantizers[:nq], rvq.py:160)");
  return FRT2_OK;
}

int frt2_decode(frt2_handle* hh, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B, int nq,
                int L, const int32_t* lengths, float* audio, int64_t audio_pitch, void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_TRY(check_decode_args(h, tokens, idx_bytes, B, nq, L, audio));
  FRT2_REQUIRE(audio_pitch >= static_cast<int64_t>(8) * h.hop * L, FRT2_ERR_BAD_ARG, "audio_pitch too small");
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  return h.pipeline(tokens, idx_bytes, sB, sQ, sL, B, nq, L, lengths, audio, audio_pitch, nullptr, 1,
                    static_cast<cudaStream_t>(cuda_stream));
}

int frt2_decode_pcm16(frt2_handle* hh, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                      int nq, int L, const int32_t* lengths, int16_t* pcm, int64_t pcm_pitch, void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_TRY(check_decode_args(h, tokens, idx_bytes, B, nq, L, reinterpret_cast<const float*>(pcm)));
  FRT2_REQUIRE(pcm_pitch >= static_cast<int64_t>(8) * h.hop * L, FRT2_ERR_BAD_ARG, "pcm_pitch too small");
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  h.pcm16_out = pcm;
  const int rc = h.pipeline(tokens, idx_bytes, sB, sQ, sL, B, nq, L, lengths, nullptr, pcm_pitch, nullptr, 1,
                            static_cast<cudaStream_t>(cuda_stream));
  h.pcm16_out = nullptr;
  return rc;
}

int frt2_decode_scatter(frt2_handle* hh, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                        int nq, int L, const int32_t* lengths, void* out_base, int out_pcm16, const int64_t* out_off,
                        void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_TRY(check_decode_args(h, tokens, idx_bytes, B, nq, L, reinterpret_cast<const float*>(out_base)));
  FRT2_REQUIRE(out_off != nullptr, FRT2_ERR_BAD_ARG, "frt2_decode_scatter: null out_off");
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  h.scatter_off = reinterpret_cast<const long long*>(out_off);
  h.pcm16_out = out_pcm16 ? static_cast<int16_t*>(out_base) : nullptr;
  const int rc = h.pipeline(tokens, idx_bytes, sB, sQ, sL, B, nq, L, lengths, out_pcm16 ? nullptr : static_cast<float*>(out_base),
                            0, nullptr, 1, static_cast<cudaStream_t>(cuda_stream));
  h.scatter_off = nullptr;
  h.pcm16_out = nullptr;
  return rc;
}

// ---- peer memory: a waveform buffer on one GPU that the other ranks' overlap-add kernels write into ----
int frt2_peer_alloc(int device, int64_t bytes, void** ptr, unsigned char* handle) {
  FRT2_REQUIRE(ptr && handle && bytes > 0, FRT2_ERR_BAD_ARG, "frt2_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == FRT2_PEER_HANDLE_BYTES, "IPC handle size");
  FRT2_CUDA_OK(cudaSetDevice(device));
  void* p = nullptr;
  FRT2_CUDA_OK(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t hd;
  const cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    FRT2_CUDA_OK(e);
  }
  std::memcpy(handle, &hd, sizeof(hd));
  *ptr = p;
  return FRT2_OK;
}

int frt2_peer_open(int device, const unsigned char* handle, void** ptr) {
  FRT2_REQUIRE(ptr && handle, FRT2_ERR_BAD_ARG, "frt2_peer_open: bad argument");
  FRT2_CUDA_OK(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, sizeof(hd));
  void* p = nullptr;
  FRT2_CUDA_OK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return FRT2_OK;
}

int frt2_peer_close(int device, void* ptr) {
  FRT2_REQUIRE(ptr, FRT2_ERR_BAD_ARG, "frt2_peer_close: null pointer");
  FRT2_CUDA_OK(cudaSetDevice(device));
  FRT2_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return FRT2_OK;
}

int frt2_peer_free(int device, void* ptr) {
  if (ptr == nullptr) return FRT2_OK;
  FRT2_CUDA_OK(cudaSetDevice(device));
  FRT2_CUDA_OK(cudaFree(ptr));
  return FRT2_OK;
}

// Spare stream states kept by the handle (see Handle::free_streams): at most this many / this many bytes.
static constexpr size_t STREAM_POOL_MAX = 16;
static constexpr size_t STREAM_POOL_MAX_BYTES = static_cast<size_t>(8) << 30;

static int stream_create_fresh(Handle& h, int B, int max_tokens, frt2_stream** out) {
  auto* w = new frt2_stream();
  Stream& s = w->s;
  s.h = &h; s.B = B; s.max_tokens = max_tokens;
  s.kv.resize(h.nl, nullptr);
  int st = FRT2_OK;
  auto fail = [&](int code) { delete w; return code; };
  for (int i = 0; i < h.nl; ++i) {
    if (cudaMalloc(reinterpret_cast<void**>(&s.kv[i]), static_cast<size_t>(B) * s.kv_pitch() * 2) != cudaSuccess) {
      set_error("frt2_stream_create: out of memory for the KV state");
      return fail(FRT2_ERR_CUDA);
    }
  }
  if (cudaMalloc(reinterpret_cast<void**>(&s.tail), static_cast<size_t>(B) * 3 * h.n_fft * 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&s.ctrl), static_cast<size_t>(B) * CTRL_INTS * sizeof(int)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&s.err_words), static_cast<size_t>(1 + B) * sizeof(unsigned int)) != cudaSuccess ||
      create_step_stream(&s.cap_stream) != cudaSuccess) {
    set_error("frt2_stream_create: out of memory");
    return fail(FRT2_ERR_CUDA);
  }
  if (h.E / h.H == 64 && static_cast<long long>(B) * h.H <= 1024) {
    const size_t pf = static_cast<size_t>(B) * h.H * ATTN_KSPLIT * 8 * (64 + 2) * sizeof(float);
    if (cudaMalloc(reinterpret_cast<void**>(&s.attn_part), pf) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&s.attn_count), static_cast<size_t>(B) * h.H * sizeof(int)) != cudaSuccess ||
        cudaMemset(s.attn_count, 0, static_cast<size_t>(B) * h.H * sizeof(int)) != cudaSuccess) {
      set_error("frt2_stream_create: out of memory for the attention split workspace");
      return fail(FRT2_ERR_CUDA);
    }
  }
  st = s.ensure_chunk_cap(1, nullptr);
  if (st != FRT2_OK) return fail(st);
  st = s.ensure_ws(h.ws_bytes_for(B, 1));
  if (st != FRT2_OK) return fail(st);
  s.mark_reset();   // the reset kernel runs on the stream of the first decode call
  // everything above ran on the null stream (allocation-time memsets): make it visible to any stream
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("frt2_stream_create: device synchronisation failed");
    return fail(FRT2_ERR_CUDA);
  }
  *out = w;
  return FRT2_OK;
}

int frt2_stream_create(frt2_handle* hh, int B, int max_tokens, frt2_stream** out) {
  FRT2_REQUIRE(hh && out, FRT2_ERR_BAD_ARG, "null argument");
  Handle& h = hh->h;
  FRT2_REQUIRE(h.finalized, FRT2_ERR_NOT_FINALIZED, "handle not finalized");
  FRT2_REQUIRE(B >= 1 && max_tokens >= 1, FRT2_ERR_BAD_ARG, "B and max_tokens must be >= 1");
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  {
    // a spare state of this shape (returned by frt2_stream_destroy, or made by frt2_stream_reserve): no allocation, no
    // graph capture — its reset is a kernel on the stream of the first decode call
    std::lock_guard<std::mutex> lk(h.mu);
    for (size_t i = 0; i < h.free_streams.size(); ++i) {
      Stream& c = h.free_streams[i]->s;
      if (c.B == B && c.max_tokens == max_tokens && !c.pooled) {
        *out = h.free_streams[i];
        h.free_stream_bytes -= c.footprint();
        h.free_streams.erase(h.free_streams.begin() + i);
        c.mark_reset();
        return FRT2_OK;
      }
    }
  }
  return stream_create_fresh(h, B, max_tokens, out);
}

static int capture_step(Handle& h, Stream& s, int nq, int Lc);

int frt2_stream_reserve(frt2_handle* hh, int B, int max_tokens, int count) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_REQUIRE(h.finalized, FRT2_ERR_NOT_FINALIZED, "handle not finalized");
  FRT2_REQUIRE(B >= 1 && max_tokens >= 1 && count >= 0, FRT2_ERR_BAD_ARG, "frt2_stream_reserve: bad argument");
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  for (int i = 0; i < count; ++i) {
    {
      std::lock_guard<std::mutex> lk(h.mu);
      int have = 0;
      for (frt2_stream* fs : h.free_streams) have += (fs->s.B == B && fs->s.max_tokens == max_tokens && !fs->s.pooled);
      if (have >= count || h.free_streams.size() >= STREAM_POOL_MAX) return FRT2_OK;
    }
    frt2_stream* fs = nullptr;
    FRT2_TRY(stream_create_fresh(h, B, max_tokens, &fs));
    std::lock_guard<std::mutex> lk(h.mu);
    if (B * 8 < 32) {   // the per-token step of this shape, captured now instead of inside the first request
      const int rc = capture_step(h, fs->s, h.nq, 1);
      if (rc != FRT2_OK) {
        delete fs;
        return rc;
      }
    }
    h.free_stream_bytes += fs->s.footprint();
    h.free_streams.push_back(fs);
  }
  return FRT2_OK;
}

int frt2_stream_reset(frt2_stream* ss) {
  FRT2_REQUIRE(ss, FRT2_ERR_BAD_ARG, "null stream");
  std::lock_guard<std::mutex> lk(ss->s.h->mu);
  ss->s.mark_reset();   // stream-ordered: the reset kernel runs on the stream of the next decode call
  return FRT2_OK;
}

void frt2_stream_destroy(frt2_stream* ss) {
  if (ss == nullptr) return;
  Handle& h = *ss->s.h;
  {
    std::lock_guard<std::mutex> lk(h.mu);
    const size_t fp = ss->s.footprint();
    if (!ss->s.pooled && h.free_streams.size() < STREAM_POOL_MAX && h.free_stream_bytes + fp <= STREAM_POOL_MAX_BYTES) {
      ss->s.mark_reset();
      h.free_stream_bytes += fp;
      h.free_streams.push_back(ss);
      return;
    }
  }
  cudaSetDevice(h.device);
  cudaDeviceSynchronize();   // kernels of the last request may still be using the state
  delete ss;
}
int frt2_stream_tokens(const frt2_stream* ss) { return ss ? ss->s.n_tokens : 0; }

int frt2_stream_check_error(frt2_handle* hh, frt2_stream* ss, int32_t* item_flags, void* cuda_stream) {
  FRT2_REQUIRE(hh && ss, FRT2_ERR_BAD_ARG, "null handle/stream");
  Handle& h = hh->h;
  Stream& s = ss->s;
  FRT2_REQUIRE(s.h == &h, FRT2_ERR_BAD_ARG, "stream belongs to another handle");
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::vector<unsigned int> w(1 + s.B, 0u);
  FRT2_CUDA_OK(cudaMemcpyAsync(w.data(), s.err_words, w.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  FRT2_CUDA_OK(cudaStreamSynchronize(st));
  if (item_flags != nullptr)
    for (int b = 0; b < s.B; ++b) item_flags[b] = static_cast<int32_t>(w[1 + b]);
  if (w[0] != 0) {
    FRT2_CUDA_OK(cudaMemsetAsync(s.err_words, 0, w.size() * sizeof(unsigned int), st));
    if (w[0] & DEV_ERR_INDEX_OOR) {
      set_error("index out of range in self");
      return FRT2_ERR_INDEX_OUT_OF_RANGE;
    }
  }
  return FRT2_OK;
}

int frt2_stream_fetch_errors(frt2_handle* hh, frt2_stream* ss, uint32_t* host_words, void* cuda_stream) {
  FRT2_REQUIRE(hh && ss && host_words, FRT2_ERR_BAD_ARG, "null handle/stream/buffer");
  FRT2_REQUIRE(ss->s.h == &hh->h, FRT2_ERR_BAD_ARG, "stream belongs to another handle");
  FRT2_CUDA_OK(cudaSetDevice(hh->h.device));
  FRT2_CUDA_OK(cudaMemcpyAsync(host_words, ss->s.err_words, static_cast<size_t>(1 + ss->s.B) * sizeof(unsigned int),
                               cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(cuda_stream)));
  return FRT2_OK;
}

// Capture the control-block-driven step for (Lc, nq) on the stream's private workspace (h.mu held by the caller).
static int capture_step(Handle& h, Stream& s, int nq, int Lc) {
  auto key = std::make_pair(Lc, nq);
  if (s.graphs.count(key)) return FRT2_OK;
  FRT2_TRY(s.ensure_chunk_cap(Lc, nullptr));
  FRT2_TRY(s.ensure_ws(h.ws_bytes_for(s.B, Lc)));  // before capture: no allocation may happen while recording
  const long long before = h.launches;
  FRT2_CUDA_OK(cudaStreamBeginCapture(s.cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = h.pipeline(s.tok_stage, 4, static_cast<int64_t>(nq) * Lc, Lc, 1, s.B, nq, Lc, nullptr, s.audio_stage,
                            s.audio_stage_pitch, &s, 0, s.cap_stream, true);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(s.cap_stream, &graph);
  if (rc != FRT2_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  FRT2_CUDA_OK(ce);
  Stream::GraphRec rec{};
  rec.kernels = h.launches - before;
  h.launches = before;
  const cudaError_t ie = cudaGraphInstantiate(&rec.exec, graph, 0);
  cudaGraphDestroy(graph);
  FRT2_CUDA_OK(ie);
  s.graphs.emplace(key, rec);
  return FRT2_OK;
}

// Run the control-block-driven step for (Lc, nq): tokens come from the stream's staging buffer, audio goes to its
// staging buffer, positions / flags are read from HBM.  Normally one CUDA-graph replay (captured on first use).
static int run_ctrl_step(Handle& h, Stream& s, int nq, int Lc, cudaStream_t st, bool use_graph) {
  if (!use_graph)
    return h.pipeline(s.tok_stage, 4, static_cast<int64_t>(nq) * Lc, Lc, 1, s.B, nq, Lc, nullptr, s.audio_stage,
                      s.audio_stage_pitch, &s, 0, st, true);
  FRT2_TRY(capture_step(h, s, nq, Lc));
  const Stream::GraphRec& g = s.graphs.at(std::make_pair(Lc, nq));
  FRT2_CUDA_OK(cudaGraphLaunch(g.exec, st));
  h.launches += g.kernels;
  return FRT2_OK;
}

// staged chunk -> caller's buffer (fp32 copy or int16 PCM conversion), n samples per item
static int emit_chunk(Handle& h, Stream& s, float* audio, int16_t* pcm, int64_t pitch, int n, cudaStream_t st) {
  if (pcm != nullptr) {
    emit_pcm16_kernel<<<dim3((n + 255) / 256, s.B), 256, 0, st>>>(s.audio_stage, s.audio_stage_pitch, pcm, pitch, n, s.B);
    FRT2_CUDA_OK(cudaGetLastError());
  } else {
    FRT2_CUDA_OK(cudaMemcpy2DAsync(audio, static_cast<size_t>(pitch) * 4, s.audio_stage,
                                   static_cast<size_t>(s.audio_stage_pitch) * 4, static_cast<size_t>(n) * 4, s.B,
                                   cudaMemcpyDeviceToDevice, st));
  }
  ++h.launches;
  return FRT2_OK;
}

static int decode_chunk_impl(frt2_handle* hh, frt2_stream* ss, const void* tokens, int idx_bytes, int64_t sB,
                             int64_t sQ, int64_t sL, int nq, int Lc, int last, float* audio, int16_t* pcm,
                             int64_t audio_pitch, int* n_samples, void* cuda_stream) {
  FRT2_REQUIRE(hh && ss, FRT2_ERR_BAD_ARG, "null handle/stream");
  Handle& h = hh->h;
  Stream& s = ss->s;
  FRT2_REQUIRE(s.h == &h, FRT2_ERR_BAD_ARG, "stream belongs to another handle");
  FRT2_REQUIRE(!s.pooled, FRT2_ERR_BAD_ARG, "this is a slot pool: use frt2_pool_step");
  FRT2_TRY(check_decode_args(h, tokens, idx_bytes, s.B, nq, Lc,
                             pcm ? reinterpret_cast<const float*>(pcm) : audio));
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_REQUIRE(s.n_tokens + Lc <= s.max_tokens, FRT2_ERR_STATE_OVERFLOW,
               "stream state overflow: more tokens than frt2_stream_create reserved");
  const int pad = (h.n_fft - h.hop) / 2;
  const int n = 8 * h.hop * Lc - (s.n_tokens == 0 ? pad : 0) + (last ? pad : 0);
  FRT2_REQUIRE(audio_pitch >= n, FRT2_ERR_BAD_ARG, "audio_pitch too small");
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  FRT2_TRY(s.ensure_chunk_cap(Lc, st));
  FRT2_TRY(s.begin_use(st));   // ordered after the state's previous user; pending reset applied on this stream
  // Short chunks (the per-token latency path) replay one captured CUDA graph per call; anything that needs host-side
  // parameters per kernel (taps, per-kernel event timing, the tcgen05 attention for long chunks) runs kernel by kernel.
  const bool graph_mode = !(h.debug & (DBG_NO_GRAPH | DBG_TAPS)) && !h.profile && 8 * Lc < 32;
  if (!graph_mode) {
    h.pcm16_out = pcm;
    const int rc = h.pipeline(tokens, idx_bytes, sB, sQ, sL, s.B, nq, Lc, nullptr, audio, audio_pitch, &s, last, st,
                              false);
    h.pcm16_out = nullptr;
    FRT2_TRY(rc);
  } else {
    const int ntok = s.B * nq * Lc;
    const int nthr = std::max(ntok, s.B);
    const SlotFlags none{};
    if (idx_bytes == 4) {
      stage_tokens_kernel<int, false><<<(nthr + 127) / 128, 128, 0, st>>>(
          static_cast<const int*>(tokens), sB, sQ, sL, s.B, nq, Lc, h.K, s.tok_stage, s.ctrl, last, s.err_words, none);
    } else {
      stage_tokens_kernel<long long, false><<<(nthr + 127) / 128, 128, 0, st>>>(
          static_cast<const long long*>(tokens), sB, sQ, sL, s.B, nq, Lc, h.K, s.tok_stage, s.ctrl, last, s.err_words,
          none);
    }
    FRT2_CUDA_OK(cudaGetLastError());
    ++h.launches;
    FRT2_TRY(run_ctrl_step(h, s, nq, Lc, st, true));
    FRT2_TRY(emit_chunk(h, s, audio, pcm, audio_pitch, n, st));
  }
  FRT2_TRY(s.end_use(st));
  s.n_tokens += Lc;
  if (n_samples) *n_samples = n;
  return FRT2_OK;
}

int frt2_decode_chunk(frt2_handle* hh, frt2_stream* ss, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ,
                      int64_t sL, int nq, int Lc, int last, float* audio, int64_t audio_pitch, int* n_samples,
                      void* cuda_stream) {
  return decode_chunk_impl(hh, ss, tokens, idx_bytes, sB, sQ, sL, nq, Lc, last, audio, nullptr, audio_pitch, n_samples,
                           cuda_stream);
}

int frt2_decode_chunk_pcm16(frt2_handle* hh, frt2_stream* ss, const void* tokens, int idx_bytes, int64_t sB,
                            int64_t sQ, int64_t sL, int nq, int Lc, int last, int16_t* pcm, int64_t pcm_pitch,
                            int* n_samples, void* cuda_stream) {
  return decode_chunk_impl(hh, ss, tokens, idx_bytes, sB, sQ, sL, nq, Lc, last, nullptr, pcm, pcm_pitch, n_samples,
                           cuda_stream);
}

// ---- slot pool ----
int frt2_pool_create(frt2_handle* hh, int slots, int max_tokens, frt2_stream** out) {
  FRT2_REQUIRE(slots >= 1 && slots <= FRT2_POOL_MAX_SLOTS, FRT2_ERR_BAD_ARG,
               "frt2_pool_create: slots must be in [1, FRT2_POOL_MAX_SLOTS]");
  FRT2_TRY(frt2_stream_create(hh, slots, max_tokens, out));
  Stream& s = (*out)->s;
  s.pooled = true;
  s.slot_tokens.assign(slots, 0);
  s.slot_done.assign(slots, 0);
  return FRT2_OK;
}

int frt2_pool_slot_tokens(const frt2_stream* ss, int slot) {
  if (!ss || !ss->s.pooled || slot < 0 || slot >= ss->s.B) return 0;
  return ss->s.slot_tokens[slot];
}

int frt2_pool_step(frt2_handle* hh, frt2_stream* ss, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ,
                   int nq, const int32_t* slot_flags, void* out, int out_pcm16, int64_t out_pitch,
                   int32_t* n_samples, void* cuda_stream) {
  FRT2_REQUIRE(hh && ss && slot_flags, FRT2_ERR_BAD_ARG, "null handle/pool/flags");
  Handle& h = hh->h;
  Stream& s = ss->s;
  FRT2_REQUIRE(s.h == &h, FRT2_ERR_BAD_ARG, "pool belongs to another handle");
  FRT2_REQUIRE(s.pooled, FRT2_ERR_BAD_ARG, "not a slot pool (frt2_pool_create)");
  FRT2_TRY(check_decode_args(h, tokens, idx_bytes, s.B, nq, 1, static_cast<const float*>(out)));
  const int pad = (h.n_fft - h.hop) / 2;
  const int width = 8 * h.hop + pad;
  FRT2_REQUIRE(out_pitch >= width, FRT2_ERR_BAD_ARG, "out_pitch must be >= 8*hop + pad");
  std::lock_guard<std::mutex> lk(h.mu);
  // validate the whole step before touching any state
  SlotFlags flags{};
  bool any_reset = false;
  for (int b = 0; b < s.B; ++b) {
    const int f = slot_flags[b];
    FRT2_REQUIRE((f & ~(FRT2_SLOT_ACTIVE | FRT2_SLOT_LAST | FRT2_SLOT_RESET)) == 0, FRT2_ERR_BAD_ARG,
                 "frt2_pool_step: unknown slot flag");
    if (!(f & FRT2_SLOT_ACTIVE)) {
      FRT2_REQUIRE(f == 0, FRT2_ERR_BAD_ARG, "frt2_pool_step: RESET / LAST need ACTIVE");
      continue;
    }
    const int consumed = (f & FRT2_SLOT_RESET) ? 0 : s.slot_tokens[b];
    FRT2_REQUIRE((f & FRT2_SLOT_RESET) || !s.slot_done[b], FRT2_ERR_BAD_ARG,
                 "frt2_pool_step: the slot's stream has ended (LAST): set FRT2_SLOT_RESET to start a new one");
    FRT2_REQUIRE(consumed + 1 <= s.max_tokens, FRT2_ERR_STATE_OVERFLOW,
                 "pool slot overflow: more tokens than frt2_pool_create reserved");
    flags.f[b] = static_cast<unsigned char>(f);
    any_reset = any_reset || (f & FRT2_SLOT_RESET);
  }
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  FRT2_TRY(s.begin_use(st));
  if (any_reset) {
    reset_history_kernel<<<dim3(8, 11), 256, 0, st>>>(s.shift_table(0), h.E, s.B, flags);
    ++h.launches;
  }
  const int nthr = s.B * nq;
  if (idx_bytes == 4) {
    stage_tokens_kernel<int, true><<<(nthr + 127) / 128, 128, 0, st>>>(
        static_cast<const int*>(tokens), sB, sQ, 0, s.B, nq, 1, h.K, s.tok_stage, s.ctrl, 0, s.err_words, flags);
  } else {
    stage_tokens_kernel<long long, true><<<(nthr + 127) / 128, 128, 0, st>>>(
        static_cast<const long long*>(tokens), sB, sQ, 0, s.B, nq, 1, h.K, s.tok_stage, s.ctrl, 0, s.err_words, flags);
  }
  FRT2_CUDA_OK(cudaGetLastError());
  ++h.launches;
  FRT2_TRY(run_ctrl_step(h, s, nq, 1, st, !(h.debug & DBG_NO_GRAPH) && !h.profile));
  FRT2_TRY(emit_chunk(h, s, out_pcm16 ? nullptr : static_cast<float*>(out),
                      out_pcm16 ? static_cast<int16_t*>(out) : nullptr, out_pitch, width, st));
  FRT2_TRY(s.end_use(st));
  for (int b = 0; b < s.B; ++b) {
    const int f = flags.f[b];
    int n = 0;
    if (f & FRT2_SLOT_ACTIVE) {
      if (f & FRT2_SLOT_RESET) s.slot_tokens[b] = 0;
      n = 8 * h.hop - (s.slot_tokens[b] == 0 ? pad : 0) + ((f & FRT2_SLOT_LAST) ? pad : 0);
      s.slot_tokens[b] += 1;
      s.slot_done[b] = (f & FRT2_SLOT_LAST) ? 1 : 0;
    }
    if (n_samples) n_samples[b] = n;
  }
  return FRT2_OK;
}

int frt2_export_state(frt2_handle* hh, const frt2_stream* ss, float* up_conv_cache, float* bb_conv_cache1,
                      float* bb_conv_cache2, float* bb_kv_cache, float* is_cache, void* cuda_stream) {
  FRT2_REQUIRE(hh && ss, FRT2_ERR_BAD_ARG, "null handle/stream");
  Handle& h = hh->h;
  Stream& s = const_cast<Stream&>(ss->s);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  FRT2_TRY(s.begin_use(st));
  const int E = h.E, B = s.B;
  if (up_conv_cache) {  // (B,E,3) = [last x50 frame | last two post-GELU frames]  (decoder.py:624-655)
    export_tm_to_cm_kernel<<<grid_for(1LL * B * E), 256, 0, st>>>(s.conv[0], s.conv_pitch(0), 1, E, up_conv_cache, E, 0, 3, 0, B);
    export_tm_to_cm_kernel<<<grid_for(2LL * B * E), 256, 0, st>>>(s.conv[1], s.conv_pitch(1), 2, E, up_conv_cache, E, 0, 3, 1, B);
  }
  if (bb_conv_cache1)  // (B,E,6)
    export_tm_to_cm_kernel<<<grid_for(6LL * B * E), 256, 0, st>>>(s.conv[2], s.conv_pitch(2), 6, E, bb_conv_cache1, E, 0, 6, 0, B);
  if (bb_conv_cache2)  // (B,8E,2): 4 blocks x [conv1 input | conv2 input]  (decoder.py:150-171,317-319)
    for (int i = 0; i < 8; ++i)
      export_tm_to_cm_kernel<<<grid_for(2LL * B * E), 256, 0, st>>>(s.conv[3 + i], s.conv_pitch(3 + i), 2, E, bb_conv_cache2, 8 * E, i * E, 2, 0, B);
  if (bb_kv_cache && s.n_tokens > 0)
    for (int l = 0; l < h.nl; ++l)
      export_kv_kernel<<<grid_for(2LL * B * 8 * s.n_tokens * E), 256, 0, st>>>(s.kv[l], s.kv_pitch(), 8 * s.n_tokens, E, h.H, h.hd, bb_kv_cache, h.nl, l, B);
  if (is_cache)
    transpose_tail_kernel<<<(B * 3 * h.n_fft + 255) / 256, 256, 0, st>>>(s.tail, is_cache, B, h.n_fft, 1);
  FRT2_CUDA_OK(cudaGetLastError());
  FRT2_TRY(s.end_use(st));
  return FRT2_OK;
}

int frt2_import_state(frt2_handle* hh, frt2_stream* ss, int n_tokens, const float* up_conv_cache,
                      const float* bb_conv_cache1, const float* bb_conv_cache2, const float* bb_kv_cache,
                      const float* is_cache, void* cuda_stream) {
  FRT2_REQUIRE(hh && ss, FRT2_ERR_BAD_ARG, "null handle/stream");
  Handle& h = hh->h;
  Stream& s = ss->s;
  FRT2_REQUIRE(n_tokens >= 0 && n_tokens <= s.max_tokens, FRT2_ERR_STATE_OVERFLOW, "n_tokens exceeds the stream capacity");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  FRT2_TRY(s.begin_use(st));
  const int E = h.E, B = s.B;
  if (up_conv_cache) {
    import_cm_to_tm_kernel<<<grid_for(1LL * B * E), 256, 0, st>>>(up_conv_cache, E, 0, 3, 0, s.conv[0], s.conv_pitch(0), 1, E, B);
    import_cm_to_tm_kernel<<<grid_for(2LL * B * E), 256, 0, st>>>(up_conv_cache, E, 0, 3, 1, s.conv[1], s.conv_pitch(1), 2, E, B);
  }
  if (bb_conv_cache1)
    import_cm_to_tm_kernel<<<grid_for(6LL * B * E), 256, 0, st>>>(bb_conv_cache1, E, 0, 6, 0, s.conv[2], s.conv_pitch(2), 6, E, B);
  if (bb_conv_cache2)
    for (int i = 0; i < 8; ++i)
      import_cm_to_tm_kernel<<<grid_for(2LL * B * E), 256, 0, st>>>(bb_conv_cache2, 8 * E, i * E, 2, 0, s.conv[3 + i], s.conv_pitch(3 + i), 2, E, B);
  if (bb_kv_cache && n_tokens > 0)
    for (int l = 0; l < h.nl; ++l)
      import_kv_kernel<<<grid_for(2LL * B * 8 * n_tokens * E), 256, 0, st>>>(bb_kv_cache, h.nl, l, s.kv[l], s.kv_pitch(), 8 * n_tokens, E, h.H, h.hd, B);
  if (is_cache)
    transpose_tail_kernel<<<(B * 3 * h.n_fft + 255) / 256, 256, 0, st>>>(is_cache, s.tail, B, h.n_fft, 0);
  FRT2_CUDA_OK(cudaGetLastError());
  s.n_tokens = n_tokens;
  std::vector<int> ctrl_h(static_cast<size_t>(B) * CTRL_INTS, 0);
  for (int b = 0; b < B; ++b) ctrl_h[b * CTRL_INTS + CTRL_POS] = 8 * n_tokens;
  FRT2_CUDA_OK(cudaMemcpyAsync(s.ctrl, ctrl_h.data(), ctrl_h.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  FRT2_CUDA_OK(cudaStreamSynchronize(st));
  FRT2_TRY(s.end_use(st));
  return FRT2_OK;
}

int frt2_rvq_gather(frt2_handle* hh, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                    int nq, int L, float* rows, float* sum, void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_REQUIRE(h.finalized, FRT2_ERR_NOT_FINALIZED, "handle not finalized");
  FRT2_REQUIRE(nq >= 1 && nq <= h.nq && B >= 1 && L >= 1 && tokens, FRT2_ERR_BAD_ARG, "bad argument");
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  return rvq_gather_sum(tokens, idx_bytes, sB, sQ, sL, B, nq, L, h.codebooks, h.K, h.cd, sum, nullptr, rows, h.err_word,
                        static_cast<cudaStream_t>(cuda_stream));
}

int frt2_rvq_encode(frt2_handle* hh, const float* z, int64_t sB, int64_t sD, int64_t sT, int B, int input_dim, int T,
                    int nq, int64_t* codes, void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_REQUIRE(h.finalized, FRT2_ERR_NOT_FINALIZED, "handle not finalized");
  FRT2_REQUIRE(h.enc_ready, FRT2_ERR_MISSING_TENSOR,
               "frt2_rvq_encode: the encode-side tensors (rvq.quantizers.*.in_project, rvq.input_proj) were not loaded");
  FRT2_REQUIRE(z != nullptr && codes != nullptr && B >= 0 && T >= 0, FRT2_ERR_BAD_ARG, "frt2_rvq_encode: bad argument");
  FRT2_REQUIRE(input_dim == h.enc_input_dim, FRT2_ERR_BAD_ARG, "frt2_rvq_encode: z must have rvq.input_dim channels");
  FRT2_REQUIRE(nq >= 1 && nq <= h.nq, FRT2_ERR_BAD_ARG, "nq must be in [1, num_quantizers]");
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  RvqEncDesc d{};
  d.z = z; d.sB = sB; d.sD = sD; d.sT = sT; d.B = B; d.T = T; d.nq = nq;
  d.input_dim = h.enc_input_dim; d.rd = h.rd; d.cd = h.cd; d.K = h.K;
  d.WinpT = h.enc_WinpT; d.binp = h.enc_binp; d.WinT = h.enc_WinT; d.bin = h.enc_bin; d.CT = h.enc_CT; d.c2 = h.enc_c2;
  d.C = h.codebooks; d.WoutT = h.enc_WoutT; d.bout = h.enc_bout;
  d.codes = reinterpret_cast<long long*>(codes);
  d.s_inp = h.enc_s_inp; d.s_in = h.enc_s_in; d.s_C = h.enc_s_C; d.s_out = h.enc_s_out; d.nbout = h.enc_nbout;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // product path: the chain on the tensor cores (split-fp16 GEMMs); FRT2_RVQ_ENC_SIMT=1 / DBG_GEMM_REF select the
  // CUDA-core kernel (A/B, checker), which also serves widths that are not multiples of 64
  static const bool force_simt = getenv("FRT2_RVQ_ENC_SIMT") != nullptr && atoi(getenv("FRT2_RVQ_ENC_SIMT")) != 0;
  if (force_simt || (h.debug & DBG_GEMM_REF) || !rvq_encode_tc_applicable(d)) return rvq_encode(d, st);
  std::lock_guard<std::mutex> lk(h.mu);
  // slabs of tokens bound the score buffer (K floats per token)
  const long long R = static_cast<long long>(B) * T;
  if (R == 0) return FRT2_OK;
  const long long SLAB = 32768;
  if (R <= SLAB) {
    FRT2_TRY(h.ensure_ws(rvq_encode_tc_ws_bytes(d, R)));
    FRT2_TRY(h.ws_acquire(st));
    long long n = 0;
    FRT2_TRY(rvq_encode_tc(d, h.ws, st, &n));
    h.launches += n;
    return h.ws_release(st);
  }
  // larger inputs: item by item groups whose token count fits a slab (codes are (nq, B, T): a group is a B-slice,
  // written through a temporary because its rows are not contiguous in the output)
  const int per = static_cast<int>(std::max<long long>(1, SLAB / std::max(1, T)));
  FRT2_REQUIRE(static_cast<long long>(T) <= SLAB, FRT2_ERR_BAD_ARG, "frt2_rvq_encode: more than 32768 frames per item");
  FRT2_TRY(h.ensure_ws(rvq_encode_tc_ws_bytes(d, static_cast<long long>(per) * T) +
                       static_cast<size_t>(nq) * per * T * sizeof(long long)));
  FRT2_TRY(h.ws_acquire(st));
  long long* tmp = reinterpret_cast<long long*>(h.ws + rvq_encode_tc_ws_bytes(d, static_cast<long long>(per) * T));
  for (int b0 = 0; b0 < B; b0 += per) {
    RvqEncDesc g = d;
    g.B = std::min(per, B - b0);
    g.z = z + static_cast<int64_t>(b0) * sB;
    g.codes = tmp;
    long long n = 0;
    FRT2_TRY(rvq_encode_tc(g, h.ws, st, &n));
    h.launches += n;
    FRT2_CUDA_OK(cudaMemcpy2DAsync(d.codes + static_cast<long long>(b0) * T, static_cast<size_t>(B) * T * 8, tmp,
                                   static_cast<size_t>(g.B) * T * 8, static_cast<size_t>(g.B) * T * 8, nq,
                                   cudaMemcpyDeviceToDevice, st));
  }
  return h.ws_release(st);
}

// FIR bank of torchaudio's sinc_interp_hann resampler, computed in float32 in the same operation order as
// torchaudio.functional._get_sinc_resample_kernel is run by resample() on a float32 waveform
namespace {
struct ResampleBank { float* taps = nullptr; int K = 0, width = 0, orig = 0, nnew = 0; };
std::mutex g_resample_mu;
std::map<std::tuple<int, int, int>, ResampleBank> g_resample_banks;   // (device, orig/gcd, new/gcd)

int get_resample_bank(int device, int orig_freq, int new_freq, ResampleBank* out) {
  int a = orig_freq, b = new_freq;
  while (b) { const int t = a % b; a = b; b = t; }
  const int orig = orig_freq / a, nnew = new_freq / a;
  std::lock_guard<std::mutex> lk(g_resample_mu);
  auto key = std::make_tuple(device, orig, nnew);
  auto it = g_resample_banks.find(key);
  if (it != g_resample_banks.end()) { *out = it->second; return FRT2_OK; }
  const int lpw = 6;
  const double rolloff = 0.99;
  const float base = static_cast<float>(std::min(orig, nnew) * rolloff);
  const int width = static_cast<int>(std::ceil(lpw * orig / (std::min(orig, nnew) * rolloff)));
  const int K = 2 * width + orig;
  FRT2_REQUIRE(K <= 4096 && static_cast<int64_t>(K) * nnew <= (1 << 24), FRT2_ERR_BAD_ARG,
               "frt2_resample: rate pair needs too large a filter bank (reduce the rates by their gcd first)");
  std::vector<float> taps(static_cast<size_t>(nnew) * K);
  const float pi = static_cast<float>(M_PI);
  const float scale = static_cast<float>(std::min(orig, nnew) * rolloff / orig);
  for (int p = 0; p < nnew; ++p) {
    for (int k = 0; k < K; ++k) {
      const float idx = static_cast<float>(k - width) / static_cast<float>(orig);
      float t = static_cast<float>(-p) / static_cast<float>(nnew) + idx;
      t *= base;
      t = std::min(std::max(t, static_cast<float>(-lpw)), static_cast<float>(lpw));
      const float c = std::cos(t * pi / static_cast<float>(lpw) / 2.0f);
      const float window = c * c;
      t *= pi;
      const float sinc = (t == 0.0f) ? 1.0f : std::sin(t) / t;
      taps[static_cast<size_t>(p) * K + k] = sinc * (window * scale);
    }
  }
  ResampleBank bank;
  bank.K = K; bank.width = width; bank.orig = orig; bank.nnew = nnew;
  FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&bank.taps), taps.size() * 4));
  FRT2_CUDA_OK(cudaMemcpy(bank.taps, taps.data(), taps.size() * 4, cudaMemcpyHostToDevice));
  g_resample_banks[key] = bank;
  *out = bank;
  return FRT2_OK;
}
}  // namespace

int frt2_resample(int device, const float* in, int64_t in_pitch, int B, int64_t n_in, const int32_t* lengths,
                  int orig_freq, int new_freq, float* out, int64_t out_pitch, int64_t* n_out, void* cuda_stream) {
  FRT2_REQUIRE(in != nullptr && out != nullptr, FRT2_ERR_BAD_ARG, "frt2_resample: null pointer");
  FRT2_REQUIRE(orig_freq > 0 && new_freq > 0, FRT2_ERR_BAD_ARG,
               "Original frequency and desired frequecy should be positive");   // torchaudio's message
  FRT2_REQUIRE(B >= 0 && n_in >= 0, FRT2_ERR_BAD_ARG, "frt2_resample: negative size");
  int ndev = 0;
  FRT2_CUDA_OK(cudaGetDeviceCount(&ndev));
  FRT2_REQUIRE(device >= 0 && device < ndev, FRT2_ERR_BAD_ARG, "frt2_resample: no such CUDA device");
  FRT2_CUDA_OK(cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (orig_freq == new_freq) {   // torchaudio returns the waveform unchanged
    FRT2_REQUIRE(out_pitch >= n_in, FRT2_ERR_BAD_ARG, "frt2_resample: out_pitch too small");
    if (B > 0 && n_in > 0)
      FRT2_CUDA_OK(cudaMemcpy2DAsync(out, out_pitch * 4, in, in_pitch * 4, n_in * 4, B, cudaMemcpyDeviceToDevice, st));
    if (n_out) *n_out = n_in;
    return FRT2_OK;
  }
  ResampleBank bank;
  FRT2_TRY(get_resample_bank(device, orig_freq, new_freq, &bank));
  const int64_t n = (n_in * bank.nnew + bank.orig - 1) / bank.orig;
  FRT2_REQUIRE(out_pitch >= n && in_pitch >= n_in, FRT2_ERR_BAD_ARG, "frt2_resample: pitch too small");
  if (n_out) *n_out = n;
  return resample_rows(in, in_pitch, B, n_in, lengths, bank.taps, bank.K, bank.width, bank.orig, bank.nnew, out,
                       out_pitch, st);
}

int frt2_decode_resampled(frt2_handle* hh, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                          int nq, int L, const int32_t* lengths, float* audio, int64_t audio_pitch, int orig_freq,
                          int new_freq, float* audio_rs, int64_t rs_pitch, int64_t* n_rs, void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_TRY(check_decode_args(h, tokens, idx_bytes, B, nq, L, audio_rs));
  FRT2_REQUIRE(orig_freq > 0 && new_freq > 0, FRT2_ERR_BAD_ARG,
               "Original frequency and desired frequecy should be positive");   // torchaudio's message
  FRT2_REQUIRE(orig_freq != new_freq, FRT2_ERR_BAD_ARG, "frt2_decode_resampled: equal rates (use frt2_decode)");
  const int64_t n_in = static_cast<int64_t>(8) * h.hop * L;
  FRT2_REQUIRE(audio == nullptr || audio_pitch >= n_in, FRT2_ERR_BAD_ARG, "audio_pitch too small");
  ResampleBank bank;
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  FRT2_TRY(get_resample_bank(h.device, orig_freq, new_freq, &bank));
  FRT2_REQUIRE(bank.nnew <= 3, FRT2_ERR_BAD_ARG,
               "frt2_decode_resampled: new_freq / gcd(orig_freq, new_freq) must be 1, 2 or 3 (decode, then frt2_resample)");
  const int64_t n = (n_in * bank.nnew + bank.orig - 1) / bank.orig;
  FRT2_REQUIRE(rs_pitch >= n, FRT2_ERR_BAD_ARG, "frt2_decode_resampled: rs_pitch too small");
  if (n_rs) *n_rs = n;
  std::lock_guard<std::mutex> lk(h.mu);
  h.rs_out = audio_rs; h.rs_pitch = rs_pitch; h.rs_taps = bank.taps; h.rs_K = bank.K; h.rs_width = bank.width;
  h.rs_orig = bank.orig; h.rs_new = bank.nnew;
  const int rc = h.pipeline(tokens, idx_bytes, sB, sQ, sL, B, nq, L, lengths, audio, audio_pitch, nullptr, 1,
                            static_cast<cudaStream_t>(cuda_stream));
  h.rs_out = nullptr;
  return rc;
}

int frt2_set_debug(frt2_handle* hh, int flags) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  hh->h.debug = flags;
  return FRT2_OK;
}

int frt2_get_tap(frt2_handle* hh, const char* name, float* out, int64_t capacity, int64_t* n, void* cuda_stream) {
  FRT2_REQUIRE(hh && name && out && n, FRT2_ERR_BAD_ARG, "bad argument");
  Handle& h = hh->h;
  auto it = h.taps.find(name);
  FRT2_REQUIRE(it != h.taps.end() && it->second.first != nullptr, FRT2_ERR_BAD_ARG,
               "no such tap recorded (call frt2_set_debug(h, 1) before decoding)");
  FRT2_REQUIRE(capacity >= it->second.second, FRT2_ERR_BAD_ARG, "tap buffer too small");
  FRT2_CUDA_OK(cudaMemcpyAsync(out, it->second.first, it->second.second * 4, cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(cuda_stream)));
  *n = it->second.second;
  return FRT2_OK;
}

int frt2_profile(frt2_handle* hh, int enable) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  std::lock_guard<std::mutex> lk(h.mu);
  h.profile = enable != 0;
  h.prof_n = 0;
  h.launches = 0;
  return FRT2_OK;
}

int frt2_profile_get(frt2_handle* hh, int cls, double* ms, int64_t* launches, double* flops, double* bytes) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  std::lock_guard<std::mutex> lk(h.mu);
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  FRT2_CUDA_OK(cudaDeviceSynchronize());
  double t = 0, f = 0, by = 0;
  int64_t n = 0;
  for (size_t i = 0; i < h.prof_n; ++i) {
    const auto& r = h.prof[i];
    if (cls != FRT2_PROF_ALL && r.cls != cls) continue;
    float e = 0.f;
    FRT2_CUDA_OK(cudaEventElapsedTime(&e, r.a, r.b));
    t += e; f += r.flops; by += r.bytes; ++n;
  }
  if (cls == FRT2_PROF_ALL) n = h.launches;
  if (ms) *ms = t;
  if (launches) *launches = n;
  if (flops) *flops = f;
  if (bytes) *bytes = by;
  return FRT2_OK;
}

int frt2_check_error(frt2_handle* hh, void* cuda_stream) {
  FRT2_REQUIRE(hh, FRT2_ERR_BAD_ARG, "null handle");
  Handle& h = hh->h;
  FRT2_CUDA_OK(cudaSetDevice(h.device));
  unsigned int word = 0;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  FRT2_CUDA_OK(cudaMemcpyAsync(&word, h.err_word, 4, cudaMemcpyDeviceToHost, st));
  FRT2_CUDA_OK(cudaStreamSynchronize(st));
  if (word != 0) {
    FRT2_CUDA_OK(cudaMemsetAsync(h.err_word, 0, 4, st));
    if (word & DEV_ERR_INDEX_OOR) {
      set_error("index out of range in self");
      return FRT2_ERR_INDEX_OUT_OF_RANGE;
    }
  }
  return FRT2_OK;
}

// ---- single-operator entry points ----
int frt2_op_gemm(int impl, const void* A16, const void* W16, int batches, int rows_per_batch, int Kc, int ntaps, int N,
                 float alpha, const float* bias, int act, const float* resid, float* out32, void* out16,
                 void* cuda_stream) {
  GemmDesc g{};
  g.A = static_cast<const __half*>(A16); g.a_row_pitch = Kc; g.a_batch_pitch = static_cast<int64_t>(rows_per_batch) * Kc;
  g.rows_a = rows_per_batch; g.batches = batches; g.Kc = Kc; g.ntaps = ntaps; g.row_shift = -(ntaps - 1);
  g.W = static_cast<const __half*>(W16); g.N = N; g.rows_out = rows_per_batch;
  g.pitch32 = static_cast<int64_t>(rows_per_batch) * N; g.pitch16 = g.pitch32; g.alpha = alpha; g.bias = bias; g.act = act;
  g.resid = resid; g.out32 = out32; g.ld32 = N; g.out16 = static_cast<__half*>(out16); g.ld16 = N;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (impl == 2) {
    FRT2_TRY(gemm_skinny_init());
    return gemm_skinny(g, st);
  }
  if (impl == 3) {   // K2w: the weights are repacked here (tile-blocked order), test-only convenience; synchronises
    FRT2_REQUIRE(batches == 1 && ntaps == 1 && alpha == 1.0f && gemm_stream_applicable(N, Kc, rows_per_batch), FRT2_ERR_BAD_ARG,
                 "frt2_op_gemm(impl 3): one batch of <= 16 rows that fit the kernel's activation tile, one tap, alpha 1, K a multiple of 32");
    FRT2_TRY(gemm_stream_init());
    const size_t nw = static_cast<size_t>(N) * Kc;
    std::vector<__half> hw(nw);
    FRT2_CUDA_OK(cudaMemcpy(hw.data(), W16, nw * 2, cudaMemcpyDeviceToHost));
    std::vector<float> fw(nw);
    for (size_t i = 0; i < nw; ++i) fw[i] = __half2float(hw[i]);
    std::vector<__half> packed(gemm_stream_packed_elems(N, Kc));
    gemm_stream_pack_host(fw.data(), N, Kc, packed.data());
    __half* dw = nullptr;
    FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dw), packed.size() * 2));
    cudaError_t e = cudaMemcpy(dw, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice);
    int rc = FRT2_OK;
    if (e == cudaSuccess) {
      StreamGemm d{};
      d.Wt = dw; d.N = N; d.K = Kc; d.B = rows_per_batch; d.A = g.A; d.lda = Kc; d.bias = bias; d.act = act; d.resid = resid;
      d.out32 = out32; d.ld32 = N; d.out16 = g.out16; d.ld16 = act == ACT_SWIGLU ? N / 2 : N;
      rc = gemm_stream(d, st);
      if (rc == FRT2_OK) e = cudaStreamSynchronize(st);
    }
    cudaFree(dw);
    FRT2_CUDA_OK(e);
    return rc;
  }
  return impl == 0 ? gemm_tc(g, st) : gemm_ref(g, st);
}

int frt2_op_layer_norm(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                       int apply_silu, void* out16, void* cuda_stream) {
  return layer_norm_rows(x, C, rows, C, gamma, beta, eps, apply_silu, static_cast<__half*>(out16), C,
                         static_cast<cudaStream_t>(cuda_stream));
}

int frt2_op_attention(int impl, const void* q16, const void* k16, const void* v16, void* out16, int B, int H, int hd,
                      int Tq, int Tk, int q_pos0, int block_causal, void* cuda_stream) {
  AttnDesc a{};
  const int64_t E = static_cast<int64_t>(H) * hd;
  a.q = static_cast<const __half*>(q16); a.q_row_pitch = E; a.q_batch_pitch = E * Tq;
  a.k = static_cast<const __half*>(k16); a.v = static_cast<const __half*>(v16); a.kv_row_pitch = E; a.kv_batch_pitch = E * Tk;
  a.out = static_cast<__half*>(out16); a.o_row_pitch = E; a.o_batch_pitch = E * Tq;
  a.B = B; a.H = H; a.hd = hd; a.Tq = Tq; a.Tk = Tk; a.q_pos0 = q_pos0; a.block_causal = block_causal;
  a.scale = 1.0f / std::sqrt(static_cast<float>(hd));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  return impl == 0 ? attention_tc(a, st) : attention_warp(a, st);
}

int frt2_op_attention_trace(void* dev_buf) {
  attention_tc_set_trace(dev_buf);
  return FRT2_OK;
}

int frt2_op_overlap_add(const float* frames, const float* tail, const float* window, const int32_t* lengths,
                        float* audio, int64_t audio_pitch, int B, int T, int n_fft, int hop, int first, int last,
                        void* cuda_stream) {
  OlaDesc d{};
  d.frames = frames; d.frames_batch_pitch = static_cast<int64_t>(T) * n_fft; d.tail = tail; d.window = window;
  d.lengths = lengths; d.len_mul = 8; d.audio = audio; d.audio_pitch = audio_pitch; d.B = B; d.T = T; d.n_fft = n_fft;
  d.hop = hop; d.first = first; d.last = last; d.pcm16 = nullptr;
  return istft_overlap_add(d, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
