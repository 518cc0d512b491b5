// Codec ENCODE side behind the two feature encoders: everything `RedCodecInfer._encode_one_batch`
// (reference codec/model.py:218-236) runs between the Whisper encoders and `ResidualVQ.encode_codes`:
//
//   sem  = SslAdaptor(ssl)                      model.py:19-77   Linear -> N x WhisperEncoderLayer (full attention)
//                                                                 -> LayerNorm -> Linear
//   x    = cat([sem, aco], dim=2)               model.py:230
//   vq   = ResidualDownConv(x)                  model.py:80-121  gate/up Conv1d(k = s = pooler), SiLU(g)*u, down_proj,
//                                                                 LayerNorm(c + x.reshape), out_proj
//
// and, when the checkpoint's feature encoders are loaded too, the whole of `_encode_one_batch` from the 16 kHz waveform:
//
//   mel  = WhisperMelExtractor(audio16k)        whisper.py:276-302  STFT 400/160 (Hann, centre, reflect), power,
//                                                                    slaney mel bank, log10, per-item max - 8 floor, (x+4)/4
//   ssl  = PretrainedWhisperEncoder(mel)        whisper.py:195-258,352-385   conv k3 + GELU, conv k3 s2 + GELU, + positions,
//   aco  = WhisperAcousticEncoder(mel)          whisper.py:388-431           N x WhisperEncoderLayer, LayerNorm
//
// Built from the decode path's kernels: every Linear / strided conv is a tcgen05 GEMM (gemm_tc; the k = s = pooler
// convolutions are plain GEMMs on the (T/pooler, pooler*D) view of the time-major rows, gate and up share ONE launch),
// attention is the tcgen05 flash kernel with the mask switched off (make_nonpad_mask of full-length chunks,
// model.py:220-222), LayerNorm is the two-rows-per-warp kernel.  fp16 operands, fp32 accumulation / residual stream /
// statistics, exactly as on the decode side.  The result feeds frt2_rvq_encode.
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

struct HostT {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

// fp32 rows -> fp32 and/or fp16 rows at a column offset of wider buffers (the torch.cat of model.py:230 and the
// operand conversion in front of a GEMM in one pass); 4 elements per thread
__global__ void __launch_bounds__(256) cvt_rows_kernel(const float* __restrict__ src, long long ld_src, long long rows,
                                                       int C4, float* __restrict__ dst32, long long ld32,
                                                       __half* __restrict__ dst16, long long ld16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const long long r = i / C4;
  const int c = static_cast<int>(i - r * C4) * 4;
  const float4 v = __ldcs(reinterpret_cast<const float4*>(src + r * ld_src + c));
  if (dst32 != nullptr) *reinterpret_cast<float4*>(dst32 + r * ld32 + c) = v;
  if (dst16 != nullptr) {
    uint2 h;
    h.x = pack_half2(v.x, v.y);
    h.y = pack_half2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst16 + r * ld16 + c) = h;
  }
}

// act = SiLU(gate) * up on the (rows, 2C) output of the merged gate|up GEMM (model.py:116): 8 elements per thread
__global__ void __launch_bounds__(256) silu_mul_kernel(const __half* __restrict__ gu, long long rows, int C8,
                                                       __half* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C8) return;
  const long long r = i / C8;
  const int c = static_cast<int>(i - r * C8);
  const uint4 g = __ldcs(reinterpret_cast<const uint4*>(gu + r * 16 * C8) + c);
  const uint4 u = __ldcs(reinterpret_cast<const uint4*>(gu + r * 16 * C8 + 8 * C8) + c);
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 gf = __half22float2(*reinterpret_cast<const __half2*>(&gw[j]));
    const float2 uf = __half22float2(*reinterpret_cast<const __half2*>(&uw[j]));
    o[j] = pack_half2(silu(gf.x) * uf.x, silu(gf.y) * uf.y);
  }
  reinterpret_cast<uint4*>(out + r * 8 * C8)[c] = make_uint4(o[0], o[1], o[2], o[3]);
}

// x[b, t, :] += pos[t, :]  (WhisperEncoder.forward, whisper.py:238-242); 4 elements per thread
__global__ void __launch_bounds__(256) add_pos_kernel(float* __restrict__ x, const float* __restrict__ pos, long long rows,
                                                      int T, int C4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const long long r = i / C4;
  const int c = static_cast<int>(i - r * C4);
  const int t = static_cast<int>(r % T);
  float4 v = reinterpret_cast<float4*>(x)[i];
  const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + static_cast<long long>(t) * C4 + c);
  v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  reinterpret_cast<float4*>(x)[i] = v;
}

// LayerNorm with an fp32 result (the encoders' final layer_norm, whisper.py:250: its output is a residual input of
// ResidualDownConv, model.py:117) plus the fp16 copy the next GEMM reads; one warp per row, two passes over the row
__global__ void __launch_bounds__(256) layer_norm_f32_kernel(const float* __restrict__ x, long long rows, int C,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float eps, float* __restrict__ out32, long long ld32,
                                                             __half* __restrict__ out16, long long ld16) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  const int C4 = C >> 2;
  float s = 0.f;
  for (int c = lane; c < C4; c += 32) {
    const float4 v = xr[c];
    s += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
  for (int c = lane; c < C4; c += 32) {
    const float4 v = xr[c];
    const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (cc * cc + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
  for (int c = lane; c < C4; c += 32) {
    const float4 v = xr[c];
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
    const float4 y = make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y,
                                 (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w);
    if (out32 != nullptr) *reinterpret_cast<float4*>(out32 + row * ld32 + 4 * c) = y;
    if (out16 != nullptr) {
      uint2 h;
      h.x = pack_half2(y.x, y.y);
      h.y = pack_half2(y.z, y.w);
      *reinterpret_cast<uint2*>(out16 + row * ld16 + 4 * c) = h;
    }
  }
}

// ---- log-mel front end (WhisperMelExtractor.extract_fbank, whisper.py:276-296) ----
// torch.stft(audio, n_fft, hop, window=hann(n_fft), center=True, pad_mode="reflect") -> |X|^2 of bins 0..n_fft/2 of every
// frame but the last -> mel bank -> log10(max(., 1e-10)).  One CTA per MEL_FR frames of one item: the reflect-padded,
// windowed frames sit in shared memory, every thread evaluates whole DFT bins (fp32, twiddles from a shared table walked
// with an integer phase, so no argument reduction error), then (frame, mel) dot products over the power spectrum.
// The per-item maximum (whisper.py:293) is collected with an integer atomicMax on log10 + 10 >= 0.
constexpr int MEL_FR = 8;
__global__ void __launch_bounds__(256) mel_power_kernel(const float* __restrict__ audio, long long audio_pitch, long long n,
                                                        int T, int n_fft, int hop, const float* __restrict__ window,
                                                        const float* __restrict__ bank /*(bins, n_mels)*/, int n_mels,
                                                        float* __restrict__ logmel /*(B, T, n_mels)*/,
                                                        int* __restrict__ item_max) {
  extern __shared__ float mel_smem[];
  const int bins = n_fft / 2 + 1;
  float* s_cos = mel_smem;                 // [n_fft]
  float* s_sin = s_cos + n_fft;            // [n_fft]
  float* s_fr = s_sin + n_fft;             // [MEL_FR][n_fft] windowed frames
  float* s_pw = s_fr + MEL_FR * n_fft;     // [MEL_FR][bins]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * MEL_FR;
  const float* a = audio + static_cast<long long>(b) * audio_pitch;
  for (int j = threadIdx.x; j < n_fft; j += blockDim.x) {
    float sn, cs;
    sincospif(2.0f * static_cast<float>(j) / static_cast<float>(n_fft), &sn, &cs);
    s_cos[j] = cs;
    s_sin[j] = sn;
  }
  const int half = n_fft / 2;
  for (int e = threadIdx.x; e < MEL_FR * n_fft; e += blockDim.x) {
    const int f = e / n_fft, j = e - f * n_fft;
    const int t = t0 + f;
    float v = 0.f;
    if (t < T) {
      long long i = static_cast<long long>(t) * hop - half + j;     // index into the unpadded waveform
      if (i < 0) i = -i;                                            // reflect (no edge repeat)
      if (i >= n) i = 2 * (n - 1) - i;
      v = a[i] * __ldg(window + j);
    }
    s_fr[e] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < MEL_FR * bins; e += blockDim.x) {
    const int f = e / bins, k = e - f * bins;
    const float* fr = s_fr + f * n_fft;
    float re = 0.f, im = 0.f;
    int ph = 0;
    for (int j = 0; j < n_fft; ++j) {
      const float v = fr[j];
      re = fmaf(v, s_cos[ph], re);
      im = fmaf(v, s_sin[ph], im);      // sign irrelevant for the power
      ph += k;
      if (ph >= n_fft) ph -= n_fft;
    }
    s_pw[e] = re * re + im * im;
  }
  __syncthreads();
  float local_max = 0.f;
  for (int e = threadIdx.x; e < MEL_FR * n_mels; e += blockDim.x) {
    const int f = e / n_mels, m = e - f * n_mels;
    const int t = t0 + f;
    if (t >= T) continue;
    const float* pw = s_pw + f * bins;
    float acc = 0.f;
    for (int k = 0; k < bins; ++k) acc = fmaf(__ldg(bank + static_cast<long long>(k) * n_mels + m), pw[k], acc);
    const float lg = log10f(fmaxf(acc, 1e-10f));
    logmel[(static_cast<long long>(b) * T + t) * n_mels + m] = lg;
    local_max = fmaxf(local_max, lg + 10.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if ((threadIdx.x & 31) == 0) atomicMax(item_max + b, __float_as_int(local_max));
}

// ---- the same front end with the DFT on the tensor cores (product path) ----
// The 400-point real DFT of every frame is a GEMM: frames (rows) x [cos | sin] twiddles (2 x 201 columns), evaluated with
// SPLIT fp16 operands like the RVQ chain (rvq_encode_tc.cu): x = x_hi + x_lo, three-term product in ONE GEMM over a
// reduction of 3 * KP (KP = n_fft padded to a multiple of 64), fp32 accumulation — ~22 significant bits, the precision
// of the reference's fp32 FFT — instead of 2 * 400 * 201 FMAs per frame on the CUDA cores (the direct kernel above took
// 1.64 ms per 32 chunks, 8 % of the whole encode; it stays as the checker, FRT2_MEL_SIMT=1).
constexpr float MEL_S = 256.0f;

// reflect-padded, Hann-windowed frames -> split rows [hi | hi/S | lo*S] of KP columns each; one thread per (frame, column)
__global__ void __launch_bounds__(256) mel_frames_kernel(const float* __restrict__ audio, long long audio_pitch, long long n,
                                                         int T, int n_fft, int hop, int KP, const float* __restrict__ window,
                                                         __half* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_item = static_cast<long long>(T) * KP;
  const int b = blockIdx.y;
  if (i >= per_item) return;
  const int t = static_cast<int>(i / KP), j = static_cast<int>(i - static_cast<long long>(t) * KP);
  float v = 0.f;
  if (j < n_fft) {
    long long k = static_cast<long long>(t) * hop - n_fft / 2 + j;
    if (k < 0) k = -k;
    if (k >= n) k = 2 * (n - 1) - k;
    v = audio[static_cast<long long>(b) * audio_pitch + k] * __ldg(window + j);
  }
  const __half hi = to_half_sat(v);
  const float h = __half2float(hi);
  __half* o = out + (static_cast<long long>(b) * T + t) * 3 * KP + j;
  o[0] = hi;
  o[KP] = __float2half_rn(h * (1.0f / MEL_S));
  o[2 * KP] = __float2half_rn((v - h) * MEL_S);
}

// spec (frames, ld) fp32 = [Re X_0..X_bins-1 | Im X_0..X_bins-1] -> power -> mel bank -> log10, per-item maximum
__global__ void __launch_bounds__(256) mel_bank_kernel(const float* __restrict__ spec, long long ld, int T, int bins,
                                                       const float* __restrict__ bank, int n_mels,
                                                       float* __restrict__ logmel, int* __restrict__ item_max) {
  extern __shared__ float mel_smem[];
  float* s_pw = mel_smem;                  // [MEL_FR][bins]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * MEL_FR;
  for (int e = threadIdx.x; e < MEL_FR * bins; e += blockDim.x) {
    const int f = e / bins, k = e - f * bins;
    float pw = 0.f;
    if (t0 + f < T) {
      const float* row = spec + (static_cast<long long>(b) * T + t0 + f) * ld;
      const float re = row[k], im = row[bins + k];
      pw = re * re + im * im;
    }
    s_pw[e] = pw;
  }
  __syncthreads();
  float local_max = 0.f;
  for (int e = threadIdx.x; e < MEL_FR * n_mels; e += blockDim.x) {
    const int f = e / n_mels, m = e - f * n_mels;
    const int t = t0 + f;
    if (t >= T) continue;
    const float* pw = s_pw + f * bins;
    float acc = 0.f;
    for (int k = 0; k < bins; ++k) acc = fmaf(__ldg(bank + static_cast<long long>(k) * n_mels + m), pw[k], acc);
    const float lg = log10f(fmaxf(acc, 1e-10f));
    logmel[(static_cast<long long>(b) * T + t) * n_mels + m] = lg;
    local_max = fmaxf(local_max, lg + 10.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if ((threadIdx.x & 31) == 0) atomicMax(item_max + b, __float_as_int(local_max));
}

// log_spec = max(log_spec, item_max - 8); (log_spec + 4) / 4  (whisper.py:294-295) -> fp32 (parity hook) and fp16 (conv1's operand)
__global__ void __launch_bounds__(256) mel_norm_kernel(const float* __restrict__ logmel, const int* __restrict__ item_max,
                                                       long long per_item, long long total, float* __restrict__ out32,
                                                       __half* __restrict__ out16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float mx = __int_as_float(item_max[i / per_item]) - 10.0f;
  const float v = (fmaxf(logmel[i], mx - 8.0f) + 4.0f) / 4.0f;
  if (out32 != nullptr) out32[i] = v;
  if (out16 != nullptr) out16[i] = to_half_sat(v);
}

struct EncLayer {
  __half *w_qkv = nullptr, *w_o = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;
  float *b_qkv = nullptr, *b_o = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  // LayerNorm folded across the GEMMs as on the decode side: LN(x) W^T = rstd (x (gamma.W)^T - mean colsum)
  // + (b + W beta); the producer of the residual stream (out-projection, fc2) also writes its fp16 copy
  __half *w_qkv_f = nullptr, *w_fc1_f = nullptr;
  float *s_qkv = nullptr, *c_qkv = nullptr, *s_fc1 = nullptr, *c_fc1 = nullptr;
};

// N x WhisperEncoderLayer (whisper.py:121-163) + the LayerNorm behind them.  Head dims the attention kernels do not
// serve (e.g. 96 = the acoustic encoder's 768 / 8 default) are zero-padded to hdp = 64 or 128 in the packed q|k|v and
// out-projection weights: padded q.k products and padded v columns are exactly zero, so the result is unchanged.
struct EncStack {
  int E = 0, H = 0, hd = 0, hdp = 0, F = 0;
  std::vector<EncLayer> layers;
  float *lnf_g = nullptr, *lnf_b = nullptr;
};

// WhisperEncoder front (whisper.py:206-242): conv1 k3 p1 + GELU, conv2 k3 s2 p1 + GELU, + sinusoidal positions
struct EncFront {
  int in_dim = 0, max_pos = 0;
  __half *w_c1 = nullptr, *w_c2 = nullptr;   // (E, 3*in_dim) tap-major; (E, 4E) = taps over the paired-frame view
  float *b_c1 = nullptr, *b_c2 = nullptr, *pos = nullptr;
};

}  // namespace

struct Encoder {
  int device = 0;
  std::mutex mu;
  std::map<std::string, HostT> raw;
  bool finalized = false;
  // config
  int ssl_in = 0, ssl_out = 0, aco = 0, pool = 4, D = 0;
  int n_mels = 128, n_fft = 400, hop = 160;
  bool has_front = false;        // the two Whisper encoders + mel front end are loaded (frt2_enc_audio_features)
  // weights
  std::vector<void*> owned;
  EncStack ada;                                       // ssl_adaptor.layers
  __half* w_in = nullptr;   float* b_in = nullptr;
  __half* w_out = nullptr;  float* b_out = nullptr;
  __half* w_gu = nullptr;                            // (2*pool*D, pool*D): gate rows, then up rows, tap-major K
  __half* w_down = nullptr;                          // (pool*D, pool*D)
  float *dln_g = nullptr, *dln_b = nullptr;
  __half* w_dout = nullptr; float* b_dout = nullptr; // (D, pool*D)
  EncStack ssl_stack, aco_stack;                     // ssl.layers / acoustic_encoder.layers
  EncFront ssl_front, aco_front;
  float *mel_window = nullptr, *mel_bank = nullptr;  // hann(n_fft); (n_fft/2+1, n_mels) slaney bank
  __half* mel_dft_s = nullptr;                       // (2*bins, 3*KP) split-fp16 [cos | sin] twiddle rows
  int mel_kp = 0;                                    // n_fft padded to a multiple of 64
  unsigned int* sched = nullptr;                     // item-scheduler words of the persistent attention kernel
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  cudaEvent_t ws_event = nullptr;
  cudaStream_t ws_last = nullptr;
  bool ws_used = false;
  long long launches = 0;

  ~Encoder() {
    cudaSetDevice(device);
    for (void* p : owned) cudaFree(p);
    if (ws) cudaFree(ws);
    if (ws_event) cudaEventDestroy(ws_event);
  }
  int dev_alloc(void** p, size_t bytes) {
    FRT2_CUDA_OK(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    owned.push_back(*p);
    return FRT2_OK;
  }
  int up32(const std::vector<float>& v, float** out) {
    FRT2_TRY(dev_alloc(reinterpret_cast<void**>(out), v.size() * 4));
    FRT2_CUDA_OK(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return FRT2_OK;
  }
  int up16(const std::vector<float>& v, __half** out) {
    std::vector<__half> hb(v.size());
    const long long n = static_cast<long long>(v.size());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) hb[i] = __float2half_rn(std::min(65504.0f, std::max(-65504.0f, v[i])));
    FRT2_TRY(dev_alloc(reinterpret_cast<void**>(out), hb.size() * 2));
    FRT2_CUDA_OK(cudaMemcpy(*out, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
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
  int load_stack(const std::string& prefix, const std::string& ln_key, EncStack& s);
  int load_front(const std::string& prefix, int E, EncFront& f);
  int finalize();
  int ensure_ws(size_t bytes);
  int begin(cudaStream_t st) {
    // calls are asynchronous and share the arena: a call on another CUDA stream waits for the previous one on the device
    if (ws_used && st != ws_last) FRT2_CUDA_OK(cudaStreamWaitEvent(st, ws_event, 0));
    return FRT2_OK;
  }
  int end(cudaStream_t st) {
    if (ws_event == nullptr) FRT2_CUDA_OK(cudaEventCreateWithFlags(&ws_event, cudaEventDisableTiming));
    FRT2_CUDA_OK(cudaEventRecord(ws_event, st));
    ws_last = st;
    ws_used = true;
    return FRT2_OK;
  }
  // ---- building blocks (all asynchronous on `st`) ----
  int gemm(const __half* A, int64_t rows, int K, const __half* W, int N, const float* bias, int act, const float* resid,
           float* out32, int64_t ld32, __half* out16, int64_t ld16, cudaStream_t st, __half* x16_copy = nullptr,
           const float2* stats = nullptr, const float* colsum = nullptr) {
    GemmDesc g{};
    g.x16_out = x16_copy; g.ld_x16 = N; g.stats_in = stats; g.colsum = colsum;
    g.A = A; g.a_row_pitch = K; g.a_batch_pitch = 0; g.rows_a = static_cast<int>(rows); g.batches = 1;
    g.Kc = K; g.ntaps = 1; g.row_shift = 0; g.W = W; g.N = N; g.rows_out = static_cast<int>(rows);
    g.alpha = 1.0f; g.bias = bias; g.act = act; g.resid = resid; g.out32 = out32; g.ld32 = ld32; g.out16 = out16;
    g.ld16 = ld16;
    ++launches;
    return gemm_tc(g, st);
  }
  int cvt(const float* src, int64_t ld_src, int64_t rows, int C, float* d32, int64_t ld32, __half* d16, int64_t ld16,
          cudaStream_t st) {
    const long long n = rows * (C / 4);
    cvt_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, ld_src, rows, C / 4, d32, ld32, d16, ld16);
    ++launches;
    FRT2_CUDA_OK(cudaGetLastError());
    return FRT2_OK;
  }
  int ln16(const float* x, int64_t rows, int C, const float* g, const float* b, __half* out, cudaStream_t st) {
    ++launches;
    return layer_norm_rows_batched(x, C, rows, static_cast<int>(rows), C, g, b, 1e-5f, 0, out, C, 0, st);
  }
  int ln32(const float* x, int64_t rows, int C, const float* g, const float* b, float* out32, int64_t ld32, __half* out16,
           int64_t ld16, cudaStream_t st) {
    layer_norm_f32_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(x, rows, C, g, b, 1e-5f, out32, ld32,
                                                                               out16, ld16);
    ++launches;
    FRT2_CUDA_OK(cudaGetLastError());
    return FRT2_OK;
  }
  struct StackBufs { __half *n16, *qkv16, *o16, *g16; float2* stats; };
  static size_t stack_bytes(const EncStack& s, int64_t M, size_t (&o)[5]) {
    size_t off = 0;
    auto take = [&](size_t b) { const size_t at = off; off += (b + 255) & ~static_cast<size_t>(255); return at; };
    o[0] = take(M * s.E * 2);
    o[1] = take(M * 3 * s.H * s.hdp * 2);
    o[2] = take(M * s.H * s.hdp * 2);
    o[3] = take(M * s.F * 2);
    o[4] = take(M * sizeof(float2));
    return off;
  }
  static StackBufs stack_bufs(uint8_t* base, const size_t (&o)[5]) {
    return StackBufs{reinterpret_cast<__half*>(base + o[0]), reinterpret_cast<__half*>(base + o[1]),
                     reinterpret_cast<__half*>(base + o[2]), reinterpret_cast<__half*>(base + o[3]),
                     reinterpret_cast<float2*>(base + o[4])};
  }
  int fold_ln(const std::vector<float>& W, const std::vector<float>& bias, const std::vector<float>& gamma,
              const std::vector<float>& beta, int64_t Nn, int64_t Kk, __half** w_f, float** colsum, float** bias_f) {
    std::vector<float> Wf(static_cast<size_t>(Nn) * Kk), cs(Nn), bf(Nn);
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < Nn; ++n) {
      double sacc = 0.0, bacc = bias[n];
      for (int64_t k = 0; k < Kk; ++k) {
        const float w = W[n * Kk + k];
        const float wf = w * gamma[k];
        Wf[n * Kk + k] = wf;
        // the column sum runs over the ROUNDED fp16 weights the tensor cores see
        sacc += static_cast<double>(__half2float(__float2half_rn(std::min(65504.0f, std::max(-65504.0f, wf)))));
        bacc += static_cast<double>(beta[k]) * w;
      }
      cs[n] = static_cast<float>(sacc);
      bf[n] = static_cast<float>(bacc);
    }
    FRT2_TRY(up16(Wf, w_f));
    FRT2_TRY(up32(cs, colsum));
    FRT2_TRY(up32(bf, bias_f));
    return FRT2_OK;
  }
  int run_stack(const EncStack& s, float* x32, int B, int T, const StackBufs& w, cudaStream_t st);
  int run_front(const EncFront& f, const EncStack& s, const __half* mel16, int B, int Tm, __half* c1, float* x32,
                cudaStream_t st);
  int mel(const float* audio, int64_t pitch, int B, int64_t n, float* logmel, int* item_max, float* out32, __half* out16,
          __half* fr_s, float* spec, cudaStream_t st);
  int downstream(const __half* ssl16, int64_t M, int B, int T, uint8_t* base, float* vq_in, cudaStream_t st,
                 const float* aco32);
  int features(const float* ssl, const float* aco_feats, int B, int T, float* vq_in, cudaStream_t st);
  int audio_features(const float* audio, int64_t pitch, int B, int64_t n, float* vq_in, float* mel_out, float* ssl_out32,
                     float* aco_out32, cudaStream_t st);
};

int Encoder::ensure_ws(size_t bytes) {
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

int Encoder::load_stack(const std::string& prefix, const std::string& ln_key, EncStack& s) {
  const int64_t E = s.E, F = s.F, H = s.H, hd = s.hd, hdp = s.hdp, EP = H * hdp;
  const HostT *w, *b;
  for (size_t i = 0; i < s.layers.size(); ++i) {
    const std::string p = prefix + "layers." + std::to_string(i) + ".";
    EncLayer& L = s.layers[i];
    const HostT *wq, *bq, *wk, *wv, *bv;
    FRT2_TRY(need(p + "self_attn.q_proj.weight", &wq, {E, E}));
    FRT2_TRY(need(p + "self_attn.q_proj.bias", &bq, {E}));
    FRT2_TRY(need(p + "self_attn.k_proj.weight", &wk, {E, E}));     // no bias (whisper.py:37)
    FRT2_TRY(need(p + "self_attn.v_proj.weight", &wv, {E, E}));
    FRT2_TRY(need(p + "self_attn.v_proj.bias", &bv, {E}));
    std::vector<float> wqkv(static_cast<size_t>(3 * EP * E), 0.f), bqkv(static_cast<size_t>(3 * EP), 0.f);
    const HostT* ws3[3] = {wq, wk, wv};
    const HostT* bs3[3] = {bq, nullptr, bv};
    for (int sec = 0; sec < 3; ++sec)
      for (int64_t h = 0; h < H; ++h)
        for (int64_t d = 0; d < hd; ++d) {
          const int64_t dst = sec * EP + h * hdp + d, src = h * hd + d;
          std::copy(ws3[sec]->data.begin() + src * E, ws3[sec]->data.begin() + (src + 1) * E, wqkv.begin() + dst * E);
          if (bs3[sec] != nullptr) bqkv[dst] = bs3[sec]->data[src];
        }
    FRT2_TRY(up16(wqkv, &L.w_qkv));
    FRT2_TRY(up32(bqkv, &L.b_qkv));
    {
      const HostT *g1, *b1;
      FRT2_TRY(need(p + "self_attn_layer_norm.weight", &g1, {E}));
      FRT2_TRY(need(p + "self_attn_layer_norm.bias", &b1, {E}));
      FRT2_TRY(fold_ln(wqkv, bqkv, g1->data, b1->data, 3 * EP, E, &L.w_qkv_f, &L.s_qkv, &L.c_qkv));
    }
    FRT2_TRY(need(p + "self_attn.out_proj.weight", &w, {E, E}));
    FRT2_TRY(need(p + "self_attn.out_proj.bias", &b, {E}));
    std::vector<float> wo(static_cast<size_t>(E * EP), 0.f);
    for (int64_t o = 0; o < E; ++o)
      for (int64_t h = 0; h < H; ++h)
        for (int64_t d = 0; d < hd; ++d) wo[o * EP + h * hdp + d] = w->data[o * E + h * hd + d];
    FRT2_TRY(up16(wo, &L.w_o));
    FRT2_TRY(up32(b->data, &L.b_o));
    FRT2_TRY(need(p + "fc1.weight", &w, {F, E}));
    FRT2_TRY(need(p + "fc1.bias", &b, {F}));
    FRT2_TRY(up16(w->data, &L.w_fc1));
    FRT2_TRY(up32(b->data, &L.b_fc1));
    {
      const HostT *g2, *b2;
      FRT2_TRY(need(p + "final_layer_norm.weight", &g2, {E}));
      FRT2_TRY(need(p + "final_layer_norm.bias", &b2, {E}));
      FRT2_TRY(fold_ln(w->data, b->data, g2->data, b2->data, F, E, &L.w_fc1_f, &L.s_fc1, &L.c_fc1));
    }
    FRT2_TRY(need(p + "fc2.weight", &w, {E, F}));
    FRT2_TRY(need(p + "fc2.bias", &b, {E}));
    FRT2_TRY(up16(w->data, &L.w_fc2));
    FRT2_TRY(up32(b->data, &L.b_fc2));
    FRT2_TRY(need(p + "self_attn_layer_norm.weight", &w, {E}));
    FRT2_TRY(need(p + "self_attn_layer_norm.bias", &b, {E}));
    FRT2_TRY(up32(w->data, &L.ln1_g));
    FRT2_TRY(up32(b->data, &L.ln1_b));
    FRT2_TRY(need(p + "final_layer_norm.weight", &w, {E}));
    FRT2_TRY(need(p + "final_layer_norm.bias", &b, {E}));
    FRT2_TRY(up32(w->data, &L.ln2_g));
    FRT2_TRY(up32(b->data, &L.ln2_b));
  }
  FRT2_TRY(need(prefix + ln_key + ".weight", &w, {E}));
  FRT2_TRY(need(prefix + ln_key + ".bias", &b, {E}));
  FRT2_TRY(up32(w->data, &s.lnf_g));
  FRT2_TRY(up32(b->data, &s.lnf_b));
  return FRT2_OK;
}

int Encoder::load_front(const std::string& prefix, int E, EncFront& f) {
  const HostT *w, *b;
  const int64_t C = f.in_dim;
  FRT2_TRY(need(prefix + "conv1.weight", &w, {E, C, 3}));
  FRT2_TRY(need(prefix + "conv1.bias", &b, {E}));
  {   // (out, in, k) -> (out, k*in + c): tap j reads input row t - 1 + j
    std::vector<float> wf(static_cast<size_t>(E) * 3 * C);
    for (int64_t o = 0; o < E; ++o)
      for (int64_t c = 0; c < C; ++c)
        for (int j = 0; j < 3; ++j) wf[(o * 3 + j) * C + c] = w->data[(o * C + c) * 3 + j];
    FRT2_TRY(up16(wf, &f.w_c1));
    FRT2_TRY(up32(b->data, &f.b_c1));
  }
  FRT2_TRY(need(prefix + "conv2.weight", &w, {E, E, 3}));
  FRT2_TRY(need(prefix + "conv2.bias", &b, {E}));
  {   // stride 2 on the paired-frame view x2[t] = [x[2t] | x[2t+1]]: out[t] = W0 x[2t-1] + W1 x[2t] + W2 x[2t+1]
      //   = tap 0 over x2[t-1] with [0 | W0]  +  tap 1 over x2[t] with [W1 | W2]
    const int64_t Ee = E;
    std::vector<float> wf(static_cast<size_t>(Ee) * 4 * Ee, 0.f);
    for (int64_t o = 0; o < Ee; ++o)
      for (int64_t c = 0; c < Ee; ++c) {
        wf[o * 4 * Ee + 1 * Ee + c] = w->data[(o * Ee + c) * 3 + 0];
        wf[o * 4 * Ee + 2 * Ee + c] = w->data[(o * Ee + c) * 3 + 1];
        wf[o * 4 * Ee + 3 * Ee + c] = w->data[(o * Ee + c) * 3 + 2];
      }
    FRT2_TRY(up16(wf, &f.w_c2));
    FRT2_TRY(up32(b->data, &f.b_c2));
  }
  FRT2_TRY(need(prefix + "embed_positions.weight", &w, {f.max_pos, E}));
  FRT2_TRY(up32(w->data, &f.pos));
  return FRT2_OK;
}

// slaney mel scale (reference codec/audio.py:24-75, mel_scale="slaney")
static double hz_to_mel_slaney(double f) {
  const double min_log_hertz = 1000.0, min_log_mel = 15.0, logstep = 27.0 / std::log(6.4);
  return f >= min_log_hertz ? min_log_mel + std::log(f / min_log_hertz) * logstep : 3.0 * f / 200.0;
}
static double mel_to_hz_slaney(double m) {
  const double min_log_hertz = 1000.0, min_log_mel = 15.0, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hertz * std::exp(logstep * (m - min_log_mel)) : 200.0 * m / 3.0;
}

int Encoder::finalize() {
  FRT2_CUDA_OK(cudaSetDevice(device));
  FRT2_TRY(gemm_tc_init());
  const HostT *w, *b;
  const int64_t E = ada.E, P = static_cast<int64_t>(pool) * D;
  FRT2_TRY(need("ssl_adaptor.in_proj.weight", &w, {E, ssl_in}));
  FRT2_TRY(need("ssl_adaptor.in_proj.bias", &b, {E}));
  FRT2_TRY(up16(w->data, &w_in));
  FRT2_TRY(up32(b->data, &b_in));
  FRT2_TRY(load_stack("ssl_adaptor.", "layer_norm", ada));
  FRT2_TRY(need("ssl_adaptor.out_proj.weight", &w, {ssl_out, E}));
  FRT2_TRY(need("ssl_adaptor.out_proj.bias", &b, {ssl_out}));
  FRT2_TRY(up16(w->data, &w_out));
  FRT2_TRY(up32(b->data, &b_out));
  // ResidualDownConv: Conv1d weights (out = P, in = D, k = pool) -> GEMM rows over the (T/pool, pool*D) view whose
  // column index is k*D + c (time-major rows: `pool` consecutive frames side by side); gate rows first, then up rows
  {
    const HostT *wg, *wu;
    FRT2_TRY(need("downsample.gate_proj.weight", &wg, {P, D, pool}));
    FRT2_TRY(need("downsample.up_proj.weight", &wu, {P, D, pool}));
    std::vector<float> gu(static_cast<size_t>(2 * P * P));
    for (int half = 0; half < 2; ++half) {
      const std::vector<float>& src = half == 0 ? wg->data : wu->data;
#pragma omp parallel for schedule(static)
      for (long long o = 0; o < P; ++o)
        for (int64_t c = 0; c < D; ++c)
          for (int k = 0; k < pool; ++k)
            gu[(half * P + o) * P + static_cast<int64_t>(k) * D + c] = src[(o * D + c) * pool + k];
    }
    FRT2_TRY(up16(gu, &w_gu));
  }
  FRT2_TRY(need("downsample.down_proj.weight", &w, {P, P}));
  FRT2_TRY(up16(w->data, &w_down));
  FRT2_TRY(need("downsample.layer_norm.weight", &w, {P}));
  FRT2_TRY(need("downsample.layer_norm.bias", &b, {P}));
  FRT2_TRY(up32(w->data, &dln_g));
  FRT2_TRY(up32(b->data, &dln_b));
  FRT2_TRY(need("downsample.out_proj.weight", &w, {D, P}));
  FRT2_TRY(need("downsample.out_proj.bias", &b, {D}));
  FRT2_TRY(up16(w->data, &w_dout));
  FRT2_TRY(up32(b->data, &b_dout));
  if (has_front) {
    FRT2_TRY(load_front("ssl.", ssl_stack.E, ssl_front));
    FRT2_TRY(load_stack("ssl.", "layer_norm", ssl_stack));
    FRT2_TRY(load_front("acoustic_encoder.", aco_stack.E, aco_front));
    FRT2_TRY(load_stack("acoustic_encoder.", "layer_norm", aco_stack));
    // torch.hann_window(n_fft) (periodic) and the slaney-normalised slaney-scale bank of WhisperMelExtractor
    // (whisper.py:276-279,266-274; codec/audio.py:102-148), computed in double like the numpy reference, used as fp32
    const int bins = n_fft / 2 + 1;
    std::vector<float> win(n_fft), bank(static_cast<size_t>(bins) * n_mels);
    for (int j = 0; j < n_fft; ++j) win[j] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * M_PI * j / n_fft));
    const double fmin = 0.0, fmax = 8000.0, sr = 16000.0;
    const double mel_min = hz_to_mel_slaney(fmin), mel_max = hz_to_mel_slaney(fmax);
    std::vector<double> ff(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) ff[i] = mel_to_hz_slaney(mel_min + (mel_max - mel_min) * i / (n_mels + 1));
    for (int k = 0; k < bins; ++k) {
      const double fk = (sr / 2.0) * k / (bins - 1);            // np.linspace(0, sr // 2, bins)
      for (int m = 0; m < n_mels; ++m) {
        const double down = -(ff[m] - fk) / (ff[m + 1] - ff[m]);
        const double up = (ff[m + 2] - fk) / (ff[m + 2] - ff[m + 1]);
        const double tri = std::max(0.0, std::min(down, up));
        bank[static_cast<size_t>(k) * n_mels + m] = static_cast<float>(tri * (2.0 / (ff[m + 2] - ff[m])));
      }
    }
    FRT2_TRY(up32(win, &mel_window));
    FRT2_TRY(up32(bank, &mel_bank));
    // twiddles of the real DFT as GEMM weights: row k < bins = cos(2 pi k j / n_fft), row bins + k = sin(...), exact phase
    mel_kp = (n_fft + 63) / 64 * 64;
    std::vector<float> tw(static_cast<size_t>(2 * bins) * mel_kp, 0.f);
    for (int k = 0; k < bins; ++k)
      for (int j = 0; j < n_fft; ++j) {
        const double ang = 2.0 * M_PI * static_cast<double>((static_cast<long long>(k) * j) % n_fft) / n_fft;
        tw[static_cast<size_t>(k) * mel_kp + j] = static_cast<float>(std::cos(ang));
        tw[static_cast<size_t>(bins + k) * mel_kp + j] = static_cast<float>(std::sin(ang));
      }
    std::vector<__half> tws(tw.size() * 3);
    rvq_split_weight_host(tw.data(), 2 * bins, mel_kp, tws.data());
    FRT2_TRY(dev_alloc(reinterpret_cast<void**>(&mel_dft_s), tws.size() * 2));
    FRT2_CUDA_OK(cudaMemcpy(mel_dft_s, tws.data(), tws.size() * 2, cudaMemcpyHostToDevice));
  }
  FRT2_TRY(dev_alloc(reinterpret_cast<void**>(&sched), 16));
  FRT2_CUDA_OK(cudaMemset(sched, 0, 16));
  FRT2_CUDA_OK(cudaDeviceSynchronize());
  raw.clear();
  finalized = true;
  return FRT2_OK;
}

// x32 (B*T, E) fp32 residual stream, updated in place; leaves nothing else behind (the caller applies the final LayerNorm).
// LayerNorm is folded across the GEMMs like on the decode side: the out-projection and fc2 also write the fp16 copy of the
// rows they produce, row_stats reads that copy (2 B per element instead of LayerNorm's 4 in + 2 out) and the q|k|v / fc1
// GEMM finishes the normalisation in its epilogue.  Only the first LayerNorm of a stack (its rows come from a producer
// without that copy) and the final one are kernels.  FRT2_ENC_NO_LNFOLD=1: every LayerNorm as a kernel (A/B).
int Encoder::run_stack(const EncStack& s, float* x32, int B, int T, const StackBufs& w, cudaStream_t st) {
  const int64_t M = static_cast<int64_t>(B) * T, E = s.E, EP = static_cast<int64_t>(s.H) * s.hdp;
  static const bool no_fold = getenv("FRT2_ENC_NO_LNFOLD") != nullptr && atoi(getenv("FRT2_ENC_NO_LNFOLD")) != 0;
  const bool fold = !no_fold && s.E <= 2048;     // the same arithmetic whatever the batch (rows are never packed here)
  auto stats = [&]() {
    ++launches;
    return row_stats(w.n16, E, M, s.E, 1e-5f, w.stats, st, nullptr);
  };
  const size_t nl = s.layers.size();
  for (size_t i = 0; i < nl; ++i) {
    const EncLayer& L = s.layers[i];
    if (fold && i > 0) {     // n16 holds the fp16 copy fc2 of the previous layer wrote
      FRT2_TRY(stats());
      FRT2_TRY(gemm(w.n16, M, s.E, L.w_qkv_f, static_cast<int>(3 * EP), L.c_qkv, ACT_NONE, nullptr, nullptr, 0, w.qkv16,
                    3 * EP, st, nullptr, w.stats, L.s_qkv));
    } else {
      FRT2_TRY(ln16(x32, M, s.E, L.ln1_g, L.ln1_b, w.n16, st));
      FRT2_TRY(gemm(w.n16, M, s.E, L.w_qkv, static_cast<int>(3 * EP), L.b_qkv, ACT_NONE, nullptr, nullptr, 0, w.qkv16, 3 * EP, st));
    }
    AttnDesc a{};
    a.B = B; a.H = s.H; a.hd = s.hdp; a.Tq = T; a.Tk = T; a.q_pos0 = 0; a.block_causal = 0;
    a.q = w.qkv16; a.q_row_pitch = 3 * EP; a.q_batch_pitch = static_cast<int64_t>(T) * 3 * EP;
    a.k = w.qkv16 + EP; a.v = w.qkv16 + 2 * EP; a.kv_row_pitch = 3 * EP; a.kv_batch_pitch = a.q_batch_pitch;
    a.out = w.o16; a.o_row_pitch = EP; a.o_batch_pitch = static_cast<int64_t>(T) * EP;
    a.scale = 1.0f / std::sqrt(static_cast<float>(s.hd));     // the TRUE head dim (padding adds zeros only)
    a.sched = sched;
    ++launches;
    if ((s.hdp == 64 || s.hdp == 128) && T >= 32) FRT2_TRY(attention_tc(a, st));
    else FRT2_TRY(attention_warp(a, st));
    FRT2_TRY(gemm(w.o16, M, static_cast<int>(EP), L.w_o, s.E, L.b_o, ACT_NONE, x32, x32, E, nullptr, 0, st,
                  fold ? w.n16 : nullptr));
    if (fold) {
      FRT2_TRY(stats());
      FRT2_TRY(gemm(w.n16, M, s.E, L.w_fc1_f, s.F, L.c_fc1, ACT_GELU, nullptr, nullptr, 0, w.g16, s.F, st, nullptr, w.stats,
                    L.s_fc1));
    } else {
      FRT2_TRY(ln16(x32, M, s.E, L.ln2_g, L.ln2_b, w.n16, st));
      FRT2_TRY(gemm(w.n16, M, s.E, L.w_fc1, s.F, L.b_fc1, ACT_GELU, nullptr, nullptr, 0, w.g16, s.F, st));
    }
    FRT2_TRY(gemm(w.g16, M, s.F, L.w_fc2, s.E, L.b_fc2, ACT_NONE, x32, x32, E, nullptr, 0, st,
                  (fold && i + 1 < nl) ? w.n16 : nullptr));
  }
  return FRT2_OK;
}

// mel16 (B, Tm, in_dim) fp16 -> x32 (B, Tm/2, E) fp32 = GELU(conv2(GELU(conv1(mel)))) + positions (whisper.py:229-242)
int Encoder::run_front(const EncFront& f, const EncStack& s, const __half* mel16, int B, int Tm, __half* c1, float* x32,
                       cudaStream_t st) {
  const int E = s.E, T = Tm / 2;
  {
    GemmDesc g{};
    g.A = mel16; g.a_row_pitch = f.in_dim; g.a_batch_pitch = static_cast<int64_t>(Tm) * f.in_dim; g.rows_a = Tm;
    g.batches = B; g.Kc = f.in_dim; g.ntaps = 3; g.row_shift = -1;     // padding = 1: tap j reads row t - 1 + j
    g.W = f.w_c1; g.N = E; g.rows_out = Tm; g.alpha = 1.0f; g.bias = f.b_c1; g.act = ACT_GELU;
    g.out16 = c1; g.ld16 = E; g.pitch16 = static_cast<int64_t>(Tm) * E;
    ++launches;
    FRT2_TRY(gemm_tc(g, st));
  }
  {
    GemmDesc g{};
    g.A = c1; g.a_row_pitch = 2 * E; g.a_batch_pitch = static_cast<int64_t>(T) * 2 * E; g.rows_a = T;
    g.batches = B; g.Kc = 2 * E; g.ntaps = 2; g.row_shift = -1;
    g.W = f.w_c2; g.N = E; g.rows_out = T; g.alpha = 1.0f; g.bias = f.b_c2; g.act = ACT_GELU;
    g.out32 = x32; g.ld32 = E; g.pitch32 = static_cast<int64_t>(T) * E;
    ++launches;
    FRT2_TRY(gemm_tc(g, st));
  }
  const long long n = static_cast<long long>(B) * T * (E / 4);
  add_pos_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x32, f.pos, static_cast<long long>(B) * T, T, E / 4);
  ++launches;
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

int Encoder::mel(const float* audio, int64_t pitch, int B, int64_t n, float* logmel, int* item_max, float* out32,
                 __half* out16, __half* fr_s, float* spec, cudaStream_t st) {
  const int T = static_cast<int>(n / hop);
  FRT2_CUDA_OK(cudaMemsetAsync(item_max, 0, static_cast<size_t>(B) * sizeof(int), st));
  const int bins = n_fft / 2 + 1;
  static const bool simt = getenv("FRT2_MEL_SIMT") != nullptr && atoi(getenv("FRT2_MEL_SIMT")) != 0;   // checker / A-B
  if (simt) {
    const size_t smem = static_cast<size_t>(2 * n_fft + MEL_FR * n_fft + MEL_FR * bins) * sizeof(float);
    mel_power_kernel<<<dim3((T + MEL_FR - 1) / MEL_FR, B), 256, smem, st>>>(audio, pitch, n, T, n_fft, hop, mel_window,
                                                                          mel_bank, n_mels, logmel, item_max);
    FRT2_CUDA_OK(cudaGetLastError());
    launches += 1;
  } else {
    const int KP = mel_kp, ld = (2 * bins + 3) / 4 * 4;
    const long long per_item = static_cast<long long>(T) * KP;
    mel_frames_kernel<<<dim3(static_cast<unsigned>((per_item + 255) / 256), B), 256, 0, st>>>(audio, pitch, n, T, n_fft, hop,
                                                                                            KP, mel_window, fr_s);
    FRT2_CUDA_OK(cudaGetLastError());
    GemmDesc g{};
    const int64_t Mm = static_cast<int64_t>(B) * T;
    g.A = fr_s; g.a_row_pitch = 3 * KP; g.rows_a = static_cast<int>(Mm); g.batches = 1; g.Kc = 3 * KP; g.ntaps = 1;
    g.W = mel_dft_s; g.N = 2 * bins; g.rows_out = static_cast<int>(Mm); g.alpha = 1.0f; g.act = ACT_NONE;
    g.out32 = spec; g.ld32 = ld;
    FRT2_TRY(gemm_tc(g, st));
    mel_bank_kernel<<<dim3((T + MEL_FR - 1) / MEL_FR, B), 256, static_cast<size_t>(MEL_FR) * bins * sizeof(float), st>>>(
        spec, ld, T, bins, mel_bank, n_mels, logmel, item_max);
    FRT2_CUDA_OK(cudaGetLastError());
    launches += 3;
  }
  const long long total = static_cast<long long>(B) * T * n_mels;
  mel_norm_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(logmel, item_max, static_cast<long long>(T) * n_mels,
                                                                            total, out32, out16);
  FRT2_CUDA_OK(cudaGetLastError());
  launches += 1;
  return FRT2_OK;
}

// SslAdaptor + cat + ResidualDownConv.  ssl16: (M, ssl_in) fp16 SSL features; the acoustic features are already in the
// concatenated buffers (aco32 == nullptr) or are copied in from aco32 (M, aco) fp32.  `base` = workspace behind ssl16.
int Encoder::downstream(const __half* ssl16, int64_t M, int B, int T, uint8_t* base, float* vq_in, cudaStream_t st,
                        const float* aco32) {
  const int64_t M4 = M / pool, P = static_cast<int64_t>(pool) * D, E = ada.E;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~static_cast<size_t>(255); return base + o; };
  float* cat32 = reinterpret_cast<float*>(take(M * D * 4));      // FIRST: audio_features() pre-fills the acoustic half
  __half* cat16 = reinterpret_cast<__half*>(take(M * D * 2));
  float* x32 = reinterpret_cast<float*>(take(M * E * 4));
  size_t so[5];
  const size_t sb = stack_bytes(ada, M, so);
  uint8_t* sbase = take(sb);
  StackBufs w = stack_bufs(sbase, so);
  __half* gu16 = reinterpret_cast<__half*>(take(M4 * 2 * P * 2));
  __half* act16 = reinterpret_cast<__half*>(take(M4 * P * 2));
  float* c32 = reinterpret_cast<float*>(take(M4 * P * 4));
  __half* cn16 = reinterpret_cast<__half*>(take(M4 * P * 2));
  // ---- SslAdaptor (model.py:53-66) ----
  FRT2_TRY(gemm(ssl16, M, ssl_in, w_in, ada.E, b_in, ACT_NONE, nullptr, x32, E, nullptr, 0, st));
  FRT2_TRY(run_stack(ada, x32, B, T, w, st));
  FRT2_TRY(ln16(x32, M, ada.E, ada.lnf_g, ada.lnf_b, w.n16, st));
  // out_proj writes the semantic half of the concatenated features (fp32 for the residual of model.py:117, fp16 as the
  // operand of the gate / up convolutions); the acoustic half sits beside it: torch.cat without a pass of its own
  FRT2_TRY(gemm(w.n16, M, ada.E, w_out, ssl_out, b_out, ACT_NONE, nullptr, cat32, D, cat16, D, st));
  if (aco32 != nullptr) FRT2_TRY(cvt(aco32, aco, M, aco, cat32 + ssl_out, D, cat16 + ssl_out, D, st));
  // ---- ResidualDownConv (model.py:106-121) on the (M/pool, pool*D) view ----
  FRT2_TRY(gemm(cat16, M4, static_cast<int>(P), w_gu, static_cast<int>(2 * P), nullptr, ACT_NONE, nullptr, nullptr, 0,
                gu16, 2 * P, st));
  {
    const long long n = M4 * (P / 8);
    silu_mul_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(gu16, M4, static_cast<int>(P / 8), act16);
    ++launches;
    FRT2_CUDA_OK(cudaGetLastError());
  }
  FRT2_TRY(gemm(act16, M4, static_cast<int>(P), w_down, static_cast<int>(P), nullptr, ACT_NONE, cat32, c32, P, nullptr, 0, st));
  FRT2_TRY(ln16(c32, M4, static_cast<int>(P), dln_g, dln_b, cn16, st));
  FRT2_TRY(gemm(cn16, M4, static_cast<int>(P), w_dout, D, b_dout, ACT_NONE, nullptr, vq_in, D, nullptr, 0, st));
  return FRT2_OK;
}

static size_t downstream_bytes(const Encoder& e, int64_t M) {
  const int64_t M4 = M / e.pool, P = static_cast<int64_t>(e.pool) * e.D;
  size_t so[5];
  auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  return al(M * e.D * 4) + al(M * e.D * 2) + al(M * e.ada.E * 4) + al(Encoder::stack_bytes(e.ada, M, so)) +
         al(M4 * 2 * P * 2) + al(M4 * P * 2) + al(M4 * P * 4) + al(M4 * P * 2);
}

int Encoder::features(const float* ssl, const float* aco_feats, int B, int T, float* vq_in, cudaStream_t st) {
  const int64_t M = static_cast<int64_t>(B) * T;
  const size_t ssl_bytes = (static_cast<size_t>(M) * ssl_in * 2 + 255) & ~static_cast<size_t>(255);
  FRT2_TRY(ensure_ws(ssl_bytes + downstream_bytes(*this, M)));
  FRT2_TRY(begin(st));
  __half* ssl16 = reinterpret_cast<__half*>(ws);
  FRT2_TRY(cvt(ssl, ssl_in, M, ssl_in, nullptr, 0, ssl16, ssl_in, st));
  FRT2_TRY(downstream(ssl16, M, B, T, ws + ssl_bytes, vq_in, st, aco_feats));
  return end(st);
}

// audio (B, n) fp32 16 kHz, n a multiple of hop * 2 * pool -> vq_in (B, n / (hop*2*pool), D).  Optional parity taps:
// mel_out (B, n/hop, n_mels), ssl_out32 (B, T, ssl_in), aco_out32 (B, T, aco), all fp32.
int Encoder::audio_features(const float* audio, int64_t pitch, int B, int64_t n, float* vq_in, float* mel_out,
                            float* ssl_out32, float* aco_out32, cudaStream_t st) {
  const int Tm = static_cast<int>(n / hop), T = Tm / 2;
  const int64_t Mm = static_cast<int64_t>(B) * Tm, M = static_cast<int64_t>(B) * T;
  auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  // layout: [ssl16 | downstream arena (cat32 first) | encoder scratch]
  const size_t ssl_bytes = al(M * ssl_in * 2), down_bytes = downstream_bytes(*this, M);
  size_t so_s[5], so_a[5];
  const size_t sb_s = stack_bytes(ssl_stack, M, so_s), sb_a = stack_bytes(aco_stack, M, so_a);
  const int Emax = std::max(ssl_stack.E, aco_stack.E);
  // the split frame rows / DFT output of the log-mel front end are dead once mel16 exists: they share the bytes of the
  // encoders' scratch (c1, x32, layer-stack buffers)
  const int spec_ld = (n_fft + 2 + 3) / 4 * 4;
  const size_t mel_tmp = al(Mm * 3 * static_cast<size_t>(mel_kp) * 2) + al(Mm * static_cast<size_t>(spec_ld) * 4);
  const size_t o_logmel = 0, o_imax = o_logmel + al(Mm * n_mels * 4), o_mel16 = o_imax + al(B * 4),
               o_c1 = o_mel16 + al(Mm * n_mels * 2), o_x32 = o_c1 + al(Mm * Emax * 2), o_stack = o_x32 + al(M * Emax * 4),
               scratch = std::max(o_stack + std::max(sb_s, sb_a), o_c1 + mel_tmp);
  FRT2_TRY(ensure_ws(ssl_bytes + down_bytes + scratch));
  FRT2_TRY(begin(st));
  __half* ssl16 = reinterpret_cast<__half*>(ws);
  uint8_t* down = ws + ssl_bytes;
  uint8_t* sc = down + down_bytes;
  float* logmel = reinterpret_cast<float*>(sc + o_logmel);
  int* imax = reinterpret_cast<int*>(sc + o_imax);
  __half* mel16 = reinterpret_cast<__half*>(sc + o_mel16);
  __half* c1 = reinterpret_cast<__half*>(sc + o_c1);
  float* x32 = reinterpret_cast<float*>(sc + o_x32);
  uint8_t* sbase = sc + o_stack;
  float* cat32 = reinterpret_cast<float*>(down);                                // downstream()'s first two buffers
  __half* cat16 = reinterpret_cast<__half*>(down + al(M * D * 4));
  FRT2_TRY(mel(audio, pitch, B, n, logmel, imax, mel_out, mel16, reinterpret_cast<__half*>(sc + o_c1),
               reinterpret_cast<float*>(sc + o_c1 + al(Mm * 3 * static_cast<size_t>(mel_kp) * 2)), st));
  // semantic encoder: its final LayerNorm output is only ever a GEMM operand (ssl_adaptor.in_proj) -> fp16
  FRT2_TRY(run_front(ssl_front, ssl_stack, mel16, B, Tm, c1, x32, st));
  {
    StackBufs w = stack_bufs(sbase, so_s);
    FRT2_TRY(run_stack(ssl_stack, x32, B, T, w, st));
    FRT2_TRY(ln32(x32, M, ssl_stack.E, ssl_stack.lnf_g, ssl_stack.lnf_b, ssl_out32, ssl_in, ssl16, ssl_in, st));
  }
  // acoustic encoder: its output is the right half of the concatenated features (fp32 residual + fp16 operand)
  FRT2_TRY(run_front(aco_front, aco_stack, mel16, B, Tm, c1, x32, st));
  {
    StackBufs w = stack_bufs(sbase, so_a);
    FRT2_TRY(run_stack(aco_stack, x32, B, T, w, st));
    FRT2_TRY(ln32(x32, M, aco_stack.E, aco_stack.lnf_g, aco_stack.lnf_b, cat32 + ssl_out, D, cat16 + ssl_out, D, st));
    if (aco_out32 != nullptr)
      FRT2_CUDA_OK(cudaMemcpy2DAsync(aco_out32, static_cast<size_t>(aco) * 4, cat32 + ssl_out, static_cast<size_t>(D) * 4,
                                     static_cast<size_t>(aco) * 4, M, cudaMemcpyDeviceToDevice, st));
  }
  FRT2_TRY(downstream(ssl16, M, B, T, down, vq_in, st, nullptr));
  return end(st);
}

}  // namespace frt2

using namespace frt2;

struct frt2_encoder { Encoder e; };

static int padded_head_dim(int hd) { return hd == 32 ? 32 : (hd <= 64 ? 64 : 128); }

static int init_stack(EncStack& s, int E, int H, int F, int layers, const char* what) {
  FRT2_REQUIRE(E > 0 && E % 64 == 0, FRT2_ERR_BAD_ARG, std::string(what) + ": embed_dim must be a positive multiple of 64");
  FRT2_REQUIRE(H > 0 && E % H == 0, FRT2_ERR_BAD_ARG, std::string(what) + ": embed_dim must be divisible by num_heads");
  FRT2_REQUIRE(E / H <= 128 && (E / H) % 8 == 0, FRT2_ERR_BAD_ARG, std::string(what) + ": head_dim must be a multiple of 8, at most 128");
  FRT2_REQUIRE(layers >= 0, FRT2_ERR_BAD_ARG, std::string(what) + ": negative num_layers");
  s.E = E; s.H = H; s.hd = E / H; s.hdp = padded_head_dim(s.hd);
  s.F = F > 0 ? F : 4 * E;                                   // whisper.py:137
  FRT2_REQUIRE(s.F % 64 == 0, FRT2_ERR_BAD_ARG, std::string(what) + ": ffn_dim must be a multiple of 64");
  s.layers.resize(layers);
  return FRT2_OK;
}

extern "C" {

int frt2_enc_create(const frt2_enc_config* cfg, int device, frt2_encoder** out) {
  FRT2_REQUIRE(cfg != nullptr && out != nullptr, FRT2_ERR_BAD_ARG, "frt2_enc_create: null argument");
  FRT2_REQUIRE(cfg->ssl_in_dim > 0 && cfg->ssl_in_dim % 64 == 0, FRT2_ERR_BAD_ARG,
               "frt2_enc_create: ssl_adaptor.in_dim must be a positive multiple of 64");
  FRT2_REQUIRE(cfg->ssl_out_dim > 0 && cfg->ssl_out_dim % 8 == 0 && cfg->aco_dim > 0 && cfg->aco_dim % 8 == 0,
               FRT2_ERR_BAD_ARG, "frt2_enc_create: ssl_adaptor.out_dim and acoustic_encoder.embed_dim must be multiples of 8");
  FRT2_REQUIRE(cfg->avg_pooler >= 1 && cfg->avg_pooler <= 8, FRT2_ERR_BAD_ARG, "frt2_enc_create: avg_pooler must be in [1, 8]");
  const int D = cfg->ssl_out_dim + cfg->aco_dim;
  FRT2_REQUIRE((static_cast<int64_t>(D) * cfg->avg_pooler) % 64 == 0, FRT2_ERR_BAD_ARG,
               "frt2_enc_create: avg_pooler * (out_dim + aco_dim) must be a multiple of 64");
  int ndev = 0;
  FRT2_CUDA_OK(cudaGetDeviceCount(&ndev));
  FRT2_REQUIRE(device >= 0 && device < ndev, FRT2_ERR_BAD_ARG, "frt2_enc_create: bad device index");
  cudaDeviceProp prop;
  FRT2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  FRT2_REQUIRE(prop.major == 10, FRT2_ERR_BAD_ARG, "frt2_enc_create: this library is sm_100a only (no fallback path)");
  auto* fe = new frt2_encoder();
  Encoder& e = fe->e;
  auto fail = [&](int rc) { delete fe; return rc; };
  e.device = device;
  e.ssl_in = cfg->ssl_in_dim; e.ssl_out = cfg->ssl_out_dim; e.aco = cfg->aco_dim; e.pool = cfg->avg_pooler; e.D = D;
  int rc = init_stack(e.ada, cfg->ssl_embed_dim, cfg->ssl_num_heads, cfg->ssl_ffn_dim, cfg->ssl_num_layers, "ssl_adaptor");
  if (rc != FRT2_OK) return fail(rc);
  e.has_front = cfg->ssl_enc_layers > 0 || cfg->aco_layers > 0;
  if (e.has_front) {
    rc = init_stack(e.ssl_stack, cfg->ssl_in_dim, cfg->ssl_enc_heads, cfg->ssl_enc_ffn_dim, cfg->ssl_enc_layers, "ssl encoder");
    if (rc == FRT2_OK)
      rc = init_stack(e.aco_stack, cfg->aco_dim, cfg->aco_heads, cfg->aco_ffn_dim, cfg->aco_layers, "acoustic_encoder");
    if (rc != FRT2_OK) return fail(rc);
    if (!(cfg->num_mels > 0 && cfg->num_mels % 64 == 0 && cfg->max_positions > 0 && cfg->aco_dim % 64 == 0)) {
      set_error("frt2_enc_create: num_mels and acoustic_encoder.embed_dim must be multiples of 64, max_positions > 0");
      return fail(FRT2_ERR_BAD_ARG);
    }
    e.n_mels = cfg->num_mels;
    e.ssl_front.in_dim = e.aco_front.in_dim = cfg->num_mels;
    e.ssl_front.max_pos = e.aco_front.max_pos = cfg->max_positions;
  }
  *out = fe;
  return FRT2_OK;
}

int frt2_enc_load_tensor(frt2_encoder* fe, const char* key, const float* data, int ndim, const int64_t* shape,
                         int on_device) {
  FRT2_REQUIRE(fe && key && data && shape && ndim >= 0 && ndim <= 4, FRT2_ERR_BAD_ARG, "frt2_enc_load_tensor: bad argument");
  Encoder& e = fe->e;
  FRT2_REQUIRE(!e.finalized, FRT2_ERR_BAD_ARG, "frt2_enc_load_tensor: already finalized");
  const std::string k(key);
  const bool mine = k.rfind("ssl_adaptor.", 0) == 0 || k.rfind("downsample.", 0) == 0 ||
                    (e.has_front && (k.rfind("ssl.", 0) == 0 || k.rfind("acoustic_encoder.", 0) == 0));
  if (!mine) return FRT2_OK;   // not part of the encode side
  HostT t;
  t.shape.assign(shape, shape + ndim);
  int64_t n = 1;
  for (auto d : t.shape) n *= d;
  FRT2_REQUIRE(n >= 0, FRT2_ERR_BAD_ARG, "frt2_enc_load_tensor: negative dimension");
  t.data.resize(n);
  if (on_device) {
    FRT2_CUDA_OK(cudaSetDevice(e.device));
    FRT2_CUDA_OK(cudaMemcpy(t.data.data(), data, n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::memcpy(t.data.data(), data, n * 4);
  }
  std::lock_guard<std::mutex> lk(e.mu);
  e.raw[k] = std::move(t);
  return FRT2_OK;
}

int frt2_enc_finalize(frt2_encoder* fe) {
  FRT2_REQUIRE(fe, FRT2_ERR_BAD_ARG, "null encoder");
  std::lock_guard<std::mutex> lk(fe->e.mu);
  FRT2_REQUIRE(!fe->e.finalized, FRT2_ERR_BAD_ARG, "already finalized");
  return fe->e.finalize();
}

void frt2_enc_destroy(frt2_encoder* fe) { delete fe; }

int frt2_enc_features(frt2_encoder* fe, const float* ssl, const float* aco, int B, int T, float* vq_in,
                      int64_t* launches, void* cuda_stream) {
  FRT2_REQUIRE(fe, FRT2_ERR_BAD_ARG, "null encoder");
  Encoder& e = fe->e;
  FRT2_REQUIRE(e.finalized, FRT2_ERR_NOT_FINALIZED, "encoder not finalized");
  FRT2_REQUIRE(ssl != nullptr && aco != nullptr && vq_in != nullptr, FRT2_ERR_BAD_ARG, "frt2_enc_features: null pointer");
  FRT2_REQUIRE(B >= 1 && T >= e.pool, FRT2_ERR_BAD_ARG, "frt2_enc_features: B >= 1 and T >= avg_pooler required");
  // x.reshape(batch_size, -1, intermediate_dim) (model.py:113) needs whole groups of `pooler` frames
  FRT2_REQUIRE(T % e.pool == 0, FRT2_ERR_BAD_ARG, "frt2_enc_features: T must be a multiple of avg_pooler (model.py:113)");
  FRT2_REQUIRE(e.ada.hdp != 32 || T % 8 == 0, FRT2_ERR_BAD_ARG, "frt2_enc_features: head_dim 32 needs T % 8 == 0");
  FRT2_REQUIRE((reinterpret_cast<uintptr_t>(ssl) & 15) == 0 && (reinterpret_cast<uintptr_t>(aco) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(vq_in) & 15) == 0,
               FRT2_ERR_BAD_ARG, "frt2_enc_features: pointers must be 16-byte aligned");
  std::lock_guard<std::mutex> lk(e.mu);
  FRT2_CUDA_OK(cudaSetDevice(e.device));
  const long long before = e.launches;
  const int rc = e.features(ssl, aco, B, T, vq_in, static_cast<cudaStream_t>(cuda_stream));
  if (launches != nullptr) *launches = e.launches - before;
  return rc;
}

int frt2_enc_audio_features(frt2_encoder* fe, const float* audio16k, int64_t audio_pitch, int B, int64_t n, float* vq_in,
                            float* mel_out, float* ssl_out, float* aco_out, int64_t* launches, void* cuda_stream) {
  FRT2_REQUIRE(fe, FRT2_ERR_BAD_ARG, "null encoder");
  Encoder& e = fe->e;
  FRT2_REQUIRE(e.finalized, FRT2_ERR_NOT_FINALIZED, "encoder not finalized");
  FRT2_REQUIRE(e.has_front, FRT2_ERR_MISSING_TENSOR,
               "frt2_enc_audio_features: the feature encoders (ssl.*, acoustic_encoder.*) were not configured");
  FRT2_REQUIRE(audio16k != nullptr && vq_in != nullptr && B >= 1, FRT2_ERR_BAD_ARG, "frt2_enc_audio_features: bad argument");
  const int64_t unit = static_cast<int64_t>(e.hop) * 2 * e.pool;
  FRT2_REQUIRE(n >= unit && n % unit == 0 && audio_pitch >= n, FRT2_ERR_BAD_ARG,
               "frt2_enc_audio_features: the sample count must be a positive multiple of hop*2*avg_pooler (1280): the "
               "reference pads every chunk to 6 s (model.py:262-275)");
  FRT2_REQUIRE(n / (2 * e.hop) <= e.ssl_front.max_pos, FRT2_ERR_BAD_ARG,
               "frt2_enc_audio_features: more frames than max_positions (whisper.py:222)");
  FRT2_REQUIRE(n > e.n_fft / 2, FRT2_ERR_BAD_ARG, "frt2_enc_audio_features: reflect padding needs more than n_fft/2 samples");
  std::lock_guard<std::mutex> lk(e.mu);
  FRT2_CUDA_OK(cudaSetDevice(e.device));
  const long long before = e.launches;
  const int rc = e.audio_features(audio16k, audio_pitch, B, n, vq_in, mel_out, ssl_out, aco_out,
                                  static_cast<cudaStream_t>(cuda_stream));
  if (launches != nullptr) *launches = e.launches - before;
  return rc;
}

}  // extern "C"
