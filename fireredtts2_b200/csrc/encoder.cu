// Codec ENCODE side behind the two feature encoders: everything `RedCodecInfer._encode_one_batch`
// (reference codec/model.py:218-236) runs between the Whisper encoders and `ResidualVQ.encode_codes`:
//
//   sem  = SslAdaptor(ssl)                      model.py:19-77   Linear -> N x WhisperEncoderLayer (full attention)
//                                                                 -> LayerNorm -> Linear
//   x    = cat([sem, aco], dim=2)               model.py:230
//   vq   = ResidualDownConv(x)                  model.py:80-121  gate/up Conv1d(k = s = pooler), SiLU(g)*u, down_proj,
//                                                                 LayerNorm(c + x.reshape), out_proj
//
// Built from the decode path's kernels: every Linear / strided conv is a tcgen05 GEMM (gemm_tc; the k = s = pooler
// convolutions are plain GEMMs on the (T/pooler, pooler*D) view of the time-major rows, gate and up share ONE launch),
// attention is the tcgen05 flash kernel with the mask switched off (make_nonpad_mask of full-length chunks,
// model.py:220-222), LayerNorm is the two-rows-per-warp kernel.  fp16 operands, fp32 accumulation / residual stream /
// statistics, exactly as on the decode side.  The result feeds frt2_rvq_encode.
#include <algorithm>
#include <cmath>
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

struct EncLayer {
  __half *w_qkv = nullptr, *w_o = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;
  float *b_qkv = nullptr, *b_o = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

}  // namespace

struct Encoder {
  int device = 0;
  std::mutex mu;
  std::map<std::string, HostT> raw;
  bool finalized = false;
  // config
  int ssl_in = 0, Es = 0, ssl_out = 0, nl = 0, H = 0, F = 0, aco = 0, pool = 4, D = 0, hd = 0;
  // weights
  std::vector<void*> owned;
  __half* w_in = nullptr;   float* b_in = nullptr;
  std::vector<EncLayer> layers;
  float *lnf_g = nullptr, *lnf_b = nullptr;
  __half* w_out = nullptr;  float* b_out = nullptr;
  __half* w_gu = nullptr;                            // (2*pool*D, pool*D): gate rows, then up rows, tap-major K
  __half* w_down = nullptr;                          // (pool*D, pool*D)
  float *dln_g = nullptr, *dln_b = nullptr;
  __half* w_dout = nullptr; float* b_dout = nullptr; // (D, pool*D)
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
  int finalize();
  int features(const float* ssl, const float* aco_feats, int B, int T, float* vq_in, cudaStream_t st);
};

int Encoder::finalize() {
  FRT2_CUDA_OK(cudaSetDevice(device));
  FRT2_TRY(gemm_tc_init());
  const HostT *w, *b;
  const int64_t E = Es, P = static_cast<int64_t>(pool) * D;
  FRT2_TRY(need("ssl_adaptor.in_proj.weight", &w, {E, ssl_in}));
  FRT2_TRY(need("ssl_adaptor.in_proj.bias", &b, {E}));
  FRT2_TRY(up16(w->data, &w_in));
  FRT2_TRY(up32(b->data, &b_in));
  layers.resize(nl);
  for (int i = 0; i < nl; ++i) {
    const std::string p = "ssl_adaptor.layers." + std::to_string(i) + ".";
    EncLayer& L = layers[i];
    const HostT *wq, *bq, *wk, *wv, *bv;
    FRT2_TRY(need(p + "self_attn.q_proj.weight", &wq, {E, E}));
    FRT2_TRY(need(p + "self_attn.q_proj.bias", &bq, {E}));
    FRT2_TRY(need(p + "self_attn.k_proj.weight", &wk, {E, E}));     // no bias (whisper.py:37)
    FRT2_TRY(need(p + "self_attn.v_proj.weight", &wv, {E, E}));
    FRT2_TRY(need(p + "self_attn.v_proj.bias", &bv, {E}));
    std::vector<float> wqkv(static_cast<size_t>(3 * E * E)), bqkv(static_cast<size_t>(3 * E), 0.f);
    std::copy(wq->data.begin(), wq->data.end(), wqkv.begin());
    std::copy(wk->data.begin(), wk->data.end(), wqkv.begin() + E * E);
    std::copy(wv->data.begin(), wv->data.end(), wqkv.begin() + 2 * E * E);
    std::copy(bq->data.begin(), bq->data.end(), bqkv.begin());
    std::copy(bv->data.begin(), bv->data.end(), bqkv.begin() + 2 * E);
    FRT2_TRY(up16(wqkv, &L.w_qkv));
    FRT2_TRY(up32(bqkv, &L.b_qkv));
    FRT2_TRY(need(p + "self_attn.out_proj.weight", &w, {E, E}));
    FRT2_TRY(need(p + "self_attn.out_proj.bias", &b, {E}));
    FRT2_TRY(up16(w->data, &L.w_o));
    FRT2_TRY(up32(b->data, &L.b_o));
    FRT2_TRY(need(p + "fc1.weight", &w, {F, E}));
    FRT2_TRY(need(p + "fc1.bias", &b, {F}));
    FRT2_TRY(up16(w->data, &L.w_fc1));
    FRT2_TRY(up32(b->data, &L.b_fc1));
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
  FRT2_TRY(need("ssl_adaptor.layer_norm.weight", &w, {E}));
  FRT2_TRY(need("ssl_adaptor.layer_norm.bias", &b, {E}));
  FRT2_TRY(up32(w->data, &lnf_g));
  FRT2_TRY(up32(b->data, &lnf_b));
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
  FRT2_TRY(dev_alloc(reinterpret_cast<void**>(&sched), 16));
  FRT2_CUDA_OK(cudaMemset(sched, 0, 16));
  FRT2_CUDA_OK(cudaDeviceSynchronize());
  raw.clear();
  finalized = true;
  return FRT2_OK;
}

int Encoder::features(const float* ssl, const float* aco_feats, int B, int T, float* vq_in, cudaStream_t st) {
  const int64_t M = static_cast<int64_t>(B) * T, M4 = M / pool, P = static_cast<int64_t>(pool) * D, E = Es;
  // ---- workspace carve-up (256-byte aligned) ----
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return o;
  };
  const size_t o_ssl16 = take(M * ssl_in * 2), o_x32 = take(M * E * 4), o_n16 = take(M * E * 2);
  const size_t o_qkv = take(M * 3 * E * 2), o_o16 = take(M * E * 2), o_g16 = take(M * F * 2);
  const size_t o_cat32 = take(M * D * 4), o_cat16 = take(M * D * 2);
  const size_t o_gu = take(M4 * 2 * P * 2), o_act = take(M4 * P * 2), o_c32 = take(M4 * P * 4), o_cn16 = take(M4 * P * 2);
  if (off > ws_bytes) {
    if (ws) {
      FRT2_CUDA_OK(cudaDeviceSynchronize());
      FRT2_CUDA_OK(cudaFree(ws));
      ws = nullptr;
      ws_bytes = 0;
    }
    FRT2_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ws), off));
    ws_bytes = off;
  }
  // calls are asynchronous and share the arena: a call on another CUDA stream waits for the previous one on the device
  if (ws_used && st != ws_last) FRT2_CUDA_OK(cudaStreamWaitEvent(st, ws_event, 0));
  __half* ssl16 = reinterpret_cast<__half*>(ws + o_ssl16);
  float* x32 = reinterpret_cast<float*>(ws + o_x32);
  __half* n16 = reinterpret_cast<__half*>(ws + o_n16);
  __half* qkv16 = reinterpret_cast<__half*>(ws + o_qkv);
  __half* o16 = reinterpret_cast<__half*>(ws + o_o16);
  __half* g16 = reinterpret_cast<__half*>(ws + o_g16);
  float* cat32 = reinterpret_cast<float*>(ws + o_cat32);
  __half* cat16 = reinterpret_cast<__half*>(ws + o_cat16);
  __half* gu16 = reinterpret_cast<__half*>(ws + o_gu);
  __half* act16 = reinterpret_cast<__half*>(ws + o_act);
  float* c32 = reinterpret_cast<float*>(ws + o_c32);
  __half* cn16 = reinterpret_cast<__half*>(ws + o_cn16);

  auto gemm = [&](const __half* A, int64_t rows, int K, const __half* W, int N, const float* bias, int act,
                  const float* resid, float* out32, int64_t ld32, __half* out16, int64_t ld16) {
    GemmDesc g{};
    g.A = A; g.a_row_pitch = K; g.a_batch_pitch = 0; g.rows_a = static_cast<int>(rows); g.batches = 1;
    g.Kc = K; g.ntaps = 1; g.row_shift = 0; g.W = W; g.N = N; g.rows_out = static_cast<int>(rows);
    g.alpha = 1.0f; g.bias = bias; g.act = act; g.resid = resid; g.out32 = out32; g.ld32 = ld32; g.out16 = out16;
    g.ld16 = ld16;
    ++launches;
    return gemm_tc(g, st);
  };
  auto cvt = [&](const float* src, int64_t ld_src, int64_t rows, int C, float* d32, int64_t ld32, __half* d16,
                 int64_t ld16) {
    const long long n = rows * (C / 4);
    cvt_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, ld_src, rows, C / 4, d32, ld32, d16, ld16);
    ++launches;
    return cudaGetLastError();
  };
  auto ln = [&](const float* x, int64_t rows, int C, const float* g, const float* b, __half* out) {
    ++launches;
    return layer_norm_rows_batched(x, C, rows, static_cast<int>(rows), C, g, b, 1e-5f, 0, out, C, 0, st);
  };

  // ---- SslAdaptor (model.py:53-66) ----
  FRT2_CUDA_OK(cvt(ssl, ssl_in, M, ssl_in, nullptr, 0, ssl16, ssl_in));
  FRT2_TRY(gemm(ssl16, M, ssl_in, w_in, Es, b_in, ACT_NONE, nullptr, x32, E, nullptr, 0));
  for (int i = 0; i < nl; ++i) {
    const EncLayer& L = layers[i];
    FRT2_TRY(ln(x32, M, Es, L.ln1_g, L.ln1_b, n16));
    FRT2_TRY(gemm(n16, M, Es, L.w_qkv, 3 * Es, L.b_qkv, ACT_NONE, nullptr, nullptr, 0, qkv16, 3 * E));
    AttnDesc a{};
    a.B = B; a.H = H; a.hd = hd; a.Tq = T; a.Tk = T; a.q_pos0 = 0; a.block_causal = 0;
    a.q = qkv16; a.q_row_pitch = 3 * E; a.q_batch_pitch = static_cast<int64_t>(T) * 3 * E;
    a.k = qkv16 + E; a.v = qkv16 + 2 * E; a.kv_row_pitch = 3 * E; a.kv_batch_pitch = a.q_batch_pitch;
    a.out = o16; a.o_row_pitch = E; a.o_batch_pitch = static_cast<int64_t>(T) * E;
    a.scale = 1.0f / std::sqrt(static_cast<float>(hd));
    a.sched = sched;
    ++launches;
    if ((hd == 64 || hd == 128) && T >= 32) FRT2_TRY(attention_tc(a, st));
    else FRT2_TRY(attention_warp(a, st));
    FRT2_TRY(gemm(o16, M, Es, L.w_o, Es, L.b_o, ACT_NONE, x32, x32, E, nullptr, 0));
    FRT2_TRY(ln(x32, M, Es, L.ln2_g, L.ln2_b, n16));
    FRT2_TRY(gemm(n16, M, Es, L.w_fc1, F, L.b_fc1, ACT_GELU, nullptr, nullptr, 0, g16, F));
    FRT2_TRY(gemm(g16, M, F, L.w_fc2, Es, L.b_fc2, ACT_NONE, x32, x32, E, nullptr, 0));
  }
  FRT2_TRY(ln(x32, M, Es, lnf_g, lnf_b, n16));
  // out_proj writes the semantic half of the concatenated features (fp32 for the residual of model.py:117, fp16 as the
  // operand of the gate / up convolutions); the acoustic half is copied in beside it: torch.cat without a pass of its own
  FRT2_TRY(gemm(n16, M, Es, w_out, ssl_out, b_out, ACT_NONE, nullptr, cat32, D, cat16, D));
  FRT2_CUDA_OK(cvt(aco_feats, aco, M, aco, cat32 + ssl_out, D, cat16 + ssl_out, D));
  // ---- ResidualDownConv (model.py:106-121) on the (M/pool, pool*D) view ----
  FRT2_TRY(gemm(cat16, M4, static_cast<int>(P), w_gu, static_cast<int>(2 * P), nullptr, ACT_NONE, nullptr, nullptr, 0,
                gu16, 2 * P));
  {
    const long long n = M4 * (P / 8);
    silu_mul_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(gu16, M4, static_cast<int>(P / 8), act16);
    ++launches;
    FRT2_CUDA_OK(cudaGetLastError());
  }
  FRT2_TRY(gemm(act16, M4, static_cast<int>(P), w_down, static_cast<int>(P), nullptr, ACT_NONE, cat32, c32, P, nullptr, 0));
  FRT2_TRY(ln(c32, M4, static_cast<int>(P), dln_g, dln_b, cn16));
  FRT2_TRY(gemm(cn16, M4, static_cast<int>(P), w_dout, D, b_dout, ACT_NONE, nullptr, vq_in, D, nullptr, 0));
  if (ws_event == nullptr) FRT2_CUDA_OK(cudaEventCreateWithFlags(&ws_event, cudaEventDisableTiming));
  FRT2_CUDA_OK(cudaEventRecord(ws_event, st));
  ws_last = st;
  ws_used = true;
  return FRT2_OK;
}

}  // namespace frt2

using namespace frt2;

struct frt2_encoder { Encoder e; };

extern "C" {

int frt2_enc_create(const frt2_enc_config* cfg, int device, frt2_encoder** out) {
  FRT2_REQUIRE(cfg != nullptr && out != nullptr, FRT2_ERR_BAD_ARG, "frt2_enc_create: null argument");
  FRT2_REQUIRE(cfg->ssl_in_dim > 0 && cfg->ssl_in_dim % 64 == 0 && cfg->ssl_embed_dim > 0 && cfg->ssl_embed_dim % 64 == 0,
               FRT2_ERR_BAD_ARG, "frt2_enc_create: ssl_adaptor in_dim / embed_dim must be positive multiples of 64");
  FRT2_REQUIRE(cfg->ssl_num_heads > 0 && cfg->ssl_embed_dim % cfg->ssl_num_heads == 0, FRT2_ERR_BAD_ARG,
               "frt2_enc_create: embed_dim must be divisible by num_heads");
  const int hd = cfg->ssl_embed_dim / cfg->ssl_num_heads;
  FRT2_REQUIRE(hd == 32 || hd == 64 || hd == 128, FRT2_ERR_BAD_ARG, "frt2_enc_create: head_dim must be 32, 64 or 128");
  FRT2_REQUIRE(cfg->ssl_out_dim > 0 && cfg->ssl_out_dim % 8 == 0 && cfg->aco_dim > 0 && cfg->aco_dim % 8 == 0,
               FRT2_ERR_BAD_ARG, "frt2_enc_create: ssl_adaptor.out_dim and acoustic_encoder.embed_dim must be multiples of 8");
  FRT2_REQUIRE(cfg->avg_pooler >= 1 && cfg->avg_pooler <= 8, FRT2_ERR_BAD_ARG, "frt2_enc_create: avg_pooler must be in [1, 8]");
  const int D = cfg->ssl_out_dim + cfg->aco_dim;
  FRT2_REQUIRE((static_cast<int64_t>(D) * cfg->avg_pooler) % 64 == 0, FRT2_ERR_BAD_ARG,
               "frt2_enc_create: avg_pooler * (out_dim + aco_dim) must be a multiple of 64");
  const int F = cfg->ssl_ffn_dim > 0 ? cfg->ssl_ffn_dim : 4 * cfg->ssl_embed_dim;   // whisper.py:137
  FRT2_REQUIRE(F % 64 == 0, FRT2_ERR_BAD_ARG, "frt2_enc_create: ffn_dim must be a multiple of 64");
  int ndev = 0;
  FRT2_CUDA_OK(cudaGetDeviceCount(&ndev));
  FRT2_REQUIRE(device >= 0 && device < ndev, FRT2_ERR_BAD_ARG, "frt2_enc_create: bad device index");
  cudaDeviceProp prop;
  FRT2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  FRT2_REQUIRE(prop.major == 10, FRT2_ERR_BAD_ARG, "frt2_enc_create: this library is sm_100a only (no fallback path)");
  auto* fe = new frt2_encoder();
  Encoder& e = fe->e;
  e.device = device;
  e.ssl_in = cfg->ssl_in_dim; e.Es = cfg->ssl_embed_dim; e.ssl_out = cfg->ssl_out_dim; e.nl = cfg->ssl_num_layers;
  e.H = cfg->ssl_num_heads; e.F = F; e.aco = cfg->aco_dim; e.pool = cfg->avg_pooler; e.D = D; e.hd = hd;
  *out = fe;
  return FRT2_OK;
}

int frt2_enc_load_tensor(frt2_encoder* fe, const char* key, const float* data, int ndim, const int64_t* shape,
                         int on_device) {
  FRT2_REQUIRE(fe && key && data && shape && ndim >= 0 && ndim <= 4, FRT2_ERR_BAD_ARG, "frt2_enc_load_tensor: bad argument");
  Encoder& e = fe->e;
  FRT2_REQUIRE(!e.finalized, FRT2_ERR_BAD_ARG, "frt2_enc_load_tensor: already finalized");
  const std::string k(key);
  if (k.rfind("ssl_adaptor.", 0) != 0 && k.rfind("downsample.", 0) != 0) return FRT2_OK;   // not part of this stage
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

}  // extern "C"
