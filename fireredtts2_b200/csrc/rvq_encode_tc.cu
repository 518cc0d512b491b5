// K7t — residual vector quantisation, ENCODE side, on the tensor cores: ResidualVQ.encode_codes (reference
// rvq.py:128-143) over VectorQuantize.encode_code (rvq.py:62-89) with every contraction of the chain — input_proj,
// in_project_i, the distance products z_e . C_i^T, out_project_i — as a tcgen05 GEMM (gemm_tc) on SPLIT fp16 operands:
//
//     x = x_hi + x_lo,  x_hi = fp16(x),  x_lo = x - x_hi            (|x_lo| <= 2^-12 |x|)
//     a . w  ~=  a_hi.w_hi + a_hi.w_lo + a_lo.w_hi                  (the dropped a_lo.w_lo is 2^-24 relative)
//
// evaluated by ONE GEMM over a three times longer reduction: A row = [a_hi | a_hi/S | a_lo*S], W row = [w_hi | w_lo*S |
// w_hi/S] with S = 2^8 (keeps the small parts in fp16's normal range), fp32 accumulation in tensor memory.  The result
// carries ~22 significant bits like the reference's fp32 products, so an index differs from the reference's only at a
// near-tie — the same contract as the CUDA-core kernel (rvq_encode.cu), which stays as the checker and serves widths
// that are not multiples of 64.  Distances, arg-max (first maximum), z_q and the residual update keep the reference's
// own expression and rounding order:  -((|z_e|^2 - 2 s) + |c|^2),  z_q = z_e + (C[idx] - z_e),  residual -= (W z_q + b).
//
// Per quantizer: GEMM in_project -> split_rows (+ |z_e|^2) -> GEMM distances -> argmax_update (index, z_q, split z_q)
// -> GEMM out_project with the residual update in its epilogue (alpha = -1, bias = -b, resid = residual) -> split_rows.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace frt2 {

namespace {

constexpr float SPLIT_S = 256.0f;   // 2^8

// residual (R, C) fp32 <- z (B, C, T) with element strides (channel-major producers: sT == 1, time-major: sD == 1)
__global__ void __launch_bounds__(256) gather_z_kernel(const float* __restrict__ z, long long sB, long long sD, long long sT,
                                                       int T, int C, long long R, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const long long r0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (sT == 1 || sT < sD) {
    // tokens are the fast axis of the source: read with tx along tokens, transpose through shared memory
    for (int j = ty; j < 32; j += 8) {
      const long long r = r0 + tx;
      const int c = c0 + j;
      float v = 0.f;
      if (r < R && c < C) v = z[(r / T) * sB + static_cast<long long>(c) * sD + (r % T) * sT];
      tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const long long r = r0 + j;
      const int c = c0 + tx;
      if (r < R && c < C) out[r * C + c] = tile[tx][j];
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      const long long r = r0 + j;
      const int c = c0 + tx;
      if (r < R && c < C) out[r * C + c] = z[(r / T) * sB + static_cast<long long>(c) * sD + (r % T) * sT];
    }
  }
}

__device__ __forceinline__ uint2 pack4(const __half (&h)[4]) {
  uint2 u;
  u.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
  u.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
  return u;
}

// x (rows, C) fp32 -> out (rows, 3C) fp16 = [hi | hi/S | lo*S]; a2[row] = sum x^2 (optional).  One warp per row.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, long long ldx, long long rows, int C,
                                                         __half* __restrict__ out, float* __restrict__ a2) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * ldx;
  __half* o = out + row * 3 * C;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a[4] = {v.x, v.y, v.z, v.w};
    __half hi[4], hs[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = to_half_sat(a[j]);
      const float h = __half2float(hi[j]);
      hs[j] = __float2half_rn(h * (1.0f / SPLIT_S));
      lo[j] = __float2half_rn((a[j] - h) * SPLIT_S);
      s = fmaf(a[j], a[j], s);
    }
    *reinterpret_cast<uint2*>(o + c) = pack4(hi);
    *reinterpret_cast<uint2*>(o + C + c) = pack4(hs);
    *reinterpret_cast<uint2*>(o + 2 * C + c) = pack4(lo);
  }
  if (a2 != nullptr) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) a2[row] = s;
  }
}

// One warp per token: idx = first maximum of -((a2 - 2 s_k) + c2_k) over the K codes (rvq.py:71-78), the index goes to
// codes[(i, b, t)], z_q = z_e + (C[idx] - z_e) (rvq.py:82-85) to zq32 (when the out-projection is the Identity: the
// residual update residual -= z_q happens here) and, split, to zq_s for the out-projection GEMM.
__global__ void __launch_bounds__(256) argmax_update_kernel(const float* __restrict__ scores, long long R, int K,
                                                            const float* __restrict__ a2, const float* __restrict__ c2,
                                                            const float* __restrict__ ze, int cd,
                                                            const float* __restrict__ Cb /*(K, cd)*/,
                                                            long long* __restrict__ codes_i, __half* __restrict__ zq_s,
                                                            float* __restrict__ resid_identity, int rd) {
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float a = a2[r];
  const float* s = scores + r * K;
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 sv = *reinterpret_cast<const float4*>(s + k);
    const float4 cv = __ldg(reinterpret_cast<const float4*>(c2 + k));
    const float sj[4] = {sv.x, sv.y, sv.z, sv.w}, cj[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = -(__fadd_rn(__fsub_rn(a, 2.0f * sj[j]), cj[j]));
      if (v > bv) {            // increasing k within a lane: strict > keeps the first maximum
        bv = v;
        bi = k + j;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, d);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if (bi == 0x7fffffff) bi = 0;     // a row of NaN scores (non-finite input): keep the index in range
  if (lane == 0) codes_i[r] = bi;
  const float* zr = ze + r * cd;
  const float* q = Cb + static_cast<long long>(bi) * cd;
  for (int c = lane; c < cd; c += 32) {
    const float zv = zr[c];
    const float zq = __fadd_rn(zv, __fsub_rn(__ldg(q + c), zv));
    if (zq_s != nullptr) {
      const __half hi = to_half_sat(zq);
      const float h = __half2float(hi);
      zq_s[r * 3 * cd + c] = hi;
      zq_s[r * 3 * cd + cd + c] = __float2half_rn(h * (1.0f / SPLIT_S));
      zq_s[r * 3 * cd + 2 * cd + c] = __float2half_rn((zq - h) * SPLIT_S);
    }
    if (resid_identity != nullptr) resid_identity[r * rd + c] = __fsub_rn(resid_identity[r * rd + c], zq);
  }
}

}  // namespace

// host helper (engine.cu, at load): W (N, Kk) fp32 row-major -> (N, 3 Kk) fp16 rows [w_hi | w_lo*S | w_hi/S]
void rvq_split_weight_host(const float* W, int64_t N, int64_t Kk, __half* out) {
#pragma omp parallel for schedule(static)
  for (long long n = 0; n < N; ++n) {
    for (int64_t k = 0; k < Kk; ++k) {
      const float w = W[n * Kk + k];
      const __half hi = __float2half_rn(std::min(65504.0f, std::max(-65504.0f, w)));
      const float h = __half2float(hi);
      out[n * 3 * Kk + k] = hi;
      out[n * 3 * Kk + Kk + k] = __float2half_rn((w - h) * SPLIT_S);
      out[n * 3 * Kk + 2 * Kk + k] = __float2half_rn(h * (1.0f / SPLIT_S));
    }
  }
}

bool rvq_encode_tc_applicable(const RvqEncDesc& d) {
  return d.rd % 64 == 0 && d.cd % 64 == 0 && d.input_dim % 64 == 0 && d.K % 4 == 0 && d.s_C != nullptr &&
         (d.WinpT == nullptr || d.s_inp != nullptr) && (d.WinT == nullptr || d.s_in != nullptr) &&
         (d.WoutT == nullptr || d.s_out != nullptr);
}

size_t rvq_encode_tc_ws_bytes(const RvqEncDesc& d, long long R) {
  auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  const int maxc = std::max(std::max(d.input_dim, d.rd), d.cd);
  return al(R * d.input_dim * 4) + al(R * d.rd * 4) + al(R * 3 * maxc * 2) + al(R * d.cd * 4) + al(R * 3 * d.cd * 2) +
         al(R * 4) + al(R * static_cast<size_t>(d.K) * 4) + al(R * 3 * d.cd * 2);
}

// ws: rvq_encode_tc_ws_bytes(d, R) bytes.  *launches receives the number of kernels launched.
int rvq_encode_tc(const RvqEncDesc& d, uint8_t* ws, cudaStream_t st, long long* launches) {
  const long long R = static_cast<long long>(d.B) * d.T;
  if (R == 0 || d.nq == 0) return FRT2_OK;
  auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  const int maxc = std::max(std::max(d.input_dim, d.rd), d.cd);
  size_t off = 0;
  auto take = [&](size_t b) { uint8_t* p = ws + off; off += al(b); return p; };
  float* zin = reinterpret_cast<float*>(take(R * d.input_dim * 4));
  float* res = reinterpret_cast<float*>(take(R * d.rd * 4));
  __half* a_s = reinterpret_cast<__half*>(take(R * 3 * maxc * 2));     // split operand of input_proj / in_project
  float* ze = reinterpret_cast<float*>(take(R * d.cd * 4));
  __half* ze_s = reinterpret_cast<__half*>(take(R * 3 * d.cd * 2));
  float* a2 = reinterpret_cast<float*>(take(R * 4));
  float* scores = reinterpret_cast<float*>(take(R * static_cast<size_t>(d.K) * 4));
  __half* zq_s = reinterpret_cast<__half*>(take(R * 3 * d.cd * 2));
  long long n = 0;
  auto gemm = [&](const __half* A, int K3, const __half* W, int N, float alpha, const float* bias, const float* resid,
                  float* out32) {
    GemmDesc g{};
    g.A = A; g.a_row_pitch = K3; g.a_batch_pitch = 0; g.rows_a = static_cast<int>(R); g.batches = 1; g.Kc = K3; g.ntaps = 1;
    g.W = W; g.N = N; g.rows_out = static_cast<int>(R); g.alpha = alpha; g.bias = bias; g.act = ACT_NONE; g.resid = resid;
    g.out32 = out32; g.ld32 = N;
    ++n;
    return gemm_tc(g, st);
  };
  auto split = [&](const float* x, int C, __half* out, float* sq) {
    split_rows_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, st>>>(x, C, R, C, out, sq);
    ++n;
    return cudaGetLastError();
  };
  // ---- residual = input_proj(z)   (rvq.py:129-130)
  {
    float* dst = d.WinpT == nullptr ? res : zin;
    gather_z_kernel<<<dim3(static_cast<unsigned>((R + 31) / 32), (d.input_dim + 31) / 32), 256, 0, st>>>(
        d.z, d.sB, d.sD, d.sT, d.T, d.input_dim, R, dst);
    ++n;
    FRT2_CUDA_OK(cudaGetLastError());
    if (d.WinpT != nullptr) {
      FRT2_CUDA_OK(split(zin, d.input_dim, a_s, nullptr));
      FRT2_TRY(gemm(a_s, 3 * d.input_dim, d.s_inp, d.rd, 1.0f, d.binp, nullptr, res));
    }
  }
  for (int i = 0; i < d.nq; ++i) {
    // ---- z_e = in_project_i(residual)   (rvq.py:65)
    const float* ze_i = res;     // Identity: z_e is the residual itself (rd == cd)
    if (d.WinT != nullptr) {
      FRT2_CUDA_OK(split(res, d.rd, a_s, nullptr));
      FRT2_TRY(gemm(a_s, 3 * d.rd, d.s_in + static_cast<long long>(i) * d.cd * 3 * d.rd, d.cd, 1.0f, d.bin + i * d.cd,
                    nullptr, ze));
      ze_i = ze;
    }
    FRT2_CUDA_OK(split(ze_i, d.cd, ze_s, a2));
    // ---- s = z_e . C_i^T, distances + arg-max, z_q   (rvq.py:71-85)
    FRT2_TRY(gemm(ze_s, 3 * d.cd, d.s_C + static_cast<long long>(i) * d.K * 3 * d.cd, d.K, 1.0f, nullptr, nullptr, scores));
    argmax_update_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, st>>>(
        scores, R, d.K, a2, d.c2 + static_cast<long long>(i) * d.K, ze_i, d.cd,
        d.C + static_cast<long long>(i) * d.K * d.cd, d.codes + static_cast<long long>(i) * R,
        d.WoutT != nullptr ? zq_s : nullptr, d.WoutT != nullptr ? nullptr : res, d.rd);
    ++n;
    FRT2_CUDA_OK(cudaGetLastError());
    // ---- residual -= out_project_i(z_q)   (rvq.py:86,138): alpha = -1, bias = -b, + residual in the GEMM epilogue
    if (d.WoutT != nullptr && i + 1 < d.nq)
      FRT2_TRY(gemm(zq_s, 3 * d.cd, d.s_out + static_cast<long long>(i) * d.rd * 3 * d.cd, d.rd, -1.0f,
                    d.nbout + i * d.rd, res, res));
  }
  if (launches != nullptr) *launches = n;
  return FRT2_OK;
}

}  // namespace frt2
