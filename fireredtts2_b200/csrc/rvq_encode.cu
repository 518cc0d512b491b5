// K7 — residual vector quantisation, ENCODE side: ResidualVQ.encode_codes (reference rvq.py:128-143) over
// VectorQuantize.encode_code (rvq.py:62-89).  The producer of the token tensors the decode path consumes.
//
//   residual = input_proj(z)
//   for i in quantizers:  z_e = in_project_i(residual)
//                         dist = |z_e|^2 - (2 z_e) . C_i^T + |C_i|^2 ;  idx_i = argmax(-dist)   (first maximum)
//                         z_q  = z_e + (C_i[idx_i] - z_e) ;  residual -= out_project_i(z_q)
//
// The 16 quantizers of a token form a dependent chain, but tokens are independent: one CTA owns a tile of 32 tokens,
// keeps their residual (32 x rvq_dim fp32) and the current z_e (32 x codebook_dim) in shared memory for the whole chain
// and streams the (L2-resident, 48 MB at C0) projection matrices and transposed codebooks past them: one launch, no
// intermediate tensor in HBM, indices written once.  Arithmetic is fp32 FMA on the CUDA cores ON PURPOSE: the result
// is an integer index, and the distance expression is evaluated in the reference's own form and rounding order
// ((a - 2s) + c2, negated, lowest index on ties) so that an index only differs from the reference's where two codes
// are closer than the fp32 rounding of the distance itself.  Thread = one output column (projection) or one code
// (distance), 32 token accumulators in registers, weights read coalesced ([k][column] layouts built at load), the
// tokens' activations broadcast from shared memory as float4.
#include <math_constants.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace frt2 {

namespace {

constexpr int RE_TM_MAX = 32;    // tokens per CTA: 32, or 16 / 8 for small batches (more CTAs; a token's arithmetic does not depend on it)
constexpr int RE_THREADS = 256;
constexpr int RE_WARPS = RE_THREADS / 32;
constexpr int RE_MAXCD = 256;    // codebook_dim limit (one thread per dim); also the k-chunk of the input projection

// acc[m] += sum_k act[m][k] * W[k][col]   for k in [0, kn), act rows have pitch `ap` floats (16-byte aligned, kn % 4 == 0)
template <int RE_TM>
__device__ __forceinline__ void tile_fma(float (&acc)[RE_TM], const float* __restrict__ act, int ap,
                                         const float* __restrict__ Wcol, int wld, int kn) {
  for (int k = 0; k < kn; k += 4) {
    const float w0 = __ldg(Wcol + static_cast<long long>(k) * wld);
    const float w1 = __ldg(Wcol + static_cast<long long>(k + 1) * wld);
    const float w2 = __ldg(Wcol + static_cast<long long>(k + 2) * wld);
    const float w3 = __ldg(Wcol + static_cast<long long>(k + 3) * wld);
#pragma unroll
    for (int m = 0; m < RE_TM; ++m) {
      const float4 a = *reinterpret_cast<const float4*>(act + m * ap + k);   // same address in every lane: broadcast
      acc[m] = fmaf(a.x, w0, acc[m]);
      acc[m] = fmaf(a.y, w1, acc[m]);
      acc[m] = fmaf(a.z, w2, acc[m]);
      acc[m] = fmaf(a.w, w3, acc[m]);
    }
  }
}

template <int RE_TM>
__global__ void __launch_bounds__(RE_THREADS) rvq_encode_kernel(RvqEncDesc d) {
  extern __shared__ __align__(16) uint8_t re_smem[];
  const int rp = d.rd + 4;                       // residual row pitch (floats): +4 keeps rows 16-byte aligned, de-phased
  const int zp = RE_MAXCD + 4;
  float* res = reinterpret_cast<float*>(re_smem);                  // [RE_TM][rp]
  float* ze = res + RE_TM * rp;                                     // [RE_TM][zp]   z_e, then z_q; input chunk staging
  float* a2 = ze + RE_TM * zp;                                      // [RE_TM]  |z_e|^2
  float* best_v = a2 + RE_TM;                                       // [RE_TM]
  int* best_i = reinterpret_cast<int*>(best_v + RE_TM);             // [RE_TM]
  float* red_v = reinterpret_cast<float*>(best_i + RE_TM);          // [RE_WARPS][RE_TM]
  int* red_i = reinterpret_cast<int*>(red_v + RE_WARPS * RE_TM);    // [RE_WARPS][RE_TM]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long R = static_cast<long long>(d.B) * d.T;
  const long long r0 = static_cast<long long>(blockIdx.x) * RE_TM;
  const int valid = static_cast<int>(min(static_cast<long long>(RE_TM), R - r0));

  // ---------------------------------------------------------------- residual = input_proj(z)   (rvq.py:129-130)
  if (d.WinpT == nullptr) {
    for (int e = tid; e < RE_TM * d.rd; e += RE_THREADS) {
      // channel-major producers (sT == 1): consecutive threads take consecutive tokens; time-major: consecutive dims
      const int m = (d.sT == 1) ? e % RE_TM : e / d.rd;
      const int k = (d.sT == 1) ? e / RE_TM : e % d.rd;
      float v = 0.f;
      if (m < valid) {
        const long long r = r0 + m;
        v = d.z[(r / d.T) * d.sB + static_cast<long long>(k) * d.sD + (r % d.T) * d.sT];
      }
      res[m * rp + k] = v;
    }
  } else {
    float acc0[RE_TM], acc1[RE_TM];   // columns tid and tid + 256 of the rvq_dim outputs
#pragma unroll
    for (int m = 0; m < RE_TM; ++m) acc0[m] = acc1[m] = 0.f;
    for (int k0 = 0; k0 < d.input_dim; k0 += RE_MAXCD) {
      const int kn = min(RE_MAXCD, d.input_dim - k0);
      __syncthreads();
      for (int e = tid; e < RE_TM * kn; e += RE_THREADS) {
        const int m = (d.sT == 1) ? e % RE_TM : e / kn;
        const int k = (d.sT == 1) ? e / RE_TM : e % kn;
        float v = 0.f;
        if (m < valid) {
          const long long r = r0 + m;
          v = d.z[(r / d.T) * d.sB + static_cast<long long>(k0 + k) * d.sD + (r % d.T) * d.sT];
        }
        ze[m * zp + k] = v;
      }
      __syncthreads();
      if (tid < d.rd) tile_fma(acc0, ze, zp, d.WinpT + static_cast<long long>(k0) * d.rd + tid, d.rd, kn);
      if (tid + RE_THREADS < d.rd)
        tile_fma(acc1, ze, zp, d.WinpT + static_cast<long long>(k0) * d.rd + tid + RE_THREADS, d.rd, kn);
    }
    if (tid < d.rd) {
      const float b = __ldg(d.binp + tid);
#pragma unroll
      for (int m = 0; m < RE_TM; ++m) res[m * rp + tid] = acc0[m] + b;
    }
    if (tid + RE_THREADS < d.rd) {
      const float b = __ldg(d.binp + tid + RE_THREADS);
#pragma unroll
      for (int m = 0; m < RE_TM; ++m) res[m * rp + tid + RE_THREADS] = acc1[m] + b;
    }
  }
  __syncthreads();

  for (int i = 0; i < d.nq; ++i) {
    // ---------------------------------------------------------------- z_e = in_project_i(residual)   (rvq.py:65)
    if (d.WinT != nullptr) {
      if (tid < d.cd) {
        float acc[RE_TM];
#pragma unroll
        for (int m = 0; m < RE_TM; ++m) acc[m] = 0.f;
        tile_fma(acc, res, rp, d.WinT + static_cast<long long>(i) * d.rd * d.cd + tid, d.cd, d.rd);
        const float b = __ldg(d.bin + i * d.cd + tid);
#pragma unroll
        for (int m = 0; m < RE_TM; ++m) ze[m * zp + tid] = acc[m] + b;
      }
    } else {
      for (int e = tid; e < RE_TM * d.cd; e += RE_THREADS) ze[(e / d.cd) * zp + e % d.cd] = res[(e / d.cd) * rp + e % d.cd];
    }
    if (tid < RE_TM) {
      best_v[tid] = -CUDART_INF_F;
      best_i[tid] = 0;
    }
    __syncthreads();
    // |z_e|^2 per token (encodings.pow(2).sum(1), rvq.py:72): one warp per token
    for (int m = warp; m < RE_TM; m += RE_WARPS) {
      float s = 0.f;
      for (int k = lane; k < d.cd; k += 32) s = fmaf(ze[m * zp + k], ze[m * zp + k], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) a2[m] = s;
    }
    __syncthreads();
    // ---------------------------------------------------------------- distances + argmax   (rvq.py:71-78)
    const float* CT = d.CT + static_cast<long long>(i) * d.cd * d.K;
    for (int c0 = 0; c0 < d.K; c0 += RE_THREADS) {
      const int c = c0 + tid;
      const bool c_ok = c < d.K;
      float acc[RE_TM];
#pragma unroll
      for (int m = 0; m < RE_TM; ++m) acc[m] = 0.f;
      if (c_ok) tile_fma(acc, ze, zp, CT + c, d.K, d.cd);
      const float c2 = c_ok ? __ldg(d.c2 + static_cast<long long>(i) * d.K + c) : 0.f;
#pragma unroll
      for (int m = 0; m < RE_TM; ++m) {
        // -dist in the reference's rounding order: (|z_e|^2 - (2 z_e).c) + |c|^2
        float v = c_ok ? -(__fadd_rn(__fsub_rn(a2[m], 2.0f * acc[m]), c2)) : -CUDART_INF_F;
        int vi = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
          if (ov > v || (ov == v && oi < vi)) {
            v = ov;
            vi = oi;
          }
        }
        if (lane == 0) {
          red_v[warp * RE_TM + m] = v;
          red_i[warp * RE_TM + m] = vi;
        }
      }
      __syncthreads();
      if (tid < RE_TM) {
        float bv = best_v[tid];
        int bi = best_i[tid];
#pragma unroll
        for (int w = 0; w < RE_WARPS; ++w) {   // increasing code index: strict > keeps the FIRST maximum
          const float v = red_v[w * RE_TM + tid];
          if (v > bv) {
            bv = v;
            bi = red_i[w * RE_TM + tid];
          }
        }
        best_v[tid] = bv;
        best_i[tid] = bi;
      }
      __syncthreads();
    }
    if (tid < valid) {
      const long long r = r0 + tid;
      d.codes[(static_cast<long long>(i) * d.B + r / d.T) * d.T + r % d.T] = best_i[tid];
    }
    // ---------------------------------------------------------------- z_q = z_e + (C[idx] - z_e)   (rvq.py:82-85)
    const float* Cb = d.C + static_cast<long long>(i) * d.K * d.cd;
    for (int e = tid; e < RE_TM * d.cd; e += RE_THREADS) {
      const int m = e / d.cd, k = e - m * d.cd;
      const float z = ze[m * zp + k];
      const float q = __ldg(Cb + static_cast<long long>(best_i[m]) * d.cd + k);
      ze[m * zp + k] = __fadd_rn(z, __fsub_rn(q, z));
    }
    __syncthreads();
    // ---------------------------------------------------------------- residual -= out_project_i(z_q)   (rvq.py:86,138)
    if (d.WoutT != nullptr) {
      for (int c0 = 0; c0 < d.rd; c0 += RE_THREADS) {
        const int c = c0 + tid;
        if (c < d.rd) {
          float acc[RE_TM];
#pragma unroll
          for (int m = 0; m < RE_TM; ++m) acc[m] = 0.f;
          tile_fma(acc, ze, zp, d.WoutT + static_cast<long long>(i) * d.cd * d.rd + c, d.rd, d.cd);
          const float b = __ldg(d.bout + i * d.rd + c);
#pragma unroll
          for (int m = 0; m < RE_TM; ++m) res[m * rp + c] = __fsub_rn(res[m * rp + c], acc[m] + b);
        }
      }
    } else {
      for (int e = tid; e < RE_TM * d.cd; e += RE_THREADS) {
        const int m = e / d.cd, k = e - m * d.cd;
        res[m * rp + k] = __fsub_rn(res[m * rp + k], ze[m * zp + k]);
      }
    }
    __syncthreads();
  }
}

size_t re_smem_bytes(int rd, int RE_TM) {
  return static_cast<size_t>(RE_TM) * (rd + 4) * 4 + static_cast<size_t>(RE_TM) * (RE_MAXCD + 4) * 4 + RE_TM * 4 * 3 +
         static_cast<size_t>(RE_WARPS) * RE_TM * 8;
}

}  // namespace

int rvq_encode(const RvqEncDesc& d, cudaStream_t stream) {
  FRT2_REQUIRE(d.rd >= 4 && d.rd <= 2 * RE_THREADS && d.rd % 4 == 0, FRT2_ERR_BAD_ARG,
               "rvq_encode: rvq_dim must be a multiple of 4, <= 512");
  FRT2_REQUIRE(d.cd >= 4 && d.cd <= RE_MAXCD && d.cd % 4 == 0, FRT2_ERR_BAD_ARG,
               "rvq_encode: codebook_dim must be a multiple of 4, <= 256");
  FRT2_REQUIRE(d.WinpT == nullptr || d.input_dim % 4 == 0, FRT2_ERR_BAD_ARG, "rvq_encode: input_dim must be a multiple of 4");
  FRT2_REQUIRE(d.WinT != nullptr || d.rd == d.cd, FRT2_ERR_BAD_ARG, "rvq_encode: Identity in_project needs rvq_dim == codebook_dim");
  const long long R = static_cast<long long>(d.B) * d.T;
  if (R == 0 || d.nq == 0) return FRT2_OK;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(rvq_encode_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(re_smem_bytes(2 * RE_THREADS, 32)));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rvq_encode_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(re_smem_bytes(2 * RE_THREADS, 16)));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rvq_encode_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(re_smem_bytes(2 * RE_THREADS, 8)));
  });
  FRT2_CUDA_OK(attr_err);
  // tokens per CTA: 32 once every SM gets a tile, else 8 (more, smaller tiles).  Measured at C0: 7200 tokens 9.5 ms with
  // 32 per CTA, 13.6 ms with 16; 2400 tokens 9.8 / 8.1 / 7.3 ms with 32 / 16 / 8 (smaller tiles re-stream the L2-resident
  // tables; a software prefetch of the next weights cost registers and time: 30.3 -> 33.3 ms at 24 000 tokens)
  static const int force_tm = getenv("FRT2_RVQ_ENC_TM") ? atoi(getenv("FRT2_RVQ_ENC_TM")) : 0;   // A/B
  const long long sms = num_sms();
  int tm = R >= 32 * sms ? 32 : 8;
  if (force_tm == 8 || force_tm == 16 || force_tm == 32) tm = force_tm;
  const unsigned grid = static_cast<unsigned>((R + tm - 1) / tm);
  const size_t smem = re_smem_bytes(d.rd, tm);
  if (tm == 32) rvq_encode_kernel<32><<<grid, RE_THREADS, smem, stream>>>(d);
  else if (tm == 16) rvq_encode_kernel<16><<<grid, RE_THREADS, smem, stream>>>(d);
  else rvq_encode_kernel<8><<<grid, RE_THREADS, smem, stream>>>(d);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
