// Persistent per-token step kernel ("megakernel") for the batch-1 / small-pool streaming step.
//
// One decode_one_token step (reference model.py:326-376) at <= 16 rows is ~90 dependent kernels of a few microseconds:
// replayed as a CUDA graph it costs ~9 us per node (launch gap + CTA ramp + first-load latency) against a 66 us HBM floor
// for the 429 MB of weights.  Here the SAME op list — recorded once per (chunk length, nq) by running the host pipeline
// in recording mode — is interpreted by ONE cooperative kernel, one 256-thread CTA per SM (so that the op bodies keep the
// 255-register budget they have as stand-alone kernels: a 512-thread CTA capped them at 128 and spilled the skinny GEMM's
// weight vectors), with a grid barrier (one atomic + an acquire spin per CTA) between dependent ops instead of a kernel
// boundary.  The op bodies are the very
// device functions the stand-alone kernels run (gemm_skinny_body, attention_warp_body, layer_norm_body,
// overlap_add_sample), so a step is bit-identical to the kernel-by-kernel path.
//
// Memory model: producers store with plain stores, arrive with  __syncthreads -> __threadfence -> atomicAdd ; consumers
// spin with ld.acquire.gpu, then __syncthreads — everything an op wrote is visible to every thread of the next op.  No
// op body reads kernel-written data through the non-coherent path (ld.global.nc); weights / biases / tables may.
#include <cooperative_groups.h>

#include "attention_warp_body.cuh"
#include "common.cuh"
#include "gemm_skinny_body.cuh"
#include "misc_bodies.cuh"

namespace frt2 {

namespace {

constexpr int MG_THREADS = 256;                       // == the skinny GEMM's tile group; 8 warps for attention / LayerNorm
constexpr int MG_OP_WORDS = (sizeof(MegaOp) + 15) / 16;

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (co-resident) grid have arrived `target` times in total
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int spins = 0;
    while (static_cast<int>(ld_acquire(bar) - target) < 0) {
      if (++spins > (1u << 22)) __trap();   // a lost CTA must not hang the box
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void op_rvq(const MegaRvq& r, int* s_idx /* [4][64] */) {
  // zero fill (the padding columns of the spectrum buffer): grid-stride, independent of the gather
  if (r.zero_ptr != nullptr) {
    const long long total = static_cast<long long>(r.zero_rows) * r.zero_cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      r.zero_ptr[(i / r.zero_cols) * r.zero_pitch + i % r.zero_cols] = __float2half_rn(0.f);
  }
  // K1 (kernels_misc.cu rvq_gather_sum_kernel): 4 tokens per CTA, index-ordered fp32 sum
  const long long R = static_cast<long long>(r.B) * r.L;
  const int D4 = r.D >> 2;
  const float4* tab4 = reinterpret_cast<const float4*>(r.tables);
  for (long long r0 = static_cast<long long>(blockIdx.x) * 4; r0 < R; r0 += static_cast<long long>(gridDim.x) * 4) {
    for (int e = threadIdx.x; e < 4 * r.nq; e += blockDim.x) {
      const int t = e / r.nq, i = e - t * r.nq;
      const long long row = r0 + t;
      int idx = 0;
      if (row < R) {
        const long long b = row / r.L, l = row - b * r.L;
        const int raw = r.tokens[b * r.sB + i * r.sQ + l * r.sL];
        if (raw < 0 || raw >= r.K) atomicOr(r.err_word, DEV_ERR_INDEX_OOR);
        else idx = raw;
      }
      s_idx[t * 64 + i] = idx;
    }
    __syncthreads();
    for (int t = 0; t < 4; ++t) {
      const long long row = r0 + t;
      if (row >= R) break;
      for (int c = threadIdx.x; c < D4; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i0 = 0; i0 < r.nq; i0 += 8) {      // 8 rows in flight; the sum stays in index order
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = min(i0 + u, r.nq - 1);
            v[u] = __ldg(tab4 + (static_cast<long long>(i) * r.K + s_idx[t * 64 + i]) * D4 + c);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (i0 + u < r.nq) {
              acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
          }
        }
        uint2 h;
        h.x = pack_half2(acc.x, acc.y);
        h.y = pack_half2(acc.z, acc.w);
        reinterpret_cast<uint2*>(r.sum16)[row * D4 + c] = h;
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void op_roll(const MegaRoll& ro) {
  const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  // One flat index space over the tail copy (16-byte pieces) and every history row (16-byte pieces): each thread issues
  // all its loads before its first store — one round trip for the whole roll instead of one per buffer.
  // new iSTFT tail = last 3 windowed frames (kernels_misc.cu update_tail_kernel); conv history: the last `hist` rows of
  // [hist | chunk] move to the head (engine.cu shift_history_kernel; rows >= hist, so source and destination are disjoint)
  const long long tail_v = 3LL * ro.n_fft / 4 * ro.B;             // float4 pieces
  long long start[17];
  start[0] = tail_v;
  const int E8 = ro.E / 8;
  for (int ei = 0; ei < ro.tb.n; ++ei) start[ei + 1] = start[ei] + static_cast<long long>(ro.B) * ro.tb.e[ei].hist * E8;
  const long long total = start[ro.tb.n];
  for (long long i0 = gtid; i0 < total; i0 += 4 * gthreads) {
    uint4 val[4];
    uint4* dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * gthreads;
      dst[u] = nullptr;
      if (i >= total) continue;
      if (i < tail_v) {
        const long long per = 3LL * ro.n_fft / 4;
        const int b = static_cast<int>(i / per);
        const long long k = i - b * per;
        if (!ro.all_items && ro.ctrl[b * CTRL_INTS + CTRL_ACTIVE] == 0) continue;
        val[u] = *reinterpret_cast<const uint4*>(ro.frames + b * ro.frames_batch_pitch +
                                                 static_cast<long long>(ro.T - 3) * ro.n_fft + 4 * k);
        dst[u] = reinterpret_cast<uint4*>(ro.tail + b * per * 4 + 4 * k);
      } else {
        int ei = 0;
        while (i >= start[ei + 1]) ++ei;
        const ShiftEntry& en = ro.tb.e[ei];
        const long long j = i - start[ei];
        const int c = static_cast<int>(j % E8);
        const int r = static_cast<int>((j / E8) % en.hist);
        const int b = static_cast<int>(j / (static_cast<long long>(E8) * en.hist));
        if (!ro.all_items && ro.ctrl[b * CTRL_INTS + CTRL_ACTIVE] == 0) continue;
        __half* base = en.p + b * en.batch_pitch;
        val[u] = *reinterpret_cast<const uint4*>(base + static_cast<long long>(r + en.rows) * ro.E + c * 8);
        dst[u] = reinterpret_cast<uint4*>(base + static_cast<long long>(r) * ro.E + c * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dst[u] != nullptr) *dst[u] = val[u];
  }
  // position advance (engine.cu advance_ctrl_kernel, pool form: active items only).  Nothing in this op reads CTRL_POS.
  if (gtid < ro.B && (ro.all_items || ro.ctrl[gtid * CTRL_INTS + CTRL_ACTIVE] != 0))
    ro.ctrl[gtid * CTRL_INTS + CTRL_POS] += ro.advance_frames;
}

__global__ void __launch_bounds__(MG_THREADS, 1)
stream_step_kernel(const MegaOp* __restrict__ ops, int nops, unsigned int* bar, unsigned int epoch0, int group_smem,
                   long long* trace /* debug: per op (start, end) %globaltimer of CTA 0, or null */) {
  extern __shared__ __align__(16) uint8_t mg_smem[];
  __shared__ __align__(16) uint4 s_op[2][MG_OP_WORDS];
  __shared__ int s_idx[4 * 64];
  auto fetch = [&](int i) {   // op descriptors are written by the host before the launch: read-only here
    const uint4* src = reinterpret_cast<const uint4*>(ops + i);
    for (int w = threadIdx.x; w < MG_OP_WORDS; w += blockDim.x) s_op[i & 1][w] = __ldg(src + w);
  };
  fetch(0);
  __syncthreads();
  unsigned int target = epoch0;
  for (int i = 0; i < nops; ++i) {
    const MegaOp& op = *reinterpret_cast<const MegaOp*>(s_op[i & 1]);
    if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      trace[2 * i] = t;
    }
    if (i + 1 < nops) fetch(i + 1);   // lands before the barrier's __syncthreads; op i+2 overwrites this slot only after it
    switch (op.kind) {
      case MK_SKINNY: {
        // op.pad = n-tiles (8 columns each) per CTA, chosen at recording time so that a layer is ONE round of tiles
        auto sync = [] { __syncthreads(); };
        for (int vb = blockIdx.x; vb < op.nblocks; vb += gridDim.x) {
          if (op.mtot <= 8) {
            switch (op.pad) {
              case 1: gemm_skinny_body<8, false, 1>(op.u.g, op.mtot, vb, threadIdx.x, mg_smem, sync); break;
              case 2: gemm_skinny_body<8, false, 2>(op.u.g, op.mtot, vb, threadIdx.x, mg_smem, sync); break;
              case 3: gemm_skinny_body<8, false, 3>(op.u.g, op.mtot, vb, threadIdx.x, mg_smem, sync); break;
              default: gemm_skinny_body<8, false, 4>(op.u.g, op.mtot, vb, threadIdx.x, mg_smem, sync); break;
            }
          } else if (op.pad == 1) {
            gemm_skinny_body<16, false, 1>(op.u.g, op.mtot, vb, threadIdx.x, mg_smem, sync);
          } else {
            gemm_skinny_body<16, false, 2>(op.u.g, op.mtot, vb, threadIdx.x, mg_smem, sync);
          }
          __syncthreads();   // shared memory is reused by the CTA's next tile
        }
        break;
      }
      case MK_ATTN:
        for (int vb = blockIdx.x; vb < op.nblocks; vb += gridDim.x) {
          attention_warp_body<64, 8, true, ATTN_KSPLIT>(op.u.a, vb);   // recorded only with the split workspace, Tq == 8
          __syncthreads();
        }
        break;
      case MK_LN: {
        const MegaLn& l = op.u.l;
        const long long gw = static_cast<long long>(blockIdx.x) * (MG_THREADS / 32) + (threadIdx.x >> 5);
        for (long long row = gw; row < l.rows; row += static_cast<long long>(gridDim.x) * (MG_THREADS / 32))
          layer_norm_body<1>(l.x, l.ldx, l.rows, l.rows_per_batch, l.C, l.gamma, l.beta, l.eps, l.silu, l.out16, l.ld16,
                             l.out_batch_pitch, row);
        break;
      }
      case MK_RVQ:
        op_rvq(op.u.r, s_idx);
        break;
      case MK_OLA: {
        const OlaDesc& d = op.u.o;
        const int pad = (d.n_fft - d.hop) / 2;
        const int n_grid = d.T * d.hop + pad;          // upper bound over first / last (istft_overlap_add, ctrl form)
        const long long total = static_cast<long long>(d.B) * n_grid;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
             i += static_cast<long long>(gridDim.x) * blockDim.x)
          overlap_add_sample(d, 0, 0, n_grid, static_cast<int>(i / n_grid), static_cast<int>(i % n_grid));
        break;
      }
      case MK_ROLL:
        op_roll(op.u.ro);
        break;
      default:
        break;
    }
    if (trace != nullptr && blockIdx.x == 0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        trace[2 * i + 1] = t;
      }
    }
    if (i + 1 < nops) {
      target += gridDim.x;
      grid_barrier(bar, target);
    }
  }
}

__global__ void __launch_bounds__(256) state_roll_kernel(const MegaRoll ro) { op_roll(ro); }

int g_mega_grid = 0;
int g_mega_group_smem = 0;
int g_mega_smem = 0;

}  // namespace

// Never fatal: the persistent step kernel is an opt-in path; if it cannot run here the grid stays 0 and the engine keeps
// replaying the CUDA graph.
int stream_mega_init() {
  g_mega_grid = 0;
  int dev = 0, sms = 0, coop = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || coop == 0) {
    cudaGetLastError();
    return FRT2_OK;
  }
  // the largest tile: K = 4096 (plain) or a fused-LayerNorm tile with up to 2048 channels, up to 4 n-tiles
  g_mega_group_smem = 0;
  g_mega_smem = static_cast<int>(std::max(sk_smem_bytes(8, SK_KCHUNK, 2048, 4), sk_smem_bytes(16, SK_KCHUNK, 2048, 2)));
  if (cudaFuncSetAttribute(stream_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_mega_smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream_step_kernel, MG_THREADS, g_mega_smem) != cudaSuccess ||
      per_sm < 1) {
    cudaGetLastError();
    return FRT2_OK;
  }
  g_mega_grid = sms;   // one CTA per SM
  return FRT2_OK;
}

int stream_mega_grid() { return g_mega_grid; }

int stream_state_roll(const MegaRoll& ro, cudaStream_t stream) {
  FRT2_REQUIRE(ro.E % 8 == 0 && ro.n_fft % 4 == 0 && ro.T >= 3 && ro.tb.n <= 16, FRT2_ERR_BAD_ARG, "state roll: bad shape");
  // enough threads that every 16-byte piece has its own (one round trip for the whole roll), at least one per item
  long long pieces = 3LL * ro.n_fft / 4 * ro.B;
  for (int i = 0; i < ro.tb.n; ++i) pieces += static_cast<long long>(ro.B) * ro.tb.e[i].hist * (ro.E / 8);
  const long long threads = std::max<long long>(pieces, ro.B);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((threads + 255) / 256, 4096));
  state_roll_kernel<<<grid, 256, 0, stream>>>(ro);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

int stream_mega_launch(const MegaOp* ops, int nops, unsigned int* bar, unsigned int epoch0, cudaStream_t stream,
                       long long* trace) {
  FRT2_REQUIRE(g_mega_grid > 0, FRT2_ERR_NOT_FINALIZED, "stream_mega_init was not called");
  if (nops <= 0) return FRT2_OK;
  void* args[] = {&ops, &nops, &bar, &epoch0, &g_mega_group_smem, &trace};
  FRT2_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(stream_step_kernel), dim3(g_mega_grid),
                                           dim3(MG_THREADS), args, static_cast<size_t>(g_mega_smem), stream));
  return FRT2_OK;
}

}  // namespace frt2
