// K2s — skinny GEMM for the batch-1 streaming step (<= 16 output rows in total).
//
// With 8 rows per codec token the per-token step is a sequence of GEMVs: every weight byte is used once, so the
// step is bound by streaming the 429 MB of fp16 weights from HBM (>= 66 us at 6.5 TB/s), not by tensor throughput.
// A 128-row tcgen05 tile would leave all but N/256 SMs idle; here each CTA owns 8 output columns so a 1024-column
// layer spreads over 128 SMs, the 8 warps of the CTA split K, every lane streams 16-byte weight vectors (all of a
// warp's vectors in flight at once), the (im2col'd, causal) activation rows sit in shared memory, and the products run
// on mma.sync.m16n8k16 (m = the <= 16 rows, n = the 8 columns) with fp32 accumulation.  Same contract / epilogues as
// gemm_tc.
#include <cstdlib>

#include "common.cuh"
#include "gemm_skinny_body.cuh"

namespace frt2 {

namespace {

template <int MR, int NT>
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(GemmDesc g, int mtot) {
  extern __shared__ __align__(16) uint8_t sk_smem_dyn[];
  gemm_skinny_body<MR, true, NT>(g, mtot, blockIdx.x, threadIdx.x, sk_smem_dyn, [] { __syncthreads(); });
}

}  // namespace

int gemm_skinny_init() {
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 1))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(16, SK_KCHUNK, SK_KCHUNK, 1))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 2))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(16, SK_KCHUNK, SK_KCHUNK, 2))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 3))));
  FRT2_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sk_smem_bytes(8, SK_KCHUNK, SK_KCHUNK, 4))));
  return FRT2_OK;
}

bool gemm_skinny_applicable(const GemmDesc& g) {
  return static_cast<long long>(g.batches) * g.rows_out <= SK_MAXROWS && g.Kc % 8 == 0 &&
         (g.act != ACT_POLAR || g.N % 2 == 0);
}

int gemm_skinny(const GemmDesc& g, cudaStream_t stream) {
  FRT2_REQUIRE(gemm_skinny_applicable(g), FRT2_ERR_BAD_ARG, "gemm_skinny: more than 16 rows");
  FRT2_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0 &&
                   g.a_row_pitch % 8 == 0 && g.a_batch_pitch % 8 == 0,
               FRT2_ERR_BAD_ARG, "gemm_skinny: operands must be 16-byte aligned");
  const int mtot = g.batches * g.rows_out;
  const int Ktot = g.ntaps * g.Kc;
  const int kchunk = Ktot < SK_KCHUNK ? Ktot : SK_KCHUNK;
  const int grid = (g.N + SK_COLS - 1) / SK_COLS;
  FRT2_REQUIRE(g.ln_gamma == nullptr || (g.ntaps == 1 && g.Kc <= SK_KCHUNK && g.Kc % 4 == 0 && g.ln_x != nullptr),
               FRT2_ERR_BAD_ARG, "gemm_skinny: fused LayerNorm needs ntaps == 1 and Kc <= 4096");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_WARPS * 32);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap the weight prefetch with the
  attr[0].val.programmaticStreamSerializationAllowed = 1;            // tail of the previous kernel of the step
  // Programmatic dependent launch: the next skinny kernel's prologue (all of its weight vectors, gamma/beta) is in flight
  // while the current kernel drains.  With the round-1 kernels this cost ~100 us per step inside the captured graph
  // (early-resident dependents competed for SM slots); with the current ones it saves ~20 (557 -> 538 us per token, full
  // GPU suite green with it), so it is on by default; FRT2_NO_PDL=1 restores plain stream order (A/B).
  static const bool use_pdl = (getenv("FRT2_NO_PDL") == nullptr);
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  // wide layers (QKV, fc1): 16 columns per CTA — one wave of CTAs, activation rows staged / normalised half as often
  static const bool no_nt2 = (getenv("FRT2_SKINNY_NT1") != nullptr);   // A/B switch for measurements
  static const int nt2_min_n = [] {
    const char* e = getenv("FRT2_SKINNY_NT2_MIN_N");
    return e != nullptr ? atoi(e) : 2048;
  }();
  static const int nt_max = [] {   // up to 3 / 4 n-tiles per CTA: one CTA per SM for N = 3072 / 4096 (A/B: =2 -> 562 vs 554 us)
    const char* e = getenv("FRT2_SKINNY_NTMAX");
    return e != nullptr ? atoi(e) : 4;
  }();
  int nt = (g.N >= nt2_min_n && !no_nt2) ? 2 : 1;
  if (nt == 2 && nt_max > 2 && mtot <= 8) nt = std::max(2, std::min(nt_max, (g.N / 8 + 147) / 148));
  cfg.gridDim = dim3((g.N + SK_COLS * nt - 1) / (SK_COLS * nt));
  const int ln_c = g.ln_gamma != nullptr ? g.Kc : 0;
  if (mtot <= 8) {
    cfg.dynamicSmemBytes = sk_smem_bytes(8, kchunk, ln_c, nt);
    switch (nt) {
      case 4: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 4>, g, mtot)); break;
      case 3: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 3>, g, mtot)); break;
      case 2: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 2>, g, mtot)); break;
      default: FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<8, 1>, g, mtot)); break;
    }
  } else {
    cfg.dynamicSmemBytes = sk_smem_bytes(16, kchunk, ln_c, nt);
    if (nt == 2) FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<16, 2>, g, mtot));
    else FRT2_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<16, 1>, g, mtot));
  }
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
