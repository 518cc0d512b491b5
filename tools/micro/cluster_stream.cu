// Feasibility probe for a per-token step that lives on ONE thread-block cluster: how fast can the CTAs of a
// cluster stream a weight set through shared-memory rings with cp.async.bulk (L2 prefetch running ahead), and what does
// a cluster barrier cost?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_stream cluster_stream.cu
//   ./cluster_stream [MB=429] [cluster=16] [clusters=1] [stage_kb=32] [stages=6] [prefetch_mb_ahead=8]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// every CTA streams its contiguous share of `bytes`; 1 producer thread, 8 consumer warps that read every byte from smem
__global__ void __launch_bounds__(288) stream_kernel(const uint8_t* w, size_t bytes, int stage_bytes, int stages, size_t pf_ahead,
                                                     int barrier_every, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  uint8_t* ring = smem + 256;
  const int nctas = gridDim.x;
  const size_t share = (bytes / nctas) & ~static_cast<size_t>(stage_bytes - 1);
  const uint8_t* mine = w + share * blockIdx.x;
  const int nstage = static_cast<int>(share / stage_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) {
    if (lane == 0) {
      size_t pf = 0;
      for (int i = 0; i < nstage; ++i) {
        const int s = i % stages;
        const uint32_t ph = (i / stages) & 1;
        // keep the L2 prefetch `pf_ahead` bytes in front of the smem fills
        while (pf < share && pf < static_cast<size_t>(i) * stage_bytes + pf_ahead) { bulk_prefetch_l2(mine + pf, stage_bytes); pf += stage_bytes; }
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect(&full[s], stage_bytes);
        bulk_load(ring + static_cast<size_t>(s) * stage_bytes, mine + static_cast<size_t>(i) * stage_bytes, stage_bytes, &full[s]);
      }
    }
  } else {
    unsigned acc = 0;
    for (int i = 0; i < nstage; ++i) {
      const int s = i % stages;
      const uint32_t ph = (i / stages) & 1;
      mbar_wait(&full[s], ph);
      const uint4* p = reinterpret_cast<const uint4*>(ring + static_cast<size_t>(s) * stage_bytes);
      const int n16 = stage_bytes / 16;
      for (int j = warp * 32 + lane; j < n16; j += 256) { const uint4 v = p[j]; acc += v.x ^ v.y ^ v.z ^ v.w; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (barrier_every > 0 && (i + 1) % barrier_every == 0) {
        // stand-in for the end of an op: all consumer warps of all CTAs of the cluster meet
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    if (acc == 0x12345678u) sink[0] = acc;
  }
  cluster_sync();
}

__global__ void barrier_kernel(int iters, long long* out) {
  cluster_sync();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) cluster_sync();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 429;
  const int csize = argc > 2 ? atoi(argv[2]) : 16;
  const int nclusters = argc > 3 ? atoi(argv[3]) : 1;
  const int stage_kb = argc > 4 ? atoi(argv[4]) : 32;
  const int stages = argc > 5 ? atoi(argv[5]) : 6;
  const size_t pf_mb = argc > 6 ? atoi(argv[6]) : 8;
  const size_t bytes = mb << 20;
  uint8_t* w;
  unsigned* sink;
  long long* clk;
  cudaMalloc(&w, bytes);
  cudaMemset(w, 1, bytes);
  cudaMalloc(&sink, 64);
  cudaMalloc(&clk, 64);
  uint8_t* flush;
  cudaMalloc(&flush, 512u << 20);
  const int smem = 256 + stages * stage_kb * 1024;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(barrier_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(csize * nclusters); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int barrier_every = 0; barrier_every <= 4; barrier_every += 4) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemset(flush, rep, 512u << 20);   // evict the weights from L2
      cudaEventRecord(e0);
      cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel, (const uint8_t*)w, bytes, stage_kb * 1024, stages, pf_mb << 20, barrier_every, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("{\"probe\": \"cluster_stream\", \"mb\": %zu, \"cluster\": %d, \"clusters\": %d, \"stage_kb\": %d, \"stages\": %d, \"pf_mb\": %zu, \"cta_barrier_every\": %d, \"us\": %.1f, \"gbs\": %.0f}\n",
           mb, csize, nclusters, stage_kb, stages, pf_mb, barrier_every, best * 1e3, bytes / (best * 1e-3) / 1e9);
  }
  cfg.gridDim = dim3(csize); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
  cudaLaunchKernelEx(&cfg, barrier_kernel, 1000, clk);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"probe\": \"cluster_barrier\", \"cluster\": %d, \"clocks\": %lld, \"err\": \"%s\"}\n", csize, h, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
