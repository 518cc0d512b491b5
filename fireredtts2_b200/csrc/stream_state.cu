// End-of-step state kernels of the streaming decode (reference RedCodecInfer.decode_one_token + cache_dict,
// codec/model.py:326-376): what the reference does with cat / slice / stack on its five cache tensors is, on an in-HBM
// state that is updated in place, one "roll" kernel per step (new iSTFT tail, conv histories to the head of their
// [history | chunk] buffers, position advance) and one "reset" kernel per new stream (zero history = the causal left
// padding, position 0).  Both are stream-ordered: no host synchronisation, no cudaMemset on the null stream.
#include <algorithm>

#include "common.cuh"

namespace frt2 {

namespace {

__device__ __forceinline__ void state_roll(const StateRoll& ro) {
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

__global__ void __launch_bounds__(256) state_roll_kernel(const StateRoll ro) { state_roll(ro); }

// A stream (or every item of a batch of streams) back to "no token consumed": zero conv history rows (the K/V state and
// the iSTFT tail need no clearing: nothing before position 0 is ever read), control blocks and error words cleared.
__global__ void __launch_bounds__(256) state_reset_kernel(const ShiftTable tb, int E, int B, int* ctrl,
                                                          unsigned int* err_words, int n_err) {
  const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const int E8 = E / 8;
  for (int ei = 0; ei < tb.n; ++ei) {
    const ShiftEntry& en = tb.e[ei];
    const long long total = static_cast<long long>(B) * en.hist * E8;
    for (long long i = gtid; i < total; i += gthreads) {
      const int c = static_cast<int>(i % E8);
      const int r = static_cast<int>((i / E8) % en.hist);
      const int b = static_cast<int>(i / (static_cast<long long>(E8) * en.hist));
      *reinterpret_cast<uint4*>(en.p + b * en.batch_pitch + static_cast<long long>(r) * E + c * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  for (long long i = gtid; i < static_cast<long long>(B) * CTRL_INTS; i += gthreads) ctrl[i] = 0;
  for (long long i = gtid; i < n_err; i += gthreads) err_words[i] = 0u;
}

}  // namespace

int stream_state_roll(const StateRoll& ro, cudaStream_t stream) {
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

int stream_state_reset(const ShiftTable& tb, int E, int B, int* ctrl, unsigned int* err_words, int n_err,
                       cudaStream_t stream) {
  FRT2_REQUIRE(E % 8 == 0 && tb.n <= 16, FRT2_ERR_BAD_ARG, "state reset: bad shape");
  long long pieces = static_cast<long long>(B) * CTRL_INTS;
  for (int i = 0; i < tb.n; ++i) pieces = std::max(pieces, static_cast<long long>(B) * tb.e[i].hist * (E / 8));
  const unsigned grid = static_cast<unsigned>(std::min<long long>((pieces + 255) / 256, 1024));
  state_reset_kernel<<<grid, 256, 0, stream>>>(tb, E, B, ctrl, err_words, n_err);
  FRT2_CUDA_OK(cudaGetLastError());
  return FRT2_OK;
}

}  // namespace frt2
