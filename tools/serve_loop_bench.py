"""Serving loop behind the LLM backbone for N concurrent streams (SURVEY 8f.2 + 8f.4): per 80 ms step the frame tail
(frt2_fd_generate, one frame for all N streams) produces every stream's 16 codes on the device, the slot pool
(frt2_pool_step) turns them into every stream's audio chunk, the chunks go to pinned host memory as int16 PCM.  The codes
never visit the host; the backbone in front (one last_h per stream and step) is simulated by resident random states.

    python tools/serve_loop_bench.py [--slots 1,8,16,64,128] [--steps 40]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fireredtts2_b200 import _native as N  # noqa: E402
from fireredtts2_b200.codec import RedCodecB200  # noqa: E402
from fireredtts2_b200.config import C0  # noqa: E402
from fireredtts2_b200.frame_decoder import FD_200M, FrameDecoderB200, synthetic_frame_decoder_state_dict  # noqa: E402
from fireredtts2_b200.weights import synthetic_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slots", default="1,8,16,64,128")
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    slots_list = [int(x) for x in a.slots.split(",")]
    codec = RedCodecB200(C0, synthetic_state_dict(C0, 0), check_indices=False, stream_max_tokens=a.steps + 16)
    tail = FrameDecoderB200(FD_200M, synthetic_frame_decoder_state_dict(FD_200M, 0), max_batch=max(slots_list))
    g = torch.Generator(device="cuda").manual_seed(3)
    for slots in slots_list:
        pool = codec.new_pool(slots)
        host = torch.empty((slots, pool.width), dtype=torch.int16).pin_memory()
        states = torch.randn(a.steps + 8, slots, FD_200M.backbone_dim, device="cuda", generator=g)
        first, mid = [N.SLOT_ACTIVE | N.SLOT_RESET] * slots, [N.SLOT_ACTIVE] * slots
        wall, t_tail, t_pool = [], [], []
        for i in range(a.steps + 8):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev[0].record()
            codes = tail.generate_codes(states[i], 30, 0.9, seed=7)           # (slots, 16) int32, on the device
            ev[1].record()
            out, n = pool.step_dense(codes, first if i == 0 else mid, pcm16=True)
            ev[2].record()
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            t_tail.append(ev[0].elapsed_time(ev[1]))
            t_pool.append(ev[1].elapsed_time(ev[2]))
        w = float(np.median(wall[8:]))
        print(json.dumps({"slots": slots, "steps": a.steps, "step_ms_p50": w, "step_ms_p99": float(np.percentile(wall[8:], 99)),
                          "frame_tail_ms": float(np.median(t_tail[8:])), "codec_pool_ms": float(np.median(t_pool[8:])),
                          "audio_s_per_s": slots * 0.08 / (w * 1e-3), "x_realtime_per_stream": 80.0 / w,
                          "samples_per_chunk": int(n[0]), "pcm_nonzero": bool((host != 0).any())}), flush=True)
        pool.destroy()
        del pool
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
