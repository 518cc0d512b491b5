"""Frame tail of the speech LM (csrc/frame_decoder.cu): device time per frame and the HBM roofline of its weight streams.

    python tools/frame_decoder_bench.py [--preset FD_200M] [--batch 1] [--frames 50]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fireredtts2_b200.frame_decoder import (FD_PRESETS, FrameDecoderB200, synthetic_frame_decoder_state_dict,  # noqa: E402
                                            synthetic_frame_inputs)


def run(preset="FD_200M", batch=1, frames=50, warmup=5, peak_gbs=None):
    cfg = FD_PRESETS[preset]
    fd = FrameDecoderB200(cfg, synthetic_frame_decoder_state_dict(cfg, 0), max_batch=max(8, batch))
    last_h, _ = synthetic_frame_inputs(cfg, batch, 0)
    last_h = torch.from_numpy(last_h).cuda()
    for _ in range(warmup):
        fd.generate_codes(last_h, 30, 0.9, seed=1)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(frames + 1)]
    ev[0].record()
    for i in range(frames):
        fd.generate_codes(last_h, 30, 0.9, seed=1)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(frames)])
    bytes_frame = cfg.weight_bytes_per_frame()
    rec = {"preset": preset, "batch": batch, "frames": frames, "ms_per_frame_median": float(np.median(ms)),
           "ms_per_frame_min": float(ms.min()), "launches_per_frame": fd.last_launches,
           "us_per_launch": float(np.median(ms)) * 1e3 / fd.last_launches,
           "weight_bytes_per_frame": bytes_frame, "achieved_gbs": bytes_frame / (float(np.median(ms)) * 1e-3) / 1e9,
           "frames_per_second": batch * 1e3 / float(np.median(ms)), "realtime_factor": batch * 80.0 / float(np.median(ms))}
    if peak_gbs:
        rec["frac_of_hbm_peak"] = rec["achieved_gbs"] / peak_gbs
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="FD_200M")
    ap.add_argument("--batch", type=int, nargs="+", default=[1])
    ap.add_argument("--frames", type=int, default=50)
    a = ap.parse_args()
    peak = None
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs")
    except Exception:
        pass
    for b in a.batch:
        print(json.dumps(run(a.preset, b, a.frames, peak_gbs=peak)))
