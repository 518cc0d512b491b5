"""Slot-pool (continuous batching) benchmark, C0 architecture: N concurrent streams advance one token per step.

Reports per step: GPU time (CUDA events) and end-to-end wall time from pinned host tokens to samples on the host,
the aggregate real-time factor (N x 80 ms of audio per step) and the worst parity of a sampled stream against the
unbatched decode of the same tokens.  usage: python tools/pool_bench.py [--slots 1,2,16,32,64] [--steps 48] [--ctx 0]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fireredtts2_b200 import _native as N
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.weights import synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--slots", default="1,2,8,16,32,64,128")
ap.add_argument("--steps", type=int, default=48)
ap.add_argument("--ctx", type=int, default=0, help="tokens every stream has already consumed when timing starts")
ap.add_argument("--eager", action="store_true")
ap.add_argument("--pcm16", action="store_true")
args = ap.parse_args()

cfg = C0
codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), check_indices=False,
                     stream_max_tokens=args.ctx + args.steps + 8)
if args.eager:
    codec.set_debug(N.DBG_NO_GRAPH)
rng = np.random.default_rng(7)
results = []
for slots in [int(x) for x in args.slots.split(",")]:
    total = args.ctx + args.steps
    tok_h = torch.from_numpy(rng.integers(0, cfg.codebook_size, size=(total, slots, cfg.num_quantizers))
                             .astype(np.int32)).pin_memory()
    pool = codec.new_pool(slots)
    width = pool.width
    host_out = torch.empty((slots, width), dtype=torch.int16 if args.pcm16 else torch.float32).pin_memory()
    first = [N.SLOT_ACTIVE | N.SLOT_RESET] * slots
    mid = [N.SLOT_ACTIVE] * slots
    got = []
    # context + warm-up (graph capture happens on the first step)
    for i in range(args.ctx):
        pool.step_dense(tok_h[i].cuda(non_blocking=True), first if i == 0 else mid)
    torch.cuda.synchronize()
    gpu_us, wall_us = [], []
    for i in range(args.ctx, total):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        d_tok = tok_h[i].cuda(non_blocking=True)
        e0.record()
        out, n = pool.step_dense(d_tok, first if i == 0 else mid, pcm16=args.pcm16)
        e1.record()
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        wall_us.append(1e6 * (time.perf_counter() - t0))
        gpu_us.append(1e3 * e0.elapsed_time(e1))
        if not args.pcm16:
            got.append(host_out[0, :n[0]].clone().numpy())
    warm = 8
    g, w = np.array(gpu_us[warm:]), np.array(wall_us[warm:])
    rec = {"slots": slots, "ctx_tokens": args.ctx, "steps": len(g), "gpu_us_p50": float(np.median(g)),
           "wall_us_p50": float(np.median(w)), "wall_us_p99": float(np.percentile(w, 99)),
           "audio_s_per_s": float(slots * 0.08 / (np.median(w) * 1e-6)),
           "rtf_per_stream": float(0.08 / (np.median(w) * 1e-6))}
    if got and args.ctx == 0:
        # parity of slot 0's stream against the offline decode of the same tokens
        t0_ = tok_h[:, 0, :].numpy().T[None]                        # (1, nq, total)
        ref = codec.decode(torch.from_numpy(np.ascontiguousarray(t0_)).cuda()).cpu().numpy()[0]
        cat = np.concatenate(got)
        ref = ref[:cat.shape[0]]
        err = ref.astype(np.float64) - cat
        rec["snr_db_vs_offline"] = float(10 * np.log10((ref.astype(np.float64) ** 2).sum() / (err ** 2).sum()))
    results.append(rec)
    print(json.dumps(rec), flush=True)
    pool.destroy()
    del pool
    torch.cuda.empty_cache()
