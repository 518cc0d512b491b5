"""BASELINE.json configs[4]: bulk synthetic-data generation, 4096 x 20 s utterances sharded over the ranks
(512 per rank at 8 GPUs -> 8 batches of 64).  No collective in the data path; `--gather` additionally sends every
rank's waveforms to rank 0 over NCCL.  Launch with torchrun.  One JSON line on rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.weights import synthetic_state_dict


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_utt = int(os.environ.get("N_UTT", "4096"))
    L = 250
    gather = "--gather" in sys.argv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = C0
    codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), device=f"cuda:{local}", check_indices=False)
    mine = list(range(rank, n_utt, world))          # equal-length units: round-robin == LPT
    g = torch.Generator().manual_seed(100 + rank)
    tok = torch.randint(0, cfg.codebook_size, (len(mine), cfg.num_quantizers, L), generator=g, dtype=torch.int32).to(dev)
    out = torch.empty((len(mine), cfg.samples_per_token * L), dtype=torch.float32, device=dev)
    codec.decode(tok[:64])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(0, len(mine), 64):
        out[i:i + 64] = codec.decode(tok[i:i + 64])
    torch.cuda.synchronize()
    t_dec = time.perf_counter() - t0
    t_gather = 0.0
    if gather and world > 1:
        dist.barrier()
        t1 = time.perf_counter()
        if rank == 0:
            bufs = [out] + [torch.empty_like(out) for _ in range(world - 1)]
            for r in range(1, world):
                dist.recv(bufs[r], src=r)
        else:
            dist.send(out, dst=0)
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - t1
    t = torch.tensor([t_dec, t_gather], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        audio_s = n_utt * L / 12.5
        print(json.dumps({"workload": f"configs[4]: {n_utt} x 20 s utterances", "n_gpus": world,
                          "decode_seconds": float(t[0]), "audio_s_per_s": audio_s / float(t[0]),
                          "gather_seconds": float(t[1]), "gathered_bytes": int(out.numel() * 4 * (world - 1)) if gather else 0,
                          "finite": bool(torch.isfinite(out).all())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
