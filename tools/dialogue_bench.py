"""BASELINE.json configs[3]: a 3-minute 4-speaker dialogue (24 turns, 2250 tokens) with the turns sharded over the
ranks and the waveform chunks gathered to rank 0 in turn order over NCCL — or, with `--peer`, written by every rank's
overlap-add kernel straight to the turn's place in the concatenated waveform on rank 0 (NVLink peer memory,
sharding.decode_sharded_peer).  Launch with torchrun (or plain python for 1 GPU).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.sharding import decode_sharded, decode_sharded_peer, dialogue_turn_lengths, partition_units
from fireredtts2_b200.weights import synthetic_state_dict


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = C0
    codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), device=f"cuda:{local}", check_indices=False)
    lens = dialogue_turn_lengths()
    g = torch.Generator().manual_seed(11)
    units = [torch.randint(0, cfg.codebook_size, (cfg.num_quantizers, L), generator=g, dtype=torch.int32) for L in lens]
    fn = lambda tok, lengths: codec.decode(tok, lengths)
    peer = "--peer" in sys.argv
    pbuf = None
    if peer:
        _, _, pbuf = decode_sharded_peer(codec, units, dev)

    def run():
        if peer:
            flat, _, _ = decode_sharded_peer(codec, units, dev, buffer=pbuf)
            return flat
        res = decode_sharded(fn, units, dev)
        return torch.cat(res) if rank == 0 else None   # the concatenated dialogue (fireredtts2.py:401)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps, ts = 10, []
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        full = run()                        # (4 320 000,) on rank 0
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        # check against an unsharded decode of every turn on this rank
        ref = torch.cat([codec.decode(u[None].to(dev))[0] for u in units])
        err = float((ref - full).abs().max())
        plan = partition_units(lens, world)
        print(json.dumps({"workload": "configs[3]: 180 s dialogue, 24 turns, 2250 tokens", "n_gpus": world,
                          "gather": "peer-memory scatter over NVLink" if peer else "nccl send/recv",
                          "seconds": float(t[0]), "audio_s_per_s": 180.0 / float(t[0]),
                          "samples": int(full.numel()), "max_abs_vs_unsharded": err,
                          "rank_loads_tokens": [sum(lens[i] for i in p) for p in plan]}), flush=True)
    if world > 1:
        dist.barrier()
    if pbuf is not None:
        pbuf.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
