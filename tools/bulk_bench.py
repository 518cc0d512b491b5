"""BASELINE.json configs[4]: bulk synthetic-data generation, 4096 x 20 s utterances sharded over the ranks
(512 per rank at 8 GPUs -> 8 batches of 64).  No collective in the data path.  `--gather` additionally sends every
rank's waveforms to rank 0 over NCCL after the decode (one warm-up exchange first: the first send/recv of a pair pays
the NCCL connection set-up).  `--peer` instead decodes straight INTO rank 0's buffer: the overlap-add kernel's stores
go over NVLink peer memory (frt2_decode_scatter + sharding.PeerBuffer), so there is no exchange step after the compute;
`--pcm16` emits int16 PCM (half the bytes).  Launch with torchrun.  One JSON line on rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.sharding import PeerBuffer
from fireredtts2_b200.weights import synthetic_state_dict


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_utt = int(os.environ.get("N_UTT", "4096"))
    L = 250
    gather = "--gather" in sys.argv
    peer = "--peer" in sys.argv
    pcm16 = "--pcm16" in sys.argv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = C0
    codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), device=f"cuda:{local}", check_indices=False)
    mine = list(range(rank, n_utt, world))          # equal-length units: round-robin == LPT
    g = torch.Generator().manual_seed(100 + rank)
    tok = torch.randint(0, cfg.codebook_size, (len(mine), cfg.num_quantizers, L), generator=g, dtype=torch.int32).to(dev)
    n_per = cfg.samples_per_token * L
    odt = torch.int16 if pcm16 else torch.float32
    esz = 2 if pcm16 else 4
    codec.decode(tok[:64])
    torch.cuda.synchronize()
    if peer:
        # unit u = rank + world*k lives at element offset u*n_per of ONE (n_utt, n_per) buffer on rank 0
        buf = PeerBuffer(n_utt * n_per, odt, dev, None, 0)
        offs = torch.tensor([u * n_per for u in mine], dtype=torch.int64, device=dev)
        codec.decode_into(tok[:64], buf.ptr, offs[:64], pcm16=pcm16)      # warm-up incl. the peer mapping
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(0, len(mine), 64):
            codec.decode_into(tok[i:i + 64], buf.ptr, offs[i:i + 64], pcm16=pcm16)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_dec = time.perf_counter() - t0
        t = torch.tensor([t_dec, 0.0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            full = buf.tensor().view(n_utt, n_per)
            # spot check: a few units of every rank against a local decode of the same tokens
            err = 0.0
            for r in range(world):
                gr = torch.Generator().manual_seed(100 + r)
                tr = torch.randint(0, cfg.codebook_size, (len(range(r, n_utt, world)), cfg.num_quantizers, L),
                                   generator=gr, dtype=torch.int32)[:2].to(dev)
                ref = codec.decode(tr, pcm16=pcm16)
                got = torch.stack([full[r], full[r + world]])
                err = max(err, float((ref.float() - got.float()).abs().max()))
            audio_s = n_utt * L / 12.5
            print(json.dumps({"workload": f"configs[4]: {n_utt} x 20 s utterances", "n_gpus": world,
                              "mode": "peer-memory scatter (overlap-add stores over NVLink into rank 0's buffer)",
                              "dtype": "int16 pcm" if pcm16 else "fp32",
                              "decode_and_gather_seconds": float(t[0]), "audio_s_per_s": audio_s / float(t[0]),
                              "gathered_bytes": int(n_utt * n_per * esz * (world - 1) // world),
                              "max_abs_vs_local_decode": err,
                              "finite": bool(torch.isfinite(full.float()).all())}), flush=True)
        if world > 1:
            dist.barrier()
        buf.close()
        if world > 1:
            dist.destroy_process_group()
        return
    out = torch.empty((len(mine), n_per), dtype=odt, device=dev)
    if gather and world > 1:      # connection set-up of every (r, 0) pair outside the timed region
        w = torch.zeros(1024, device=dev)
        if rank == 0:
            for r in range(1, world):
                dist.recv(w, src=r)
        else:
            dist.send(w, dst=0)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(0, len(mine), 64):
        out[i:i + 64] = codec.decode(tok[i:i + 64], pcm16=pcm16)
    torch.cuda.synchronize()
    t_dec = time.perf_counter() - t0
    t_gather = 0.0
    if gather and world > 1:
        if rank == 0:
            bufs = [out] + [torch.empty_like(out) for _ in range(world - 1)]
            torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        if rank == 0:
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
                          "gather_seconds": float(t[1]), "gathered_bytes": int(out.numel() * esz * (world - 1)) if gather else 0,
                          "mode": "nccl send/recv after the decode" if gather else "no gather",
                          "dtype": "int16 pcm" if pcm16 else "fp32",
                          "finite": bool(torch.isfinite(out.float()).all())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
