"""One-GPU proxy of the sharded dialogue (BASELINE configs[3]): for N = 1, 2, 4, 8 decode every rank's share of the
24 turns on THIS GPU, one share at a time, exactly as `decode_sharded_peer` would on rank r (local scatter instead of
peer stores), and report the slowest share = what an N-GPU run takes apart from the barrier.  Also prints the kernel
classes of the slowest 8-way share.  usage: python tools/dialogue_proxy.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.sharding import dialogue_turn_lengths, make_batches, partition_units, unit_offsets
from fireredtts2_b200.weights import synthetic_state_dict

dev = torch.device("cuda", 0)
cfg = C0
codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), device="cuda:0", check_indices=False)
lens = dialogue_turn_lengths()
g = torch.Generator().manual_seed(11)
units = [torch.randint(0, cfg.codebook_size, (cfg.num_quantizers, L), generator=g, dtype=torch.int32).to(dev) for L in lens]
offs = unit_offsets(lens, cfg.samples_per_token)
out = torch.empty(offs[-1], dtype=torch.float32, device=dev)


def share(idx):
    for batch in make_batches(idx, lens, 64, 64 * 375):
        L = max(lens[i] for i in batch)
        tok = torch.zeros((len(batch), cfg.num_quantizers, L), dtype=torch.int32, device=dev)
        for k, i in enumerate(batch):
            tok[k, :, :lens[i]] = units[i]
        codec.decode_into(tok, out.data_ptr(), torch.tensor([offs[i] for i in batch], dtype=torch.int64, device=dev),
                          torch.tensor([lens[i] for i in batch], dtype=torch.int32, device=dev))
    torch.cuda.synchronize()


def best(idx, reps=10):
    share(idx); share(idx)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        share(idx)
        ts.append(time.perf_counter() - t0)
    return min(ts)


res = {}
for n in (1, 2, 4, 8):
    plan = partition_units(lens, n)
    ts = [best(p) for p in plan]
    res[n] = max(ts)
    print(f"N={n}: slowest share {max(ts) * 1e3:.3f} ms (shares {[round(t * 1e3, 2) for t in ts]}; tokens {[sum(lens[i] for i in p) for p in plan]})")
print(json.dumps({"proxy_speedup_vs_1": {n: res[1] / res[n] for n in res}}))
plan = partition_units(lens, 8)
codec.profile(True)
share(plan[0])
names = {0: "gemm_tc", 1: "attention_tc", 2: "attention_warp", 3: "layer_norm", 4: "rvq", 5: "ola", 6: "gemm_skinny"}
for c, nm in names.items():
    try:
        r = codec.profile_get(c)
        if r.get("launches", 0):
            print(nm, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
    except Exception as e:  # noqa: BLE001
        print(nm, "n/a", e)
codec.profile(False)
