"""Debug: every unit of the config-4 dialogue decoded inside a padded var-len batch must equal its standalone decode, under
the product path and the check paths (found the folded-LayerNorm aliasing fault)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200 import _native as N
from fireredtts2_b200.sharding import dialogue_turn_lengths, partition_units, make_batches
from fireredtts2_b200.weights import synthetic_state_dict
cfg = C0
codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), device="cuda:0", check_indices=False)
lens = dialogue_turn_lengths()
g = torch.Generator().manual_seed(11)
units = [torch.randint(0, cfg.codebook_size, (cfg.num_quantizers, L), generator=g, dtype=torch.int32) for L in lens]
dev = torch.device("cuda:0")
for world in (1, 2, 8):
  plan = partition_units(lens, world)
  for flags, name in ((0, "product"), (N.DBG_NO_LNFOLD, "nofold"), (N.DBG_ATTN_WARP, "warp_attn"), (N.DBG_GEMM_REF, "gemm_ref")):
    if name == "gemm_ref" and world != 2: continue
    codec.set_debug(flags)
    refs = [codec.decode(u[None].to(dev))[0] for u in units]
    for r in range(min(world, 2)):
        for batch in make_batches(plan[r], lens, 64, 64 * 375):
            L = max(lens[i] for i in batch)
            tok = torch.zeros((len(batch), 16, L), dtype=torch.int32, device=dev)
            for k, i in enumerate(batch):
                tok[k, :, :lens[i]] = units[i].to(dev)
            ln = torch.tensor([lens[i] for i in batch], dtype=torch.int32, device=dev)
            a = codec.decode(tok, ln)
            errs = [float((a[k, :1920 * lens[i]] - refs[i]).abs().max()) for k, i in enumerate(batch)]
            print(name, "world", world, "rank", r, "B", len(batch), "L", L, "lens", [lens[i] for i in batch])
            print("   errs", ["%.2e" % e for e in errs], flush=True)
