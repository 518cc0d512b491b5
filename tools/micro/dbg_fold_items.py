"""Debug: a batch of IDENTICAL items must decode to identical rows; prints, per tap, the first row where item k differs
from item 0 (folded-LayerNorm path with taps: FRT2_FOLD_WITH_TAPS=1)."""
import os, sys
os.environ["FRT2_FOLD_WITH_TAPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200 import _native as N
from fireredtts2_b200.weights import synthetic_state_dict
cfg = C0
codec = RedCodecB200(cfg, synthetic_state_dict(cfg, 0), device="cuda:0", check_indices=False)
B, L = int(os.environ.get("DB", 12)), int(os.environ.get("DL", 159))
g = torch.Generator().manual_seed(11)
one = torch.randint(0, cfg.codebook_size, (1, 16, L), generator=g, dtype=torch.int32)
tok = one.expand(B, 16, L).contiguous().cuda()
T, E = 8 * L, cfg.embed_dim
for flags, name in ((N.DBG_TAPS, "fold+taps"), (N.DBG_TAPS | N.DBG_NO_LNFOLD, "nofold+taps")):
    codec.set_debug(flags)
    a = codec.decode(tok)
    print(name, "audio item diffs", [float((a[k] - a[0]).abs().max()) for k in range(B)])
    for tap in ("prior", "layer0", "layers", "frames"):
        C_ = E if tap != "frames" else 960
        x = codec.get_tap(tap, (B, T, C_))
        out = []
        for k in range(1, B):
            d = (x[k] - x[0]).abs().amax(dim=1)
            bad = torch.nonzero(d > 0).flatten()
            out.append((k, int(bad[0]) if bad.numel() else -1, int(bad.numel()), float(d.max())))
        print("  ", tap, [o for o in out if o[1] >= 0])
