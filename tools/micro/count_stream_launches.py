import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fireredtts2_b200 import _native as N
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
codec = RedCodecB200(C0, synthetic_state_dict(C0, 0), check_indices=False, stream_max_tokens=64)
tok = torch.from_numpy(synthetic_tokens(C0, 1, 4, 3)).cuda()
cache = codec.new_stream(1)
a, cache = codec.decode_one_token(tok[:, :, 0:1], cache, False)
codec.profile(True)
a, cache = codec.decode_one_token(tok[:, :, 1:2], cache, False)
torch.cuda.synchronize()
for c, n in N.PROF_NAMES.items():
    r = codec.profile_get(c)
    if r["launches"]:
        print(n, r["launches"], round(r["ms"] * 1e3, 1), "us")
print("all", codec.profile_get(N.PROF_ALL)["launches"])
