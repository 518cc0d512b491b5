"""Per-token streaming step micro-benchmark (BASELINE configs[1]); `--eager` disables the CUDA graph (for ncu lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fireredtts2_b200 import _native as N
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens

eager = "--eager" in sys.argv
steps = int(os.environ.get('STEPS', '16'))
codec = RedCodecB200(C0, synthetic_state_dict(C0, 0), check_indices=False, stream_max_tokens=max(64, steps + 8))
if eager:
    codec.set_debug(N.DBG_NO_GRAPH)
tok = torch.from_numpy(synthetic_tokens(C0, 1, steps, 3)).cuda()
state = codec.new_stream(1)
for rep in range(3):
    codec.reset_stream(state)
    cache = state
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps - 16):
        a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, False)
    e0.record()
    for i in range(steps - 16, steps):
        a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, False)
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: tokens {steps - 16}..{steps} (context {8 * (steps - 16)} frames): GPU {e0.elapsed_time(e1) / 16 * 1e3:.1f} us/token")
