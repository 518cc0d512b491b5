"""In-kernel timeline of the persistent tcgen05 attention kernel (CTA 0): clock64 stamps of every softmax warp and of
the MMA-issuing thread per key tile.  usage: python tools/attn_trace.py [first_idx] [count]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fireredtts2_b200 import _native as N

lib = N.load()
P = lambda t: C.c_void_p(t.data_ptr())
B, H, T = 64, 16, 3000
E = H * 64
qkv = torch.randn(B, T, 3 * E, device="cuda").half()
q, k, v = [qkv[..., i * E:(i + 1) * E].contiguous() for i in range(3)]
out = torch.empty(B, T, E, device="cuda", dtype=torch.half)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    N.check(lib.frt2_op_attention(0, P(q), P(k), P(v), P(out), B, H, 64, T, T, 0, 1, st))
torch.cuda.synchronize()
buf = torch.zeros(16 * 128 * 8, device="cuda", dtype=torch.int32)
N.check(lib.frt2_op_attention_trace(P(buf)))
N.check(lib.frt2_op_attention(0, P(q), P(k), P(v), P(out), B, H, 64, T, T, 0, 1, st))
torch.cuda.synchronize()
N.check(lib.frt2_op_attention_trace(None))
tr = buf.cpu().numpy().astype(np.uint32).reshape(16, 128, 8).astype(np.int64)
i0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 12
base = tr[0, i0, 0]
names = ["top", "s_full", "s_empty", "max", "exp", "pv_wait", "p_full"]
print("softmax warps: stamps relative to warp 0's loop top of key tile", i0, "(clk); columns:", names)
for w in (0, 4, 8, 1, 5, 9):
    print(f"warp {w} (tile {w >> 2}, sub-partition {w & 3})")
    for i in range(i0, i0 + cnt):
        r = (tr[w, i, :7] - base) & 0xffffffff
        d = np.diff(np.concatenate([[(tr[w, i - 1, 6] - base) & 0xffffffff], r]))
        print(f"  kt {i:3d}: " + " ".join(f"{int(x):7d}" for x in r) + "   | deltas " + " ".join(f"{int(x):5d}" for x in d[1:]))
print("MMA thread: issue time of QK_i and PV_i per tile, relative to the same origin")
for t in range(3):
    for i in range(i0, i0 + cnt):
        qk, pv = [(tr[12 + t, i, c] - base) & 0xffffffff for c in (0, 1)]
        print(f"  tile {t} kt {i:3d}: QK {int(qk):7d}  PV {int(pv):7d}")
per = [(tr[w, i0 + cnt, 0] - tr[w, i0, 0]) / cnt for w in range(12)]
print("mean period per key tile (clk) per softmax warp:", [int(x) for x in per])
