"""frt2_rvq_encode at the benchmark architecture (C0): 64 x 375 tokens (the tokens of BASELINE configs[2]) per launch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict

from fireredtts2_b200 import _native as N

cfg = C0
sd = dict(synthetic_state_dict(cfg, 0)); sd.update(synthetic_encode_tensors(cfg, 0))
codec = RedCodecB200(cfg, sd, check_indices=False)
flop_tok = 2.0 * (cfg.embed_dim * cfg.rvq_dim + cfg.num_quantizers * (2 * cfg.rvq_dim * cfg.codebook_dim + cfg.codebook_dim * cfg.codebook_size))
for B, T in ((64, 375), (96, 75), (32, 75)):
    z = torch.randn(B, cfg.embed_dim, T, device="cuda")
    rec = {"op": "rvq_encode", "tokens": B * T, "audio_s": B * T / 12.5}
    for name, dbg in (("tensor_core_chain", 0), ("cuda_core_kernel", N.DBG_GEMM_REF)):
        codec.set_debug(dbg)
        for _ in range(3):
            codes = codec.rvq_encode_codes(z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            codes = codec.rvq_encode_codes(z)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec[name] = {"ms": ms, "algorithmic_tflops": B * T * flop_tok / ms / 1e9, "audio_s_per_s": B * T / 12.5 / (ms / 1e3)}
    codec.set_debug(0)
    print(json.dumps(rec))
