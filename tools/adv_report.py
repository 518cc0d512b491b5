"""Adversarial-weights report: the adv4 fixtures (LayerNorm gamma in [0.1, 5], residual rows at ~100 +- 2, four
outlier channels x 100, GELU activations in the thousands; real-reference goldens) decoded by the product path, by the
same path with separate LayerNorm kernels, and — when run with FRT2_NO_LNSHIFT=1 — with the plain fp16(x) copy the
mean-shifted copy replaces.   usage: python tools/adv_report.py ; FRT2_NO_LNSHIFT=1 python tools/adv_report.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fireredtts2_b200 import _native as N
from fireredtts2_b200.codec import RedCodecB200
from oracle import codec_oracle as O
from tests.helpers import cases, load_case

out = {"lnshift": "off (FRT2_NO_LNSHIFT=1: plain fp16(x) copy)" if os.environ.get("FRT2_NO_LNSHIFT") else "on (product)"}
for case in cases("offline") + cases("stream"):
    if not case["name"].startswith("adv4"):
        continue
    cfg, sd, g = load_case(case)
    codec = RedCodecB200(cfg, sd, device="cuda:0", stream_max_tokens=64)
    tok = torch.from_numpy(g["tokens"]).cuda()
    rec = {}
    if case["kind"] == "offline":
        for name, dbg in (("folded_layernorm", 0), ("separate_layernorm_kernels", N.DBG_NO_LNFOLD)):
            codec.set_debug(dbg)
            a = codec.decode(tok).cpu().numpy()
            rec[name] = {"snr_db": O.snr_db(g["audio"], a), "max_abs": float(np.abs(g["audio"] - a).max()),
                         "ref_peak": float(np.abs(g["audio"]).max())}
    else:
        chunks = list(g["chunks"])
        cache, pos, outs = {}, 0, []
        for i, lc in enumerate(chunks):
            a, cache = codec.decode_one_token(tok[:, :, pos:pos + lc], cache, i == len(chunks) - 1)
            outs.append(a.cpu().numpy())
            pos += lc
        a = np.concatenate(outs, axis=1)
        ref = np.concatenate([g[f"audio_{i}"] for i in range(len(chunks))], axis=1)
        rec["streaming_token_step"] = {"snr_db": O.snr_db(ref, a), "max_abs": float(np.abs(ref - a).max()),
                                       "ref_peak": float(np.abs(ref).max())}
    out[case["name"]] = rec
print(json.dumps(out))
