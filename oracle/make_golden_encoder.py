"""Golden vectors for the codec ENCODE side behind the feature encoders, from the REAL reference (build container only).

    python oracle/make_golden_encoder.py     # writes tests/golden/enc_*.npz

Builds the reference's own ``SslAdaptor`` / ``ResidualDownConv`` / ``ResidualVQ`` (/root/reference/fireredtts2/codec,
unmodified), loads the numpy-seeded synthetic weights with ``load_state_dict`` and records, for seeded feature tensors,
what ``RedCodecInfer._encode_one_batch`` (model.py:225-236) computes after the two Whisper encoders: ``sem_feats``,
``vq_in_feats`` and the RVQ indices, plus the reference's top-2 margin of every arg-max decision.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("FRT2_REFERENCE", "/root/reference"))

from fireredtts2.codec.model import ResidualDownConv, SslAdaptor  # noqa: E402  (reference)

from fireredtts2_b200.config import PRESETS  # noqa: E402
from fireredtts2_b200.encoder import ENC_PRESETS, synthetic_encoder_state_dict, synthetic_features  # noqa: E402
from oracle.make_golden_rvq_encode import build as build_rvq, margins  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name, encoder preset, codec preset (its rvq sees input_dim = down_dim), B, T (50 Hz frames), weight seed, data seed
CASES = [
    ("enc_etiny", "ETINY", "TINY", 2, 48, 3, 11),
    ("enc_esmall", "ESMALL", "SMALL", 3, 300, 5, 12),      # T = 300: the reference's 6 s chunk (model.py:262)
    ("enc_ec0", "EC0", "C0", 1, 100, 0, 13),
]


def build_modules(ecfg, wseed):
    d = ecfg.to_reference_dict()
    sd = synthetic_encoder_state_dict(ecfg, wseed)
    ada = SslAdaptor(**d["ssl_adaptor"]).eval()
    down = ResidualDownConv(**d["downsample"]).eval()
    ada.load_state_dict({k[len("ssl_adaptor."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("ssl_adaptor.")})
    down.load_state_dict({k[len("downsample."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("downsample.")})
    return ada, down, sd


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name, en, cn, B, T, wseed, dseed in CASES:
        ecfg, cfg = ENC_PRESETS[en], PRESETS[cn]
        ada, down, _ = build_modules(ecfg, wseed)
        rvq = build_rvq(cfg, wseed, ecfg.down_dim)
        ssl, aco = synthetic_features(ecfg, B, T, dseed)
        length = torch.full((B,), T, dtype=torch.long)
        with torch.inference_mode():
            sem, sem_len = ada(torch.from_numpy(ssl), length)                      # model.py:225
            vq_in, vq_len = down(torch.cat([sem, torch.from_numpy(aco)], dim=2), length)   # model.py:230-231
            codes = rvq.encode_codes(vq_in.transpose(1, 2)).permute(1, 0, 2)        # model.py:233-234
        mg = margins(rvq, vq_in.transpose(1, 2).contiguous())
        assert int(vq_len[0]) == T // ecfg.avg_pooler
        out = os.path.join(GOLDEN, name + ".npz")
        # the inputs are not stored: tests regenerate them from the data seed (synthetic_features)
        np.savez_compressed(out, sem=sem.numpy(), vq_in=vq_in.numpy(),
                            codes=codes.numpy().astype(np.int32), margins=mg.astype(np.float32),
                            meta=np.array([B, T, wseed, dseed]))
        print(name, "vq_in", tuple(vq_in.shape), "codes", tuple(codes.shape), "rms", float(vq_in.pow(2).mean().sqrt()),
              os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
