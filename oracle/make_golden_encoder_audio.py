"""Golden vectors for the WHOLE codec encode path from the waveform, from the REAL reference (build container only).

    python oracle/make_golden_encoder_audio.py     # writes tests/golden/encaudio_*.npz

Builds the reference's own ``PretrainedWhisperEncoder`` (at a reduced width; the class constructor, not
``from_pretrained``, which hard-codes whisper-large-v3), ``WhisperAcousticEncoder``, ``SslAdaptor``, ``ResidualDownConv``
and ``ResidualVQ`` (/root/reference/fireredtts2/codec, unmodified), loads the numpy-seeded synthetic weights with
``load_state_dict`` and records ``RedCodecInfer._encode_one_batch`` (model.py:218-236) on seeded audio: the log-mel
features, both encoder outputs, ``vq_in_feats`` and the indices with the reference's top-2 margins.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("FRT2_REFERENCE", "/root/reference"))

from fireredtts2.codec.model import RedCodecInfer, ResidualDownConv, SslAdaptor  # noqa: E402  (reference)
from fireredtts2.codec.whisper import PretrainedWhisperEncoder, WhisperAcousticEncoder, WhisperMelExtractor  # noqa: E402

from fireredtts2_b200.config import PRESETS  # noqa: E402
from fireredtts2_b200.encoder import (ENC_PRESETS, synthetic_audio, synthetic_encoder_state_dict,  # noqa: E402
                                      synthetic_front_state_dict)
from oracle.make_golden_rvq_encode import build as build_rvq, margins  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name, encoder preset, codec preset, B, samples (multiple of 1280), weight seed, data seed
CASES = [
    ("encaudio_etinyf", "ETINYF", "TINY", 2, 1280 * 12, 2, 31),
    ("encaudio_epadf", "EPADF", "SMALL", 2, 1280 * 30, 4, 32),
]


class RefEncodeOnly(RedCodecInfer):
    """RedCodecInfer with only the encode-side sub-modules (its own __init__ wants a full RedCodec)."""

    def __init__(self, ecfg, cfg, wseed):
        nn.Module.__init__(self)
        d = ecfg.to_reference_dict()
        self.ssl = PretrainedWhisperEncoder(**{k: v for k, v in d["ssl"].items()})
        self.ssl.feature_extractor = WhisperMelExtractor(num_mels=ecfg.num_mels)      # whisper.py:371-378
        self.ssl_adaptor = SslAdaptor(**d["ssl_adaptor"])
        self.acoustic_encoder = WhisperAcousticEncoder(**d["acoustic_encoder"])
        self.downsample = ResidualDownConv(**d["downsample"])
        self.rvq = build_rvq(cfg, wseed, ecfg.down_dim)
        sd = dict(synthetic_encoder_state_dict(ecfg, wseed))
        sd.update(synthetic_front_state_dict(ecfg, wseed))
        for name in ("ssl", "ssl_adaptor", "acoustic_encoder", "downsample"):
            sub = {k[len(name) + 1:]: torch.from_numpy(v) for k, v in sd.items() if k.startswith(name + ".")}
            getattr(self, name).load_state_dict(sub)
        self.eval()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name, en, cn, B, n, wseed, dseed in CASES:
        ecfg, cfg = ENC_PRESETS[en], PRESETS[cn]
        m = RefEncodeOnly(ecfg, cfg, wseed)
        audio = synthetic_audio(B, n, dseed)
        a = torch.from_numpy(audio)
        length = torch.full((B,), n, dtype=torch.long)
        with torch.inference_mode():
            codes = m._encode_one_batch(a)                                          # (B, nq, L)  model.py:218-236
            mel, _ = m.ssl.feature_extractor(a, length)                             # the intermediates, same modules
            ssl, ssl_len = m.ssl.forward(a, length)
            aco, aco_len = m.acoustic_encoder(a, length)
            sem, _ = m.ssl_adaptor(ssl, ssl_len)
            vq_in, _ = m.downsample(torch.cat([sem, aco], dim=2), aco_len)
            again = m.rvq.encode_codes(vq_in.transpose(1, 2)).permute(1, 0, 2)
        assert torch.equal(codes, again)
        mg = margins(m.rvq, vq_in.transpose(1, 2).contiguous())
        out = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(out, mel=mel.numpy().astype(np.float32), ssl=ssl.numpy(), aco=aco.numpy(), vq_in=vq_in.numpy(),
                            codes=codes.numpy().astype(np.int32), margins=mg.astype(np.float32),
                            meta=np.array([B, n, wseed, dseed]))
        print(name, "mel", tuple(mel.shape), "ssl", tuple(ssl.shape), "vq_in", tuple(vq_in.shape), "codes", tuple(codes.shape),
              "mel range", float(mel.min()), float(mel.max()), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
