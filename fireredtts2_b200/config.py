"""Codec-decoder configuration.

Mirrors the three sub-dicts of the reference's ``config_codec.json["codec"]`` that the decode
path consumes (``rvq``, ``upsample``, ``acoustic_decoder`` — reference
``fireredtts2/codec/model.py:173-194``).  The same field order is mirrored by ``frt2_config`` in
``include/frt2.h``.
"""
from __future__ import annotations

import dataclasses
import json
from typing import Any, Dict

TOKEN_RATE_HZ = 12.5          # reference model.py:301 (16 kHz / 1280)
SAMPLE_RATE = 24000           # reference fireredtts2.py:390
UPCONV_STRIDE = 4             # reference model.py:127
DECODER_UPSAMPLE = 2          # reference decoder.py:572-587
ATTN_BLOCK = 8                # reference decoder.py:266 (block-causal chunk of 8 frames)


@dataclasses.dataclass(frozen=True)
class CodecConfig:
    # rvq (reference rvq.py:93-130)
    rvq_dim: int = 512
    output_dim: int = 1024
    num_quantizers: int = 16
    codebook_size: int = 2048
    codebook_dim: int = 256
    # upsample / acoustic_decoder (reference model.py:123-140, decoder.py:550-597)
    embed_dim: int = 1024
    num_layers: int = 12
    num_heads: int = 16
    hop_length: int = 240
    upconv_stride: int = UPCONV_STRIDE

    def __post_init__(self):
        if self.output_dim != self.embed_dim:
            raise ValueError("rvq.output_dim must equal upsample/acoustic_decoder embed_dim "
                             "(the tensor flows unprojected, reference model.py:316-323)")
        if self.embed_dim % self.num_heads:
            raise ValueError("embed_dim must be divisible by num_heads")
        if self.upconv_stride != UPCONV_STRIDE:
            raise ValueError("only the reference's UpConv stride 4 is supported")

    # ---- derived ----
    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def n_fft(self) -> int:
        return 4 * self.hop_length

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def frames_per_token(self) -> int:
        return self.upconv_stride * DECODER_UPSAMPLE

    @property
    def samples_per_token(self) -> int:
        return self.frames_per_token * self.hop_length

    @property
    def istft_pad(self) -> int:
        return (self.n_fft - self.hop_length) // 2

    @property
    def has_out_project(self) -> bool:
        return self.codebook_dim != self.rvq_dim       # reference rvq.py:35-41

    @property
    def has_output_proj(self) -> bool:
        return self.rvq_dim != self.output_dim         # reference rvq.py:115-119

    # ---- reference-JSON interop ----
    @classmethod
    def from_reference_dict(cls, codec: Dict[str, Any]) -> "CodecConfig":
        """Accept ``json.load(config_codec.json)["codec"]`` (or the whole file) verbatim."""
        if "codec" in codec:
            codec = codec["codec"]
        rvq, up, ad = codec["rvq"], codec["upsample"], codec["acoustic_decoder"]
        if not ad.get("causal", False):
            raise AssertionError("Only AcousticDecoder with causal=True is supported "
                                 "(reference decoder.py:675-677)")
        if up["embed_dim"] != ad["embed_dim"]:
            raise ValueError("upsample.embed_dim != acoustic_decoder.embed_dim")
        return cls(
            rvq_dim=rvq.get("rvq_dim") or rvq["input_dim"],
            output_dim=rvq.get("output_dim") or rvq.get("rvq_dim") or rvq["input_dim"],
            num_quantizers=rvq.get("num_quantizers", 8),
            codebook_size=rvq.get("codebook_size", 1024),
            codebook_dim=rvq.get("codebook_dim", 256),
            embed_dim=ad["embed_dim"],
            num_layers=ad["num_layers"],
            num_heads=ad["num_heads"],
            hop_length=ad.get("hop_length", 240),
            upconv_stride=up.get("stride", 4),
        )

    @classmethod
    def from_json(cls, path: str) -> "CodecConfig":
        with open(path, "r") as f:
            return cls.from_reference_dict(json.load(f))

    def to_reference_dict(self) -> Dict[str, Any]:
        return {
            "rvq": dict(input_dim=self.embed_dim, rvq_dim=self.rvq_dim, output_dim=self.output_dim,
                        num_quantizers=self.num_quantizers, codebook_size=self.codebook_size,
                        codebook_dim=self.codebook_dim),
            "upsample": dict(embed_dim=self.embed_dim, stride=self.upconv_stride),
            "acoustic_decoder": dict(embed_dim=self.embed_dim, num_layers=self.num_layers,
                                     num_heads=self.num_heads, hop_length=self.hop_length, causal=True),
        }


# Canonical benchmark config and its identity-projection variant.
C0 = CodecConfig()
C1 = dataclasses.replace(C0, rvq_dim=256, codebook_dim=256)
# Small configs for fast CPU/GPU parity tests (same architecture, reduced widths).
TINY = CodecConfig(rvq_dim=64, output_dim=128, num_quantizers=4, codebook_size=64, codebook_dim=32,
                   embed_dim=128, num_layers=2, num_heads=2)
TINY_IDENT = dataclasses.replace(TINY, rvq_dim=64, codebook_dim=64)
SMALL = CodecConfig(rvq_dim=128, output_dim=256, num_quantizers=8, codebook_size=256, codebook_dim=64,
                    embed_dim=256, num_layers=3, num_heads=4)

MICRO = CodecConfig(rvq_dim=64, output_dim=64, num_quantizers=2, codebook_size=16, codebook_dim=16,
                    embed_dim=64, num_layers=1, num_heads=1)

# C0 widths with 4 layers: the adversarial-weights parity fixtures (tests/golden/adv4_*)
ADV4 = dataclasses.replace(C0, num_layers=4)

PRESETS = {"ADV4": ADV4, "MICRO": MICRO, "C0": C0, "C1": C1, "TINY": TINY, "TINY_IDENT": TINY_IDENT, "SMALL": SMALL}
