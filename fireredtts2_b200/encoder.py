"""Codec ENCODE side behind the two feature encoders.

``CodecEncoderB200`` runs what ``RedCodecInfer._encode_one_batch`` (reference ``fireredtts2/codec/model.py:218-236``)
does after the Whisper encoders: ``SslAdaptor`` (model.py:19-77), ``torch.cat([sem, aco], dim=2)`` (model.py:230),
``ResidualDownConv`` (model.py:80-121) and — through the codec handle — ``ResidualVQ.encode_codes`` (rvq.py:128-143).
All compute runs in libfrt2_b200.so (``frt2_enc_*`` in include/frt2.h): tcgen05 GEMMs and attention, fp16 operands, fp32
accumulation.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any, Dict, Iterable, List, Optional

import numpy as np
import torch

from . import _native as N
from .weights import normalise_state_dict


@dataclasses.dataclass(frozen=True)
class EncoderConfig:
    """``config_codec.json["codec"]["ssl_adaptor"|"acoustic_encoder"|"downsample"]`` (reference model.py:176-180)."""
    ssl_in_dim: int = 1280        # PretrainedWhisperEncoder embed_dim (whisper.py:363)
    ssl_embed_dim: int = 768
    ssl_out_dim: int = 256
    ssl_num_layers: int = 4
    ssl_num_heads: int = 12
    ssl_ffn_dim: int = 0          # 0 = 4 * embed_dim (whisper.py:137)
    aco_dim: int = 768            # WhisperAcousticEncoder embed_dim default (whisper.py:389)
    avg_pooler: int = 4           # ResidualDownConv default (model.py:84)

    @property
    def ffn_dim(self) -> int:
        return self.ssl_ffn_dim or 4 * self.ssl_embed_dim

    @property
    def down_dim(self) -> int:    # downsample.embed_dim == rvq.input_dim
        return self.ssl_out_dim + self.aco_dim

    @classmethod
    def from_reference_dict(cls, codec: Dict[str, Any]) -> "EncoderConfig":
        if "codec" in codec:
            codec = codec["codec"]
        a, d = codec["ssl_adaptor"], codec["downsample"]
        aco = codec.get("acoustic_encoder", {}).get("embed_dim", 768)
        if d.get("embed_dim", 768) != a["out_dim"] + aco:
            raise ValueError("downsample.embed_dim must equal ssl_adaptor.out_dim + acoustic_encoder.embed_dim "
                             "(the two feature streams are concatenated, reference model.py:230)")
        return cls(ssl_in_dim=a["in_dim"], ssl_embed_dim=a["embed_dim"], ssl_out_dim=a["out_dim"],
                   ssl_num_layers=a["num_layers"], ssl_num_heads=a["num_heads"], ssl_ffn_dim=a.get("ffn_dim") or 0,
                   aco_dim=aco, avg_pooler=d.get("avg_pooler", 4))

    def to_reference_dict(self) -> Dict[str, Any]:
        return {"ssl_adaptor": dict(in_dim=self.ssl_in_dim, embed_dim=self.ssl_embed_dim, out_dim=self.ssl_out_dim,
                                    num_layers=self.ssl_num_layers, num_heads=self.ssl_num_heads,
                                    ffn_dim=self.ssl_ffn_dim or None),
                "downsample": dict(embed_dim=self.down_dim, avg_pooler=self.avg_pooler)}


# EC0: the encode side that fits the canonical decode config C0 (rvq.input_dim = 1024 = 256 semantic + 768 acoustic)
EC0 = EncoderConfig()
ETINY = EncoderConfig(ssl_in_dim=128, ssl_embed_dim=128, ssl_out_dim=64, ssl_num_layers=2, ssl_num_heads=2, aco_dim=64)
ESMALL = EncoderConfig(ssl_in_dim=256, ssl_embed_dim=256, ssl_out_dim=128, ssl_num_layers=3, ssl_num_heads=4, aco_dim=128)
ENC_PRESETS = {"EC0": EC0, "ETINY": ETINY, "ESMALL": ESMALL}


def encoder_keys(cfg: EncoderConfig) -> List[str]:
    """Reference state_dict keys this stage consumes (``RedCodec`` naming, model.py:163-170)."""
    names = ["ssl_adaptor.in_proj.weight", "ssl_adaptor.in_proj.bias"]
    for i in range(cfg.ssl_num_layers):
        t = f"ssl_adaptor.layers.{i}."
        names += [t + "self_attn.k_proj.weight", t + "self_attn.v_proj.weight", t + "self_attn.v_proj.bias",
                  t + "self_attn.q_proj.weight", t + "self_attn.q_proj.bias",
                  t + "self_attn.out_proj.weight", t + "self_attn.out_proj.bias",
                  t + "self_attn_layer_norm.weight", t + "self_attn_layer_norm.bias",
                  t + "fc1.weight", t + "fc1.bias", t + "fc2.weight", t + "fc2.bias",
                  t + "final_layer_norm.weight", t + "final_layer_norm.bias"]
    names += ["ssl_adaptor.layer_norm.weight", "ssl_adaptor.layer_norm.bias",
              "ssl_adaptor.out_proj.weight", "ssl_adaptor.out_proj.bias",
              "downsample.gate_proj.weight", "downsample.up_proj.weight", "downsample.down_proj.weight",
              "downsample.layer_norm.weight", "downsample.layer_norm.bias",
              "downsample.out_proj.weight", "downsample.out_proj.bias"]
    return names


def synthetic_encoder_state_dict(cfg: EncoderConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded stand-in for the encode-side checkpoint tensors (reference layouts and key names; normal(0, 0.02) weights
    like ``SslAdaptor._init_weights`` model.py:68-77, but non-trivial biases and LayerNorm parameters so that every
    term of the path is exercised)."""
    rng = np.random.default_rng(seed + 4243)
    E, F, D, P = cfg.ssl_embed_dim, cfg.ffn_dim, cfg.down_dim, cfg.avg_pooler * cfg.down_dim
    sd: Dict[str, np.ndarray] = {}

    def lin(name, out_f, in_f, bias=True, std=None):
        std = std if std is not None else 1.0 / np.sqrt(in_f)
        sd[name + ".weight"] = (rng.standard_normal((out_f, in_f)) * std).astype(np.float32)
        if bias:
            sd[name + ".bias"] = (rng.standard_normal(out_f) * 0.05).astype(np.float32)

    def lnp(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[name + ".bias"] = (0.05 * rng.standard_normal(c)).astype(np.float32)

    lin("ssl_adaptor.in_proj", E, cfg.ssl_in_dim)
    for i in range(cfg.ssl_num_layers):
        t = f"ssl_adaptor.layers.{i}."
        lin(t + "self_attn.q_proj", E, E)
        lin(t + "self_attn.k_proj", E, E, bias=False)
        lin(t + "self_attn.v_proj", E, E)
        lin(t + "self_attn.out_proj", E, E)
        lnp(t + "self_attn_layer_norm", E)
        lin(t + "fc1", F, E)
        lin(t + "fc2", E, F)
        lnp(t + "final_layer_norm", E)
    lnp("ssl_adaptor.layer_norm", E)
    lin("ssl_adaptor.out_proj", cfg.ssl_out_dim, E)
    for nm in ("gate_proj", "up_proj"):
        sd[f"downsample.{nm}.weight"] = (rng.standard_normal((P, D, cfg.avg_pooler)) / np.sqrt(P)).astype(np.float32)
    lin("downsample.down_proj", P, P, bias=False)
    lnp("downsample.layer_norm", P)
    lin("downsample.out_proj", D, P)
    return sd


def synthetic_features(cfg: EncoderConfig, batch: int, frames: int, seed: int = 0):
    """Seeded stand-ins for the outputs of the two feature encoders: ssl (B, T, ssl_in_dim), aco (B, T, aco_dim)."""
    rng = np.random.default_rng(seed)
    ssl = rng.standard_normal((batch, frames, cfg.ssl_in_dim)).astype(np.float32)
    aco = rng.standard_normal((batch, frames, cfg.aco_dim)).astype(np.float32)
    return ssl, aco


class CodecEncoderB200:
    """``features(ssl, aco)`` -> the input of the RVQ; ``encode_features(ssl, aco, codec)`` -> codes ``(B, nq, L)``."""

    def __init__(self, cfg: EncoderConfig, state_dict, device="cuda:0"):
        self._lib = N.load()
        if not torch.cuda.is_available():
            raise RuntimeError("CodecEncoderB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device)
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._e = C.c_void_p()
        c = N.Frt2EncConfig(cfg.ssl_in_dim, cfg.ssl_embed_dim, cfg.ssl_out_dim, cfg.ssl_num_layers, cfg.ssl_num_heads,
                            cfg.ssl_ffn_dim, cfg.aco_dim, cfg.avg_pooler)
        N.check(self._lib.frt2_enc_create(C.byref(c), self.device_index, C.byref(self._e)))
        sd = normalise_state_dict(state_dict)
        for key in encoder_keys(cfg):
            if key not in sd:
                raise KeyError(f"state_dict is missing encode-path tensor {key!r}")
            a = np.ascontiguousarray(sd[key], dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            N.check(self._lib.frt2_enc_load_tensor(self._e, key.encode(), a.ctypes.data_as(C.c_void_p), a.ndim, shape, 0))
        N.check(self._lib.frt2_enc_finalize(self._e))
        self.last_launches = 0

    def __del__(self):
        try:
            if self._e:
                self._lib.frt2_enc_destroy(self._e)
                self._e = C.c_void_p()
        except Exception:
            pass

    def features(self, ssl: torch.Tensor, aco: torch.Tensor) -> torch.Tensor:
        """ssl ``(B, T, ssl_in_dim)``, aco ``(B, T, aco_dim)`` fp32 -> ``vq_in_feats`` ``(B, T // pooler, down_dim)``
        (reference model.py:225-232: ssl_adaptor, cat, downsample)."""
        cfg = self.cfg
        if ssl.dim() != 3 or aco.dim() != 3 or ssl.shape[:2] != aco.shape[:2]:
            raise ValueError(f"ssl (B,T,{cfg.ssl_in_dim}) and aco (B,T,{cfg.aco_dim}) expected, got {tuple(ssl.shape)} "
                             f"and {tuple(aco.shape)}")
        if ssl.shape[2] != cfg.ssl_in_dim or aco.shape[2] != cfg.aco_dim:
            raise ValueError("feature widths do not match the encoder config")
        dev = torch.device("cuda", self.device_index)
        ssl = ssl.to(device=dev, dtype=torch.float32).contiguous()
        aco = aco.to(device=dev, dtype=torch.float32).contiguous()
        B, T, _ = ssl.shape
        if T % cfg.avg_pooler:
            raise RuntimeError(f"T={T} is not a multiple of avg_pooler={cfg.avg_pooler} (reference model.py:113 reshape)")
        with torch.cuda.device(self.device_index):
            out = torch.empty((B, T // cfg.avg_pooler, cfg.down_dim), dtype=torch.float32, device=dev)
            n = C.c_int64(0)
            N.check(self._lib.frt2_enc_features(self._e, C.c_void_p(ssl.data_ptr()), C.c_void_p(aco.data_ptr()), B, T,
                                                C.c_void_p(out.data_ptr()), C.byref(n),
                                                C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)))
            self.last_launches = int(n.value)
        return out

    def encode_features(self, ssl: torch.Tensor, aco: torch.Tensor, codec, nq: Optional[int] = None) -> torch.Tensor:
        """-> indices ``(B, nq, L)`` int64 as ``_encode_one_batch`` returns them (model.py:233-236); ``codec`` is the
        ``RedCodecB200`` that holds the RVQ (its checkpoint must include the encode-side RVQ tensors)."""
        vq_in = self.features(ssl, aco)                                  # (B, L, D) time-major
        codes = codec.rvq_encode_codes(vq_in.transpose(1, 2), nq)        # (nq, B, L); strided view, no copy
        return codes.permute(1, 0, 2)
