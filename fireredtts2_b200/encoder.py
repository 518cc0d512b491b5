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
from typing import Any, Dict, List, Optional

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
    # the two feature encoders (0 layers = not configured: only ``features`` / ``encode_features`` are available)
    ssl_enc_layers: int = 0       # PretrainedWhisperEncoder: 32 layers, 20 heads, ffn 5120 at width ssl_in_dim (whisper.py:359-369)
    ssl_enc_heads: int = 0
    ssl_enc_ffn_dim: int = 0
    aco_layers: int = 0           # WhisperAcousticEncoder defaults: 12 layers, 8 heads (whisper.py:398-401)
    aco_heads: int = 0
    aco_ffn_dim: int = 0
    num_mels: int = 128           # whisper.py:372,391
    max_positions: int = 1500     # whisper.py:366,403

    @property
    def has_front(self) -> bool:
        return self.ssl_enc_layers > 0 or self.aco_layers > 0

    @property
    def samples_per_token(self) -> int:   # 16 kHz samples per RVQ frame: hop 160 * conv stride 2 * pooler (model.py:301)
        return 160 * 2 * self.avg_pooler

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
        ae = codec.get("acoustic_encoder", {})
        front = {}
        if codec.get("with_feature_encoders", False):
            # PretrainedWhisperEncoder.from_pretrained hard-codes whisper-large-v3 (whisper.py:359-369)
            front = dict(ssl_enc_layers=32, ssl_enc_heads=20, ssl_enc_ffn_dim=5120, aco_layers=ae.get("num_layers", 12),
                         aco_heads=ae.get("num_heads", 8), aco_ffn_dim=ae.get("ffn_dim") or 0,
                         num_mels=ae.get("num_mels", 128), max_positions=ae.get("max_positions", 1500))
        return cls(ssl_in_dim=a["in_dim"], ssl_embed_dim=a["embed_dim"], ssl_out_dim=a["out_dim"],
                   ssl_num_layers=a["num_layers"], ssl_num_heads=a["num_heads"], ssl_ffn_dim=a.get("ffn_dim") or 0,
                   aco_dim=aco, avg_pooler=d.get("avg_pooler", 4), **front)

    @classmethod
    def from_reference_module(cls, ref) -> "EncoderConfig":
        """Read the encode-side architecture off a live reference ``RedCodec`` / ``RedCodecInfer`` (model.py:150-170)."""
        a, d, ae, ssl = ref.ssl_adaptor, ref.downsample, ref.acoustic_encoder, ref.ssl

        def ffn(stack, width):
            f = stack.layers[0].fc1.out_features if len(stack.layers) else 0
            return 0 if f == 4 * width else f

        def heads(stack):
            return stack.layers[0].self_attn.num_heads if len(stack.layers) else 1

        return cls(ssl_in_dim=a.in_dim, ssl_embed_dim=a.embed_dim, ssl_out_dim=a.out_proj.out_features,
                   ssl_num_layers=len(a.layers), ssl_num_heads=heads(a), ssl_ffn_dim=ffn(a, a.embed_dim),
                   aco_dim=ae.embed_dim, avg_pooler=d.avg_pooler,
                   ssl_enc_layers=len(ssl.layers), ssl_enc_heads=heads(ssl), ssl_enc_ffn_dim=ffn(ssl, ssl.embed_dim),
                   aco_layers=len(ae.layers), aco_heads=heads(ae), aco_ffn_dim=ffn(ae, ae.embed_dim),
                   num_mels=ae.in_dim, max_positions=ae.max_positions)

    def to_reference_dict(self) -> Dict[str, Any]:
        return {"ssl_adaptor": dict(in_dim=self.ssl_in_dim, embed_dim=self.ssl_embed_dim, out_dim=self.ssl_out_dim,
                                    num_layers=self.ssl_num_layers, num_heads=self.ssl_num_heads,
                                    ffn_dim=self.ssl_ffn_dim or None),
                "downsample": dict(embed_dim=self.down_dim, avg_pooler=self.avg_pooler),
                "acoustic_encoder": dict(num_mels=self.num_mels, embed_dim=self.aco_dim, num_layers=self.aco_layers,
                                         num_heads=self.aco_heads or 1, ffn_dim=self.aco_ffn_dim or None,
                                         max_positions=self.max_positions),
                "ssl": dict(in_dim=self.num_mels, embed_dim=self.ssl_in_dim, num_layers=self.ssl_enc_layers,
                            num_heads=self.ssl_enc_heads or 1, ffn_dim=self.ssl_enc_ffn_dim or None,
                            max_positions=self.max_positions)}


# EC0: the encode side that fits the canonical decode config C0 (rvq.input_dim = 1024 = 256 semantic + 768 acoustic)
EC0 = EncoderConfig()
ETINY = EncoderConfig(ssl_in_dim=128, ssl_embed_dim=128, ssl_out_dim=64, ssl_num_layers=2, ssl_num_heads=2, aco_dim=64)
ESMALL = EncoderConfig(ssl_in_dim=256, ssl_embed_dim=256, ssl_out_dim=128, ssl_num_layers=3, ssl_num_heads=4, aco_dim=128)
# with the feature encoders: EC0F = whisper-large-v3 SSL encoder + the default acoustic encoder (head_dim 96, padded)
EC0F = dataclasses.replace(EC0, ssl_enc_layers=32, ssl_enc_heads=20, ssl_enc_ffn_dim=5120, aco_layers=12, aco_heads=8)
ETINYF = dataclasses.replace(ETINY, ssl_enc_layers=2, ssl_enc_heads=2, aco_layers=2, aco_heads=1, num_mels=64,
                             max_positions=300)
# acoustic width 192 with 2 heads: head_dim 96, zero-padded to 128 at load like the 768 / 8 default
EPADF = EncoderConfig(ssl_in_dim=256, ssl_embed_dim=256, ssl_out_dim=64, ssl_num_layers=2, ssl_num_heads=4, aco_dim=192,
                      ssl_enc_layers=2, ssl_enc_heads=4, aco_layers=2, aco_heads=2, num_mels=128, max_positions=300)
ENC_PRESETS = {"EC0": EC0, "ETINY": ETINY, "ESMALL": ESMALL, "EC0F": EC0F, "ETINYF": ETINYF, "EPADF": EPADF}


def _layer_keys(t: str) -> List[str]:
    return [t + "self_attn.k_proj.weight", t + "self_attn.v_proj.weight", t + "self_attn.v_proj.bias",
            t + "self_attn.q_proj.weight", t + "self_attn.q_proj.bias",
            t + "self_attn.out_proj.weight", t + "self_attn.out_proj.bias",
            t + "self_attn_layer_norm.weight", t + "self_attn_layer_norm.bias",
            t + "fc1.weight", t + "fc1.bias", t + "fc2.weight", t + "fc2.bias",
            t + "final_layer_norm.weight", t + "final_layer_norm.bias"]


def front_keys(cfg: EncoderConfig) -> List[str]:
    """Keys of the two ``WhisperEncoder`` modules (``ssl.*``, ``acoustic_encoder.*``; whisper.py:206-226)."""
    names: List[str] = []
    for p, n in (("ssl.", cfg.ssl_enc_layers), ("acoustic_encoder.", cfg.aco_layers)):
        names += [p + "conv1.weight", p + "conv1.bias", p + "conv2.weight", p + "conv2.bias", p + "embed_positions.weight"]
        for i in range(n):
            names += _layer_keys(f"{p}layers.{i}.")
        names += [p + "layer_norm.weight", p + "layer_norm.bias"]
    return names


def encoder_keys(cfg: EncoderConfig) -> List[str]:
    """Reference state_dict keys this stage consumes (``RedCodec`` naming, model.py:163-170)."""
    names = ["ssl_adaptor.in_proj.weight", "ssl_adaptor.in_proj.bias"]
    for i in range(cfg.ssl_num_layers):
        names += _layer_keys(f"ssl_adaptor.layers.{i}.")
    names += ["ssl_adaptor.layer_norm.weight", "ssl_adaptor.layer_norm.bias",
              "ssl_adaptor.out_proj.weight", "ssl_adaptor.out_proj.bias",
              "downsample.gate_proj.weight", "downsample.up_proj.weight", "downsample.down_proj.weight",
              "downsample.layer_norm.weight", "downsample.layer_norm.bias",
              "downsample.out_proj.weight", "downsample.out_proj.bias"]
    if cfg.has_front:
        names += front_keys(cfg)
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


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """The fixed positional table of ``WhisperEncoder`` (whisper.py:11-20,226), as float32."""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2, dtype=np.float32)).astype(np.float32)
    st = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(st), np.cos(st)], axis=1).astype(np.float32)


def synthetic_front_state_dict(cfg: EncoderConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded stand-ins for ``ssl.*`` and ``acoustic_encoder.*`` (own generator: the tensors of
    ``synthetic_encoder_state_dict`` and the goldens made from them do not change)."""
    rng = np.random.default_rng(seed + 90001)
    sd: Dict[str, np.ndarray] = {}

    def lin(name, out_f, in_f, bias=True):
        sd[name + ".weight"] = (rng.standard_normal((out_f, in_f), dtype=np.float32) / np.float32(np.sqrt(in_f)))
        if bias:
            sd[name + ".bias"] = (rng.standard_normal(out_f, dtype=np.float32) * np.float32(0.05))

    def lnp(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[name + ".bias"] = (0.05 * rng.standard_normal(c)).astype(np.float32)

    for p, E, n, F in (("ssl.", cfg.ssl_in_dim, cfg.ssl_enc_layers, cfg.ssl_enc_ffn_dim or 4 * cfg.ssl_in_dim),
                       ("acoustic_encoder.", cfg.aco_dim, cfg.aco_layers, cfg.aco_ffn_dim or 4 * cfg.aco_dim)):
        sd[p + "conv1.weight"] = (rng.standard_normal((E, cfg.num_mels, 3), dtype=np.float32) /
                                  np.float32(np.sqrt(3 * cfg.num_mels)))
        sd[p + "conv1.bias"] = rng.standard_normal(E, dtype=np.float32) * np.float32(0.05)
        sd[p + "conv2.weight"] = rng.standard_normal((E, E, 3), dtype=np.float32) / np.float32(np.sqrt(3 * E))
        sd[p + "conv2.bias"] = rng.standard_normal(E, dtype=np.float32) * np.float32(0.05)
        sd[p + "embed_positions.weight"] = sinusoids(cfg.max_positions, E)
        for i in range(n):
            t = f"{p}layers.{i}."
            lin(t + "self_attn.q_proj", E, E)
            lin(t + "self_attn.k_proj", E, E, bias=False)
            lin(t + "self_attn.v_proj", E, E)
            lin(t + "self_attn.out_proj", E, E)
            lnp(t + "self_attn_layer_norm", E)
            lin(t + "fc1", F, E)
            lin(t + "fc2", E, F)
            lnp(t + "final_layer_norm", E)
        lnp(p + "layer_norm", E)
    return sd


def synthetic_audio(batch: int, samples: int, seed: int = 0) -> np.ndarray:
    """Seeded 16 kHz test signal (B, n): a few drifting tones over noise, so the log-mel has structure and dynamic range."""
    rng = np.random.default_rng(seed)
    t = np.arange(samples, dtype=np.float64) / 16000.0
    out = np.empty((batch, samples), dtype=np.float32)
    for b in range(batch):
        x = 0.02 * rng.standard_normal(samples)
        for _ in range(4):
            f0, f1, a = rng.uniform(80, 3500), rng.uniform(-400, 400), rng.uniform(0.05, 0.25)
            x += a * np.sin(2 * np.pi * (f0 * t + 0.5 * f1 * t * t) + rng.uniform(0, 6.28))
        out[b] = x.astype(np.float32)
    return out


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
                            cfg.ssl_ffn_dim, cfg.aco_dim, cfg.avg_pooler, cfg.ssl_enc_layers, cfg.ssl_enc_heads,
                            cfg.ssl_enc_ffn_dim, cfg.aco_layers, cfg.aco_heads, cfg.aco_ffn_dim, cfg.num_mels,
                            cfg.max_positions)
        N.check(self._lib.frt2_enc_create(C.byref(c), self.device_index, C.byref(self._e)))
        sd = normalise_state_dict(state_dict)
        for key in encoder_keys(cfg):      # includes ssl.* / acoustic_encoder.* when the front end is configured
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

    def audio_features(self, audio16k: torch.Tensor, taps: bool = False):
        """audio16k ``(B, n)`` fp32, n a multiple of 1280 -> ``vq_in_feats`` ``(B, n // 1280, down_dim)``: the log-mel front
        end, both Whisper encoders, ssl_adaptor, cat, downsample (reference model.py:218-232).  ``taps=True`` also returns
        ``{"mel", "ssl", "aco"}`` = the outputs of the feature extractor and the two encoders (parity hooks)."""
        cfg = self.cfg
        if not cfg.has_front:
            raise ValueError("this encoder was configured without the feature encoders (ssl.*, acoustic_encoder.*)")
        if audio16k.dim() != 2:
            raise ValueError(f"audio16k must be (B, n), got {tuple(audio16k.shape)}")
        dev = torch.device("cuda", self.device_index)
        audio16k = audio16k.to(device=dev, dtype=torch.float32).contiguous()
        B, n = audio16k.shape
        if n == 0 or n % cfg.samples_per_token:
            raise ValueError(f"n={n} is not a positive multiple of {cfg.samples_per_token} (pad chunks as the reference "
                             "does, model.py:238-242)")
        L, T = n // cfg.samples_per_token, n // 320
        with torch.cuda.device(self.device_index):
            out = torch.empty((B, L, cfg.down_dim), dtype=torch.float32, device=dev)
            tp = {}
            if taps:
                tp = {"mel": torch.empty((B, n // 160, cfg.num_mels), dtype=torch.float32, device=dev),
                      "ssl": torch.empty((B, T, cfg.ssl_in_dim), dtype=torch.float32, device=dev),
                      "aco": torch.empty((B, T, cfg.aco_dim), dtype=torch.float32, device=dev)}
            ptr = lambda k: C.c_void_p(tp[k].data_ptr()) if taps else None
            cnt = C.c_int64(0)
            pitch = audio16k.stride(0) if B > 1 else n      # the stride of a size-1 dimension is arbitrary
            N.check(self._lib.frt2_enc_audio_features(self._e, C.c_void_p(audio16k.data_ptr()), pitch, B, n,
                                                      C.c_void_p(out.data_ptr()), ptr("mel"), ptr("ssl"), ptr("aco"),
                                                      C.byref(cnt),
                                                      C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)))
            self.last_launches = int(cnt.value)
        return (out, tp) if taps else out

    def encode(self, audio16k: torch.Tensor, audio16k_length: Optional[torch.Tensor], codec, batch_size: int = 96):
        """``RedCodecInfer.encode`` (reference model.py:243-305): every item is cut to its length, zero-padded to whole
        6 s chunks, all chunks are encoded in batches of ``batch_size`` and each item's chunk tokens are concatenated
        again -> ``(token (B, nq, L_max) int64, token_length (B,))`` with ``token_length = ceil(length / 1280)``."""
        if audio16k_length is None:
            assert audio16k.shape[0] == 1
            audio16k_length = torch.tensor([audio16k.shape[1]], dtype=torch.long, device=audio16k.device)
        CHUNK = 6 * 16000
        B = audio16k.shape[0]
        lens = [int(v) for v in audio16k_length.tolist()]
        chunks, counts = [], []
        for i in range(B):
            a = audio16k[i, :lens[i]]
            n_chunks = max(1, -(-lens[i] // CHUNK)) if lens[i] > 0 else 0
            a = torch.nn.functional.pad(a, (0, n_chunks * CHUNK - lens[i]))
            chunks += list(a.reshape(n_chunks, CHUNK)) if n_chunks else []
            counts.append(n_chunks)
        batch = torch.stack(chunks, dim=0)
        toks = []
        for i in range(0, batch.shape[0], batch_size):
            vq_in = self.audio_features(batch[i:i + batch_size])
            toks.append(codec.rvq_encode_codes(vq_in.transpose(1, 2)).permute(1, 0, 2))      # (b, nq, 75)
        toks = torch.cat(toks, dim=0)
        per_item = torch.split(toks, counts, dim=0)
        per_item = [t.permute(1, 0, 2).reshape(t.shape[1], -1) for t in per_item]              # (nq, n_chunks * 75)
        token_length = torch.tensor([-(-n // 1280) for n in lens], dtype=torch.long, device=toks.device)
        Lmax = int(token_length.max())
        out = torch.zeros((B, toks.shape[1], max(Lmax, max(t.shape[1] for t in per_item))), dtype=torch.int64, device=toks.device)
        for i, t in enumerate(per_item):
            out[i, :, :t.shape[1]] = t
        return out[..., :Lmax], token_length

    def encode_features(self, ssl: torch.Tensor, aco: torch.Tensor, codec, nq: Optional[int] = None) -> torch.Tensor:
        """-> indices ``(B, nq, L)`` int64 as ``_encode_one_batch`` returns them (model.py:233-236); ``codec`` is the
        ``RedCodecB200`` that holds the RVQ (its checkpoint must include the encode-side RVQ tensors)."""
        vq_in = self.features(ssl, aco)                                  # (B, L, D) time-major
        codes = codec.rvq_encode_codes(vq_in.transpose(1, 2), nq)        # (nq, B, L); strided view, no copy
        return codes.permute(1, 0, 2)
