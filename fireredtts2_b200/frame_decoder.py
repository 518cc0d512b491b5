"""Frame decoder + sampler of the speech LM on the library (reference ``fireredtts2/llm/llm.py:304-330``).

``Model.generate_frame`` ends with the part that turns the backbone's last hidden state into the frame's
``audio_num_codebooks`` codes: ``codebook0_head`` + ``sample_topk`` for codebook 0, then fifteen dependent passes of the
small "decoder" transformer (torchtune ``qwen2``: RMSNorm, grouped-query attention with q/k/v bias and rotary positions,
SwiGLU) over positions 0..15 of a per-frame K/V state, each followed by ``audio_head[i-1]``, ``sample_topk(·, 10, 0.75)``
and the embedding lookup of the next input.  ``FrameDecoderB200.generate_codes`` runs all of that as ONE captured CUDA
graph per frame on ``libfrt2_b200`` (``csrc/frame_decoder.cu``): the codes never visit the host and can be handed to
``RedCodecB200.decode_one_token`` / ``StreamDecoder.push`` as a device tensor.  The backbone itself is out of scope.

There is no CPU fallback: without the library / a CUDA device the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as N


@dataclasses.dataclass(frozen=True)
class FrameDecoderConfig:
    """Widths of ``Model`` that the frame tail sees (reference llm.py:75-118, modules.py:5-82)."""
    backbone_dim: int = 1536          # FLAVORS[backbone_flavor] embed_dim (qwen-1.5b)
    dim: int = 1536                   # decoder flavor embed_dim (qwen-200m)
    num_layers: int = 4
    num_heads: int = 12
    num_kv_heads: int = 2
    intermediate_dim: int = 8960
    audio_vocab_size: int = 2048
    audio_num_codebooks: int = 16
    rope_base: float = 1000000.0
    norm_eps: float = 1e-6

    def __post_init__(self):
        if self.dim % self.num_heads or self.num_heads % self.num_kv_heads:
            raise ValueError("dim % num_heads and num_heads % num_kv_heads must be 0")
        hd = self.dim // self.num_heads
        if hd % 2 or hd > 128:
            raise ValueError("head_dim must be even and <= 128")
        if self.dim % 8 or self.backbone_dim % 8 or self.intermediate_dim % 8:
            raise ValueError("widths must be multiples of 8")
        if self.dim > 4096:
            raise ValueError("decoder dim > 4096 is not supported (fused RMSNorm row in shared memory)")

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads

    @property
    def qkv_dim(self) -> int:
        return (self.num_heads + 2 * self.num_kv_heads) * self.head_dim

    def weight_bytes_per_frame(self) -> int:
        """Algorithmic fp16 weight bytes of one frame (the HBM roofline of the frame tail at small batch): the reference runs
        ``audio_num_codebooks - 1`` decoder calls per frame (llm.py:318-322, the first over two positions), each streams the
        projection and every layer once; plus the codebook-0 head and one ``audio_head`` slice per call."""
        D, Db, I, V, n = self.dim, self.backbone_dim, self.intermediate_dim, self.audio_vocab_size, self.audio_num_codebooks
        layer = D * self.qkv_dim + D * D + 3 * D * I
        per_pass = Db * D + self.num_layers * layer
        return 2 * ((n - 1) * per_pass + (n - 1) * D * V + V * Db)


# decoder flavors of the reference (modules.py:5-35) next to the backbone width they are paired with
FD_200M = FrameDecoderConfig()                                                            # qwen-1.5b backbone + qwen-200m decoder
FD_500M = FrameDecoderConfig(backbone_dim=2048, dim=896, num_layers=24, num_heads=14, num_kv_heads=2,
                             intermediate_dim=4864)                                      # llm.py:356-357 (qwen-3b + qwen-500m)
FD_TINY = FrameDecoderConfig(backbone_dim=96, dim=64, num_layers=2, num_heads=4, num_kv_heads=2, intermediate_dim=160,
                             audio_vocab_size=64, audio_num_codebooks=6)
FD_SMALL = FrameDecoderConfig(backbone_dim=256, dim=384, num_layers=3, num_heads=6, num_kv_heads=2, intermediate_dim=1024,
                              audio_vocab_size=512, audio_num_codebooks=16)
FD_PRESETS = {"FD_200M": FD_200M, "FD_500M": FD_500M, "FD_TINY": FD_TINY, "FD_SMALL": FD_SMALL}


def layer_keys(i: int) -> List[str]:
    """torchtune ``TransformerSelfAttentionLayer`` parameter names under the reference's ``decoder.`` prefix."""
    p = f"decoder.layers.{i}."
    return [p + "sa_norm.scale", p + "attn.q_proj.weight", p + "attn.q_proj.bias", p + "attn.k_proj.weight",
            p + "attn.k_proj.bias", p + "attn.v_proj.weight", p + "attn.v_proj.bias", p + "attn.output_proj.weight",
            p + "mlp_norm.scale", p + "mlp.w1.weight", p + "mlp.w2.weight", p + "mlp.w3.weight"]


def frame_decoder_keys(cfg: FrameDecoderConfig) -> List[str]:
    keys = ["projection.weight", "audio_embeddings.weight", "codebook0_head.weight", "audio_head", "decoder.norm.scale"]
    for i in range(cfg.num_layers):
        keys += layer_keys(i)
    return keys


def synthetic_frame_decoder_state_dict(cfg: FrameDecoderConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random weights under the reference ``Model.state_dict()`` names (numpy-seeded: the same on every machine).  Scales
    are chosen so that the logits have a spread of a few units (top-k sampling then has real decisions to make) and the
    norm scales / biases are not trivial."""
    rng = np.random.default_rng(seed)
    D, Db, I, V, n = cfg.dim, cfg.backbone_dim, cfg.intermediate_dim, cfg.audio_vocab_size, cfg.audio_num_codebooks
    hd, H, Hk = cfg.head_dim, cfg.num_heads, cfg.num_kv_heads

    def lin(o, i, s=1.0):
        return (rng.standard_normal((o, i)) * (s / np.sqrt(i))).astype(np.float32)

    sd = {
        "projection.weight": lin(D, Db),
        "audio_embeddings.weight": rng.standard_normal((V * n, Db)).astype(np.float32),
        "codebook0_head.weight": lin(V, Db, 2.0),
        "audio_head": (rng.standard_normal((n - 1, D, V)) * (2.0 / np.sqrt(D))).astype(np.float32),
        "decoder.norm.scale": (1.0 + 0.2 * rng.standard_normal(D)).astype(np.float32),
    }
    for i in range(cfg.num_layers):
        p = f"decoder.layers.{i}."
        sd[p + "sa_norm.scale"] = (1.0 + 0.2 * rng.standard_normal(D)).astype(np.float32)
        sd[p + "attn.q_proj.weight"] = lin(H * hd, D, 1.5)
        sd[p + "attn.q_proj.bias"] = (0.3 * rng.standard_normal(H * hd)).astype(np.float32)
        sd[p + "attn.k_proj.weight"] = lin(Hk * hd, D, 1.5)
        sd[p + "attn.k_proj.bias"] = (0.3 * rng.standard_normal(Hk * hd)).astype(np.float32)
        sd[p + "attn.v_proj.weight"] = lin(Hk * hd, D)
        sd[p + "attn.v_proj.bias"] = (0.1 * rng.standard_normal(Hk * hd)).astype(np.float32)
        sd[p + "attn.output_proj.weight"] = lin(D, H * hd)
        sd[p + "mlp_norm.scale"] = (1.0 + 0.2 * rng.standard_normal(D)).astype(np.float32)
        sd[p + "mlp.w1.weight"] = lin(I, D)
        sd[p + "mlp.w2.weight"] = lin(D, I)
        sd[p + "mlp.w3.weight"] = lin(I, D)
    return sd


def adversarial_frame_decoder_state_dict(cfg: FrameDecoderConfig, seed: int = 0, outlier: float = 50.0) -> Dict[str, np.ndarray]:
    """Weights that stress fp16 operands the way trained LMs do: four residual channels carry values ``outlier`` times the
    others ("massive activations": the projection's rows of those channels are scaled), RMSNorm scales between 0.1 and 5,
    gate / up projections three times wider (SwiGLU products in the hundreds)."""
    sd = synthetic_frame_decoder_state_dict(cfg, seed)
    rng = np.random.default_rng(7000 + seed)
    ch = rng.choice(cfg.dim, size=4, replace=False)
    sd["projection.weight"][ch] *= outlier
    for k in list(sd):
        if k.endswith("norm.scale"):
            sd[k] = rng.uniform(0.1, 5.0, sd[k].shape).astype(np.float32)
        if k.endswith("mlp.w1.weight") or k.endswith("mlp.w3.weight"):
            sd[k] = (sd[k] * 3.0).astype(np.float32)
    return sd


def synthetic_frame_inputs(cfg: FrameDecoderConfig, batch: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """-> ``last_h (B, backbone_dim)`` fp32 and Exp(1) draws ``noise (B, ncb, V)`` (what ``exponential_(1)`` produces in
    ``_multinomial_sample_one_no_sync``, llm.py:34-36)."""
    rng = np.random.default_rng(1000 + seed)
    last_h = rng.standard_normal((batch, cfg.backbone_dim)).astype(np.float32)
    noise = rng.exponential(1.0, (batch, cfg.audio_num_codebooks, cfg.audio_vocab_size)).astype(np.float32)
    return last_h, np.maximum(noise, np.float32(1e-30))


class Frt2FdConfig(C.Structure):
    _fields_ = [("backbone_dim", C.c_int32), ("dim", C.c_int32), ("num_layers", C.c_int32), ("num_heads", C.c_int32),
                ("num_kv_heads", C.c_int32), ("intermediate_dim", C.c_int32), ("audio_vocab_size", C.c_int32),
                ("audio_num_codebooks", C.c_int32), ("rope_base", C.c_float), ("norm_eps", C.c_float),
                ("max_batch", C.c_int32)]


class FrameDecoderB200:
    """``generate_codes(last_h, ...)`` -> ``(B, audio_num_codebooks)`` int32 codes of one frame (llm.py:304-330)."""

    def __init__(self, cfg: FrameDecoderConfig, state_dict, device="cuda:0", max_batch: int = 8):
        """Frames of up to 16 items always run on the weight-streaming kernels; ``max_batch`` > 16 also keeps row-major weight
        copies so that frames of up to ``max_batch`` items (a pool of concurrent streams) run on the tcgen05 GEMM."""
        import torch
        self._lib = N.load()
        if not torch.cuda.is_available():
            raise RuntimeError("FrameDecoderB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device)
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = C.c_void_p()
        self.max_batch = max(16, int(max_batch))
        c = Frt2FdConfig(cfg.backbone_dim, cfg.dim, cfg.num_layers, cfg.num_heads, cfg.num_kv_heads, cfg.intermediate_dim,
                         cfg.audio_vocab_size, cfg.audio_num_codebooks, cfg.rope_base, cfg.norm_eps, int(max_batch))
        N.check(self._lib.frt2_fd_create(C.byref(c), self.device_index, C.byref(self._h)))
        for key in frame_decoder_keys(cfg):
            if key not in state_dict:
                raise KeyError(f"state_dict is missing frame-decoder tensor {key!r}")
            v = state_dict[key]
            if hasattr(v, "detach"):
                v = v.detach().to("cpu", dtype=torch.float32).numpy()
            a = np.ascontiguousarray(v, dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            N.check(self._lib.frt2_fd_load_tensor(self._h, key.encode(), a.ctypes.data_as(C.c_void_p), a.ndim, shape, 0))
        N.check(self._lib.frt2_fd_finalize(self._h))
        self.last_launches = 0

    def __del__(self):
        try:
            if self._h:
                self._lib.frt2_fd_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @staticmethod
    def config_from_reference(model) -> Tuple[FrameDecoderConfig, dict]:
        """Widths of a live reference ``Model`` (llm.py:85) read off its modules / state_dict (torchtune parameter names)."""
        sd = {k: v for k, v in model.state_dict().items()}
        n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))
        D = sd["decoder.norm.scale"].shape[0]
        hd_total = sd["decoder.layers.0.attn.q_proj.weight"].shape[0]
        kv_total = sd["decoder.layers.0.attn.k_proj.weight"].shape[0]
        H = int(model.decoder.layers[0].attn.num_heads)
        hd = hd_total // H
        cfg = FrameDecoderConfig(backbone_dim=sd["projection.weight"].shape[1], dim=D, num_layers=n_layers, num_heads=H,
                                 num_kv_heads=kv_total // hd, intermediate_dim=sd["decoder.layers.0.mlp.w1.weight"].shape[0],
                                 audio_vocab_size=model.config.audio_vocab_size,
                                 audio_num_codebooks=model.config.audio_num_codebooks)
        return cfg, sd

    @classmethod
    def from_reference(cls, model, device="cuda:0", max_batch: int = 8) -> "FrameDecoderB200":
        """Build from a live reference ``Model`` (llm.py:85): widths from its modules, tensors from its state_dict."""
        cfg, sd = cls.config_from_reference(model)
        return cls(cfg, sd, device, max_batch)

    def check_error(self):
        """Synchronise and raise IndexError if a given code was outside ``[0, audio_vocab_size)`` since the last check (the
        reference raises inside ``nn.Embedding``, llm.py:336-337)."""
        import torch
        N.check(self._lib.frt2_fd_check_error(self._h, C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)))

    def generate_codes(self, last_h, topk: int, temperature: float, c0=None, noise=None, seed: int = 0, forced=None,
                       return_logits: bool = False):
        """last_h ``(B, backbone_dim)`` fp32 = ``h[:, -1, :]`` of the backbone (llm.py:304).  ``c0`` ``(B,)`` int: codebook-0
        codes already sampled by the caller (else ``codebook0_head`` + ``sample_topk(topk, temperature)`` run here,
        llm.py:305-306).  ``noise`` ``(B, ncb, V)`` fp32: the Exp(1) draws ``q`` of ``_multinomial_sample_one_no_sync`` per
        codebook (parity tests); ``None`` -> the library's counter-based generator keyed by ``seed``.  ``forced``
        ``(B, ncb)`` int: teacher forcing — the returned codes are these, the logits are still computed from them.
        -> codes ``(B, ncb)`` int32 on the device [, logits ``(B, ncb, V)`` fp32]."""
        import torch
        cfg = self.cfg
        dev = torch.device("cuda", self.device_index)
        if last_h.dim() != 2 or last_h.shape[1] != cfg.backbone_dim:
            raise ValueError(f"last_h must be (B, {cfg.backbone_dim}), got {tuple(last_h.shape)}")
        B = last_h.shape[0]
        if not 1 <= B <= self.max_batch:
            raise ValueError(f"batch {B} outside 1..{self.max_batch} (max_batch of the constructor)")
        if topk < 1:
            raise ValueError("topk must be >= 1")
        last_h = last_h.to(device=dev, dtype=torch.float32).contiguous()
        ncb, V = cfg.audio_num_codebooks, cfg.audio_vocab_size

        def opt(t, shape, dtype):
            if t is None:
                return None
            t = t.to(device=dev, dtype=dtype).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"expected shape {shape}, got {tuple(t.shape)}")
            return t

        c0 = opt(c0, (B,), torch.int32)
        noise = opt(noise, (B, ncb, V), torch.float32)
        forced = opt(forced, (B, ncb), torch.int32)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        with torch.cuda.device(self.device_index):
            codes = torch.empty((B, ncb), dtype=torch.int32, device=dev)
            logits = torch.empty((B, ncb, V), dtype=torch.float32, device=dev) if return_logits else None
            cnt = C.c_int64(0)
            N.check(self._lib.frt2_fd_generate(self._h, C.c_void_p(last_h.data_ptr()), B, ptr(c0), ptr(noise),
                                               C.c_uint64(seed), int(topk), C.c_float(temperature), ptr(forced),
                                               C.c_void_p(codes.data_ptr()), ptr(logits), C.byref(cnt),
                                               C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)))
            self.last_launches = int(cnt.value)
        return (codes, logits) if return_logits else codes


class GenerateFrameB200:
    """Drop-in for the bound method ``Model.generate_frame`` (llm.py:274-330): the same five arguments, the same
    ``(batch_size, audio_num_codebooks)`` int32 result, so ``FireRedTTS2``'s frame loops (``fireredtts2.py:173-192,
    228-247``) run unchanged after ``GenerateFrameB200.install(model)``.

    Lines 292-302 — the causal-mask lookup, the masked sum of the text / audio token embeddings, the backbone and its K/V
    state — stay the reference's OWN modules, called exactly as the reference calls them (the backbone is outside SURVEY
    §8).  Everything behind ``last_h = h[:, -1, :]`` (llm.py:304-330) is one ``FrameDecoderB200.generate_codes`` call, i.e.
    one captured CUDA graph on the library; ``self.decoder`` of the reference is never run (its per-frame
    ``reset_caches()`` at llm.py:317 has nothing left to reset: the library owns the frame's K/V rows).

    Random draws: the reference draws ``q ~ Exp(1)`` from torch's global generator (llm.py:34-36); here the library's
    counter-based generator is keyed by ``seed`` and its own frame counter, so a seeded run is reproducible and
    independent of torch's generator state.  ``noise`` (tests) feeds given draws instead.
    """

    def __init__(self, model, tail: Optional["FrameDecoderB200"] = None, device=None, max_batch: int = 8, seed: int = 0):
        self.model = model
        if tail is None:
            if device is None:
                device = next(model.parameters()).device
            tail = FrameDecoderB200.from_reference(model, device, max_batch)
        self.tail = tail
        self.seed = int(seed)
        self.noise = None       # (B, audio_num_codebooks, audio_vocab_size) Exp(1) draws for the NEXT call only (parity tests)

    @classmethod
    def install(cls, model, **kw) -> "GenerateFrameB200":
        """``model.generate_frame`` becomes this object (an instance attribute shadowing the class's method);
        ``uninstall()`` gives the reference's method back."""
        g = cls(model, **kw)
        model.__dict__["generate_frame"] = g
        return g

    def uninstall(self) -> None:
        if self.model.__dict__.get("generate_frame") is self:
            del self.model.__dict__["generate_frame"]

    def __call__(self, tokens, tokens_mask, input_pos, temperature: float, topk: int):
        import torch
        m = self.model
        dtype = next(m.parameters()).dtype                                             # llm.py:292
        assert m.backbone.caches_are_enabled(), "backbone caches are not enabled"      # llm.py:295
        curr_backbone_mask = m.backbone_causal_mask[input_pos, :]                      # _index_causal_mask, llm.py:20-31,296
        embeds = m._embed_tokens(tokens)                                               # llm.py:297, 339-352
        h = (embeds * tokens_mask.unsqueeze(-1)).sum(dim=2)                            # llm.py:298-299
        h = m.backbone(h, input_pos=input_pos, mask=curr_backbone_mask).to(dtype=dtype)  # llm.py:300-302
        last_h = h[:, -1, :]                                                           # llm.py:304
        noise, self.noise = self.noise, None
        codes = self.tail.generate_codes(last_h.detach().to(torch.float32), int(topk), float(temperature), noise=noise,
                                         seed=self.seed)                               # llm.py:305-330
        return codes.to(device=tokens.device)

    generate_frame = __call__


def sample_topk(logits, topk: int, temperature: float, noise=None, seed: int = 0):
    """``sample_topk`` of the reference (llm.py:39-49) on the library: logits ``(B, V)`` fp32 on a CUDA device, ``noise``
    ``(B, V)`` = the Exp(1) draws (None: the library's generator) -> ``(B,)`` int32."""
    import torch
    lib = N.load()
    logits = logits.to(dtype=torch.float32).contiguous()
    B, V = logits.shape
    if noise is not None:
        noise = noise.to(device=logits.device, dtype=torch.float32).contiguous()
        assert tuple(noise.shape) == (B, V)
    with torch.cuda.device(logits.device):
        codes = torch.empty((B,), dtype=torch.int32, device=logits.device)
        N.check(lib.frt2_op_sample_topk(C.c_void_p(logits.data_ptr()), B, V, int(topk), C.c_float(temperature),
                                        C.c_void_p(noise.data_ptr()) if noise is not None else None, C.c_uint64(seed),
                                        C.c_void_p(codes.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)))
    return codes
