"""Decode-side weights in the reference's ``state_dict`` naming (SURVEY.md §8a "weights consumed").

``synthetic_state_dict`` builds random-init weights of the reference codec architecture with a numpy
generator, so the *same* weights can be regenerated bit-for-bit on the GPU box (where
``/root/reference`` does not exist) and loaded into the reference modules here (``load_state_dict``)
to produce the golden vectors.  Distributions follow the reference constructors' own init:
``nn.Linear`` / ``nn.ConvTranspose1d`` / weight-normed 1x1 convs use torch's default
kaiming-uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)); every ``nn.Conv1d`` inside
``AcousticDecoder`` is re-initialised trunc_normal(std=0.02) with zero bias
(reference decoder.py:599-602); codebooks are N(0,1) (they are zeros by default, rvq.py:44-46).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List

import numpy as np

from .config import CodecConfig

PREFIX_RVQ = "rvq."
PREFIX_UP = "upsample."
PREFIX_AD = "acoustic_decoder."


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _trunc_normal(rng, shape, std=0.02):
    x = rng.standard_normal(size=shape).astype(np.float32)
    np.clip(x, -2.0 / std, 2.0 / std, out=x)  # torch trunc_normal_ default cut is +-2 (absolute)
    return (x * std).astype(np.float32)


def _linear(rng, sd, name, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(rng, (out_f, in_f), b)
    if bias:
        sd[name + ".bias"] = _uniform(rng, (out_f,), b)


def _wn_conv1x1(rng, sd, name, out_c, in_c):
    """weight_norm(nn.Conv1d(in,out,1)): original0 = g (out,1,1), original1 = v (out,in,1)."""
    b = 1.0 / math.sqrt(in_c)
    v = _uniform(rng, (out_c, in_c, 1), b)
    g = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
    # perturb g so that the g/||v|| factor is actually exercised
    g = (g * rng.uniform(0.7, 1.3, size=g.shape)).astype(np.float32)
    sd[name + ".parametrizations.weight.original0"] = g
    sd[name + ".parametrizations.weight.original1"] = v
    sd[name + ".bias"] = _uniform(rng, (out_c,), b)


def _layer_norm(rng, sd, name, dim):
    # torch default is ones/zeros; perturb so affine terms are exercised by the parity tests
    sd[name + ".weight"] = (1.0 + 0.1 * rng.standard_normal(dim)).astype(np.float32)
    sd[name + ".bias"] = (0.05 * rng.standard_normal(dim)).astype(np.float32)


def _causal_conv(rng, sd, name, out_c, in_c, k):
    sd[name + ".weight"] = _trunc_normal(rng, (out_c, in_c, k))
    sd[name + ".bias"] = (0.02 * rng.standard_normal(out_c)).astype(np.float32)


def synthetic_state_dict(cfg: CodecConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}
    E = cfg.embed_dim
    # ---- rvq ----
    for i in range(cfg.num_quantizers):
        q = f"{PREFIX_RVQ}quantizers.{i}"
        sd[q + ".codebook"] = rng.standard_normal((cfg.codebook_size, cfg.codebook_dim)).astype(np.float32)
        if cfg.has_out_project:
            _wn_conv1x1(rng, sd, q + ".out_project", cfg.rvq_dim, cfg.codebook_dim)
    if cfg.has_output_proj:
        _wn_conv1x1(rng, sd, PREFIX_RVQ + "output_proj", cfg.output_dim, cfg.rvq_dim)
    # ---- upsample (UpConv) ----
    S = cfg.upconv_stride
    _linear(rng, sd, PREFIX_UP + "in_proj", S * E, E)
    bt = 1.0 / math.sqrt(E * S)  # ConvTranspose1d fan_in = weight.size(1) * k = out_channels * k
    sd[PREFIX_UP + "up_conv.weight"] = _uniform(rng, (S * E, E, S), bt)
    # ---- acoustic decoder ----
    a = PREFIX_AD
    for idx in (0, 2):
        b = 1.0 / math.sqrt(E * 3)
        sd[f"{a}upsample_conv.{idx}.weight"] = _uniform(rng, (E, E, 3), b)
        sd[f"{a}upsample_conv.{idx}.bias"] = _uniform(rng, (E,), b)
    bb = a + "backbone."
    _causal_conv(rng, sd, bb + "in_proj", E, E, 7)
    for net in ("prior_net", "post_net"):
        for j in (0, 1):
            p = f"{bb}{net}.{j}."
            _layer_norm(rng, sd, p + "block1.1", E)
            _causal_conv(rng, sd, p + "block1.4", E, E, 3)
            _layer_norm(rng, sd, p + "block2.1", E)
            _causal_conv(rng, sd, p + "block2.5", E, E, 3)
    for i in range(cfg.num_layers):
        t = f"{bb}transformers.{i}."
        _linear(rng, sd, t + "self_attn.k_proj", E, E, bias=False)
        _linear(rng, sd, t + "self_attn.v_proj", E, E)
        _linear(rng, sd, t + "self_attn.q_proj", E, E)
        _linear(rng, sd, t + "self_attn.out_proj", E, E)
        _layer_norm(rng, sd, t + "self_attn_layer_norm", E)
        _linear(rng, sd, t + "fc1", 4 * E, E)
        _linear(rng, sd, t + "fc2", E, 4 * E)
        _layer_norm(rng, sd, t + "final_layer_norm", E)
    _layer_norm(rng, sd, bb + "final_norm", E)
    _linear(rng, sd, a + "isift.out", cfg.n_fft + 2, E)
    n = np.arange(cfg.n_fft, dtype=np.float64)
    # torch.hann_window(N) (periodic): 0.5 - 0.5 cos(2 pi n / N), computed in fp32 by torch;
    # the value loaded from a real checkpoint is used verbatim, this is only the synthetic stand-in.
    sd[a + "isift.istft.window"] = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / cfg.n_fft)).astype(np.float32)
    return sd


def adversarial_state_dict(cfg: CodecConfig, seed: int = 0, mean_offset: float = 100.0, outlier_scale: float = 100.0,
                           fc1_scale: float = 1000.0, fc2_scale: float = 1.0 / 1000.0) -> Dict[str, np.ndarray]:
    """Weights of the same architecture that stress the fp16-operand / folded-LayerNorm path the way a trained
    checkpoint can (a random init with sigma = 0.02 is the kindest possible input):

    * every LayerNorm gamma log-uniform in [0.1, 5], beta ~ N(0, 0.5);
    * a constant added to the bias of the convolution that opens the fp32 residual stream (``backbone.in_proj``), so
      that every row of the stream carries a mean of ~50 times its spread (100 against a row std of ~2);
    * four outlier channels of that convolution scaled by ``outlier_scale`` ("massive activations");
    * fc1 scaled up / fc2 scaled down by the same factor so that the GELU activations reach 1e3 - 1e4 (fp16 operand
      range) while the function stays well conditioned.

    Same numpy generator discipline as ``synthetic_state_dict``: regenerated bit for bit on the GPU box and loaded into
    the real reference in the build container (oracle/make_golden.py) for the golden waveforms."""
    sd = synthetic_state_dict(cfg, seed)
    rng = np.random.default_rng(seed + 104729)
    E = cfg.embed_dim
    for k in sorted(sd):
        is_ln = (k.endswith(("block1.1.weight", "block2.1.weight", "layer_norm.weight", "final_norm.weight")))
        if is_ln:
            sd[k] = np.exp(rng.uniform(math.log(0.1), math.log(5.0), size=sd[k].shape)).astype(np.float32)
            sd[k[:-6] + "bias"] = (0.5 * rng.standard_normal(sd[k].shape)).astype(np.float32)
    bb = PREFIX_AD + "backbone."
    ch = rng.choice(E, size=4, replace=False)
    w = sd[bb + "in_proj.weight"].copy()
    b = sd[bb + "in_proj.bias"].copy()
    w[ch] *= outlier_scale
    b[ch] *= outlier_scale
    b += np.float32(mean_offset)
    sd[bb + "in_proj.weight"], sd[bb + "in_proj.bias"] = w, b.astype(np.float32)
    for i in range(cfg.num_layers):
        t = f"{bb}transformers.{i}."
        sd[t + "fc1.weight"] = (sd[t + "fc1.weight"] * np.float32(fc1_scale)).astype(np.float32)
        sd[t + "fc1.bias"] = (sd[t + "fc1.bias"] * np.float32(fc1_scale)).astype(np.float32)
        sd[t + "fc2.weight"] = (sd[t + "fc2.weight"] * np.float32(fc2_scale)).astype(np.float32)
    return sd


def synthetic_encode_tensors(cfg: CodecConfig, seed: int = 0, input_dim: int = None) -> Dict[str, np.ndarray]:
    """The encode-side tensors of ``ResidualVQ`` that ``synthetic_state_dict`` leaves out (they are not part of the
    decode path): ``rvq.input_proj`` (present iff input_dim != rvq_dim, reference rvq.py:110-114) and every
    ``rvq.quantizers.{i}.in_project`` (present iff rvq_dim != codebook_dim, rvq.py:28-34).  Own generator, so the
    decode-side weights (and the golden vectors made from them) do not change."""
    rng = np.random.default_rng(seed + 7919)
    sd: Dict[str, np.ndarray] = {}
    input_dim = cfg.embed_dim if input_dim is None else input_dim
    if input_dim != cfg.rvq_dim:
        _wn_conv1x1(rng, sd, PREFIX_RVQ + "input_proj", cfg.rvq_dim, input_dim)
    if cfg.has_out_project:
        for i in range(cfg.num_quantizers):
            _wn_conv1x1(rng, sd, f"{PREFIX_RVQ}quantizers.{i}.in_project", cfg.codebook_dim, cfg.rvq_dim)
    return sd


def decode_keys(cfg: CodecConfig) -> List[str]:
    """Every state_dict key the decode path consumes, in the order they are handed to the C-ABI."""
    return list(synthetic_state_dict_keys(cfg))


def synthetic_state_dict_keys(cfg: CodecConfig) -> Iterable[str]:
    # key names do not depend on widths, only on nq / num_layers / projection presence
    names: List[str] = []
    for i in range(cfg.num_quantizers):
        q = f"{PREFIX_RVQ}quantizers.{i}"
        names.append(q + ".codebook")
        if cfg.has_out_project:
            names += [q + ".out_project.parametrizations.weight.original0",
                      q + ".out_project.parametrizations.weight.original1",
                      q + ".out_project.bias"]
    if cfg.has_output_proj:
        o = PREFIX_RVQ + "output_proj"
        names += [o + ".parametrizations.weight.original0", o + ".parametrizations.weight.original1", o + ".bias"]
    names += [PREFIX_UP + "in_proj.weight", PREFIX_UP + "in_proj.bias", PREFIX_UP + "up_conv.weight"]
    a = PREFIX_AD
    for idx in (0, 2):
        names += [f"{a}upsample_conv.{idx}.weight", f"{a}upsample_conv.{idx}.bias"]
    bb = a + "backbone."
    names += [bb + "in_proj.weight", bb + "in_proj.bias"]
    for net in ("prior_net", "post_net"):
        for j in (0, 1):
            p = f"{bb}{net}.{j}."
            for m in ("block1.1", "block1.4", "block2.1", "block2.5"):
                names += [p + m + ".weight", p + m + ".bias"]
    for i in range(cfg.num_layers):
        t = f"{bb}transformers.{i}."
        names += [t + "self_attn.k_proj.weight",
                  t + "self_attn.v_proj.weight", t + "self_attn.v_proj.bias",
                  t + "self_attn.q_proj.weight", t + "self_attn.q_proj.bias",
                  t + "self_attn.out_proj.weight", t + "self_attn.out_proj.bias",
                  t + "self_attn_layer_norm.weight", t + "self_attn_layer_norm.bias",
                  t + "fc1.weight", t + "fc1.bias", t + "fc2.weight", t + "fc2.bias",
                  t + "final_layer_norm.weight", t + "final_layer_norm.bias"]
    names += [bb + "final_norm.weight", bb + "final_norm.bias",
              a + "isift.out.weight", a + "isift.out.bias", a + "isift.istft.window"]
    return names


def weight_norm_materialise(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """W[o] = g[o] * v[o] / ||v[o]||_2 over (in,k) — torch weight_norm dim=0 (reference rvq.py:8-13)."""
    norm = np.sqrt((v.astype(np.float32) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True, dtype=np.float32))
    return (v * (g / norm)).astype(np.float32)


def normalise_state_dict(sd) -> Dict[str, np.ndarray]:
    """Accept a torch or numpy state_dict (full RedCodec or decode-only) and return fp32 numpy arrays."""
    out = {}
    for k, v in sd.items():
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        v = np.asarray(v)
        if v.dtype.kind == "f":
            v = np.ascontiguousarray(v, dtype=np.float32)
        out[k] = v
    return out


def synthetic_tokens(cfg: CodecConfig, batch: int, length: int, seed: int = 1234, dtype=np.int64) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, cfg.codebook_size, size=(batch, cfg.num_quantizers, length)).astype(dtype)
