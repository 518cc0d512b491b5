"""CPU port of the reference decode path on torch CPU ops — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Same algorithm as ``oracle/codec_oracle.py`` (every function cites the same reference lines) but expressed with the
ATen CPU kernels the reference itself dispatches to (conv1d, conv_transpose1d, linear, SDPA, layer_norm, irfft,
fold), so that ``bench.py --impl reference`` / ``cpu_baseline`` time what the reference's own CPU decode costs on
the box's host cores.  /root/reference cannot travel to the GPU box; this port (pinned to the same golden vectors in
``tests/test_oracle_golden.py``) does.  fp32, ``torch.inference_mode``.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

RVQ, UP, AD = "rvq.", "upsample.", "acoustic_decoder."
BB = AD + "backbone."


def to_torch(sd) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v for k, v in sd.items()}


def _wn(sd, prefix):
    """weight_norm dim=0 (reference rvq.py:8-13)."""
    g, v = sd[prefix + "parametrizations.weight.original0"], sd[prefix + "parametrizations.weight.original1"]
    return v * (g / v.norm(dim=(1, 2), keepdim=True))


def rvq_decode_codes(sd, tokens):
    """ResidualVQ.decode_codes (reference rvq.py:145-164); tokens (B,nq,L) -> (B,E,L)."""
    B, nq, L = tokens.shape
    emb = None
    for i in range(nq):
        q = F.embedding(tokens[:, i, :], sd[f"{RVQ}quantizers.{i}.codebook"]).transpose(1, 2).float()
        p = f"{RVQ}quantizers.{i}.out_project."
        if p + "bias" in sd:
            q = F.conv1d(q, _wn(sd, p), sd[p + "bias"])
        emb = q if emb is None else emb + q
    if RVQ + "output_proj.bias" in sd:
        emb = F.conv1d(emb, _wn(sd, RVQ + "output_proj."), sd[RVQ + "output_proj.bias"])
    return emb


def causal_conv(x, w, b):
    """CausalConv1d.forward (reference decoder.py:88-91): left pad k-1."""
    return F.conv1d(F.pad(x, (w.shape[2] - 1, 0)), w, b)


def resnet_block(sd, p, x):
    """CausalResnetBlock.forward (reference decoder.py:133-148)."""
    E = x.shape[1]
    h = F.layer_norm(x.transpose(1, 2), (E,), sd[p + "block1.1.weight"], sd[p + "block1.1.bias"], 1e-5).transpose(1, 2)
    h = causal_conv(F.silu(h), sd[p + "block1.4.weight"], sd[p + "block1.4.bias"])
    h = F.layer_norm(h.transpose(1, 2), (E,), sd[p + "block2.1.weight"], sd[p + "block2.1.bias"], 1e-5).transpose(1, 2)
    h = causal_conv(F.silu(h), sd[p + "block2.5.weight"], sd[p + "block2.5.bias"])
    return x + h


def block_causal_mask(T, device):
    """make_block_causal_mask (reference utils.py:19-38): key j visible to query i iff j <= (i | 7)."""
    i = torch.arange(T, device=device)
    return i[None, :] <= (i[:, None] | 7)


def transformer_layer(sd, p, x, H, mask):
    """WhisperEncoderLayer.forward (reference whisper.py:142-162)."""
    B, T, E = x.shape
    a = F.layer_norm(x, (E,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5)
    q = F.linear(a, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"])
    k = F.linear(a, sd[p + "self_attn.k_proj.weight"])
    v = F.linear(a, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
    sh = lambda t: t.view(B, T, H, E // H).transpose(1, 2)
    o = F.scaled_dot_product_attention(sh(q), sh(k), sh(v), attn_mask=mask)
    o = o.transpose(1, 2).reshape(B, T, E)
    x = x + F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    f = F.layer_norm(x, (E,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
    f = F.linear(F.gelu(F.linear(f, sd[p + "fc1.weight"], sd[p + "fc1.bias"])), sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    return x + f


@torch.inference_mode()
def decode(sd, tokens, num_heads, hop=240):
    """RedCodecInfer.decode (reference model.py:307-324): tokens (B,nq,L) -> (B, 8*hop*L) fp32."""
    if isinstance(tokens, np.ndarray):
        tokens = torch.from_numpy(tokens)
    tokens = tokens.long()
    if tokens.numel() and (int(tokens.min()) < 0 or int(tokens.max()) >= sd[f"{RVQ}quantizers.0.codebook"].shape[0]):
        raise IndexError("index out of range in self")
    z = rvq_decode_codes(sd, tokens)                                             # (B,E,L)
    # UpConv (reference model.py:142-148)
    h = F.linear(z.transpose(1, 2), sd[UP + "in_proj.weight"], sd[UP + "in_proj.bias"]).transpose(1, 2)
    x = F.conv_transpose1d(h, sd[UP + "up_conv.weight"], None, stride=4)           # (B,E,4L)
    # upsample_conv (reference decoder.py:571-589, 610-616)
    T = 2 * x.shape[2]
    x = F.gelu(F.conv_transpose1d(x, sd[AD + "upsample_conv.0.weight"], sd[AD + "upsample_conv.0.bias"], stride=2))
    x = F.gelu(F.conv_transpose1d(x, sd[AD + "upsample_conv.2.weight"], sd[AD + "upsample_conv.2.bias"], stride=1))
    x = x[:, :, :T]
    # backbone (reference decoder.py:248-274)
    x = causal_conv(x, sd[BB + "in_proj.weight"], sd[BB + "in_proj.bias"])
    for j in (0, 1):
        x = resnet_block(sd, f"{BB}prior_net.{j}.", x)
    x = x.transpose(1, 2)
    mask = block_causal_mask(T, x.device)
    i = 0
    while f"{BB}transformers.{i}.fc1.weight" in sd:
        x = transformer_layer(sd, f"{BB}transformers.{i}.", x, num_heads, mask)
        i += 1
    x = x.transpose(1, 2)
    for j in (0, 1):
        x = resnet_block(sd, f"{BB}post_net.{j}.", x)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[BB + "final_norm.weight"], sd[BB + "final_norm.bias"], 1e-6)
    # ISTFTHead / ISTFT "same" (reference decoder.py:503-518, 350-405)
    p = F.linear(x, sd[AD + "isift.out.weight"], sd[AD + "isift.out.bias"]).transpose(1, 2)
    mag, ph = p.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(ph) + 1j * torch.sin(ph))
    n_fft = 4 * hop
    win = sd[AD + "isift.istft.window"]
    fr = torch.fft.irfft(S, n_fft, dim=1, norm="backward") * win[None, :, None]
    out_size = (T - 1) * hop + n_fft
    pad = (n_fft - hop) // 2
    y = F.fold(fr, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:-pad]
    env = F.fold(win.square().expand(1, T, -1).transpose(1, 2), output_size=(1, out_size), kernel_size=(1, n_fft),
                 stride=(1, hop)).squeeze()[pad:-pad]
    return y / env
