"""CPU model of the product path's ROUNDING POINTS (design tool; not product code, not a parity oracle).

The offline decode of oracle/codec_oracle_torch.py with the fp16 roundings libfrt2_b200 applies (fp16 GEMM operands,
fp32 accumulation, fp32 residual stream) and a choice of LayerNorm forms in front of QKV / fc1 / the iSTFT head:

  separate    n16 = fp16(LN(x32))                     (the LayerNorm kernels, DBG_NO_LNFOLD)
  fold        x16 = fp16(x32); rstd*(x16 W'^T - mean*colsum) + b'   (round-1 product path)
  fold_shift  x16 = fp16(x32 - s_row), s_row = the last known row mean of the residual stream (this round)

Prints the SNR of each form against the fp32 decode on the chosen weights, so that a numerics change can be judged
before GPU time is spent:   python tools/emulate_numerics.py [--adv] [--layers 4] [--tokens 40]
"""
from __future__ import annotations

import argparse
import dataclasses
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fireredtts2_b200.config import C0  # noqa: E402
from fireredtts2_b200.weights import adversarial_state_dict, synthetic_state_dict, synthetic_tokens  # noqa: E402
from oracle import codec_oracle as O  # noqa: E402
from oracle import codec_oracle_torch as OT  # noqa: E402

RVQ, UP, AD, BB = OT.RVQ, OT.UP, OT.AD, OT.BB


def r16(t):
    return t.clamp(-65504.0, 65504.0).half().float()


def lin16(x16, W, b=None):
    return F.linear(x16, r16(W), b)


def ln_stats(x, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return mu, torch.rsqrt(var + eps)


class Norm:
    """The LayerNorm in front of a Linear, in the three forms."""

    def __init__(self, mode):
        self.mode = mode
        self.shift = None   # (B,T,1) last known row mean of the residual stream

    def note_mean(self, x32):          # an LN(+SiLU) kernel over x32 leaves its row means behind
        self.shift = x32.mean(-1, keepdim=True)

    def apply(self, x32, g, b, eps, W, bias):
        if self.mode == "separate":
            n16 = r16(F.layer_norm(x32, (x32.shape[-1],), g, b, eps))
            return lin16(n16, W, bias)
        s = self.shift if self.mode == "fold_shift" else torch.zeros_like(x32[..., :1])
        x16 = r16(x32 - s)
        mu, rstd = ln_stats(x16, eps)                       # statistics of the fp16 copy (what the GEMM sees)
        if self.mode == "fold_shift":
            self.shift = s + mu
        Wf = r16(W * g[None, :])
        colsum = Wf.sum(1)
        bf = (bias if bias is not None else 0) + W @ b
        acc = F.linear(x16, Wf)
        return rstd * (acc - mu * colsum[None, None, :]) + bf


@torch.inference_mode()
def decode_emulated(sd, tokens, H, hop, mode):
    tokens = torch.from_numpy(np.asarray(tokens)).long()
    nm = Norm(mode)
    # RVQ: folded tables in fp32, the sum rounded once
    emb = None
    for i in range(tokens.shape[1]):
        q = F.embedding(tokens[:, i, :], sd[f"{RVQ}quantizers.{i}.codebook"]).float()
        p = f"{RVQ}quantizers.{i}.out_project."
        if p + "bias" in sd:
            q = F.linear(q, OT._wn(sd, p)[:, :, 0], sd[p + "bias"])
        emb = q if emb is None else emb + q
    z = r16(emb)
    if RVQ + "output_proj.bias" in sd:
        z = r16(lin16(z, OT._wn(sd, RVQ + "output_proj.")[:, :, 0], sd[RVQ + "output_proj.bias"]))
    h = r16(lin16(z, sd[UP + "in_proj.weight"], sd[UP + "in_proj.bias"]))
    x = r16(F.conv_transpose1d(h.transpose(1, 2), r16(sd[UP + "up_conv.weight"]), None, stride=4))
    T = 2 * x.shape[2]
    x = r16(F.gelu(F.conv_transpose1d(x, r16(sd[AD + "upsample_conv.0.weight"]), sd[AD + "upsample_conv.0.bias"], stride=2)))
    x = r16(F.gelu(F.conv_transpose1d(x, r16(sd[AD + "upsample_conv.2.weight"]), sd[AD + "upsample_conv.2.bias"], stride=1)))
    x = x[:, :, :T]
    x = OT.causal_conv(x, r16(sd[BB + "in_proj.weight"]), sd[BB + "in_proj.bias"])          # fp32 residual stream

    def resblock(p, x):
        E = x.shape[1]
        nm.note_mean(x.transpose(1, 2))
        hh = F.layer_norm(x.transpose(1, 2), (E,), sd[p + "block1.1.weight"], sd[p + "block1.1.bias"], 1e-5).transpose(1, 2)
        hh = OT.causal_conv(r16(F.silu(hh)), r16(sd[p + "block1.4.weight"]), sd[p + "block1.4.bias"])
        hh = F.layer_norm(hh.transpose(1, 2), (E,), sd[p + "block2.1.weight"], sd[p + "block2.1.bias"], 1e-5).transpose(1, 2)
        hh = OT.causal_conv(r16(F.silu(hh)), r16(sd[p + "block2.5.weight"]), sd[p + "block2.5.bias"])
        return x + hh

    for j in (0, 1):
        x = resblock(f"{BB}prior_net.{j}.", x)
    x = x.transpose(1, 2)
    B, T_, E = x.shape
    mask = OT.block_causal_mask(T, x.device)
    i = 0
    stats = []
    while f"{BB}transformers.{i}.fc1.weight" in sd:
        p = f"{BB}transformers.{i}."
        mu, rstd = ln_stats(x, 1e-5)
        stats.append((float(mu.abs().mean()), float((1 / rstd).mean()), float(x.abs().max())))
        Wqkv = torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.v_proj.weight"]])
        bqkv = torch.cat([sd[p + "self_attn.q_proj.bias"], torch.zeros(E), sd[p + "self_attn.v_proj.bias"]])
        qkv = r16(nm.apply(x, sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5, Wqkv, bqkv))
        q, k, v = qkv.split(E, dim=-1)
        sh = lambda t: t.view(B, T_, H, E // H).transpose(1, 2)
        o = F.scaled_dot_product_attention(sh(q), sh(k), sh(v), attn_mask=mask).transpose(1, 2).reshape(B, T_, E)
        x = x + lin16(r16(o), sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        g = r16(F.gelu(nm.apply(x, sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5,
                                sd[p + "fc1.weight"], sd[p + "fc1.bias"])))
        x = x + lin16(g, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        i += 1
    x = x.transpose(1, 2)
    for j in (0, 1):
        x = resblock(f"{BB}post_net.{j}.", x)
    x = x.transpose(1, 2)
    pp = nm.apply(x, sd[BB + "final_norm.weight"], sd[BB + "final_norm.bias"], 1e-6, sd[AD + "isift.out.weight"],
                  sd[AD + "isift.out.bias"]).transpose(1, 2)
    mag, ph = pp.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = r16(mag * torch.cos(ph)) + 1j * r16(mag * torch.sin(ph))
    n_fft = 4 * hop
    win = sd[AD + "isift.istft.window"]
    fr = torch.fft.irfft(S, n_fft, dim=1, norm="backward") * win[None, :, None]
    out_size = (T - 1) * hop + n_fft
    pad = (n_fft - hop) // 2
    y = F.fold(fr, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:-pad]
    env = F.fold(win.square().expand(1, T, -1).transpose(1, 2), output_size=(1, out_size), kernel_size=(1, n_fft),
                 stride=(1, hop)).squeeze()[pad:-pad]
    return (y / env).numpy(), stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--adv", action="store_true", help="adversarial weights (weights.adversarial_state_dict)")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--tokens", type=int, default=40)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    cfg = dataclasses.replace(C0, num_layers=args.layers)
    sd_np = (adversarial_state_dict if args.adv else synthetic_state_dict)(cfg, args.seed)
    sd = OT.to_torch(sd_np)
    tok = synthetic_tokens(cfg, args.batch, args.tokens, 1234)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = OT.decode(sd, tok, cfg.num_heads, cfg.hop_length).numpy()
    print(f"fp32 decode: peak {np.abs(ref).max():.3e} rms {np.sqrt((ref ** 2).mean()):.3e} finite {np.isfinite(ref).all()}")
    for mode in ("separate", "fold", "fold_shift"):
        y, stats = decode_emulated(sd, tok, cfg.num_heads, cfg.hop_length, mode)
        print(f"{mode:11s}: SNR {O.snr_db(ref, y):6.1f} dB  max-abs {np.abs(ref - y).max():.3e}")
    for i, (m, s, mx) in enumerate(stats):
        print(f"  layer {i}: mean|row mean| {m:9.3f}  mean row std {s:9.3f}  max|x| {mx:9.1f}")


if __name__ == "__main__":
    main()
