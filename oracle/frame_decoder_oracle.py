"""CPU restatement (numpy, fp32) of the frame tail of ``Model.generate_frame`` — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s checker legs may import this module; the product path
(``fireredtts2_b200/``) never does.

What is restated: ``fireredtts2/llm/llm.py:304-330`` (codebook-0 head, ``sample_topk``, the fifteen dependent decoder
passes with ``projection`` / ``audio_head`` / ``_embed_audio``) and ``llm.py:34-49`` (``_multinomial_sample_one_no_sync``,
``sample_topk``).  The transformer inside (``self.decoder``) is NOT in /root/reference: it is ``torchtune.models.qwen2.
qwen2`` (requirements.txt:1 ``torchtune``, unpinned, absent from this image) — the published Qwen2 decoder block: RMSNorm
(eps 1e-6, fp32), grouped-query attention with biased q/k/v projections and an unbiased output projection, rotary
positions on the two halves of each head (base 1e6), SwiGLU ``w2(silu(w1 x) * w3 x)``, final RMSNorm.

PIN: ``oracle/make_golden_frame_decoder.py`` runs the reference's UNMODIFIED ``Model.generate_frame`` (imported from
/root/reference) with ``transformers.models.qwen2.Qwen2Model`` standing in for the absent torchtune modules behind a
``torchtune`` shim, and records last_h / noise / logits / codes (``tests/golden/fd_*.npz``); ``tests/test_oracle_golden.py``
holds this restatement to those.  So: the frame loop and the sampler are pinned to the reference itself, the Qwen2 block
to Hugging Face's implementation of the same published architecture — "parity pinned to a stand-in" for the torchtune part.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def rms_norm(x, scale, eps):
    """torchtune RMSNorm: fp32 ``x * rsqrt(mean(x^2) + eps) * scale``."""
    x = x.astype(F32)
    r = F32(1.0) / np.sqrt(np.mean(x * x, axis=-1, keepdims=True, dtype=F32) + F32(eps))
    return (x * r).astype(F32) * scale.astype(F32)


def rope_tables(head_dim, n_pos, base):
    """cos / sin (n_pos, head_dim/2): theta_i = base^(-2i/head_dim) (Qwen2RotaryPositionalEmbeddings)."""
    i = np.arange(0, head_dim, 2, dtype=np.float64)
    theta = 1.0 / (float(base) ** (i / head_dim))
    ang = np.arange(n_pos, dtype=np.float64)[:, None] * theta[None, :]
    return np.cos(ang).astype(F32), np.sin(ang).astype(F32)


def rope(x, cos, sin):
    """x (..., head_dim): halves (x1, x2) -> (x1 cos - x2 sin, x2 cos + x1 sin)."""
    h = x.shape[-1] // 2
    x1, x2 = x[..., :h], x[..., h:]
    return np.concatenate([x1 * cos - x2 * sin, x2 * cos + x1 * sin], axis=-1).astype(F32)


def silu(x):
    return (x / (F32(1.0) + np.exp(-x.astype(F32)))).astype(F32)


def decoder_position(sd, cfg, x, pos, kc, vc):
    """One position of ``self.decoder`` (tok_embeddings / output are Identity, llm.py:9-13).  x (B, dim); kc / vc are
    per-layer lists of (B, pos, Hk, hd) arrays holding positions < pos, extended in place.  -> (B, dim) after ``norm``."""
    H, Hk, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    cos, sin = rope_tables(hd, pos + 1, cfg.rope_base)
    B = x.shape[0]
    x = x.astype(F32)
    for i in range(cfg.num_layers):
        p = f"decoder.layers.{i}."
        h = rms_norm(x, sd[p + "sa_norm.scale"], cfg.norm_eps)
        q = (h @ sd[p + "attn.q_proj.weight"].T + sd[p + "attn.q_proj.bias"]).reshape(B, H, hd)
        k = (h @ sd[p + "attn.k_proj.weight"].T + sd[p + "attn.k_proj.bias"]).reshape(B, Hk, hd)
        v = (h @ sd[p + "attn.v_proj.weight"].T + sd[p + "attn.v_proj.bias"]).reshape(B, Hk, hd)
        q = rope(q, cos[pos], sin[pos])
        k = rope(k, cos[pos], sin[pos])
        kc[i] = np.concatenate([kc[i], k[:, None]], axis=1)
        vc[i] = np.concatenate([vc[i], v[:, None]], axis=1)
        rep = H // Hk
        kk = np.repeat(kc[i], rep, axis=2)                       # (B, T, H, hd): q head h reads kv head h // rep
        vv = np.repeat(vc[i], rep, axis=2)
        s = np.einsum("bhd,bthd->bht", q, kk).astype(F32) * F32(1.0 / np.sqrt(hd))
        s = s - s.max(axis=-1, keepdims=True)
        e = np.exp(s)
        a = (e / e.sum(axis=-1, keepdims=True)).astype(F32)
        o = np.einsum("bht,bthd->bhd", a, vv).astype(F32).reshape(B, H * hd)
        x = x + o @ sd[p + "attn.output_proj.weight"].T
        h = rms_norm(x, sd[p + "mlp_norm.scale"], cfg.norm_eps)
        g = h @ sd[p + "mlp.w1.weight"].T
        u = h @ sd[p + "mlp.w3.weight"].T
        x = (x + (silu(g) * u).astype(F32) @ sd[p + "mlp.w2.weight"].T).astype(F32)
    return rms_norm(x, sd["decoder.norm.scale"], cfg.norm_eps)


def sample_topk(logits, topk, temperature, q):
    """llm.py:39-49 with the exponential draws ``q`` given (llm.py:34-36).  logits, q: (B, V) -> (B,) int32."""
    s = (logits.astype(F32) / F32(temperature)).astype(F32)
    kth = np.sort(s, axis=-1)[:, ::-1][:, topk - 1][:, None]     # torch.topk(...)[0][..., -1, None]
    s = np.where(s < kth, F32(-np.inf), s)
    m = s.max(axis=-1, keepdims=True)
    ls = (s - m) - np.log(np.exp(s - m).sum(axis=-1, keepdims=True, dtype=F32))      # log_softmax
    e = np.exp(ls - ls.max(axis=-1, keepdims=True))
    probs = (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)               # softmax
    return np.argmax(probs / q.astype(F32), axis=-1).astype(np.int32)


def generate_codes(sd, cfg, last_h, topk, temperature, noise, c0=None, forced=None):
    """llm.py:304-330 from ``last_h`` on.  noise (B, ncb, V) = the draws q of codebook i at [:, i].  ``forced`` (B, ncb)
    replaces every sampled code (teacher forcing; the logits are still those of the forced history).
    -> codes (B, ncb) int32, logits (B, ncb, V) fp32 (row 0 = c0_logits; zeros when c0 is given)."""
    sd = {k: np.asarray(v, dtype=F32) for k, v in sd.items()}
    B = last_h.shape[0]
    V, n = cfg.audio_vocab_size, cfg.audio_num_codebooks
    logits = np.zeros((B, n, V), F32)
    codes = np.zeros((B, n), np.int32)
    last_h = last_h.astype(F32)
    if c0 is None:
        logits[:, 0] = last_h @ sd["codebook0_head.weight"].T                          # llm.py:305
        c0 = sample_topk(logits[:, 0], topk, temperature, noise[:, 0])                 # llm.py:306
    codes[:, 0] = forced[:, 0] if forced is not None else c0
    kc = [np.zeros((B, 0, cfg.num_kv_heads, cfg.head_dim), F32) for _ in range(cfg.num_layers)]   # llm.py:317 reset_caches
    vc = [np.zeros((B, 0, cfg.num_kv_heads, cfg.head_dim), F32) for _ in range(cfg.num_layers)]
    P = sd["projection.weight"].T
    decoder_position(sd, cfg, last_h @ P, 0, kc, vc)                                   # position 0 of llm.py:308,320-322
    for i in range(1, n):
        emb = sd["audio_embeddings.weight"][codes[:, i - 1] + (i - 1) * V]             # llm.py:307,325 (_embed_audio)
        h = decoder_position(sd, cfg, emb @ P, i, kc, vc)
        logits[:, i] = h @ sd["audio_head"][i - 1]                                     # llm.py:323
        ci = sample_topk(logits[:, i], 10, 0.75, noise[:, i])                          # llm.py:324
        codes[:, i] = forced[:, i] if forced is not None else ci
    return codes, logits
