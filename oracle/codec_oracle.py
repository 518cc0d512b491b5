"""CPU oracle for the FireRedTTS-2 codec *decode* path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's algorithm (reference = /root/reference/fireredtts2/codec/*,
pure PyTorch).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path (``fireredtts2_b200``) never does
and has no CPU fallback.

Parity pin: the reference ships no golden vectors / known-answer tests for this path,
so the oracle is pinned against outputs of the reference itself, generated in the build container by
``oracle/make_golden.py`` (imports ``/root/reference``) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks the oracle against them on every run.

All tensors are time-major ``(B, T, C)`` internally (the reference is channel-major ``(B, C, T)`` in the
conv stages); the five streaming caches are exported in the reference's own layouts.
Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.special import erf as _erf

RVQ, UP, AD = "rvq.", "upsample.", "acoustic_decoder."
BB = AD + "backbone."


# --------------------------------------------------------------------------------------------
# elementary ops
# --------------------------------------------------------------------------------------------
def _f(sd, key, dt):
    return np.asarray(sd[key], dtype=dt)


def weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """torch weight_norm, dim=0: W[o] = g[o] * v[o] / ||v[o]||  (reference rvq.py:8-13)."""
    n = np.sqrt((v * v).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return v * (g / n)


def gelu(x):
    """Exact erf GELU (reference decoder.py:580,588; whisper.py:157 F.gelu default)."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def silu(x):
    """x * sigmoid(x) (reference decoder.py:121,129 nn.SiLU).  For x << 0 exp(-x) overflows to +inf and x / inf = -0.0, the
    value torch returns; the overflow flag is expected there (adversarial-weights goldens), so it is not reported."""
    with np.errstate(over="ignore"):
        return (x / (1.0 + np.exp(-x))).astype(x.dtype)


def layer_norm(x, w, b, eps):
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return (xc / np.sqrt(var + eps) * w + b).astype(x.dtype)


def causal_conv1d(x, W, b, hist=None):
    """CausalConv1d.forward / forward_chunk (reference decoder.py:78-101).

    x: (B,T,Cin) time-major.  W: (Cout,Cin,k) [nn.Conv1d layout].  hist: (B,k-1,Cin) previous inputs
    (zeros == left padding when None).  Returns (y (B,T,Cout), new_hist (B,k-1,Cin)).
    y[t] = b + sum_kk W[:,:,kk] @ xp[t+kk],  xp = hist ++ x.
    """
    B, T, Cin = x.shape
    k = W.shape[2]
    if hist is None:
        hist = np.zeros((B, k - 1, Cin), dtype=x.dtype)
    xp = np.concatenate([hist, x], axis=1)
    y = np.broadcast_to(b, (B, T, W.shape[0])).astype(x.dtype).copy()
    for kk in range(k):
        y += xp[:, kk:kk + T, :] @ W[:, :, kk].T
    return y, xp[:, xp.shape[1] - (k - 1):, :].copy()


def conv_transpose1d(x, W, b, stride):
    """nn.ConvTranspose1d, padding 0.  x (B,T,Cin); W (Cin,Cout,k).  y[t*stride+kk] += x[t] @ W[:,:,kk]."""
    B, T, Cin = x.shape
    k = W.shape[2]
    Tout = (T - 1) * stride + k
    y = np.zeros((B, Tout, W.shape[1]), dtype=x.dtype)
    for kk in range(k):
        y[:, kk:kk + (T - 1) * stride + 1:stride, :] += x @ W[:, :, kk]
    if b is not None:
        y += b
    return y


# --------------------------------------------------------------------------------------------
# RVQ dequantisation  (reference rvq.py:56-60, 145-164)
# --------------------------------------------------------------------------------------------
def check_indices(tokens, K):
    """F.embedding raises IndexError on CPU for idx outside [0,K) (SURVEY.md §8b [probe])."""
    if tokens.size and (tokens.min() < 0 or tokens.max() >= K):
        raise IndexError("index out of range in self")


def rvq_gather(sd, tokens):
    """VectorQuantize.decode_code for every codebook: rows[b,l,i,:] = codebook_i[tokens[b,i,l]]
    (bit-exact fp32 copy).  tokens (B,nq,L) -> rows (B,L,nq,cd)."""
    B, nq, L = tokens.shape
    cb0 = sd[f"{RVQ}quantizers.0.codebook"]
    check_indices(tokens, cb0.shape[0])
    rows = np.empty((B, L, nq, cb0.shape[1]), dtype=np.float32)
    for i in range(nq):
        rows[:, :, i, :] = np.asarray(sd[f"{RVQ}quantizers.{i}.codebook"], dtype=np.float32)[tokens[:, i, :]]
    return rows


def rvq_decode_codes(sd, tokens, dt=np.float32):
    """ResidualVQ.decode_codes (reference rvq.py:145-164), time-major.

    emb = 0; for i<nq: emb += out_project_i(codebook_i[idx_i]); z = output_proj(emb).
    Returns (emb (B,L,rd), z (B,L,E)).  The sum runs in index order from +0.0, so with Identity
    projections it is bit-reproducible.
    """
    B, nq, L = tokens.shape
    rows = rvq_gather(sd, tokens).astype(dt)
    has_proj = f"{RVQ}quantizers.0.out_project.bias" in sd
    emb = None
    for i in range(nq):
        q = rows[:, :, i, :]
        if has_proj:
            p = f"{RVQ}quantizers.{i}.out_project."
            W = weight_norm(_f(sd, p + "parametrizations.weight.original0", dt),
                            _f(sd, p + "parametrizations.weight.original1", dt))[:, :, 0]
            q = q @ W.T + _f(sd, p + "bias", dt)
        emb = (np.zeros_like(q) + q) if emb is None else emb + q
    z = emb
    if f"{RVQ}output_proj.bias" in sd:
        p = f"{RVQ}output_proj."
        W = weight_norm(_f(sd, p + "parametrizations.weight.original0", dt),
                        _f(sd, p + "parametrizations.weight.original1", dt))[:, :, 0]
        z = emb @ W.T + _f(sd, p + "bias", dt)
    return emb, z


def _wn_1x1(sd, prefix, dt=np.float32):
    W = weight_norm(_f(sd, prefix + "parametrizations.weight.original0", dt),
                    _f(sd, prefix + "parametrizations.weight.original1", dt))[:, :, 0]
    return W.astype(dt), _f(sd, prefix + "bias", dt)


def rvq_encode_codes(sd, z, nq: Optional[int] = None, dt=np.float32):
    """ResidualVQ.encode_codes (reference rvq.py:128-143) over VectorQuantize.encode_code (rvq.py:62-89).

    z (B, input_dim, T) channel-major as the reference receives it (model.py:240).  Returns
    (codes (nq,B,T) int64, margin (nq,B,T)): margin = best - second best of -dist for every decision, i.e. how far the
    chosen code is from a tie — a CUDA (or any other) implementation with a different summation order may legitimately
    pick the runner-up only where margin is within fp32 rounding of the distance.
      residual = input_proj(z)                                   rvq.py:129-130
      z_e = in_project(residual)                                 rvq.py:65
      dist = |z_e|^2 - (2 z_e) @ C^T + |C|^2 ; idx = argmax(-dist)   rvq.py:71-78 (first maximum)
      z_q = z_e + (C[idx] - z_e) ; residual -= out_project(z_q)  rvq.py:82-86,138
    """
    z = np.asarray(z, dtype=dt)
    B, D, T = z.shape
    x = np.ascontiguousarray(z.transpose(0, 2, 1)).reshape(B * T, D)      # (B*T, D): "(b t) d", rvq.py:68
    if f"{RVQ}input_proj.bias" in sd:
        W, b = _wn_1x1(sd, RVQ + "input_proj.", dt)
        x = x @ W.T + b
    residual = x.astype(dt)
    n_all = 0
    while f"{RVQ}quantizers.{n_all}.codebook" in sd:
        n_all += 1
    nq = n_all if nq is None else nq
    has_proj = f"{RVQ}quantizers.0.in_project.bias" in sd
    codes = np.empty((nq, B, T), dtype=np.int64)
    margin = np.empty((nq, B, T), dtype=dt)
    for i in range(nq):
        q = f"{RVQ}quantizers.{i}."
        C = _f(sd, q + "codebook", dt)
        z_e = residual
        if has_proj:
            Wi, bi = _wn_1x1(sd, q + "in_project.", dt)
            z_e = (residual @ Wi.T + bi).astype(dt)
        a = (z_e * z_e).sum(axis=1, keepdims=True, dtype=dt)
        c2 = (C * C).sum(axis=1, keepdims=True, dtype=dt).T
        dist = (a - ((dt(2) * z_e) @ C.T).astype(dt)).astype(dt) + c2
        neg = -dist
        idx = neg.argmax(axis=1)                                          # first maximum, like torch.max(1)[1] on CPU
        part = np.partition(neg, -2, axis=1)
        margin[i] = (part[:, -1] - part[:, -2]).reshape(B, T)
        z_q = z_e + (C[idx] - z_e)
        if has_proj:
            Wo, bo = _wn_1x1(sd, q + "out_project.", dt)
            z_q = (z_q @ Wo.T + bo).astype(dt)
        residual = (residual - z_q).astype(dt)
        codes[i] = idx.reshape(B, T)
    return codes, margin


# --------------------------------------------------------------------------------------------
# UpConv 12.5 -> 50 Hz  (reference model.py:142-148)
# --------------------------------------------------------------------------------------------
def upconv(sd, z, dt=np.float32):
    """h = in_proj(z); x50 = ConvTranspose1d(k=s=4, no bias)(h).  (B,L,E) -> (B,4L,E)."""
    h = z @ _f(sd, UP + "in_proj.weight", dt).T + _f(sd, UP + "in_proj.bias", dt)
    W = _f(sd, UP + "up_conv.weight", dt)
    return conv_transpose1d(h, W, None, W.shape[2])


# --------------------------------------------------------------------------------------------
# upsample_conv 50 -> 100 Hz  (reference decoder.py:571-589, 610-616, 624-655)
# --------------------------------------------------------------------------------------------
def upsample_conv(sd, x50, cache=None, dt=np.float32):
    """Two ConvTranspose1d (k3 s2, k3 s1) each followed by exact GELU, strictly causal.

    Offline (cache None): output trimmed to 2*T50 frames (decoder.py:615).  Streaming: ``cache`` is the
    reference's up_conv_cache in time-major form (B,3,E) = [last x50 frame | last two post-GELU frames of
    the first conv] (decoder.py:624-655); returns the same for the next chunk.
    """
    W0, b0 = _f(sd, AD + "upsample_conv.0.weight", dt), _f(sd, AD + "upsample_conv.0.bias", dt)
    W2, b2 = _f(sd, AD + "upsample_conv.2.weight", dt), _f(sd, AD + "upsample_conv.2.bias", dt)
    T50 = x50.shape[1]
    x = x50
    if cache is not None:
        x = np.concatenate([cache[:, 0:1, :], x], axis=1)
    new_c1 = x[:, -1:, :]
    a = conv_transpose1d(x, W0, b0, 2)[:, :-1, :]          # decoder.py:640 "remove extra 1 frame"
    if cache is not None:
        a = a[:, 2:, :]
    a = gelu(a)
    if cache is not None:
        a = np.concatenate([cache[:, 1:3, :], a], axis=1)
    new_c2 = a[:, -2:, :]
    v = conv_transpose1d(a, W2, b2, 1)[:, :-2, :]          # decoder.py:649 "remove extra 2 frame"
    if cache is not None:
        v = v[:, 2:, :]
    y = gelu(v)
    assert y.shape[1] == 2 * T50
    return y, np.concatenate([new_c1, new_c2], axis=1)


# --------------------------------------------------------------------------------------------
# backbone  (reference decoder.py:105-171, 225-320; whisper.py:23-192; utils.py:19-38)
# --------------------------------------------------------------------------------------------
def resnet_block(sd, prefix, x, cache=None, dt=np.float32):
    """CausalResnetBlock.forward / forward_chunk (reference decoder.py:105-171).

    out = x + conv3(SiLU(LN(conv3(SiLU(LN(x)))))); cache (B,2,2E) time-major = last two post-LN-SiLU
    inputs of conv1 ++ conv2 (channel-concatenated like the reference's (B,2E,2))."""
    E = x.shape[-1]
    c1 = None if cache is None else cache[:, :, :E]
    c2 = None if cache is None else cache[:, :, E:]
    h = silu(layer_norm(x, _f(sd, prefix + "block1.1.weight", dt), _f(sd, prefix + "block1.1.bias", dt), 1e-5))
    h, n1 = causal_conv1d(h, _f(sd, prefix + "block1.4.weight", dt), _f(sd, prefix + "block1.4.bias", dt), c1)
    h = silu(layer_norm(h, _f(sd, prefix + "block2.1.weight", dt), _f(sd, prefix + "block2.1.bias", dt), 1e-5))
    h, n2 = causal_conv1d(h, _f(sd, prefix + "block2.5.weight", dt), _f(sd, prefix + "block2.5.bias", dt), c2)
    return x + h, np.concatenate([n1, n2], axis=2)


def block_causal_visible_end(i: np.ndarray, block: int = 8) -> np.ndarray:
    """make_block_causal_mask (reference utils.py:19-38): key j visible to query i iff j <= i or
    floor(j/8) == floor(i/8)  <=>  j <= (i | 7)."""
    return i | (block - 1)


def attention(q, k, v, num_heads, q_pos0: int, mask_mode: str):
    """softmax(q k^T / sqrt(hd) [+ mask]) v per head (reference whisper.py:49-79 / 81-118).

    q (B,Tq,E); k,v (B,Tk,E).  mask_mode 'block_causal': query row r has absolute position q_pos0+r and
    sees keys j <= (pos|7) (offline, utils.py:19-38).  'none': every key visible (forward_chunk passes
    attn_mask=None, whisper.py:107-113)."""
    B, Tq, E = q.shape
    Tk = k.shape[1]
    hd = E // num_heads
    qh = q.reshape(B, Tq, num_heads, hd).transpose(0, 2, 1, 3)
    kh = k.reshape(B, Tk, num_heads, hd).transpose(0, 2, 1, 3)
    vh = v.reshape(B, Tk, num_heads, hd).transpose(0, 2, 1, 3)
    out = np.empty_like(qh)
    scale = 1.0 / math.sqrt(hd)
    step = 512
    for s in range(0, Tq, step):
        e = min(Tq, s + step)
        pos = np.arange(s, e) + q_pos0
        kend = Tk if mask_mode == "none" else min(Tk, int(block_causal_visible_end(pos).max()) + 1)
        sc = (qh[:, :, s:e, :] @ kh[:, :, :kend, :].transpose(0, 1, 3, 2)) * q.dtype.type(scale)
        if mask_mode != "none":
            vis = np.arange(kend)[None, :] <= block_causal_visible_end(pos)[:, None]
            sc = np.where(vis[None, None], sc, -np.inf)
        sc = sc - sc.max(axis=-1, keepdims=True)
        p = np.exp(sc)
        p /= p.sum(axis=-1, keepdims=True)
        out[:, :, s:e, :] = p.astype(q.dtype) @ vh[:, :, :kend, :]
    return out.transpose(0, 2, 1, 3).reshape(B, Tq, E)


def transformer_layer(sd, prefix, x, num_heads, kv=None, mask_mode="block_causal", dt=np.float32):
    """WhisperEncoderLayer.forward / forward_chunk (reference whisper.py:121-192).

    kv: optional (k_cache (B,Tc,E), v_cache (B,Tc,E)).  Returns (x, (k_all, v_all))."""
    a = layer_norm(x, _f(sd, prefix + "self_attn_layer_norm.weight", dt),
                   _f(sd, prefix + "self_attn_layer_norm.bias", dt), 1e-5)
    q = a @ _f(sd, prefix + "self_attn.q_proj.weight", dt).T + _f(sd, prefix + "self_attn.q_proj.bias", dt)
    k = a @ _f(sd, prefix + "self_attn.k_proj.weight", dt).T                      # no bias (whisper.py:37)
    v = a @ _f(sd, prefix + "self_attn.v_proj.weight", dt).T + _f(sd, prefix + "self_attn.v_proj.bias", dt)
    q_pos0 = 0
    if kv is not None:
        q_pos0 = kv[0].shape[1]
        k = np.concatenate([kv[0], k], axis=1)
        v = np.concatenate([kv[1], v], axis=1)
    o = attention(q, k, v, num_heads, q_pos0, mask_mode)
    x = x + (o @ _f(sd, prefix + "self_attn.out_proj.weight", dt).T + _f(sd, prefix + "self_attn.out_proj.bias", dt))
    f = layer_norm(x, _f(sd, prefix + "final_layer_norm.weight", dt), _f(sd, prefix + "final_layer_norm.bias", dt), 1e-5)
    g = gelu(f @ _f(sd, prefix + "fc1.weight", dt).T + _f(sd, prefix + "fc1.bias", dt))
    x = x + (g @ _f(sd, prefix + "fc2.weight", dt).T + _f(sd, prefix + "fc2.bias", dt))
    return x, (k, v)


def num_layers_of(sd) -> int:
    n = 0
    while f"{BB}transformers.{n}.fc1.weight" in sd:
        n += 1
    return n


def backbone(sd, x, num_heads, state=None, dt=np.float32, taps=None):
    """CausalVocosBackbone.forward / forward_chunk (reference decoder.py:248-320).

    state None -> offline (block-causal mask).  state dict -> streaming: keys 'in_proj' (B,6,E),
    'res' list of 4 (B,2,2E), 'kv' list of per-layer (k,v) or None; attention is unmasked over
    cache ++ chunk exactly as forward_chunk does."""
    streaming = state is not None
    st = state if streaming else {}
    x, st_in = causal_conv1d(x, _f(sd, BB + "in_proj.weight", dt), _f(sd, BB + "in_proj.bias", dt), st.get("in_proj"))
    res_caches = []
    rc = st.get("res") or [None] * 4
    for j in (0, 1):
        x, c = resnet_block(sd, f"{BB}prior_net.{j}.", x, rc[j], dt)
        res_caches.append(c)
    if taps is not None:
        taps["prior"] = x.copy()
    kvs = []
    old = st.get("kv") or [None] * num_layers_of(sd)
    for i in range(num_layers_of(sd)):
        x, kv = transformer_layer(sd, f"{BB}transformers.{i}.", x, num_heads, old[i],
                                  "none" if streaming else "block_causal", dt)
        kvs.append(kv)
        if taps is not None and i == 0:
            taps["layer0"] = x.copy()
    if taps is not None:
        taps["layers"] = x.copy()
    for j in (0, 1):
        x, c = resnet_block(sd, f"{BB}post_net.{j}.", x, rc[2 + j], dt)
        res_caches.append(c)
    x = layer_norm(x, _f(sd, BB + "final_norm.weight", dt), _f(sd, BB + "final_norm.bias", dt), 1e-6)
    return x, {"in_proj": st_in, "res": res_caches, "kv": kvs}


# --------------------------------------------------------------------------------------------
# iSTFT head  (reference decoder.py:323-546)
# --------------------------------------------------------------------------------------------
def head_spectrum(sd, x, dt=np.float32):
    """ISTFTHead.forward up to S (reference decoder.py:503-518): p = Linear(x); mag = min(exp(p[:481]),100);
    S = mag * (cos(phi) + i sin(phi)).  Returns complex (B,T,481)."""
    p = x @ _f(sd, AD + "isift.out.weight", dt).T + _f(sd, AD + "isift.out.bias", dt)
    nb = p.shape[-1] // 2
    mag = np.minimum(np.exp(p[..., :nb]), dt(100.0))
    ph = p[..., nb:]
    return (mag * np.cos(ph)) + 1j * (mag * np.sin(ph))


def windowed_frames(sd, S, dt=np.float32):
    """irfft(n=n_fft, norm='backward') along bins, times the window (reference decoder.py:380-381)."""
    n_fft = 2 * (S.shape[-1] - 1)
    fr = np.fft.irfft(S, n=n_fft, axis=-1).astype(dt)
    return fr * _f(sd, AD + "isift.istft.window", dt)


def overlap_add(frames, window, hop):
    """F.fold overlap-add of windowed frames + window-square envelope (reference decoder.py:384-399).
    frames (B,T,n_fft) -> (y_full (B,(T-1)*hop+n_fft), env ((T-1)*hop+n_fft,))."""
    B, T, N = frames.shape
    out = np.zeros((B, (T - 1) * hop + N), dtype=frames.dtype)
    env = np.zeros(((T - 1) * hop + N,), dtype=frames.dtype)
    w2 = (window * window).astype(frames.dtype)
    for t in range(T):
        out[:, t * hop:t * hop + N] += frames[:, t, :]
        env[t * hop:t * hop + N] += w2
    return out, env


def istft_offline(sd, S, hop, dt=np.float32):
    """ISTFT.forward, padding='same' (reference decoder.py:350-405)."""
    fr = windowed_frames(sd, S, dt)
    n_fft = fr.shape[-1]
    pad = (n_fft - hop) // 2
    y, env = overlap_add(fr, _f(sd, AD + "isift.istft.window", dt), hop)
    y, env = y[:, pad:-pad], env[pad:-pad]
    assert (env > 1e-11).all()
    return y / env


def istft_chunk(sd, S, hop, cache, last_chunk, dt=np.float32):
    """ISTFT.forward_chunk (reference decoder.py:407-468).  cache (B,3,n_fft) windowed frames or None."""
    fr = windowed_frames(sd, S, dt)
    n_fft = fr.shape[-1]
    pad = (n_fft - hop) // 2
    first = cache is None
    if not first:
        fr = np.concatenate([cache, fr], axis=1)
    new_cache = fr[:, -(n_fft // hop - 1):, :].copy()
    y, env = overlap_add(fr, _f(sd, AD + "isift.istft.window", dt), hop)
    with np.errstate(invalid="ignore", divide="ignore"):
        y = y / env      # env == 0 only at full-domain sample 0, which is always trimmed (no assert: decoder.py:456)
    y = y[:, pad:] if first else y[:, (n_fft - hop):]
    y = y[:, :-pad] if last_chunk else y[:, :-(n_fft - hop)]
    return y, new_cache


# --------------------------------------------------------------------------------------------
# entry points  (reference model.py:307-376)
# --------------------------------------------------------------------------------------------
def decode(sd, tokens, num_heads, hop=240, dt=np.float32, taps: Optional[dict] = None):
    """RedCodecInfer.decode (reference model.py:307-324): tokens (B,nq,L) int -> audio (B, 8*hop*L)."""
    tokens = np.asarray(tokens)
    emb, z = rvq_decode_codes(sd, tokens, dt)
    x50 = upconv(sd, z, dt)
    x, _ = upsample_conv(sd, x50, None, dt)
    h, _ = backbone(sd, x, num_heads, None, dt, taps)
    S = head_spectrum(sd, h, dt)
    y = istft_offline(sd, S, hop, dt)
    if taps is not None:
        taps.update(emb=emb, z=z, x50=x50, up=x, final=h, spec=S)
    return y


class StreamState:
    """The reference's 5-tensor cache_dict (model.py:346-375) held time-major."""

    def __init__(self):
        self.up = None      # (B,3,E)
        self.bb = None      # dict for backbone()
        self.istft = None   # (B,3,n_fft)
        self.n_tokens = 0

    def to_reference_layout(self, num_heads) -> Dict[str, np.ndarray]:
        """Export as the reference layouts: up_conv_cache (B,E,3), bb_conv_cache1 (B,E,6),
        bb_conv_cache2 (B,8E,2), bb_kv_cache (B,nl,H,T,2hd), is_cache (B,n_fft,3)."""
        B, _, E = self.up.shape
        hd = E // num_heads
        kv = []
        for (k, v) in self.bb["kv"]:
            T = k.shape[1]
            kh = k.reshape(B, T, num_heads, hd).transpose(0, 2, 1, 3)
            vh = v.reshape(B, T, num_heads, hd).transpose(0, 2, 1, 3)
            kv.append(np.concatenate([kh, vh], axis=-1))
        return {
            "up_conv_cache": self.up.transpose(0, 2, 1).copy(),
            "bb_conv_cache1": self.bb["in_proj"].transpose(0, 2, 1).copy(),
            "bb_conv_cache2": np.concatenate([c.transpose(0, 2, 1) for c in self.bb["res"]], axis=1),
            "bb_kv_cache": np.stack(kv, axis=1),
            "is_cache": self.istft.transpose(0, 2, 1).copy(),
        }


def decode_chunk(sd, tokens, state: Optional[StreamState], last: bool, num_heads, hop=240,
                 dt=np.float32) -> Tuple[np.ndarray, StreamState]:
    """RedCodecInfer.decode_one_token (reference model.py:326-376) for a chunk of Lc >= 1 tokens.
    n = 8*hop*Lc - pad*[first] + pad*[last] samples (decoder.py:459-467)."""
    tokens = np.asarray(tokens)
    st = state if state is not None else StreamState()
    _, z = rvq_decode_codes(sd, tokens, dt)
    x50 = upconv(sd, z, dt)
    new = StreamState()
    x, new.up = upsample_conv(sd, x50, st.up, dt)   # cache None on the first chunk (decoder.py:631-655)
    h, new.bb = backbone(sd, x, num_heads, st.bb if st.bb is not None else {}, dt)
    S = head_spectrum(sd, h, dt)
    y, new.istft = istft_chunk(sd, S, hop, st.istft, last, dt)
    new.n_tokens = st.n_tokens + tokens.shape[2]
    return y, new


def resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """FIR bank of torchaudio.functional.resample (sinc_interp_hann), restated from torchaudio 2.11
    functional._get_sinc_resample_kernel as the reference calls it (fireredtts2.py:65,389-391: defaults, the float32
    waveform's dtype is used for the whole computation).  -> (taps (new, 2*width + orig) float32, width, orig, new)
    with orig/new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = np.float32(min(orig, new) * rolloff)
    width = int(math.ceil(lowpass_filter_width * orig / (min(orig, new) * rolloff)))
    f32 = np.float32
    idx = np.arange(-width, width + orig, dtype=f32)[None, :] / f32(orig)
    t = np.arange(0, -new, -1, dtype=f32)[:, None] / f32(new) + idx
    t = t * base
    t = np.clip(t, f32(-lowpass_filter_width), f32(lowpass_filter_width))
    window = np.cos(t * f32(math.pi) / f32(lowpass_filter_width) / f32(2)) ** 2
    t = t * f32(math.pi)
    scale = f32(min(orig, new) * rolloff / orig)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, f32(1.0), np.sin(t) / t)
    k = (k * (window * scale)).astype(f32)
    return k, width, orig, new


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """torchaudio.functional.resample on (..., n) float32 (functional._apply_sinc_resample_kernel): zero-pad
    (width, width + orig), strided correlation with the `new` phase filters, keep ceil(new * n / orig) samples."""
    if orig_freq == new_freq:
        return x
    taps, width, orig, new = resample_taps(orig_freq, new_freq)
    shape = x.shape
    x2 = np.asarray(x, dtype=np.float32).reshape(-1, shape[-1])
    n = x2.shape[1]
    xp = np.pad(x2, ((0, 0), (width, width + orig)))
    K = taps.shape[1]
    frames = (xp.shape[1] - K) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, K, axis=1)[:, ::orig][:, :frames]   # (B, frames, K)
    y = np.einsum("bfk,pk->bfp", win, taps, dtype=np.float32).reshape(x2.shape[0], -1)
    target = int(math.ceil(new * n / orig))
    return y[:, :target].reshape(shape[:-1] + (target,))


def snr_db(ref: np.ndarray, out: np.ndarray) -> float:
    """SNR = 10 log10( sum ref^2 / sum (ref-out)^2 )  (SURVEY.md §8d parity gate)."""
    ref = np.asarray(ref, dtype=np.float64)
    out = np.asarray(out, dtype=np.float64)
    num = float((ref * ref).sum())
    den = float(((ref - out) ** 2).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * math.log10(num / den)
