"""CPU restatement (numpy) of the codec ENCODE side behind the feature encoders — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s checker legs may import this module; the product path
(``fireredtts2_b200/``) never does.  Pinned against outputs of the unmodified reference (``oracle/make_golden_encoder.py``
-> ``tests/golden/enc_*.npz``; ``tests/test_oracle_golden.py``).

Time-major ``(B, T, C)`` like ``codec_oracle``; each function cites the reference lines it restates.
"""
from __future__ import annotations

import numpy as np

from . import codec_oracle as O


def ssl_adaptor(sd, x, num_heads, dt=np.float32):
    """SslAdaptor.forward (reference codec/model.py:53-66): in_proj, WhisperEncoderLayer x N under a non-pad mask
    (every item full length -> no key is masked, model.py:220-222), LayerNorm (eps 1e-5), out_proj."""
    p = "ssl_adaptor."
    x = x.astype(dt) @ O._f(sd, p + "in_proj.weight", dt).T + O._f(sd, p + "in_proj.bias", dt)
    i = 0
    while f"{p}layers.{i}.fc1.weight" in sd:
        x, _ = O.transformer_layer(sd, f"{p}layers.{i}.", x, num_heads, None, "none", dt)
        i += 1
    x = O.layer_norm(x, O._f(sd, p + "layer_norm.weight", dt), O._f(sd, p + "layer_norm.bias", dt), 1e-5)
    return x @ O._f(sd, p + "out_proj.weight", dt).T + O._f(sd, p + "out_proj.bias", dt)


def residual_down_conv(sd, x, pooler, dt=np.float32):
    """ResidualDownConv.forward (reference codec/model.py:106-121).  x (B, T, D), T a multiple of ``pooler``.
    Conv1d(D, pooler*D, k = s = pooler, bias=False) on (B, D, T): out[o, t'] = sum_{c,k} W[o,c,k] x[pooler*t'+k, c]."""
    p = "downsample."
    B, T, D = x.shape
    x = x.astype(dt)
    xr = x.reshape(B, T // pooler, pooler, D)                       # [b, t', k, c]
    Wg = O._f(sd, p + "gate_proj.weight", dt)                       # (P, D, k)
    Wu = O._f(sd, p + "up_proj.weight", dt)
    g = np.einsum("btkc,ock->bto", xr, Wg).astype(dt)
    u = np.einsum("btkc,ock->bto", xr, Wu).astype(dt)
    xres = x.reshape(B, T // pooler, pooler * D)                    # model.py:113
    c = (O.silu(g) * u) @ O._f(sd, p + "down_proj.weight", dt).T    # model.py:116
    res = O.layer_norm(c + xres, O._f(sd, p + "layer_norm.weight", dt), O._f(sd, p + "layer_norm.bias", dt), 1e-5)
    return res @ O._f(sd, p + "out_proj.weight", dt).T + O._f(sd, p + "out_proj.bias", dt)


def encode_features(sd, ssl, aco, num_heads, pooler, dt=np.float32):
    """model.py:225-232: vq_in_feats = downsample(cat([ssl_adaptor(ssl), aco], dim=2)) -> (B, T/pooler, D)."""
    sem = ssl_adaptor(sd, ssl, num_heads, dt)
    return residual_down_conv(sd, np.concatenate([sem, aco.astype(dt)], axis=2), pooler, dt)
