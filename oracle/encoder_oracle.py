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


# ---- feature encoders: log-mel front end + WhisperEncoder (reference codec/whisper.py:195-302, codec/audio.py) ----
def _hz_to_mel_slaney(f):
    """hertz_to_mel(mel_scale="slaney") (reference codec/audio.py:24-48)."""
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)


def _mel_to_hz_slaney(m):
    """mel_to_hertz(mel_scale="slaney") (reference codec/audio.py:51-75)."""
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)


def mel_filter_bank(bins: int, n_mels: int, fmin=0.0, fmax=8000.0, sr=16000):
    """mel_filter_bank(norm="slaney", mel_scale="slaney") (reference codec/audio.py:102-148) -> (bins, n_mels) float64."""
    ff = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fft = np.linspace(0, sr // 2, bins)
    d = np.diff(ff)
    slopes = ff[None, :] - fft[:, None]
    down = -slopes[:, :-2] / d[:-1]
    up = slopes[:, 2:] / d[1:]
    bank = np.maximum(0.0, np.minimum(down, up))
    return bank * (2.0 / (ff[2:n_mels + 2] - ff[:n_mels]))[None, :]


def log_mel(audio, n_mels=128, n_fft=400, hop=160, dt=np.float32):
    """WhisperMelExtractor.extract_fbank + transpose (reference whisper.py:276-302): audio (B, n) -> (B, n // hop, n_mels).
    torch.stft defaults: center=True, pad_mode="reflect", periodic Hann window, onesided; the last frame is dropped."""
    audio = np.asarray(audio, dtype=np.float64)
    B, n = audio.shape
    pad = n_fft // 2
    x = np.pad(audio, ((0, 0), (pad, pad)), mode="reflect")
    T = n // hop
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)
    frames = x[:, idx] * w                                         # (B, T, n_fft)
    power = np.abs(np.fft.rfft(frames, axis=-1)) ** 2              # (B, T, bins)
    bank = mel_filter_bank(n_fft // 2 + 1, n_mels).astype(np.float32).astype(np.float64)
    mel = power @ bank
    lg = np.log10(np.maximum(mel, 1e-10))
    mx = lg.max(axis=(1, 2), keepdims=True)
    lg = np.maximum(lg, mx - 8.0)
    return ((lg + 4.0) / 4.0).astype(dt)


def whisper_encoder(sd, prefix, mel, num_heads, dt=np.float32):
    """WhisperEncoder.forward (reference whisper.py:222-251): conv1 k3 p1 + GELU, conv2 k3 s2 p1 + GELU, + positions,
    N layers under an all-true mask, LayerNorm.  mel (B, Tm, C) -> (B, Tm // 2, E)."""
    x = mel.astype(dt)
    B, Tm, C = x.shape
    W1, b1 = O._f(sd, prefix + "conv1.weight", dt), O._f(sd, prefix + "conv1.bias", dt)     # (E, C, 3)
    xp = np.pad(x, ((0, 0), (1, 1), (0, 0)))
    h = sum(xp[:, j:j + Tm, :] @ W1[:, :, j].T for j in range(3)) + b1
    h = O.gelu(h.astype(dt))
    W2, b2 = O._f(sd, prefix + "conv2.weight", dt), O._f(sd, prefix + "conv2.bias", dt)     # (E, E, 3), stride 2
    hp = np.pad(h, ((0, 0), (1, 1), (0, 0)))
    T = (Tm + 2 - 3) // 2 + 1
    y = sum(hp[:, j:j + 2 * T:2, :][:, :T] @ W2[:, :, j].T for j in range(3)) + b2
    y = O.gelu(y.astype(dt)) + O._f(sd, prefix + "embed_positions.weight", dt)[:T]
    i = 0
    while f"{prefix}layers.{i}.fc1.weight" in sd:
        y, _ = O.transformer_layer(sd, f"{prefix}layers.{i}.", y, num_heads, None, "none", dt)
        i += 1
    return O.layer_norm(y, O._f(sd, prefix + "layer_norm.weight", dt), O._f(sd, prefix + "layer_norm.bias", dt), 1e-5)


def encode_audio_features(sd, audio, cfg, dt=np.float32, taps=None):
    """RedCodecInfer._encode_one_batch up to the RVQ input (reference model.py:218-232).  cfg: EncoderConfig-like."""
    mel = log_mel(audio, cfg.num_mels, dt=dt)
    ssl = whisper_encoder(sd, "ssl.", mel, cfg.ssl_enc_heads, dt)
    aco = whisper_encoder(sd, "acoustic_encoder.", mel, cfg.aco_heads, dt)
    if taps is not None:
        taps.update(mel=mel, ssl=ssl, aco=aco)
    return encode_features(sd, ssl, aco, cfg.ssl_num_heads, cfg.avg_pooler, dt)
