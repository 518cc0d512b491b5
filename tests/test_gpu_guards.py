"""Own bounds checks for every output the library writes (compute-sanitizer is not available on the GPU pool).

Every output buffer of the Python wrappers comes from ``torch.empty``; for the duration of the test that call hands out the
middle of a larger allocation whose 2 KiB head and tail are filled with a sentinel.  After a pass over every entry point —
offline / ragged / int16 / resampled decode, the token step (first, middle, last chunk; batch 2), state export, a slot pool,
RVQ gather and encode, taps, the encode features and the LM's frame tail — the sentinels must be untouched and every result
must equal, bit for bit, the one the same call produced into an ordinary allocation."""
import numpy as np
import pytest
import torch

from fireredtts2_b200.codec import resample
from fireredtts2_b200.config import C0, SMALL, TINY
from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict, synthetic_tokens
from .gpu_common import build_codec

pytestmark = pytest.mark.gpu

PAD_BYTES = 2048


class GuardedEmpty:
    """Stand-in for ``torch.empty``: CUDA requests get sentinel-filled guard zones on both sides."""

    SENTINEL = {torch.float32: 777.25, torch.int16: 0x5A5A, torch.int32: 0x5A5A5A5A, torch.int64: 0x5A5A5A5A5A5A5A5A}

    def __init__(self, cuda_only=True):
        self.real, self.bufs, self.cuda_only = torch.empty, [], cuda_only

    def __call__(self, *size, dtype=None, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dtype = dtype or torch.float32
        dev = torch.device(device) if device is not None else torch.device("cpu")
        n = int(np.prod(shape)) if shape else 1
        if (self.cuda_only and dev.type != "cuda") or dtype not in self.SENTINEL or n == 0 or kw:
            return self.real(*size, dtype=dtype, device=device, **kw)
        pad = PAD_BYTES // dtype.itemsize
        raw = torch.full((pad + n + pad,), self.SENTINEL[dtype], dtype=dtype, device=dev)
        self.bufs.append((raw, pad, n))
        return raw[pad:pad + n].view(shape)

    def intact(self):
        bad = []
        for i, (raw, pad, n) in enumerate(self.bufs):
            s = self.SENTINEL[raw.dtype]
            if not (bool((raw[:pad] == s).all()) and bool((raw[pad + n:] == s).all())):
                bad.append((i, str(raw.dtype), n))
        return bad


def _pass(codec, cfg, enc=None, fd=None, L=9):
    """One call of every entry point; returns the results in a fixed order."""
    out = []
    tok = torch.from_numpy(synthetic_tokens(cfg, 3, L, 1)).cuda()
    lens = torch.tensor([L, L // 2, 1], dtype=torch.int32)
    out.append(codec.decode(tok))
    out.append(codec.decode(tok.to(torch.int32).permute(0, 2, 1).contiguous().permute(0, 2, 1), lengths=lens))
    out.append(codec.decode(tok, pcm16=True))
    a24, a16 = codec.decode_resampled(tok, 16000, lengths=lens)
    out += [a24, a16]
    out.append(resample(a24, 24000, 16000))
    cache = {}
    for i in range(4):                                     # batch 2: first, middle, middle, last chunk
        a, cache = codec.decode_one_token(tok[:2, :, i:i + 1], cache, i == 3)
        out.append(a)
        if i == 2:
            out += list(codec.export_cache(cache).values())
    cache = {}
    for i in range(2):                                     # int16 chunks: first, last
        a, cache = codec.decode_one_token(tok[2:, :, i:i + 1], cache, i == 1, pcm16=True)
        out.append(a)
    pool = codec.new_pool(3, 8)
    s0, s1 = pool.open(), pool.open()
    r = pool.step({s0: tok[0, :, 0], s1: tok[1, :, 0]})
    out += [r[s0], r[s1]]
    r = pool.step({s0: tok[0, :, 1], s1: tok[1, :, 1]}, last=(s1,), pcm16=True)
    out += [r[s0], r[s1]]
    pool.destroy()
    rows, ssum = codec.rvq_gather(tok)
    out += [rows, ssum]
    g = torch.Generator().manual_seed(3)
    out.append(codec.rvq_encode_codes(torch.randn(2, cfg.embed_dim, 11, generator=g).cuda()))
    if enc is not None:
        e, (ssl, aco) = enc
        out.append(e.features(ssl, aco))
    if fd is not None:
        f, (last_h, noise) = fd
        codes, logits = f.generate_codes(last_h, 8, 0.9, noise=noise, return_logits=True)
        out += [codes, logits, f.generate_codes(last_h[:1], 8, 0.9, noise=noise[:1])]
        f.check_error()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("preset", ["TINY", "SMALL", "C0"])
def test_no_entry_point_writes_outside_its_output(preset, monkeypatch):
    from fireredtts2_b200.encoder import ETINY, CodecEncoderB200, synthetic_encoder_state_dict, synthetic_features
    from fireredtts2_b200.frame_decoder import (FD_TINY, FrameDecoderB200, synthetic_frame_decoder_state_dict,
                                                synthetic_frame_inputs)
    cfg = {"TINY": TINY, "SMALL": SMALL, "C0": C0}[preset]
    L = 70 if preset == "C0" else 9       # C0: 560 rows per item, the CTA-pair GEMM tiles and several attention tiles per head
    sd = dict(synthetic_state_dict(cfg, 4))
    sd.update(synthetic_encode_tensors(cfg, 4))
    codec = build_codec(cfg, sd, stream_max_tokens=8)
    enc = fd = None
    if preset == "TINY":                                   # the encode features and the frame tail once are enough
        ssl, aco = synthetic_features(ETINY, 2, 48, 2)
        enc = (CodecEncoderB200(ETINY, synthetic_encoder_state_dict(ETINY, 1), device="cuda:0"),
               (torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda()))
        last_h, noise = synthetic_frame_inputs(FD_TINY, 2, 4)
        fd = (FrameDecoderB200(FD_TINY, synthetic_frame_decoder_state_dict(FD_TINY, 3), device="cuda:0"),
              (torch.from_numpy(last_h).cuda(), torch.from_numpy(noise).cuda()))
    plain = [t.clone() for t in _pass(codec, cfg, enc, fd, L)]
    guard = GuardedEmpty()
    monkeypatch.setattr(torch, "empty", guard)
    try:
        guarded = _pass(codec, cfg, enc, fd, L)
    finally:
        monkeypatch.undo()
    assert len(guard.bufs) >= 20, (len(guard.bufs), len(plain))
    del codec       # the outputs did come from the guard
    assert guard.intact() == []
    assert len(guarded) == len(plain)
    for i, (a, b) in enumerate(zip(guarded, plain)):
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b), i
    print(f"[guards] {preset}: {len(guard.bufs)} guarded outputs, {sum(n for _, _, n in guard.bufs)} elements, sentinels intact")
