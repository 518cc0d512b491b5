"""Concurrency, life-cycle and error semantics of the streaming states (GPU).

* The reference's own call pattern — ``decode_one_token(tok, {}, last)`` with a fresh cache_dict per utterance
  (codec/model.py:346) — must be as cheap as a pooled state: the handle recycles states, resets are stream-ordered.
* Streams that decode concurrently on different CUDA streams (StreamDecoder runs every codec step on its own side stream)
  and an offline decode on the caller's stream must not disturb each other: each result is bit-identical to its solo run.
* An out-of-range code raises IndexError to the request that sent it — per stream, per pool slot — and nobody else.
"""
import numpy as np
import pytest
import torch

from fireredtts2_b200.codec import StreamDecoder, StreamPoolIndexError
from fireredtts2_b200.config import SMALL, TINY
from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
from oracle import codec_oracle as O
from .gpu_common import build_codec, report, to_np

pytestmark = pytest.mark.gpu


def _stream_all(dec, frames):
    out = []
    for f in frames:
        c = dec.push(f)
        if c is not None:
            out.append(c)
    out.append(dec.finish())
    return out


def _cat(chunks):
    for c in chunks:
        c.ready.synchronize()
    return np.concatenate([c.samples.numpy()[0].copy() for c in chunks])


def test_two_stream_decoders_and_an_offline_decode_interleaved_are_bit_identical_to_solo_runs():
    cfg = SMALL
    sd = synthetic_state_dict(cfg, 21)
    codec = build_codec(cfg, sd, stream_max_tokens=40)
    n = 24
    ta, tb = synthetic_tokens(cfg, 1, n, 1), synthetic_tokens(cfg, 1, n, 2)
    toff = torch.from_numpy(synthetic_tokens(cfg, 3, 50, 3)).cuda()
    fa = [torch.from_numpy(ta[0, :, i]).cuda() for i in range(n)]
    fb = [torch.from_numpy(tb[0, :, i]).cuda() for i in range(n)]
    solo_a = _cat(_stream_all(StreamDecoder(codec, pcm16=False, ring=n + 2), fa))
    solo_b = _cat(_stream_all(StreamDecoder(codec, pcm16=False, ring=n + 2), fb))
    solo_off = codec.decode(toff).clone()
    torch.cuda.synchronize()
    for rep in range(3):                      # nothing between the calls synchronises the device
        da, db = StreamDecoder(codec, pcm16=False, ring=n + 2), StreamDecoder(codec, pcm16=False, ring=n + 2)
        ca, cb, offs = [], [], []
        for i in range(n):
            x = da.push(fa[i])
            y = db.push(fb[i])
            if x is not None:
                ca.append(x)
            if y is not None:
                cb.append(y)
            if i % 5 == 2:
                offs.append(codec.decode(toff))
        ca.append(da.finish())
        cb.append(db.finish())
        assert np.array_equal(_cat(ca), solo_a), rep
        assert np.array_equal(_cat(cb), solo_b), rep
        for o in offs:
            assert torch.equal(o, solo_off), rep
    ref = O.decode(sd, ta, cfg.num_heads, cfg.hop_length)[0]
    _, snr = report("interleaved/stream-a vs oracle", ref, solo_a)
    assert snr >= 40.0


def test_reference_call_pattern_recycles_states():
    """decode_one_token(tok, {}, last) utterance after utterance: the state of a finished utterance goes back to the
    handle and the next {} call gets it (same native object), reset on the caller's stream; every utterance decodes the
    same samples as the first time, on the default stream and on a side stream without any synchronisation between."""
    cfg = TINY
    sd = synthetic_state_dict(cfg, 22)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    codec.reserve_streams(2)
    tok = torch.from_numpy(synthetic_tokens(cfg, 1, 6, 9)).cuda()

    def utterance():
        cache, out = {}, []
        for i in range(6):
            a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, i == 5)
            out.append(a)
        ptr = cache["frt2_state"].ptr.value
        del cache
        return torch.cat(out, dim=1), ptr

    first, p0 = utterance()
    side = torch.cuda.Stream()
    ptrs = {p0}
    for rep in range(6):
        if rep % 2:
            with torch.cuda.stream(side):
                got, p = utterance()
            side.synchronize()
        else:
            got, p = utterance()
        ptrs.add(p)
        assert torch.equal(got, first), rep
    assert len(ptrs) <= 2, "states must be recycled, not re-allocated"
    ref = O.decode(sd, tok.cpu().numpy(), cfg.num_heads, cfg.hop_length)
    _, snr = report("recycled-state utterance vs oracle", ref, to_np(first))
    assert snr >= 40.0


def test_reset_is_stream_ordered():
    """reset_stream followed at once by a StreamDecoder-style decode on a side stream, while the previous utterance's
    kernels may still be running on another stream: same samples as a fresh state."""
    cfg = TINY
    sd = synthetic_state_dict(cfg, 23)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    tok = torch.from_numpy(synthetic_tokens(cfg, 1, 8, 3)).cuda()
    state = codec.new_stream(1)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(stream):
        cache, out = state, []
        with torch.cuda.stream(stream):
            for i in range(8):
                a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, i == 7, _check=False)
                out.append(a)
            res = torch.cat(out, dim=1)
        return res

    torch.cuda.synchronize()
    first = run(s1)
    outs = []
    for rep in range(4):
        codec.reset_stream(state)
        outs.append(run(s2 if rep % 2 == 0 else s1))
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, first)


def test_index_error_goes_to_the_stream_that_sent_it():
    cfg = TINY
    sd = synthetic_state_dict(cfg, 24)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    good = torch.from_numpy(synthetic_tokens(cfg, 1, 4, 1)).cuda()
    bad = good.clone()
    bad[0, 1, 2] = cfg.codebook_size
    # decode_one_token: IndexError inside the offending call (reference rvq.py:58), the other stream never sees it
    ca, cb = {}, {}
    for i in range(4):
        _, ca = codec.decode_one_token(good[:, :, i:i + 1], ca, i == 3)
        if i == 2:
            with pytest.raises(IndexError):
                codec.decode_one_token(bad[:, :, i:i + 1], cb, False)
            break
        _, cb = codec.decode_one_token(bad[:, :, i:i + 1], cb, False)
    _, ca2 = codec.decode_one_token(good[:, :, :1], {}, False)     # healthy streams keep working, offline decode too
    a = codec.decode(good)
    assert torch.isfinite(a).all()
    # StreamDecoder: the words ride along with the chunks; the error surfaces at a later push / finish of THAT decoder
    d_good, d_bad = StreamDecoder(codec, pcm16=False), StreamDecoder(codec, pcm16=False)
    raised = False
    try:
        for i in range(4):
            d_good.push(good[0, :, i])
            d_bad.push(bad[0, :, i])
        d_bad.finish()
    except IndexError:
        raised = True
    assert raised
    last = d_good.finish()
    last.ready.synchronize()
    assert last.samples.shape[1] == cfg.samples_per_token + cfg.istft_pad


def test_pool_reports_bad_codes_per_slot_and_keeps_the_other_slots():
    cfg = TINY
    sd = synthetic_state_dict(cfg, 25)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    rng = np.random.default_rng(6)
    ta = rng.integers(0, cfg.codebook_size, size=(cfg.num_quantizers, 5))
    tb = rng.integers(0, cfg.codebook_size, size=(cfg.num_quantizers, 5))
    # reference run of stream a alone
    solo_pool = codec.new_pool(3)
    s = solo_pool.open()
    solo = [to_np(solo_pool.step({s: torch.from_numpy(ta[:, i])}, last=[s] if i == 4 else [])[s]) for i in range(5)]
    pool = codec.new_pool(3)
    a, b = pool.open(), pool.open()
    got = []
    for i in range(5):
        tok_b = tb[:, i].copy()
        if i == 2:
            tok_b[0] = -7
        toks = {a: torch.from_numpy(ta[:, i])}
        if b is not None:
            toks[b] = torch.from_numpy(tok_b)
        try:
            out = pool.step(toks, last=[a] if i == 4 else [])
        except StreamPoolIndexError as e:
            assert i == 2 and e.slots == [b]
            out = e.results                      # the healthy slot's chunk is still delivered
            b = None                             # the offending stream has been closed
        got.append(to_np(out[a]))
    for x, y in zip(got, solo):
        assert np.array_equal(x, y)              # stream a never noticed
    assert pool.n_open == 0
    c = pool.open()                              # the slot is usable again
    out = pool.step({c: torch.from_numpy(tb[:, 0])})
    assert out[c].shape[0] == cfg.samples_per_token - cfg.istft_pad
