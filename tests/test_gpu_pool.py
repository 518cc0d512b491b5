"""Slot pool (continuous batching of concurrent decode_one_token streams, C ABI frt2_pool_*) on the GPU: every
stream that passes through a pool slot must match the oracle's decode of its own tokens, whatever its neighbours do."""
import numpy as np
import pytest
import torch

from fireredtts2_b200 import _native as N
from fireredtts2_b200.config import SMALL, TINY
from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
from oracle import codec_oracle as O
from .gpu_common import build_codec, report, to_np

pytestmark = pytest.mark.gpu
SNR_GATE_DB = 40.0   # BASELINE.json north_star: waveform SNR >= 40 dB vs the fp32 reference


def _run_schedule(pool, streams, schedule, pcm16=False):
    """streams: {name: tokens (nq, L)}; schedule: {name: first step}.  One token per open stream per step.
    -> {name: concatenated chunks}, {name: slot}"""
    slots, pos, chunks = {}, {k: 0 for k in streams}, {k: [] for k in streams}
    step = 0
    while any(pos[k] < streams[k].shape[1] for k in streams):
        toks, last = {}, []
        for k, t in streams.items():
            if step < schedule[k] or pos[k] >= t.shape[1]:
                continue
            if k not in slots:
                slots[k] = pool.open()
            toks[slots[k]] = torch.from_numpy(t[:, pos[k]])
            if pos[k] == t.shape[1] - 1:
                last.append(slots[k])
        out = pool.step(toks, last=last, pcm16=pcm16)
        for k in streams:
            if k in slots and slots[k] in out and (step >= schedule[k]) and pos[k] < streams[k].shape[1]:
                chunks[k].append(out[slots[k]].cpu().numpy())
                pos[k] += 1
        step += 1
        assert step < 1000
    return {k: np.concatenate(v) for k, v in chunks.items()}, slots


@pytest.mark.parametrize("cfg,slots", [(TINY, 4), (SMALL, 20)], ids=["tiny_4slots", "small_20slots"])
def test_pool_streams_match_the_oracle(cfg, slots):
    sd = synthetic_state_dict(cfg, 11)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    pool = codec.new_pool(slots)
    rng = np.random.default_rng(3)
    lens = {"a": 9, "b": 4, "c": 1, "d": 6, "e": 12}
    streams = {k: rng.integers(0, cfg.codebook_size, size=(cfg.num_quantizers, n)) for k, n in lens.items()}
    schedule = {"a": 0, "b": 2, "c": 3, "d": 7, "e": 1}      # d starts after b and c ended: re-uses a released slot
    got, used = _run_schedule(pool, streams, schedule)
    assert len(set(used.values())) < len(used), "a released slot should have been re-used"
    for k, t in streams.items():
        ref = O.decode(sd, t[None], cfg.num_heads, cfg.hop_length)[0]
        assert got[k].shape == ref.shape, k                 # 1560 + 1920*(L-2) + 2280 == 1920*L
        _, snr = report(f"pool/{k}", ref, got[k])
        assert snr >= SNR_GATE_DB, k
    assert pool.n_open == 0


def test_pool_slot_is_independent_of_its_neighbours():
    """Bit-identical samples for a stream whether the other slots are idle, busy, starting or ending."""
    cfg = TINY
    sd = synthetic_state_dict(cfg, 12)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    rng = np.random.default_rng(4)
    main = rng.integers(0, cfg.codebook_size, size=(cfg.num_quantizers, 8))
    alone, _ = _run_schedule(codec.new_pool(6), {"m": main}, {"m": 0})
    others = {f"o{i}": rng.integers(0, cfg.codebook_size, size=(cfg.num_quantizers, n))
              for i, n in enumerate((3, 5, 2, 7))}
    pool = codec.new_pool(6)
    crowd, _ = _run_schedule(pool, {"m": main, **others}, {"m": 0, "o0": 0, "o1": 1, "o2": 4, "o3": 2})
    assert np.array_equal(alone["m"], crowd["m"])
    # and a second pass through the SAME pool (slots re-used after LAST, graph already captured) repeats it
    again, _ = _run_schedule(pool, {"m": main, **others}, {"m": 2, "o0": 0, "o1": 0, "o2": 1, "o3": 3})
    assert np.array_equal(alone["m"], again["m"])


def test_pool_equals_decode_one_token_stream():
    """A pool slot and a plain batch-1 stream run the same kernels on one token: same chunk sizes, and samples equal
    within the fp16-operand tolerance (the GEMM tiles differ, so not bit for bit)."""
    cfg = TINY
    sd = synthetic_state_dict(cfg, 13)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    tok = synthetic_tokens(cfg, 1, 6, 21)
    cache, single = {}, []
    for i in range(6):
        a, cache = codec.decode_one_token(torch.from_numpy(tok[:, :, i:i + 1]).cuda(), cache, i == 5)
        single.append(to_np(a)[0])
    pool = codec.new_pool(3)
    s = pool.open()
    for i in range(6):
        out = pool.step({s: torch.from_numpy(tok[0, :, i])}, last=[s] if i == 5 else [])
        assert out[s].shape[0] == single[i].shape[0]
        _, snr = report(f"pool-vs-stream/chunk{i}", single[i], to_np(out[s]))
        assert snr >= 50.0


def test_pool_pcm16_and_graph_vs_eager():
    cfg = TINY
    sd = synthetic_state_dict(cfg, 14)
    rng = np.random.default_rng(5)
    streams = {k: rng.integers(0, cfg.codebook_size, size=(cfg.num_quantizers, n)) for k, n in (("a", 5), ("b", 3))}
    sched = {"a": 0, "b": 1}
    codec = build_codec(cfg, sd, stream_max_tokens=8)
    f32, _ = _run_schedule(codec.new_pool(2), streams, sched)
    pcm, _ = _run_schedule(codec.new_pool(2), streams, sched, pcm16=True)
    for k in streams:
        assert pcm[k].dtype == np.int16
        assert np.array_equal(pcm[k], (f32[k] * 32767).astype(np.int16))      # the reference's wire conversion
    codec.set_debug(N.DBG_NO_GRAPH)
    eager, _ = _run_schedule(codec.new_pool(2), streams, sched)
    for k in streams:
        assert np.array_equal(eager[k], f32[k])                              # graph replay == kernel by kernel


def test_pool_int64_tokens_strides_and_errors():
    cfg = TINY
    sd = synthetic_state_dict(cfg, 15)
    codec = build_codec(cfg, sd, stream_max_tokens=3)
    pool = codec.new_pool(2)
    nq = cfg.num_quantizers
    tok = torch.from_numpy(synthetic_tokens(cfg, 2, 1, 1)[:, :, 0]).cuda()              # (2, nq) int64
    wide = torch.zeros((nq, 4), dtype=torch.int64, device="cuda")
    wide[:, :2] = tok.T
    out64, n = pool.step_dense(wide.T[:2], [N.SLOT_ACTIVE | N.SLOT_RESET, 0])          # strided int64 view
    assert n == [cfg.samples_per_token - cfg.istft_pad, 0]
    assert float(out64[1].abs().max()) == 0.0                                           # idle slot: zeros
    pool2 = codec.new_pool(2)
    out32, _ = pool2.step_dense(tok.int(), [N.SLOT_ACTIVE | N.SLOT_RESET, 0])
    assert torch.equal(out64[0], out32[0])
    with pytest.raises(ValueError):
        pool.step_dense(tok, [N.SLOT_LAST, 0])                                          # LAST without ACTIVE
    pool.step_dense(tok, [N.SLOT_ACTIVE | N.SLOT_LAST, 0])
    with pytest.raises(ValueError):
        pool.step_dense(tok, [N.SLOT_ACTIVE, 0])                                        # ended: needs RESET
    pool.step_dense(tok, [N.SLOT_ACTIVE | N.SLOT_RESET, 0])
    pool.step_dense(tok, [N.SLOT_ACTIVE, 0])
    pool.step_dense(tok, [N.SLOT_ACTIVE, 0])
    with pytest.raises(OverflowError):
        pool.step_dense(tok, [N.SLOT_ACTIVE, 0])                                        # max_tokens = 3
    assert pool.slot_tokens(0) == 3 and pool.slot_tokens(1) == 0
    bad = tok.clone()
    bad[1, 0] = cfg.codebook_size
    pool.step_dense(bad, [0, 0])                                                        # idle slots are not read
    with pytest.raises(IndexError):
        pool.step_dense(bad, [0, N.SLOT_ACTIVE | N.SLOT_RESET])


def test_stream_chunk_pcm16():
    cfg = TINY
    sd = synthetic_state_dict(cfg, 16)
    codec = build_codec(cfg, sd, stream_max_tokens=8)
    tok = torch.from_numpy(synthetic_tokens(cfg, 2, 4, 2)).cuda()
    c32, c16 = {}, {}
    for lc_plan in ([1, 1, 1, 1], [4]):                      # graph path and kernel-by-kernel path
        c32, c16, pos = {}, {}, 0
        for i, lc in enumerate(lc_plan):
            last = i == len(lc_plan) - 1
            a, c32 = codec.decode_one_token(tok[:, :, pos:pos + lc], c32, last)
            p, c16 = codec.decode_one_token(tok[:, :, pos:pos + lc], c16, last, pcm16=True)
            assert p.dtype == torch.int16 and p.shape == a.shape
            assert np.array_equal(p.cpu().numpy(), (to_np(a) * 32767).astype(np.int16))
            pos += lc


def test_decode_stream_front_matches_offline_and_wire_format():
    """codec.decode_stream (the codec half of the reference's generate_stream, fireredtts2.py:259-343): chunks come
    back one frame late, in pinned host memory, as int16 PCM; concatenated they equal the offline decode's PCM up to
    the fp16-operand tolerance, and the fp32 variant passes the SNR gate against the oracle."""
    cfg = TINY
    sd = synthetic_state_dict(cfg, 17)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    tok = synthetic_tokens(cfg, 1, 7, 5)
    frames = [torch.from_numpy(tok[0, :, i]).cuda() for i in range(7)]

    def llm():
        for f in frames:
            yield f

    chunks = list(codec.decode_stream(llm(), pcm16=False))
    assert [c.index for c in chunks] == list(range(7))
    sizes = [c.samples.shape[1] for c in chunks]
    spt, pad = cfg.samples_per_token, cfg.istft_pad
    assert sizes == [spt - pad] + [spt] * 5 + [spt + pad]
    for c in chunks:
        c.ready.synchronize()
        assert c.samples.is_pinned()
    cat = np.concatenate([c.samples.numpy()[0] for c in chunks])
    ref = O.decode(sd, tok, cfg.num_heads, cfg.hop_length)[0]
    _, snr = report("decode_stream/fp32", ref, cat)
    assert snr >= SNR_GATE_DB
    pcm = list(codec.decode_stream(llm(), pcm16=True))
    for c in pcm:
        c.ready.synchronize()
    pcm_cat = np.concatenate([c.samples.numpy()[0] for c in pcm])
    assert pcm_cat.dtype == np.int16
    assert np.array_equal(pcm_cat, (cat * 32767).astype(np.int16))
    # the push/finish form used by a server loop
    from fireredtts2_b200.codec import StreamDecoder
    dec = StreamDecoder(codec, pcm16=False)
    assert dec.push(frames[0]) is None
    c0 = dec.push(frames[1])
    c1 = dec.finish()
    c0.ready.synchronize(); c1.ready.synchronize()
    assert c0.samples.shape[1] == spt - pad and c1.samples.shape[1] == spt + pad
    assert dec.finish() is None
    with pytest.raises(ValueError):
        dec.push(frames[2])


def test_generate_stream_revived_on_the_library():
    """dropin.generate_stream (fireredtts2.py:259-343) on a FireRedTTS2-shaped object whose ``_audio_tokenizer`` is the
    library's codec and whose LM is a stand-in emitting given frames on the device, then an all-zero frame: one chunk per
    frame, one frame late, int16 PCM in pinned host memory — bit-identical to ``decode_stream`` over the same frames, and
    the fp32 form within the gate of the oracle's offline decode."""
    from fireredtts2_b200 import dropin
    cfg = TINY
    sd = synthetic_state_dict(cfg, 17)
    codec = build_codec(cfg, sd, stream_max_tokens=16)
    nq, n = cfg.num_quantizers, 6
    tok = synthetic_tokens(cfg, 1, n, 8)
    tok[0, :, :] = np.maximum(tok[0, :, :], 1)                      # no all-zero frame before the end
    frames = [torch.from_numpy(tok[0, :, i].astype(np.int32)).cuda().unsqueeze(0) for i in range(n)]

    class Model:
        def __init__(self):
            self.n = 0

        def reset_caches(self):
            self.n = 0

        def generate_frame(self, tokens, mask, pos, temperature, topk):
            assert tokens.shape[-1] == nq + 1 and tokens.is_cuda
            self.n += 1
            return frames[self.n - 1] if self.n <= n else torch.zeros(1, nq, dtype=torch.int32, device="cuda")

    class TTS:
        device = "cuda:0"

        def __init__(self):
            self._model, self._audio_tokenizer = Model(), codec

        def _tokenize_segment(self, seg):
            return torch.zeros(3, nq + 1), torch.ones(3, nq + 1)

        def _tokenize_text_segment(self, text, speaker):
            return torch.zeros(2, nq + 1), torch.ones(2, nq + 1)

    tts = TTS()
    spt, pad = cfg.samples_per_token, cfg.istft_pad
    got = {}
    for pcm16 in (True, False):
        chunks = list(dropin.generate_stream(tts, "hi", "[S1]", [None], max_audio_length_ms=1600, pcm16=pcm16))
        assert tts._model.n == n + 1 and [c.index for c in chunks] == list(range(n))
        assert [c.samples.shape[1] for c in chunks] == [spt - pad] + [spt] * (n - 2) + [spt + pad]
        for c in chunks:
            c.ready.synchronize()
            assert c.samples.is_pinned()
        got[pcm16] = np.concatenate([c.samples.numpy()[0].copy() for c in chunks])
        ref_chunks = list(codec.decode_stream((f[0] for f in frames), pcm16=pcm16))
        for c in ref_chunks:
            c.ready.synchronize()
        assert np.array_equal(got[pcm16], np.concatenate([c.samples.numpy()[0] for c in ref_chunks]))
    assert got[True].dtype == np.int16 and np.array_equal(got[True], (got[False] * 32767).astype(np.int16))
    ref = O.decode(sd, tok, cfg.num_heads, cfg.hop_length)[0]
    _, snr = report("generate_stream/fp32 vs the oracle's offline decode", ref, got[False])
    assert snr >= SNR_GATE_DB
    # the generation budget ends the stream too: 3 frames at 240 ms, the third flushed with last_token
    tts._model = Model()
    short = list(dropin.generate_stream(tts, "hi", "[S1]", [], max_audio_length_ms=240))
    assert [c.samples.shape[1] for c in short] == [spt - pad, spt, spt + pad]
