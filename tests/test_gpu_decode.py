"""End-to-end parity (GPU): RedCodecB200.decode / decode_one_token through the C ABI against the golden
vectors of the real reference and against the oracle on the same seeded inputs.

Gates (BASELINE.json north_star): RVQ indices / gathered rows bit-exact; waveform SNR >= 40 dB and the
max-abs error reported (tolerance MAXABS_TOL relative to the waveform peak)."""
import numpy as np
import pytest
import torch

from fireredtts2_b200 import _native as N
from oracle import codec_oracle as O
from tests.gpu_common import build_codec, report, to_np
from tests.helpers import cases, load_case

pytestmark = pytest.mark.gpu

SNR_GATE_DB = 40.0
MAXABS_TOL = 0.05   # of the reference waveform peak

MODES = {"simt_gemm+warp_attn": N.DBG_GEMM_REF | N.DBG_ATTN_WARP, "tc_gemm+warp_attn": N.DBG_ATTN_WARP, "product": 0,
         "no_graph": N.DBG_NO_GRAPH, "no_skinny": N.DBG_NO_SKINNY}


def _gate(name, ref, out):
    maxabs, snr = report(name, ref, out)
    assert np.isfinite(out).all(), name
    assert snr >= SNR_GATE_DB, f"{name}: SNR {snr:.1f} dB < {SNR_GATE_DB}"
    assert maxabs <= MAXABS_TOL * np.abs(ref).max(), f"{name}: max-abs {maxabs}"


@pytest.mark.parametrize("mode", ["simt_gemm+warp_attn", "tc_gemm+warp_attn", "product"])
@pytest.mark.parametrize("case", [c for c in cases("offline") + cases("reference_init")
                                  if c["preset"] not in ("C0", "ADV4")], ids=lambda c: c["name"])
def test_offline_decode_vs_reference_golden(case, mode):
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    codec.set_debug(MODES[mode] | N.DBG_TAPS)
    tok = torch.from_numpy(g["tokens"]).cuda()
    if case.get("idx") == "int32_permuted":   # production form (fireredtts2.py:196)
        tok = torch.from_numpy(np.ascontiguousarray(g["tokens"].transpose(2, 0, 1)).astype(np.int32)).cuda().permute(1, 2, 0)
        assert not tok.is_contiguous()
    audio = to_np(codec.decode(tok))
    assert audio.shape == g["audio"].shape
    B, L = case["B"], case["L"]
    E = cfg.embed_dim
    # stage-by-stage report against the oracle (same weights), then the gate on the waveform
    taps = {}
    O.decode(sd, g["tokens"], cfg.num_heads, cfg.hop_length, taps=taps)
    for name, shape in (("emb", (B, L, cfg.rvq_dim)), ("z", (B, L, E)), ("x50", (B, 4 * L, E)), ("up", (B, 8 * L, E)),
                        ("prior", (B, 8 * L, E)), ("layer0", (B, 8 * L, E)), ("layers", (B, 8 * L, E)),
                        ("final", (B, 8 * L, E))):
        got = to_np(codec.get_tap(name, shape))
        _, snr = report(f"{case['name']}/{mode}/{name}", taps[name], got)
        assert snr > 45.0, f"stage {name}: {snr:.1f} dB"
    _gate(f"{case['name']}/{mode}/audio", g["audio"], audio)


@pytest.mark.parametrize("case", cases("offline") + cases("reference_init"), ids=lambda c: c["name"])
def test_folded_layernorm_vs_reference_golden(case):
    """Product offline path (LayerNorm folded across the GEMMs: producer epilogue emits the fp16 residual copy + row
    partials, consumer epilogue finishes the normalisation) and the same decode with separate LayerNorm kernels, both
    against the reference's golden waveform."""
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    folded = to_np(codec.decode(tok))
    codec.set_debug(N.DBG_NO_LNFOLD)
    plain = to_np(codec.decode(tok))
    _gate(f"{case['name']}/ln-folded", g["audio"], folded)
    _gate(f"{case['name']}/ln-kernels", g["audio"], plain)
    _, snr = report(f"{case['name']}/folded-vs-kernels", plain, folded)
    assert snr >= (45.0 if case.get("weights") == "adversarial" else 50.0)


def test_adversarial_weights_keep_the_gate():
    """Weights that stress fp16 operands and the folded LayerNorm the way a trained checkpoint can (LayerNorm gamma in
    [0.1, 5], residual rows at ~100 +- 2, four outlier channels x 100, GELU activations in the thousands; 4 layers at the
    C0 widths; weights.adversarial_state_dict) against the real reference's waveform on the same weights.  Reported: the
    product path (fp16 copy rounded after subtracting the row mean), the same path with the plain fp16(x) copy it
    replaces (FRT2_NO_LNSHIFT=1 is read once per process, so that variant is only reported by tools/adv_report.py), and
    separate LayerNorm kernels."""
    case = [c for c in cases("offline") if c["name"] == "adv4_offline"][0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    folded = to_np(codec.decode(tok))
    _, snr_f = report("adv4/product (folded LayerNorm, mean-shifted copy)", g["audio"], folded)
    codec.set_debug(N.DBG_NO_LNFOLD)
    plain = to_np(codec.decode(tok))
    _, snr_p = report("adv4/separate LayerNorm kernels", g["audio"], plain)
    assert snr_f >= SNR_GATE_DB and snr_p >= SNR_GATE_DB
    assert snr_f >= snr_p - 3.0, "the folded form must not cost more than 3 dB against separate LayerNorm kernels"
    # batch rows are independent here too
    codec.set_debug(0)
    one = to_np(codec.decode(tok[1:2]))
    assert np.array_equal(one[0], folded[1])


@pytest.mark.parametrize("case", cases("rvq_emb"), ids=lambda c: c["name"])
def test_c1_rvq_sum_bit_exact_vs_reference(case):
    """C1 = C0 with Identity out_project: gathered rows and the index-ordered fp32 sum at the C0 dimensions
    (16 codebooks x 2048 x 256), bit for bit against the REAL reference's tensor (captured in front of rvq.output_proj)."""
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    for dtype in (torch.int64, torch.int32):
        tok = torch.from_numpy(g["tokens"]).cuda().to(dtype)
        rows, s = codec.rvq_gather(tok)
        assert np.array_equal(to_np(s), g["emb"])
        assert np.array_equal(to_np(rows), O.rvq_gather(sd, g["tokens"]))
    codec.set_debug(N.DBG_TAPS)
    codec.decode(torch.from_numpy(g["tokens"]).cuda())
    B, L = case["B"], case["L"]
    assert np.array_equal(to_np(codec.get_tap("emb", (B, L, cfg.rvq_dim))), g["emb"])     # the decode path's own sum
    _, snr = report("c1/z", g["z"], to_np(codec.get_tap("z", (B, L, cfg.embed_dim))))
    assert snr > 45.0


def test_c0_varlen_distinct_items_multi_wave():
    """14 DISTINCT C0 items of different lengths in one padded batch (every GEMM of the step runs several waves of
    tiles): every item must equal the standalone decode of its own first L_b tokens bit for bit — rows never see their
    neighbours — and two of them are checked against the oracle."""
    from fireredtts2_b200.config import C0
    from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
    cfg = C0
    sd = synthetic_state_dict(cfg, 0)
    codec = build_codec(cfg, sd)
    rng = np.random.default_rng(5)
    Bn, L = 14, 170
    lens = [int(x) for x in rng.integers(60, L + 1, size=Bn)]
    lens[3] = L
    tok_np = synthetic_tokens(cfg, Bn, L, 4321)
    tok = torch.from_numpy(tok_np).cuda()
    spt = cfg.samples_per_token
    for flags in (0, N.DBG_NO_LNFOLD):
        codec.set_debug(flags)
        batch = codec.decode(tok, lengths=torch.tensor(lens, dtype=torch.int32))
        worst = 0.0
        for b in range(Bn):
            single = codec.decode(tok[b:b + 1, :, :lens[b]])
            worst = max(worst, float((batch[b, :lens[b] * spt] - single[0]).abs().max()))
            if lens[b] < L:
                assert float(batch[b, lens[b] * spt:].abs().max()) == 0.0
        print(f"[parity] c0 14 distinct var-len items, debug {flags}: max |item - standalone| = {worst:.3e}")
        assert worst == 0.0
    codec.set_debug(0)
    batch = to_np(codec.decode(tok, lengths=torch.tensor(lens, dtype=torch.int32)))
    for b in (0, 9):
        ref = O.decode(sd, tok_np[b:b + 1, :, :lens[b]], cfg.num_heads, cfg.hop_length)
        _gate(f"c0-varlen/item{b}", ref, batch[b:b + 1, :lens[b] * spt])


def test_c0_offline_vs_reference_golden():
    case = [c for c in cases("offline") if c["preset"] == "C0"][0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    audio = to_np(codec.decode(torch.from_numpy(g["tokens"]).cuda()))
    _gate("c0_L25/product/audio", g["audio"], audio)
    # batch row == single decode, and int32 strided view == int64 contiguous (bit-identical: same kernels)
    tok = torch.from_numpy(g["tokens"]).cuda()
    tok3 = tok.repeat(3, 1, 1)
    a3 = to_np(codec.decode(tok3))
    _gate("c0_L25/product/batch-row", g["audio"], a3[2:3])
    t32 = tok.to(torch.int32).permute(2, 0, 1).contiguous().permute(1, 2, 0)
    assert np.array_equal(to_np(codec.decode(t32)), audio)


def test_rvq_gather_bit_exact():
    case = [c for c in cases("offline") if c["name"] == "tiny_ident_offline"][0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    for dtype in (torch.int64, torch.int32):
        tok = torch.from_numpy(g["tokens"]).cuda().to(dtype)
        rows, s = codec.rvq_gather(tok)
        ref_rows = O.rvq_gather(sd, g["tokens"])
        ref_emb, _ = O.rvq_decode_codes(sd, g["tokens"])
        assert np.array_equal(to_np(rows), ref_rows)          # gathered embeddings: bit-exact
        assert np.array_equal(to_np(s), ref_emb)              # index-ordered fp32 sum: bit-exact
    # the decode path's own sum (Identity projection config) is the same kernel
    codec.set_debug(N.DBG_TAPS)
    codec.decode(torch.from_numpy(g["tokens"]).cuda())
    emb = to_np(codec.get_tap("emb", ref_emb.shape))
    assert np.array_equal(emb, ref_emb)


def test_index_errors_and_prefix():
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    bad = tok.clone()
    bad[0, 1, 2] = cfg.codebook_size
    with pytest.raises(IndexError):
        codec.decode(bad)
    bad[0, 1, 2] = -1
    with pytest.raises(IndexError):
        codec.decode(bad)
    with pytest.raises(TypeError):
        codec.decode(tok.float())
    with pytest.raises(ValueError):
        codec.decode(tok[0])
    # quantizers[:nq]: fewer codebooks than configured (rvq.py:160)
    a = to_np(codec.decode(tok[:, :2, :]))
    ref = O.decode(sd, g["tokens"][:, :2, :], cfg.num_heads, cfg.hop_length)
    _gate("prefix-nq2/audio", ref, a)
    # still healthy after the errors
    _gate("after-errors/audio", g["audio"], to_np(codec.decode(tok)))


def test_varlen_lengths_extension():
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    L = tok.shape[2]
    lens = torch.tensor([L, L - 4], dtype=torch.int32)
    a = to_np(codec.decode(tok, lengths=lens))
    spt = cfg.samples_per_token
    ref0 = O.decode(sd, g["tokens"][0:1], cfg.num_heads, cfg.hop_length)
    ref1 = O.decode(sd, g["tokens"][1:2, :, :L - 4], cfg.num_heads, cfg.hop_length)
    _gate("varlen/item0", ref0, a[0:1])
    _gate("varlen/item1", ref1, a[1:2, :(L - 4) * spt])
    assert np.all(a[1, (L - 4) * spt:] == 0)


def test_varlen_lengths_edge_values():
    """Ragged batches at the edges of the ``lengths`` extension: an item of zero tokens (an empty turn) and a negative
    count decode to all zeros, a count beyond L is clamped to L; the neighbours are bit-identical to their padded-batch
    decode, for the fp32 and the int16 form and for decode + resample; a shortened item matches its standalone decode."""
    from fireredtts2_b200.config import SMALL
    from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
    cfg = SMALL
    codec = build_codec(cfg, synthetic_state_dict(cfg, 9))
    L = 11
    tok = torch.from_numpy(synthetic_tokens(cfg, 5, L, 4)).cuda()
    lens = torch.tensor([L, 0, L + 5, -3, 4], dtype=torch.int32)
    full = codec.decode(tok)
    spt = cfg.samples_per_token
    for pcm16 in (False, True):
        ref = codec.decode(tok, pcm16=True) if pcm16 else full
        a = codec.decode(tok, lengths=lens, pcm16=pcm16)
        assert torch.equal(a[0], ref[0]) and torch.equal(a[2], ref[2])
        assert not bool(a[1].any()) and not bool(a[3].any())
        assert not bool(a[4, 4 * spt:].any()) and bool(a[4, :4 * spt].any())
    # against the standalone decode of the same 4 tokens: another batch shape, so other kernels may serve it (32 rows here)
    # — equal within the parity tolerance, not necessarily bit for bit
    short = codec.decode(tok[4:5, :, :4])
    _, snr = report("ragged item vs standalone decode", to_np(short[0]), to_np(codec.decode(tok, lengths=lens)[4, :4 * spt]))
    assert snr >= 40.0
    a24, a16 = codec.decode_resampled(tok, 16000, lengths=lens)
    assert torch.equal(a24, codec.decode(tok, lengths=lens))
    assert not bool(a16[1].any()) and not bool(a16[3].any()) and bool(a16[0].any())
    # every item empty
    z = codec.decode(tok, lengths=torch.zeros(5, dtype=torch.int32))
    assert z.shape == full.shape and not bool(z.any())


def test_c0_multi_wave_batch_items_are_independent():
    """A batch large enough that every GEMM of the step runs several waves of tiles over the 148 SMs (12 x 159 tokens
    at C0 = 15 264 frames): identical items must decode to bit-identical waveforms, equal to the standalone decode of
    one item — with and without the folded LayerNorm.  (Regression: the folded-LayerNorm producer once wrote its fp16
    copy into the buffer later tiles of the same launch were still reading as A.)"""
    from fireredtts2_b200.config import C0
    from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
    import fireredtts2_b200._native as N
    cfg = C0
    codec = build_codec(cfg, synthetic_state_dict(cfg, 0), check_indices=False)
    one = torch.from_numpy(synthetic_tokens(cfg, 1, 159, 77)).cuda()
    tok = one.expand(12, -1, -1).contiguous()
    outs = {}
    for name, flags in (("fold", 0), ("nofold", N.DBG_NO_LNFOLD)):
        codec.set_debug(flags)
        single = codec.decode(one)
        batch = codec.decode(tok)
        diffs = [float((batch[k] - single[0]).abs().max()) for k in range(12)]
        print(f"[parity] c0 12x159 {name}: max |item_k - standalone| = {max(diffs):.3e}")
        assert max(diffs) == 0.0, (name, diffs)
        outs[name] = to_np(single)
    codec.set_debug(0)
    _, snr = report("c0 12x159 fold vs separate LayerNorm kernels", outs["nofold"], outs["fold"])
    assert snr > 50.0


def test_scatter_decode_equals_padded_decode_bitwise():
    """frt2_decode_scatter: each item's samples land at its own offset of one flat buffer (the concatenated dialogue,
    reference fireredtts2.py:399-401), bit-identical to the padded-batch decode, and nothing else is touched."""
    from fireredtts2_b200.sharding import PeerBuffer, unit_offsets
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    B, _, L = tok.shape
    spt = cfg.samples_per_token
    lens = [L, L - 4][:B] + [L] * max(0, B - 2)
    lens_t = torch.tensor(lens, dtype=torch.int32)
    a = codec.decode(tok, lengths=lens_t)
    order = list(reversed(range(B)))                       # units are laid out in reverse item order, with a gap
    offs = unit_offsets([lens[i] for i in order], spt)
    gap = 7
    item_off = {i: offs[k] + gap * (k + 1) for k, i in enumerate(order)}
    total = offs[-1] + gap * (B + 1)
    for pcm16 in (False, True):
        ref = codec.decode(tok, lengths=lens_t, pcm16=pcm16)
        sentinel = 12345 if pcm16 else 777.0
        buf = PeerBuffer(total, torch.int16 if pcm16 else torch.float32, torch.device("cuda", 0))
        try:
            flat = buf.tensor()
            flat.fill_(sentinel)
            codec.decode_into(tok, buf.ptr, torch.tensor([item_off[i] for i in range(B)]), lens_t, pcm16=pcm16)
            torch.cuda.synchronize()
            mask = torch.ones(total, dtype=torch.bool, device="cuda")
            for i in range(B):
                n = spt * lens[i]
                assert torch.equal(flat[item_off[i]:item_off[i] + n], ref[i, :n]), (i, pcm16)
                mask[item_off[i]:item_off[i] + n] = False
            assert bool((flat[mask] == sentinel).all())    # gaps and the padding of short items stay untouched
        finally:
            buf.close()
    assert torch.equal(a, codec.decode(tok, lengths=lens_t))
    with pytest.raises(ValueError):
        codec.decode_into(tok, 0, torch.zeros(B, dtype=torch.int64), lens_t)


@pytest.mark.parametrize("mode", ["tc_gemm+warp_attn", "product", "no_graph", "no_skinny"])
@pytest.mark.parametrize("case", cases("stream"), ids=lambda c: c["name"])
def test_streaming_vs_reference_golden(case, mode):
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd, stream_max_tokens=max(32, case["L"] + 1))
    codec.set_debug(MODES[mode])
    tok = torch.from_numpy(g["tokens"]).cuda()
    chunks = list(g["chunks"])
    cache, pos, outs = {}, 0, []
    for i, lc in enumerate(chunks):
        a, cache = codec.decode_one_token(tok[:, :, pos:pos + lc], cache, i == len(chunks) - 1)
        ref = g[f"audio_{i}"]
        assert tuple(a.shape) == ref.shape
        outs.append(to_np(a))
        pos += lc
    cat = np.concatenate(outs, axis=1)
    refcat = np.concatenate([g[f"audio_{i}"] for i in range(len(chunks))], axis=1)
    _gate(f"{case['name']}/{mode}/stream-audio", refcat, cat)
    exported = codec.export_cache(cache)
    for k, v in exported.items():
        if "cache_" + k not in g.files:      # the K/V cache of the C0-width fixtures is not stored (12.6 MB per 16 tokens)
            assert k == "bb_kv_cache"
            continue
        ref = g["cache_" + k]
        assert tuple(v.shape) == ref.shape, k
        _, snr = report(f"{case['name']}/{mode}/cache/{k}", ref, to_np(v))
        assert snr > 45.0, k
    with pytest.raises(ValueError):
        codec.decode_one_token(tok[:, :, :1], cache, False)   # stream already finished


def test_stream_state_roundtrip_and_overflow():
    case = [c for c in cases("stream") if c["name"] == "tiny_stream_1"][0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd, stream_max_tokens=4)
    tok = torch.from_numpy(g["tokens"]).cuda()
    cache = {}
    for i in range(3):
        a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, False)
    exported = codec.export_cache(cache)
    # hand-off: continue from the reference-layout tensors in a fresh stream
    a4, c2 = codec.decode_one_token(tok[:, :, 3:4], dict(exported), False)
    a4_direct, cache = codec.decode_one_token(tok[:, :, 3:4], cache, False)
    report("state-roundtrip", to_np(a4_direct), to_np(a4))
    assert np.abs(to_np(a4) - to_np(a4_direct)).max() < 1e-3
    with pytest.raises(OverflowError):
        codec.decode_one_token(tok[:, :, 4:5], cache, False)


def test_graph_replay_equals_kernel_by_kernel():
    """The captured per-token CUDA graph (position / flags read from HBM) must reproduce the eager launch sequence
    bit for bit, across a reset and for many steps."""
    case = [c for c in cases("stream") if c["name"] == "tiny_stream_1"][0]
    cfg, sd, g = load_case(case)
    rng = np.random.default_rng(0)
    tok = torch.from_numpy(rng.integers(0, cfg.codebook_size, size=(2, cfg.num_quantizers, 24))).cuda()
    outs = {}
    for mode in ("product", "no_graph"):
        codec = build_codec(cfg, sd, stream_max_tokens=32)
        codec.set_debug(MODES[mode])
        state = codec.new_stream(2)
        for rep in range(2):          # second pass re-uses the same state (and the same graph) after a reset
            codec.reset_stream(state)
            cache, chunks = state, []
            for i in range(24):
                a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, i == 23)
                chunks.append(to_np(a))
            outs[(mode, rep)] = np.concatenate(chunks, axis=1)
    assert np.array_equal(outs[("product", 0)], outs[("no_graph", 0)])
    assert np.array_equal(outs[("product", 1)], outs[("product", 0)])
    ref = O.decode(sd, tok.cpu().numpy(), cfg.num_heads, cfg.hop_length)
    _gate("graph-stream-24-vs-offline", ref, outs[("product", 0)])


def test_streaming_long_context_splits_attention_over_ctas():
    """Past 512 frames of K/V state the step's attention runs on several CTAs per (item, head) whose partial softmax
    states are merged by the last CTA to finish: 150 tokens (1200 frames -> 3 working CTAs) streamed one at a time must
    still match the offline decode of the same tokens (the reference: streaming == offline, model.py:326-376), and
    the graph replay must equal the eager launch sequence bit for bit (deterministic merge order)."""
    case = [c for c in cases("stream") if c["name"] == "tiny_stream_1"][0]
    cfg, sd, g = load_case(case)
    assert cfg.head_dim == 64
    n = 150
    tok = torch.from_numpy(np.random.default_rng(3).integers(0, cfg.codebook_size, size=(2, cfg.num_quantizers, n))).cuda()
    outs = {}
    for mode in ("product", "no_graph"):
        codec = build_codec(cfg, sd, stream_max_tokens=n + 2)
        codec.set_debug(MODES[mode])
        cache, chunks = {}, []
        for i in range(n):
            a, cache = codec.decode_one_token(tok[:, :, i:i + 1], cache, i == n - 1)
            chunks.append(to_np(a))
        outs[mode] = np.concatenate(chunks, axis=1)
    assert np.array_equal(outs["product"], outs["no_graph"])
    codec = build_codec(cfg, sd)
    offline = to_np(codec.decode(tok))
    _, snr = report("stream-150-tokens vs offline decode", offline, outs["product"])
    assert snr > 50.0
    ref = O.decode(sd, tok.cpu().numpy(), cfg.num_heads, cfg.hop_length)
    _gate("stream-150-tokens vs oracle", ref, outs["product"])


def test_pcm16_output_is_the_reference_wire_format():
    """decode(pcm16=True) == (decode() * 32767).astype(int16), the reference's wire conversion
    (enhanced_fireredtts2.py:603,655), bit for bit."""
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    f32 = to_np(codec.decode(tok))
    pcm = codec.decode(tok, pcm16=True).cpu().numpy()
    assert pcm.dtype == np.int16 and pcm.shape == f32.shape
    assert np.array_equal(pcm, (f32 * 32767).astype(np.int16))


def test_long_utterance_against_oracle():
    """A 24 s utterance (T = 2400 frames): many query-tile pairs / key tiles in the block-causal attention and many
    M tiles per item in the batched conv GEMMs; compared with the oracle on the same seeded inputs."""
    from fireredtts2_b200.config import SMALL
    from fireredtts2_b200.weights import synthetic_state_dict, synthetic_tokens
    cfg = SMALL
    sd = synthetic_state_dict(cfg, 5)
    tok = synthetic_tokens(cfg, 2, 300, 9)
    codec = build_codec(cfg, sd)
    a = to_np(codec.decode(torch.from_numpy(tok).cuda()))
    ref = O.decode(sd, tok, cfg.num_heads, cfg.hop_length)
    _gate("small_B2_L300/audio", ref, a)


@pytest.mark.parametrize("new_freq", [16000, 8000, 12000])
def test_decode_resampled_is_bit_identical_to_decode_then_resample(new_freq):
    """frt2_decode_resampled (overlap-add + resampler in one kernel; the context loop's decode -> torchaudio resample,
    fireredtts2.py:386-391): both outputs equal the two separate calls bit for bit, ragged lengths included, and the
    resampled waveform matches the oracle's torchaudio restatement on the oracle's own decode."""
    from fireredtts2_b200.codec import resample
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    codec = build_codec(cfg, sd)
    tok = torch.from_numpy(g["tokens"]).cuda()
    B, _, L = tok.shape
    a24 = codec.decode(tok)
    a_rs = resample(a24, 24000, new_freq)
    f24, f_rs = codec.decode_resampled(tok, new_freq)
    assert torch.equal(f24, a24) and torch.equal(f_rs, a_rs)
    only24, only_rs = codec.decode_resampled(tok, new_freq, return_native=False)
    assert only24 is None and torch.equal(only_rs, a_rs)
    if B >= 2:
        lens = torch.tensor([L] + [max(1, L - 3)] * (B - 1), dtype=torch.int32)
        v24 = codec.decode(tok, lengths=lens)
        v_rs = resample(v24, 24000, new_freq, lengths=(lens * cfg.samples_per_token))
        w24, w_rs = codec.decode_resampled(tok, new_freq, lengths=lens)
        assert torch.equal(w24, v24) and torch.equal(w_rs, v_rs)
    ref = O.resample(O.decode(sd, g["tokens"], cfg.num_heads, cfg.hop_length), 24000, new_freq)
    _, snr = report(f"decode_resampled {new_freq} Hz vs oracle", ref, to_np(f_rs))
    assert f_rs.shape == ref.shape and snr >= 40.0
    with pytest.raises(ValueError):
        codec.decode_resampled(tok, 22050)        # 24000 -> 22050: 147 phases per frame, not served by the fused kernel
