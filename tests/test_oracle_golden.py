"""The oracle (numpy restatement) against golden vectors produced by the REAL reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest

from oracle import codec_oracle as O
from tests.helpers import cases, load_case

TOL_AUDIO = 2e-6      # fp32 vs fp32, different summation orders (observed <= 2.2e-7)
TOL_FEAT = 2e-5       # intermediates have magnitudes of a few units (observed <= 4.3e-6)
# adversarial weights (weights.adversarial_state_dict): activations in the thousands and a residual stream whose rows sit
# at ~100 +- 2 put the fp32 noise floor of ANY summation order at ~1e-4 of a 0.4 peak — the reference's own streaming vs
# offline decode differ by 1.0e-4 on these weights (MANIFEST stream_vs_offline_maxabs)
TOL_AUDIO_ADV = 1e-3
SNR_ADV_DB = 60.0


def _tol(case):
    return TOL_AUDIO_ADV if case.get("weights") == "adversarial" else TOL_AUDIO


@pytest.mark.parametrize("case", cases("offline") + cases("reference_init"), ids=lambda c: c["name"])
def test_offline_decode_matches_reference(case):
    cfg, sd, g = load_case(case)
    taps = {}
    y = O.decode(sd, g["tokens"], cfg.num_heads, cfg.hop_length, taps=taps)
    assert y.shape == g["audio"].shape == (case["B"], case["L"] * cfg.samples_per_token)
    assert np.abs(y - g["audio"]).max() < _tol(case)
    assert O.snr_db(g["audio"], y) > (SNR_ADV_DB if case.get("weights") == "adversarial" else 100.0)
    for k in ("z", "x50", "prior", "layer0", "final"):
        if k in g.files:
            assert np.abs(taps[k] - g[k]).max() < TOL_FEAT, k
    if "up_full" in g.files:  # reference keeps 3 extra frames before the trim (decoder.py:615)
        up = taps["up"]
        assert np.abs(up - g["up_full"][:, :up.shape[1]]).max() < TOL_FEAT


@pytest.mark.parametrize("case", cases("stream"), ids=lambda c: c["name"])
def test_streaming_decode_matches_reference(case):
    cfg, sd, g = load_case(case)
    tok, chunks = g["tokens"], list(g["chunks"])
    st, pos, outs = None, 0, []
    for i, lc in enumerate(chunks):
        y, st = O.decode_chunk(sd, tok[:, :, pos:pos + lc], st, i == len(chunks) - 1, cfg.num_heads, cfg.hop_length)
        ref = g[f"audio_{i}"]
        first, last = i == 0, i == len(chunks) - 1
        n = cfg.samples_per_token * lc - cfg.istft_pad * first + cfg.istft_pad * last
        assert y.shape == ref.shape == (case["B"], n)
        assert np.abs(y - ref).max() < _tol(case)
        outs.append(y)
        pos += lc
    adv = case.get("weights") == "adversarial"
    for k, v in st.to_reference_layout(cfg.num_heads).items():
        if "cache_" + k not in g.files:     # the large K/V cache is not stored for the C0-width fixtures
            assert k == "bb_kv_cache"
            continue
        ref = g["cache_" + k]
        assert v.shape == ref.shape, k
        assert np.abs(v - ref).max() < (TOL_FEAT if not adv else 1e-3 * max(1.0, float(np.abs(ref).max()))), k
    # one token per call == offline (block of 8 frames == one token); multi-token chunks are unmasked
    # inside the chunk (whisper.py:107-113) and legitimately differ from offline.
    cat = np.concatenate(outs, axis=1)
    assert cat.shape == (case["B"], case["L"] * cfg.samples_per_token)
    if "offline" in g.files and all(c == 1 for c in chunks):
        assert np.abs(cat - g["offline"]).max() < TOL_AUDIO


@pytest.mark.parametrize("case", cases("rvq_emb"), ids=lambda c: c["name"])
def test_rvq_sum_matches_reference_bit_for_bit(case):
    """C1 (C0 with Identity out_project): the index-ordered fp32 sum that reaches rvq.output_proj in the real
    reference (captured with a forward pre-hook by oracle/make_golden.py) equals the oracle's, bit for bit."""
    cfg, sd, g = load_case(case)
    emb, z = O.rvq_decode_codes(sd, g["tokens"])
    assert emb.shape == g["emb"].shape and np.array_equal(emb, g["emb"])
    assert np.abs(z - g["z"]).max() < TOL_FEAT


def test_identity_projection_sum_is_index_ordered():
    case = [c for c in cases("offline") if c["name"] == "tiny_ident_offline"][0]
    cfg, sd, g = load_case(case)
    rows = O.rvq_gather(sd, g["tokens"])
    emb, _ = O.rvq_decode_codes(sd, g["tokens"])
    acc = np.zeros_like(rows[:, :, 0, :])
    for i in range(cfg.num_quantizers):
        acc = acc + rows[:, :, i, :]
    assert np.array_equal(acc, emb)
    for i in range(cfg.num_quantizers):
        cb = sd[f"rvq.quantizers.{i}.codebook"]
        assert np.array_equal(rows[:, :, i, :], cb[g["tokens"][:, i, :]])


def test_out_of_range_index_raises():
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    tok = g["tokens"].copy()
    tok[0, 1, 2] = cfg.codebook_size
    with pytest.raises(IndexError):
        O.decode(sd, tok, cfg.num_heads)
    tok[0, 1, 2] = -1
    with pytest.raises(IndexError):
        O.decode(sd, tok, cfg.num_heads)


def test_prefix_codebooks_allowed():
    # quantizers[:nq] — fewer codebooks than configured are accepted (rvq.py:160)
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    y = O.decode(sd, g["tokens"][:, :2, :], cfg.num_heads)
    assert y.shape == g["audio"].shape and np.isfinite(y).all()


def test_batch_row_equals_single():
    case = cases("offline")[0]
    cfg, sd, g = load_case(case)
    y = O.decode(sd, g["tokens"], cfg.num_heads)
    y0 = O.decode(sd, g["tokens"][1:2], cfg.num_heads)
    assert np.abs(y[1:2] - y0).max() < TOL_AUDIO


@pytest.mark.parametrize("case", cases("offline") + cases("reference_init"), ids=lambda c: c["name"])
def test_torch_cpu_port_matches_reference(case):
    """The torch-CPU port used as the timed CPU baseline (bench.py) against the same golden vectors."""
    from oracle import codec_oracle_torch as OT
    cfg, sd, g = load_case(case)
    y = OT.decode(OT.to_torch(sd), g["tokens"], cfg.num_heads, cfg.hop_length).numpy()
    assert y.shape == g["audio"].shape
    assert np.abs(y - g["audio"]).max() < _tol(case)


def _resample_cases():
    import os
    from .helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "resample.npz"))
    names = sorted({k.split("::")[0] for k in g.files if "::" in k})
    return [(n, g[n + "::x"], g[n + "::y"], int(g[n + "::rates"][0]), int(g[n + "::rates"][1])) for n in names]


@pytest.mark.parametrize("case", _resample_cases(), ids=lambda c: c[0])
def test_resample_oracle_matches_torchaudio_golden(case):
    """oracle.resample restates torchaudio.functional.resample (third-party, torchaudio 2.11.0, not vendored in the
    reference: fireredtts2.py:65,389-391); pinned to vectors produced by torchaudio (oracle/make_golden_resample.py)."""
    name, x, y, orig, new = case
    out = O.resample(x, orig, new)
    assert out.shape == y.shape and out.dtype == np.float32
    assert np.abs(out - y).max() <= 2e-7 * max(1.0, float(np.abs(y).max())) + 2e-7


# ---------------------------------------------------------------- RVQ encode side (first stage)
RVQ_ENC_CASES = [("rvq_encode_tiny", "TINY"), ("rvq_encode_tiny_ident", "TINY_IDENT"),
                 ("rvq_encode_tiny_noinput", "TINY"), ("rvq_encode_small", "SMALL"), ("rvq_encode_c0", "C0")]


def load_rvq_encode_case(name, preset):
    import os
    from fireredtts2_b200.config import PRESETS
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    wseed, _, input_dim = (int(x) for x in g["meta"])
    cfg = PRESETS[preset]
    sd = dict(synthetic_state_dict(cfg, wseed))
    sd.update(synthetic_encode_tensors(cfg, wseed, input_dim))
    return cfg, sd, g


@pytest.mark.parametrize("name,preset", RVQ_ENC_CASES)
def test_oracle_rvq_encode_matches_reference_indices(name, preset):
    """oracle.rvq_encode_codes against ResidualVQ.encode_codes of the real reference: every index identical (the
    fixtures' smallest top-2 margin is 1e-3, far above fp32 rounding of the distances), margins agree."""
    cfg, sd, g = load_rvq_encode_case(name, preset)
    codes, margin = O.rvq_encode_codes(sd, g["z"])
    assert codes.shape == g["codes"].shape and codes.dtype == np.int64
    assert np.array_equal(codes, g["codes"])
    assert np.allclose(margin, g["margin"], rtol=0, atol=2e-3 * max(1.0, float(np.abs(g["margin"]).max())))
    # encode -> decode_codes round trip is a contraction: the residual shrinks with every quantizer (property)
    # and prefix nq reproduces the first rows
    c2, _ = O.rvq_encode_codes(sd, g["z"], nq=2)
    assert np.array_equal(c2, g["codes"][:2])


def test_oracle_rvq_encode_properties():
    """Size-independent properties of ResidualVQ.encode_codes on the oracle: (1) with Identity projections a codebook row is
    its own nearest neighbour (distance 0), (2) every stage picks the codebook row closest to the current residual, (3) tokens are independent (any slice encodes the same)."""
    from fireredtts2_b200.config import TINY_IDENT
    from fireredtts2_b200.weights import synthetic_state_dict
    cfg = TINY_IDENT
    sd = synthetic_state_dict(cfg, 3)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, cfg.codebook_size, size=(2, 40))
    z = np.asarray(sd["rvq.quantizers.0.codebook"])[idx].transpose(0, 2, 1).copy()      # (B, cd, T), input_dim == rvq_dim
    codes, _ = O.rvq_encode_codes(sd, z, nq=1)
    assert np.array_equal(codes[0], idx)
    z = rng.standard_normal((2, cfg.rvq_dim, 30)).astype(np.float32)
    codes, _ = O.rvq_encode_codes(sd, z)
    # every stage picks the row closest to the CURRENT residual (brute force in float64; Identity projections)
    resid = z.transpose(0, 2, 1).astype(np.float64).reshape(-1, cfg.rvq_dim)
    for i in range(cfg.num_quantizers):
        C = np.asarray(sd[f"rvq.quantizers.{i}.codebook"], dtype=np.float64)
        d = ((resid[:, None, :] - C[None, :, :]) ** 2).sum(-1)
        chosen = codes[i].reshape(-1)
        assert np.all(d[np.arange(len(chosen)), chosen] <= d.min(axis=1) + 1e-4)
        resid = resid - C[chosen]
    part, _ = O.rvq_encode_codes(sd, z[1:, :, 5:17])
    assert np.array_equal(part, codes[:, 1:, 5:17])


# ---- codec encode side behind the feature encoders (SslAdaptor + cat + ResidualDownConv, model.py:19-121,225-232) ----
ENC_CASES = [("enc_etiny", "ETINY", "TINY"), ("enc_esmall", "ESMALL", "SMALL"), ("enc_ec0", "EC0", "C0")]


def load_encoder_case(name, enc_preset, preset):
    """-> (encoder cfg, codec cfg, encoder state_dict, codec state_dict incl. RVQ encode tensors, ssl, aco, golden)"""
    import os
    from tests.helpers import GOLDEN
    from fireredtts2_b200.config import PRESETS
    from fireredtts2_b200.encoder import ENC_PRESETS, synthetic_encoder_state_dict, synthetic_features
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, T, wseed, dseed = [int(v) for v in g["meta"]]
    ecfg, cfg = ENC_PRESETS[enc_preset], PRESETS[preset]
    esd = synthetic_encoder_state_dict(ecfg, wseed)
    sd = dict(synthetic_state_dict(cfg, wseed))
    sd.update(synthetic_encode_tensors(cfg, wseed, ecfg.down_dim))
    ssl, aco = synthetic_features(ecfg, B, T, dseed)
    return ecfg, cfg, esd, sd, ssl, aco, g


@pytest.mark.parametrize("name,enc_preset,preset", ENC_CASES)
def test_encoder_oracle_matches_reference_golden(name, enc_preset, preset):
    """oracle.encoder_oracle against SslAdaptor / ResidualDownConv / ResidualVQ.encode_codes of the real reference."""
    from oracle import encoder_oracle as EO
    ecfg, cfg, esd, sd, ssl, aco, g = load_encoder_case(name, enc_preset, preset)
    sem = EO.ssl_adaptor(esd, ssl, ecfg.ssl_num_heads)
    assert sem.shape == g["sem"].shape
    assert np.abs(sem - g["sem"]).max() < 2e-5 and O.snr_db(g["sem"], sem) > 100.0
    vq = EO.encode_features(esd, ssl, aco, ecfg.ssl_num_heads, ecfg.avg_pooler)
    assert vq.shape == g["vq_in"].shape
    assert np.abs(vq - g["vq_in"]).max() < 2e-5 and O.snr_db(g["vq_in"], vq) > 100.0
    # the indices the reference derives from ITS features, reproduced by the oracle from the same features
    codes, _ = O.rvq_encode_codes(sd, np.ascontiguousarray(g["vq_in"].transpose(0, 2, 1)))
    assert np.array_equal(codes.transpose(1, 0, 2), g["codes"])


def test_encoder_config_roundtrip():
    from fireredtts2_b200.encoder import EC0, EncoderConfig, encoder_keys, synthetic_encoder_state_dict
    d = EC0.to_reference_dict()
    d["acoustic_encoder"] = {"embed_dim": EC0.aco_dim}
    assert EncoderConfig.from_reference_dict({"codec": d}) == EC0
    with pytest.raises(ValueError):
        bad = dict(d, downsample=dict(d["downsample"], embed_dim=EC0.down_dim + 8))
        EncoderConfig.from_reference_dict(bad)
    # with the feature encoders: the SSL encoder's shape is fixed by PretrainedWhisperEncoder.from_pretrained (whisper.py:359-369)
    from fireredtts2_b200.encoder import EC0F, front_keys, synthetic_front_state_dict, ETINYF
    df = EC0F.to_reference_dict()
    df["with_feature_encoders"] = True
    assert EncoderConfig.from_reference_dict({"codec": df}) == EC0F
    assert sorted(synthetic_front_state_dict(ETINYF, 0)) == sorted(front_keys(ETINYF))
    sd = synthetic_encoder_state_dict(EC0, 0)
    assert sorted(sd) == sorted(encoder_keys(EC0))
    assert sd["downsample.gate_proj.weight"].shape == (4 * EC0.down_dim, EC0.down_dim, 4)


# ---- the whole encode path from the waveform (log-mel, both Whisper encoders, ssl_adaptor, downsample; model.py:218-236) ----
ENC_AUDIO_CASES = [("encaudio_etinyf", "ETINYF", "TINY"), ("encaudio_epadf", "EPADF", "SMALL")]


def load_encoder_audio_case(name, enc_preset, preset):
    """-> (encoder cfg, codec cfg, encoder state_dict incl. the feature encoders, codec state_dict, audio, golden)"""
    import os
    from tests.helpers import GOLDEN
    from fireredtts2_b200.config import PRESETS
    from fireredtts2_b200.encoder import (ENC_PRESETS, synthetic_audio, synthetic_encoder_state_dict,
                                          synthetic_front_state_dict)
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, n, wseed, dseed = [int(v) for v in g["meta"]]
    ecfg, cfg = ENC_PRESETS[enc_preset], PRESETS[preset]
    esd = dict(synthetic_encoder_state_dict(ecfg, wseed))
    esd.update(synthetic_front_state_dict(ecfg, wseed))
    sd = dict(synthetic_state_dict(cfg, wseed))
    sd.update(synthetic_encode_tensors(cfg, wseed, ecfg.down_dim))
    return ecfg, cfg, esd, sd, synthetic_audio(B, n, dseed), g


@pytest.mark.parametrize("name,enc_preset,preset", ENC_AUDIO_CASES)
def test_encoder_audio_oracle_matches_reference_golden(name, enc_preset, preset):
    """oracle.encoder_oracle (log_mel, whisper_encoder, ...) against RedCodecInfer._encode_one_batch of the real reference."""
    from oracle import encoder_oracle as EO
    ecfg, cfg, esd, sd, audio, g = load_encoder_audio_case(name, enc_preset, preset)
    taps = {}
    vq = EO.encode_audio_features(esd, audio, ecfg, taps=taps)
    assert np.abs(taps["mel"] - g["mel"]).max() < 2e-4 and O.snr_db(g["mel"], taps["mel"]) > 100.0
    for k in ("ssl", "aco"):
        assert taps[k].shape == g[k].shape
        assert np.abs(taps[k] - g[k]).max() < 2e-5 and O.snr_db(g[k], taps[k]) > 100.0
    assert np.abs(vq - g["vq_in"]).max() < 2e-5 and O.snr_db(g["vq_in"], vq) > 100.0


def test_mel_filter_bank_and_positions_match_reference_formulas():
    """The slaney bank and the sinusoidal table restated in oracle/ and fireredtts2_b200/: basic invariants (the golden
    log-mel above pins the values)."""
    from oracle import encoder_oracle as EO
    from fireredtts2_b200.encoder import sinusoids
    bank = EO.mel_filter_bank(201, 128)
    assert bank.shape == (201, 128) and (bank >= 0).all() and (bank.max(axis=0) > 0).all()
    assert bank[0].sum() == 0.0                       # DC belongs to no filter's interior
    peaks = bank.argmax(axis=0)
    assert (np.diff(peaks) >= 0).all()                # centres ascend
    pos = sinusoids(1500, 1280)
    assert pos.shape == (1500, 1280) and np.allclose(pos[0, :640], 0.0) and np.allclose(pos[0, 640:], 1.0)


# ---- frame tail of the speech LM (oracle/frame_decoder_oracle.py; goldens from oracle/make_golden_frame_decoder.py) ----
FD_CASES = [("fd_tiny", "FD_TINY"), ("fd_small", "FD_SMALL"), ("fd_small_b1", "FD_SMALL"), ("fd_200m", "FD_200M")]


def load_fd_case(name, preset):
    import os
    from fireredtts2_b200.frame_decoder import FD_PRESETS, synthetic_frame_decoder_state_dict
    from tests.helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = FD_PRESETS[preset]
    B, wseed, dseed, topk = (int(v) for v in g["meta"])
    return cfg, synthetic_frame_decoder_state_dict(cfg, wseed), g, topk, float(g["temperature"])


@pytest.mark.parametrize("name,preset", FD_CASES)
def test_frame_decoder_oracle_matches_reference_generate_frame(name, preset):
    """oracle.frame_decoder_oracle.generate_codes against the reference's own Model.generate_frame (llm.py:274-330; Qwen2
    blocks from Hugging Face transformers behind a torchtune shim): same codes, logits to fp32 summation order."""
    from oracle import frame_decoder_oracle as FO
    cfg, sd, g, topk, temperature = load_fd_case(name, preset)
    codes, logits = FO.generate_codes(sd, cfg, g["last_h"], topk, temperature, g["noise"])
    assert codes.shape == g["codes"].shape == (g["last_h"].shape[0], cfg.audio_num_codebooks)
    assert np.array_equal(codes, g["codes"])
    assert np.abs(logits - g["logits"]).max() < 1e-4 and O.snr_db(g["logits"], logits) > 100.0
    # teacher forcing reproduces the same logits, and a given c0 skips the codebook-0 head
    codes_f, logits_f = FO.generate_codes(sd, cfg, g["last_h"], topk, temperature, g["noise"], forced=g["codes"])
    assert np.array_equal(codes_f, g["codes"]) and np.array_equal(logits_f, logits)
    codes_c, logits_c = FO.generate_codes(sd, cfg, g["last_h"], topk, temperature, g["noise"], c0=g["codes"][:, 0])
    assert np.array_equal(codes_c, g["codes"]) and not logits_c[:, 0].any()


def test_sample_topk_oracle_edge_cases():
    """ties at the k-th value stay in the candidate set (llm.py:43 uses `<`), topk = 1 is arg-max, topk = V keeps all."""
    from oracle import frame_decoder_oracle as FO
    logits = np.array([[0.0, 3.0, 3.0, 1.0, 3.0, -2.0]], np.float32)
    q = np.array([[1.0, 9.0, 1.0, 1.0, 0.5, 1e-6]], np.float32)
    assert FO.sample_topk(logits, 2, 1.0, q)[0] == 4            # three tied maxima survive topk = 2; the smallest q wins
    assert FO.sample_topk(logits, 1, 0.5, np.ones_like(q))[0] == 1
    assert FO.sample_topk(logits, 6, 1.0, q)[0] == 5            # nothing filtered: the tiny q of the last entry wins


def test_sample_topk_oracle_matches_the_reference_sampler():
    """oracle.sample_topk against the reference's own sample_topk / _multinomial_sample_one_no_sync (llm.py:34-49) on seeded
    logits with exact ties at the k-th value; the Exp(1) draws are the ones the reference made (tests/golden/fd_sampler.npz)."""
    import os
    from oracle import frame_decoder_oracle as FO
    from tests.helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "fd_sampler.npz"))
    for i in range(int(g["n_cases"])):
        V, topk, rows = (int(v) for v in g[f"c{i}_meta"])
        got = FO.sample_topk(g[f"c{i}_logits"], topk, float(g[f"c{i}_temperature"]), g[f"c{i}_q"])
        assert got.shape == (rows,) and np.array_equal(got, g[f"c{i}_codes"]), (i, V, topk)
