"""GPU parity of the codec ENCODE side behind the feature encoders (frt2_enc_features -> frt2_rvq_encode; reference
codec/model.py:19-121,225-236) through the C ABI, against golden vectors of the REAL reference."""
import numpy as np
import pytest
import torch

from oracle import codec_oracle as O
from oracle import encoder_oracle as EO
from tests.gpu_common import build_codec, report
from tests.test_oracle_golden import ENC_CASES, load_encoder_case

pytestmark = pytest.mark.gpu
SNR_GATE_DB = 40.0


def build_encoder(ecfg, esd):
    from fireredtts2_b200.encoder import CodecEncoderB200
    return CodecEncoderB200(ecfg, esd, device="cuda:0")


@pytest.mark.parametrize("name,enc_preset,preset", ENC_CASES)
def test_encoder_features_vs_reference_golden(name, enc_preset, preset):
    ecfg, cfg, esd, sd, ssl, aco, g = load_encoder_case(name, enc_preset, preset)
    enc = build_encoder(ecfg, esd)
    vq = enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
    assert tuple(vq.shape) == g["vq_in"].shape and vq.dtype == torch.float32
    assert enc.last_launches == 10 + 7 * ecfg.ssl_num_layers
    maxabs, snr = report(name + "/vq_in_feats", g["vq_in"], vq.cpu().numpy())
    assert snr >= SNR_GATE_DB and maxabs <= 0.05 * np.abs(g["vq_in"]).max()
    # deterministic, and every item independent of its batch neighbours
    vq2 = enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
    assert torch.equal(vq, vq2)
    one = enc.features(torch.from_numpy(ssl[-1:]).cuda(), torch.from_numpy(aco[-1:]).cuda())
    _, snr1 = report(name + "/last item alone vs in batch", vq[-1:].cpu().numpy(), one.cpu().numpy())
    assert snr1 >= 80.0     # tile shapes differ with the row count: fp32 summation order only


@pytest.mark.parametrize("name,enc_preset,preset", ENC_CASES)
def test_encoder_codes_vs_reference_golden(name, enc_preset, preset):
    """Indices from OUR features: identical to the reference's wherever its top-2 margin exceeds what the feature error
    can move a distance by; a token that diverged at a near-tie sees another residual afterwards, so only the FIRST
    divergence of a token is judged."""
    ecfg, cfg, esd, sd, ssl, aco, g = load_encoder_case(name, enc_preset, preset)
    enc = build_encoder(ecfg, esd)
    codec = build_codec(cfg, sd)
    codes = enc.encode_features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda(), codec)
    assert tuple(codes.shape) == g["codes"].shape and codes.dtype == torch.int64
    codes = codes.cpu().numpy()
    ref, margin = g["codes"], g["margins"]                       # (B, nq, L), (nq, B, L)
    nq = ref.shape[1]
    diff = codes != ref
    first = np.where(diff.any(axis=1), diff.argmax(axis=1), nq)  # (B, L) first diverging quantizer
    # distance error bound: d = |z|^2 - 2 z.c + |c|^2; a feature error e moves the gap between two codes by
    # 2 e.(c1 - c2) <= 2 |e| |c1 - c2|
    vq = enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda()).cpu().numpy()
    err = np.sqrt(((vq - g["vq_in"]) ** 2).sum(axis=2))          # (B, L) |e| in the rvq input space
    bad = 0
    for b, t in zip(*np.nonzero(first < nq)):
        tol = 40.0 * err[b, t] + 1e-4       # the in/out projections and |c1 - c2| scale the bound; 40x is generous
        if margin[first[b, t], b, t] > tol:
            bad += 1
    agree0 = float((codes[:, 0] == ref[:, 0]).mean())
    print(f"[parity] {name}: first codebook agrees on {agree0 * 100:.1f} % of the tokens, "
          f"{int((first == nq).sum())}/{first.size} tokens identical over all {nq} codebooks, {bad} diverge off a near-tie")
    assert bad == 0 and agree0 >= 0.9


def test_encoder_rejects_ragged_frames_and_bad_widths():
    from fireredtts2_b200.encoder import ETINY, synthetic_encoder_state_dict, synthetic_features
    enc = build_encoder(ETINY, synthetic_encoder_state_dict(ETINY, 1))
    ssl, aco = synthetic_features(ETINY, 1, 10, 0)               # 10 % 4 != 0: the reference's reshape fails (model.py:113)
    with pytest.raises(RuntimeError):
        enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
    ssl, aco = synthetic_features(ETINY, 1, 8, 0)
    with pytest.raises(ValueError):
        enc.features(torch.from_numpy(ssl[:, :, :64]).cuda(), torch.from_numpy(aco).cuda())
    sd = synthetic_encoder_state_dict(ETINY, 1)
    del sd["downsample.up_proj.weight"]
    with pytest.raises(KeyError):
        build_encoder(ETINY, sd)


def test_encoder_full_size_chunk_batch_against_oracle():
    """EC0 at the reference's batch shape: 6 s chunks (T = 300 frames at 50 Hz, model.py:262), 8 of them; one item
    against the numpy oracle, the others through the batch-independence property."""
    from fireredtts2_b200.encoder import EC0, synthetic_encoder_state_dict, synthetic_features
    esd = synthetic_encoder_state_dict(EC0, 0)
    enc = build_encoder(EC0, esd)
    ssl, aco = synthetic_features(EC0, 8, 300, 99)
    vq = enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
    assert tuple(vq.shape) == (8, 75, 1024)
    ref = EO.encode_features(esd, ssl[5:6], aco[5:6], EC0.ssl_num_heads, EC0.avg_pooler)
    maxabs, snr = report("EC0 8x300 item 5 vs oracle", ref, vq[5:6].cpu().numpy())
    assert snr >= SNR_GATE_DB
    for b in (0, 7):
        one = enc.features(torch.from_numpy(ssl[b:b + 1]).cuda(), torch.from_numpy(aco[b:b + 1]).cuda())
        assert O.snr_db(one.cpu().numpy(), vq[b:b + 1].cpu().numpy()) >= 80.0


# ---- the whole path from the waveform: log-mel, both Whisper encoders, ssl_adaptor, downsample, RVQ ----
from tests.test_oracle_golden import ENC_AUDIO_CASES, load_encoder_audio_case  # noqa: E402


@pytest.mark.parametrize("name,enc_preset,preset", ENC_AUDIO_CASES)
def test_encoder_audio_path_vs_reference_golden(name, enc_preset, preset):
    """frt2_enc_audio_features against RedCodecInfer._encode_one_batch of the real reference (model.py:218-236): the
    log-mel features (fp32 kernel: tight), both encoder outputs and the RVQ input (fp16-operand GEMMs: SNR gate)."""
    ecfg, cfg, esd, sd, audio, g = load_encoder_audio_case(name, enc_preset, preset)
    enc = build_encoder(ecfg, esd)
    vq, taps = enc.audio_features(torch.from_numpy(audio).cuda(), taps=True)
    mel = taps["mel"].cpu().numpy()
    maxabs, snr = report(name + "/log-mel", g["mel"], mel)
    assert maxabs < 2e-3 and snr >= 70.0
    for k in ("ssl", "aco"):
        _, snr = report(name + "/" + k + " encoder output", g[k], taps[k].cpu().numpy())
        assert snr >= SNR_GATE_DB
    maxabs, snr = report(name + "/vq_in_feats", g["vq_in"], vq.cpu().numpy())
    assert snr >= SNR_GATE_DB and maxabs <= 0.05 * np.abs(g["vq_in"]).max()
    # without the taps: same result; second call deterministic
    vq2 = enc.audio_features(torch.from_numpy(audio).cuda())
    assert torch.equal(vq, vq2)
    # indices: first codebook mostly identical (later ones see the fp16-operand feature error through the residual chain)
    codec = build_codec(cfg, sd)
    codes = codec.rvq_encode_codes(vq.transpose(1, 2)).permute(1, 0, 2).cpu().numpy()
    agree0 = float((codes[:, 0] == g["codes"][:, 0]).mean())
    print(f"[parity] {name}: first codebook agrees with the reference on {agree0 * 100:.1f} % of the tokens")
    assert codes.shape == g["codes"].shape and agree0 >= 0.85


def test_encode_chunks_pads_and_reassembles_like_the_reference():
    """CodecEncoderB200.encode (RedCodecInfer.encode, model.py:243-305): ragged items -> 6 s chunks -> batches -> tokens.
    An item's tokens equal the concatenation of its chunks' tokens encoded on their own, token_length = ceil(n / 1280)."""
    from fireredtts2_b200.config import TINY
    from fireredtts2_b200.encoder import ETINYF, synthetic_audio, synthetic_encoder_state_dict, synthetic_front_state_dict
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    esd = dict(synthetic_encoder_state_dict(ETINYF, 2))
    esd.update(synthetic_front_state_dict(ETINYF, 2))
    enc = build_encoder(ETINYF, esd)
    sd = dict(synthetic_state_dict(TINY, 2))
    sd.update(synthetic_encode_tensors(TINY, 2, ETINYF.down_dim))
    codec = build_codec(TINY, sd)
    lens = [7 * 16000 + 123, 13 * 16000 + 8000, 96000]            # 2, 3 and exactly 1 chunk
    audio = synthetic_audio(3, max(lens), 77)
    tok, tok_len = enc.encode(torch.from_numpy(audio).cuda(), torch.tensor(lens), codec, batch_size=4)
    assert tok_len.tolist() == [-(-n // 1280) for n in lens]
    assert tuple(tok.shape) == (3, TINY.num_quantizers, max(tok_len.tolist())) and tok.dtype == torch.int64
    for i, n in enumerate(lens):
        a = np.zeros(-(-n // 96000) * 96000, dtype=np.float32)
        a[:n] = audio[i, :n]
        parts = []
        for c in a.reshape(-1, 96000):
            vq = enc.audio_features(torch.from_numpy(c[None]).cuda())
            parts.append(codec.rvq_encode_codes(vq.transpose(1, 2))[:, 0])          # (nq, 75)
        ref = torch.cat(parts, dim=1)[:, :int(tok_len[i])]
        same = (tok[i, :, :int(tok_len[i])] == ref).float().mean().item()
        # batch neighbours change GEMM tile shapes (fp32 summation order): near-ties may flip, nothing else
        assert same >= 0.98, same
        # as in the reference, positions behind token_length still hold the tokens of the chunk's zero padding
        # (pad_sequence only zero-fills behind an item's LAST chunk, model.py:291-297)
        assert (tok[i, :, 75 * (-(-n // 96000)):] == 0).all()
    with pytest.raises(ValueError):
        enc.audio_features(torch.zeros(1, 1000).cuda())


def test_encoder_head_dim_32_runs_on_the_warp_kernel():
    """head_dim 32 (embed 128 / 4 heads) is served by the CUDA-core attention kernel (8-query blocks: T % 8 == 0)."""
    import dataclasses
    from fireredtts2_b200.encoder import ETINY, synthetic_encoder_state_dict, synthetic_features
    cfg = dataclasses.replace(ETINY, ssl_num_heads=4)
    esd = synthetic_encoder_state_dict(cfg, 6)
    enc = build_encoder(cfg, esd)
    ssl, aco = synthetic_features(cfg, 2, 40, 4)
    vq = enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
    ref = EO.encode_features(esd, ssl, aco, cfg.ssl_num_heads, cfg.avg_pooler)
    _, snr = report("hd32 encoder features vs oracle", ref, vq.cpu().numpy())
    assert snr >= SNR_GATE_DB
    ssl, aco = synthetic_features(cfg, 1, 44, 4)                  # 44 % 8 != 0
    with pytest.raises(ValueError):
        enc.features(torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
