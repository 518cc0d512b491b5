"""Throughput of the codec ENCODE side behind the feature encoders: SslAdaptor + cat + ResidualDownConv
(frt2_enc_features) and ResidualVQ.encode_codes (frt2_rvq_encode) at the reference's batch shape — `batch_size` 96
chunks of 6 s (model.py:247,262: T = 300 frames at 50 Hz -> 75 tokens per chunk) — EC0 / C0 widths, random weights.
Device time from CUDA events; prints one JSON line.  usage: python tools/encode_bench.py [--batch 96] [--reps 10]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.encoder import EC0, CodecEncoderB200, synthetic_encoder_state_dict, synthetic_features
from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict


def flops(cfg, M):
    E, F, D, P = cfg.ssl_embed_dim, cfg.ffn_dim, cfg.down_dim, cfg.avg_pooler * cfg.down_dim
    f = 2.0 * M * cfg.ssl_in_dim * E + cfg.ssl_num_layers * 2.0 * M * E * (4 * E + 2 * F) + 2.0 * M * E * cfg.ssl_out_dim
    M4 = M / cfg.avg_pooler
    f += 2.0 * M4 * P * (2 * P) + 2.0 * M4 * P * P + 2.0 * M4 * P * D
    return f


def stack_flops(E, F, layers, M, T):
    return layers * (2.0 * M * E * (4 * E + 2 * F) + 4.0 * E * T * M)


def audio_bench(a):
    """frt2_enc_audio_features + frt2_rvq_encode on `batch` 6 s chunks (the reference's _encode_one_batch, model.py:218-236)."""
    from fireredtts2_b200.encoder import EC0F, synthetic_audio, synthetic_front_state_dict
    dev = torch.device("cuda", 0)
    cfg = EC0F
    t0 = time.perf_counter()
    esd = dict(synthetic_encoder_state_dict(cfg, 0))
    esd.update(synthetic_front_state_dict(cfg, 0))
    enc = CodecEncoderB200(cfg, esd, device="cuda:0")
    sd = dict(synthetic_state_dict(C0, 0))
    sd.update(synthetic_encode_tensors(C0, 0, cfg.down_dim))
    codec = RedCodecB200(C0, sd, device="cuda:0", check_indices=False)
    load_s = time.perf_counter() - t0
    n = 96000
    audio = torch.from_numpy(synthetic_audio(a.batch, n, 5)).to(dev)
    for _ in range(2):
        vq = enc.audio_features(audio)
        codes = codec.rvq_encode_codes(vq.transpose(1, 2))
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tq = [], []
    for _ in range(a.reps):
        e[0].record()
        vq = enc.audio_features(audio)
        e[1].record()
        codes = codec.rvq_encode_codes(vq.transpose(1, 2))
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tq.append(e[1].elapsed_time(e[2]))
    Tm, T = n // 160, n // 320
    Mm, M = a.batch * Tm, a.batch * T
    fl = 0.0
    for E, F, L in ((cfg.ssl_in_dim, cfg.ssl_enc_ffn_dim or 4 * cfg.ssl_in_dim, cfg.ssl_enc_layers),
                    (cfg.aco_dim, cfg.aco_ffn_dim or 4 * cfg.aco_dim, cfg.aco_layers)):
        fl += 2.0 * Mm * 3 * cfg.num_mels * E + 2.0 * M * 3 * E * E + stack_flops(E, F, L, M, T)
    fl += flops(cfg, M) + cfg.ssl_num_layers * 4.0 * cfg.ssl_embed_dim * T * M
    ms_f, ms_q = float(np.median(tf)), float(np.median(tq))
    pk = {}
    try:
        pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = pk.get("bf16_tflops_sustained") or 1400.0
    audio_s = a.batch * 6.0
    # parity at the full size: chunk 0 against the numpy oracle (log-mel, both encoders, RVQ input), indices from our
    # features against the oracle's search on the same features
    from oracle import codec_oracle as O
    from oracle import encoder_oracle as EO
    t0 = time.perf_counter()
    taps = {}
    a_np = audio[:1].cpu().numpy()
    ref = EO.encode_audio_features(esd, a_np, cfg, taps=taps)
    vq1, tp = enc.audio_features(audio[:1], taps=True)
    got = vq1.cpu().numpy()
    ref_codes, margin = O.rvq_encode_codes(sd, np.ascontiguousarray(got.transpose(0, 2, 1)))
    codes1 = codec.rvq_encode_codes(vq1.transpose(1, 2)).cpu().numpy()
    parity = {"chunk": 0, "log_mel_snr_db": O.snr_db(taps["mel"], tp["mel"].cpu().numpy()),
              "ssl_encoder_snr_db": O.snr_db(taps["ssl"], tp["ssl"].cpu().numpy()),
              "acoustic_encoder_snr_db": O.snr_db(taps["aco"], tp["aco"].cpu().numpy()),
              "vq_in_feats_snr_db": O.snr_db(ref, got), "gate_snr_db": 40.0,
              "indices_identical_to_oracle_on_gpu_features": float((codes1 == ref_codes).mean()),
              "oracle_seconds": time.perf_counter() - t0}
    print(json.dumps({
        "workload": f"whole encode path from the waveform: {a.batch} chunks x 6 s @16 kHz = {audio_s:.0f} audio-s per batch; "
                    "log-mel, SSL encoder (whisper-large-v3 size: 32 x 1280, 20 heads), acoustic encoder (12 x 768, 8 heads, "
                    "head_dim 96 padded to 128), ssl_adaptor, downsample, C0 RVQ; random weights",
        "features_ms": ms_f, "rvq_encode_ms": ms_q, "total_ms": ms_f + ms_q, "audio_s_per_s": audio_s / ((ms_f + ms_q) * 1e-3),
        "launches_features": enc.last_launches, "algorithmic_flops": fl,
        "roofline": {"bound": "tensor", "achieved": fl / (ms_f * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": fl / (ms_f * 1e-3) / 1e12 / peak},
        "load_seconds": load_s, "codes_shape": list(codes.shape), "parity": parity}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=96)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--audio", action="store_true", help="whole path from the 16 kHz waveform (EC0F: whisper-large-v3 "
                    "sized SSL encoder + 12-layer acoustic encoder), 6 s chunks")
    a = ap.parse_args()
    if a.audio:
        return audio_bench(a)
    dev = torch.device("cuda", 0)
    esd = synthetic_encoder_state_dict(EC0, 0)
    enc = CodecEncoderB200(EC0, esd, device="cuda:0")
    sd = dict(synthetic_state_dict(C0, 0))
    sd.update(synthetic_encode_tensors(C0, 0, EC0.down_dim))
    codec = RedCodecB200(C0, sd, device="cuda:0", check_indices=False)
    ssl, aco = synthetic_features(EC0, a.batch, a.frames, 3)
    ssl, aco = torch.from_numpy(ssl).to(dev), torch.from_numpy(aco).to(dev)
    for _ in range(3):
        codes = enc.encode_features(ssl, aco, codec)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tq = [], []
    for _ in range(a.reps):
        e[0].record()
        vq = enc.features(ssl, aco)
        e[1].record()
        codes = codec.rvq_encode_codes(vq.transpose(1, 2))
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tq.append(e[1].elapsed_time(e[2]))
    M = a.batch * a.frames
    audio_s = a.batch * a.frames / 50.0
    # attention: full mask, 4*hd FLOP per (query, key) pair per head
    attn = EC0.ssl_num_layers * 4.0 * EC0.ssl_embed_dim * a.frames * M
    fl = flops(EC0, M)
    ms_f, ms_q = float(np.median(tf)), float(np.median(tq))
    pk = {}
    try:
        pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = pk.get("bf16_tflops_sustained") or 1400.0
    out = {"workload": f"encode side behind the feature encoders: {a.batch} chunks x {a.frames} frames (50 Hz) = "
                       f"{audio_s:.0f} audio-s per batch; EC0 (ssl_adaptor 1280->768 x{EC0.ssl_num_layers} layers ->256, "
                       "acoustic 768, downsample 1024 x4) + C0 RVQ (16 x 2048 x 256)",
           "features_ms": ms_f, "rvq_encode_ms": ms_q, "total_ms": ms_f + ms_q,
           "audio_s_per_s": audio_s / ((ms_f + ms_q) * 1e-3), "launches_features": enc.last_launches,
           "features_tflops": (fl + attn) / (ms_f * 1e-3) / 1e12, "features_gemm_flops": fl, "features_attn_flops": attn,
           "roofline": {"bound": "tensor", "achieved": (fl + attn) / (ms_f * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": (fl + attn) / (ms_f * 1e-3) / 1e12 / peak},
           "codes_shape": list(codes.shape)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
