"""A small pass over every kernel family of the library, meant to run UNDER compute-sanitizer (memcheck / synccheck):

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py

Offline decode (tcgen05 / TMA GEMMs, tcgen05 attention, LayerNorm, overlap-add; ragged lengths, scatter), the captured token
step on a side stream right after ``new_stream`` / ``reset_stream`` (skinny GEMMs, split-KV warp attention, stream-ordered
reset), a slot pool step, decode + resample, RVQ encode, the encode features and one frame of the LM's frame tail feeding
the codec step.  Sizes are the smallest presets: the point is address / barrier checking, not numbers.  Prints one line
per stage so a timeout shows how far it got."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from fireredtts2_b200.codec import RedCodecB200, StreamDecoder
from fireredtts2_b200.config import TINY
from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict, synthetic_tokens

T0 = time.time()


def stage(msg):
    torch.cuda.synchronize()
    print(f"[sanitize +{time.time() - T0:5.1f}s] {msg}", flush=True)


def main():
    cfg = TINY
    sd = dict(synthetic_state_dict(cfg, 0))
    sd.update(synthetic_encode_tensors(cfg, 0))
    codec = RedCodecB200(cfg, sd, device="cuda:0", stream_max_tokens=16)
    stage("handle built")
    tok = torch.from_numpy(synthetic_tokens(cfg, 3, 9, 1)).cuda()
    a = codec.decode(tok)
    b = codec.decode(tok, lengths=torch.tensor([9, 4, 1], dtype=torch.int32, device="cuda"))
    assert a.shape == (3, 9 * 8 * cfg.hop_length) and bool(torch.isfinite(a).all()) and bool(torch.isfinite(b).all())
    stage("offline decode, ragged lengths")
    frames = [tok[0, :, i] for i in range(6)]
    for rep in range(2):                        # second round: recycled state, stream-ordered reset
        dec = StreamDecoder(codec, pcm16=True, ring=8)
        chunks = [c for c in (dec.push(f) for f in frames) if c is not None]
        chunks.append(dec.finish())
        for c in chunks:
            c.ready.synchronize()
        assert sum(c.samples.shape[-1] for c in chunks) == 6 * 8 * cfg.hop_length
    cache = codec.new_stream()
    x, cache = codec.decode_one_token(tok[:1, :, :1], cache, False)
    cache = codec.reset_stream(cache)
    y, cache = codec.decode_one_token(tok[:1, :, :1], cache, False)
    assert torch.equal(x, y)
    stage("token step: StreamDecoder on its side stream x2, new_stream / reset_stream")
    pool = codec.new_pool(4, 16)
    s0, s1 = pool.open(), pool.open()
    out = pool.step({s0: tok[0, :, 0], s1: tok[1, :, 0]})
    out = pool.step({s0: tok[0, :, 1], s1: tok[1, :, 1]}, last=(s1,))
    assert set(out) == {s0, s1}
    pool.destroy()
    stage("slot pool")
    from fireredtts2_b200.frame_decoder import (FD_TINY, FrameDecoderB200, synthetic_frame_decoder_state_dict,
                                                synthetic_frame_inputs)
    fd = FrameDecoderB200(FD_TINY, synthetic_frame_decoder_state_dict(FD_TINY, 3), device="cuda:0")
    last_h, noise = synthetic_frame_inputs(FD_TINY, 2, 4)
    c = fd.generate_codes(torch.from_numpy(last_h).cuda(), 8, 0.9, noise=torch.from_numpy(noise).cuda())
    c2 = fd.generate_codes(torch.from_numpy(last_h).cuda(), 8, 0.9)
    fd.check_error()
    assert c.shape == c2.shape == (2, FD_TINY.audio_num_codebooks)
    stage("frame tail")
    a24, a16 = codec.decode_resampled(tok, 16000)
    p = codec.decode(tok, pcm16=True)
    assert a16.shape[0] == 3 and p.dtype == torch.int16
    stage("decode + resample, pcm16")
    z = torch.randn(2, cfg.embed_dim, 12, device="cuda")
    codes = codec.rvq_encode_codes(z)
    assert codes.shape[0] == cfg.num_quantizers
    stage("rvq encode")
    from fireredtts2_b200.encoder import ETINY, CodecEncoderB200, synthetic_encoder_state_dict, synthetic_features
    ssl, aco = synthetic_features(ETINY, 2, 48, 2)
    vq = CodecEncoderB200(ETINY, synthetic_encoder_state_dict(ETINY, 1), device="cuda:0").features(
        torch.from_numpy(ssl).cuda(), torch.from_numpy(aco).cuda())
    assert bool(torch.isfinite(vq).all())
    stage("encode features")
    print("[sanitize] all stages done", flush=True)


if __name__ == "__main__":
    main()
