"""Frame tail of the speech LM on the library (csrc/frame_decoder.cu through the C ABI) against the oracle and the goldens
recorded from the reference's own Model.generate_frame (llm.py:274-330)."""
import numpy as np
import pytest
import torch

from fireredtts2_b200.frame_decoder import (FD_PRESETS, FrameDecoderB200, sample_topk, synthetic_frame_decoder_state_dict,
                                            synthetic_frame_inputs)
from oracle import codec_oracle as O
from oracle import frame_decoder_oracle as FO
from tests.gpu_common import report, to_np
from tests.test_oracle_golden import FD_CASES, load_fd_case

pytestmark = pytest.mark.gpu

SNR_GATE_DB = 40.0      # fp16 operands / fp32 accumulation against the fp32 reference (north_star tolerance)

_cache = {}


def build(preset, wseed):
    key = (preset, wseed)
    if key not in _cache:
        cfg = FD_PRESETS[preset]
        sd = synthetic_frame_decoder_state_dict(cfg, wseed)
        _cache.clear()          # one set of full-size weights at a time
        _cache[key] = (cfg, sd, FrameDecoderB200(cfg, sd))
    return _cache[key]


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


@pytest.mark.parametrize("name,preset", FD_CASES)
def test_frame_tail_vs_reference_golden(name, preset):
    """Teacher-forced logits within the SNR gate of the reference's; the free-running frame reproduces the reference's codes
    from the reference's own Exp(1) draws."""
    cfg0, _, g, topk, temperature = load_fd_case(name, preset)
    cfg, sd, fd = build(preset, int(g["meta"][1]))
    last_h, noise = cuda(g["last_h"]), cuda(g["noise"])
    codes_f, logits = fd.generate_codes(last_h, topk, temperature, noise=noise, forced=cuda(g["codes"]), return_logits=True)
    assert np.array_equal(to_np(codes_f).astype(np.int32), g["codes"])
    _, snr = report(f"{name} teacher-forced logits", g["logits"], to_np(logits))
    assert snr >= SNR_GATE_DB
    codes = fd.generate_codes(last_h, topk, temperature, noise=noise)
    fd.check_error()
    got = codes.cpu().numpy()
    same = got == g["codes"]
    print(f"[parity] {name}: free-running codes equal to the reference's: {int(same.sum())} / {same.size}")
    assert same.all()
    # staging + codebook-0 head + projection of last_h + per decoder pass 5 kernels per layer + sampler per codebook (it
    # also writes the next position's projected embedding row) + head per codebook >= 1; positions 0 and 1 are ONE pass while 2 B rows fit the GEMM's 16
    passes = cfg.audio_num_codebooks - (1 if 2 * last_h.shape[0] <= 16 else 0)
    assert fd.last_launches == 3 + passes * 5 * cfg.num_layers + cfg.audio_num_codebooks + (cfg.audio_num_codebooks - 1)


def test_given_c0_and_batch_independence():
    """c0 from the caller skips the codebook-0 head (logits row 0 untouched = zeros here); an item's codes and logits do not
    depend on its neighbours in the batch."""
    cfg, sd, fd = build("FD_SMALL", 5)
    last_h, noise = synthetic_frame_inputs(cfg, 4, seed=7)
    c0 = np.array([3, 500, 17, 0], np.int32)
    ref_codes, ref_logits = FO.generate_codes(sd, cfg, last_h, 25, 0.85, noise, c0=c0)
    codes, logits = fd.generate_codes(cuda(last_h), 25, 0.85, c0=cuda(c0), noise=cuda(noise), return_logits=True)
    assert np.array_equal(codes.cpu().numpy()[:, 0], c0)
    assert np.array_equal(codes.cpu().numpy(), ref_codes)
    _, snr = report("FD_SMALL given-c0 logits", ref_logits[:, 1:], to_np(logits)[:, 1:])
    assert snr >= SNR_GATE_DB
    for b in (0, 3):
        cb, lb = fd.generate_codes(cuda(last_h[b:b + 1]), 25, 0.85, c0=cuda(c0[b:b + 1]), noise=cuda(noise[b:b + 1]),
                                   return_logits=True)
        assert torch.equal(cb[0], codes[b]) and torch.equal(lb[0, 1:], logits[b, 1:])


def test_two_row_first_pass_and_single_row_layout_agree():
    """Batch <= 8: positions 0 and 1 run as one two-row pass (2 B rows fit the MMA tile's 16); batch 9..16: one position
    per pass.  The single-row layout against the oracle (teacher-forced), then the same items through both layouts: GEMM
    rows and attention queries are independent, so codes and logits are identical bit for bit."""
    cfg, sd, fd = build("FD_SMALL", 5)
    last_h, noise = synthetic_frame_inputs(cfg, 11, seed=11)
    ref_codes, ref_logits = FO.generate_codes(sd, cfg, last_h, 20, 0.9, noise)
    n, L = cfg.audio_num_codebooks, cfg.num_layers
    _, forced = fd.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise), forced=cuda(ref_codes), return_logits=True)
    assert fd.last_launches == 3 + n * 5 * L + n + (n - 1)                       # one position per pass
    _, snr = report("FD_SMALL batch 11 (one position per pass, 16-row MMA tiles) logits", ref_logits, to_np(forced))
    assert snr >= SNR_GATE_DB
    codes11, logits11 = fd.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise), return_logits=True)
    codes3, logits3 = fd.generate_codes(cuda(last_h[8:11]), 20, 0.9, noise=cuda(noise[8:11]), return_logits=True)
    assert fd.last_launches == 3 + (n - 1) * 5 * L + n + (n - 1)                 # positions 0 and 1 in one pass
    assert torch.equal(codes3, codes11[8:11]) and torch.equal(logits3, logits11[8:11])
    same = (codes11.cpu().numpy() == ref_codes).all(axis=1)
    print(f"[parity] FD_SMALL batch 11: free-running frames identical to the oracle's: {int(same.sum())} / {same.size}")
    assert same.sum() >= 8          # a frame may leave the oracle's trajectory at an fp16-vs-fp32 near-tie of the sampler


def test_qwen_500m_decoder_flavor():
    """The flavor pair of the reference's own example (llm.py:356-357: qwen-3b backbone width, qwen-500m decoder: 24 layers
    x 896, 14 / 2 heads of 64, 4864): 7 query heads per kv group, a 1817-kernel graph; against the oracle."""
    cfg, sd, fd = build("FD_500M", 2)
    last_h, noise = synthetic_frame_inputs(cfg, 1, seed=3)
    ref_codes, ref_logits = FO.generate_codes(sd, cfg, last_h, 30, 0.9, noise)
    _, logits = fd.generate_codes(cuda(last_h), 30, 0.9, noise=cuda(noise), forced=cuda(ref_codes), return_logits=True)
    _, snr = report("FD_500M teacher-forced logits", ref_logits, to_np(logits))
    assert snr >= SNR_GATE_DB
    codes = fd.generate_codes(cuda(last_h), 30, 0.9, noise=cuda(noise)).cpu().numpy()
    print(f"[parity] FD_500M: free-running codes equal to the oracle's: {int((codes == ref_codes).sum())} / {codes.size}")
    assert (codes[:, :4] == ref_codes[:, :4]).all()


@pytest.mark.parametrize("preset,wseed,B,per_layer", [("FD_SMALL", 5, 24, 9), ("FD_200M", 0, 20, None)])
def test_large_batch_runs_on_the_tcgen05_gemm(preset, wseed, B, per_layer):
    """max_batch > 16: frames of 17 .. max_batch items (a pool of concurrent streams) take the tcgen05 GEMM path (row-major
    weight copies, RMSNorm / SwiGLU row kernels; at the qwen-200m widths the reductions of the narrow layers are split over
    batch items of one launch + a fixed-order reduce); against the oracle, and the same items through the <= 16 path."""
    cfg = FD_PRESETS[preset]
    sd = synthetic_frame_decoder_state_dict(cfg, wseed)
    _cache.clear()
    fd = FrameDecoderB200(cfg, sd, max_batch=40)
    last_h, noise = synthetic_frame_inputs(cfg, B, seed=13)
    ref_codes, ref_logits = FO.generate_codes(sd, cfg, last_h, 20, 0.9, noise)
    _, logits = fd.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise), forced=cuda(ref_codes), return_logits=True)
    n, L = cfg.audio_num_codebooks, cfg.num_layers
    if per_layer is not None:      # positions 0 and 1 in one pass on this path too: n - 1 decoder passes
        assert fd.last_launches == 3 + (n - 1) * per_layer * L + n + 2 * (n - 1)
    _, snr = report(f"{preset} batch {B} (tcgen05 GEMM path) teacher-forced logits", ref_logits, to_np(logits))
    assert snr >= SNR_GATE_DB
    codes = fd.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise))
    fd.check_error()
    same = (codes.cpu().numpy() == ref_codes).all(axis=1)
    print(f"[parity] {preset} batch {B}: free-running frames identical to the oracle's: {int(same.sum())} / {same.size}")
    assert same.sum() >= B - 4
    _, small = fd.generate_codes(cuda(last_h[:8]), 20, 0.9, noise=cuda(noise[:8]), forced=cuda(ref_codes[:8]), return_logits=True)
    _, snr = report("tcgen05 path vs weight-streaming path, same items", to_np(small), to_np(logits[:8]))
    assert snr >= 55.0
    with pytest.raises(ValueError):
        fd.generate_codes(cuda(np.zeros((41, cfg.backbone_dim), np.float32)), 20, 0.9)


def test_gemm_skinny_fallback_for_widths_the_stream_kernel_does_not_take(monkeypatch):
    """Widths that are not multiples of 32 (or FRT2_FD_SKINNY=1, used here) run every GEMM of the frame on gemm_skinny with
    its fused RMSNorm / SwiGLU modes (row-major weights): same parity, and 60+ dB against the stream-kernel instance."""
    cfg = FD_PRESETS["FD_SMALL"]
    sd = synthetic_frame_decoder_state_dict(cfg, 5)
    last_h, noise = synthetic_frame_inputs(cfg, 3, seed=17)
    ref_codes, ref_logits = FO.generate_codes(sd, cfg, last_h, 20, 0.9, noise)
    monkeypatch.setenv("FRT2_FD_SKINNY", "1")
    fb = FrameDecoderB200(cfg, sd)
    monkeypatch.delenv("FRT2_FD_SKINNY")
    _, lg_fb = fb.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise), forced=cuda(ref_codes), return_logits=True)
    _, snr = report("FD_SMALL on the gemm_skinny fallback", ref_logits, to_np(lg_fb))
    assert snr >= SNR_GATE_DB
    _, _, fd = build("FD_SMALL", 5)
    _, lg = fd.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise), forced=cuda(ref_codes), return_logits=True)
    _, snr = report("fallback vs stream kernel", to_np(lg), to_np(lg_fb))
    assert snr >= 60.0


def test_adversarial_weights_keep_the_gate():
    """Massive residual channels (x 50), RMSNorm scales in [0.1, 5], wide SwiGLU products: teacher-forced logits of the
    fp16-operand path stay within the gate of the fp32 oracle."""
    from fireredtts2_b200.frame_decoder import adversarial_frame_decoder_state_dict
    cfg = FD_PRESETS["FD_SMALL"]
    sd = adversarial_frame_decoder_state_dict(cfg, 1)
    _cache.clear()
    fd = FrameDecoderB200(cfg, sd)
    last_h, noise = synthetic_frame_inputs(cfg, 2, seed=19)
    ref_codes, ref_logits = FO.generate_codes(sd, cfg, last_h, 20, 0.9, noise)
    _, logits = fd.generate_codes(cuda(last_h), 20, 0.9, noise=cuda(noise), forced=cuda(ref_codes), return_logits=True)
    _, snr = report("FD_SMALL adversarial weights, teacher-forced logits", ref_logits[:, 1:], to_np(logits)[:, 1:])
    assert snr >= SNR_GATE_DB


def test_out_of_range_code_raises_index_error():
    cfg, sd, fd = build("FD_TINY", 3)
    last_h, noise = synthetic_frame_inputs(cfg, 2, seed=1)
    forced = np.zeros((2, cfg.audio_num_codebooks), np.int32)
    forced[1, 2] = cfg.audio_vocab_size       # nn.Embedding would raise IndexError (llm.py:336-337)
    fd.generate_codes(cuda(last_h), 5, 1.0, noise=cuda(noise), forced=cuda(forced))
    with pytest.raises(IndexError):
        fd.check_error()
    fd.generate_codes(cuda(last_h), 5, 1.0, noise=cuda(noise))
    fd.check_error()                           # the word was cleared
    with pytest.raises(ValueError):
        fd.generate_codes(cuda(last_h[:, :-8]), 5, 1.0)
    with pytest.raises(ValueError):
        fd.generate_codes(cuda(last_h), 0, 1.0)


def test_library_generator_is_counter_based():
    """noise=None: Philox draws keyed by (seed, frame counter): a fresh handle with the same seed repeats the sequence of
    frames, another seed does not; consecutive frames differ."""
    cfg = FD_PRESETS["FD_TINY"]
    sd = synthetic_frame_decoder_state_dict(cfg, 3)
    last_h, _ = synthetic_frame_inputs(cfg, 2, seed=2)
    runs = []
    for seed in (11, 11, 12):
        fd = FrameDecoderB200(cfg, sd)
        runs.append(torch.stack([fd.generate_codes(cuda(last_h), cfg.audio_vocab_size, 1.5, seed=seed) for _ in range(6)]).cpu())
    assert torch.equal(runs[0], runs[1])
    assert not torch.equal(runs[0], runs[2])
    assert not torch.equal(runs[0][0], runs[0][1])
    assert int(runs[0].min()) >= 0 and int(runs[0].max()) < cfg.audio_vocab_size


@pytest.mark.parametrize("V,topk,temperature", [(64, 8, 0.9), (2048, 10, 0.75), (2048, 30, 0.9), (2051, 50, 1.0), (512, 512, 1.3),
                                                (2048, 1, 0.5)])
def test_sample_topk_matches_oracle(V, topk, temperature):
    """Index work: identical to the restated llm.py:34-49 on the same logits and draws, duplicates at the k-th value included."""
    rng = np.random.default_rng(V + topk)
    B = 64
    logits = (2.0 * rng.standard_normal((B, V))).astype(np.float32)
    logits[: B // 2] = np.round(logits[: B // 2] * 2) / 2          # many exact ties, also at the k-th value
    noise = np.maximum(rng.exponential(1.0, (B, V)), 1e-30).astype(np.float32)
    ref = FO.sample_topk(logits, topk, temperature, noise)
    got = sample_topk(cuda(logits), topk, temperature, cuda(noise)).cpu().numpy()
    assert np.array_equal(got, ref)


def test_sample_topk_matches_the_reference_sampler_golden():
    """fd_sample_kernel against decisions of the reference's own sample_topk (llm.py:39-49) with the draws it made."""
    import os
    from tests.helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "fd_sampler.npz"))
    for i in range(int(g["n_cases"])):
        V, topk, rows = (int(v) for v in g[f"c{i}_meta"])
        got = sample_topk(cuda(g[f"c{i}_logits"]), topk, float(g[f"c{i}_temperature"]), cuda(g[f"c{i}_q"])).cpu().numpy()
        assert np.array_equal(got, g[f"c{i}_codes"]), (i, V, topk)


def test_sample_topk_library_draws_follow_the_distribution():
    """noise=None: the sampled frequencies follow softmax(logits / T) restricted to the top k (chi-square-like bound)."""
    V, topk, T, B = 64, 8, 0.8, 20000
    rng = np.random.default_rng(5)
    row = (1.5 * rng.standard_normal(V)).astype(np.float32)
    logits = np.tile(row, (B, 1))
    got = sample_topk(cuda(logits), topk, T, None, seed=99).cpu().numpy()
    s = row / np.float32(T)
    keep = s >= np.sort(s)[::-1][topk - 1]
    p = np.where(keep, np.exp(s - s.max()), 0.0)
    p /= p.sum()
    freq = np.bincount(got, minlength=V) / B
    assert freq[~keep].sum() == 0.0
    assert np.abs(freq - p).max() < 4.0 * np.sqrt(0.25 / B)


def test_codes_feed_the_codec_step_on_the_device():
    """generate_frame -> decode_one_token (fireredtts2.py:303-326 of generate_stream): the frame's codes go from the frame
    decoder to the codec's streaming step as a strided int32 device view, frame after frame on one stream; the audio equals
    the oracle's streaming decode of the same codes."""
    from fireredtts2_b200.codec import RedCodecB200
    from fireredtts2_b200.config import TINY
    from fireredtts2_b200.weights import synthetic_state_dict
    cfg, sd, fd = build("FD_TINY", 3)
    csd = synthetic_state_dict(TINY, 0)
    codec = RedCodecB200(TINY, csd, device="cuda:0")
    nq = TINY.num_quantizers
    assert cfg.audio_vocab_size == TINY.codebook_size and nq <= cfg.audio_num_codebooks
    rng = np.random.default_rng(9)
    cache, chunks, frames = {}, [], []
    n_frames = 5
    for i in range(n_frames):
        last_h = cuda(rng.standard_normal((1, cfg.backbone_dim)).astype(np.float32))
        codes = fd.generate_codes(last_h, 16, 1.0, seed=4)                       # (1, ncb) int32, stays on the device
        audio, cache = codec.decode_one_token(codes[:, :nq].unsqueeze(-1), cache, i == n_frames - 1)
        frames.append(codes)
        chunks.append(audio)
    fd.check_error()
    tok = torch.stack(frames, dim=-1)[:, :nq].cpu().numpy().astype(np.int64)      # (1, nq, n_frames)
    state, ref = None, []
    for i in range(n_frames):
        a, state = O.decode_chunk(csd, tok[:, :, i:i + 1], state, i == n_frames - 1, TINY.num_heads, TINY.hop_length)
        ref.append(a)
    _, snr = report("frame tail -> codec step audio", np.concatenate(ref, axis=1), to_np(torch.cat(chunks, dim=1)))
    assert snr >= SNR_GATE_DB


def test_frame_tail_feeds_a_pool_of_concurrent_streams():
    """One frame for all slots -> frt2_pool_step: every slot's chunks equal the oracle's streaming decode of that slot's
    codes (continuous batching behind the frame tail, no host round trip of the codes)."""
    from fireredtts2_b200 import _native as NN
    from fireredtts2_b200.codec import RedCodecB200
    from fireredtts2_b200.config import TINY
    from fireredtts2_b200.weights import synthetic_state_dict
    cfg, sd, fd = build("FD_TINY", 3)
    csd = synthetic_state_dict(TINY, 0)
    codec = RedCodecB200(TINY, csd, device="cuda:0")
    nq, slots, steps = TINY.num_quantizers, 3, 4
    pool = codec.new_pool(slots)
    rng = np.random.default_rng(10)
    frames, chunks = [], [[] for _ in range(slots)]
    for i in range(steps):
        last_h = cuda(rng.standard_normal((slots, cfg.backbone_dim)).astype(np.float32))
        codes = fd.generate_codes(last_h, 16, 1.0, seed=5)
        flags = [NN.SLOT_ACTIVE | (NN.SLOT_RESET if i == 0 else 0) | (NN.SLOT_LAST if i == steps - 1 else 0)] * slots
        out, n = pool.step_dense(codes[:, :nq], flags)
        frames.append(codes)
        for s in range(slots):
            chunks[s].append(out[s, :n[s]])
    tok = torch.stack(frames, dim=-1)[:, :nq].cpu().numpy().astype(np.int64)       # (slots, nq, steps)
    for s in range(slots):
        state, ref = None, []
        for i in range(steps):
            a, state = O.decode_chunk(csd, tok[s:s + 1, :, i:i + 1], state, i == steps - 1, TINY.num_heads, TINY.hop_length)
            ref.append(a)
        _, snr = report(f"frame tail -> pool slot {s}", np.concatenate(ref, axis=1)[0], to_np(torch.cat(chunks[s])))
        assert snr >= SNR_GATE_DB


def test_generate_frame_drop_in_on_a_reference_shaped_model():
    """GenerateFrameB200.install(model) on a module tree with the reference's names (Model.projection / audio_embeddings /
    codebook0_head / audio_head / decoder.layers[i], torchtune parameter names) living on the GPU: the tail is built by
    from_reference from the live tensors, ``model.generate_frame(tokens, tokens_mask, input_pos, temperature, topk)``
    (llm.py:274-330's signature) runs the model's own embedding sum and backbone and hands ``h[:, -1, :]`` to the library.
    The backbone is a stand-in (out of scope); the codes must equal, bit for bit, those of the library called directly on
    the same last_h, and the oracle's.  (The same host logic against the REAL generate_frame: tests/test_host.py.)"""
    from fireredtts2_b200.frame_decoder import GenerateFrameB200
    cfg, sd, fd = build("FD_TINY", 3)
    n, V, D = cfg.audio_num_codebooks, cfg.audio_vocab_size, cfg.backbone_dim
    dev = torch.device("cuda:0")

    class Holder(torch.nn.Module):
        pass

    def grow(root, dotted, tensor):
        *path, leaf = dotted.split(".")
        m = root
        for p in path:
            if not hasattr(m, p):
                setattr(m, p, Holder())
            m = getattr(m, p)
        setattr(m, leaf, torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(tensor))))

    class Backbone(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(11)
            self.w = torch.nn.Parameter(torch.randn(D, D, generator=g) / D ** 0.5)
            self.enabled, self.seen = True, []

        def caches_are_enabled(self):
            return self.enabled

        def forward(self, h, input_pos=None, mask=None):
            self.seen.append((tuple(h.shape), tuple(mask.shape)))
            return torch.tanh(h @ self.w) + 0.01 * input_pos.unsqueeze(-1).to(h.dtype)

    class Model(Holder):
        def _embed_tokens(self, tokens):                                   # llm.py:339-352
            text = torch.nn.functional.embedding(tokens[:, :, -1], self.text_embeddings.weight).unsqueeze(-2)
            idx = tokens[:, :, :-1] + V * torch.arange(n, device=tokens.device)
            audio = torch.nn.functional.embedding(idx.reshape(-1), self.audio_embeddings.weight).reshape(
                tokens.size(0), tokens.size(1), n, -1)
            return torch.cat([audio, text], dim=-2)

    model = Model()
    model.decoder = Holder()
    model.decoder.layers = torch.nn.ModuleList([Holder() for _ in range(cfg.num_layers)])
    for k, v in sd.items():
        if k.startswith("decoder.layers."):
            _, _, i, rest = k.split(".", 3)
            grow(model.decoder.layers[int(i)], rest, v)
        else:
            grow(model, k, v)
    for l in model.decoder.layers:
        l.attn.num_heads = cfg.num_heads
    model.config = type("Cfg", (), {"audio_vocab_size": V, "audio_num_codebooks": n})()
    model.backbone = Backbone()
    g = torch.Generator().manual_seed(12)
    grow(model, "text_embeddings.weight", torch.randn(32, D, generator=g).numpy())
    model.register_buffer("backbone_causal_mask", torch.tril(torch.ones(64, 64, dtype=torch.bool)))
    model = model.to(dev)

    gen = GenerateFrameB200.install(model, seed=5)
    assert model.generate_frame is gen and gen.tail.cfg == cfg

    B, S = 2, 5
    tokens = torch.zeros(B, S, n + 1, dtype=torch.long)
    tokens[:, :, :n] = torch.randint(0, V, (B, S, n), generator=g)
    tokens[:, :, n] = torch.randint(0, 32, (B, S), generator=g)
    mask = torch.ones(B, S, n + 1, dtype=torch.bool)
    tokens, mask = tokens.to(dev), mask.to(dev)
    pos = torch.arange(S, device=dev)[None].repeat(B, 1)
    rng = np.random.default_rng(13)
    for f in range(3):
        noise = rng.exponential(1.0, (B, n, V)).astype(np.float32)
        with torch.inference_mode():
            h = (model._embed_tokens(tokens) * mask.unsqueeze(-1)).sum(dim=2)
            last_h = (torch.tanh(h @ model.backbone.w) + 0.01 * pos.unsqueeze(-1).float())[:, -1, :]
            gen.noise = cuda(noise)
            sample = model.generate_frame(tokens, mask, pos, 0.9, 8)
        assert sample.dtype == torch.int32 and tuple(sample.shape) == (B, n) and sample.device == tokens.device
        direct = fd.generate_codes(last_h, 8, 0.9, noise=cuda(noise))
        assert torch.equal(sample, direct)
        ref_codes, _ = FO.generate_codes(sd, cfg, to_np(last_h), 8, 0.9, noise)
        same = to_np(sample) == ref_codes
        print(f"[parity] generate_frame drop-in, frame {f}: codes equal to the oracle's: {int(same.sum())} / {same.size}")
        assert same.mean() >= 0.9
        tokens = torch.cat([sample, torch.zeros(B, 1, device=dev).long()], dim=1).unsqueeze(1)          # fireredtts2.py:183-191
        mask = torch.cat([torch.ones_like(sample).bool(), torch.zeros(B, 1, device=dev).bool()], dim=1).unsqueeze(1)
        pos = pos[:, -1:] + 1
    assert model.backbone.seen[0] == ((B, S, D), (B, S, 64)) and model.backbone.seen[-1] == ((B, 1, D), (B, 1, 64))
    # library draws: reproducible for a seed, the frame counter advances them
    a = model.generate_frame(tokens, mask, pos, 0.9, 8)
    b = model.generate_frame(tokens, mask, pos, 0.9, 8)
    assert a.shape == b.shape and not torch.equal(a, b)
    gen.tail.check_error()
    model.backbone.enabled = False
    with pytest.raises(AssertionError):
        model.generate_frame(tokens, mask, pos, 0.9, 8)
    gen.uninstall()
    assert "generate_frame" not in model.__dict__
