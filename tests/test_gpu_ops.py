"""Single-kernel parity tests (GPU): each hand-written kernel through the C ABI against a torch fp32
reference of the same op on the same fp16-rounded operands, and tcgen05 kernels against the SIMT/warp
check kernels."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from fireredtts2_b200 import _native as N
    return N.load()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(lib, status):
    from fireredtts2_b200 import _native as N
    N.check(status)


def gemm_reference(A, W, taps, bias, act, resid, alpha):
    """A (b, rows, Kc) fp16, W (N, taps*Kc) fp16 -> fp32 (b, rows, N); causal taps with zero left padding."""
    b, rows, Kc = A.shape
    Af = A.float()
    Wf = W.float()
    out = torch.zeros(b, rows, W.shape[0], device=A.device)
    for tap in range(taps):
        shift = taps - 1 - tap
        x = torch.zeros_like(Af)
        if shift < rows:
            x[:, shift:, :] = Af[:, :rows - shift, :]
        out += x @ Wf[:, tap * Kc:(tap + 1) * Kc].T
    out = out * alpha
    if bias is not None:
        out = out + bias
    if act == 1:
        out = torch.nn.functional.gelu(out)
    elif act == 2:
        mag = torch.clamp(torch.exp(out[..., 0::2]), max=100.0)
        ph = out[..., 1::2]
        out = torch.stack([mag * torch.cos(ph), mag * torch.sin(ph)], dim=-1).reshape(out.shape)
    if resid is not None:
        out = out + resid
    return out


GEMM_CASES = [
    # batches, rows, Kc, taps, N, bias, act, resid, alpha
    (1, 128, 64, 1, 128, False, 0, False, 1.0),
    (1, 300, 128, 1, 256, True, 0, False, 1.0),
    (1, 1000, 1024, 1, 1024, True, 1, False, 1.0),
    (1, 777, 1024, 1, 3072, True, 0, False, 1.0),
    (1, 512, 4096, 1, 1024, True, 0, True, 1.0),
    (2, 77, 128, 3, 128, True, 0, True, 1.0),
    (3, 200, 256, 7, 256, True, 0, False, 1.0),
    (2, 500, 1024, 2, 2048, True, 1, False, 1.0),
    (1, 1000, 1024, 1, 962, True, 2, False, 1.0),
    (1, 640, 1024, 1, 960, False, 0, False, 1.0 / 960),
    (4, 8, 1024, 3, 1024, True, 0, True, 1.0),
    # packed-item tiles: several short items share one 128-row tile (one 3-D TMA box {64, rows, items})
    (20, 8, 256, 3, 256, True, 0, True, 1.0),
    (64, 8, 1024, 7, 1024, True, 0, False, 1.0),
    (130, 1, 512, 1, 512, False, 0, False, 1.0),
    (5, 24, 128, 3, 192, True, 1, False, 1.0),
    (3, 64, 128, 2, 962, True, 2, False, 1.0),
    (33, 4, 1024, 2, 2048, True, 1, False, 1.0),
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: "b%d_m%d_k%d_t%d_n%d_b%d_a%d_r%d" % c[:8])
def test_gemm_tc(lib, case):
    batches, rows, Kc, taps, Nn, use_bias, act, use_resid, alpha = case
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(batches, rows, Kc, device="cuda", generator=g)).half()
    W = (torch.randn(Nn, taps * Kc, device="cuda", generator=g) / math.sqrt(taps * Kc)).half()
    if act == 2:
        W = W * 0.5
    bias = torch.randn(Nn, device="cuda", generator=g) * 0.1 if use_bias else None
    resid = torch.randn(batches, rows, Nn, device="cuda", generator=g) if use_resid else None
    outs = {}
    for impl in (0, 1):
        o32 = torch.full((batches, rows, Nn), float("nan"), device="cuda")
        o16 = torch.full((batches, rows, Nn), float("nan"), device="cuda", dtype=torch.half) if Nn % 8 == 0 else None
        _check(lib, lib.frt2_op_gemm(impl, _p(A), _p(W), batches, rows, Kc, taps, Nn, alpha, _p(bias), act,
                                     _p(resid), _p(o32), _p(o16), _stream()))
        torch.cuda.synchronize()
        outs[impl] = (o32, o16)
    ref = gemm_reference(A, W, taps, bias, act, resid, alpha)
    scale = ref.abs().max().item() + 1e-6
    e_ref = (outs[1][0] - ref).abs().max().item() / scale
    e_tc = (outs[0][0] - ref).abs().max().item() / scale
    e_tc_vs_simt = (outs[0][0] - outs[1][0]).abs().max().item() / scale
    print(f"gemm {case}: simt-vs-torch {e_ref:.2e} tc-vs-torch {e_tc:.2e} tc-vs-simt {e_tc_vs_simt:.2e}")
    assert torch.isfinite(outs[0][0]).all()
    assert e_ref < 2e-3, "SIMT check kernel disagrees with torch"
    assert e_tc < 2e-3, "tcgen05 GEMM disagrees with torch"
    if outs[0][1] is not None:
        e16 = (outs[0][1].float() - ref).abs().max().item() / scale
        assert e16 < 3e-3


@pytest.mark.parametrize("rows,Cc,silu,eps", [(1000, 1024, 0, 1e-5), (333, 128, 1, 1e-5), (64, 64, 0, 1e-6),
                                                (17, 2048, 1, 1e-5)])
def test_layer_norm(lib, rows, Cc, silu, eps):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(rows, Cc, device="cuda", generator=g) * 3 + 0.5
    gamma = torch.randn(Cc, device="cuda", generator=g)
    beta = torch.randn(Cc, device="cuda", generator=g)
    out = torch.empty(rows, Cc, device="cuda", dtype=torch.half)
    _check(lib, lib.frt2_op_layer_norm(_p(x), rows, Cc, _p(gamma), _p(beta), eps, silu, _p(out), _stream()))
    ref = torch.nn.functional.layer_norm(x, (Cc,), gamma, beta, eps)
    if silu:
        ref = torch.nn.functional.silu(ref)
    err = (out.float() - ref).abs().max().item()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), err   # fp16 output rounding


def attention_reference(q, k, v, H, q_pos0, block_causal):
    B, Tq, E = q.shape
    Tk = k.shape[1]
    hd = E // H
    qh = q.float().view(B, Tq, H, hd).transpose(1, 2)
    kh = k.float().view(B, Tk, H, hd).transpose(1, 2)
    vh = v.float().view(B, Tk, H, hd).transpose(1, 2)
    mask = None
    if block_causal:
        qi = torch.arange(Tq, device=q.device) + q_pos0
        kj = torch.arange(Tk, device=q.device)
        mask = kj[None, :] <= (qi[:, None] | 7)
    o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask)
    return o.transpose(1, 2).reshape(B, Tq, E)


ATTN_CASES = [
    # B, H, hd, Tq, Tk, q_pos0, block_causal
    (1, 2, 64, 72, 72, 0, 1),
    (2, 4, 64, 1000, 1000, 0, 1),
    (1, 16, 64, 3000, 3000, 0, 1),
    (2, 2, 64, 32, 232, 200, 0),
    (1, 2, 64, 8, 8, 0, 0),
    (3, 2, 64, 8, 808, 800, 0),
    (1, 2, 64, 128, 128, 0, 1),
    (1, 1, 64, 64, 320, 256, 0),
    (1, 2, 64, 200, 200, 0, 1),
    (2, 1, 64, 264, 264, 0, 1),
    (1, 1, 64, 136, 392, 256, 0),
    (1, 3, 64, 520, 520, 0, 1),
    # more work items than SMs: every CTA of the persistent kernel walks through several (group, head, item) items
    (4, 16, 64, 1504, 1504, 0, 1),
    (40, 8, 64, 264, 264, 0, 1),
    # head dim 128 (two query tiles per CTA)
    (1, 2, 128, 72, 72, 0, 1),
    (2, 8, 128, 1000, 1000, 0, 1),
    (3, 8, 128, 1504, 1504, 0, 1),
    (2, 2, 128, 32, 232, 200, 0),
    (1, 1, 128, 136, 392, 256, 0),
    (24, 8, 128, 200, 200, 0, 1),
]


@pytest.mark.parametrize("case", ATTN_CASES, ids=lambda c: "B%d_H%d_hd%d_q%d_k%d_p%d_bc%d" % c)
@pytest.mark.parametrize("impl", [1, 0], ids=["warp", "tc"])
def test_attention(lib, impl, case):
    B, H, hd, Tq, Tk, q_pos0, bc = case
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B, Tq, E, device="cuda", generator=g).half()
    k = torch.randn(B, Tk, E, device="cuda", generator=g).half()
    v = torch.randn(B, Tk, E, device="cuda", generator=g).half()
    out = torch.full((B, Tq, E), float("nan"), device="cuda", dtype=torch.half)
    _check(lib, lib.frt2_op_attention(impl, _p(q), _p(k), _p(v), _p(out), B, H, hd, Tq, Tk, q_pos0, bc, _stream()))
    torch.cuda.synchronize()
    ref = attention_reference(q, k, v, H, q_pos0, bc)
    err = (out.float() - ref).abs().max().item()
    print(f"attention impl={impl} {case}: max-abs err {err:.3e}")
    assert torch.isfinite(out.float()).all()
    assert err < (2e-3 if impl == 1 else 4e-3), err


@pytest.mark.parametrize("case", [(3, 4, 64, 300, 300), (2, 12, 64, 300, 300), (2, 2, 128, 300, 300), (1, 2, 64, 52, 52),
                                  (5, 20, 64, 1500, 1500)],
                         ids=lambda c: "B%d_H%d_hd%d_q%d_k%d" % c)
def test_attention_tc_full_mask_odd_lengths(lib, case):
    """The encode side's attention (whisper.py:49-79 under make_nonpad_mask of full-length items): no mask, sequence
    lengths that are not multiples of 8 (T = 300 frames per 6 s chunk, 1500 per 30 s) — tcgen05 kernel only (the warp
    kernel works on 8-query blocks)."""
    B, H, hd, Tq, Tk = case
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(B, Tq, E, device="cuda", generator=g).half()
    k = torch.randn(B, Tk, E, device="cuda", generator=g).half()
    v = torch.randn(B, Tk, E, device="cuda", generator=g).half()
    out = torch.full((B, Tq, E), float("nan"), device="cuda", dtype=torch.half)
    _check(lib, lib.frt2_op_attention(0, _p(q), _p(k), _p(v), _p(out), B, H, hd, Tq, Tk, 0, 0, _stream()))
    torch.cuda.synchronize()
    ref = attention_reference(q, k, v, H, 0, 0)
    err = (out.float() - ref).abs().max().item()
    print(f"attention tc full mask {case}: max-abs err {err:.3e}")
    assert torch.isfinite(out.float()).all()
    assert err < 4e-3, err


@pytest.mark.parametrize("hd", [64, 128])
def test_attention_tc_sharp_rows_and_repeated_launches(lib, hd):
    """Logits with a large spread: row maxima keep growing by more than the lazy-rescale threshold (2^8) along the
    keys, so the speculative exponentials are redone after rescaling O.  Launched three times back to back: the item
    scheduler's device counters re-arm themselves at the end of every launch."""
    B, H, Tq = 3, 8, 1100
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(11)
    ramp = torch.linspace(0.5, 5.0, Tq, device="cuda")[None, :, None]
    q = (torch.randn(B, Tq, E, device="cuda", generator=g) * 2.0).half()
    k = (torch.randn(B, Tq, E, device="cuda", generator=g) * ramp).half()   # later keys score higher and higher
    v = torch.randn(B, Tq, E, device="cuda", generator=g).half()
    ref = attention_reference(q, k, v, H, 0, 1)
    outs = []
    for _ in range(3):
        out = torch.full((B, Tq, E), float("nan"), device="cuda", dtype=torch.half)
        _check(lib, lib.frt2_op_attention(0, _p(q), _p(k), _p(v), _p(out), B, H, hd, Tq, Tq, 0, 1, _stream()))
        outs.append(out)
    torch.cuda.synchronize()
    err = (outs[0].float() - ref).abs().max().item()
    print(f"attention tc sharp hd={hd}: max-abs err {err:.3e}")
    assert torch.isfinite(outs[0].float()).all() and err < 8e-3, err
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("B,T,first,last,use_tail", [(2, 24, 1, 1, 0), (1, 8, 1, 0, 1), (2, 8, 0, 0, 1),
                                                      (1, 16, 0, 1, 1), (1, 8, 1, 1, 1)])
def test_overlap_add(lib, B, T, first, last, use_tail):
    from oracle import codec_oracle as O
    n_fft, hop = 960, 240
    rng = np.random.default_rng(5)
    frames = rng.standard_normal((B, T, n_fft)).astype(np.float32)
    tail = rng.standard_normal((B, 3, n_fft)).astype(np.float32)
    n = np.arange(n_fft)
    window = (0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)).astype(np.float32)
    fr = frames * window
    tl = tail * window
    pad = (n_fft - hop) // 2
    # oracle (reference decoder.py:384-405 / 431-467)
    cat = fr if first else np.concatenate([tl, fr], axis=1)
    y, env = O.overlap_add(cat, window, hop)
    with np.errstate(invalid="ignore", divide="ignore"):
        y = y / env
    y = y[:, pad:] if first else y[:, n_fft - hop:]
    y = y[:, :-pad] if last else y[:, :-(n_fft - hop)]
    d_fr = torch.from_numpy(fr).cuda()
    d_tl = torch.from_numpy(tl).cuda() if use_tail else None
    d_w = torch.from_numpy(window).cuda()
    out = torch.full((B, y.shape[1]), float("nan"), device="cuda")
    _check(lib, lib.frt2_op_overlap_add(_p(d_fr), _p(d_tl), _p(d_w), None, _p(out), out.stride(0), B, T, n_fft, hop,
                                        first, last, _stream()))
    err = np.abs(out.cpu().numpy() - y).max()
    assert err < 1e-5, err


SKINNY_CASES = [
    # batches, rows, Kc, taps, N, bias, act, resid, alpha
    (1, 8, 1024, 1, 1024, True, 0, True, 1.0),
    (1, 8, 1024, 1, 4096, True, 1, False, 1.0),
    (1, 8, 4096, 1, 1024, True, 0, True, 1.0),
    (1, 8, 1024, 7, 1024, True, 0, False, 1.0),
    (2, 8, 128, 3, 128, True, 0, True, 1.0),
    (1, 8, 1024, 1, 962, True, 2, False, 1.0),
    (1, 8, 1024, 1, 960, False, 0, False, 1.0 / 960),
    (1, 4, 512, 2, 1024, True, 1, False, 1.0),
    (1, 1, 512, 1, 1024, True, 0, False, 1.0),
    (3, 5, 64, 1, 70, True, 0, False, 1.0),
]


@pytest.mark.parametrize("case", SKINNY_CASES, ids=lambda c: "b%d_m%d_k%d_t%d_n%d_b%d_a%d_r%d" % c[:8])
def test_gemm_skinny(lib, case):
    batches, rows, Kc, taps, Nn, use_bias, act, use_resid, alpha = case
    g = torch.Generator(device="cuda").manual_seed(7)
    A = (torch.randn(batches, rows, Kc, device="cuda", generator=g)).half()
    W = (torch.randn(Nn, taps * Kc, device="cuda", generator=g) / math.sqrt(taps * Kc)).half()
    if act == 2:
        W = W * 0.5
    bias = torch.randn(Nn, device="cuda", generator=g) * 0.1 if use_bias else None
    resid = torch.randn(batches, rows, Nn, device="cuda", generator=g) if use_resid else None
    o32 = torch.full((batches, rows, Nn), float("nan"), device="cuda")
    o16 = torch.full((batches, rows, Nn), float("nan"), device="cuda", dtype=torch.half)
    _check(lib, lib.frt2_op_gemm(2, _p(A), _p(W), batches, rows, Kc, taps, Nn, alpha, _p(bias), act, _p(resid),
                                 _p(o32), _p(o16), _stream()))
    torch.cuda.synchronize()
    ref = gemm_reference(A, W, taps, bias, act, resid, alpha)
    scale = ref.abs().max().item() + 1e-6
    assert (o32 - ref).abs().max().item() / scale < 2e-3
    assert (o16.float() - ref).abs().max().item() / scale < 3e-3


STREAM_CASES = [
    # rows, K, N, bias, act (0 none, 1 GELU, 3 SwiGLU on interleaved gate/up rows), resid
    (1, 1536, 2048, True, 0, False),
    (1, 1536, 1536, False, 0, True),
    (8, 1536, 17920, False, 3, False),      # every warp of the grid takes part: 2240 column tiles on 148 x 16 warps
    (4, 8960, 1536, False, 0, True),        # 280 k-blocks per tile, 192 tiles on 148 CTAs
    (3, 64, 70, True, 0, False),            # fewer k-blocks than warps, N not a multiple of 8
    (8, 160, 64, True, 1, True),
    (2, 32, 8, False, 0, False),
    (5, 896, 9728, False, 3, False),
    (16, 1536, 2048, True, 0, True),        # 9..16 rows: both halves of the m16 tile
    (12, 1536, 17920, False, 3, False),
    (10, 8960, 1536, False, 0, True),       # the widest activation tile that fits (10 rows of 8960)
    (9, 64, 70, True, 1, False),
]


@pytest.mark.parametrize("case", STREAM_CASES, ids=lambda c: "m%d_k%d_n%d_b%d_a%d_r%d" % c)
def test_gemm_stream(lib, case):
    """K2w (gemm_stream.cu) against a torch fp32 reference of the same op on the same fp16-rounded operands."""
    rows, K, Nn, use_bias, act, use_resid = case
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(1, rows, K, device="cuda", generator=g).half()
    W = (torch.randn(Nn, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(Nn, device="cuda", generator=g) * 0.1 if use_bias else None
    resid = torch.randn(1, rows, Nn, device="cuda", generator=g) if use_resid else None
    n_out = Nn // 2 if act == 3 else Nn
    o32 = torch.full((1, rows, Nn), float("nan"), device="cuda") if act != 3 else None
    o16 = torch.full((1, rows, n_out), float("nan"), device="cuda", dtype=torch.half)
    _check(lib, lib.frt2_op_gemm(3, _p(A), _p(W), 1, rows, K, 1, Nn, 1.0, _p(bias), act, _p(resid), _p(o32), _p(o16), _stream()))
    torch.cuda.synchronize()
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act == 3:
        ref = torch.nn.functional.silu(y[..., 0::2]) * y[..., 1::2]
    else:
        ref = torch.nn.functional.gelu(y) if act == 1 else y
        if resid is not None:
            ref = ref + resid
    scale = ref.abs().max().item() + 1e-6
    if o32 is not None:
        assert (o32 - ref).abs().max().item() / scale < 2e-3
    assert torch.isfinite(o16.float()).all() and (o16.float() - ref).abs().max().item() / scale < 3e-3


def test_fp16_outputs_saturate_instead_of_overflowing(lib):
    """fp32 -> fp16 conversions saturate at +-65504 (F2FP.SATFINITE): no inf / NaN from an out-of-range activation."""
    M, K, Nn = 256, 64, 512
    A = torch.full((1, M, K), 300.0, device="cuda").half()
    W = torch.full((Nn, K), 300.0, device="cuda").half()
    W[1::2] = -300.0
    for impl in (0, 1):
        o16 = torch.zeros(1, M, Nn, device="cuda", dtype=torch.half)
        _check(lib, lib.frt2_op_gemm(impl, _p(A), _p(W), 1, M, K, 1, Nn, 1.0, None, 0, None, None, _p(o16), _stream()))
        torch.cuda.synchronize()
        assert torch.isfinite(o16.float()).all()
        assert (o16[0, :, 0::2] == 65504).all() and (o16[0, :, 1::2] == -65504).all()


def _resample_golden():
    import os
    from .helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "resample.npz"))
    names = sorted({k.split("::")[0] for k in g.files if "::" in k})
    return [(n, g[n + "::x"], g[n + "::y"], int(g[n + "::rates"][0]), int(g[n + "::rates"][1])) for n in names]


@pytest.mark.parametrize("case", _resample_golden(), ids=lambda c: c[0])
def test_resample_matches_torchaudio_golden(case):
    """K6 resampler through the C ABI against vectors produced by torchaudio.functional.resample (the call the
    reference makes at fireredtts2.py:65,389-391).  fp32 FIR, K sequential FMAs: tolerance 1e-6 of full scale."""
    from fireredtts2_b200.codec import resample
    name, x, y, orig, new = case
    out = resample(torch.from_numpy(x).cuda(), orig, new).cpu().numpy()
    assert out.shape == y.shape
    err = np.abs(out - y).max()
    print(f"resample {name}: max-abs {err:.2e} at peak {np.abs(y).max():.2e}")
    assert err <= 1e-6 * max(1.0, float(np.abs(y).max()))


def test_resample_ragged_long_and_errors():
    from fireredtts2_b200.codec import resample
    from oracle import codec_oracle as O
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((3, 50001)) * 0.2).astype(np.float32)       # many blocks per row
    lens = np.asarray([50001, 12345, 1], dtype=np.int32)
    out = resample(torch.from_numpy(x).cuda(), 24000, 16000, lengths=torch.from_numpy(lens)).cpu().numpy()
    for b, n in enumerate(lens):
        ref = O.resample(x[b:b + 1, :n], 24000, 16000)[0]
        assert np.abs(out[b, :ref.shape[0]] - ref).max() <= 1e-6
        assert np.all(out[b, ref.shape[0]:] == 0)
    lead = resample(torch.from_numpy(x).cuda().reshape(3, 1, 50001), 24000, 16000)   # leading dims kept
    assert lead.shape == (3, 1, 33334)
    assert resample(torch.zeros(2, 0, device="cuda"), 24000, 16000).shape == (2, 0)
    same = torch.from_numpy(x).cuda()
    assert resample(same, 16000, 16000) is same
    with pytest.raises(ValueError):
        resample(same, 0, 16000)
    with pytest.raises(TypeError):
        resample(same.int(), 24000, 16000)
    with pytest.raises(ValueError):
        resample(torch.zeros(1, 8), 24000, 16000)                         # CPU tensor: no fallback path
