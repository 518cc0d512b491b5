"""GPU parity of frt2_rvq_encode (ResidualVQ.encode_codes, reference rvq.py:128-143) through the C ABI: indices against
the golden vectors of the REAL reference and against the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import codec_oracle as O
from tests.gpu_common import build_codec
from tests.test_oracle_golden import RVQ_ENC_CASES, load_rvq_encode_case

pytestmark = pytest.mark.gpu


def _compare(name, codes, ref_codes, margin, tol):
    """Index parity: identical wherever the reference's own top-2 margin exceeds `tol` (a different fp32 summation
    order may pick the runner-up of a tie; once a token diverges its later quantizers see another residual)."""
    codes = codes.cpu().numpy()
    assert codes.shape == ref_codes.shape and codes.dtype == np.int64
    nq = codes.shape[0]
    diff = codes != ref_codes                                   # (nq, B, T)
    first = np.where(diff.any(axis=0), diff.argmax(axis=0), nq)  # first diverging quantizer per token
    bad = 0
    for b, t in zip(*np.nonzero(first < nq)):
        if margin[first[b, t], b, t] > tol:
            bad += 1
    n_tok = first.size
    print(f"[parity] {name}: {int((first == nq).sum())}/{n_tok} tokens identical over all {nq} codebooks, "
          f"{int((first < nq).sum())} diverge at a near-tie, {bad} diverge elsewhere (min margin {margin.min():.2e})")
    assert bad == 0
    return int((first < nq).sum())


@pytest.mark.parametrize("name,preset", RVQ_ENC_CASES)
def test_rvq_encode_vs_reference_golden(name, preset):
    cfg, sd, g = load_rvq_encode_case(name, preset)
    codec = build_codec(cfg, sd)
    z = torch.from_numpy(g["z"]).cuda()
    n_div = _compare(name + "/channel-major", codec.rvq_encode_codes(z), g["codes"], g["margin"], 1e-4)
    assert n_div == 0      # the fixtures hold no near-ties: every index must match
    zt = z.transpose(1, 2).contiguous().transpose(1, 2)         # time-major storage, same logical (B, D, T)
    assert zt.stride(1) == 1
    assert torch.equal(codec.rvq_encode_codes(zt), codec.rvq_encode_codes(z))
    c2 = codec.rvq_encode_codes(z, nq=2)
    assert np.array_equal(c2.cpu().numpy(), g["codes"][:2])


def test_rvq_encode_c0_against_oracle_and_batch_independence():
    from fireredtts2_b200.config import C0
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    cfg = C0
    sd = dict(synthetic_state_dict(cfg, 0))
    sd.update(synthetic_encode_tensors(cfg, 0))
    codec = build_codec(cfg, sd, check_indices=False)
    rng = np.random.default_rng(5)
    z = rng.standard_normal((4, cfg.embed_dim, 250)).astype(np.float32)      # 1000 tokens, ragged last tile
    ref, margin = O.rvq_encode_codes(sd, z)
    zc = torch.from_numpy(z).cuda()
    codes = codec.rvq_encode_codes(zc)
    _compare("c0 4x250 vs oracle", codes, ref, margin, 2e-4)
    # tokens are independent: any slice of the batch encodes to the same indices; repeated runs are deterministic
    assert torch.equal(codec.rvq_encode_codes(zc[1:3, :, 17:200]), codes[:, 1:3, 17:200])
    assert torch.equal(codec.rvq_encode_codes(zc), codes)
    # encode -> decode: the indices are valid input of the decode path
    audio = codec.decode(codes.permute(1, 0, 2))
    assert audio.shape == (4, 250 * cfg.samples_per_token) and bool(torch.isfinite(audio).all())


def test_rvq_encode_tensor_core_chain_vs_cuda_core_kernel():
    """The product path runs the chain as split-fp16 tcgen05 GEMMs (rvq_encode_tc.cu); the fp32 CUDA-core kernel
    (DBG_GEMM_REF) is its checker: same indices wherever the oracle's top-2 margin exceeds the fp32 rounding scale, on
    channel-major and time-major inputs, for prefixes of the quantizers and for more tokens than one slab."""
    from fireredtts2_b200 import _native as N
    from fireredtts2_b200.config import C0
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    cfg = C0
    sd = dict(synthetic_state_dict(cfg, 0))
    sd.update(synthetic_encode_tensors(cfg, 0))
    codec = build_codec(cfg, sd, check_indices=False)
    rng = np.random.default_rng(21)
    z = rng.standard_normal((3, cfg.embed_dim, 333)).astype(np.float32)
    ref, margin = O.rvq_encode_codes(sd, z)
    zc = torch.from_numpy(z).cuda()
    tc = codec.rvq_encode_codes(zc)
    _compare("c0 3x333 tensor-core chain vs oracle", tc, ref, margin, 2e-4)
    codec.set_debug(N.DBG_GEMM_REF)
    simt = codec.rvq_encode_codes(zc)
    codec.set_debug(0)
    _compare("c0 3x333 cuda-core kernel vs oracle", simt, ref, margin, 2e-4)
    same = float((tc == simt).float().mean())
    print(f"[parity] tensor-core chain == cuda-core kernel on {same * 100:.2f} % of the indices")
    assert same >= 0.995
    zt = zc.transpose(1, 2).contiguous().transpose(1, 2)                 # time-major storage
    assert torch.equal(codec.rvq_encode_codes(zt), tc)
    assert torch.equal(codec.rvq_encode_codes(zc, nq=3), tc[:3])
    # more than one slab of 32768 tokens: groups of items, same indices as item by item
    zb = torch.from_numpy(rng.standard_normal((5, cfg.embed_dim, 9000)).astype(np.float32)).cuda()
    big = codec.rvq_encode_codes(zb, nq=4)
    for b in (0, 3, 4):
        assert torch.equal(big[:, b], codec.rvq_encode_codes(zb[b:b + 1], nq=4)[:, 0])


def test_rvq_encode_tile_size_does_not_change_the_indices():
    """The kernel picks 32 or 8 tokens per CTA from the batch size (small batches: more, smaller tiles); a token's
    arithmetic does not depend on the tile it sits in, so the variants agree bit for bit."""
    from fireredtts2_b200.config import SMALL
    from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict
    cfg = SMALL
    sd = dict(synthetic_state_dict(cfg, 1))
    sd.update(synthetic_encode_tensors(cfg, 1))
    codec = build_codec(cfg, sd, check_indices=False)
    from fireredtts2_b200 import _native as N
    codec.set_debug(N.DBG_GEMM_REF)                                     # the CUDA-core kernel (the product path is the GEMM chain)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    T = 32 * sms + 13                                                   # >= 32 tokens per SM: 32 per CTA
    z = torch.from_numpy(np.random.default_rng(8).standard_normal((1, cfg.embed_dim, T)).astype(np.float32)).cuda()
    big = codec.rvq_encode_codes(z)
    mid = codec.rvq_encode_codes(z[:, :, :20 * sms])                    # 8 per CTA, many tiles
    small = codec.rvq_encode_codes(z[:, :, :999])                       # 8 per CTA
    assert torch.equal(mid, big[:, :, :20 * sms]) and torch.equal(small, big[:, :, :999])


def test_rvq_encode_identity_first_code_is_exact():
    """Identity projections: z = codebook_0[idx] is reproduced exactly by the first quantizer (distance 0)."""
    from fireredtts2_b200.config import TINY_IDENT
    from fireredtts2_b200.weights import synthetic_state_dict
    cfg = TINY_IDENT
    sd = synthetic_state_dict(cfg, 3)
    codec = build_codec(cfg, sd)
    idx = np.random.default_rng(0).integers(0, cfg.codebook_size, size=(2, 40))
    z = np.asarray(sd["rvq.quantizers.0.codebook"])[idx].transpose(0, 2, 1).copy()   # (B, cd, T), input_dim == rvq_dim
    codes = codec.rvq_encode_codes(torch.from_numpy(z).cuda(), nq=1)
    assert np.array_equal(codes[0].cpu().numpy(), idx)


def test_rvq_encode_needs_the_encode_side_tensors():
    from fireredtts2_b200.config import TINY
    from fireredtts2_b200.weights import synthetic_state_dict
    codec = build_codec(TINY, synthetic_state_dict(TINY, 0))                 # decode-only state_dict
    with pytest.raises(ValueError):
        codec.rvq_encode_codes(torch.zeros(1, TINY.embed_dim, 4).cuda())
    with pytest.raises(ValueError):
        codec.rvq_encode_codes(torch.zeros(1, 4).cuda())
