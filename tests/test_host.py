"""Host-side logic that needs no GPU: config interop with the reference JSON, weight naming, status mapping."""
import numpy as np
import pytest

from fireredtts2_b200.config import C0, C1, PRESETS, TINY, CodecConfig
from fireredtts2_b200.weights import decode_keys, synthetic_state_dict, weight_norm_materialise


def test_c0_matches_survey():
    assert (C0.embed_dim, C0.num_layers, C0.num_heads, C0.head_dim) == (1024, 12, 16, 64)
    assert (C0.n_fft, C0.n_bins, C0.istft_pad, C0.samples_per_token, C0.frames_per_token) == (960, 481, 360, 1920, 8)
    assert C0.has_out_project and C0.has_output_proj
    assert not C1.has_out_project and C1.has_output_proj


def test_reference_json_roundtrip():
    d = {"codec": C0.to_reference_dict()}
    assert CodecConfig.from_reference_dict(d) == C0
    bad = C0.to_reference_dict()
    bad["acoustic_decoder"]["causal"] = False
    with pytest.raises(AssertionError):      # reference decoder.py:675-677
        CodecConfig.from_reference_dict(bad)
    with pytest.raises(ValueError):
        CodecConfig(output_dim=512)


@pytest.mark.parametrize("name", ["TINY", "TINY_IDENT", "SMALL", "MICRO"])
def test_synthetic_weights_cover_decode_keys(name):
    cfg = PRESETS[name]
    sd = synthetic_state_dict(cfg, 0)
    assert sorted(sd) == sorted(decode_keys(cfg))
    sd2 = synthetic_state_dict(cfg, 0)
    assert all(np.array_equal(sd[k], sd2[k]) for k in sd)          # reproducible on any box
    assert all(v.dtype == np.float32 for v in sd.values())


def test_weight_norm_materialise():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((6, 5, 1)).astype(np.float32)
    g = rng.uniform(0.5, 2, (6, 1, 1)).astype(np.float32)
    W = weight_norm_materialise(g, v)
    assert np.allclose(np.sqrt((W ** 2).sum(axis=(1, 2))), g[:, 0, 0], rtol=1e-5)


def test_status_to_exception_mapping():
    from fireredtts2_b200 import _native as N
    N.load()
    with pytest.raises(IndexError):
        N.check(N.ERR_INDEX_OOR)
    with pytest.raises(TypeError):
        N.check(N.ERR_BAD_DTYPE)
    with pytest.raises(ValueError):
        N.check(N.ERR_BAD_ARG)
    with pytest.raises(OverflowError):
        N.check(N.ERR_STATE_OVERFLOW)
    with pytest.raises(N.Frt2Error):
        N.check(N.ERR_CUDA)
    N.check(N.FRT2_OK)
