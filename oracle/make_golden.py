"""Generate golden vectors from the REAL reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz + tests/golden/MANIFEST.json

Imports the reference's own codec modules from /root/reference (read-only, unmodified), loads the
synthetic weights of ``fireredtts2_b200.weights.synthetic_state_dict`` into them with
``load_state_dict`` and records ``RedCodecInfer.decode`` / ``.decode_one_token`` outputs plus a few
intermediates.  /root/reference does not exist on the GPU box, so nothing in tests/ imports this
script; the tests only read the committed fixtures (and regenerate the same numpy-seeded weights).

Construction recipe follows SURVEY.md §8c: build ResidualVQ / UpConv / AcousticDecoder directly and a
RedCodecInfer subclass that skips the (640 M-param, not needed for decode) Whisper encoder.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
os.environ.setdefault("FRT2_REFERENCE", "/root/reference")   # golden vectors come from the read-only checkout itself

from fireredtts2_b200.config import PRESETS, CodecConfig  # noqa: E402
from fireredtts2_b200.weights import adversarial_state_dict, synthetic_state_dict, synthetic_tokens  # noqa: E402
from oracle.reference_runner import _import as _import_reference, build_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
RefDecodeOnly, _REF_ROOT = _import_reference()
assert _REF_ROOT == "/root/reference", _REF_ROOT
WEIGHTS = {"synthetic": synthetic_state_dict, "adversarial": adversarial_state_dict}


def record_intermediates(m, tokens):
    """Tap a few intermediates of the reference decode via forward hooks (time-major on save)."""
    taps = {}

    def hook(name, tm):
        def fn(_mod, _inp, out):
            o = out[0] if isinstance(out, tuple) else out
            taps[name] = (o.transpose(1, 2) if tm else o).detach().numpy().copy()
        return fn

    hs = [
        m.upsample.register_forward_hook(hook("x50", False)),
        m.acoustic_decoder.upsample_conv.register_forward_hook(hook("up_full", True)),
        m.acoustic_decoder.backbone.prior_net.register_forward_hook(hook("prior", True)),
        m.acoustic_decoder.backbone.transformers[0].register_forward_hook(hook("layer0", False)),
        m.acoustic_decoder.backbone.register_forward_hook(hook("final", False)),
    ]
    with torch.inference_mode():
        emb_cm = None
        audio = m.decode(tokens)
        z = m.rvq.decode_codes(tokens.permute(1, 0, 2)).transpose(1, 2)
    for h in hs:
        h.remove()
    del emb_cm
    taps["z"] = z.numpy().copy()
    taps["audio"] = audio.numpy().copy()
    return taps


def case_offline(name, preset, B, L, wseed, tseed, idx_dtype, keep, weights="synthetic"):
    cfg = PRESETS[preset]
    sd = WEIGHTS[weights](cfg, wseed)
    m = build_reference(cfg, sd)
    tok = synthetic_tokens(cfg, B, L, tseed, np.int64)
    t = torch.from_numpy(tok)
    if idx_dtype == "int32_permuted":
        # production form: torch.stack(samples).permute(1,2,0) int32 non-contiguous (fireredtts2.py:196)
        t = torch.from_numpy(np.ascontiguousarray(tok.transpose(2, 0, 1)).astype(np.int32)).permute(1, 2, 0)
    taps = record_intermediates(m, t)
    out = {k: taps[k] for k in keep}
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), tokens=tok, **out)
    return dict(name=name, kind="offline", preset=preset, B=B, L=L, wseed=wseed, tseed=tseed,
                idx=idx_dtype, keys=sorted(out), weights=weights)


def case_stream(name, preset, B, L, wseed, tseed, chunks, weights="synthetic", keep_kv=True, keep_offline=True):
    """decode_one_token over `chunks` (list of chunk lengths summing to L).  keep_kv=False drops bb_kv_cache from the
    fixture (12.6 MB after 16 tokens at C0); the four small caches are always kept."""
    cfg = PRESETS[preset]
    sd = WEIGHTS[weights](cfg, wseed)
    m = build_reference(cfg, sd)
    tok = synthetic_tokens(cfg, B, L, tseed, np.int64)
    t = torch.from_numpy(tok)
    cache = {}
    outs = []
    pos = 0
    with torch.inference_mode():
        for ci, lc in enumerate(chunks):
            a, cache = m.decode_one_token(t[:, :, pos:pos + lc], cache, ci == len(chunks) - 1)
            outs.append(a.numpy().copy())
            pos += lc
        offline = m.decode(t).numpy()
    save = {f"audio_{i}": o for i, o in enumerate(outs)}
    save.update({"cache_" + k: v.numpy().copy() for k, v in cache.items() if keep_kv or k != "bb_kv_cache"})
    if keep_offline:
        save["offline"] = offline
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), tokens=tok, chunks=np.asarray(chunks), **save)
    cat = np.concatenate(outs, axis=1)
    return dict(name=name, kind="stream", preset=preset, B=B, L=L, wseed=wseed, tseed=tseed, weights=weights,
                chunks=list(chunks), stream_vs_offline_maxabs=float(np.abs(cat - offline).max()))


def case_rvq_emb(name, preset, B, L, wseed, tseed):
    """ResidualVQ.decode_codes on an Identity-out_project config (C1): the index-ordered fp32 sum `emb` is what reaches
    rvq.output_proj (captured with a forward pre-hook) and must be reproduced bit for bit (rvq.py:145-164)."""
    cfg = PRESETS[preset]
    assert not cfg.has_out_project
    sd = synthetic_state_dict(cfg, wseed)
    m = build_reference(cfg, sd)
    tok = synthetic_tokens(cfg, B, L, tseed, np.int64)
    got = {}
    h = m.rvq.output_proj.register_forward_pre_hook(lambda _m, inp: got.__setitem__("emb", inp[0].detach().numpy().copy()))
    with torch.inference_mode():
        z = m.rvq.decode_codes(torch.from_numpy(tok).permute(1, 0, 2))
    h.remove()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), tokens=tok, emb=got["emb"].transpose(0, 2, 1).copy(),
                        z=z.numpy().transpose(0, 2, 1).copy())
    return dict(name=name, kind="rvq_emb", preset=preset, B=B, L=L, wseed=wseed, tseed=tseed)


def case_reference_init(name, preset, L):
    """Weights from the reference constructors' OWN init (torch.manual_seed(0)) + N(0,1) codebooks: checks the
    state_dict importer against a genuine reference state_dict (weight-norm parametrisation keys etc.).
    Weights are too large to commit for C0, so this case uses a small preset and stores the state_dict."""
    cfg = PRESETS[preset]
    torch.manual_seed(0)
    m = RefDecodeOnly(cfg).eval()
    g = torch.Generator().manual_seed(7)
    for q in m.rvq.quantizers:
        q.codebook.copy_(torch.randn(q.codebook.shape, generator=g))
    tok = synthetic_tokens(cfg, 1, L, 99, np.int64)
    with torch.inference_mode():
        audio = m.decode(torch.from_numpy(tok)).numpy()
    sd = {("sd::" + k): v.numpy() for k, v in m.state_dict().items()
          if v.dtype.is_floating_point and not any(e in k for e in ("input_proj", "in_project", "cluster_size", "embed_avg"))}
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), tokens=tok, audio=audio, **sd)
    return dict(name=name, kind="reference_init", preset=preset, B=1, L=L)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    man = []
    inter = ["z", "x50", "up_full", "prior", "layer0", "final", "audio"]
    man.append(case_offline("tiny_offline", "TINY", 2, 9, 0, 1234, "int64", inter))
    man.append(case_offline("tiny_ident_offline", "TINY_IDENT", 1, 7, 1, 5, "int64", ["z", "audio"]))
    man.append(case_offline("small_offline_i32perm", "SMALL", 3, 21, 2, 77, "int32_permuted", ["z", "final", "audio"]))
    man.append(case_stream("tiny_stream_1", "TINY", 2, 6, 0, 4321, [1, 1, 1, 1, 1, 1]))
    man.append(case_stream("tiny_stream_multi", "TINY", 1, 7, 0, 99, [2, 1, 3, 1]))
    man.append(case_stream("tiny_stream_single_last", "TINY", 1, 1, 0, 3, [1]))
    man.append(case_reference_init("micro_refinit", "MICRO", 5))
    # C0 (the benchmark architecture), config-1 shape shortened to 25 tokens (2 s) to keep the fixture small
    man.append(case_offline("c0_offline_L25", "C0", 1, 25, 0, 1234, "int64", ["audio"]))
    # ---- the BASELINE.json configs themselves ----
    # configs[0]: one 10 s monologue, C0
    man.append(case_offline("c0_offline_L125", "C0", 1, 125, 0, 1234, "int64", ["audio"]))
    # configs[2] item shape: one 30 s utterance, C0 (the benchmark decodes 64 of these per step)
    man.append(case_offline("c0_offline_L375", "C0", 1, 375, 0, 4242, "int64", ["audio"]))
    # configs[1]: streaming, batch 1, one token per call, C0 — 16 tokens, and 72 tokens (> 512 frames of K/V state: the
    # step's attention is then split over several CTAs per head)
    man.append(case_stream("c0_stream_16", "C0", 1, 16, 0, 777, [1] * 16, keep_kv=False, keep_offline=False))
    man.append(case_stream("c0_stream_72", "C0", 1, 72, 0, 778, [1] * 72, keep_kv=False, keep_offline=False))
    # C1 = C0 with Identity out_project: the bit-exact index-ordered sum at C0 dimensions
    man.append(case_rvq_emb("c1_rvq_emb", "C1", 2, 20, 0, 31))
    # adversarial weights (weights.adversarial_state_dict): LayerNorm gamma in [0.1, 5], row mean ~ 50 x row spread,
    # outlier channels x 100, GELU activations in the thousands — 4 layers at the C0 widths
    man.append(case_offline("adv4_offline", "ADV4", 2, 40, 0, 99, "int64", ["audio"], weights="adversarial"))
    man.append(case_stream("adv4_stream_12", "ADV4", 1, 12, 0, 98, [1] * 12, weights="adversarial", keep_kv=False,
                           keep_offline=False))
    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "torch": torch.__version__, "cases": man}, f, indent=1)
    for c in man:
        print(c)


if __name__ == "__main__":
    main()
