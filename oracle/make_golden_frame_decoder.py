"""Golden vectors for the frame tail of the speech LM, from the reference's own ``Model.generate_frame`` (build container only).

    python oracle/make_golden_frame_decoder.py     # writes tests/golden/fd_*.npz

``fireredtts2/llm/llm.py`` is imported UNMODIFIED from /root/reference and ``Model.generate_frame`` (llm.py:274-330) runs as
it is.  Its transformer blocks come from ``torchtune`` (requirements.txt:1), which is not in this image and cannot be
installed (no network); a ``torchtune`` shim registered in ``sys.modules`` below provides ``qwen2(...)`` /
``TransformerDecoder`` with exactly the surface the reference touches (``tok_embeddings``, ``output``, ``max_seq_len``,
``setup_caches``, ``caches_are_enabled``, ``reset_caches``, ``forward(h, input_pos=, mask=)``) on top of Hugging Face
``transformers``' ``Qwen2Model`` — an independent implementation of the same published Qwen2 block.  The decoder's weights
are the numpy-seeded synthetic ones (``synthetic_frame_decoder_state_dict``, torchtune parameter names) mapped onto the HF
names, so the tests can rebuild them; the backbone is a one-layer stand-in whose only role is to produce ``last_h``, which
is recorded.  ``sample_topk`` / ``_multinomial_sample_one_no_sync`` are the reference's own; the Exp(1) draws are recorded
by wrapping ``torch.Tensor.exponential_`` for the duration of the call.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("FRT2_REFERENCE", "/root/reference"))

from transformers import DynamicCache, Qwen2Config, Qwen2Model  # noqa: E402

from fireredtts2_b200.frame_decoder import FD_PRESETS, synthetic_frame_decoder_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class TransformerDecoder(torch.nn.Module):
    """The part of torchtune.modules.transformer.TransformerDecoder the reference uses, on Hugging Face's Qwen2Model."""

    def __init__(self, vocab_size, num_layers, num_heads, num_kv_heads, embed_dim, intermediate_dim, max_seq_len,
                 attn_dropout=0.0, norm_eps=1e-6, rope_base=1000000.0, tie_word_embeddings=True):
        super().__init__()
        self.hf_config = Qwen2Config(vocab_size=8, hidden_size=embed_dim, intermediate_size=intermediate_dim,
                                     num_hidden_layers=num_layers, num_attention_heads=num_heads,
                                     num_key_value_heads=num_kv_heads, max_position_embeddings=max_seq_len,
                                     rms_norm_eps=norm_eps, rope_theta=rope_base, tie_word_embeddings=True,
                                     attention_dropout=attn_dropout, attn_implementation="eager")
        self.hf = Qwen2Model(self.hf_config)
        self.tok_embeddings = torch.nn.Embedding(8, embed_dim)     # replaced by Identity in _prepare_transformer
        self.output = torch.nn.Identity()
        self.max_seq_len = max_seq_len
        self._cache = None

    def setup_caches(self, batch_size, dtype, decoder_max_seq_len=None):
        self._cache = DynamicCache(config=self.hf_config)

    def caches_are_enabled(self):
        return self._cache is not None

    def reset_caches(self):
        self._cache = DynamicCache(config=self.hf_config)

    def forward(self, h, input_pos=None, mask=None):
        h = self.tok_embeddings(h)
        past = self._cache.get_seq_length()
        kv_len = past + h.shape[1]
        add = torch.zeros(mask.shape[0], 1, mask.shape[1], kv_len, dtype=h.dtype)
        add.masked_fill_(~mask[:, None, :, :kv_len], torch.finfo(h.dtype).min)
        out = self.hf(inputs_embeds=h, position_ids=input_pos, attention_mask=add, past_key_values=self._cache,
                      use_cache=True).last_hidden_state
        return self.output(out)


def _install_shim():
    tt = types.ModuleType("torchtune")
    models = types.ModuleType("torchtune.models")
    q2 = types.ModuleType("torchtune.models.qwen2")
    modules = types.ModuleType("torchtune.modules")
    tr = types.ModuleType("torchtune.modules.transformer")
    q2.qwen2 = lambda **kw: TransformerDecoder(**kw)
    tr.TransformerDecoder = TransformerDecoder
    sys.modules.update({"torchtune": tt, "torchtune.models": models, "torchtune.models.qwen2": q2,
                        "torchtune.modules": modules, "torchtune.modules.transformer": tr})


_install_shim()
import fireredtts2.llm.llm as ref_llm  # noqa: E402  (reference, unmodified)
from fireredtts2.llm import modules as ref_modules  # noqa: E402


def load_decoder_weights(model, sd):
    """synthetic (torchtune-named) tensors -> the reference Model: top-level ones by their own names, the decoder's onto
    the Hugging Face names of the stand-in."""
    t = lambda k: torch.from_numpy(np.asarray(sd[k], dtype=np.float32))
    with torch.no_grad():
        model.projection.weight.copy_(t("projection.weight"))
        model.audio_embeddings.weight.copy_(t("audio_embeddings.weight"))
        model.codebook0_head.weight.copy_(t("codebook0_head.weight"))
        model.audio_head.copy_(t("audio_head"))
        hf = model.decoder.hf
        hf.norm.weight.copy_(t("decoder.norm.scale"))
        for i, layer in enumerate(hf.layers):
            p = f"decoder.layers.{i}."
            layer.input_layernorm.weight.copy_(t(p + "sa_norm.scale"))
            layer.post_attention_layernorm.weight.copy_(t(p + "mlp_norm.scale"))
            for a, b in (("q_proj", "q_proj"), ("k_proj", "k_proj"), ("v_proj", "v_proj")):
                getattr(layer.self_attn, a).weight.copy_(t(p + f"attn.{b}.weight"))
                getattr(layer.self_attn, a).bias.copy_(t(p + f"attn.{b}.bias"))
            layer.self_attn.o_proj.weight.copy_(t(p + "attn.output_proj.weight"))
            layer.mlp.gate_proj.weight.copy_(t(p + "mlp.w1.weight"))
            layer.mlp.down_proj.weight.copy_(t(p + "mlp.w2.weight"))
            layer.mlp.up_proj.weight.copy_(t(p + "mlp.w3.weight"))


def build_model(cfg, wseed):
    ref_modules.FLAVORS["fd-backbone"] = lambda: TransformerDecoder(
        vocab_size=8, num_layers=1, num_heads=2, num_kv_heads=1, embed_dim=cfg.backbone_dim,
        intermediate_dim=2 * cfg.backbone_dim, max_seq_len=64)
    ref_modules.FLAVORS["fd-decoder"] = lambda: TransformerDecoder(
        vocab_size=8, num_layers=cfg.num_layers, num_heads=cfg.num_heads, num_kv_heads=cfg.num_kv_heads,
        embed_dim=cfg.dim, intermediate_dim=cfg.intermediate_dim, max_seq_len=64, norm_eps=cfg.norm_eps,
        rope_base=cfg.rope_base)
    torch.manual_seed(wseed)
    args = ref_llm.ModelArgs(backbone_flavor="fd-backbone", decoder_flavor="fd-decoder", text_vocab_size=32,
                             audio_vocab_size=cfg.audio_vocab_size, audio_num_codebooks=cfg.audio_num_codebooks,
                             decoder_loss_weight=0.5, use_text_loss=True)
    model = ref_llm.Model(args).eval()
    with torch.no_grad():
        model.text_embeddings.weight.normal_(0, 1.0)
    load_decoder_weights(model, synthetic_frame_decoder_state_dict(cfg, wseed))
    return model


# name, preset, batch, weight seed, data seed, topk, temperature
CASES = [
    ("fd_tiny", "FD_TINY", 2, 3, 21, 8, 0.9),
    ("fd_small", "FD_SMALL", 3, 5, 22, 30, 0.9),
    ("fd_small_b1", "FD_SMALL", 1, 5, 23, 20, 0.8),
    ("fd_200m", "FD_200M", 1, 0, 24, 30, 0.9),
]


def run_case(cfg, B, wseed, dseed, topk, temperature):
    model = build_model(cfg, wseed)
    model.setup_caches(B)
    n, V = cfg.audio_num_codebooks, cfg.audio_vocab_size
    g = torch.Generator().manual_seed(dseed)
    S = 5
    tokens = torch.zeros(B, S, n + 1, dtype=torch.long)
    tokens[:, :, :n] = torch.randint(0, V, (B, S, n), generator=g)
    tokens[:, :, n] = torch.randint(0, 32, (B, S), generator=g)
    tokens_mask = torch.ones(B, S, n + 1, dtype=torch.bool)
    input_pos = torch.arange(S)[None].repeat(B, 1)

    rec = {"last_h": None, "logits": [], "noise": []}
    hook = model.backbone.register_forward_hook(lambda m, a, out: rec.__setitem__("last_h", out[:, -1, :].detach().clone()))
    real_sample, real_exp = ref_llm.sample_topk, torch.Tensor.exponential_

    def sample_spy(logits, k, temp):
        rec["logits"].append(logits.detach().clone())
        return real_sample(logits, k, temp)

    def exp_spy(self, *a, **kw):
        out = real_exp(self, *a, **kw)
        rec["noise"].append(out.detach().clone())
        return out

    ref_llm.sample_topk, torch.Tensor.exponential_ = sample_spy, exp_spy
    try:
        torch.manual_seed(dseed)
        with torch.inference_mode():
            codes = model.generate_frame(tokens, tokens_mask, input_pos, temperature, topk)     # llm.py:274-330
    finally:
        ref_llm.sample_topk, torch.Tensor.exponential_ = real_sample, real_exp
        hook.remove()
    assert len(rec["logits"]) == n and len(rec["noise"]) == n
    return {"last_h": rec["last_h"].numpy(), "codes": codes.numpy().astype(np.int32),
            "logits": torch.stack(rec["logits"], 1).numpy(), "noise": torch.stack(rec["noise"], 1).numpy(),
            "meta": np.array([B, wseed, dseed, topk], dtype=np.int64), "temperature": np.float32(temperature)}


# V, topk, temperature, rows — sample_topk of the reference on its own (llm.py:39-49), exact ties at the k-th value included
SAMPLER_CASES = [(64, 8, 0.9, 64), (2048, 10, 0.75, 32), (2048, 30, 0.9, 32), (2051, 50, 1.0, 24), (512, 512, 1.3, 16),
                 (2048, 1, 0.5, 16)]


def sampler_golden():
    """The reference's ``sample_topk`` / ``_multinomial_sample_one_no_sync`` on seeded logits; the Exp(1) draws it makes
    are recorded by wrapping ``Tensor.exponential_``."""
    out = {}
    real_exp = torch.Tensor.exponential_
    for i, (V, topk, temp, rows) in enumerate(SAMPLER_CASES):
        g = torch.Generator().manual_seed(100 + i)
        logits = 2.0 * torch.randn(rows, V, generator=g)
        logits[: rows // 2] = torch.round(logits[: rows // 2] * 2) / 2        # many exact ties
        drawn = []

        def exp_spy(self, *a, **kw):
            r = real_exp(self, *a, **kw)
            drawn.append(r.detach().clone())
            return r

        torch.Tensor.exponential_ = exp_spy
        try:
            torch.manual_seed(200 + i)
            codes = ref_llm.sample_topk(logits, topk, temp)
        finally:
            torch.Tensor.exponential_ = real_exp
        assert len(drawn) == 1
        out[f"c{i}_logits"] = logits.numpy()
        out[f"c{i}_q"] = drawn[0].numpy()
        out[f"c{i}_codes"] = codes.numpy().astype(np.int32).reshape(-1)
        out[f"c{i}_meta"] = np.array([V, topk, rows], dtype=np.int64)
        out[f"c{i}_temperature"] = np.float32(temp)
    out["n_cases"] = np.int64(len(SAMPLER_CASES))
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import frame_decoder_oracle as FO_
    sg = sampler_golden()
    np.savez_compressed(os.path.join(GOLDEN, "fd_sampler.npz"), **sg)
    for i in range(int(sg["n_cases"])):
        V, topk, rows = (int(v) for v in sg[f"c{i}_meta"])
        got = FO_.sample_topk(sg[f"c{i}_logits"], topk, float(sg[f"c{i}_temperature"]), sg[f"c{i}_q"])
        print("sampler case", i, (V, topk, rows), "oracle == reference:", bool((got == sg[f"c{i}_codes"]).all()))
    from oracle import frame_decoder_oracle as FO
    for name, preset, B, wseed, dseed, topk, temperature in CASES:
        cfg = FD_PRESETS[preset]
        r = run_case(cfg, B, wseed, dseed, topk, temperature)
        out = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(out, **r)
        sd = synthetic_frame_decoder_state_dict(cfg, wseed)
        codes, logits = FO.generate_codes(sd, cfg, r["last_h"], topk, temperature, r["noise"])
        err = float(np.abs(logits - r["logits"]).max())
        print(name, "codes", r["codes"].tolist()[0][:8], "logit std", float(r["logits"].std()), "oracle max-abs", err,
              "codes equal", bool((codes == r["codes"]).all()), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
