"""Golden vectors for the RVQ ENCODE side, from the REAL reference (run in the build container only).

    python oracle/make_golden_rvq_encode.py     # writes tests/golden/rvq_encode_*.npz

Builds the reference's own ``ResidualVQ`` (/root/reference/fireredtts2/codec/rvq.py, unmodified), loads the numpy-seeded
synthetic weights (decode side: ``synthetic_state_dict``; encode side: ``synthetic_encode_tensors``) with
``load_state_dict`` and records ``encode_codes(z)`` for seeded inputs, plus the reference's own top-2 margin of every
arg-max decision (recomputed with the reference modules) so that the tests can tell a legitimate fp32 tie from a bug.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("FRT2_REFERENCE", "/root/reference"))

from fireredtts2.codec.rvq import ResidualVQ  # noqa: E402  (reference)

from fireredtts2_b200.config import PRESETS  # noqa: E402
from fireredtts2_b200.weights import synthetic_encode_tensors, synthetic_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def build(cfg, wseed, input_dim):
    d = cfg.to_reference_dict()["rvq"]
    d["input_dim"] = input_dim
    m = ResidualVQ(**d).eval()
    sd = {k[len("rvq."):]: torch.from_numpy(np.asarray(v)) for k, v in synthetic_state_dict(cfg, wseed).items()
          if k.startswith("rvq.")}
    sd.update({k[len("rvq."):]: torch.from_numpy(v) for k, v in synthetic_encode_tensors(cfg, wseed, input_dim).items()})
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(any(e in k for e in ("inited", "cluster_size", "embed_avg")) for k in missing), missing
    return m


def margins(m, z):
    """The reference's own decisions, step by step with its own modules (rvq.py:62-89,128-143)."""
    out = []
    with torch.inference_mode():
        residual = m.input_proj(z).clone().float()
        for q in m.quantizers:
            z_e = q.in_project(residual.float()).float()
            enc = z_e.transpose(1, 2).reshape(-1, z_e.shape[1])
            dist = (enc.pow(2).sum(1, keepdim=True) - 2 * enc @ q.codebook.float().t()
                    + q.codebook.float().pow(2).sum(1, keepdim=True).t())
            top = (-dist).topk(2, dim=1).values
            out.append((top[:, 0] - top[:, 1]).reshape(z.shape[0], z.shape[2]).numpy())
            z_q, _ = q.encode_code(residual)
            residual = residual - z_q
    return np.stack(out)


def case(name, preset, B, T, wseed, zseed, input_dim=None, scale=1.0):
    cfg = PRESETS[preset]
    input_dim = cfg.embed_dim if input_dim is None else input_dim
    m = build(cfg, wseed, input_dim)
    z = (np.random.default_rng(zseed).standard_normal((B, input_dim, T)) * scale).astype(np.float32)
    with torch.inference_mode():
        codes = m.encode_codes(torch.from_numpy(z)).numpy()
    mg = margins(m, torch.from_numpy(z))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), z=z, codes=codes, margin=mg,
                        meta=np.asarray([wseed, zseed, input_dim]))
    print(name, preset, "codes", codes.shape, "min margin", float(mg.min()), "decisions", mg.size)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    case("rvq_encode_tiny", "TINY", 2, 9, 0, 11)
    case("rvq_encode_tiny_ident", "TINY_IDENT", 1, 33, 1, 12)                    # Identity in/out_project
    case("rvq_encode_tiny_noinput", "TINY", 2, 17, 0, 13, input_dim=64)          # Identity input_proj
    case("rvq_encode_small", "SMALL", 3, 21, 2, 14)
    case("rvq_encode_c0", "C0", 1, 50, 0, 15)                                    # the benchmark architecture, 4 s


if __name__ == "__main__":
    main()
