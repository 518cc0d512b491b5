"""Drive the UNMODIFIED reference codec decode (test / baseline infrastructure, not product code).

Where the reference lives:
  * ``/root/reference``                 the read-only checkout in the build container (golden generation);
  * ``<repo>/baseline/_ref``            an offline ``pip install --target`` of that checkout made by
                                        ``__graft_entry__.build()`` — git-ignored, but it travels to the GPU box with the
                                        gpurun snapshot, so ``bench.py --impl reference`` / ``cpu_baseline`` time the real
                                        ``RedCodecInfer.decode`` on the box's host cores.
Construction recipe: build ``ResidualVQ`` / ``UpConv`` / ``AcousticDecoder`` directly and a
``RedCodecInfer`` subclass that skips the 640 M-parameter Whisper encoder the decode path never touches; every decode /
decode_one_token call below is the reference's own method (codec/model.py:307-376).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.environ.get("FRT2_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def reference_root():
    """First location that holds ``fireredtts2/codec/model.py``, or None."""
    for c in CANDIDATES:
        if c and os.path.exists(os.path.join(c, "fireredtts2", "codec", "model.py")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def _import():
    root = reference_root()
    if root is None:
        raise ImportError("the reference codec is not available (neither baseline/_ref nor /root/reference)")
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch.nn as nn
    from fireredtts2.codec.decoder import AcousticDecoder
    from fireredtts2.codec.model import RedCodecInfer, UpConv
    from fireredtts2.codec.rvq import ResidualVQ

    class RefDecodeOnly(RedCodecInfer):
        """RedCodecInfer with only the three decode-side sub-modules."""

        def __init__(self, cfg):
            nn.Module.__init__(self)
            d = cfg.to_reference_dict()
            self.rvq = ResidualVQ(**d["rvq"])
            self.upsample = UpConv(**d["upsample"])
            self.acoustic_decoder = AcousticDecoder(**d["acoustic_decoder"])

    return RefDecodeOnly, root


def build_reference(cfg, sd_np=None):
    """The reference module for ``cfg`` in eval mode; ``sd_np`` (numpy state_dict in the reference's key naming) is loaded
    with ``load_state_dict`` — only encode-side keys may be missing."""
    import torch
    RefDecodeOnly, _ = _import()
    torch.manual_seed(0)
    m = RefDecodeOnly(cfg).eval()
    if sd_np is not None:
        sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        enc_only = ("input_proj", "in_project", "inited", "cluster_size", "embed_avg")
        bad = [k for k in missing if not any(e in k for e in enc_only)]
        assert not bad, bad
    return m
