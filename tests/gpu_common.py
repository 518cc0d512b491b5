"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import codec_oracle as O


def build_codec(cfg, sd, **kw):
    from fireredtts2_b200.codec import RedCodecB200
    return RedCodecB200(cfg, sd, device="cuda:0", **kw)


def report(name, ref, out):
    ref = np.asarray(ref, dtype=np.float64)
    out = np.asarray(out, dtype=np.float64)
    maxabs = float(np.abs(ref - out).max())
    snr = O.snr_db(ref, out)
    print(f"[parity] {name}: max-abs {maxabs:.3e}  ref-peak {np.abs(ref).max():.3e}  SNR {snr:.1f} dB")
    return maxabs, snr


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()
