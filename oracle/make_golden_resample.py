"""Golden vectors for the waveform resampler — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference resamples with the third-party ``torchaudio.functional.resample`` (fireredtts2.py:65 prompt audio ->
16 kHz, fireredtts2.py:389-391 every generated turn 24 kHz -> 16 kHz for the context loop).  torchaudio (2.11.0 in
this image) is not vendored in /root/reference, so the oracle restates its published algorithm
(``_get_sinc_resample_kernel`` / ``_apply_sinc_resample_kernel``, sinc_interp_hann, lowpass_filter_width 6, rolloff
0.99) and is pinned to the outputs of torchaudio itself generated here:

    python oracle/make_golden_resample.py     # writes tests/golden/resample.npz
"""
import os
import sys

import numpy as np
import torch
import torchaudio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CASES = [  # name, orig, new, batch, length
    ("turn_24k_16k", 24000, 16000, 2, 4807),
    ("chunk_24k_16k", 24000, 16000, 1, 1560),
    ("tiny_24k_16k", 24000, 16000, 3, 5),
    ("prompt_48k_16k", 48000, 16000, 1, 3001),
    ("prompt_44k1_16k", 44100, 16000, 1, 2205),
    ("prompt_22k05_16k", 22050, 16000, 2, 1103),
    ("up_16k_24k", 16000, 24000, 1, 1001),
]


def main():
    rng = np.random.default_rng(20260101)
    out = {}
    for name, orig, new, B, n in CASES:
        x = (rng.standard_normal((B, n)) * 0.1).astype(np.float32)
        y = torchaudio.functional.resample(torch.from_numpy(x), orig, new).numpy()
        out[name + "::x"] = x
        out[name + "::y"] = y
        out[name + "::rates"] = np.asarray([orig, new], dtype=np.int64)
        print(name, x.shape, "->", y.shape)
    np.savez_compressed(os.path.join(GOLDEN, "resample.npz"), torchaudio=np.asarray(torchaudio.__version__), **out)


if __name__ == "__main__":
    sys.exit(main())
