"""One call that puts the library behind a live reference ``FireRedTTS2`` object (reference ``fireredtts2/fireredtts2.py:14-59``).

    tts = FireRedTTS2(pretrained_dir, gen_type, device)       # the reference, unmodified
    fireredtts2_b200.dropin.install(tts)                      # codec decode / streaming step + frame tail on libfrt2_b200

What is swapped — and nothing else:
  * ``tts._audio_tokenizer`` (``RedCodecInfer``, fireredtts2.py:51-53) -> ``RedCodecB200.from_reference`` of it: ``decode``
    (fireredtts2.py:195-199, 441), ``decode_one_token`` and — with ``native_encode=True`` — ``encode`` (fireredtts2.py:96)
    run on the library; without the flag ``encode`` stays delegated to the wrapped reference module;
  * ``tts._model.generate_frame`` (llm.py:274-330) -> ``GenerateFrameB200``: the backbone stays the reference's, everything
    behind ``last_h`` is one ``frt2_fd_generate`` call.
The text tokenizer, prompt preparation, the LLM backbone and the dialogue loop are the reference's own code.  There is no
CPU fallback: on a machine without the library / a CUDA device the constructors raise and the object is left untouched.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Optional


@dataclasses.dataclass
class Installed:
    """What ``install`` replaced; ``uninstall()`` puts the reference's objects back."""
    tts: Any
    reference_codec: Any
    codec: Any
    generate_frame: Optional[Any]

    def uninstall(self) -> None:
        if self.tts._audio_tokenizer is self.codec:
            self.tts._audio_tokenizer = self.reference_codec
        if self.generate_frame is not None:
            self.generate_frame.uninstall()


def install(tts, native_encode: bool = False, frame_tail: bool = True, max_batch: int = 8, seed: int = 0,
            num_heads: Optional[int] = None) -> Installed:
    """Both objects are built BEFORE anything is assigned, so a failure (no GPU, unsupported widths) leaves ``tts`` as it was."""
    from .codec import RedCodecB200
    from .frame_decoder import FrameDecoderB200, GenerateFrameB200
    device = getattr(tts, "device", "cuda:0")
    ref_codec = tts._audio_tokenizer
    codec = RedCodecB200.from_reference(ref_codec, num_heads=num_heads, device=str(device), native_encode=native_encode)
    tail = FrameDecoderB200.from_reference(tts._model, device=str(device), max_batch=max_batch) if frame_tail else None
    tts._audio_tokenizer = codec
    gen = GenerateFrameB200.install(tts._model, tail=tail, seed=seed) if frame_tail else None
    return Installed(tts, ref_codec, codec, gen)
