"""One call that puts the library behind a live reference ``FireRedTTS2`` object (reference ``fireredtts2/fireredtts2.py:14-59``).

    tts = FireRedTTS2(pretrained_dir, gen_type, device)       # the reference, unmodified
    fireredtts2_b200.dropin.install(tts)                      # codec decode / streaming step + frame tail on libfrt2_b200

What is swapped — and nothing else:
  * ``tts._audio_tokenizer`` (``RedCodecInfer``, fireredtts2.py:51-53) -> ``RedCodecB200.from_reference`` of it: ``decode``
    (fireredtts2.py:195-199, 441), ``decode_one_token`` and — with ``native_encode=True`` — ``encode`` (fireredtts2.py:96)
    run on the library; without the flag ``encode`` stays delegated to the wrapped reference module;
  * ``tts._model.generate_frame`` (llm.py:274-330) -> ``GenerateFrameB200``: the backbone stays the reference's, everything
    behind ``last_h`` is one ``frt2_fd_generate`` call.
  * the name ``torchaudio`` inside the reference's module (``fireredtts2.py:5``) -> a pass-through proxy whose
    ``functional.resample`` sends float CUDA waveforms to the library's resampler: the 24 kHz -> 16 kHz conversion of
    every generated turn in ``generate_dialogue`` (fireredtts2.py:389-391) stays on the device; CPU tensors (the prompt
    audio, fireredtts2.py:65) and every other torchaudio attribute go to torchaudio itself.  Only that module's binding
    changes, ``torchaudio`` as imported anywhere else is untouched.
The text tokenizer, prompt preparation, the LLM backbone and the dialogue loop are the reference's own code.  There is no
CPU fallback: on a machine without the library / a CUDA device the constructors raise and the object is left untouched.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Optional

import torch


class _FunctionalProxy:
    """``torchaudio.functional`` as the reference's module sees it after ``install``."""

    def __init__(self, real, resample_fn):
        self._real, self._resample = real, resample_fn

    def __getattr__(self, name):
        return getattr(self._real, name)

    def resample(self, waveform, orig_freq, new_freq, *args, **kw):
        """The call at fireredtts2.py:389-391 (positional rates, torchaudio's defaults).  Anything else — a CPU tensor,
        non-default filter arguments, an integer waveform — is torchaudio's."""
        if not args and not kw and getattr(waveform, "is_cuda", False) and waveform.is_floating_point():
            return self._resample(waveform, int(orig_freq), int(new_freq))
        return self._real.resample(waveform, orig_freq, new_freq, *args, **kw)


class _TorchaudioProxy:
    def __init__(self, real, resample_fn):
        self._real = real
        self.functional = _FunctionalProxy(real.functional, resample_fn)

    def __getattr__(self, name):
        return getattr(self._real, name)


@dataclasses.dataclass
class Installed:
    """What ``install`` replaced; ``uninstall()`` puts the reference's objects back."""
    tts: Any
    reference_codec: Any
    codec: Any
    generate_frame: Optional[Any]
    module: Optional[Any] = None            # the reference module whose ``torchaudio`` binding was swapped
    reference_torchaudio: Optional[Any] = None

    def uninstall(self) -> None:
        if self.tts._audio_tokenizer is self.codec:
            self.tts._audio_tokenizer = self.reference_codec
        if self.generate_frame is not None:
            self.generate_frame.uninstall()
        if self.module is not None and isinstance(getattr(self.module, "torchaudio", None), _TorchaudioProxy):
            self.module.torchaudio = self.reference_torchaudio


def install(tts, native_encode: bool = False, frame_tail: bool = True, max_batch: int = 8, seed: int = 0,
            num_heads: Optional[int] = None, resample: bool = True) -> Installed:
    """Both objects are built BEFORE anything is assigned, so a failure (no GPU, unsupported widths) leaves ``tts`` as it was."""
    import sys
    from . import codec as codec_mod
    from .codec import RedCodecB200
    from .frame_decoder import FrameDecoderB200, GenerateFrameB200
    device = getattr(tts, "device", "cuda:0")
    ref_codec = tts._audio_tokenizer
    codec = RedCodecB200.from_reference(ref_codec, num_heads=num_heads, device=str(device), native_encode=native_encode)
    tail = FrameDecoderB200.from_reference(tts._model, device=str(device), max_batch=max_batch) if frame_tail else None
    tts._audio_tokenizer = codec
    gen = GenerateFrameB200.install(tts._model, tail=tail, seed=seed) if frame_tail else None
    module = sys.modules.get(type(tts).__module__) if resample else None
    real_ta = getattr(module, "torchaudio", None) if module is not None else None
    if real_ta is None or isinstance(real_ta, _TorchaudioProxy) or not hasattr(real_ta, "functional"):
        module = real_ta = None                 # nothing to swap (or already swapped by an earlier install)
    else:
        module.torchaudio = _TorchaudioProxy(real_ta, lambda w, o, n: codec_mod.resample(w, o, n))
    return Installed(tts, ref_codec, codec, gen, module, real_ta)


def _feed_back(sample, pos):
    """The next frame's LM input from the frame just sampled (fireredtts2.py:326-336, the live loop's :183-192): one row
    ``[code_0 .. code_{nq-1}, 0]`` with the text column masked out, at the next position."""
    b, nq = sample.shape
    tokens = torch.zeros((b, 1, nq + 1), dtype=torch.long, device=sample.device)
    tokens[:, 0, :nq] = sample
    mask = torch.ones((b, 1, nq + 1), dtype=torch.bool, device=sample.device)
    mask[:, 0, nq] = False
    return tokens, mask, pos[:, -1:] + 1


@torch.inference_mode()      # on a generator function torch enters the mode around every resumption, not across yields
def generate_stream(tts, text: str, speaker: str, context, max_audio_length_ms: float = 90_000, temperature: float = 0.9,
                    topk: int = 50, pcm16: bool = True, decoder_factory=None):
    """The reference's commented-out ``FireRedTTS2.generate_stream`` (fireredtts2.py:259-343) revived on the library
    (SURVEY §8f.1): same arguments, same prompt assembly through the object's own ``_tokenize_segment`` /
    ``_tokenize_text_segment``, same frame loop and stopping rule — and one chunk per generated frame, delivered with the
    same one-frame delay (the chunk of frame *i* is yielded once frame *i+1* exists, the last one with ``last_token=True``).

    What differs from the commented-out code: the codec step does not run between two LLM frames on the caller's stream
    (``decode_one_token`` at fireredtts2.py:317-321, 338-342) but on ``StreamDecoder``'s side stream — one captured graph replay
    straight to int16 PCM in pinned host memory — so it overlaps ``generate_frame`` of the next frame; what is yielded is a
    ``StreamChunk`` (``.samples`` (1, n) pinned host tensor, n = 1560 first / 1920 / 2280 last at hop 240; ``.ready`` the
    CUDA event to wait on before reading; ``.index``) instead of a device tensor.  ``pcm16=False`` gives fp32 samples.
    ``tts._audio_tokenizer`` must be a ``RedCodecB200`` (``dropin.install(tts)``); an immediate end-of-speech frame yields
    nothing (the commented-out code would fail on ``prev_sample.unsqueeze``)."""
    from .codec import RedCodecB200, StreamDecoder
    max_generation_len = int(max_audio_length_ms / 80)                                  # fireredtts2.py:271
    if decoder_factory is None:
        codec = tts._audio_tokenizer
        if not isinstance(codec, RedCodecB200):
            raise TypeError("generate_stream needs the library's codec behind tts._audio_tokenizer: call dropin.install(tts) first")
        decoder_factory = lambda: StreamDecoder(codec, batch=1, pcm16=pcm16, max_tokens=max(max_generation_len, 1))  # noqa: E731
    tts._model.reset_caches()                                                       # fireredtts2.py:269
    tokens, tokens_mask = [], []
    for segment in context:                                                         # fireredtts2.py:273-276
        segment_tokens, segment_tokens_mask = tts._tokenize_segment(segment)
        tokens.append(segment_tokens)
        tokens_mask.append(segment_tokens_mask)
    gen_segment_tokens, gen_segment_tokens_mask = tts._tokenize_text_segment(text, speaker)   # fireredtts2.py:278-282
    tokens.append(gen_segment_tokens)
    tokens_mask.append(gen_segment_tokens_mask)
    prompt_tokens = torch.cat(tokens, dim=0).long().to(tts.device)
    prompt_tokens_mask = torch.cat(tokens_mask, dim=0).bool().to(tts.device)
    curr_tokens = prompt_tokens.unsqueeze(0)
    curr_tokens_mask = prompt_tokens_mask.unsqueeze(0)
    curr_pos = torch.arange(0, prompt_tokens.size(0)).unsqueeze(0).long().to(tts.device)
    max_seq_len = 3100                                                              # fireredtts2.py:294-299
    max_context_len = max_seq_len - max_generation_len
    if curr_tokens.size(1) >= max_context_len:
        raise ValueError(f"Inputs too long, must be below max_seq_len - max_generation_len: {max_context_len}")
    dec = decoder_factory()
    for _ in range(max_generation_len):                                             # fireredtts2.py:305-336
        sample = tts._model.generate_frame(curr_tokens, curr_tokens_mask, curr_pos, temperature, topk)
        if torch.all(sample == 0):                                                  # eos
            break
        chunk = dec.push(sample)            # codec step of the PREVIOUS frame, on the side stream
        if chunk is not None:
            yield chunk
        curr_tokens, curr_tokens_mask, curr_pos = _feed_back(sample, curr_pos)
    last = dec.finish()                                                             # last_token=True, fireredtts2.py:338-343
    if last is not None:
        yield last
