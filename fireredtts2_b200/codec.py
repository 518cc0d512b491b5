"""Drop-in replacement for the reference's codec decode entry points.

``RedCodecB200`` duck-types what ``FireRedTTS2`` uses of ``RedCodecInfer`` (reference
``fireredtts2/codec/model.py:197-376``; call sites ``fireredtts2/fireredtts2.py:51-53,96,196,441``):

* ``decode(tokens (B,nq,L) int32|int64, any strides) -> (B, 1920*L) float32``
* ``decode_one_token(token (B,nq,Lc), cache_dict, last_token) -> (audio (B,n), new_cache_dict)``
* ``encode(...)`` — delegated to a wrapped reference module when one is given (encode is out of scope).

All compute runs in libfrt2_b200.so (hand-written sm_100a CUDA) through the C ABI in ``include/frt2.h``;
PyTorch only provides device memory and the current CUDA stream.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Iterable, Iterator, NamedTuple, Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .config import SAMPLE_RATE, CodecConfig
from .weights import decode_keys, normalise_state_dict

_STATE_KEY = "frt2_state"
_REF_CACHE_KEYS = ("up_conv_cache", "bb_conv_cache1", "bb_conv_cache2", "bb_kv_cache", "is_cache")


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int, lengths: Optional[torch.Tensor] = None
             ) -> torch.Tensor:
    """Drop-in for ``torchaudio.functional.resample(waveform, orig_freq, new_freq)`` with its defaults, as the reference
    calls it on every generated turn (24 kHz -> 16 kHz for the context loop, fireredtts2.py:389-391) and on the prompt
    (fireredtts2.py:65): ``(..., time)`` float32 CUDA tensor -> ``(..., ceil(new * time / orig))``.  Extension:
    ``lengths`` (int32, one per row) resamples ragged rows as if each were ``lengths[i]`` long (rest of the row 0)."""
    if orig_freq <= 0 or new_freq <= 0:
        raise ValueError("Original frequency and desired frequecy should be positive")
    if not waveform.is_floating_point():
        raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
    if waveform.device.type != "cuda":
        raise ValueError("fireredtts2_b200.codec.resample runs on CUDA tensors only (no CPU fallback)")
    lib = N.load()
    if int(orig_freq) == int(new_freq):
        return waveform
    shape = waveform.shape
    x = waveform.to(torch.float32).reshape(int(np.prod(shape[:-1], dtype=np.int64)), shape[-1])
    if x.stride(1) != 1:
        x = x.contiguous()
    B, n = x.shape
    g = np.gcd(int(orig_freq), int(new_freq))
    o, w = int(orig_freq) // g, int(new_freq) // g
    n_out = -(-w * n // o)
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(dev):
        y = torch.empty((B, n_out), dtype=torch.float32, device=x.device)
        lptr = None
        if lengths is not None:
            lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
            if lengths.numel() != B:
                raise ValueError("lengths must have one entry per row")
            lptr = C.c_void_p(lengths.data_ptr())
        got = C.c_int64(0)
        if B and n:
            N.check(lib.frt2_resample(dev, C.c_void_p(x.data_ptr()), x.stride(0), B, n, lptr, int(orig_freq),
                                      int(new_freq), C.c_void_p(y.data_ptr()), y.stride(0) if n_out else 1,
                                      C.byref(got), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            assert got.value == n_out
    return y.reshape(shape[:-1] + (n_out,))


class _NativeStream:
    """Owns a frt2_stream (the in-HBM equivalent of the reference's cache_dict)."""

    def __init__(self, codec: "RedCodecB200", batch: int, max_tokens: int):
        self.codec = codec
        self.batch = batch
        self.max_tokens = max_tokens
        self.ptr = C.c_void_p()
        N.check(codec._lib.frt2_stream_create(codec._h, batch, max_tokens, C.byref(self.ptr)))
        self.finished = False

    @property
    def n_tokens(self) -> int:
        return int(self.codec._lib.frt2_stream_tokens(self.ptr))

    def close(self):
        if self.ptr:
            self.codec._lib.frt2_stream_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RedCodecB200(torch.nn.Module):
    def __init__(self, cfg: CodecConfig, state_dict, device="cuda:0", encoder: Optional[torch.nn.Module] = None,
                 stream_max_tokens: int = 1200, check_indices: bool = True):
        super().__init__()
        self._lib = N.load()   # raises if the CUDA extension is missing
        if not torch.cuda.is_available():
            raise RuntimeError("RedCodecB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("RedCodecB200 runs on CUDA devices only")
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.stream_max_tokens = stream_max_tokens
        self.check_indices = check_indices
        self._encoder = [encoder]  # list: keep the reference module out of nn.Module registration
        self._h = C.c_void_p()
        c = N.Frt2Config(cfg.rvq_dim, cfg.output_dim, cfg.num_quantizers, cfg.codebook_size, cfg.codebook_dim,
                         cfg.embed_dim, cfg.num_layers, cfg.num_heads, cfg.hop_length, cfg.upconv_stride)
        N.check(self._lib.frt2_create(C.byref(c), self.device_index, C.byref(self._h)))
        sd = normalise_state_dict(state_dict)
        for key in decode_keys(cfg):
            if key not in sd:
                raise KeyError(f"state_dict is missing decode-path tensor {key!r}")
            a = np.ascontiguousarray(sd[key], dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            N.check(self._lib.frt2_load_tensor(self._h, key.encode(), a.ctypes.data_as(C.c_void_p), a.ndim, shape, 0))
        # encode-side tensors of the RVQ (optional: present in a full RedCodec checkpoint, absent from decode-only dicts)
        self.has_rvq_encoder = False
        for key, v in sd.items():
            if key.startswith("rvq.") and (".in_project." in key or key.startswith("rvq.input_proj.")):
                a = np.ascontiguousarray(v, dtype=np.float32)
                shape = (C.c_int64 * a.ndim)(*a.shape)
                N.check(self._lib.frt2_load_tensor(self._h, key.encode(), a.ctypes.data_as(C.c_void_p), a.ndim, shape, 0))
                self.has_rvq_encoder = True
        if not cfg.has_out_project:
            self.has_rvq_encoder = True      # Identity in_project: the codebooks are all the encoder needs
        N.check(self._lib.frt2_finalize(self._h))

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_reference(cls, ref: torch.nn.Module, num_heads: Optional[int] = None, **kw) -> "RedCodecB200":
        """Build from a live reference ``RedCodecInfer`` (weights copied once; ``encode`` delegates to it)."""
        ad = ref.acoustic_decoder
        q0 = ref.rvq.quantizers[0]
        cfg = CodecConfig(rvq_dim=ref.rvq.rvq_dim, output_dim=ad.embed_dim, num_quantizers=len(ref.rvq.quantizers),
                          codebook_size=q0.codebook.shape[0], codebook_dim=q0.codebook.shape[1],
                          embed_dim=ad.embed_dim, num_layers=ad.num_layers,
                          num_heads=num_heads or ad.num_heads, hop_length=ad.hop_length)
        if not getattr(ad, "causal", True):
            raise AssertionError("Only AcousticDecoder with causal=True supports forward_chunk method.")
        native_encode = kw.pop("native_encode", False)
        codec = cls(cfg, ref.state_dict(), encoder=ref, **kw)
        if native_encode:       # the whole encode path on the GPU library too; the reference stays unused
            from .encoder import CodecEncoderB200, EncoderConfig
            codec.attach_encoder(CodecEncoderB200(EncoderConfig.from_reference_module(ref), ref.state_dict(),
                                                  device=str(codec.device)))
        return codec

    @classmethod
    def from_pretrained(cls, conf_path: str, ckpt_path: str, **kw) -> "RedCodecB200":
        """Same arguments as ``RedCodecInfer.from_pretrained`` (reference model.py:210-216); decode-only unless
        ``native_encode=True`` (then ``encode`` runs on the library too, fireredtts2_b200/encoder.py)."""
        native_encode = kw.pop("native_encode", False)
        with open(conf_path, "r") as f:
            conf = json.load(f)
        cfg = CodecConfig.from_reference_dict(conf)
        ckpt = torch.load(ckpt_path, map_location="cpu")["generator"]
        codec = cls(cfg, ckpt, **kw)
        if native_encode:       # encode() on the library as well: the checkpoint holds ssl.*, ssl_adaptor.*, acoustic_encoder.*, downsample.*
            from .encoder import CodecEncoderB200, EncoderConfig
            c = dict(conf.get("codec", conf), with_feature_encoders=True)
            codec.attach_encoder(CodecEncoderB200(EncoderConfig.from_reference_dict(c), ckpt, device=str(codec.device)))
        return codec

    # ------------------------------------------------------------------ nn.Module compatibility
    def to(self, *args, **kwargs):  # weights live in the native handle on self.device
        return self

    def __del__(self):
        try:
            if self._h:
                self._lib.frt2_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _cuda_stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)

    def _prep_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        if tokens.dim() != 3:
            raise ValueError(f"tokens must be (B, nq, L), got {tuple(tokens.shape)}")
        if tokens.dtype not in (torch.int32, torch.int64):
            if tokens.dtype.is_floating_point or tokens.dtype == torch.bool:
                raise TypeError(f"tokens must be an integer tensor, got {tokens.dtype}")
            tokens = tokens.long()
        if tokens.device.type != "cuda" or tokens.device.index != self.device_index:
            tokens = tokens.to(torch.device("cuda", self.device_index))
        return tokens

    def _maybe_check(self):
        if self.check_indices:
            N.check(self._lib.frt2_check_error(self._h, self._cuda_stream()))

    # ------------------------------------------------------------------ reference API
    @torch.inference_mode()
    def decode(self, tokens: torch.Tensor, lengths: Optional[torch.Tensor] = None, pcm16: bool = False) -> torch.Tensor:
        """RedCodecInfer.decode (reference model.py:307-324).  Extensions: ``lengths`` (B,) int32 token counts for
        ragged batches (item b equals a standalone decode of its first lengths[b] tokens — within the parity tolerance where the
        two batch shapes are served by different kernels; a count <= 0 gives zeros, one beyond L is L); ``pcm16=True`` returns
        the int16 PCM of the reference's wire format, ``(audio * 32767).astype(int16)``, straight from the kernel."""
        tokens = self._prep_tokens(tokens)
        B, nq, L = tokens.shape
        out_dtype = torch.int16 if pcm16 else torch.float32
        if L == 0 or B == 0:
            return torch.zeros((B, 0), dtype=out_dtype, device=tokens.device)
        with torch.cuda.device(self.device_index):
            audio = torch.empty((B, self.cfg.samples_per_token * L), dtype=out_dtype, device=tokens.device)
            lptr = None
            if lengths is not None:
                lengths = lengths.to(device=tokens.device, dtype=torch.int32).contiguous()
                if lengths.numel() != B:
                    raise ValueError("lengths must have B entries")
                lptr = C.c_void_p(lengths.data_ptr())
            sB, sQ, sL = tokens.stride()
            fn = self._lib.frt2_decode_pcm16 if pcm16 else self._lib.frt2_decode
            N.check(fn(self._h, C.c_void_p(tokens.data_ptr()), tokens.element_size(), sB, sQ, sL,
                       B, nq, L, lptr, C.c_void_p(audio.data_ptr()), audio.stride(0), self._cuda_stream()))
            self._maybe_check()
        return audio

    @torch.inference_mode()
    def decode_resampled(self, tokens: torch.Tensor, new_freq: int = 16000, lengths: Optional[torch.Tensor] = None,
                         return_native: bool = True):
        """``decode`` followed by ``torchaudio.functional.resample(audio, 24000, new_freq)`` as the context loop runs them on
        every generated turn (reference fireredtts2.py:386-391), with the resampler fused into the overlap-add kernel: the
        24 kHz waveform makes no HBM round trip in between.  -> ``(audio24k (B, 1920 L) or None, audio_rs (B, ceil(new *
        1920 L / 24000)))``, both bit-identical to the two separate calls."""
        tokens = self._prep_tokens(tokens)
        B, nq, L = tokens.shape
        if int(new_freq) <= 0:
            raise ValueError("Original frequency and desired frequecy should be positive")
        g = int(np.gcd(SAMPLE_RATE, int(new_freq)))
        o, w = SAMPLE_RATE // g, int(new_freq) // g
        n_in = self.cfg.samples_per_token * L
        n_rs = -(-w * n_in // o)
        if L == 0 or B == 0:
            z = torch.zeros((B, 0), dtype=torch.float32, device=tokens.device)
            return (z if return_native else None), z
        with torch.cuda.device(self.device_index):
            audio = torch.empty((B, n_in), dtype=torch.float32, device=tokens.device) if return_native else None
            audio_rs = torch.empty((B, n_rs), dtype=torch.float32, device=tokens.device)
            lptr = None
            if lengths is not None:
                lengths = lengths.to(device=tokens.device, dtype=torch.int32).contiguous()
                if lengths.numel() != B:
                    raise ValueError("lengths must have B entries")
                lptr = C.c_void_p(lengths.data_ptr())
            sB, sQ, sL = tokens.stride()
            got = C.c_int64(0)
            N.check(self._lib.frt2_decode_resampled(
                self._h, C.c_void_p(tokens.data_ptr()), tokens.element_size(), sB, sQ, sL, B, nq, L, lptr,
                C.c_void_p(audio.data_ptr()) if audio is not None else None, audio.stride(0) if audio is not None else 0,
                SAMPLE_RATE, int(new_freq), C.c_void_p(audio_rs.data_ptr()), audio_rs.stride(0), C.byref(got),
                self._cuda_stream()))
            assert got.value == n_rs
            self._maybe_check()
        return audio, audio_rs

    @torch.inference_mode()
    def decode_into(self, tokens: torch.Tensor, out_ptr: int, out_off: torch.Tensor,
                    lengths: Optional[torch.Tensor] = None, pcm16: bool = False) -> None:
        """Offline decode whose items are scattered: item b's ``1920 * L_b`` samples are written at element offset
        ``out_off[b]`` of the buffer at device address ``out_ptr`` (fp32, or int16 PCM with ``pcm16``) and nothing
        beyond them.  ``out_ptr`` may be a local tensor's ``data_ptr()`` — the turns of a dialogue land at their place
        in the concatenated waveform (reference fireredtts2.py:399-401) — or a buffer of another GPU mapped with
        ``sharding.PeerBuffer``, in which case the overlap-add kernel's stores go over NVLink (frt2_decode_scatter)."""
        tokens = self._prep_tokens(tokens)
        B, nq, L = tokens.shape
        if L == 0 or B == 0:
            return
        with torch.cuda.device(self.device_index):
            out_off = out_off.to(device=tokens.device, dtype=torch.int64).contiguous()
            if out_off.numel() != B:
                raise ValueError("out_off must have B entries")
            lptr = None
            if lengths is not None:
                lengths = lengths.to(device=tokens.device, dtype=torch.int32).contiguous()
                if lengths.numel() != B:
                    raise ValueError("lengths must have B entries")
                lptr = C.c_void_p(lengths.data_ptr())
            sB, sQ, sL = tokens.stride()
            N.check(self._lib.frt2_decode_scatter(
                self._h, C.c_void_p(tokens.data_ptr()), tokens.element_size(), sB, sQ, sL, B, nq, L, lptr,
                C.c_void_p(int(out_ptr)), 1 if pcm16 else 0, C.c_void_p(out_off.data_ptr()), self._cuda_stream()))
            self._maybe_check()

    @torch.inference_mode()
    def decode_one_token(self, token: torch.Tensor, cache_dict: Dict[str, object], last_token: bool,
                         pcm16: bool = False, _check: bool = True) -> Tuple[torch.Tensor, Dict[str, object]]:
        """RedCodecInfer.decode_one_token (reference model.py:326-376).  Extension: ``pcm16=True`` returns the chunk
        as the int16 PCM the reference's streaming front puts on the wire (enhanced_fireredtts2.py:603,655).

        ``cache_dict`` is ``{}`` on the first call.  The returned dict carries the opaque in-HBM state under
        ``"frt2_state"`` (updated in place — re-using an *old* dict to fork a stream is not supported; use
        ``export_cache`` / a reference-layout dict for that).  A dict holding the reference's five tensors
        (hand-off from the reference implementation) is imported."""
        token = self._prep_tokens(token)
        B, nq, Lc = token.shape
        if Lc < 1:
            raise ValueError("decode_one_token needs at least one token")
        with torch.cuda.device(self.device_index):
            st = cache_dict.get(_STATE_KEY) if cache_dict else None
            if st is None:
                st = _NativeStream(self, B, self.stream_max_tokens)
                if cache_dict and all(k in cache_dict for k in _REF_CACHE_KEYS):
                    self._import_reference_cache(st, cache_dict)
            if st.batch != B:
                raise ValueError(f"stream was created for batch {st.batch}, got {B}")
            if st.finished:
                raise ValueError("stream already received its last token")
            first = st.n_tokens == 0
            n = self.cfg.samples_per_token * Lc - self.cfg.istft_pad * first + self.cfg.istft_pad * bool(last_token)
            audio = torch.empty((B, n), dtype=torch.int16 if pcm16 else torch.float32, device=token.device)
            n_out = C.c_int(0)
            sB, sQ, sL = token.stride()
            fn = self._lib.frt2_decode_chunk_pcm16 if pcm16 else self._lib.frt2_decode_chunk
            N.check(fn(self._h, st.ptr, C.c_void_p(token.data_ptr()), token.element_size(),
                       sB, sQ, sL, nq, Lc, int(bool(last_token)),
                       C.c_void_p(audio.data_ptr()), audio.stride(0), C.byref(n_out), self._cuda_stream()))
            assert n_out.value == n
            st.finished = bool(last_token)
            if _check and self.check_indices:   # IndexError inside the offending call, like the reference (rvq.py:58)
                N.check(self._lib.frt2_stream_check_error(self._h, st.ptr, None, self._cuda_stream()))
        return audio, {_STATE_KEY: st}

    def new_stream(self, batch: int = 1, max_tokens: Optional[int] = None) -> Dict[str, object]:
        """A stream state up front; pass the returned dict as ``cache_dict`` of the first ``decode_one_token`` call.
        Not needed for latency: the native handle recycles the states of finished utterances, so the reference's own
        call pattern ``decode_one_token(tok, {}, last)`` pays no allocation either (see ``reserve_streams``)."""
        with torch.cuda.device(self.device_index):
            return {_STATE_KEY: _NativeStream(self, batch, max_tokens or self.stream_max_tokens)}

    def reserve_streams(self, count: int = 1, batch: int = 1, max_tokens: Optional[int] = None) -> None:
        """Keep ``count`` spare stream states (device buffers + the captured per-token step) in the handle, so that the
        first ``decode_one_token(tok, {}, ...)`` calls of a freshly loaded model are as fast as every later one."""
        with torch.cuda.device(self.device_index):
            N.check(self._lib.frt2_stream_reserve(self._h, batch, max_tokens or self.stream_max_tokens, count))

    def decode_stream(self, frames: Iterable[torch.Tensor], pcm16: bool = True, batch: int = 1,
                      ring: int = 16) -> Iterator["StreamChunk"]:
        """The codec half of ``FireRedTTS2.generate_stream`` (reference fireredtts2.py:259-343; commented out there):
        consume the frames the LLM emits one by one and yield the waveform chunk of frame *i-1* when frame *i* arrives
        (one frame of delay, exactly as the reference's loop holds ``prev_sample`` back so that ``last_token`` is known),
        then the final chunk.  See ``StreamDecoder``."""
        dec = StreamDecoder(self, batch=batch, pcm16=pcm16, ring=ring)
        for f in frames:
            c = dec.push(f)
            if c is not None:
                yield c
        c = dec.finish()
        if c is not None:
            yield c

    def new_pool(self, slots: int, max_tokens: Optional[int] = None) -> "StreamPool":
        """A pool of ``slots`` concurrent streams that decode one token each per step (continuous batching)."""
        return StreamPool(self, slots, max_tokens or self.stream_max_tokens)

    def reset_stream(self, cache_dict: Dict[str, object]) -> Dict[str, object]:
        """Return a used state to its initial (empty) condition for the next utterance.  Stream-ordered: the state is
        cleared by a kernel on the CUDA stream of the next ``decode_one_token`` call, nothing synchronises."""
        st: _NativeStream = cache_dict[_STATE_KEY]
        N.check(self._lib.frt2_stream_reset(st.ptr))
        st.finished = False
        return cache_dict

    def export_cache(self, cache_dict: Dict[str, object]) -> Dict[str, torch.Tensor]:
        """The state in the reference's own cache_dict layouts (model.py:346-375), for parity / hand-off."""
        st: _NativeStream = cache_dict[_STATE_KEY]
        c, B, E = self.cfg, st.batch, self.cfg.embed_dim
        T = 8 * st.n_tokens
        dev = torch.device("cuda", self.device_index)
        out = {
            "up_conv_cache": torch.empty((B, E, 3), device=dev),
            "bb_conv_cache1": torch.empty((B, E, 6), device=dev),
            "bb_conv_cache2": torch.empty((B, 8 * E, 2), device=dev),
            "bb_kv_cache": torch.empty((B, c.num_layers, c.num_heads, T, 2 * c.head_dim), device=dev),
            "is_cache": torch.empty((B, c.n_fft, 3), device=dev),
        }
        with torch.cuda.device(self.device_index):
            N.check(self._lib.frt2_export_state(self._h, st.ptr, *[C.c_void_p(out[k].data_ptr()) for k in _REF_CACHE_KEYS],
                                                self._cuda_stream()))
        return out

    def _import_reference_cache(self, st: _NativeStream, cache: Dict[str, torch.Tensor]):
        dev = torch.device("cuda", self.device_index)
        ts = [cache[k].to(device=dev, dtype=torch.float32).contiguous() for k in _REF_CACHE_KEYS]
        T = ts[3].shape[3]
        if T % 8:
            raise ValueError("bb_kv_cache length must be a multiple of 8 frames (whole tokens)")
        N.check(self._lib.frt2_import_state(self._h, st.ptr, T // 8, *[C.c_void_p(t.data_ptr()) for t in ts],
                                            self._cuda_stream()))
        torch.cuda.current_stream(self.device_index).synchronize()

    resample = staticmethod(resample)   # torchaudio.functional.resample of the context loop, on the same device

    def attach_encoder(self, native_encoder) -> None:
        """Route ``encode`` through a ``CodecEncoderB200`` (fireredtts2_b200/encoder.py) instead of the reference module."""
        self._native_encoder = native_encoder

    def encode(self, *args, **kwargs):
        """``RedCodecInfer.encode`` (model.py:243-305): through the attached ``CodecEncoderB200`` when there is one
        (SURVEY.md §8f.3), else delegated to the wrapped reference module."""
        ne = getattr(self, "_native_encoder", None)
        if ne is not None:
            audio16k = args[0] if args else kwargs["audio16k"]
            length = args[1] if len(args) > 1 else kwargs.get("audio16k_length")
            bs = args[2] if len(args) > 2 else kwargs.get("batch_size", 96)
            return ne.encode(audio16k, length, self, bs)
        enc = self._encoder[0]
        if enc is None:
            raise NotImplementedError("encode() needs the reference RedCodecInfer (use RedCodecB200.from_reference)")
        return enc.encode(*args, **kwargs)

    @torch.inference_mode()
    def rvq_encode_codes(self, z: torch.Tensor, nq: Optional[int] = None) -> torch.Tensor:
        """``ResidualVQ.encode_codes`` (reference rvq.py:128-143): z (B, input_dim, T) fp32, any strides (the
        reference passes a transposed view, model.py:240) -> codes (nq, B, T) int64.  One CUDA kernel
        (frt2_rvq_encode); needs the checkpoint's ``in_project`` / ``input_proj`` tensors in the state_dict."""
        if z.dim() != 3:
            raise ValueError(f"z must be (B, input_dim, T), got {tuple(z.shape)}")
        dev = torch.device("cuda", self.device_index)
        z = z.to(device=dev, dtype=torch.float32)
        B, D, T = z.shape
        nq = self.cfg.num_quantizers if nq is None else nq
        with torch.cuda.device(self.device_index):
            codes = torch.empty((nq, B, T), dtype=torch.int64, device=dev)
            if B * T > 0:
                sB, sD, sT = z.stride()
                N.check(self._lib.frt2_rvq_encode(self._h, C.c_void_p(z.data_ptr()), sB, sD, sT, B, D, T, nq,
                                                  C.c_void_p(codes.data_ptr()), self._cuda_stream()))
        return codes

    # ------------------------------------------------------------------ parity hooks
    def rvq_gather(self, tokens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Raw codebook rows (B,L,nq,cd) and their index-ordered sum (B,L,cd), bit-exact with F.embedding."""
        tokens = self._prep_tokens(tokens)
        B, nq, L = tokens.shape
        with torch.cuda.device(self.device_index):
            rows = torch.empty((B, L, nq, self.cfg.codebook_dim), dtype=torch.float32, device=tokens.device)
            s = torch.empty((B, L, self.cfg.codebook_dim), dtype=torch.float32, device=tokens.device)
            sB, sQ, sL = tokens.stride()
            N.check(self._lib.frt2_rvq_gather(self._h, C.c_void_p(tokens.data_ptr()), tokens.element_size(), sB, sQ, sL,
                                              B, nq, L, C.c_void_p(rows.data_ptr()), C.c_void_p(s.data_ptr()),
                                              self._cuda_stream()))
            N.check(self._lib.frt2_check_error(self._h, self._cuda_stream()))
        return rows, s

    def profile(self, enable: bool):
        """Bracket every kernel of the following decode calls with CUDA events (bench.py roofline numbers)."""
        N.check(self._lib.frt2_profile(self._h, int(enable)))

    def profile_get(self, cls: int) -> Dict[str, float]:
        ms, n, fl, by = C.c_double(0), C.c_int64(0), C.c_double(0), C.c_double(0)
        N.check(self._lib.frt2_profile_get(self._h, cls, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)))
        return {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}

    def set_debug(self, flags: int):
        N.check(self._lib.frt2_set_debug(self._h, flags))

    def get_tap(self, name: str, shape) -> torch.Tensor:
        out = torch.empty(tuple(shape), dtype=torch.float32, device=torch.device("cuda", self.device_index))
        n = C.c_int64(0)
        N.check(self._lib.frt2_get_tap(self._h, name.encode(), C.c_void_p(out.data_ptr()), out.numel(), C.byref(n),
                                       self._cuda_stream()))
        if n.value != out.numel():
            raise ValueError(f"tap {name} has {n.value} elements, expected {out.numel()}")
        return out


class StreamChunk(NamedTuple):
    """One streamed chunk: ``samples`` (B, n) in pinned HOST memory (int16 PCM or fp32), valid once ``ready`` (a CUDA
    event) has completed — call ``ready.synchronize()`` before reading the bytes; ``index`` = token index.  The buffer
    belongs to a ring and is re-used ``ring`` chunks later: send or copy it before then."""
    samples: torch.Tensor
    ready: torch.cuda.Event
    index: int


class StreamDecoder:
    """Live streaming front for one request.

    The reference's ``generate_stream`` (fireredtts2.py:259-343) calls ``decode_one_token(prev_sample, cache, last)``
    between two LLM frames on the default stream, so the LLM waits for the codec and the host for both.  Here each
    codec step (one CUDA-graph replay, int16 PCM straight from the kernel) and its device->host copy run on a
    dedicated CUDA stream that only waits for the event marking the frame's tokens: the LLM's next frame, enqueued on
    the caller's stream, overlaps the codec step, and the host gets the wire-format bytes without synchronising the
    caller's stream.  ``push(frame)`` returns the chunk of the PREVIOUS frame (None for the first frame);
    ``finish()`` flushes the last frame with ``last_token=True``."""

    def __init__(self, codec: RedCodecB200, batch: int = 1, pcm16: bool = True, max_tokens: Optional[int] = None,
                 ring: int = 16, timing: bool = False):
        self.codec = codec
        self.batch = batch
        self.pcm16 = pcm16
        self.timing = timing      # chunk events carry timestamps (measurements only)
        dev = torch.device("cuda", codec.device_index)
        with torch.cuda.device(codec.device_index):
            self._cache = codec.new_stream(batch, max_tokens)
            self._side = torch.cuda.Stream(device=dev, priority=-1)   # the step is tiny and latency-critical
            width = codec.cfg.samples_per_token + codec.cfg.istft_pad
            dt = torch.int16 if pcm16 else torch.float32
            self._ring = [torch.empty((batch, width), dtype=dt).pin_memory() for _ in range(max(2, ring))]
            # the stream's device-side error words (word 0: any item sent an out-of-range code) ride along with every
            # chunk into pinned memory and are tested one step late, once the chunk's event has completed
            self._err_ring = [torch.zeros(1 + batch, dtype=torch.int32).pin_memory() for _ in range(max(2, ring))]
        self._held: Optional[Tuple[torch.Tensor, torch.cuda.Event]] = None
        self._unchecked = []      # (ready event, error words, chunk index) of chunks whose words were not tested yet
        self._n = 0
        self._done = False

    def _check_completed(self, wait: bool = False) -> None:
        """IndexError for an out-of-range code, raised to THIS request (the words belong to this stream) at the first
        push / finish after the offending chunk has left the device."""
        keep = []
        for ev, words, idx in self._unchecked:
            if wait:
                ev.synchronize()
            if not ev.query():
                keep.append((ev, words, idx))
            elif int(words[0]) != 0:
                self._unchecked = []
                c = self.codec
                st = self._cache[_STATE_KEY]
                with torch.cuda.device(c.device_index):   # clears the stream's words
                    c._lib.frt2_stream_check_error(c._h, st.ptr, None, C.c_void_p(self._side.cuda_stream))
                raise IndexError(f"index out of range in self (code frame {idx} of this stream)")
        self._unchecked = keep

    def _as_token(self, frame: torch.Tensor) -> torch.Tensor:
        nq = frame.shape[-2] if frame.dim() == 3 else frame.shape[-1]
        t = frame.reshape(self.batch, nq, 1) if frame.dim() != 3 else frame
        if t.shape[0] != self.batch or t.shape[2] != 1:
            raise ValueError(f"frame must be (nq,), (B,nq) or (B,nq,1) with B={self.batch}, got {tuple(frame.shape)}")
        return self.codec._prep_tokens(t)

    def _decode_held(self, last: bool) -> StreamChunk:
        tok, ev = self._held
        c = self.codec
        with torch.cuda.device(c.device_index), torch.cuda.stream(self._side):
            self._side.wait_event(ev)                       # the frame's tokens are complete
            # no per-step host synchronisation: the stream's error words follow each chunk to the host (below)
            audio, self._cache = c.decode_one_token(tok, self._cache, last, pcm16=self.pcm16, _check=False)
            tok.record_stream(self._side)
            host = self._ring[self._n % len(self._ring)][:, :audio.shape[1]]
            host.copy_(audio, non_blocking=True)
            words = self._err_ring[self._n % len(self._err_ring)]
            N.check(c._lib.frt2_stream_fetch_errors(c._h, self._cache[_STATE_KEY].ptr, C.c_void_p(words.data_ptr()),
                                                    C.c_void_p(self._side.cuda_stream)))
            ready = torch.cuda.Event(enable_timing=self.timing)
            ready.record(self._side)
        self._unchecked.append((ready, words, self._n))
        chunk = StreamChunk(host, ready, self._n)
        self._n += 1
        if last and c.check_indices:
            self._check_completed(wait=True)                # nothing of this stream is left unchecked
        return chunk

    def push(self, frame: torch.Tensor) -> Optional[StreamChunk]:
        if self._done:
            raise ValueError("stream already finished")
        if self.codec.check_indices:
            self._check_completed()
        tok = self._as_token(frame)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.codec.device_index))
        out = self._decode_held(False) if self._held is not None else None
        self._held = (tok, ev)
        return out

    def finish(self) -> Optional[StreamChunk]:
        if self._done:
            return None
        self._done = True
        if self._held is None:
            return None
        out = self._decode_held(True)
        self._held = None
        return out


class StreamPoolIndexError(IndexError):
    """Some slots of a ``StreamPool.step`` sent an out-of-range code (the reference raises IndexError inside the
    offending request's ``decode_one_token``, rvq.py:58).  ``slots`` = the offending slots (their streams have been
    closed); ``results`` = the chunks of every other slot of the step, which are unaffected and must still be delivered."""

    def __init__(self, slots, results):
        super().__init__(f"index out of range in self (pool slots {sorted(slots)})")
        self.slots = sorted(slots)
        self.results = results


class StreamPool:
    """Continuous batching of concurrent ``decode_one_token`` streams (C ABI ``frt2_pool_*``).

    The reference decodes concurrent requests one after the other (one worker thread, enhanced_fireredtts2.py:199-203).
    A pool keeps ``slots`` independent streaming states (reference cache_dict, model.py:346-375) in HBM and advances
    all active ones by one token per ``step`` in a single batched launch sequence — each slot at its own position of
    its own stream, joining and leaving at any step.  A slot's samples do not depend on what the other slots do.

        pool = codec.new_pool(32)
        a = pool.open()                                  # claim a free slot
        out = pool.step({a: tok_a, b: tok_b}, last=[b])  # {slot: (nq,) or (nq,1) int tensor} -> {slot: audio (n,)}
        pool.close(a)                                    # release (implicit after ``last``)
    """

    def __init__(self, codec: RedCodecB200, slots: int, max_tokens: int):
        if not 1 <= slots <= N.POOL_MAX_SLOTS:
            raise ValueError(f"slots must be in [1, {N.POOL_MAX_SLOTS}]")
        self.codec = codec
        self.slots = slots
        self.max_tokens = max_tokens
        self.ptr = C.c_void_p()
        with torch.cuda.device(codec.device_index):
            N.check(codec._lib.frt2_pool_create(codec._h, slots, max_tokens, C.byref(self.ptr)))
            dev = torch.device("cuda", codec.device_index)
            self._tok = torch.zeros((slots, codec.cfg.num_quantizers), dtype=torch.int32, device=dev)
            self._tok_host = torch.zeros((slots, codec.cfg.num_quantizers), dtype=torch.int32).pin_memory()
        self.width = codec.cfg.samples_per_token + codec.cfg.istft_pad
        self._h2d_done = None     # event: the previous step's token upload has left the pinned buffer
        self._free = list(range(slots - 1, -1, -1))
        self._fresh = set()       # opened, no token yet: the first step carries RESET
        self._open = set()

    # -- slot management (host only) --
    def open(self) -> int:
        if not self._free:
            raise RuntimeError("StreamPool: no free slot")
        s = self._free.pop()
        self._open.add(s)
        self._fresh.add(s)
        return s

    def close(self, slot: int):
        if slot in self._open:
            self._open.discard(slot)
            self._fresh.discard(slot)
            self._free.append(slot)

    @property
    def n_open(self) -> int:
        return len(self._open)

    def slot_tokens(self, slot: int) -> int:
        return int(self.codec._lib.frt2_pool_slot_tokens(self.ptr, slot))

    # -- the step --
    @torch.inference_mode()
    def step_dense(self, tokens: torch.Tensor, flags, pcm16: bool = False, bad_slots: Optional[list] = None
                   ) -> Tuple[torch.Tensor, list]:
        """Lowest level: ``tokens`` (slots, nq) int32|int64 on the device, ``flags`` one FRT2_SLOT_* int per slot.
        Returns ``(out (slots, 8*hop+pad) fp32|int16, n_samples per slot)``; slot b's chunk is ``out[b, :n[b]]``.
        Out-of-range codes are reported per slot from the pool's own error words: with ``bad_slots`` (a list) the
        offending slots are appended to it, otherwise IndexError is raised after the step has completed."""
        c = self.codec
        if tokens.dim() != 2 or tokens.shape[0] != self.slots:
            raise ValueError(f"tokens must be (slots={self.slots}, nq), got {tuple(tokens.shape)}")
        if tokens.dtype not in (torch.int32, torch.int64):
            raise TypeError(f"tokens must be int32 or int64, got {tokens.dtype}")
        if len(flags) != self.slots:
            raise ValueError("flags must have one entry per slot")
        with torch.cuda.device(c.device_index):
            out = torch.empty((self.slots, self.width), dtype=torch.int16 if pcm16 else torch.float32,
                              device=tokens.device)
            f = (C.c_int32 * self.slots)(*[int(x) for x in flags])
            n = (C.c_int32 * self.slots)()
            N.check(c._lib.frt2_pool_step(c._h, self.ptr, C.c_void_p(tokens.data_ptr()), tokens.element_size(),
                                          tokens.stride(0), tokens.stride(1), tokens.shape[1], f,
                                          C.c_void_p(out.data_ptr()), int(pcm16), out.stride(0), n, c._cuda_stream()))
            if c.check_indices:
                item = (C.c_int32 * self.slots)()
                rc = c._lib.frt2_stream_check_error(c._h, self.ptr, item, c._cuda_stream())
                if rc == N.ERR_INDEX_OOR:
                    bad = [b for b in range(self.slots) if item[b]]
                    if bad_slots is None:
                        raise IndexError(f"index out of range in self (pool slots {bad})")
                    bad_slots.extend(bad)
                else:
                    N.check(rc)
        return out, list(n)

    @torch.inference_mode()
    def step(self, tokens: Dict[int, torch.Tensor], last=(), pcm16: bool = False) -> Dict[int, torch.Tensor]:
        """One token for each slot in ``tokens``; slots listed in ``last`` end their stream with it (and are released).
        Returns ``{slot: chunk}`` with 1560 / 1920 / 2280-sample chunks exactly as decode_one_token would."""
        last = set(last)
        flags = [0] * self.slots
        nq = None
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        for s, t in tokens.items():
            if s not in self._open:
                raise ValueError(f"slot {s} is not open")
            t = t.reshape(-1)
            nq = t.numel() if nq is None else nq
            if t.numel() != nq:
                raise ValueError("all tokens of a step must use the same number of codebooks")
            self._tok_host[s, :nq] = t.to("cpu", torch.int32) if t.device.type != "cpu" else t.to(torch.int32)
            flags[s] = N.SLOT_ACTIVE | (N.SLOT_RESET if s in self._fresh else 0) | (N.SLOT_LAST if s in last else 0)
        if nq is None:
            return {}
        self._tok.copy_(self._tok_host, non_blocking=True)
        self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.codec.device_index))
        bad: list = []
        # a native-side refusal (bad flag, slot overflow) raises here before any state — native or host — has changed
        out, n = self.step_dense(self._tok[:, :nq], flags, pcm16, bad_slots=bad)
        res = {}
        for s in tokens:          # the native side has advanced every active slot: mirror that first
            self._fresh.discard(s)
            if s in last or s in bad:
                self.close(s)
            if s not in bad:
                res[s] = out[s, :n[s]]
        if bad:                   # one bad token must not cost the other slots their chunk
            raise StreamPoolIndexError([s for s in bad if s in tokens], res)
        return res

    def destroy(self):
        if self.ptr:
            self.codec._lib.frt2_stream_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
