/*
 * libfrt2_b200 — C ABI of the B200-native FireRedTTS-2 codec decoder.
 *
 * The reference has no FFI: its boundary for this path is the Python attribute
 * FireRedTTS2._audio_tokenizer (fireredtts2/fireredtts2.py:51-53) and the methods
 * RedCodecInfer.decode (fireredtts2/codec/model.py:307-324) and
 * RedCodecInfer.decode_one_token (model.py:326-376).  The entry points below are what a binding of
 * that boundary needs; fireredtts2_b200/codec.py is the ctypes binding.
 *
 * All pointers named "device" are CUDA device pointers on the handle's device.  Calls are
 * asynchronous on the caller's stream unless stated otherwise.  Return value: FRT2_OK or a negative
 * frt2_status; frt2_last_error() returns a thread-local message for the last failure.
 */
#ifndef FRT2_H_
#define FRT2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum frt2_status {
  FRT2_OK = 0,
  FRT2_ERR_BAD_ARG = -1,
  FRT2_ERR_BAD_DTYPE = -2,
  FRT2_ERR_INDEX_OUT_OF_RANGE = -3, /* reference: IndexError from F.embedding (codec/rvq.py:58) */
  FRT2_ERR_STATE_OVERFLOW = -4,     /* more tokens than frt2_stream_create reserved */
  FRT2_ERR_CUDA = -5,
  FRT2_ERR_MISSING_TENSOR = -6,
  FRT2_ERR_NOT_FINALIZED = -7
} frt2_status;

/* Mirrors config_codec.json["codec"]["rvq"|"upsample"|"acoustic_decoder"] (model.py:181-184). */
typedef struct frt2_config {
  int32_t rvq_dim;        /* rvq.rvq_dim            (rvq.py:96)  */
  int32_t output_dim;     /* rvq.output_dim == embed_dim        */
  int32_t num_quantizers; /* rvq.num_quantizers     (rvq.py:98)  */
  int32_t codebook_size;  /* rvq.codebook_size      (rvq.py:99)  */
  int32_t codebook_dim;   /* rvq.codebook_dim       (rvq.py:100) */
  int32_t embed_dim;      /* upsample / acoustic_decoder embed_dim (model.py:126, decoder.py:553) */
  int32_t num_layers;     /* acoustic_decoder.num_layers */
  int32_t num_heads;      /* acoustic_decoder.num_heads  */
  int32_t hop_length;     /* acoustic_decoder.hop_length (decoder.py:559) */
  int32_t upconv_stride;  /* upsample.stride (model.py:127), must be 4 */
} frt2_config;

typedef struct frt2_handle frt2_handle; /* weights + workspace; one per device */
typedef struct frt2_stream frt2_stream; /* streaming state of B concurrent utterances */

/* ---- lifecycle: replaces RedCodecInfer.from_pretrained + load_state_dict (model.py:210-216) ---- */
int frt2_create(const frt2_config* cfg, int device, frt2_handle** out);
/* Hand over one tensor of the reference state_dict by its reference key (e.g.
 * "acoustic_decoder.backbone.transformers.3.fc1.weight", "rvq.quantizers.0.codebook",
 * "rvq.output_proj.parametrizations.weight.original0").  fp32, reference layout, contiguous;
 * host pointer (on_device = 0) or device pointer.  Keys the decode path does not use are ignored.  Replaces
 * codec.load_state_dict(ckpt) (model.py:214-215); the weight-norm pairs original0 / original1 are materialised as
 * rvq.py:8-13 does on every forward. */
int frt2_load_tensor(frt2_handle* h, const char* key, const float* data, int ndim, const int64_t* shape,
                     int on_device);
/* One-time repack on the GPU: weight-norm materialisation (rvq.py:8-13), folded RVQ tables, tap-major
 * conv weights, fp16 operands, windowed iDFT basis.  Synchronous. */
int frt2_finalize(frt2_handle* h);
void frt2_destroy(frt2_handle* h);

/* ---- offline decode: RedCodecInfer.decode (model.py:307-324) ----
 * tokens: device, (B,nq,L) integers of idx_bytes (4 or 8) with ELEMENT strides sB,sQ,sL (any strides: the
 * production caller passes a permuted int32 view, fireredtts2.py:196).  audio: device fp32 (B, 8*hop*L) with
 * row pitch audio_pitch elements.  lengths: optional device int32 (B) of per-item token counts L_b <= L
 * (extension; item b then equals a standalone decode of its first L_b tokens — bit for bit when the same kernels serve
 * both batch shapes, within the parity tolerance otherwise — remaining samples are 0; L_b <= 0: all zeros, L_b > L: L);
 * NULL reproduces the reference (every item has L tokens). */
int frt2_decode(frt2_handle* h, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                int nq, int L, const int32_t* lengths, float* audio, int64_t audio_pitch, void* cuda_stream);

/* Same decode, but the waveform is emitted directly as the int16 PCM of the reference's wire format,
 * pcm = (int16) trunc(sample * 32767) ((audio * 32767).astype(np.int16), enhanced_fireredtts2.py:603,655), saturated
 * instead of wrapped for out-of-range samples: half the device->host bytes of the fp32 waveform. */
int frt2_decode_pcm16(frt2_handle* h, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                      int nq, int L, const int32_t* lengths, int16_t* pcm, int64_t pcm_pitch, void* cuda_stream);

/* Offline decode that SCATTERS its items: item b's 8*hop*L_b samples (L_b = lengths[b], or L when lengths is NULL) go
 * to out_base + out_off[b] (element offsets, device int64 (B)); nothing is written beyond an item's own samples.  This
 * is how the turns of a dialogue land directly at their place in the concatenated waveform (the reference
 * concatenates the decoded turns on the time axis afterwards, fireredtts2.py:399-401) — and, with out_base a buffer of
 * ANOTHER GPU opened through frt2_peer_open, how a rank's overlap-add kernel delivers its waveforms to the gathering
 * rank over NVLink without a separate collective.  out_pcm16 != 0: out_base is int16 PCM (see frt2_decode_pcm16). */
int frt2_decode_scatter(frt2_handle* h, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                        int nq, int L, const int32_t* lengths, void* out_base, int out_pcm16, const int64_t* out_off,
                        void* cuda_stream);

/* ---- peer memory for the waveform gather ("NCCL over NVLink only to gather waveform chunks") ----
 * The gathering rank allocates the destination buffer with frt2_peer_alloc and publishes the 64-byte handle (through
 * torch.distributed, a file, ... — plain bytes); every other rank (one process per GPU, same box) maps it with
 * frt2_peer_open and passes the mapped pointer as `audio` / `out_base` of its decode calls, so the last kernel of the
 * path writes over NVLink / NVSwitch peer stores while the rest of the batch is still computing.  The writer
 * synchronises its stream, then the ranks meet at a barrier; after that the owner may read the buffer.  All four
 * calls are synchronous. */
#define FRT2_PEER_HANDLE_BYTES 64
int frt2_peer_alloc(int device, int64_t bytes, void** ptr, unsigned char* handle /* FRT2_PEER_HANDLE_BYTES out */);
int frt2_peer_open(int device, const unsigned char* handle, void** ptr);
int frt2_peer_close(int device, void* ptr);
int frt2_peer_free(int device, void* ptr);

/* ---- streaming decode: RedCodecInfer.decode_one_token (model.py:326-376) ----
 * The stream object owns what the reference keeps in cache_dict (up_conv_cache, bb_conv_cache1/2, bb_kv_cache,
 * is_cache) in HBM, updated in place. */
int frt2_stream_create(frt2_handle* h, int B, int max_tokens, frt2_stream** out);
/* Back to "no token consumed" — the reference's fresh cache_dict = {} of a new utterance (model.py:346-355 take the
 * "cache is None" branches).  Stream-ordered: nothing runs here; the state is cleared by a kernel on the CUDA stream
 * of the next decode call, after everything the state's previous user enqueued (no device synchronisation). */
int frt2_stream_reset(frt2_stream* s);
/* The reference's call pattern is decode_one_token(token, {}, last) with a fresh cache_dict per utterance
 * (model.py:346): create / destroy per utterance.  frt2_stream_destroy therefore hands the state (device buffers and the
 * captured per-token step) back to a small per-handle pool and frt2_stream_create of the same (B, max_tokens) takes it
 * from there — no allocation, no capture, no synchronisation in the request path.  frt2_stream_reserve fills that pool
 * ahead of the first request (`count` spare states, per-token step captured). */
void frt2_stream_destroy(frt2_stream* s);
int frt2_stream_reserve(frt2_handle* h, int B, int max_tokens, int count);
int frt2_stream_tokens(const frt2_stream* s); /* tokens consumed so far */
/* Out-of-range codes are detected on the device, per stream and per batch item / pool slot (the reference raises
 * IndexError inside the offending decode_one_token, rvq.py:58).  frt2_stream_check_error synchronises cuda_stream, reads
 * and clears THIS stream's (or pool's) error words and returns FRT2_ERR_INDEX_OUT_OF_RANGE if any item sent a bad code
 * since the last check; item_flags (host, B entries, optional) receives a non-zero value for each offending item.  A bad
 * code is decoded as code 0.  frt2_stream_fetch_errors enqueues an asynchronous copy of the 1 + B words (word 0: any
 * item; word 1 + b: item b) to host_words (pinned host memory) on cuda_stream and neither synchronises nor clears: a
 * streaming front copies them alongside each chunk and tests word 0 once the chunk's event has completed. */
int frt2_stream_check_error(frt2_handle* h, frt2_stream* s, int32_t* item_flags, void* cuda_stream);
int frt2_stream_fetch_errors(frt2_handle* h, frt2_stream* s, uint32_t* host_words, void* cuda_stream);
/* One chunk of Lc >= 1 tokens per item.  n_samples (host, written before return) =
 * 8*hop*Lc - pad*[first chunk] + pad*[last], pad = (n_fft-hop)/2, exactly as ISTFT.forward_chunk slices
 * (decoder.py:459-467).  Attention inside the chunk is unmasked like the reference's forward_chunk
 * (whisper.py:107-113); with Lc == 1 it equals the offline block-causal mask. */
int frt2_decode_chunk(frt2_handle* h, frt2_stream* s, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ,
                      int64_t sL, int nq, int Lc, int last, float* audio, int64_t audio_pitch, int* n_samples,
                      void* cuda_stream);
/* Same chunk, emitted as the int16 PCM of the reference's wire format (see frt2_decode_pcm16): what the streaming
 * front sends per token (enhanced_fireredtts2.py:603,655). */
int frt2_decode_chunk_pcm16(frt2_handle* h, frt2_stream* s, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ,
                            int64_t sL, int nq, int Lc, int last, int16_t* pcm, int64_t pcm_pitch, int* n_samples,
                            void* cuda_stream);

/* ---- slot pool: many concurrent streams, one token each per step (continuous batching) ----
 * The reference serves concurrent requests one after the other (one daemon worker, enhanced_fireredtts2.py:199-203;
 * its asyncio BatchProcessor, performance_optimization.py:822, batches whole utterances after the fact).  A pool is
 * `slots` independent decode_one_token states (model.py:326-376) in HBM that advance together: every step decodes ONE
 * token for each active slot in a single batched launch sequence (one captured CUDA graph), each slot at its own
 * position of its own stream.  Per slot and step the caller passes FRT2_SLOT_* flags (HOST int32 array, `slots`
 * entries): ACTIVE = the slot has a token this step; RESET = the slot starts a new stream with this token (state
 * rewound first); LAST = this is the stream's final token.  tokens: device (slots, nq) integers, element strides
 * sB, sQ (idle slots are not read).  out: device (slots, out_pitch) fp32, or int16 PCM when out_pcm16 != 0; slot b
 * receives n_samples[b] (HOST, written before return) = 8*hop - pad*[first token] + pad*[LAST] samples (0 when
 * idle), zero-filled up to 8*hop + pad, so out_pitch >= 8*hop + pad.  A slot's samples are bit-identical whatever
 * the other slots are doing.  After LAST a slot must be RESET before it is stepped again. */
enum { FRT2_SLOT_ACTIVE = 1, FRT2_SLOT_LAST = 2, FRT2_SLOT_RESET = 4 };
#define FRT2_POOL_MAX_SLOTS 256
int frt2_pool_create(frt2_handle* h, int slots, int max_tokens, frt2_stream** out);
int frt2_pool_step(frt2_handle* h, frt2_stream* pool, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ,
                   int nq, const int32_t* slot_flags, void* out, int out_pcm16, int64_t out_pitch,
                   int32_t* n_samples, void* cuda_stream);
/* tokens consumed so far by one slot of a pool */
int frt2_pool_slot_tokens(const frt2_stream* pool, int slot);

/* Export / import the state in the reference's cache_dict layouts (model.py:346-375; the five tensors are produced at
 * decoder.py:624-655 [up_conv_cache], :275-320 [bb_conv_cache1/2, bb_kv_cache], :407-468 [is_cache]) — fp32, contiguous, device:
 * up_conv_cache (B,E,3), bb_conv_cache1 (B,E,6), bb_conv_cache2 (B,8E,2), bb_kv_cache (B,layers,H,T,2*hd)
 * with T = 8*frt2_stream_tokens(s), is_cache (B,n_fft,3).  Any pointer may be NULL (skipped). */
int frt2_export_state(frt2_handle* h, const frt2_stream* s, float* up_conv_cache, float* bb_conv_cache1,
                      float* bb_conv_cache2, float* bb_kv_cache, float* is_cache, void* cuda_stream);
int frt2_import_state(frt2_handle* h, frt2_stream* s, int n_tokens, const float* up_conv_cache,
                      const float* bb_conv_cache1, const float* bb_conv_cache2, const float* bb_kv_cache,
                      const float* is_cache, void* cuda_stream);

/* ---- RVQ encode (first stage): ResidualVQ.encode_codes (rvq.py:128-143) ----
 * The producer of the token tensors this library decodes: input_proj, then per quantizer in_project -> nearest
 * codebook row (VectorQuantize.encode_code, rvq.py:62-89: argmax of -(|z_e|^2 - 2 z_e.C^T + |C|^2), first maximum) ->
 * residual -= out_project(z_e + (C[idx] - z_e)).  Needs the encode-side tensors of the checkpoint
 * (rvq.quantizers.{i}.in_project.*, rvq.input_proj.*) to have been passed to frt2_load_tensor; FRT2_ERR_MISSING_TENSOR
 * otherwise.  z: device fp32 (B, input_dim, T) with ELEMENT strides sB, sD, sT (the reference passes the channel-major
 * (B, D, T) view, model.py:240; a time-major producer passes sD = 1).  codes: device int64 (nq, B, T) contiguous — the
 * reference's return value (rvq.py:142).  One kernel, fp32 on the CUDA cores, distances evaluated in the reference's
 * rounding order: an index differs from the reference's only where two codes tie within fp32 rounding. */
int frt2_rvq_encode(frt2_handle* h, const float* z, int64_t sB, int64_t sD, int64_t sT, int B, int input_dim, int T,
                    int nq, int64_t* codes, void* cuda_stream);

/* ---- codec ENCODE side behind the feature encoders ----
 * Everything RedCodecInfer._encode_one_batch (model.py:218-236) runs between the two Whisper encoders and
 * ResidualVQ.encode_codes: SslAdaptor (model.py:19-77: Linear, N x WhisperEncoderLayer with full attention, LayerNorm,
 * Linear), torch.cat([sem, aco], dim=2) (model.py:230) and ResidualDownConv (model.py:80-121).  The config mirrors
 * config_codec.json["codec"]["ssl_adaptor"|"acoustic_encoder"|"downsample"]; downsample.embed_dim must equal
 * ssl_out_dim + aco_dim and is what rvq.input_dim sees. */
typedef struct frt2_enc_config {
  int32_t ssl_in_dim;     /* ssl_adaptor.in_dim = the SSL encoder's width (1280, whisper.py:363) */
  int32_t ssl_embed_dim;  /* ssl_adaptor.embed_dim */
  int32_t ssl_out_dim;    /* ssl_adaptor.out_dim */
  int32_t ssl_num_layers; /* ssl_adaptor.num_layers */
  int32_t ssl_num_heads;  /* ssl_adaptor.num_heads (head_dim 32, 64 or 128; 64 / 128 run on tcgen05) */
  int32_t ssl_ffn_dim;    /* ssl_adaptor.ffn_dim, 0 = 4 * embed_dim (whisper.py:137) */
  int32_t aco_dim;        /* acoustic_encoder.embed_dim */
  int32_t avg_pooler;     /* downsample.avg_pooler (model.py:84) */
  /* the two feature encoders (optional: all of the following 0 = only frt2_enc_features is available) */
  int32_t ssl_enc_layers;  /* PretrainedWhisperEncoder: 32 layers, 20 heads, ffn 5120 at width ssl_in_dim (whisper.py:359-369) */
  int32_t ssl_enc_heads;
  int32_t ssl_enc_ffn_dim; /* 0 = 4 * width */
  int32_t aco_layers;      /* acoustic_encoder.num_layers / num_heads / ffn_dim (whisper.py:398-401) */
  int32_t aco_heads;       /* any head_dim <= 128 that is a multiple of 8 (zero-padded to 64 / 128 at load) */
  int32_t aco_ffn_dim;
  int32_t num_mels;        /* 128 (whisper.py:372,391); n_fft 400, hop 160, 16 kHz, 0..8000 Hz are fixed (whisper.py:373-377) */
  int32_t max_positions;   /* 1500 (whisper.py:366,403) */
} frt2_enc_config;
typedef struct frt2_encoder frt2_encoder;
int frt2_enc_create(const frt2_enc_config* cfg, int device, frt2_encoder** out);
/* Reference state_dict keys "ssl_adaptor.*", "downsample.*" and — when the feature encoders are configured — "ssl.*",
 * "acoustic_encoder.*" (fp32, reference layouts); other keys are ignored. */
int frt2_enc_load_tensor(frt2_encoder* e, const char* key, const float* data, int ndim, const int64_t* shape,
                         int on_device);
int frt2_enc_finalize(frt2_encoder* e);
void frt2_enc_destroy(frt2_encoder* e);
/* ssl (B,T,ssl_in_dim), aco (B,T,aco_dim): device fp32, contiguous, the outputs of the two feature encoders at 50 Hz
 * (every item T frames long: the reference encodes fixed 6 s chunks, model.py:262-275); T a multiple of avg_pooler.
 * vq_in: device fp32 (B, T/avg_pooler, ssl_out_dim + aco_dim) = the input of ResidualVQ.encode_codes in time-major
 * layout (hand it to frt2_rvq_encode with sD = 1, sT = dim).  *launches (host, optional) = kernels launched. */
int frt2_enc_features(frt2_encoder* e, const float* ssl, const float* aco, int B, int T, float* vq_in,
                      int64_t* launches, void* cuda_stream);

/* The whole of RedCodecInfer._encode_one_batch (model.py:218-236) up to the RVQ input, from the waveform: log-mel
 * front end (WhisperMelExtractor, whisper.py:261-302), the SSL and the acoustic WhisperEncoder (whisper.py:195-258),
 * then everything frt2_enc_features does.  audio16k: device fp32 (B, n) with row pitch audio_pitch, n a multiple of 1280
 * (the reference pads every chunk to 6 s = 96000 samples, model.py:262-275).  vq_in: (B, n/1280, dim).  Optional
 * parity taps (device fp32, may be NULL): mel_out (B, n/160, num_mels), ssl_out (B, n/320, ssl_in_dim), aco_out
 * (B, n/320, aco_dim) = the outputs of the feature extractor and of the two encoders. */
int frt2_enc_audio_features(frt2_encoder* e, const float* audio16k, int64_t audio_pitch, int B, int64_t n, float* vq_in,
                            float* mel_out, float* ssl_out, float* aco_out, int64_t* launches, void* cuda_stream);

/* ---- frame tail of the speech LM: replaces the part of Model.generate_frame behind the backbone ----
 * fireredtts2/llm/llm.py:304-330: codebook0_head + sample_topk (llm.py:305-306), the per-frame reset of the decoder's
 * K/V state (llm.py:317), and for i = 1 .. ncb-1 decoder(projection(curr_h)) -> audio_head[i-1] -> sample_topk(., 10,
 * 0.75) -> _embed_audio(i, .) (llm.py:318-328).  `decoder` = torchtune qwen2 (llm/modules.py:5-82).  One CUDA graph per
 * frame; the backbone (llm.py:292-302) stays with the caller. */
typedef struct {
  int32_t backbone_dim;        /* FLAVORS[backbone_flavor] embed_dim = width of last_h / audio_embeddings (llm.py:98-104) */
  int32_t dim;                 /* decoder flavor: embed_dim, num_layers, num_heads, num_kv_heads, intermediate_dim */
  int32_t num_layers;
  int32_t num_heads;
  int32_t num_kv_heads;
  int32_t intermediate_dim;
  int32_t audio_vocab_size;    /* ModelArgs.audio_vocab_size / audio_num_codebooks (llm.py:80-81) */
  int32_t audio_num_codebooks;
  float rope_base;             /* 1e6 (modules.py:16) */
  float norm_eps;              /* 1e-6 (modules.py:15) */
  int32_t max_batch;           /* 0 / <= 16: frames of up to 16 items (weight-streaming kernels only).  > 16 (<= 1024): also keeps
                                  row-major weight copies and runs larger batches — a pool of concurrent streams — on the
                                  tcgen05 GEMM (widths must be multiples of 64) */
} frt2_fd_config;
typedef struct frt2_frame_decoder frt2_frame_decoder;
int frt2_fd_create(const frt2_fd_config* cfg, int device, frt2_frame_decoder** out);
/* Model.state_dict() keys (fp32, reference layouts): "projection.weight" (dim, backbone_dim), "audio_embeddings.weight"
 * (ncb*V, backbone_dim), "codebook0_head.weight" (V, backbone_dim), "audio_head" (ncb-1, dim, V), "decoder.norm.scale",
 * "decoder.layers.{i}.{sa_norm.scale, mlp_norm.scale, attn.{q,k,v}_proj.{weight,bias}, attn.output_proj.weight,
 * mlp.{w1,w2,w3}.weight}" (torchtune names: w1 = gate, w3 = up, w2 = down); other keys are ignored. */
int frt2_fd_load_tensor(frt2_frame_decoder* f, const char* key, const float* data, int ndim, const int64_t* shape,
                        int on_device);
int frt2_fd_finalize(frt2_frame_decoder* f);
void frt2_fd_destroy(frt2_frame_decoder* f);
/* One frame for B <= max(16, max_batch) items.  last_h: device fp32 (B, backbone_dim) = h[:, -1, :] of the backbone (llm.py:304).
 * c0: optional device int32 (B) codebook-0 codes sampled by the caller (NULL: sampled here with topk / temperature).
 * noise: optional device fp32 (B, ncb, V), the Exp(1) draws q of _multinomial_sample_one_no_sync (llm.py:34-36) per
 * codebook (parity tests feed the reference's own draws); NULL: counter-based Philox draws keyed by (seed, frame counter
 * of the handle, item, codebook, entry).  forced: optional device int32 (B, ncb): teacher forcing (every code given; the
 * logits are those of the forced history).  codes: device int32 (B, ncb) — the curr_sample of llm.py:330 (the codec's
 * token operand: frt2_decode_chunk(tokens = codes, idx_bytes 4, sB = ncb, sQ = 1, sL = 1)).  logits: optional device fp32
 * (B, ncb, V): c0_logits at [:, 0] (only when c0 == NULL), ci_logits at [:, i].  Asynchronous on the stream; a given code
 * outside [0, V) is reported by frt2_fd_check_error (the reference raises IndexError in nn.Embedding). */
int frt2_fd_generate(frt2_frame_decoder* f, const float* last_h, int B, const int32_t* c0, const float* noise,
                     uint64_t seed, int topk, float temperature, const int32_t* forced, int32_t* codes, float* logits,
                     int64_t* launches, void* cuda_stream);
int frt2_fd_check_error(frt2_frame_decoder* f, void* cuda_stream);

/* ---- waveform resampler of the context loop ----
 * Replaces torchaudio.functional.resample(waveform, orig_freq, new_freq) with its defaults (sinc_interp_hann,
 * lowpass_filter_width 6, rolloff 0.99) as the reference calls it on every generated turn (24 kHz -> 16 kHz,
 * fireredtts2.py:389-391) and on the prompt (fireredtts2.py:65).  in: device fp32 (B, n_in) with row pitch in_pitch;
 * out: device fp32 (B, ceil(new*n_in/orig)) with row pitch out_pitch; lengths: optional device int32 (B) of per-item
 * sample counts <= n_in (ragged batch: item b is resampled as if it were lengths[b] long, the rest of its row is 0).
 * *n_out (host, optional) receives ceil(new*n_in/orig).  Needs no handle (no weights): the FIR bank of a rate pair
 * is built once per device and cached. */
int frt2_resample(int device, const float* in, int64_t in_pitch, int B, int64_t n_in, const int32_t* lengths,
                  int orig_freq, int new_freq, float* out, int64_t out_pitch, int64_t* n_out, void* cuda_stream);

/* frt2_decode with the resampler fused behind the overlap-add (the context loop's decode -> torchaudio resample,
 * fireredtts2.py:386-391): ONE kernel computes the 24 kHz samples from the iSTFT frames, stores them to audio (optional:
 * NULL = only the resampled waveform is wanted) and resamples them from shared memory into audio_rs (B, ceil(new * 1920 L
 * / orig)), row pitch rs_pitch; *n_rs (host, optional) receives that length.  Both outputs are bit-identical to
 * frt2_decode followed by frt2_resample (lengths: item b's outputs behind its own length are zero).  orig_freq is the
 * codec's 24000; new_freq / gcd must be 1, 2 or 3 (16 kHz: 2). */
int frt2_decode_resampled(frt2_handle* h, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                          int nq, int L, const int32_t* lengths, float* audio, int64_t audio_pitch, int orig_freq,
                          int new_freq, float* audio_rs, int64_t rs_pitch, int64_t* n_rs, void* cuda_stream);

/* ---- parity hooks ----
 * Raw codebook rows and their index-ordered fp32 sum, bit-exact with VectorQuantize.decode_code /
 * ResidualVQ.decode_codes for Identity projections (rvq.py:56-60,145-164).  rows (B,L,nq,cd) / sum (B,L,cd),
 * either may be NULL. */
int frt2_rvq_gather(frt2_handle* h, const void* tokens, int idx_bytes, int64_t sB, int64_t sQ, int64_t sL, int B,
                    int nq, int L, float* rows, float* sum, void* cuda_stream);
/* Copy an intermediate of the most recent frt2_decode into out (device fp32, time-major (B,T,C)) — the values the
 * reference holds inside ResidualVQ.decode_codes (rvq.py:145-164: "emb" before, "z" after output_proj), after UpConv
 * (model.py:142-148: "x50"), after the upsample convs (decoder.py:610-616: "up"), inside CausalVocosBackbone.forward
 * (decoder.py:258-273: "prior" after prior_net :260, "layer0" / "layers" after the first / last layer :267-268, "final"
 * after final_norm :273), and in the iSTFT head (decoder.py:503-518: "spec"; :380-381: "frames" = windowed irfft):
 * "emb" (B,L,rvq_dim), "z" (B,L,E), "x50" (B,4L,E), "up" (B,8L,E), "prior", "layer0", "layers", "final",
 * "spec" (B,8L,2*n_bins interleaved re,im), "frames" (B,8L,n_fft).  Requires frt2_set_debug(h, 1) before
 * the decode.  Returns the number of floats written through *n. */
int frt2_set_debug(frt2_handle* h, int flags);
int frt2_get_tap(frt2_handle* h, const char* name, float* out, int64_t capacity, int64_t* n, void* cuda_stream);
/* Synchronise the stream and translate the handle's device-side error word (out-of-range code index of an OFFLINE decode —
 * the reference raises IndexError inside F.embedding, rvq.py:58 — /
 * frt2_rvq_gather since the last check) into a status.  Streams and pools have their own words: frt2_stream_check_error. */
int frt2_check_error(frt2_handle* h, void* cuda_stream);

/* ---- measurement ----
 * frt2_profile(h, 1) brackets every kernel launch of the following decode calls with CUDA events on the
 * caller's stream and resets the counters; frt2_profile_get sums one kernel class: total device ms, number of
 * launches, algorithmic FLOPs and algorithmic bytes (DESIGN.md states the per-unit figures).  Class
 * FRT2_PROF_ALL returns the total number of kernels launched since frt2_profile(h, *) in *launches. */
enum { FRT2_PROF_GEMM = 0, FRT2_PROF_ATTN_TC = 1, FRT2_PROF_ATTN_WARP = 2, FRT2_PROF_LAYER_NORM = 3,
       FRT2_PROF_RVQ = 4, FRT2_PROF_OLA = 5, FRT2_PROF_GEMM_SKINNY = 6, FRT2_PROF_ALL = -1 };
int frt2_profile(frt2_handle* h, int enable);
int frt2_profile_get(frt2_handle* h, int cls, double* ms, int64_t* launches, double* flops, double* bytes);

/* ---- single-operator entry points (unit parity tests and per-kernel roofline benches) ----
 * Each is one kernel of the path on its own, checked against the torch fp32 op the reference calls there. */
/* The nn.Linear / Conv1d / ConvTranspose1d contractions of the path (model.py:142-148, decoder.py:78-101,571-589,
 * whisper.py:37-40,137-138, decoder.py:503): C[M,N] = act(alpha * A[M,K] * W[N,K]^T + bias) (+ resid); A,W fp16 device, fp32 accumulate.
 * impl 0 = tcgen05/TMEM/TMA kernel, 1 = SIMT check kernel, 2 = skinny weight-streaming kernel (<= 16 rows), 3 = the frame
 * tail's persistent weight-streaming kernel (one batch of <= 16 rows that fit its shared-memory tile, K a multiple of 32; act 3 = SwiGLU on interleaved
 * (gate, up) weight rows: out16 is (rows, N/2); the weights are repacked inside the call, which synchronises).  ntaps > 1: causal conv over `batches` items of
 * rows_per_batch rows, K = ntaps*Kc, zero left padding. */
int frt2_op_gemm(int impl, const void* A16, const void* W16, int batches, int rows_per_batch, int Kc, int ntaps,
                 int N, float alpha, const float* bias, int act, const float* resid, float* out32, void* out16,
                 void* cuda_stream);
/* nn.LayerNorm over channels (+ the nn.SiLU behind it in CausalResnetBlock): decoder.py:119-121,127-129,246; whisper.py:134,140. */
int frt2_op_layer_norm(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                       int apply_silu, void* out16, void* cuda_stream);
/* F.scaled_dot_product_attention of WhisperSdpaAttention (whisper.py:49-79 with the block-causal mask of utils.py:19-38:
 * key j visible to query i iff j <= (i | 7); whisper.py:81-118 chunked: Tq new queries at q_pos0 over Tk keys, no mask).
 * q,k,v,out: fp16 (B,T,H*hd) contiguous.  impl 0 = tcgen05 flash kernel, 1 = warp kernel. */
int frt2_op_attention(int impl, const void* q16, const void* k16, const void* v16, void* out16, int B, int H,
                      int hd, int Tq, int Tk, int q_pos0, int block_causal, void* cuda_stream);
/* debug: while dev_buf (16*128*8 uint32, device) is set, CTA 0 of the persistent tcgen05 attention kernel writes
 * clock64 stamps of its softmax / MMA-issue phases there (tools/attn_trace.py); NULL switches it off. */
int frt2_op_attention_trace(void* dev_buf);
/* sample_topk + _multinomial_sample_one_no_sync (llm.py:34-49): logits (B, V) device fp32, noise (B, V) Exp(1) draws or NULL
 * (Philox draws keyed by seed, item, entry), codes (B) device int32.  Synchronises the stream. */
int frt2_op_sample_topk(const float* logits, int B, int V, int topk, float temperature, const float* noise, uint64_t seed,
                        int32_t* codes, void* cuda_stream);
/* ISTFT's window / fold / envelope division / trim (decoder.py:380-405 offline; :407-468 chunked with the three cached
 * windowed frames `tail` and the first / last trimming rules). */
int frt2_op_overlap_add(const float* frames, const float* tail, const float* window, const int32_t* lengths,
                        float* audio, int64_t audio_pitch, int B, int T, int n_fft, int hop, int first, int last,
                        void* cuda_stream);

const char* frt2_last_error(void);
const char* frt2_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FRT2_H_ */
