/*
 * b200vgan -- C ABI of the B200-native BigVGAN2 speech-code decoder (libb200vgan.so).
 *
 * This is the drop-in boundary for ONE path of scwf/index-tts-dubbing: the
 * `indextts/BigVGAN` generator forward (reference: indextts/BigVGAN/models.py:201-250,
 * called from indextts/infer.py:458 and :623 as `wav, _ = self.bigvgan(latent, mel_ref)`),
 * and for the reference's only native component, the fused anti-alias activation extension
 * (indextts/BigVGAN/alias_free_activation/cuda/anti_alias_activation.cpp:19-23,
 * anti_alias_activation_cuda.cu:214-256).
 *
 * Conventions
 *   - plain C: raw device pointers, sizes and a CUDA stream handle (void*, a cudaStream_t);
 *     no torch / pybind types cross this boundary.
 *   - every function returns 0 on success, non-zero on failure; the message is available from
 *     bvg_last_error() (thread-local).  Nothing throws across the ABI.
 *   - the caller owns all tensors and the workspace; the library owns only its repacked
 *     weights and the arena of per-geometry segment tables (inside bvg_handle; a bvg_plan holds
 *     a slice of it, so destroy plans before their handle).
 *   - all work is asynchronous on the given stream; bvg_forward performs no allocation and no
 *     host synchronisation, so it is CUDA-graph capturable.
 *   - a handle is bound to the device current at bvg_create(); it is not thread-safe.
 *   - there is no CPU fallback: every entry point fails with an error if no sm_100 GPU is
 *     present.
 */
#ifndef B200VGAN_H_
#define B200VGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVG_MAX_UPS 8
#define BVG_MAX_KERNELS 4
#define BVG_MAX_DILATIONS 4

/* element types of caller tensors */
enum { BVG_F32 = 0, BVG_BF16 = 1, BVG_F16 = 2 };

/* arithmetic mode of bvg_forward */
enum {
  BVG_MODE_FP32 = 0, /* parity mode: fp32 storage, fp32 FFMA convolutions (CUDA cores)          */
  BVG_MODE_BF16 = 1, /* performance mode: bf16 storage, tcgen05 bf16 MMA with fp32 accumulation  */
  BVG_MODE_F16 = 2,  /* performance mode: fp16 storage (3 more mantissa bits, saturating stores), tcgen05 f16 MMA with
                        fp32 accumulation -- what the reference runs under torch.amp.autocast(float16), infer.py:456,613 */
  BVG_MODE_FP32_TC = 3 /* fp32 on the tensor cores: fp32 storage, bias, residuals and activations; every convolution as three
                        tcgen05 f16 products x_lo w_hi + x_hi w_lo + x_hi w_hi (x and w split into two fp16 numbers each: 22
                        significand bits, every product exact in the fp32 accumulator).  Waveform 8.8e-5 from the reference's
                        fp32 output on config 1 (gate 1e-3; BVG_MODE_FP32: 1.1e-5) at 6x the speed of BVG_MODE_FP32.  Inputs of a
                        convolution saturate at +-65504 (as in BVG_MODE_F16). */
};

/* memory layout of the per-op test entry points */
enum { BVG_LAYOUT_NCT = 0 /* [B,C,T] like the reference */ };

/* Architecture = the `bigvgan:` section of the reference's checkpoints/config.yaml:51-70
 * (only the keys the inference path reads, models.py:132-197). */
typedef struct bvg_config {
  int32_t gpt_dim;                   /* 1024  latent width                         */
  int32_t upsample_initial_channel;  /* 1536                                       */
  int32_t num_upsamples;             /* 6                                          */
  int32_t upsample_rates[BVG_MAX_UPS];        /* 4,4,4,4,2,2                       */
  int32_t upsample_kernel_sizes[BVG_MAX_UPS]; /* 8,8,4,4,4,4                       */
  int32_t num_kernels;               /* 3     resblocks per stage                  */
  int32_t resblock_kernel_sizes[BVG_MAX_KERNELS];                   /* 3,7,11      */
  int32_t resblock_dilation_sizes[BVG_MAX_KERNELS][BVG_MAX_DILATIONS]; /* 1,3,5 x3 */
  int32_t num_dilations;             /* 3                                          */
  int32_t speaker_embedding_dim;     /* 512                                        */
  int32_t cond_in_each_up_layer;     /* 1                                          */
  int32_t num_mels;                  /* 100   input width of the speaker encoder (0 = 100) */
} bvg_config;

typedef struct bvg_handle bvg_handle;
typedef struct bvg_plan bvg_plan;

const char* bvg_last_error(void);
int bvg_version(void);

/* 0 if a usable sm_100 device is current, else an error (no fallback exists). */
int bvg_device_check(void);

/* ---- generator lifetime ------------------------------------------------------------------
 * Replaces: Generator(h, use_cuda_kernel) + load_state_dict + remove_weight_norm
 * (indextts/infer.py:110-117, models.py:132-197, :252-260). */
int bvg_create(const bvg_config* cfg, bvg_handle** out);
void bvg_destroy(bvg_handle* h);

/* Upload one FOLDED (weight-norm removed) fp32 parameter by its reference state-dict name, e.g.
 * "conv_pre.weight" [1536,1024,7], "ups.0.0.weight" [1536,768,8] (ConvTranspose1d layout
 * [Cin,Cout,k]), "resblocks.4.convs1.2.bias", "resblocks.0.activations.3.act.alpha" (log scale),
 * "cond_layer.weight", "conds.2.bias", "conv_post.weight".  `data` may be a host or a device
 * pointer (is_device).  Unknown names are an error; "*.filter" buffers are not accepted (the filter
 * taps are architecture constants, models.py never trains them).  "speaker_encoder.*" parameters
 * are optional as a group: upload all of them to use bvg_speaker_embedding, or none. */
int bvg_set_weight(bvg_handle* h, const char* name, const float* data, const int64_t* shape,
                   int32_t ndim, int32_t is_device, void* stream);

/* Repack weights for the kernels (fp32 tap-major for the CUDA-core path, bf16 UMMA shared-memory
 * images for the tcgen05 path), precompute exp(alpha), 1/(exp(beta)+1e-9).  Fails if any
 * parameter is missing.  Synchronises the stream. */
int bvg_finalize(bvg_handle* h, void* stream);

/* ---- per-geometry plan ------------------------------------------------------------------
 * A plan fixes the batch geometry: B segments with frames[b] latent frames each (segment b
 * produces frames[b]*prod(upsample_rates) samples).  Segments are packed along time with
 * zero guard gaps in the workspace, so variable-length batches cost no padding FLOPs. */
int bvg_plan_create(bvg_handle* h, int32_t B, const int32_t* frames, int32_t mode, bvg_plan** out);
void bvg_plan_destroy(bvg_plan* p);
size_t bvg_plan_workspace_bytes(const bvg_plan* p);
int32_t bvg_plan_max_frames(const bvg_plan* p);
/* number of kernel launches one bvg_forward with this plan issues (for bench.py's gpu_launches) */
int32_t bvg_plan_num_launches(const bvg_plan* p);

/* ---- the hot path -------------------------------------------------------------------------
 * Replaces BigVGAN.forward after the speaker encoder (models.py:210-250):
 *   latent   [B, max_frames, gpt_dim]      latent_dtype (BVG_F32 | BVG_BF16 | BVG_F16), row-major
 *   spk_emb  [spk_batch, 1, spk_dim] fp32  ECAPA embedding, spk_batch in {1, B}
 *   wav      [B, 1, max_frames*hop] fp32   tanh output; samples beyond a segment's length are 0
 *   workspace: >= bvg_plan_workspace_bytes(plan) bytes of device memory, 256-byte aligned; its contents on
 *   entry do not matter (every call re-clears the zero guard rows of the packed layout with one small launch),
 *   so one workspace can serve any sequence of plans and may be a recycled allocation. */
int bvg_forward(bvg_handle* h, bvg_plan* plan, const void* latent, int32_t latent_dtype,
                const float* spk_emb, int32_t spk_batch, float* wav, void* workspace,
                size_t workspace_bytes, void* stream);

/* Same decode, but the final kernel also emits 16-bit PCM exactly as the reference's callers derive it from
 * the waveform: clamp(32767 * wav, -32767, 32767) and a truncating cast (infer.py:462, :627, :650).
 *   pcm [B, max_frames*hop] int16 (samples beyond a segment's length are 0);  wav_or_null: also write the
 *   fp32 waveform (layout as in bvg_forward) when non-NULL.  Halves the device-to-host bytes of a dubbing job. */
int bvg_forward_pcm16(bvg_handle* h, bvg_plan* plan, const void* latent, int32_t latent_dtype,
                      const float* spk_emb, int32_t spk_batch, int16_t* pcm, float* wav_or_null,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Optional per-launch timing for roofline reports (bench.py): when enabled, bvg_forward brackets
 * every launch with a CUDA event pair on the caller's stream.  bvg_profile_read synchronises on
 * those events and returns, per kernel class (0 = standalone Activation1d, 1 = tcgen05 conv,
 * 2 = CUDA-core conv, 3 = other), the summed device milliseconds, algorithmic FLOPs, algorithmic
 * bytes and launch counts since the previous read.  All four arrays have 4 entries. */
int bvg_profile_enable(bvg_handle* h, int32_t on);
int bvg_profile_read(bvg_handle* h, double* ms, double* flops, double* bytes, int64_t* launches);

/* Ragged (variable-length) batch I/O for multi-utterance / SRT-dubbing jobs -- replaces the per-entry loop of
 * srt_dubbing/src/strategies/stretch_strategy.py:72-83 and the time-concatenation of indextts/infer.py:439-463:
 *   latent_rows [sum_b frames[b], gpt_dim]   the segments' latent frames back to back in plan order (no padding)
 *   wav_rows    [sum_b frames[b]*hop] fp32   and / or
 *   pcm_rows    [sum_b frames[b]*hop] int16  (clamp(32767*wav) as in bvg_forward_pcm16); either may be NULL, not both.
 * Segment b's samples start at hop * (frames[0] + ... + frames[b-1]); nothing else is written.  Each segment is decoded
 * exactly as if it were alone in the batch. */
int bvg_forward_ragged(bvg_handle* h, bvg_plan* plan, const void* latent_rows, int32_t latent_dtype,
                       const float* spk_emb, int32_t spk_batch, float* wav_rows_or_null, int16_t* pcm_rows_or_null,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Plan bookkeeping: plans created on this handle so far (their tables come from a recycled device + pinned-host arena,
 * so creating one costs no device allocation and no synchronous copy once the arena is warm), and the number of
 * latent frames (sum over segments) a plan decodes. */
int64_t bvg_plans_created(const bvg_handle* h);
int64_t bvg_plan_total_frames(const bvg_plan* plan);

/* Host-buffer variant used for end-to-end timing: copies latent (host, ideally pinned) to the
 * device, runs bvg_forward, copies wav back into `wav_host` and synchronises the stream.
 * Device staging buffers must be provided by the caller (latent_dev, wav_dev). */
int bvg_forward_host(bvg_handle* h, bvg_plan* plan, const void* latent_host, int32_t latent_dtype,
                     void* latent_dev, const float* spk_emb, int32_t spk_batch, float* wav_host,
                     float* wav_dev, void* workspace, size_t workspace_bytes, void* stream);

/* ---- speaker encoder ------------------------------------------------------------------------
 * Replaces `speaker_embedding = self.speaker_encoder(mel_ref, lens)` (models.py:202-208; ECAPA-TDNN,
 * ECAPA_TDNN.py:429-581, fp32).  Needs every "speaker_encoder.*" parameter of the checkpoint uploaded
 * with bvg_set_weight before bvg_finalize (the BatchNorm "num_batches_tracked" counters are not
 * parameters and are not accepted).
 *   mel       [B, Tm, num_mels] fp32 device, Tm >= 5
 *   rel_lens  [B] fp32 device (relative lengths in (0,1], the reference's `lens`) or NULL
 *   emb       [B, 1, speaker_embedding_dim] fp32 device -- feed it to bvg_forward as spk_emb
 *   workspace >= bvg_ecapa_workspace_bytes(h, B, Tm) bytes of device memory */
size_t bvg_ecapa_workspace_bytes(const bvg_handle* h, int32_t B, int32_t Tm);
int bvg_speaker_embedding(bvg_handle* h, const float* mel, int32_t B, int32_t Tm, const float* rel_lens,
                          float* emb, void* workspace, size_t workspace_bytes, void* stream);

/* ---- prompt front-end -------------------------------------------------------------------------
 * Replaces `cond_mel = MelSpectrogramFeatures()(audio)` (indextts/infer.py:513, utils/feature_extractors.py:24-50,
 * padding="center"): torchaudio MelSpectrogram(n_fft 1024, periodic Hann, power 1, centre reflect padding, HTK mel
 * scale, norm None) + safe_log = log(clip(x, 1e-7)) (utils/common.py:110-121), fp32.
 *   audio  [B, n_samples] fp32 device, n_samples > 512
 *   mel    fp32 device: [B, n_mels, frames] like the reference (transposed = 0) or [B, frames, n_mels], the layout
 *          bvg_speaker_embedding / the module call take (transposed = 1); frames = bvg_mel_frames(n_samples, hop)
 *   sample_rate 24000, hop 256, n_mels 100 (<= 128), f_min 0, f_max <= 0 meaning sample_rate / 2: the reference's values */
int bvg_mel_frames(int32_t n_samples, int32_t hop);
int bvg_log_mel(const float* audio, int32_t B, int32_t n_samples, int32_t sample_rate, int32_t hop, int32_t n_mels,
                float f_min, float f_max, float* mel, int32_t transposed, void* stream);

/* ---- dubbing timeline ---------------------------------------------------------------------------
 * The sum / normalise part of AudioProcessor._time_synchronized_merge (srt_dubbing/src/audio_processor.py:176-232):
 *   out[t] = sum over the segments i covering t, in ascending seg_rank[i], of flat[seg_src[i] + t - seg_dst[i]]   (fp32)
 *   normalize != 0:  out /= max|out|  when that peak exceeds max_amplitude (AUDIO.MAX_AMPLITUDE = 1.0)
 * flat: the decoded segments back to back (device fp32); seg_src / seg_dst (int64) / seg_n (int32): device arrays of
 * nseg entries, seg_dst ascending, seg_rank (int32) = each segment's position in the reference's adding order
 * (start-time order; it differs from the seg_dst order when the overlap rule moved a segment); max_n = max seg_n; at most 16 segments may
 * cover one sample; out [total] device fp32; peak_scratch: 4 bytes of device memory (needed when normalize != 0).
 * The placement arithmetic (sorting, overlap rule, array growth) is host logic: b200vgan/timeline.py. */
int bvg_timeline_merge(const float* flat, const int64_t* seg_src, const int64_t* seg_dst, const int32_t* seg_n,
                       const int32_t* seg_rank, int32_t nseg, int64_t max_n, float* out, int64_t total, int32_t normalize, float max_amplitude,
                       uint32_t* peak_scratch, void* stream);

/* ---- per-op entry points (drop-in for the reference's native extension, and test hooks) ----
 *
 * bvg_activation1d: replaces anti_alias_activation_cuda.forward(x, up_f, down_f, alpha, beta)
 * (anti_alias_activation.cpp:19-23).  x,y [B,C,T] contiguous, dtype in {F32,BF16,F16}; alpha,beta
 * fp32 [C] in LOG scale (the kernel applies exp, as the reference does in .cu:89-90).  Edge
 * semantics follow the reference's torch path (alias_free_torch/act.py:24-29: replicate padding
 * of the input AND of the activated 2x signal), not the reference kernel's. */
int bvg_activation1d(const void* x, void* y, const float* log_alpha, const float* log_beta,
                     int32_t B, int32_t C, int32_t T, int32_t dtype, void* stream);

/* The same function through the kernel bvg_forward uses (packed c8 layout, per-segment lengths):
 * x,y [B,C,T] fp32 are packed / unpacked around it; C multiple of 8; mode as in bvg_forward
 * (BVG_MODE_BF16 stores the packed tensors in bf16).  Test / tooling entry, allocates temporaries. */
int bvg_activation1d_packed(const float* x, float* y, const float* log_alpha, const float* log_beta,
                            int32_t B, int32_t C, int32_t T, int32_t mode, void* stream);

/* Dense Conv1d, "same" zero padding d*(k-1)/2, stride 1 (torch.nn.Conv1d as used at
 * models.py:26-41,149).  x [B,Cin,T] fp32, w [Cout,Cin,k] fp32, bias [Cout] or NULL,
 * residual [B,Cout,T] or NULL, y [B,Cout,T] fp32.  Cin, Cout multiples of 8.
 * mode selects the CUDA-core fp32 kernel or the tcgen05 bf16 kernel.
 * Allocates temporaries (test / tooling entry, not used by bvg_forward). */
int bvg_conv1d(const float* x, const float* w, const float* bias, const float* residual, float* y,
               int32_t B, int32_t Cin, int32_t Cout, int32_t T, int32_t k, int32_t dilation,
               int32_t mode, void* stream);

/* Activation1d followed by Conv1d (one half-step of AMPBlock1.forward, models.py:69-72): y = conv(act(x)) [+ residual].
 * *fused (in/out, may be NULL): in = non-zero requests the experimental kernel that computes the activation
 * inside the tcgen05 convolution's producer stage (BVG_MODE_BF16, layer must fit one CTA tile, Cout <= 256;
 * bvg_forward uses it only with BVG_FUSE_ACT=1); out = whether that kernel ran.  Otherwise the activation
 * runs as its own pass, as in bvg_forward.  Same shapes as bvg_conv1d; log_alpha/beta [Cin]. */
int bvg_act_conv1d(const float* x, const float* log_alpha, const float* log_beta, const float* w, const float* bias,
                   const float* residual, float* y, int32_t B, int32_t Cin, int32_t Cout, int32_t T, int32_t k,
                   int32_t dilation, int32_t mode, int32_t* fused, void* stream);

/* ConvTranspose1d with stride u, kernel k, padding (k-u)/2 (models.py:155-161).
 * x [B,Cin,T] fp32, w [Cin,Cout,k] fp32, y [B,Cout,T*u] fp32. */
int bvg_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int32_t B,
                         int32_t Cin, int32_t Cout, int32_t T, int32_t k, int32_t u, int32_t mode,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VGAN_H_ */
