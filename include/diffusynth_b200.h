/* diffusynth_b200 -- C ABI of the B200 (sm_100a) implementation of DiffuSynth's
 * text-to-timbre sampling path.
 *
 * The reference (WxuanYuan/diffusynth) is pure Python/PyTorch and has no FFI; the Python
 * classes in diffusynth_b200/ keep the reference's surfaces (DiffSynthSampler,
 * ConditionedUnet.forward, VectorQuantizerEMA.forward, Decoder/Encoder.forward,
 * decode_stft/istft ...) and route them to these entry points through ctypes
 * (INTEGRATION.md shows the binding).  Each entry cites the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes only; every pointer named d_* is DEVICE memory owned
 * by the caller; `stream` is a cudaStream_t passed as void*; all calls are asynchronous on
 * that stream and capturable into a CUDA graph; return value 0 = ok, <0 = error
 * (-1 invalid argument/shape, -2 CUDA error, -3 NCCL error, -4 unsupported);
 * ds_last_error() returns the message of the calling thread's last failure.
 * Activation tensors are 16-bit ("act16", see ds_operand_dtype) channels-last ("NHWC") unless a name says f32/nchw.
 */
#ifndef DIFFUSYNTH_B200_H
#define DIFFUSYNTH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* ds_last_error(void);
int ds_version(void);
/* 16-bit storage / tensor-core operand format the library was built with: 1 = IEEE fp16 (default),
   0 = bf16 (-DDS_OPERANDS_BF16).  "act16" below means this type. */
int ds_operand_dtype(void);
/* Device sanity: returns 0 when device `dev` is compute capability 10.x. */
int ds_check_device(int dev);

/* ----------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMEM accumulators, TMA-fed).
 * Replaces every dense nn.Conv2d / nn.ConvTranspose2d on the path:
 *   ConvNextBlock.net[1], net[4], res_conv      model/diffusion_components.py:122,125,128
 *   Downsample / Upsample (4x4 stride 2)        model/diffusion_components.py:32-39
 *   LinearCrossAttentionAdd.to_qkv / to_out     model/diffusion_components.py:263-264
 *   final_conv[1]                               model/diffusion.py:174
 *   VQGAN Encoder/Decoder convs                 model/VQGAN.py:149-174,191,216,254-259
 * and, folded into its epilogue, the GroupNorm(1,C) that precedes the conv
 * (diffusion_components.py:121,124,148), bias, GELU (:123), the residual add (:139),
 * and the (sum, sum-of-squares) partials the next GroupNorm needs.
 *
 * GEMM view: M = pixels of an (Hb x Wb = 128)-pixel spatial tile of one sample, N = output
 * channels (tile BN), K = taps x input channels.  Up to two channel-concatenated sources
 * (pad_and_concat, diffusion_components.py:236-249) are read in place.
 * -------------------------------------------------------------------------------------- */
#define DS_MAX_TAPS 16
#define DS_MAX_GROUPS 4

typedef struct ds_conv_tap {
  int8_t dy, dx;   /* pixel offset of this tap inside its view */
  int8_t view;     /* which (parity) view of the source the tap reads: 0..3 */
  int8_t pad_;
} ds_conv_tap;

typedef struct ds_conv_gemm_args {
  /* sources: act16 NHWC [N, Hs, Ws, C]; source 1 optional (C1 = 0), concatenated after source 0 along C.
     A "view" v of a source is a strided pixel window, given in PIXEL units so that it applies to both
     sources: pixel(n, y, x) = view_off[v] + n*view_sn + y*view_sh + x*view_sw, extent Hv x Wv; the
     element offset in a source with C channels is pixel*C + c.  Stride-1 convs use the identity view,
     the 4x4 stride-2 conv the four (row parity, column parity) views of its input. */
  const void* d_src0;
  const void* d_src1;
  int32_t C0, C1;
  int32_t N;                /* samples */
  int32_t src_batch_mod;    /* if >0, sample n reads sample n % src_batch_mod of SOURCE 0 (and of d_stats_in); source 1 is indexed by n.
                               Classifier-free guidance: tensors computed once for both halves of the doubled batch. */
  int32_t Hv, Wv;           /* extent of every view (pixels) */
  int64_t view_sn, view_sh, view_sw;   /* pixel strides of a view per unit of (n, y, x) */
  int64_t view_off[4];      /* pixel offset of each view's origin */
  int32_t num_views;
  int32_t view_wv[4], view_hv[4];      /* per-view extents where they differ (odd input sizes of the stride-2 conv: the even-parity
                                          view holds one more column/row than the odd one); 0 = Wv / Hv */
  /* GEMM pixel grid (per sample) and tile shape */
  int32_t H, W;             /* output-tile grid extent */
  int32_t Hb, Wb;           /* Hb*Wb == 128 */
  /* weights: act16 [Z, Cout_pad, K] K-major, K = ntaps*(C0+C1) ordered tap-major then channel
     (source 0 channels first).  Z = groups, or N when per_sample_weights. */
  const void* d_weight;
  int32_t Cout_pad;         /* multiple of BN */
  int32_t Cout;             /* real output channels (<= Cout_pad) */
  int32_t BN;               /* N tile: multiple of 16, <= 256 */
  int32_t BK;               /* 32 or 64; must divide C0 and C1 */
  int32_t ntaps, groups;
  int32_t per_sample_weights;
  ds_conv_tap taps[DS_MAX_GROUPS][DS_MAX_TAPS];
  /* epilogue:  v = rstd*acc - rstd*mean*e1[cls][o] + e2[cls][o] + sbias[n][o];  act;  + residual */
  const void* d_stats_in;   /* statistics buffer of the (GroupNorm(1,C)-normalised) source, or NULL: float2
                               [N][2 + stats_in_slots], entry [n][0] = (mean, rstd) published by the producer */
  int32_t stats_in_slots;
  float stats_out_inv_count; /* 1 / (elements per sample of THIS call's output), used when d_stats_out is set */
  float eps;                /* GroupNorm eps used when publishing (mean, rstd) of the output */
  const float* d_e1;        /* [ncls][Cout_pad] or NULL */
  const float* d_e2;        /* [ncls][Cout_pad]: bias (+ folded GroupNorm beta term) */
  int32_t ncls;             /* 1, or 9 = (top/mid/bottom) x (left/mid/right) border classes of a 3x3 */
  const float* d_sbias;     /* per-sample bias [N][sbias_stride] or NULL */
  int32_t sbias_stride;
  int32_t act;              /* 0 none, 1 GELU(erf) */
  const void* d_residual;   /* act16, pixel strides below, or NULL */
  int64_t res_sn, res_sh, res_sw;      /* in elements */
  void* d_out;              /* act16 out, or NULL */
  int64_t out_sn, out_sh, out_sw;      /* in elements */
  int64_t out_goff[DS_MAX_GROUPS];     /* element offset per group (sub-pixel phase of ConvTranspose) */
  float* d_out_f32_nchw;    /* optional fp32 [N, Cout, H, W] output (final conv) */
  void* d_stats_out;        /* statistics buffer of the output, float2 [N][2 + ds_conv_gemm_stats_slots()], zero-initialised
                               once by the caller (entry [n][1] is an arrival counter that resets itself), or NULL:
                               [n][2+i] = (sum, sumsq) partial of tile-warp i, [n][0] = (mean, rstd) written by the last arriver */
} ds_conv_gemm_args;

int ds_conv_gemm(const ds_conv_gemm_args* args, void* stream);
/* number of (sum, sumsq) partial slots per sample the call writes to d_stats_out (buffer holds 2 more entries) */
int ds_conv_gemm_stats_slots(const ds_conv_gemm_args* args);
/* Same contract on plain CUDA cores; exists only to cross-check the tcgen05 kernel in tests. */
int ds_conv_gemm_reference(const ds_conv_gemm_args* args, void* stream);

/* ----------------------------------------------------------------------------------------
 * Sampler step kernels (model/DiffSynthSampler.py).
 * -------------------------------------------------------------------------------------- */
/* Fused CFG combine + DDIM/DDPM update, one coalesced fp32 pass (K11).  Replaces :320-343.
   d_coef (device) = {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, cfg_scale}.
   d_eps_u NULL <=> CFG == 1.0 (:311-312).  d_z NULL only when sigma == 0.  n = element count (%4). */
int ds_ddim_step(const float* d_eps_u, const float* d_eps_c, const float* d_x, const float* d_z,
                 const float* d_coef, float* d_out, long long n, void* stream);
/* q_sample (:271-294): out = coef[0]*x0 + coef[1]*noise; per_sample > 0: d_coef holds one {a, b} pair per sample of
   per_sample elements (per-sample timesteps, :17-22), 0: one pair for the whole tensor. */
int ds_q_sample(const float* d_x0, const float* d_noise, const float* d_coef, float* d_out, long long n, long long per_sample,
                void* stream);
/* inpaint blend (:499-510): img = m*(coef[0]*guide + coef[1]*noise) + (1-m)*img; mask [B,mask_channels,H,W] with
   mask_channels 1 (broadcast over C) or C (the reference's inpaint caller repeats it, inpaint_with_text.py:229-231). */
int ds_mask_blend(const float* d_guide, const float* d_noise, const float* d_mask, int mask_channels, const float* d_coef,
                  float* d_img, int B, int C, long long hw, void* stream);

/* ----------------------------------------------------------------------------------------
 * U-Net pieces that are not GEMMs (model/diffusion_components.py, model/diffusion.py).
 * -------------------------------------------------------------------------------------- */
/* ConvNextBlock.ds_conv (:118,131) + time-embedding bias (:133-136) + partials of net[0] GroupNorm (:121).
   src_batch_mod > 0: sample n reads sample n % src_batch_mod of source 0 (source 1 is indexed by n). */
int ds_dwconv7(const void* d_src0, const void* d_src1, int C0, int C1, int src_batch_mod, const float* d_weight,
               const float* d_tbias, long long tbias_stride, void* d_out, void* d_stats, float eps, int N, int H, int W, void* stream);
int ds_dwconv7_stats_slots(int C, int H, int W);
/* init_conv 7x7 (model/diffusion.py:82,208): fp32 NCHW in, act16 NHWC out. */
int ds_stem_conv7(const float* d_x, int x_batch_mod, const float* d_weight, const float* d_bias, void* d_out,
                  int N, int Cin, int Cout, int H, int W, void* stream);
/* init_conv as a tensor-core GEMM: 7x7xCin patches -> act16 [N, H, W, 224] (k = ky*32 + kx*4 + ci), then ds_conv_gemm 1x1. */
int ds_stem_im2col(const float* d_x, void* d_col, int N, int Cin, int H, int W, void* stream);
/* SinusoidalPositionEmbeddings (:42-56) and the small Linear layers (time_mlp, per-block mlp,
   label_embedding, label_key/label_query): out = act_out(bias + W . act_in(in)); act 1 = GELU(erf). */
int ds_sinusoidal_embedding(const long long* d_t, float* d_out, int N, int dim, void* stream);
int ds_linear(const float* d_in, long long in_stride, const float* d_w, const float* d_bias, float* d_out,
              long long out_stride, int N, int K, int O, int act_in, int act_out, void* stream);
/* Linear attention core (LinearCrossAttentionAdd.forward :271-293; VQGAN LinearAttention :261-272). */
int ds_attn_chunks(long long npix);
long long ds_attn_part_floats(int N, int heads, long long npix);
int ds_attn_ctx_partial(const void* d_qkv, void* d_q_out, float* d_part, int N, int heads, long long npix,
                        int q_mode, float scale, void* stream);
/* Fused PreNorm (folded GroupNorm(1,C)) + to_qkv + q soft-max + partial context of the U-Net attention (:148-151,263,271-289):
   the to_qkv GEMM on tcgen05 with k and v kept on chip (P = exp(k - m) and S = P^T V in the epilogue); writes only
   q' = softmax_d(q) * scale (act16 [N][npix][128]) and the chunk partials ds_attn_finalize consumes.  heads must be 4 (x 32).
   d_x act16 NHWC [x_batch_mod or N][npix][C]; d_weight act16 [384][C] (q | k | v rows, gamma folded); d_e1 (nullable) / d_e2 fp32 [384];
   d_stats_in statistics buffer of x (nullable); d_sbias fp32 [N][sbias_stride] = label_query | label_key | 0 (nullable). */
int ds_attn_qkv_ctx(const void* d_x, int C, int x_batch_mod, const void* d_stats_in, int stats_in_slots, const void* d_weight,
                    const float* d_e1, const float* d_e2, const float* d_sbias, long long sbias_stride, void* d_q_out, float* d_part,
                    int N, int heads, long long npix, float scale, void* stream);
int ds_attn_finalize(const float* d_part, const float* d_wout, void* d_M, int N, int heads, long long npix,
                     int C, int Cout_pad, void* stream);
/* LinearCrossAttention ("linear_cat", diffusion_components.py:171-207): the condition contributes one extra key / value token,
   k = cat([k, label_key(emb)]), v = cat([v, label_value(emb)]) (:187-195); d_label_k / d_label_v fp32 [N][label_stride]. */
int ds_attn_finalize_cat(const float* d_part, const float* d_label_k, const float* d_label_v, long long label_stride,
                         const float* d_wout, void* d_M, int N, int heads, long long npix, int C, int Cout_pad, void* stream);
/* to_out[1] GroupNorm(1,C) + Residual (:264, :22-29): out = GN(y)*gamma+beta + x.  x_batch_mod > 0: sample n adds x[n % x_batch_mod]. */
int ds_gn_apply_residual(const void* d_y, const void* d_x, void* d_out, const void* d_stats, int slots,
                         const float* d_gamma, const float* d_beta, int N, int C, long long hw, int x_batch_mod, void* stream);

/* ----------------------------------------------------------------------------------------
 * VQGAN (model/VQGAN.py) and the spectrogram <-> waveform transforms (tools.py + librosa call sites).
 * -------------------------------------------------------------------------------------- */
/* VectorQuantizerEMA.forward eval (:98-146): bit-exact argmin over the expanded fp32 distance; out = x+(q-x). */
int ds_vq_quantize(const float* d_x, const float* d_codebook, int K, float* d_out, long long* d_idx,
                   int B, long long hw, void* stream);
/* GroupNorm(G, eps) + activation (Normalize/nonlinearity :12-27) as statistics + apply passes; act 0/1 relu/2 swish. */
int ds_group_stats(const void* d_x, void* d_part, int N, int C, int Cp, int G, long long hw, int chunks, void* stream);
int ds_gn_act(const void* d_x, void* d_out, const void* d_part, int chunks, const float* d_gamma, const float* d_beta,
              int N, int C, int Cp, int G, long long hw, float eps, int act, void* stream);
int ds_add_bf16(const void* d_a, const void* d_b, void* d_out, long long n, void* stream);
/* x[n][pixel][c] += bias[n*bias_stride + c] in place, act16 NHWC [N, hw, C] (bias_stride 0: one row): the time embedding that
   ResnetBlock adds between its two conv + GroupNorm + SiLU blocks (model/diffusion_components.py:79-104, use_convnext=False). */
int ds_add_channel_bias(void* d_x, const float* d_bias, long long bias_stride, int N, int C, long long hw, void* stream);
/* Decoder heads (:394-398): softplus / tanh / tanh of (a + b), fp32 NCHW [N,3,H,W]; b (nin_shortcut branch) nullable. */
int ds_decoder_head(const float* d_a, const float* d_b, float* d_out, int N, long long hw, void* stream);
int ds_nchw_f32_to_nhwc_bf16(const float* d_in, void* d_out, int N, int C, int Cp, long long hw, void* stream);
int ds_nhwc_bf16_to_nchw_f32(const void* d_in, float* d_out, int N, int C, int Cp, long long hw, void* stream);
/* decode_stft + depad_STFT + librosa.istft(hop 256, win 1024) (tools.py:334-345,185-191; utils.py:241). */
long long ds_istft_length(int T);
int ds_stft_decode_istft(const float* d_spec, float* d_frames, float* d_wave, int B, int T, void* stream);
/* librosa.stft(n_fft 1024, hop 256) + pad_STFT + encode_stft (sound2sound_with_text.py:85-94; tools.py:170-182,320-331). */
int ds_stft_encode(const float* d_wave, long long L, float* d_spec, int B, int Tpad, void* stream);
/* tools.decode_stft / tools.encode_stft (tools.py:334-345, 320-331) as stand-alone elementwise kernels behind the numpy drop-ins of
   diffusynth_b200.codec: d_enc [3][plane] real <-> d_D [plane] complex (re, im interleaved); is_double: float64 instead of float32. */
int ds_decode_stft(const void* d_enc, void* d_D, long long plane, int is_double, void* stream);
int ds_encode_stft(const void* d_D, void* d_enc, long long plane, int is_double, void* stream);
/* Griffin-Lim phase update (the loop body of librosa.griffinlim as called by tools.py:63-76,194-223: hop 256, win 1024,
   momentum 0.99): d_rebuilt = ds_stft_encode(ds_stft_decode_istft(d_spec)) as [B,3,512,T]; d_tprev fp32 [B,512,T,2] holds the
   previous rebuilt STFT (ignored and initialised when first != 0); channels 1, 2 (cos, sin) of d_spec are overwritten. */
int ds_griffinlim_update(const float* d_rebuilt, float* d_tprev, float* d_spec, float momentum, int first, int B, int T,
                         void* stream);

/* ----------------------------------------------------------------------------------------
 * Image products of the decode glue (webUI/natural_language_guided_4/utils.py), batched:
 *   spectrogram_to_Gradio_image (:8-50, with tools.np_power_to_db tools.py:41-50) and phase_to_Gradio_image (:53-91)
 *   of |D| / angle(D) for D = depad_STFT(decode_stft(spec)) (:229-238): uint8 [B, 513, T, 3], flipped vertically;
 *   latent_representation_to_Gradio_image (:94-128): per-channel min-max to 0..255, 8x enlarged, flipped: uint8 [B, 8H, 8W, 4].
 * dB scale and angle are evaluated in float64 like the reference (numpy on complex128).
 * -------------------------------------------------------------------------------------- */
int ds_spec_images(const float* d_spec, void* d_mag_img, void* d_phase_img, void* d_absmax /* scratch: 8 bytes per sample */,
                   int B, int T, void* stream);
int ds_latent_image(const float* d_lat, void* d_img, void* d_minmax /* scratch: 32 bytes per sample */, int B, int H, int W,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFUSYNTH_B200_H */
