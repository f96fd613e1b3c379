// Text-conditioning front end: the pieces of the CLAP text tower that are not GEMMs.
//   text2sound.py:89-109 encodes a prompt with  text_encoder.get_text_features(**tokenizer(...))  where text_encoder is the
//   multi_modal_model (app.py:55-59): ClapModel.get_text_features (transformers' ClapTextModel = RoBERTa-base: embeddings,
//   12 post-LayerNorm layers, tanh pooler; ClapProjectionLayer; L2 normalisation) followed by ProjectionHead
//   (model/multimodal_model.py:14-47,114-116).  The dense layers run on the tcgen05 GEMM (ds_conv_gemm as a 1x1 convolution over
//   the token axis, bias / GELU / residual in its epilogue); this file holds the embedding gather + LayerNorm, the row LayerNorms,
//   the masked soft-max attention core (sequence <= 514 tokens: K and V of one (sample, head) live in shared memory) and the
//   small fp32 tail (CLS gather, L2 normalise, ProjectionLayer's residual + LayerNorm).
#include "common.cuh"
#include "../../include/diffusynth_b200.h"

namespace ds {

static constexpr int TX_THREADS = 256;

// One block per token: out[t][:] = LayerNorm(word[ids[t]] + pos[pos_id[t]] + type[0]) as act16; position ids follow
// ClapTextEmbeddings.create_position_ids_from_input_ids: cumsum(mask) * mask + padding_idx with mask = ids != padding_idx.
__global__ void __launch_bounds__(TX_THREADS)
text_embed_ln_kernel(const long long* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
                     const float* __restrict__ gamma, const float* __restrict__ beta, act_t* __restrict__ out, int L, int D, int pad_idx, float eps) {
  pdl_enter();
  __shared__ float s_red[2][TX_THREADS / 32];
  __shared__ int s_pos;
  const int t = blockIdx.x, b = t / L, l = t - b * L, tid = threadIdx.x;
  if (tid < 32) {
    int cnt = 0;
    for (int j = tid; j <= l; j += 32) cnt += ids[(size_t)b * L + j] != pad_idx;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (tid == 0) s_pos = (ids[t] != pad_idx) ? cnt + pad_idx : pad_idx;
  }
  __syncthreads();
  const float* wr = word + (size_t)ids[t] * D;
  const float* pr = pos + (size_t)s_pos * D;
  float v[4];      // D <= 4 * TX_THREADS
  float s = 0.f;
  int n = 0;
  for (int c = tid; c < D; c += TX_THREADS, ++n) { v[n] = __ldg(wr + c) + __ldg(type0 + c) + __ldg(pr + c); s += v[n]; }
  s = warp_sum(s);
  if ((tid & 31) == 0) s_red[0][tid >> 5] = s;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < TX_THREADS / 32; ++w) mean += s_red[0][w];
  mean /= (float)D;
  float q = 0.f;
  for (int i = 0; i < n; ++i) { const float d = v[i] - mean; q += d * d; }
  q = warp_sum(q);
  if ((tid & 31) == 0) s_red[1][tid >> 5] = q;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < TX_THREADS / 32; ++w) var += s_red[1][w];
  const float rstd = rsqrtf(var / (float)D + eps);
  n = 0;
  for (int c = tid; c < D; c += TX_THREADS, ++n) out[(size_t)t * D + c] = f2act((v[n] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
}

// Row LayerNorm of act16 [T][D] in place (the residual was already added by the producing GEMM's epilogue): one warp per row.
__global__ void __launch_bounds__(TX_THREADS)
layernorm_rows_kernel(act_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int T, int D, float eps) {
  pdl_enter();
  const int row = blockIdx.x * (TX_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  act_t* xr = x + (size_t)row * D;
  float v[32];      // D <= 1024
  const int n = D / 32;
  float s = 0.f;
  for (int i = 0; i < n; ++i) { v[i] = act2f(xr[lane + 32 * i]); s += v[i]; }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int i = 0; i < n; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  for (int i = 0; i < n; ++i) { const int c = lane + 32 * i; xr[c] = f2act((v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c)); }
}

// Masked soft-max attention of one (sample, head): qkv act16 [B][L][3*H*64] (q | k | v, heads contiguous inside each), mask
// int64 [B][L] (1 = token, 0 = padding; padded KEYS get probability 0 like the reference's additive -inf mask; padded query
// rows are computed but never read).  K and V (fp32, padded rows) sit in shared memory, one warp per query row.
static constexpr int TA_DH = 64, TA_PITCH = TA_DH + 1;
__global__ void __launch_bounds__(TX_THREADS)
text_attention_kernel(const act_t* __restrict__ qkv, const long long* __restrict__ mask, act_t* __restrict__ out, int L, int heads, float scale) {
  pdl_enter();
  extern __shared__ float ta_smem[];
  float* s_k = ta_smem;                          // [L][65]
  float* s_v = s_k + (size_t)L * TA_PITCH;       // [L][65]
  float* s_p = s_v + (size_t)L * TA_PITCH;       // [warps][L]
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D3 = 3 * heads * TA_DH;
  const act_t* base = qkv + (size_t)b * L * D3;
  for (int i = tid; i < L * TA_DH; i += TX_THREADS) {
    const int j = i / TA_DH, d = i - j * TA_DH;
    s_k[j * TA_PITCH + d] = act2f(base[(size_t)j * D3 + heads * TA_DH + h * TA_DH + d]);
    s_v[j * TA_PITCH + d] = act2f(base[(size_t)j * D3 + 2 * heads * TA_DH + h * TA_DH + d]);
  }
  __syncthreads();
  float* pw = s_p + (size_t)warp * L;
  for (int i = warp; i < L; i += TX_THREADS / 32) {
    const act_t* qr = base + (size_t)i * D3 + h * TA_DH;
    const float q0 = act2f(qr[lane]) * scale, q1 = act2f(qr[lane + 32]) * scale;
    float mx = -INFINITY;
    for (int j0 = 0; j0 < L; j0 += 32) {
      const int j = j0 + lane, jc = j < L ? j : L - 1;      // (every lane takes part in the shuffles; the tail lanes redo the last key)
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < 32; ++d) {
        a = fmaf(__shfl_sync(0xffffffffu, q0, d), s_k[jc * TA_PITCH + d], a);
        a = fmaf(__shfl_sync(0xffffffffu, q1, d), s_k[jc * TA_PITCH + 32 + d], a);
      }
      const float sc = (j < L && mask[(size_t)b * L + jc] != 0) ? a : -INFINITY;
      if (j < L) pw[j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = mx == -INFINITY ? 0.f : __expf(pw[j] - mx); pw[j] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < L; ++j) { const float p = pw[j]; o0 = fmaf(p, s_v[j * TA_PITCH + lane], o0); o1 = fmaf(p, s_v[j * TA_PITCH + 32 + lane], o1); }
    act_t* orow = out + ((size_t)b * L + i) * heads * TA_DH + h * TA_DH;
    orow[lane] = f2act(o0 * inv);
    orow[lane + 32] = f2act(o1 * inv);
    __syncwarp();
  }
}

// out[b][:] = float(x[b * L][:]) : the CLS rows of the last hidden state, as fp32 for the pooler
__global__ void cls_gather_kernel(const act_t* __restrict__ x, float* __restrict__ out, int B, int L, int D) {
  pdl_enter();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * D; i += gridDim.x * blockDim.x) {
    const int b = i / D, c = i - b * D;
    out[i] = act2f(x[(size_t)b * L * D + c]);
  }
}

// mode 0: x[b][:] /= max(||x[b]||_2, 1e-12)   (F.normalize, ClapModel.get_text_features)
// mode 1: x[b][:] = LayerNorm(x[b] + y[b]) * gamma + beta   (ProjectionLayer: fc(gelu(projected)) + projected, layer_norm; eps)
__global__ void __launch_bounds__(TX_THREADS)
text_rows_f32_kernel(float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta, int D, float eps, int mode) {
  pdl_enter();
  __shared__ float s_red[2][TX_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* xr = x + (size_t)b * D;
  float v[4];
  int n = 0;
  float s = 0.f, q = 0.f;
  for (int c = tid; c < D; c += TX_THREADS, ++n) {
    v[n] = xr[c] + (mode == 1 ? y[(size_t)b * D + c] : 0.f);
    s += v[n]; q += v[n] * v[n];
  }
  s = warp_sum(s); q = warp_sum(q);
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = s; s_red[1][tid >> 5] = q; }
  __syncthreads();
  s = 0.f; q = 0.f;
  for (int w = 0; w < TX_THREADS / 32; ++w) { s += s_red[0][w]; q += s_red[1][w]; }
  if (mode == 0) {
    const float inv = 1.0f / fmaxf(sqrtf(q), 1e-12f);
    n = 0;
    for (int c = tid; c < D; c += TX_THREADS, ++n) xr[c] = v[n] * inv;
    return;
  }
  const float mean = s / (float)D;
  __syncthreads();
  float qq = 0.f;
  for (int i = 0; i < n; ++i) { const float d = v[i] - mean; qq += d * d; }
  qq = warp_sum(qq);
  if ((tid & 31) == 0) s_red[0][tid >> 5] = qq;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < TX_THREADS / 32; ++w) var += s_red[0][w];
  const float rstd = rsqrtf(var / (float)D + eps);
  n = 0;
  for (int c = tid; c < D; c += TX_THREADS, ++n) xr[c] = (v[n] - mean) * rstd * gamma[c] + beta[c];
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_text_embed_ln(const long long* d_ids, const float* d_word, const float* d_pos, const float* d_type0, const float* d_gamma, const float* d_beta,
                     void* d_out, int B, int L, int D, int pad_idx, float eps, void* stream) {
  DS_REQUIRE(d_ids && d_word && d_pos && d_type0 && d_gamma && d_beta && d_out && B > 0 && L > 0, "ds_text_embed_ln: bad arguments");
  DS_REQUIRE(D > 0 && D <= 4 * TX_THREADS, "ds_text_embed_ln: hidden size %d (at most %d)", D, 4 * TX_THREADS);
  DS_CHECK_CUDA(launch_pdl(text_embed_ln_kernel, dim3(B * L), dim3(TX_THREADS), 0, (cudaStream_t)stream, d_ids, d_word, d_pos, d_type0, d_gamma, d_beta,
                           (act_t*)d_out, L, D, pad_idx, eps));
  return DS_OK;
}

int ds_layernorm_rows(void* d_x, const float* d_gamma, const float* d_beta, int T, int D, float eps, void* stream) {
  DS_REQUIRE(d_x && d_gamma && d_beta && T > 0 && D > 0 && D % 32 == 0 && D <= 1024, "ds_layernorm_rows: bad arguments (D=%d: multiple of 32, at most 1024)", D);
  const int rows = TX_THREADS / 32;
  DS_CHECK_CUDA(launch_pdl(layernorm_rows_kernel, dim3((T + rows - 1) / rows), dim3(TX_THREADS), 0, (cudaStream_t)stream, (act_t*)d_x, d_gamma, d_beta, T, D, eps));
  return DS_OK;
}

int ds_text_attention(const void* d_qkv, const long long* d_mask, void* d_out, int B, int L, int heads, float scale, void* stream) {
  DS_REQUIRE(d_qkv && d_mask && d_out && B > 0 && L > 0 && heads > 0 && B <= 65535, "ds_text_attention: bad arguments");
  const size_t smem = ((size_t)2 * L * TA_PITCH + (size_t)(TX_THREADS / 32) * L) * sizeof(float);
  DS_REQUIRE(smem <= 220 * 1024, "ds_text_attention: sequence length %d does not fit shared memory", L);
  DS_CHECK_CUDA(cudaFuncSetAttribute(text_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DS_CHECK_CUDA(launch_pdl(text_attention_kernel, dim3(heads, B), dim3(TX_THREADS), smem, (cudaStream_t)stream, (const act_t*)d_qkv, d_mask, (act_t*)d_out, L, heads, scale));
  return DS_OK;
}

int ds_cls_gather(const void* d_x, float* d_out, int B, int L, int D, void* stream) {
  DS_REQUIRE(d_x && d_out && B > 0 && L > 0 && D > 0, "ds_cls_gather: bad arguments");
  DS_CHECK_CUDA(launch_pdl(cls_gather_kernel, dim3((B * D + 255) / 256), dim3(256), 0, (cudaStream_t)stream, (const act_t*)d_x, d_out, B, L, D));
  return DS_OK;
}

int ds_l2_normalize_rows(float* d_x, int B, int D, void* stream) {
  DS_REQUIRE(d_x && B > 0 && D > 0 && D <= 4 * TX_THREADS, "ds_l2_normalize_rows: bad arguments");
  DS_CHECK_CUDA(launch_pdl(text_rows_f32_kernel, dim3(B), dim3(TX_THREADS), 0, (cudaStream_t)stream, d_x, (const float*)nullptr, (const float*)nullptr,
                           (const float*)nullptr, D, 0.f, 0));
  return DS_OK;
}

int ds_add_layernorm_rows_f32(float* d_x, const float* d_y, const float* d_gamma, const float* d_beta, int B, int D, float eps, void* stream) {
  DS_REQUIRE(d_x && d_y && d_gamma && d_beta && B > 0 && D > 0 && D <= 4 * TX_THREADS, "ds_add_layernorm_rows_f32: bad arguments");
  DS_CHECK_CUDA(launch_pdl(text_rows_f32_kernel, dim3(B), dim3(TX_THREADS), 0, (cudaStream_t)stream, d_x, d_y, d_gamma, d_beta, D, eps, 1));
  return DS_OK;
}

}  // extern "C"
