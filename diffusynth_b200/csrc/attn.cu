// Linear ("efficient") attention core.
//   U-Net  LinearCrossAttentionAdd.forward, model/diffusion_components.py:271-293
//   VQGAN  LinearAttention.forward,         model/VQGAN.py:261-272
// Given qkv = to_qkv(x) [N, n, 3*hidden] (bf16 NHWC, q | k | v blocks of `hidden` channels, head h owns
// channels [32h, 32h+32) of each block; for the U-Net the label_query/label_key biases were already added
// by the to_qkv GEMM epilogue):
//   q' = softmax_d(q) * scale            (U-Net)      or  q' = q  (VQGAN: q is neither soft-maxed nor scaled)
//   ctx[h][d][e] = sum_n softmax_n(k)[d,n] * v[e,n]
//   M[c][(h,d)]  = sum_e Wout[c][(h,e)] * ctx[h][d][e]        (per sample)
// so that  to_out(ctx^T q') = M q' + bias  is a per-sample 1x1 GEMM done by ds_conv_gemm (per_sample_weights).
// The n x n attention matrix never exists (O(n) attention): this is not a flash-attention problem.
#include "common.cuh"
#include "../../include/diffusynth_b200.h"

namespace ds {

static constexpr int AT_D = 32;          // dim_head
static constexpr int AT_PIX = 128;       // pixels per block
static constexpr int AT_PART = AT_D * AT_D + 2 * AT_D;   // S[32][32], Z[32], m[32]

// ldmatrix (transposing) and the legacy warp-level MMA: the per-chunk context S = P^T V is a 32 x 32 x 128 product -- far too
// small for a tcgen05 tile, and as FFMA2 from fp32 shared memory it was bound by shared-memory bandwidth (ncu: LSU wavefronts
// 88 % of peak).  With P and V kept 16-bit in shared memory, eight m16n8k16 MMAs per warp replace 256 FFMA2 per thread.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_m16n8k16_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t to_half2_bits(uint32_t v) {      // a pair of stored activations as IEEE half2 bits
#ifdef DS_OPERANDS_BF16
  const __half2 h = __floats2half2_rn(lo16(v), hi16(v));
  return *reinterpret_cast<const uint32_t*>(&h);
#else
  return v;
#endif
}

static constexpr int AT_PITCH = 40;      // halves per shared-memory row (80 bytes): ldmatrix reads of 8 rows are conflict-free

static constexpr int AT_SUB = 4;         // 128-pixel sub-chunks per block: one partial per 512 pixels

// grid = (chunks, heads, N), block = 256.  A block walks over up to AT_SUB sub-chunks of 128 pixels with a running column
// maximum (online softmax over n): S and Z are rescaled by exp(m_old - m_new) when the maximum moves, so one partial
// (S[32][32], Z[32], m[32]) leaves the block per 512 pixels.
__global__ void __launch_bounds__(256)
attn_ctx_partial_kernel(const act_t* __restrict__ qkv, int hidden, int npix, int q_mode, float scale,
                        act_t* __restrict__ qout /* [N, n, hidden] */, float* __restrict__ part, int chunks) {
  pdl_enter();
  __shared__ __align__(16) __half s_k[AT_PIX][AT_PITCH];     // k, then p = exp(k - m), as IEEE half
  __shared__ __align__(16) __half s_v[AT_PIX][AT_PITCH];
  __shared__ float s_red[8][AT_D];
  __shared__ float s_m[AT_D], s_sc[AT_D];
  const int chunk = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  const int ld = 3 * hidden;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = warp >> 2, nt = warp & 3;                   // this warp's 16 x 8 tile of S
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  float z_run = 0.f;                                         // running column sum of this warp's pixel rows (column = lane)
  if (tid < AT_D) s_m[tid] = -INFINITY;

  for (int sub = 0; sub < AT_SUB; ++sub) {
    const long long p0 = ((long long)chunk * AT_SUB + sub) * AT_PIX;
    if (p0 >= npix) break;
    const act_t* base = qkv + ((size_t)n * npix + p0) * ld + head * AT_D;
    if (sub > 0) __syncthreads();                            // the previous sub-chunk's MMAs are done with s_k / s_v
    // ---- stage k, v (16-bit, as stored) ; rows beyond npix are neutral (k = -inf -> p = 0, v = 0)
    {
      const int pix = tid >> 1, half = tid & 1;   // 16 channels per thread
      const bool ok = p0 + pix < npix;
      uint4 k2[2], v2[2];
      const uint32_t ninf2 = 0xFC00FC00u;          // (-inf, -inf) as half2
      k2[0] = k2[1] = make_uint4(ninf2, ninf2, ninf2, ninf2);
      v2[0] = v2[1] = make_uint4(0, 0, 0, 0);
      if (ok) {
        ldg_256(base + (size_t)pix * ld + hidden + half * 16, k2[0], k2[1]);      // one 32-byte sector per lane and tensor
        ldg_256(base + (size_t)pix * ld + 2 * hidden + half * 16, v2[0], v2[1]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          k2[i] = make_uint4(to_half2_bits(k2[i].x), to_half2_bits(k2[i].y), to_half2_bits(k2[i].z), to_half2_bits(k2[i].w));
          v2[i] = make_uint4(to_half2_bits(v2[i].x), to_half2_bits(v2[i].y), to_half2_bits(v2[i].z), to_half2_bits(v2[i].w));
        }
      }
      uint4* kd = reinterpret_cast<uint4*>(&s_k[pix][half * 16]);
      uint4* vd = reinterpret_cast<uint4*>(&s_v[pix][half * 16]);
      kd[0] = k2[0]; kd[1] = k2[1];
      vd[0] = v2[0]; vd[1] = v2[1];
      // ---- q: softmax over the 32 channels of this head (two threads per pixel), scaled; or plain copy.
      // The math runs for every thread (rows beyond npix compute on zeros) so the pair shuffles stay convergent.
      {
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
        if (ok) ldg_256(base + (size_t)pix * ld + half * 16, q0, q1);
        if (q_mode == 0) {
          const uint32_t qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
          float f[16];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) { f[2 * j] = lo16(qq[j]); f[2 * j + 1] = hi16(qq[j]); mx = fmaxf(mx, fmaxf(f[2 * j], f[2 * j + 1])); }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { f[j] = __expf(f[j] - mx); sum += f[j]; }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          const float inv = scale / sum;
          q0 = make_uint4(pack16(f[0] * inv, f[1] * inv), pack16(f[2] * inv, f[3] * inv), pack16(f[4] * inv, f[5] * inv), pack16(f[6] * inv, f[7] * inv));
          q1 = make_uint4(pack16(f[8] * inv, f[9] * inv), pack16(f[10] * inv, f[11] * inv), pack16(f[12] * inv, f[13] * inv), pack16(f[14] * inv, f[15] * inv));
        }
        if (ok) stg_256(qout + ((size_t)n * npix + p0 + pix) * hidden + head * AT_D + half * 16, q0, q1);
      }
    }
    __syncthreads();
    // ---- column max of k over the sub-chunk
    {
      float mx = -INFINITY;
      for (int p = warp * 16; p < warp * 16 + 16; ++p) mx = fmaxf(mx, __half2float(s_k[p][lane]));
      s_red[warp][lane] = mx;
    }
    __syncthreads();
    if (tid < AT_D) {
      const float m_old = s_m[tid];
      float mx = m_old;
#pragma unroll
      for (int w = 0; w < 8; ++w) mx = fmaxf(mx, s_red[w][tid]);
      s_m[tid] = mx;
      s_sc[tid] = __expf(m_old - mx);                        // 0 for the first sub-chunk (m_old = -inf)
    }
    __syncthreads();
    // ---- p = exp(k - m) in place (rounded to half: the same values enter Z and S); running column sums Z
    {
      const float m = s_m[lane];
      float z = 0.f;
      for (int p = warp * 16; p < warp * 16 + 16; ++p) {
        const __half e = __float2half_rn(__expf(__half2float(s_k[p][lane]) - m));
        s_k[p][lane] = e;
        z += __half2float(e);
      }
      z_run = fmaf(z_run, s_sc[lane], z);
    }
    __syncthreads();
    // ---- S[d][e] += sum over the sub-chunk's 128 pixels of p[pix][d] * v[pix][e]: A = P^T (d x pix), B = V (pix x e), both
    //      read transposed from their [pixel][channel] rows.  Warp w owns the 16 x 8 output tile (d tile w >> 2, e tile w & 3)
    //      for ALL pixels (8 MMAs of k = 16), so there is no cross-warp reduction.
    {
      const int r = lane & 7, mat = lane >> 3;
      const float sc_lo = s_sc[mt * 16 + (lane >> 2)], sc_hi = s_sc[mt * 16 + (lane >> 2) + 8];
      c[0] *= sc_lo; c[1] *= sc_lo; c[2] *= sc_hi; c[3] *= sc_hi;
#pragma unroll
      for (int ks = 0; ks < AT_PIX / 16; ++ks) {
        uint32_t a[4], bfr[2];
        ldmatrix_x4_trans(a, (uint32_t)__cvta_generic_to_shared(&s_k[ks * 16 + r + ((mat >> 1) & 1) * 8][mt * 16 + (mat & 1) * 8]));
        ldmatrix_x2_trans(bfr, (uint32_t)__cvta_generic_to_shared(&s_v[ks * 16 + r + (mat & 1) * 8][nt * 8]));
        mma_m16n8k16_f16(c, a, bfr[0], bfr[1]);
      }
    }
  }
  float* po = part + (((size_t)n * gridDim.y + head) * chunks + chunk) * AT_PART;
  {
    const int d = mt * 16 + (lane >> 2), e = nt * 8 + (lane & 3) * 2;
    *reinterpret_cast<float2*>(po + d * AT_D + e) = make_float2(c[0], c[1]);
    *reinterpret_cast<float2*>(po + (d + 8) * AT_D + e) = make_float2(c[2], c[3]);
  }
  s_red[warp][lane] = z_run;       // (the last read of s_red, the max merge, is behind two barriers)
  __syncthreads();
  if (tid < AT_D) {
    float z = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) z += s_red[w][tid];
    po[AT_D * AT_D + tid] = z;
    po[AT_D * AT_D + AT_D + tid] = s_m[tid];
  }
}

// attn_reduce: grid = (heads, N), block = 256: merge the chunk partials (log-sum-exp style) and normalise
//   ctx[d][e] = sum_c S_c[d][e] exp(m_c[d] - M[d]) / sum_c Z_c[d] exp(m_c[d] - M[d])        -> fp32 [N][heads][32][32]
// label_k / label_v (nullable, [N][label_stride], element head*32 + d): the extra key / value token of LinearCrossAttention
// ("linear_cat", diffusion_components.py:187-195: k = cat([k, label_k]), v = cat([v, label_v]) along the sequence) joins the merge
// like one more chunk with a single position: max candidate label_k[d], Z += exp(label_k[d] - M), S[d][e] += exp(.) * label_v[e].
// All 256 threads take part in every phase and every phase's loads are independent of each other (the first version walked the chunks
// with 32 threads in a load -> max / load -> exp -> add chain: 2 x chunks dependent global-memory round trips before the merge proper
// started, 11-24 us per launch of pure latency).  Thread (d = tid & 31, slice = tid >> 5) covers chunks slice, slice + 8, ...; the
// per-chunk weights exp(m_c[d] - M[d]) are kept in shared memory for the merge of S (up to AT_MAXW chunks; recomputed beyond that).
static constexpr int AT_MAXW = 64;
__global__ void __launch_bounds__(256)
attn_reduce_kernel(const float* __restrict__ part, int chunks, float* __restrict__ ctx, const float* __restrict__ label_k,
                   const float* __restrict__ label_v, long long label_stride) {
  pdl_enter();
  __shared__ float s_part[8][AT_D], s_M[AT_D], s_Zinv[AT_D], s_lw[AT_D];
  __shared__ float s_w[AT_MAXW][AT_D];
  const int head = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const int d = tid & 31, sl = tid >> 5;
  const float* pb = part + ((size_t)n * gridDim.x + head) * chunks * AT_PART;
  const float* pm = pb + AT_D * AT_D + AT_D + d;       // m_c[d] of chunk 0
  const float* pz = pb + AT_D * AT_D + d;              // Z_c[d] of chunk 0
  const float lk = label_k ? __ldg(label_k + (size_t)n * label_stride + head * AT_D + d) : -INFINITY;
  float mx = -INFINITY;
#pragma unroll 4
  for (int c = sl; c < chunks; c += 8) mx = fmaxf(mx, __ldg(pm + (size_t)c * AT_PART));
  s_part[sl][d] = mx;
  __syncthreads();
  mx = lk;
#pragma unroll
  for (int w = 0; w < 8; ++w) mx = fmaxf(mx, s_part[w][d]);
  __syncthreads();                                     // s_part is reused for the normaliser
  float z = 0.f;
#pragma unroll 4
  for (int c = sl; c < chunks; c += 8) {
    const float w = __expf(__ldg(pm + (size_t)c * AT_PART) - mx);
    if (c < AT_MAXW) s_w[c][d] = w;
    z = fmaf(__ldg(pz + (size_t)c * AT_PART), w, z);
  }
  s_part[sl][d] = z;
  __syncthreads();
  if (sl == 0) {
    const float lw = label_k ? __expf(lk - mx) : 0.f;
    z = lw;
#pragma unroll
    for (int w = 0; w < 8; ++w) z += s_part[w][d];
    s_M[d] = mx; s_Zinv[d] = 1.0f / z; s_lw[d] = lw;
  }
  __syncthreads();
  // merge of S: element i = tid + 256 k of the 32 x 32 matrix (row i / 32 = 8 k + slice), chunks in order
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* ps = pb + tid;
#pragma unroll 4
  for (int c = 0; c < chunks; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dr = 8 * k + sl;
      const float w = c < AT_MAXW ? s_w[c][dr] : __expf(__ldg(pb + (size_t)c * AT_PART + AT_D * AT_D + AT_D + dr) - s_M[dr]);
      acc[k] = fmaf(__ldg(ps + (size_t)c * AT_PART + 256 * k), w, acc[k]);
    }
  }
  float* co = ctx + ((size_t)n * gridDim.x + head) * AT_D * AT_D;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dr = 8 * k + sl;
    float sv = acc[k];
    if (label_v) sv = fmaf(s_lw[dr], __ldg(label_v + (size_t)n * label_stride + head * AT_D + d), sv);
    co[tid + 256 * k] = sv * s_Zinv[dr];
  }
}

// attn_fold: grid = (ceil(Cout_pad/64), heads, N), block = 256:
//   M[n][c][head*32 + d] = sum_e Wout[c][head*32 + e] * ctx[n][head][d][e]    -> 16-bit [N][Cout_pad][hidden]
__global__ void __launch_bounds__(256)
attn_fold_kernel(const float* __restrict__ ctx, const float* __restrict__ wout /* [C][hidden] */, int C, int Cout_pad, int hidden,
                 act_t* __restrict__ M) {
  pdl_enter();
  __shared__ __align__(16) float s_w[64][AT_D];
  const int c0 = blockIdx.x * 64, head = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  const int warp = tid >> 5, d = tid & 31;
  // lane d keeps row d of ctx in registers (staged through shared memory once per block: every warp needs all 32 rows);
  // the warp walks over Wout rows read as broadcast float4 from shared memory
  __shared__ float s_c[AT_D][AT_D + 1];
  {
    const float4 v = __ldg(reinterpret_cast<const float4*>(ctx + ((size_t)n * gridDim.y + head) * AT_D * AT_D) + tid);   // 256 x 16 B = 4 KB
    const int r = tid >> 3, e0 = (tid & 7) * 4;
    s_c[r][e0] = v.x; s_c[r][e0 + 1] = v.y; s_c[r][e0 + 2] = v.z; s_c[r][e0 + 3] = v.w;
  }
  for (int i = tid; i < 64 * AT_D; i += 256) {
    const int cl = i / AT_D, e = i % AT_D;
    s_w[cl][e] = (c0 + cl < C) ? __ldg(wout + (size_t)(c0 + cl) * hidden + head * AT_D + e) : 0.f;
  }
  __syncthreads();
  float cr[AT_D];
#pragma unroll
  for (int e = 0; e < AT_D; ++e) cr[e] = s_c[d][e];
  act_t* Mn = M + (size_t)n * Cout_pad * hidden;
#pragma unroll 2
  for (int cl = warp; cl < 64; cl += 8) {
    if (c0 + cl >= Cout_pad) break;
    float acc = 0.f;
#pragma unroll
    for (int e4 = 0; e4 < AT_D / 4; ++e4) {
      const float4 w4 = *reinterpret_cast<const float4*>(&s_w[cl][4 * e4]);
      acc = fmaf(w4.x, cr[4 * e4], acc); acc = fmaf(w4.y, cr[4 * e4 + 1], acc);
      acc = fmaf(w4.z, cr[4 * e4 + 2], acc); acc = fmaf(w4.w, cr[4 * e4 + 3], acc);
    }
    Mn[(size_t)(c0 + cl) * hidden + head * AT_D + d] = f2act(acc);
  }
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_attn_chunks(long long npix) { return (int)((npix + AT_PIX * AT_SUB - 1) / (AT_PIX * AT_SUB)); }
/* scratch: chunk partials followed by the merged ctx [N][heads][32][32] */
long long ds_attn_part_floats(int N, int heads, long long npix) { return (long long)N * heads * (ds_attn_chunks(npix) * AT_PART + AT_D * AT_D); }

/* q' (bf16 [N, npix, hidden]) and chunk partials of ctx.  q_mode 0: softmax over head dim * scale; 1: copy. */
int ds_attn_ctx_partial(const void* d_qkv, void* d_q_out, float* d_part, int N, int heads, long long npix, int q_mode, float scale, void* stream) {
  DS_REQUIRE(d_qkv && d_q_out && d_part && N > 0 && heads > 0 && npix > 0, "ds_attn_ctx_partial: bad arguments");
  const int chunks = ds_attn_chunks(npix);
  DS_REQUIRE(N <= 65535 && heads <= 65535, "ds_attn_ctx_partial: grid too large");
  DS_CHECK_CUDA(launch_pdl(attn_ctx_partial_kernel, dim3(dim3(chunks, heads, N)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const act_t*)d_qkv, heads * AT_D, (int)npix,
                                                                                   q_mode, scale, (act_t*)d_q_out, d_part, chunks));
  DS_CHECK_CUDA(cudaGetLastError());
  return DS_OK;
}

/* Merge partials and fold to_out's weight: d_M bf16 [N][Cout_pad][heads*32]. */
static int attn_finalize_impl(const float* d_part, const float* d_label_k, const float* d_label_v, long long label_stride, const float* d_wout,
                              void* d_M, int N, int heads, long long npix, int C, int Cout_pad, void* stream) {
  DS_REQUIRE(d_part && d_wout && d_M && N > 0 && heads > 0 && C > 0 && Cout_pad >= C, "ds_attn_finalize: bad arguments");
  DS_REQUIRE((d_label_k == nullptr) == (d_label_v == nullptr) && (!d_label_k || label_stride >= heads * AT_D), "ds_attn_finalize_cat: label_k/label_v");
  const int chunks = ds_attn_chunks(npix);
  float* ctx = const_cast<float*>(d_part) + (size_t)N * heads * chunks * AT_PART;
  DS_REQUIRE(N <= 65535 && heads <= 65535, "ds_attn_finalize: grid too large");
  DS_CHECK_CUDA(launch_pdl(attn_reduce_kernel, dim3(dim3(heads, N)), dim3(256), (size_t)(0), (cudaStream_t)stream, d_part, chunks, ctx, d_label_k, d_label_v, label_stride));
  DS_CHECK_CUDA(launch_pdl(attn_fold_kernel, dim3(dim3((Cout_pad + 63) / 64, heads, N)), dim3(256), (size_t)(0), (cudaStream_t)stream, ctx, d_wout, C, Cout_pad, heads * AT_D, (act_t*)d_M));
  DS_CHECK_CUDA(cudaGetLastError());
  return DS_OK;
}
int ds_attn_finalize(const float* d_part, const float* d_wout, void* d_M, int N, int heads, long long npix, int C, int Cout_pad, void* stream) {
  return attn_finalize_impl(d_part, nullptr, nullptr, 0, d_wout, d_M, N, heads, npix, C, Cout_pad, stream);
}
/* The same with the condition's extra key / value token of LinearCrossAttention ("linear_cat"): d_label_k, d_label_v fp32
   [N][label_stride] (element head*32 + d) = label_key(emb), label_value(emb). */
int ds_attn_finalize_cat(const float* d_part, const float* d_label_k, const float* d_label_v, long long label_stride, const float* d_wout,
                         void* d_M, int N, int heads, long long npix, int C, int Cout_pad, void* stream) {
  DS_REQUIRE(d_label_k && d_label_v, "ds_attn_finalize_cat: label_k and label_v are required");
  return attn_finalize_impl(d_part, d_label_k, d_label_v, label_stride, d_wout, d_M, N, heads, npix, C, Cout_pad, stream);
}

}  // extern "C"
