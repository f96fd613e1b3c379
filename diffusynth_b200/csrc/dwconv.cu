// Depthwise 7x7 convolution (ConvNextBlock.ds_conv, model/diffusion_components.py:118,131) fused with the
// per-sample time-embedding bias (:133-136) and the (sum, sumsq) partials of the GroupNorm(1,C) that
// follows (:121), and the 7x7 stem convolution (init_conv, model/diffusion.py:82).
#include "common.cuh"
#include "../../include/diffusynth_b200.h"

namespace ds {

// ---------------------------------------------------------------------------------------------
// dwconv7: 16-bit NHWC in (one or two channel-concatenated sources), 16-bit NHWC out, fp32 accumulation.
//
// A depthwise conv has no channel reduction, so it cannot fill a tcgen05 tile; on the fp32 pipe it is bound by FMA
// issue (49 MAC per element; measured 24 TFLOP/s = 67% of the register-operand FFMA/FFMA2 rate, 6.4 ms per U-Net
// evaluation).  Here it runs on the warp-level tensor-core path instead: mma.sync.m16n8k16 with
//   M = 16 output pixels of one image row,  N = 8 channels,  K = 2 horizontal taps x 8 channels,
//   A[m][(t,c)] = x[row+ky][m + kx0 + t][c]          -- ldmatrix.x4 straight from the NHWC halo tile in smem
//   B[(t,c)][n] = w[ky][kx0+t][n] if c == n else 0     -- per-thread constant fragments (block-diagonal weights)
// i.e. 28 MMAs (7 rows x 4 tap pairs, the 8th tap is a zero weight) per 16x8 output block; 1/8 of the MMA lanes are
// useful, which still leaves the kernel bound by shared-memory / HBM traffic rather than by math.
// block = 256 threads = 8 warps; tile = 8 rows x 16 cols x 32 channels; warp = (channel group of 8) x (4 rows).
// ---------------------------------------------------------------------------------------------
static constexpr int DW_TH = 8, DW_TW = 16, DW_CB = 32;
static constexpr int DW_HH = DW_TH + 6, DW_HW = DW_TW + 6;
static constexpr int DW_PIX = 40;     // halves per staged pixel: 32 channels + 8 pad -> conflict-free ldmatrix rows (80 B pitch)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef DS_OPERANDS_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#endif
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
dwconv7_kernel(const act_t* __restrict__ src0, const act_t* __restrict__ src1, int C0, int C1, int src_batch_mod,
               const float* __restrict__ weight,   // [49][C] (tap-major)
               const float* __restrict__ tbias, long long tbias_stride,   // [N or 1][>=C]: conv bias + time projection
               act_t* __restrict__ out, float2* __restrict__ stats, float stats_inv_count, float eps, int H, int W, int tiles_w, int tiles) {
  __shared__ __align__(16) act_t s_in[(DW_HH * DW_HW + 2) * DW_PIX];
  __shared__ float s_red[16];
  const int C = C0 + C1;
  const int tile = blockIdx.x, cblk = blockIdx.y, n = blockIdx.z;
  const int h0 = (tile / tiles_w) * DW_TH, w0 = (tile % tiles_w) * DW_TW;
  const int c0 = cblk * DW_CB;
  const int nsrc = src_batch_mod > 0 ? n % src_batch_mod : n;
  const act_t* src;
  int Cs, cs0;
  if (c0 < C0) { src = src0; Cs = C0; cs0 = c0; } else { src = src1; Cs = C1; cs0 = c0 - C0; }
  src += (size_t)nsrc * H * W * Cs;

  // stage the halo tile (zero outside the image = the conv's padding); 16-byte pieces, 4 per pixel
  for (int i = threadIdx.x; i < (DW_HH * DW_HW + 2) * 4; i += 256) {
    const int piece = i & 3, pix = i >> 2;
    const int r = pix / DW_HW, cc = pix - r * DW_HW;
    const int y = h0 + r - 3, x = w0 + cc - 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < DW_HH && y >= 0 && y < H && x >= 0 && x < W)
      v = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y * W + x) * Cs + cs0 + piece * 8));
    *reinterpret_cast<uint4*>(&s_in[pix * DW_PIX + piece * 8]) = v;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp & 3, rbase = (warp >> 2) * 4;      // channel group (8 channels), first of this warp's 4 tile rows
  // B fragments: thread (k0 = 2*(lane%4), n = lane/4) holds B[k0..k0+1][n] and B[k0+8..k0+9][n]; only c == n is non-zero
  const int bn = lane >> 2;
  const bool b_active = (lane & 3) == (bn >> 1);
  const int bch = c0 + cg * 8 + bn;
  uint32_t bfrag[7][4][2];
#pragma unroll
  for (int ky = 0; ky < 7; ++ky)
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      const float wa = b_active ? __ldg(weight + (size_t)(ky * 7 + 2 * kp) * C + bch) : 0.f;
      const float wb = (b_active && kp < 3) ? __ldg(weight + (size_t)(ky * 7 + 2 * kp + 1) * C + bch) : 0.f;
      bfrag[ky][kp][0] = (bn & 1) ? pack16(0.f, wa) : pack16(wa, 0.f);
      bfrag[ky][kp][1] = (bn & 1) ? pack16(0.f, wb) : pack16(wb, 0.f);
    }
  __syncthreads();

  // ldmatrix row address of this lane: matrix j = lane/8 -> (m0 = 8*(j&1), tap t = j>>1), row i = lane%8
  const int lj = lane >> 3, li = lane & 7;
  const uint32_t a_lane = (uint32_t)__cvta_generic_to_shared(s_in) + (uint32_t)(((lj & 1) * 8 + li + (lj >> 1)) * DW_PIX + cg * 8) * 2u;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
#pragma unroll
  for (int hrl = 0; hrl < 10; ++hrl) {                    // halo row rbase + hrl feeds output rows lr with ky = hrl - lr
    uint32_t a[4][4];
#pragma unroll
    for (int kp = 0; kp < 4; ++kp)
      ldmatrix_x4(a[kp], a_lane + (uint32_t)(((rbase + hrl) * DW_HW + 2 * kp) * DW_PIX) * 2u);
#pragma unroll
    for (int lr = 0; lr < 4; ++lr) {
      const int ky = hrl - lr;
      if (ky >= 0 && ky < 7) {
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) mma16816(acc[lr], a[kp], bfrag[ky][kp][0], bfrag[ky][kp][1]);
      }
    }
  }
  // D fragment: acc[.][0,1] -> pixel column lane/4, channels 2*(lane%4)+{0,1}; acc[.][2,3] -> column lane/4 + 8
  const int c = c0 + cg * 8 + (lane & 3) * 2;
  const float* tb = tbias + (size_t)(tbias_stride ? n : 0) * tbias_stride;
  const float b0 = __ldg(tb + c), b1 = __ldg(tb + c + 1);
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int lr = 0; lr < 4; ++lr) {
    const int y = h0 + rbase + lr;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int x = w0 + (lane >> 2) + hf * 8;
      if (y < H && x < W) {
        const float v0 = acc[lr][2 * hf] + b0, v1 = acc[lr][2 * hf + 1] + b1;
        s += v0 + v1;
        q = fmaf(v0, v0, fmaf(v1, v1, q));
        *reinterpret_cast<uint32_t*>(out + (((size_t)n * H + y) * W + x) * C + c) = pack16(v0, v1);
      }
    }
  }
  if (stats != nullptr) {
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) { s_red[warp] = s; s_red[8 + warp] = q; }
    __syncthreads();
    if (threadIdx.x < 32) {
      float ts = 0.f, tq = 0.f;
      if (threadIdx.x == 0)
        for (int i = 0; i < 8; ++i) { ts += s_red[i]; tq += s_red[8 + i]; }
      const int slots = tiles * gridDim.y;
      stats_publish(stats_sample(stats, slots, n), slots, cblk * tiles + tile, ts, tq, stats_inv_count, eps, threadIdx.x);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stem: 7x7 conv, Cin (<=4) -> Cout (multiple of 32, <= 128), fp32 NCHW in, bf16 NHWC out.
// block = 256 threads; tile = 4 rows x 32 cols; thread = (pixel, half of the output channels).
// Weights live in shared memory as [tap*Cin + ci][Cout] fp32 and are read as broadcast float4.
// ---------------------------------------------------------------------------------------------
static constexpr int ST_TH = 4, ST_TW = 32;

template <int CO_PER_THREAD>
__global__ void __launch_bounds__(256)
stem_conv7_kernel(const float* __restrict__ x, int x_batch_mod, const float* __restrict__ weight /* [49*Cin][Cout] */,
                  const float* __restrict__ bias, act_t* __restrict__ out, int N, int Cin, int Cout, int H, int W,
                  int tiles_w, int tiles_per_sample, int total_tiles) {
  extern __shared__ __align__(16) float s_mem[];
  float* s_w = s_mem;                                   // [49*Cin][Cout]
  float* s_in = s_mem + 49 * Cin * Cout;                // [Cin][ST_TH+6][ST_TW+6 (+1 pad)]
  const int IW = ST_TW + 7;
  for (int i = threadIdx.x; i < 49 * Cin * Cout; i += 256) s_w[i] = __ldg(weight + i);
  const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int pr = p / ST_TW, pc = p % ST_TW;
  const int co0 = half * CO_PER_THREAD;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / tiles_per_sample, t = tile % tiles_per_sample;
    const int h0 = (t / tiles_w) * ST_TH, w0 = (t % tiles_w) * ST_TW;
    const int nsrc = x_batch_mod > 0 ? n % x_batch_mod : n;
    __syncthreads();
    for (int i = threadIdx.x; i < Cin * (ST_TH + 6) * (ST_TW + 6); i += 256) {
      const int cc = i % (ST_TW + 6), r = (i / (ST_TW + 6)) % (ST_TH + 6), ci = i / ((ST_TW + 6) * (ST_TH + 6));
      const int y = h0 + r - 3, xx = w0 + cc - 3;
      float v = 0.f;
      if (y >= 0 && y < H && xx >= 0 && xx < W) v = __ldg(x + (((size_t)nsrc * Cin + ci) * H + y) * W + xx);
      s_in[(ci * (ST_TH + 6) + r) * IW + cc] = v;
    }
    __syncthreads();
    float2 acc2[CO_PER_THREAD / 2];
#pragma unroll
    for (int j = 0; j < CO_PER_THREAD / 2; ++j) acc2[j] = make_float2(__ldg(bias + co0 + 2 * j), __ldg(bias + co0 + 2 * j + 1));
    for (int ky = 0; ky < 7; ++ky)
      for (int kx = 0; kx < 7; ++kx)
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = s_in[(ci * (ST_TH + 6) + pr + ky) * IW + pc + kx];
          const float2 vv = make_float2(v, v);
          const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 7 + kx) * Cin + ci) * Cout + co0);
#pragma unroll
          for (int j = 0; j < CO_PER_THREAD / 4; ++j) {
            const float4 w4 = wr[j];
            ffma2(acc2[2 * j], vv, make_float2(w4.x, w4.y));
            ffma2(acc2[2 * j + 1], vv, make_float2(w4.z, w4.w));
          }
        }
    float acc[CO_PER_THREAD];
#pragma unroll
    for (int j = 0; j < CO_PER_THREAD / 2; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
    const int y = h0 + pr, xx = w0 + pc;
    if (y < H && xx < W) {
      uint4* op = reinterpret_cast<uint4*>(out + (((size_t)n * H + y) * W + xx) * Cout + co0);
#pragma unroll
      for (int j = 0; j < CO_PER_THREAD / 8; ++j)
        op[j] = make_uint4(pack16(acc[8 * j], acc[8 * j + 1]), pack16(acc[8 * j + 2], acc[8 * j + 3]),
                           pack16(acc[8 * j + 4], acc[8 * j + 5]), pack16(acc[8 * j + 6], acc[8 * j + 7]));
    }
  }
}

}  // namespace ds

using namespace ds;

extern "C" {

/* Depthwise 7x7 + time bias + GroupNorm partials.  d_weight fp32 [49][C0+C1] (tap-major);
   d_tbias fp32 [N][tbias_stride] (or one row when tbias_stride == 0) holding ds_conv.bias + mlp(time_emb).
   d_stats (nullable): statistics buffer float2 [N][2 + ds_dwconv7_stats_slots(...)], zero-initialised once (see ds_conv_gemm). */
int ds_dwconv7(const void* d_src0, const void* d_src1, int C0, int C1, int src_batch_mod, const float* d_weight, const float* d_tbias,
               long long tbias_stride, void* d_out, void* d_stats, float eps, int N, int H, int W, void* stream) {
  DS_REQUIRE(d_src0 && d_weight && d_tbias && d_out && N > 0 && H > 0 && W > 0, "ds_dwconv7: bad arguments");
  DS_REQUIRE(C0 > 0 && C0 % DW_CB == 0 && C1 >= 0 && C1 % DW_CB == 0 && (C1 == 0 || d_src1), "ds_dwconv7: C0=%d C1=%d must be multiples of 32", C0, C1);
  const int tiles_w = (W + DW_TW - 1) / DW_TW, tiles_h = (H + DW_TH - 1) / DW_TH;
  const int tiles = tiles_w * tiles_h;
  DS_REQUIRE(N <= 65535 && (C0 + C1) / DW_CB <= 65535, "ds_dwconv7: grid too large");
  dwconv7_kernel<<<dim3(tiles, (C0 + C1) / DW_CB, N), 256, 0, (cudaStream_t)stream>>>(
      (const act_t*)d_src0, (const act_t*)d_src1, C0, C1, src_batch_mod, d_weight, d_tbias, tbias_stride,
      (act_t*)d_out, (float2*)d_stats, 1.0f / ((float)H * (float)W * (float)(C0 + C1)), eps, H, W, tiles_w, tiles);
  DS_CHECK_CUDA(cudaGetLastError());
  return DS_OK;
}

int ds_dwconv7_stats_slots(int C, int H, int W) {
  return ((W + DW_TW - 1) / DW_TW) * ((H + DW_TH - 1) / DW_TH) * (C / DW_CB);
}

/* Stem 7x7 conv (init_conv).  d_x fp32 NCHW [x_batch_mod or N, Cin, H, W]; d_weight fp32 [49*Cin][Cout]
   ((ky,kx,ci)-major); output bf16 NHWC [N,H,W,Cout]. */
int ds_stem_conv7(const float* d_x, int x_batch_mod, const float* d_weight, const float* d_bias, void* d_out, int N, int Cin, int Cout,
                  int H, int W, void* stream) {
  DS_REQUIRE(d_x && d_weight && d_bias && d_out && N > 0 && Cin > 0 && Cin <= 4 && H > 0 && W > 0, "ds_stem_conv7: bad arguments");
  DS_REQUIRE(Cout == 96 || Cout == 64 || Cout == 32 || Cout == 128, "ds_stem_conv7: Cout=%d unsupported (32/64/96/128)", Cout);
  const int tiles_w = (W + ST_TW - 1) / ST_TW, tiles_h = (H + ST_TH - 1) / ST_TH;
  const int tps = tiles_w * tiles_h, total = tps * N;
  const size_t smem = (size_t)(49 * Cin * Cout + Cin * (ST_TH + 6) * (ST_TW + 7)) * sizeof(float);
  int grid = total < 2 * num_sms() ? total : 2 * num_sms();
#define LAUNCH_STEM(CPT)                                                                                                    \
  DS_CHECK_CUDA(cudaFuncSetAttribute(stem_conv7_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
  stem_conv7_kernel<CPT><<<grid, 256, smem, (cudaStream_t)stream>>>(d_x, x_batch_mod, d_weight, d_bias, (act_t*)d_out, N, \
                                                                     Cin, Cout, H, W, tiles_w, tps, total);
  if (Cout == 96) { LAUNCH_STEM(48) } else if (Cout == 64) { LAUNCH_STEM(32) } else if (Cout == 32) { LAUNCH_STEM(16) } else { LAUNCH_STEM(64) }
#undef LAUNCH_STEM
  DS_CHECK_CUDA(cudaGetLastError());
  return DS_OK;
}

}  // extern "C"
