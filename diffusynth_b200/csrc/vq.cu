// Nearest-codebook-entry quantiser: VectorQuantizerEMA.forward, eval branch (model/VQGAN.py:98-146).
// Distances are evaluated in fp32 in exactly the reference's expanded form and operation order
//   (sum_d x_d^2 + sum_d e_d^2) - 2 * (x . e)        (ascending-d sums, FMA chain for the dot product)
// so that argmin (first minimum) is bit-exact against the oracle; the [N, 8192] distance / one-hot
// matrices of the reference are never materialised.  Output = x + (q - x) (straight-through line :134).
#include "common.cuh"
#include "../../include/diffusynth_b200.h"

namespace ds {

static constexpr int VQ_PPT = 4;   // positions per thread

// dynamic smem: float4 code[K]; float se[K]
__global__ void __launch_bounds__(256)
vq_kernel(const float* __restrict__ x /* [B,4,H,W] */, const float4* __restrict__ codebook, int K, float* __restrict__ out,
          long long* __restrict__ idx_out, long long hw, long long total /* B*hw */) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t vq_smem[];
  float4* s_code = reinterpret_cast<float4*>(vq_smem);
  float* s_se = reinterpret_cast<float*>(s_code + K);
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const float4 e = __ldg(codebook + i);
    s_code[i] = e;
    s_se[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y)), __fmul_rn(e.z, e.z)), __fmul_rn(e.w, e.w));
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = blockIdx.x * (long long)blockDim.x + threadIdx.x; base < total; base += stride * VQ_PPT) {
    float xv[VQ_PPT][4], sx[VQ_PPT], best[VQ_PPT];
    int bi[VQ_PPT];
    long long pos[VQ_PPT];
#pragma unroll
    for (int p = 0; p < VQ_PPT; ++p) {
      pos[p] = base + p * stride;
      const bool ok = pos[p] < total;
      const long long b = ok ? pos[p] / hw : 0, r = ok ? pos[p] % hw : 0;
#pragma unroll
      for (int d = 0; d < 4; ++d) xv[p][d] = ok ? __ldg(x + (b * 4 + d) * hw + r) : 0.f;
      sx[p] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xv[p][0], xv[p][0]), __fmul_rn(xv[p][1], xv[p][1])), __fmul_rn(xv[p][2], xv[p][2])),
                        __fmul_rn(xv[p][3], xv[p][3]));
      best[p] = INFINITY;
      bi[p] = 0;
    }
#pragma unroll 2
    for (int j = 0; j < K; ++j) {
      const float4 e = s_code[j];
      const float se = s_se[j];
#pragma unroll
      for (int p = 0; p < VQ_PPT; ++p) {
        float dot = __fmul_rn(xv[p][0], e.x);
        dot = __fmaf_rn(xv[p][1], e.y, dot);
        dot = __fmaf_rn(xv[p][2], e.z, dot);
        dot = __fmaf_rn(xv[p][3], e.w, dot);
        const float dist = __fsub_rn(__fadd_rn(sx[p], se), __fmul_rn(2.0f, dot));
        if (dist < best[p]) { best[p] = dist; bi[p] = j; }   // strict '<' keeps the first minimum
      }
    }
#pragma unroll
    for (int p = 0; p < VQ_PPT; ++p) {
      if (pos[p] < total) {
        const long long b = pos[p] / hw, r = pos[p] % hw;
        const float4 e = s_code[bi[p]];
        const float q[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int d = 0; d < 4; ++d) out[(b * 4 + d) * hw + r] = __fadd_rn(xv[p][d], __fsub_rn(q[d], xv[p][d]));
        if (idx_out) idx_out[pos[p]] = bi[p];
      }
    }
  }
}

}  // namespace ds

using namespace ds;

extern "C" {

/* d_x / d_out fp32 NCHW [B, 4, H, W]; d_codebook fp32 [K, 4]; d_idx (nullable) int64 [B*H*W] in the
   reference's flattened BHWC order.  embedding_dim is fixed at 4 (deployed VQGAN, app.py:32). */
int ds_vq_quantize(const float* d_x, const float* d_codebook, int K, float* d_out, long long* d_idx, int B, long long hw, void* stream) {
  DS_REQUIRE(d_x && d_codebook && d_out && B > 0 && hw > 0 && K > 0, "ds_vq_quantize: bad arguments");
  const size_t smem = (size_t)K * (sizeof(float4) + sizeof(float));
  DS_REQUIRE(smem <= 200 * 1024, "ds_vq_quantize: codebook of %d entries does not fit in shared memory", K);
  DS_CHECK_CUDA(cudaFuncSetAttribute(vq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)B * hw;
  long long want = (total + 256 * VQ_PPT - 1) / (256 * VQ_PPT);
  int grid = (int)(want < num_sms() ? want : num_sms());
  DS_CHECK_CUDA(launch_pdl(vq_kernel, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, d_x, (const float4*)d_codebook, K, d_out, d_idx, hw, total));
  DS_CHECK_CUDA(cudaGetLastError());
  return DS_OK;
}

}  // extern "C"
