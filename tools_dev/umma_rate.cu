// Micro-benchmark: how fast does an SM retire back-to-back tcgen05.mma.kind::f16 (M = 128 or 64, K = 16, both operands from shared
// memory) for small N, issued by 1, 2 or 4 threads (one per warp, distinct accumulator columns)?  Decides whether a Toeplitz
// depthwise convolution (N = 8..32 per instruction) can beat the legacy mma.sync path (2.1 cycles per m16n8k16 = 975 MAC/clk/SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate tools_dev/umma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int M, int N, int issuers, int iters, int same_a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;      // halves = 1.0
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (warp < issuers && lane == 0) {
    // K-major SWIZZLE_64B operands (rows of 32 halves): A tile = 128 rows x 64 B, a fresh tile every instruction unless same_a
    const uint64_t sbo = (8 * 64) >> 4;
    const uint64_t dbase = (1ull << 16) | (sbo << 32) | (1ull << 46) | (4ull << 61);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem) + warp * 32768, b0 = smem_u32(smem) + 131072 + warp * 4096;
    const uint32_t d = tmem + warp * 128;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t aa = same_a ? a0 : a0 + (u & 3) * 8192 + (u >> 2) * 32;
        const uint64_t adesc = dbase | (uint64_t)((aa & 0x3FFFFu) >> 4), bdesc = dbase | (uint64_t)(((b0 + (u & 1) * 2048) & 0x3FFFFu) >> 4);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + warp)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar + warp)) : "memory");
    t1 = clock64();
    out[blockIdx.x * 4 + warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 148 * 4 * 8));
  const int smem = 164 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;
  printf("cycles per tcgen05.mma (K = 16, SS), per issuing thread and per SM; all 148 SMs busy\n");
  for (int M : {128, 64})
    for (int N : {8, 16, 32, 64, 128, 256}) {
      if (M == 128 && N % 16) continue;
      for (int issuers : {1, 2, 4}) {
        if (issuers * N > 512 && N > 128) continue;
        for (int same_a : {0, 1}) {
          CK(cudaMemset(d, 0, 148 * 4 * 8));
          rate_kernel<<<148, 128, smem>>>(d, M, N, issuers, iters, same_a);
          CK(cudaDeviceSynchronize());
          long long h[148 * 4];
          CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
          double worst = 0;
          for (int b = 0; b < 148; ++b) for (int w = 0; w < issuers; ++w) if (h[b * 4 + w] > worst) worst = (double)h[b * 4 + w];
          const double per_thread = worst / (iters * 8.0), per_sm = per_thread / issuers;
          printf("M=%3d N=%3d issuers=%d %s: %.1f cyc/MMA/thread, %.2f cyc/MMA/SM, %.0f MAC/clk/SM\n", M, N, issuers, same_a ? "same A " : "fresh A",
                 per_thread, per_sm, (double)M * N * 16 / per_sm);
        }
      }
    }
  return 0;
}
