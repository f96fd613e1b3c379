// Experiment: can a K-major swizzled UMMA operand start at an arbitrary ROW offset inside a TMA-written buffer
// (needed to reuse one activation halo tile for all taps of a 3x3 conv)?  For every row shift r and every
// candidate descriptor base_offset, D = A[r : r+128] * B^T is compared with the host result.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_shift_test tools_dev/umma_shift_test.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BK>
__global__ void __launch_bounds__(128, 1) shift_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                       float* out, int rows_a, int shift, int base_offset, int N) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // rows_a x BK*2 bytes
  uint8_t* sB = smem + 32768;               // N x BK*2
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint64_t* bar2 = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t bytes = rows_a * BK * 2 + N * BK * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"(&mapA), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(&mapB), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr uint64_t row_bytes = BK * 2;
    constexpr uint64_t sbo = (8 * row_bytes) >> 4;
    constexpr uint64_t layout = (BK == 64) ? 2ull : 4ull;
    const uint32_t a_addr = smem_u32(sA) + shift * row_bytes;
    const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) |
                           ((uint64_t)(base_offset & 7) << 49) | (layout << 61);
    const uint64_t bdesc = (uint64_t)((smem_u32(sB) & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
    for (int k = 0; k < BK / 16; ++k) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(k) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar2)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BK>
void run() {
  const int RA = 160, N = 64;
  std::vector<__half> A(RA * BK), B(N * BK);
  std::vector<float> Af(RA * BK), Bf(N * BK);
  srand(1);
  for (int i = 0; i < RA * BK; ++i) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2half(v); Af[i] = v; }
  for (int i = 0; i < N * BK; ++i) { float v = (rand() % 13 - 6) / 4.0f; B[i] = __float2half(v); Bf[i] = v; }
  __half *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dO, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap mA, mB;
  const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  { cuuint64_t d[2] = {(cuuint64_t)BK, (cuuint64_t)RA}, st[1] = {(cuuint64_t)BK * 2}; cuuint32_t bx[2] = {(cuuint32_t)BK, (cuuint32_t)RA}, es[2] = {1, 1};
    if (enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode A failed\n"); exit(1); } }
  { cuuint64_t d[2] = {(cuuint64_t)BK, (cuuint64_t)N}, st[1] = {(cuuint64_t)BK * 2}; cuuint32_t bx[2] = {(cuuint32_t)BK, (cuuint32_t)N}, es[2] = {1, 1};
    if (enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode B failed\n"); exit(1); } }
  CK(cudaFuncSetAttribute(shift_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  std::vector<float> O(128 * N);
  printf("BK=%d (swizzle %dB): rows = shift, columns = base_offset 0..7; '#' = exact match\n", BK, BK * 2);
  for (int shift = 0; shift <= 20; ++shift) {
    printf("  shift %2d: ", shift);
    for (int bo = 0; bo < 8; ++bo) {
      CK(cudaMemset(dO, 0, 128 * N * 4));
      shift_kernel<BK><<<1, 128, 65536>>>(mA, mB, dO, RA, shift, bo, N);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("E(%s) ", cudaGetErrorString(e)); exit(2); }
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      double err = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < BK; ++k) ref += Af[(i + shift) * BK + k] * Bf[n * BK + k];
          err += fabs(ref - O[i * N + n]);
        }
      printf("%c", err < 1e-3 ? '#' : '.');
    }
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main() {
  run<64>();
  run<32>();
  return 0;
}
