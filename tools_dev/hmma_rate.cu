// Micro-benchmark: issue rate of legacy mma.sync.m16n8k16 (f16 -> f32) and ldmatrix.x4 per SM on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/hmma_rate tools_dev/hmma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ACCS>
__global__ void hmma_loop(float* out, int iters) {
  float c[ACCS][4];
  for (int i = 0; i < ACCS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACCS; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int i = 0; i < ACCS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ldsm_loop(float* out, int iters) {
  __shared__ __align__(16) uint8_t buf[32 * 1072];
  for (int i = threadIdx.x; i < 32 * 1072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(buf)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(buf) + warp * 4 * 1072 + ((lane & 7) + ((lane >> 3) & 1) * 8) * 48 + (lane >> 4) * 16;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      uint32_t r0, r1, r2, r3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + dy * 48));
      acc += r0 ^ r1 ^ r2 ^ r3;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      hmma_loop<8><<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas_per_sm = (double)iters * 8 * warps;
    printf("hmma m16n8k16: %2d warps/SM: %.3f ms -> %.2f ns per MMA per SM (%.1f cycles at %.2f GHz nominal) = %.1f dense TFLOP/s\n", warps, ms,
           ms * 1e6 / mmas_per_sm, ms * 1e-3 * clk * 1e3 / mmas_per_sm, clk * 1e-6, 148 * mmas_per_sm * 4096 / (ms * 1e-3) / 1e12);
  }
  for (int warps = 8; warps <= 16; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      ldsm_loop<<<148, warps * 32 > 256 ? 256 : warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)iters * 7 * 8;
    printf("ldmatrix.x4 (8 warps): %.3f ms -> %.1f cycles per ldmatrix.x4 per SM (512 B each)\n", ms, ms * 1e-3 * clk * 1e3 / n);
    break;
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
