// Shared helpers for the diffusynth_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace ds {

// ---- error plumbing (C-ABI returns int; message kept per host thread) ------------------------
enum : int { DS_OK = 0, DS_ERR_INVALID = -1, DS_ERR_CUDA = -2, DS_ERR_NCCL = -3, DS_ERR_UNSUPPORTED = -4 };
void set_error(const char* fmt, ...);
const char* get_error();

#define DS_CHECK_CUDA(expr)                                                               \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ds::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ds::DS_ERR_CUDA;                                                             \
    }                                                                                     \
  } while (0)

#define DS_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ds::set_error(__VA_ARGS__);             \
      return ds::DS_ERR_INVALID;              \
    }                                         \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Reduce the per-tile (sum, sumsq) partials of one sample to (mean, rstd).  Called by a full warp;
// every lane returns the result.  Deterministic (fixed order), accumulates in double.
__device__ __forceinline__ float2 reduce_stats_warp(const float2* __restrict__ part, int slots, float inv_count,
                                                    float eps, int lane) {
  double s = 0.0, q = 0.0;
  for (int i = lane; i < slots; i += 32) {
    float2 v = __ldg(part + i);
    s += (double)v.x;
    q += (double)v.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  double mean = s * (double)inv_count;
  double var = q * (double)inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  return make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

}  // namespace ds
