/* A host with no Python in it: loads a ConditionedUnet by its reference state_dict names through the module-level C ABI and runs
   one forward (tests/test_gpu_abi_engine.py::test_unet_forward_from_plain_c compares the result with the Python host's).
   usage: abi_unet_forward weights.bin inputs.bin eps.bin
   weights.bin: records {int32 name_len, name, int32 ndim, int64 shape[ndim], float data[]}, terminated by name_len 0
   inputs.bin:  int32 N, H, W; float x[N*4*H*W]; int64 t[N]; float cond[N*64] */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "diffusynth_b200.h"

#define CHECK(e) do { int rc_ = (e); if (rc_) { fprintf(stderr, "%s failed (%d): %s\n", #e, rc_, ds_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  ds_unet_config cfg = {0};
  cfg.in_dim = 4; cfg.n_levels = 3; cfg.with_time_emb = 1; cfg.use_convnext = 1; cfg.label_emb_dim = 64;
  const int dd[3] = {32, 32, 64}, ud[3] = {64, 64, 32};
  for (int i = 0; i < 3; ++i) { cfg.down_dims[i] = dd[i]; cfg.up_dims[i] = ud[i]; }
  ds_unet* net = NULL;
  CHECK(ds_unet_create(&cfg, &net));
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  for (;;) {
    int len = 0, ndim = 0;
    char name[256];
    long long shape[8], n = 1;
    if (fread(&len, 4, 1, f) != 1 || len == 0) break;
    if (len > 255 || fread(name, 1, len, f) != (size_t)len || fread(&ndim, 4, 1, f) != 1 || fread(shape, 8, ndim, f) != (size_t)ndim) return 4;
    name[len] = 0;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    float* data = (float*)malloc(n * sizeof(float));
    if (fread(data, sizeof(float), n, f) != (size_t)n) return 5;
    CHECK(ds_unet_load(net, name, data, shape, ndim));
    free(data);
  }
  fclose(f);
  CHECK(ds_unet_finalize(net));
  int dims[3];
  f = fopen(argv[2], "rb");
  if (!f || fread(dims, 4, 3, f) != 3) return 6;
  const int N = dims[0], H = dims[1], W = dims[2];
  const size_t nx = (size_t)N * 4 * H * W, nc = (size_t)N * 64;
  float* x = (float*)malloc(nx * 4); long long* t = (long long*)malloc(N * 8); float* c = (float*)malloc(nc * 4);
  if (fread(x, 4, nx, f) != nx || fread(t, 8, N, f) != (size_t)N || fread(c, 4, nc, f) != nc) return 7;
  fclose(f);
  float *dx, *dc, *de; long long* dt;
  cudaMalloc((void**)&dx, nx * 4); cudaMalloc((void**)&dc, nc * 4); cudaMalloc((void**)&de, nx * 4); cudaMalloc((void**)&dt, N * 8);
  cudaMemcpy(dx, x, nx * 4, cudaMemcpyHostToDevice); cudaMemcpy(dc, c, nc * 4, cudaMemcpyHostToDevice); cudaMemcpy(dt, t, N * 8, cudaMemcpyHostToDevice);
  CHECK(ds_unet_forward(net, dx, dt, dc, de, N, H, W, NULL));
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "CUDA: %s\n", cudaGetErrorString(cudaGetLastError())); return 8; }
  cudaMemcpy(x, de, nx * 4, cudaMemcpyDeviceToHost);
  f = fopen(argv[3], "wb");
  fwrite(x, 4, nx, f);
  fclose(f);
  ds_unet_destroy(net);
  printf("ok\n");
  return 0;
}
