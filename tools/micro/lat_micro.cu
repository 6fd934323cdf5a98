// Single-warp latency / issue-rate probes for the serial 32x32 potf2 of k_diag_factor (tools/micro/lat_micro.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_micro lat_micro.cu ; prints cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__global__ void probe(double* out, long long* cyc, double seed) {
  const int lane = threadIdx.x;
  __shared__ double sm[64];
  sm[lane] = seed + lane; sm[lane + 32] = seed - lane;
  __syncwarp();
  double x = seed + 1.0 + lane * 1e-3;
  long long t0, t1;
  // (0) dependent rsqrt chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64(); if (lane == 0) cyc[0] = (t1 - t0) / 256;
  // (1) dependent DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; ++i) x = fma(x, 0.999999, 1e-9);
  t1 = clock64(); if (lane == 0) cyc[1] = (t1 - t0) * 100 / 1024;
  // (2) dependent 64-bit shuffle chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; ++i) x = __shfl_sync(FULL, x, (lane + 1) & 31);
  t1 = clock64(); if (lane == 0) cyc[2] = (t1 - t0) * 100 / 1024;
  // (3) independent 64-bit shuffles (16 in flight) + fma, like the trailing update
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = x + k;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(-x, __shfl_sync(FULL, x, k), a[k]);
  }
  t1 = clock64(); if (lane == 0) cyc[3] = (t1 - t0) * 100 / (64 * 16);
  // (4) same through shared memory: one store + 16 broadcast loads + fma
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    sm[lane] = x; __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(-x, sm[k], a[k]);
    __syncwarp();
  }
  t1 = clock64(); if (lane == 0) cyc[4] = (t1 - t0) * 100 / (64 * 16);
  // (5) the potf2 column step as written: shfl pivot -> rsqrt -> mul -> shfl -> fma (dependent chain only)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    double piv = __shfl_sync(FULL, x, i & 31);
    if (!(piv > 0.0)) piv = 1.0;
    const double inv = rsqrt(piv);
    const double l = x * inv;
    x = fma(-l, __shfl_sync(FULL, l, (i + 1) & 31), x + 3.0);
  }
  t1 = clock64(); if (lane == 0) cyc[5] = (t1 - t0) / 256;
  // (6) dependent sqrt + divide (the alternative to rsqrt)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) x = 1.0 / sqrt(x) + 1.5;
  t1 = clock64(); if (lane == 0) cyc[6] = (t1 - t0) / 256;
  // (7) fast reciprocal square root: float seed + two Newton steps in double
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    double y = (double)rsqrtf((float)x);
    y = y * fma(-0.5 * x * y, y, 1.5);
    y = y * fma(-0.5 * x * y, y, 1.5);
    x = y + 1.5;
  }
  t1 = clock64(); if (lane == 0) cyc[7] = (t1 - t0) / 256;
  double s = x;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  out[lane] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, 8 * sizeof(long long));
  probe<<<1, 32>>>(out, cyc, 2.0); probe<<<1, 32>>>(out, cyc, 2.0);
  long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"err\": \"%s\", \"rsqrt_dep_cycles\": %lld, \"dfma_dep_cycles_x100\": %lld, \"shfl64_dep_cycles_x100\": %lld, "
         "\"shfl64_fma_indep_cycles_x100\": %lld, \"smem_bcast_fma_cycles_x100\": %lld, \"potf2_chain_cycles\": %lld, "
         "\"sqrt_div_dep_cycles\": %lld, \"rsqrtf_newton2_dep_cycles\": %lld}\n",
         cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  return 0;
}
