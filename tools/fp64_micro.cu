// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_micro tools/fp64_micro.cu
// Prints one JSON line; the result is the measured FP64 denominator context.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
  double av = a + threadIdx.x * 1e-12, bv = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(av), "d"(bv));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
double time_ms(F f) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  for (int wps = 1; wps <= 4; wps *= 2) {  // CTAs of 256 thr per SM: 1,2,4 -> 8,16,32 warps/SM
    int grid = sms * wps;
    double ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    double tf = 2.0 * 16 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dfma_tflops_%dcta\": %.2f", wps, tf);
    ms = time_ms([&] { dmma_kernel<16><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
    tf = 2.0 * 256 * 16 * (iters / 4) * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma_tflops_%dcta\": %.2f", wps, tf);
  }
  {  // single warp per SMSP latency-bound probe: 1 accumulator chain
    double ms = time_ms([&] { dmma_kernel<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
    double ns_per = ms * 1e6 / iters;
    printf(", \"dmma_dep_chain_ns\": %.2f", ns_per);
    ms = time_ms([&] { dfma_kernel<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_dep_chain_ns\": %.2f", ms * 1e6 / iters);
  }
  {  // how far does ONE warp per SMSP get with many independent accumulators? (half-tile GEMM CTAs have one consumer warp
     // per SMSP: if a single warp cannot saturate the DMMA pipe, two co-resident CTAs cannot hide each other's idle phases)
    for (int warps = 1; warps <= 4; warps *= 2) {
      double ms = time_ms([&] { dmma_kernel<32><<<sms, 128 * warps>>>(out, iters / 8, 1.0000001, 1e-9); });
      double tf = 2.0 * 256 * 32 * (iters / 8) * 4.0 * warps * sms / (ms * 1e-3) / 1e12;
      printf(", \"dmma_tflops_%dwarp_per_smsp_32acc\": %.2f", warps, tf);
    }
    double ms = time_ms([&] { dmma_kernel<8><<<sms, 128>>>(out, iters / 2, 1.0000001, 1e-9); });
    printf(", \"dmma_tflops_1warp_per_smsp_8acc\": %.2f", 2.0 * 256 * 8 * (iters / 2) * 4.0 * sms / (ms * 1e-3) / 1e12);
  }
  printf("}\n");
  CK(cudaGetLastError());
  return 0;
}
