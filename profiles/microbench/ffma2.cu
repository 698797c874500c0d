// Microbenchmark: plain FFMA vs packed FFMA2 (sm_100) issue throughput on B200.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, aa.x, bb.x);
        acc[i].y = fmaf(acc[i].y, aa.y, bb.y);
      } else {
        acc[i] = __ffma2_rn(acc[i], aa, bb);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
      else k<1><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double flop = 2.0 * 16 * (double)iters * blocks * threads;
      if (rep) printf("%s: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, flop / ms * 1e-9);
    }
  }
  return 0;
}
