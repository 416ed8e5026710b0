// mufu_probe.cu -- MUFU.EX2 throughput per SM: f32 vs packed f16x2 / bf16x2 (two exponentials per instruction?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu_probe tools/probes/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  uint32_t a[8];
  float f[8];
  for (int i = 0; i < 8; ++i) {
    f[i] = -0.001f * (threadIdx.x + i + 1);
    a[i] = 0xb800b800u + threadIdx.x + i;  // two small negative halves
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  float* out;
  long long* clk;
  const int threads = 512, blocks = 148, iters = 4096;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaMalloc(&clk, sizeof(long long) * blocks);
  k<MODE><<<blocks, threads>>>(out, clk, iters);
  k<MODE><<<blocks, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const double instr = (double)threads * iters * 8;
  printf("PROBE mufu %-22s: %.2f instr/clk/SM = %.2f exponentials/clk/SM (%s)\n", name, instr / h[0],
         per_instr * instr / h[0], cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  return 0;
}
