// Micro-benchmark: issue rate of MUFU.EX2 (f32 and packed bf16x2) and of a degree-3 FMA-pipe exp2 on one SM, per warp-instruction,
// with 1, 2 and 4 warps per SM sub-partition.  nvcc -arch=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (MODE == 1) { unsigned u = __float_as_uint(x[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      else {   // 2^f on the FMA pipe, f in [-1, 0]: degree-3 polynomial + exponent insertion (6 ALU ops)
        float f = x[i];
        float fl = floorf(f);
        float r = f - fl;
        float p = fmaf(fmaf(fmaf(0.0555054f, r, 0.2402265f), r, 0.6931472f), r, 1.0f);
        x[i] = __uint_as_float(__float_as_uint(p) + (static_cast<int>(fl) << 23)) - 1.5f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  const char* names[3] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2", "fma-pipe poly exp2"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
        else if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
        else k<2><<<1, warps * 32>>>(out, cyc, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double per = double(h) / (double(iters) * 16);
      printf("%-24s %2d warps/SM (%d per sub-partition): %.2f cycles per warp-instruction -> %.1f results/clk/SM\n", names[mode], warps, warps / 4, per,
             (mode == 1 ? 64.0 : 32.0) * warps / per);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
