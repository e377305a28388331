// Per-SM throughput of the instructions the element-wise phases of the scan kernels are made of (sm_100a).
// One CTA of 256 threads (8 warps, 2 per scheduler -- the occupancy of the scan kernels) per SM, ILP 8 chains per thread.
// Prints lane-operations per clock per SM for each instruction.   nvcc -arch=sm_100a -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define N_IT 2048
template <int OP>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc, float seed) {
  float a[8];
  uint32_t u[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; u[i] = threadIdx.x * 8 + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
      if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(u[i]); }
      if (OP == 4) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(u[i]); }
      if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 6) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));
      if (OP == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
      if (OP == 8) a[i] = a[i] + 1.5f;
      if (OP == 9) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 10) asm volatile("shl.b32 %0, %0, 16;" : "+r"(u[i]));
      if (OP == 11) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]) + (float)(p[i] & 0xffff);
  out[blockIdx.x * 256 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int lanes_per_inst) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 256>>>(out, cyc, 1.0f);
  k<OP><<<148, 256>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double inst = 8.0 * N_IT * 8;  // warp instructions per SM
  printf("%-28s %8.1f cycles/warp-inst/SM^-1 -> %6.1f warp-inst/clk/SM = %6.1f lane-ops/clk/SM (x%d elems)\n", name, c / inst, inst / c, inst / c * 32 * lanes_per_inst, lanes_per_inst);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA", 1); run<1>("FFMA2 (fma.rn.f32x2)", 2); run<11>("FMUL2", 2); run<7>("FADD2", 2); run<8>("FADD", 1); run<2>("MUFU.TANH", 1); run<9>("MUFU.EX2", 1);
  run<3>("F2FP bf16x2 (cvt.rn)", 2); run<4>("F2FP f16x2 (cvt.rn)", 2); run<5>("PRMT", 1); run<10>("SHL", 1); run<6>("SHFL.BFLY", 1);
  return 0;
}
