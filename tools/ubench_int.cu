// Micro-benchmark of the integer instructions of the Sobel kernel (analyze_sobel_tma.cuh) on one B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_int tools/ubench_int.cu && tools/ubench_int
// Prints warp-instructions per cycle per SM sub-partition for 1..8 resident warps per sub-partition.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kIters = 2048;
constexpr int kAcc = 12;

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int OP>
__global__ void k(uint32_t* out, uint32_t a, uint32_t b, long long* cyc) {
  uint32_t acc[kAcc], p[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) { acc[i] = a * (i + threadIdx.x); p[i] = b + i * 0x01010101u; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
    if (OP == 0) {  // IDP.4A
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = (uint32_t)dp4a_us(p[i], b, (int)acc[i]);
    } else if (OP == 1) {  // VABSDIFF with accumulate (__sad)
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __sad((int)p[i], (int)b, acc[i]);
    } else if (OP == 2) {  // PRMT
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __byte_perm(acc[i], p[i], 0x5140);
    } else if (OP == 3) {  // IADD3
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = acc[i] + p[i] + b;
    } else if (OP == 4) {  // VIADD.16x2
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __vadd2(acc[i], p[i]);
    } else if (OP == 5) {  // VIMNMX.U16x2
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __vmaxu2(acc[i], p[i]);
    } else if (OP == 6) {  // the Sobel mix per channel and column pair: 4 IDP + 2 SAD + 2 add + 2 SAD, plus 7/3 PRMT
#pragma unroll
      for (int i = 0; i < kAcc; i += 4) {
        const uint32_t q = __byte_perm(acc[i], p[i], 0x5410 + it);
        const uint32_t q2 = __byte_perm(acc[i + 1], p[i + 1], 0x7632 + it);
        const int h0 = dp4a_us(q, 0x00010201u, 0), h1 = dp4a_us(q, 0x01020100u, 0), g0 = dp4a_us(q2, 0x000100FFu, 0), g1 = dp4a_us(q2, 0x0100FF00u, 0);
        acc[i] = __sad(h0, (int)p[i], acc[i]);
        acc[i + 1] = __sad(h1, (int)p[i + 1], acc[i + 1]);
        acc[i + 2] = __sad(g0 * 2 + (int)(p[i + 2] + p[i]), 0, acc[i + 2]);
        acc[i + 3] = __sad(g1 * 2 + (int)(p[i + 3] + p[i + 1]), 0, acc[i + 3]);
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  printf("%-34s", name);
  for (int wps = 1; wps <= 8; wps *= 2) {
    const int threads = wps * 4 * 32;
    k<OP><<<148, threads>>>(out, 3u, 0x01020304u, cyc);
    cudaDeviceSynchronize();
    k<OP><<<148, threads>>>(out, 3u, 0x01020304u, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += (double)h[i];
    c /= 148;
    printf("  %dw: %.3f", wps, (double)kIters * per_iter * wps / c);
  }
  printf("   (warp-instr / clk / sub-partition)\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("IDP.4A.U8.S8", kAcc);
  run<1>("VABSDIFF (+acc)", kAcc);
  run<2>("PRMT", kAcc);
  run<3>("IADD3", kAcc);
  run<4>("VIADD.16x2", kAcc);
  run<5>("VIMNMX.U16x2", kAcc);
  run<6>("Sobel mix (14 instr per group)", 3 * 14);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
