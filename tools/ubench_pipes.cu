// Micro-benchmark of the FP32 issue rates that bound the resample kernels (exact mode = FMUL + FADD per tap,
// fused mode = FFMA / FFMA2), run on one B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pipes tools/ubench_pipes.cu && tools/ubench_pipes
// Prints warp-instructions per cycle per SM sub-partition for 1..8 resident warps per sub-partition.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kIters = 2048;
constexpr int kAcc = 12;  // independent accumulators per thread (what the resample walk holds)

template <int OP>
__global__ void k(float* out, float a, float b, long long* cyc, float rt_one, float rt_negzero) {
  float acc[kAcc];
  float p[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) { acc[i] = a * (i + threadIdx.x); p[i] = b + i; }
  float2 acc2[kAcc / 2], p2[kAcc / 2];
#pragma unroll
  for (int i = 0; i < kAcc / 2; ++i) { acc2[i] = make_float2(acc[2 * i], acc[2 * i + 1]); p2[i] = make_float2(p[2 * i], p[2 * i + 1]); }
  float2 w2 = make_float2(b, b);
  uint32_t word = __float_as_uint(a) + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
    if (OP == 0) {  // FFMA, 3 register operands
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = fmaf(p[i], b, acc[i]);
    } else if (OP == 1) {  // FMUL + FADD (exact tap)
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(p[i], b));
    } else if (OP == 2) {  // FADD only
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __fadd_rn(acc[i], p[i]);
    } else if (OP == 3) {  // FMUL only
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __fmul_rn(acc[i], b);
    } else if (OP == 4) {  // FFMA2
#pragma unroll
      for (int i = 0; i < kAcc / 2; ++i) {
        unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&p2[i]), y = *reinterpret_cast<unsigned long long*>(&w2),
                              z = *reinterpret_cast<unsigned long long*>(&acc2[i]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
        *reinterpret_cast<unsigned long long*>(&acc2[i]) = r;
      }
    } else if (OP == 5) {  // exact pair through two FMA2: product = fma(p, w, -0), sum = fma(product, 1, acc)
      const float2 nz = make_float2(-0.0f, -0.0f), one = make_float2(1.0f, 1.0f);
#pragma unroll
      for (int i = 0; i < kAcc / 2; ++i) {
        unsigned long long r, q, x = *reinterpret_cast<unsigned long long*>(&p2[i]), y = *reinterpret_cast<unsigned long long*>(&w2),
                                 z = *reinterpret_cast<unsigned long long*>(&acc2[i]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(x), "l"(y), "l"(*reinterpret_cast<const unsigned long long*>(&nz)));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(q), "l"(*reinterpret_cast<const unsigned long long*>(&one)), "l"(z));
        *reinterpret_cast<unsigned long long*>(&acc2[i]) = r;
      }
    } else if (OP == 10) {  // same, but 1.0 and -0.0 are run-time values: nothing for ptxas to fold
      const float2 nz = make_float2(rt_negzero, rt_negzero), one = make_float2(rt_one, rt_one);
#pragma unroll
      for (int i = 0; i < kAcc / 2; ++i) {
        unsigned long long r, q, x = *reinterpret_cast<unsigned long long*>(&p2[i]), y = *reinterpret_cast<unsigned long long*>(&w2),
                                 z = *reinterpret_cast<unsigned long long*>(&acc2[i]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(x), "l"(y), "l"(*reinterpret_cast<const unsigned long long*>(&nz)));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(q), "l"(*reinterpret_cast<const unsigned long long*>(&one)), "l"(z));
        *reinterpret_cast<unsigned long long*>(&acc2[i]) = r;
      }
    } else if (OP == 6) {  // byte -> float conversion (PRMT + FADD) feeding an FADD
#pragma unroll
      for (int i = 0; i < kAcc; ++i) {
        const float f = __uint_as_float(__byte_perm(word + i, 0x4B000000u, 0x7440 + (i & 3))) - 8388608.0f;
        acc[i] = __fadd_rn(acc[i], f);
      }
    } else if (OP == 7) {  // FMUL + FADD with the multiplicand changing every tap (no operand reuse)
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(p[i], p[(i + 1) % kAcc]));
    } else if (OP == 8) {  // FFMA with an immediate multiplier
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = fmaf(p[i], 1.5f, acc[i]);
    } else if (OP == 9) {  // 1 FMNMX + 1 FADD mix (alu + fma)
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = __fadd_rn(fmaxf(acc[i], p[i]), b);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < kAcc / 2; ++i) s += acc2[i].x + acc2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 22);
  cudaMalloc(&cyc, 1024 * 8);
  printf("%-44s", name);
  for (int wps : {1, 2, 4, 8}) {  // warps per sub-partition; one CTA per SM
    const int threads = wps * 4 * 32;
    k<OP><<<148, threads>>>(out, 1.0001f, 0.999f, cyc, 1.0f, -0.0f);
    k<OP><<<148, threads>>>(out, 1.0001f, 0.999f, cyc, 1.0f, -0.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double ipc = (double)instr_per_iter * kIters * wps / avg;  // warp-instr per cycle per sub-partition
    printf("  w%d: %.3f", wps, ipc);
  }
  printf("   (warp-instr / clk / SMSP)\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA 3-reg", kAcc);
  run<8>("FFMA imm", kAcc);
  run<1>("FMUL+FADD (exact tap, shared weight)", 2 * kAcc);
  run<7>("FMUL+FADD (no operand reuse)", 2 * kAcc);
  run<2>("FADD", kAcc);
  run<3>("FMUL", kAcc);
  run<4>("FFMA2 (f32x2)", kAcc / 2);
  run<5>("exact tap as 2 x FFMA2", kAcc);
  run<10>("exact tap as 2 x FFMA2, run-time 1 / -0", kAcc);
  run<6>("PRMT+FADD convert, + FADD", 3 * kAcc);
  run<9>("FMNMX + FADD", 2 * kAcc);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
