// Micro-benchmark: how fast can one B200 READ a pitched 7680x4320 RGBA8 image tile by tile (64x64 px = 64 rows of 256 B)?
// Compares a linear 16-byte-per-lane sweep, warp-per-tile LDG loads, and warp-per-tile TMA rings (box rows x slots),
// at different numbers of resident warps per SM.  The resample / analysis kernels cannot beat these figures.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_stream tools/ubench_stream.cu && tools/ubench_stream
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int W = 7680, H = 4320, COLS = W / 64, ROWS = (H + 63) / 64, NT = COLS * ROWS;
constexpr size_t PITCH = (size_t)W * 4;

__global__ void k_linear(const uint4* __restrict__ p, size_t n, uint32_t* out) {
  uint32_t x = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
    x ^= a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += stride) x ^= __ldcs(p + i).x;
  if (x == 0x12345678u) out[0] = x;
}

// warp per tile, 16 B per lane, two rows per instruction, U instructions in flight
template <int U>
__global__ void k_tile_ldg(const uint8_t* __restrict__ img, uint32_t* out, uint32_t* counter, int dynamic) {
  const uint32_t lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const uint32_t total = gridDim.x * wpc;
  uint32_t t = blockIdx.x * wpc + (threadIdx.x >> 5);
  uint32_t x = 0;
  const uint32_t rr = lane >> 4, ch = lane & 15;
  if (dynamic) { if (lane == 0) t = atomicAdd(counter, 1u); t = __shfl_sync(~0u, t, 0); }
  while (t < NT) {
    const uint32_t by = t / COLS, bx = t - by * COLS;
    const uint32_t th = min(64, H - (int)by * 64);
    const uint8_t* base = img + (size_t)(by * 64 + rr) * PITCH + (size_t)bx * 256 + ch * 16;
    for (uint32_t r0 = 0; r0 < th; r0 += 2 * U) {
      uint4 v[U];
#pragma unroll
      for (int j = 0; j < U; ++j) v[j] = (r0 + 2 * j + rr < th) ? __ldcs(reinterpret_cast<const uint4*>(base + (size_t)(r0 + 2 * j) * PITCH)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < U; ++j) x ^= v[j].x ^ v[j].w;
    }
    if (dynamic) { if (lane == 0) t = atomicAdd(counter, 1u); t = __shfl_sync(~0u, t, 0); }
    else t += total;
  }
  if (x == 0x12345678u) out[0] = x;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t x, uint32_t y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar) : "memory");
}

// warp per tile, ring of S boxes of R rows; tiles statically strided or drawn from a counter; the stream continues
// into the warp's next tile when `cross` is set
template <int R, int S>
__global__ void k_tile_tma(const __grid_constant__ CUtensorMap tm, uint32_t* out, uint32_t* counter, int dynamic, int cross) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31, wpc = blockDim.x >> 5, wid = threadIdx.x >> 5;
  constexpr uint32_t BOX = R * 256, WB = S * BOX + 128;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem + wid * WB), bars = base + S * BOX;
  if (lane == 0) { for (int i = 0; i < S; ++i) mbar_init(bars + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const uint32_t total = gridDim.x * wpc;
  auto next = [&](uint32_t t) -> uint32_t {
    if (dynamic) { uint32_t r = 0; if (lane == 0) r = atomicAdd(counter, 1u); return __shfl_sync(~0u, r, 0); }
    return t + total;
  };
  uint32_t t = dynamic ? next(0) : blockIdx.x * wpc + wid;
  uint32_t x = 0, islot = 0, rslot = 0, rpar = 0, inflight = 0;
  // producer cursor: (tile, box)
  uint32_t pt = t, pb = 0, tn = 0xFFFFFFFFu;
  auto nb_of = [&](uint32_t tt) -> uint32_t { const uint32_t by = tt / COLS; return (min(64, H - (int)by * 64) + R - 1) / R; };
  auto pump = [&](uint32_t cur, uint32_t nxt) {
    while (inflight < (uint32_t)S) {
      uint32_t tt;
      if (pt == cur && pb < nb_of(cur)) tt = cur;
      else if (cross && nxt < NT) { if (pt != nxt) { pt = nxt; pb = 0; } if (pb >= nb_of(nxt)) break; tt = nxt; }
      else break;
      const uint32_t by = tt / COLS, bx = tt - by * COLS;
      if (lane == 0) { mbar_expect_tx(bars + 8 * islot, BOX); tma_load_2d(base + islot * BOX, &tm, bx * 64, by * 64 + pb * R, bars + 8 * islot); }
      ++pb; islot = islot + 1 == S ? 0 : islot + 1; ++inflight;
    }
  };
  while (t < NT) {
    tn = next(t);
    if (pt != t) { pt = t; pb = 0; }
    const uint32_t nb = nb_of(t);
    pump(t, tn);
    for (uint32_t b = 0; b < nb; ++b) {
      mbar_wait(bars + 8 * rslot, rpar);
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + rslot * BOX + lane * 4));
      x ^= v;
      if (rslot + 1 == S) { rslot = 0; rpar ^= 1; } else ++rslot;
      __syncwarp();
      --inflight;
      pump(t, tn);
    }
    t = tn;
  }
  if (x == 0x12345678u) out[0] = x;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
static float time_it(F launch, uint32_t* d_counter) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaMemset(d_counter, 0, 4));
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  uint8_t* img; uint32_t *out, *counter;
  const size_t bytes = PITCH * H;
  CK(cudaMalloc(&img, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&counter, 64));
  CK(cudaMemset(img, 1, bytes));
  // a second buffer to flush L2 between runs is not needed: the image (133 MB) is larger than L2 (126 MB)
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  auto gbs = [&](float ms) { return bytes / (ms * 1e-3) / 1e9; };
  printf("linear sweep\n");
  for (int cps : {2, 4, 8}) {
    const float ms = time_it([&] { k_linear<<<148 * cps, 256>>>((const uint4*)img, bytes / 16, out); }, counter);
    printf("  %d CTAs/SM x 256 thr: %.1f us  %.0f GB/s\n", cps, ms * 1e3, gbs(ms));
  }
  printf("warp-per-tile LDG (two rows per instruction)\n");
  for (int dyn : {0, 1})
    for (int wps : {8, 12, 16, 32, 64}) {
      const int wpc = 4, grid = 148 * wps / wpc;
      float ms = time_it([&] { k_tile_ldg<4><<<grid, wpc * 32>>>(img, out, counter, dyn); }, counter);
      printf("  U=4  (8 rows in flight)  %2d warps/SM dyn=%d: %.1f us  %.0f GB/s\n", wps, dyn, ms * 1e3, gbs(ms));
      ms = time_it([&] { k_tile_ldg<8><<<grid, wpc * 32>>>(img, out, counter, dyn); }, counter);
      printf("  U=8  (16 rows in flight) %2d warps/SM dyn=%d: %.1f us  %.0f GB/s\n", wps, dyn, ms * 1e3, gbs(ms));
      ms = time_it([&] { k_tile_ldg<16><<<grid, wpc * 32>>>(img, out, counter, dyn); }, counter);
      printf("  U=16 (32 rows in flight) %2d warps/SM dyn=%d: %.1f us  %.0f GB/s\n", wps, dyn, ms * 1e3, gbs(ms));
    }
  printf("warp-per-tile TMA ring\n");
  auto run_tma = [&](auto kern, int R, int S, int wps, int dyn, int cross) {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {W, H}; const cuuint64_t strides[1] = {PITCH};
    const cuuint32_t box[2] = {64, (cuuint32_t)R}; const cuuint32_t es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); exit(1); }
    const int wpc = 4, grid = 148 * wps / wpc;
    const size_t smem = (size_t)wpc * (S * R * 256 + 128);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float ms = time_it([&] { kern<<<grid, wpc * 32, smem>>>(tm, out, counter, dyn, cross); }, counter);
    printf("  box %2d rows x %d slots, %2d warps/SM dyn=%d cross=%d: %.1f us  %.0f GB/s\n", R, S, wps, dyn, cross, ms * 1e3, gbs(ms));
  };
  for (int wps : {8, 12, 16, 32}) {
    for (int dyn : {0, 1})
      for (int cross : {0, 1}) {
        run_tma(k_tile_tma<4, 5>, 4, 5, wps, dyn, cross);
        run_tma(k_tile_tma<8, 3>, 8, 3, wps, dyn, cross);
      }
    run_tma(k_tile_tma<4, 9>, 4, 9, wps, 1, 1);
    run_tma(k_tile_tma<16, 3>, 16, 3, wps, 1, 1);
    if (wps <= 12) run_tma(k_tile_tma<64, 2>, 64, 2, wps, 1, 1);
  }
  return 0;
}
