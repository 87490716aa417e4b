// analyze_sobel_tma.cuh — directional Sobel metric (src/operations.rs:192-259) for RGBA8 images tiled 64x64
// (included by kernels.cu after resample_tma.cuh, namespace pxz).
//
// One CTA of 8 warps owns one tile at a time; the tiles of a CTA (blockIdx.x, + gridDim.x, ...) arrive through a
// three-stage ring of 2-D tensor copies (cp.async.bulk.tensor, 64 px x 64 rows = 16 KB per copy, one mbarrier per
// stage), so the copy of the tile after the next one is in flight while this one is summed.
//   warp  = band of 8 window rows, lane = window columns 2*lane and 2*lane + 1 (lane 31 idles: 62 window columns)
//   row terms: the four pixels a lane reads per row are byte-transposed into one register per channel (7 PRMT) and
//              h = v0 + 2 v1 + v2 and g = v2 - v0 of both columns are four dp4a per channel
//   per window row: |h(y+2) - h(y)| and |g(y) + 2 g(y+1) + g(y+2)| are accumulated with one VABSDIFF each.
// Integer arithmetic throughout (so warp reductions and shared-memory atomics cannot change the result); the value is
// the reference's f64 quotient rounded to f32.  Tiles narrower or lower than 3 px yield the reference's 0/0.
#pragma once

constexpr int kSobStages = 3;
constexpr int kSobTileBytes = 64 * 64 * 4;
constexpr int kSobBarOff = kSobStages * kSobTileBytes;
constexpr int kSobAccOff = kSobBarOff + 8 * kSobStages;
constexpr int kSobSmemBytes = kSobAccOff + 4 * 4;
#ifndef PXZ_SOBEL_CTAS
#define PXZ_SOBEL_CTAS 4
#endif

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
// unsigned bytes of a times signed bytes of b, summed
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0));
  return d;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}

struct SobelRow {
  int h[3][2];  // v(x) + 2 v(x+1) + v(x+2), channel x column
  int g[3][2];  // v(x+2) - v(x)
};
__device__ __forceinline__ SobelRow sobel_row_terms(uint32_t addr) {
  const uint2 a = lds_v2(addr), b = lds_v2(addr + 8);  // pixels x .. x+3
  const uint32_t t0 = prmt(a.x, a.y, 0x5140), t1 = prmt(b.x, b.y, 0x5140);  // [R0 R1 G0 G1], [R2 R3 G2 G3]
  const uint32_t t2 = prmt(a.x, a.y, 0x6262), t3 = prmt(b.x, b.y, 0x6262);  // [B0 B1 . .], [B2 B3 . .]
  const uint32_t q[3] = {prmt(t0, t1, 0x5410), prmt(t0, t1, 0x7632), prmt(t2, t3, 0x5410)};
  SobelRow r;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r.h[c][0] = dp4a_us(q[c], 0x00010201u);
    r.h[c][1] = dp4a_us(q[c], 0x01020100u);
    r.g[c][0] = dp4a_us(q[c], 0x000100FFu);
    r.g[c][1] = dp4a_us(q[c], 0x0100FF00u);
  }
  return r;
}

__global__ void __launch_bounds__(256, PXZ_SOBEL_CTAS) k_analyze_sobel_tma(const __grid_constant__ CUtensorMap tm, Geom g,
                                                                           float* __restrict__ vx, float* __restrict__ vy) {
  extern __shared__ __align__(128) uint8_t s_sob[];
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s_sob);
  uint32_t* s_acc = reinterpret_cast<uint32_t*>(s_sob + kSobAccOff);  // [2][2]: hz / vr sums of tile k & 1
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t ntiles = g.cols * g.rows;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < kSobStages; ++b) mbar_init(s0 + kSobBarOff + 8 * b, 1);
    s_acc[0] = s_acc[1] = s_acc[2] = s_acc[3] = 0u;
    fence_proxy_async();
  }
  __syncthreads();
  auto request = [&](uint32_t tile, uint32_t stage) {  // one thread
    const Tile t = tile_of(g, tile);
    const uint32_t bar = s0 + kSobBarOff + 8 * stage;
    mbar_expect_tx(bar, kSobTileBytes);  // out-of-image parts of the box arrive as zeros and count
    tma_load_2d(s0 + stage * kSobTileBytes, &tm, t.x0, t.y0, bar);
  };
  if (tid == 0) {
#pragma unroll
    for (uint32_t b = 0; b < (uint32_t)kSobStages; ++b) {
      const uint32_t tile = blockIdx.x + b * gridDim.x;
      if (tile < ntiles) request(tile, b);
    }
  }
  // this lane's window columns; lanes 30 and 31 read the same pixels (60 .. 63), lane 31 contributes nothing
  const uint32_t col0 = 2u * lane, lane_off = min(lane, 30u) * 8u;
  uint32_t stage = 0, parity = 0;
  for (uint32_t k = 0, tile = blockIdx.x; tile < ntiles; ++k, tile += gridDim.x) {
    const Tile t = tile_of(g, tile);
    const int ww = (int)t.tw - 2, wh = (int)t.th - 2;
    mbar_wait(s0 + kSobBarOff + 8 * stage, parity);
    uint32_t ahz0 = 0, ahz1 = 0, avr0 = 0, avr1 = 0;
    const int y_lo = (int)warp * 8, y_hi = min(wh, y_lo + 8);
    if (y_lo < y_hi && ww > 0) {
      const uint32_t base = s0 + stage * kSobTileBytes + (uint32_t)y_lo * 256u + lane_off;
      SobelRow r0 = sobel_row_terms(base), r1 = sobel_row_terms(base + 256u);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (y_lo + r >= y_hi) break;  // warp-uniform
        const SobelRow r2 = sobel_row_terms(base + (uint32_t)(r + 2) * 256u);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          ahz0 = __sad(r2.h[c][0], r0.h[c][0], ahz0);  // operations.rs:240-241, 247
          ahz1 = __sad(r2.h[c][1], r0.h[c][1], ahz1);
          avr0 = __sad(r1.g[c][0] * 2 + (r0.g[c][0] + r2.g[c][0]), 0, avr0);  // operations.rs:244-245, 248
          avr1 = __sad(r1.g[c][1] * 2 + (r0.g[c][1] + r2.g[c][1]), 0, avr1);
        }
        r0 = r1;
        r1 = r2;
      }
    }
    const bool v0 = (int)col0 < ww, v1 = (int)col0 + 1 < ww;
    const uint32_t hz = __reduce_add_sync(0xffffffffu, (v0 ? ahz0 : 0u) + (v1 ? ahz1 : 0u));
    const uint32_t vr = __reduce_add_sync(0xffffffffu, (v0 ? avr0 : 0u) + (v1 ? avr1 : 0u));
    uint32_t* acc = s_acc + 2 * (k & 1u);
    if (lane == 0) {
      atomicAdd(acc, hz);  // <= 62 * 62 * 3 * 1020 per tile: no overflow
      atomicAdd(acc + 1, vr);
    }
    __syncthreads();  // sums complete; every warp is through with this stage
    if (tid == 0) {
      const uint32_t a = acc[0], b = acc[1];
      acc[0] = 0u;  // next used by tile k + 2, after the barrier of tile k + 1
      acc[1] = 0u;
      const uint32_t nxt = tile + (uint32_t)kSobStages * gridDim.x;
      if (nxt < ntiles) request(nxt, stage);
      if (ww <= 0 || wh <= 0) {
        vx[tile] = __uint_as_float(0xFFC00000u);  // 0/0 in the reference (x86: negative quiet NaN)
        vy[tile] = __uint_as_float(0xFFC00000u);
      } else {
        const double f = (double)((unsigned long long)ww * (unsigned long long)wh * 4096ull);  // :158, :253-254
        vx[tile] = __double2float_rn(__ddiv_rn((double)a, f));
        vy[tile] = __double2float_rn(__ddiv_rn((double)b, f));
      }
    }
    if (++stage == (uint32_t)kSobStages) { stage = 0; parity ^= 1u; }
  }
}
