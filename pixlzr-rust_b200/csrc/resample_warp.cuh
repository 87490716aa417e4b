// resample_warp.cuh — warp-per-tile resample kernels for RGBA8 tiles up to 64x64 (included by kernels.cu, namespace pxz).
//
// Every warp owns one tile at a time (tiles are handed out through an atomic counter), so there is no CTA barrier
// anywhere and a light tile never idles the threads of a heavy one:
//   * shrink: the source tile is streamed from global memory exactly once, row by row (lane = columns x and x + 32);
//     the vertical pass keeps all live output rows in registers ("slide" table form, pxz_internal.h) and emits one
//     finished row of f32 intermediates at a time into an 8-row shared-memory strip; whenever the strip is full the
//     warp runs the horizontal pass over it and writes the packed payload pixels.
//   * expand: the (small) source block is streamed once through a 7-row register window of converted samples; two
//     rows of intermediates at a time go through shared memory to the horizontal pass, which writes the image tile
//     (the paste is fused: 16-byte stores into the pitched image).
// Arithmetic order per output is the reference's (ascending taps, product and sum rounded separately), on the packed
// f32x2 pipe (Acc4 / Acc1 / mac2 in kernels.cu).
#pragma once

constexpr int kWarpCtaThreads = 128;  // 4 warps
constexpr int kWarpsPerCta = kWarpCtaThreads / 32;
constexpr int kStripRows = 8;         // shrink: rows of intermediates per horizontal batch
constexpr int kStripStride = 64;      // float4 per strip row
constexpr int kShrinkStripPx = kStripRows * kStripStride;
constexpr int kExpandStripPx = 2 * kStripStride;

__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// next tile of this warp (lane 0 asks, everybody gets the answer).  Every warp draws exactly one index >= ntiles; the
// warp that draws the very last one puts the counter back to zero for the next launch on this stream.
__device__ __forceinline__ uint32_t next_tile(uint32_t* counter, uint32_t ntiles, uint32_t total_warps) {
  uint32_t b = 0;
  if ((threadIdx.x & 31u) == 0) {
    b = atomicAdd(counter, 1u);
    if (b == ntiles + total_warps - 1) *counter = 0u;
  }
  return __shfl_sync(0xffffffffu, b, 0);
}

// ---- shrink ------------------------------------------------------------------------------------------------------
template <int MODE, int NP, int S>
__device__ __forceinline__ void take_slot(u64 (&acc)[2][(MODE & 1) ? 4 : 3][NP], float4& v0, float4& v1) {
  constexpr int NC = (MODE & 1) ? 4 : 3;
  constexpr int jp = S / 2;
  constexpr bool high = (S & 1) != 0;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float l, h;
    unpk2(acc[0][c][jp], l, h);
    a[c] = high ? h : l;
    acc[0][c][jp] = high ? pk2(l, 0.f) : pk2(0.f, h);
    unpk2(acc[1][c][jp], l, h);
    b[c] = high ? h : l;
    acc[1][c][jp] = high ? pk2(l, 0.f) : pk2(0.f, h);
  }
  v0 = make_float4(a[0], a[1], a[2], a[3]);
  v1 = make_float4(b[0], b[1], b[2], b[3]);
}

// horizontal pass over `nrows` strip rows (output rows row0 .. row0 + nrows - 1) -> payload block at dst
template <int MODE>
__device__ __forceinline__ void shrink_horizontal(const float4* strip, uint32_t row0, uint32_t nrows, uint32_t dw, const AxisTab& tx,
                                                  const uint32_t* __restrict__ pool, uint32_t* dst, const TapK& k) {
  const uint32_t lane = threadIdx.x & 31u;
  if (dw >= 16) {
    // groups of 4 outputs share one walk (blocked table form); item = (strip row, group), 8 rows per group so that
    // the 8 lanes of a shared-memory phase read 8 different rows (the row-XOR swizzle makes them 8 different banks)
    const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(pool + tx.boff);
    const uint32_t* lo = pool + tx.boff + 4 * tx.brows_total;
    const uint32_t* rows = lo + tx.nb;
    const uint32_t* first = rows + tx.nb;
    for (uint32_t i = lane; i < tx.nb * kStripRows; i += 32) {
      const uint32_t r = i & (kStripRows - 1), ob = i / kStripRows;
      if (r >= nrows) continue;
      const uint32_t n = __ldg(rows + ob), c0 = __ldg(lo + ob), s7 = r & 7u;
      const ulonglong2* wp = w4 + __ldg(first + ob);
      const float4* trow = strip + r * kStripStride;
      Acc4<MODE> acc;
      uint32_t c = 0;
      for (; c + 2 <= n; c += 2) {
        const float4 p0 = trow[(c0 + c) ^ s7], p1 = trow[(c0 + c + 1) ^ s7];
        const ulonglong2 w0 = __ldg(wp + c), w1 = __ldg(wp + c + 1);
        acc.step(p0, w0, k);
        acc.step(p1, w1, k);
      }
      if (c < n) acc.step(trow[(c0 + c) ^ s7], __ldg(wp + c), k);
      const uint32_t ox = ob * 4, nvalid = min(4u, dw - ox);
      uint32_t* o = dst + (size_t)(row0 + r) * dw + ox;
      o[0] = pack_px<MODE>(acc.out(0));
      if (nvalid > 1) o[1] = pack_px<MODE>(acc.out(1));
      if (nvalid > 2) o[2] = pack_px<MODE>(acc.out(2));
      if (nvalid > 3) o[3] = pack_px<MODE>(acc.out(3));
    }
  } else {
    // few outputs: one scalar chain per (row, output, channel)
    const uint32_t* left = pool + tx.off;
    const uint32_t* cnt = left + dw;
    const float* w = reinterpret_cast<const float*>(left + 2 * dw);
    const uint32_t total = nrows * dw * 4;
    for (uint32_t i = lane; i < total; i += 32) {
      const uint32_t c = i & 3u, j = i >> 2, r = j / dw, ox = j - r * dw;
      uint32_t byte = 0xFFu;
      if ((MODE & 1) || c < 3) {
        const uint32_t n = __ldg(cnt + ox), l = __ldg(left + ox), s7 = r & 7u;
        const float* wr = w + ox * tx.stride;
        const float* trow = reinterpret_cast<const float*>(strip + r * kStripStride) + c;
        float a = 0.f;
        uint32_t t = 0;
        for (; t + 4 <= n; t += 4) {
          const float p0 = trow[((l + t) ^ s7) << 2], p1 = trow[((l + t + 1) ^ s7) << 2];
          const float p2 = trow[((l + t + 2) ^ s7) << 2], p3 = trow[((l + t + 3) ^ s7) << 2];
          const float w0 = __ldg(wr + t), w1 = __ldg(wr + t + 1), w2 = __ldg(wr + t + 2), w3 = __ldg(wr + t + 3);
          a = mac1<MODE>(a, p0, w0); a = mac1<MODE>(a, p1, w1); a = mac1<MODE>(a, p2, w2); a = mac1<MODE>(a, p3, w3);
        }
        for (; t < n; ++t) a = mac1<MODE>(a, trow[((l + t) ^ s7) << 2], __ldg(wr + t));
        byte = to_u8_fast(a);
      }
      reinterpret_cast<uint8_t*>(dst)[((size_t)(row0 + r) * dw + ox) * 4 + c] = (uint8_t)byte;
    }
  }
}

// vertical pass with the slide table: A accumulator slots (A / 2 register pairs) per column and channel
template <int MODE, int A>
__device__ __forceinline__ void shrink_tile_slide(const uint8_t* __restrict__ img, size_t pitch, const Tile& t,
                                                  const pxz_block_desc& d, const AxisTab& tx, const AxisTab& ty,
                                                  const uint32_t* __restrict__ pool, float4* strip, uint8_t* __restrict__ payload,
                                                  const TapK& k) {
  constexpr int NC = (MODE & 1) ? 4 : 3;
  constexpr int NP = A / 2;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t sw = t.tw, sh = t.th, dw = d.w, dh = d.h;
  const bool has0 = lane < sw, has1 = lane + 32 < sw;
  const uint8_t* col0 = img + (size_t)t.y0 * pitch + (size_t)(t.x0 + lane) * 4;
  u64 acc[2][NC][NP];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[0][c][j] = acc[1][c][j] = 0ull;
  const ulonglong2* wrow = reinterpret_cast<const ulonglong2*>(pool + ty.soff);
  const uint32_t* done = pool + ty.soff + 8 * ty.n_in;
  uint32_t* dst = reinterpret_cast<uint32_t*>(payload + d.offset);
  uint32_t o_next = 0, slot = 0, batch0 = 0;

  auto emit = [&]() {
    float4 v0, v1;
    switch (slot) {
      case 0: take_slot<MODE, NP, 0>(acc, v0, v1); break;
      case 1: take_slot<MODE, NP, 1>(acc, v0, v1); break;
      case 2: if (A > 2) take_slot<MODE, NP, (A > 2 ? 2 : 0)>(acc, v0, v1); break;
      case 3: if (A > 2) take_slot<MODE, NP, (A > 2 ? 3 : 0)>(acc, v0, v1); break;
      case 4: if (A > 4) take_slot<MODE, NP, (A > 4 ? 4 : 0)>(acc, v0, v1); break;
      default: if (A > 4) take_slot<MODE, NP, (A > 4 ? 5 : 0)>(acc, v0, v1); break;
    }
    const uint32_t rr = o_next - batch0;
    strip[rr * kStripStride + (lane ^ (rr & 7u))] = v0;
    strip[rr * kStripStride + 32 + (lane ^ (rr & 7u))] = v1;
    ++o_next;
    slot = (slot + 1 == (uint32_t)A) ? 0u : slot + 1;
    if (o_next - batch0 == (uint32_t)kStripRows || o_next == dh) {
      __syncwarp();
      shrink_horizontal<MODE>(strip, batch0, o_next - batch0, dw, tx, pool, dst, k);
      __syncwarp();
      batch0 = o_next;
    }
  };

  // source rows run 4 ahead of the arithmetic through a small register queue
  uint32_t qa[4], qb[4];
  auto load_row = [&](uint32_t r, uint32_t& a, uint32_t& b) {
    const uint8_t* p = col0 + (size_t)r * pitch;
    a = (r < sh && has0) ? ldg_stream_u32(p) : 0u;
    b = (r < sh && has1) ? ldg_stream_u32(p + 128) : 0u;
  };
#pragma unroll
  for (int j = 0; j < 4; ++j) load_row((uint32_t)j, qa[j], qb[j]);
#pragma unroll 1
  for (uint32_t r = 0; r < sh; ++r) {
    const uint32_t wa_ = qa[0], wb_ = qb[0];
#pragma unroll
    for (int j = 0; j < 3; ++j) { qa[j] = qa[j + 1]; qb[j] = qb[j + 1]; }
    load_row(r + 4, qa[3], qb[3]);
    u64 w[NP];
    const ulonglong2 wa = __ldg(wrow + 2 * r);
    w[0] = wa.x;
    if (NP > 1) w[NP > 1 ? 1 : 0] = wa.y;
    if (NP > 2) w[NP > 2 ? 2 : 0] = __ldg(reinterpret_cast<const u64*>(wrow + 2 * r + 1));
    const uint32_t nd = __ldg(done + r);
    const float4 pa = px_to_f4<MODE>(wa_), pb = px_to_f4<MODE>(wb_);
    const float ca[4] = {pa.x, pa.y, pa.z, pa.w}, cb[4] = {pb.x, pb.y, pb.z, pb.w};
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const u64 ppa = pk2(ca[c], ca[c]), ppb = pk2(cb[c], cb[c]);
#pragma unroll
      for (int jp = 0; jp < NP; ++jp) {
        acc[0][c][jp] = mac2<MODE>(acc[0][c][jp], ppa, w[jp], k);
        acc[1][c][jp] = mac2<MODE>(acc[1][c][jp], ppb, w[jp], k);
      }
    }
    for (uint32_t i = 0; i < nd; ++i) emit();
  }
}

// vertical pass for tables without a slide form (more than 6 outputs live at once): groups of 4 output rows walk
// their source rows straight from global memory / L1
template <int MODE>
__device__ __forceinline__ void shrink_tile_blocked(const uint8_t* __restrict__ img, size_t pitch, const Tile& t,
                                                    const pxz_block_desc& d, const AxisTab& tx, const AxisTab& ty,
                                                    const uint32_t* __restrict__ pool, float4* strip, uint8_t* __restrict__ payload,
                                                    const TapK& k) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t sw = t.tw, dw = d.w, dh = d.h;
  const bool has0 = lane < sw, has1 = lane + 32 < sw;
  const uint8_t* col0 = img + (size_t)t.y0 * pitch + (size_t)(t.x0 + lane) * 4;
  uint32_t* dst = reinterpret_cast<uint32_t*>(payload + d.offset);
  const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(pool + ty.boff);
  const uint32_t* lo = pool + ty.boff + 4 * ty.brows_total;
  const uint32_t* rows = lo + ty.nb;
  const uint32_t* first = rows + ty.nb;
  for (uint32_t ob = 0; ob < ty.nb; ++ob) {
    const uint32_t n = __ldg(rows + ob), r0 = __ldg(lo + ob);
    const ulonglong2* wp = w4 + __ldg(first + ob);
    Acc4<MODE> a0, a1;
    for (uint32_t r = 0; r < n; ++r) {
      const uint8_t* p = col0 + (size_t)(r0 + r) * pitch;
      const uint32_t w0 = has0 ? __ldg(reinterpret_cast<const uint32_t*>(p)) : 0u;
      const uint32_t w1 = has1 ? __ldg(reinterpret_cast<const uint32_t*>(p + 128)) : 0u;
      const ulonglong2 w = __ldg(wp + r);
      a0.step(px_to_f4<MODE>(w0), w, k);
      a1.step(px_to_f4<MODE>(w1), w, k);
    }
    const uint32_t oy = ob * 4, rr = oy & (kStripRows - 1);
    strip[(rr + 0) * kStripStride + (lane ^ ((rr + 0) & 7u))] = a0.out(0);
    strip[(rr + 0) * kStripStride + 32 + (lane ^ ((rr + 0) & 7u))] = a1.out(0);
    strip[(rr + 1) * kStripStride + (lane ^ ((rr + 1) & 7u))] = a0.out(1);
    strip[(rr + 1) * kStripStride + 32 + (lane ^ ((rr + 1) & 7u))] = a1.out(1);
    strip[(rr + 2) * kStripStride + (lane ^ ((rr + 2) & 7u))] = a0.out(2);
    strip[(rr + 2) * kStripStride + 32 + (lane ^ ((rr + 2) & 7u))] = a1.out(2);
    strip[(rr + 3) * kStripStride + (lane ^ ((rr + 3) & 7u))] = a0.out(3);
    strip[(rr + 3) * kStripStride + 32 + (lane ^ ((rr + 3) & 7u))] = a1.out(3);
    const uint32_t filled = min(oy + 4, dh);
    if ((ob & 1u) || ob + 1 == ty.nb) {
      const uint32_t row0 = oy & ~(uint32_t)(kStripRows - 1);
      __syncwarp();
      shrink_horizontal<MODE>(strip, row0, filled - row0, dw, tx, pool, dst, k);
      __syncwarp();
    }
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(kWarpCtaThreads) k_shrink_warp(const uint8_t* __restrict__ img, size_t pitch, Geom g,
                                                                 const pxz_block_desc* __restrict__ descs,
                                                                 const uint32_t* __restrict__ tabidx,
                                                                 const uint8_t* __restrict__ opaque_flags, uint8_t* __restrict__ payload,
                                                                 const AxisTab* __restrict__ tabs, const uint32_t* __restrict__ pool,
                                                                 uint32_t* counter, float rt_one, float rt_negzero) {
  extern __shared__ float4 s_strip[];
  float4* strip = s_strip + (threadIdx.x >> 5) * kShrinkStripPx;
  const TapK k = make_tapk(rt_one, rt_negzero);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ntiles = g.cols * g.rows, total_warps = gridDim.x * kWarpsPerCta;
  constexpr int F = FUSED ? 2 : 0;
  for (;;) {
    const uint32_t b = next_tile(counter, ntiles, total_warps);
    if (b >= ntiles) break;
    const Tile t = tile_of(g, b);
    const pxz_block_desc d = descs[b];
    if (d.w == 0 || d.h == 0) continue;  // masked out (quadtree levels)
    if (d.w == t.tw && d.h == t.th) {
      // block.rs:279-281: clone.  The block is contiguous in the payload.
      const uint8_t* src = img + (size_t)t.y0 * pitch + (size_t)t.x0 * 4;
      uint32_t* dst = reinterpret_cast<uint32_t*>(payload + d.offset);
      for (uint32_t r0 = 0; r0 < t.th; r0 += 4) {
        uint32_t v0[4], v1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint8_t* p = src + (size_t)(r0 + j) * pitch + lane * 4;
          v0[j] = (r0 + j < t.th && lane < t.tw) ? ldg_stream_u32(p) : 0u;
          v1[j] = (r0 + j < t.th && lane + 32 < t.tw) ? ldg_stream_u32(p + 128) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (r0 + j < t.th && lane < t.tw) dst[(size_t)(r0 + j) * t.tw + lane] = v0[j];
          if (r0 + j < t.th && lane + 32 < t.tw) dst[(size_t)(r0 + j) * t.tw + lane + 32] = v1[j];
        }
      }
      continue;
    }
    const uint32_t ti = tabidx[b];
    const AxisTab tx = tabs[ti & 0xFFFFu], ty = tabs[ti >> 16];
    const bool opaque = opaque_flags != nullptr && opaque_flags[b] != 0;
    if (opaque) {
      switch (ty.slots) {
        case 2: shrink_tile_slide<F, 2>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
        case 4: shrink_tile_slide<F, 4>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
        case 6: shrink_tile_slide<F, 6>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
        default: shrink_tile_blocked<F>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
      }
    } else {
      switch (ty.slots) {
        case 2: shrink_tile_slide<F | 1, 2>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
        case 4: shrink_tile_slide<F | 1, 4>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
        case 6: shrink_tile_slide<F | 1, 6>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
        default: shrink_tile_blocked<F | 1>(img, pitch, t, d, tx, ty, pool, strip, payload, k); break;
      }
    }
  }
}

// ---- expand ------------------------------------------------------------------------------------------------------
// shared-memory column of intermediate sample c: odd 8-column groups flip bit 0, so that the stride-2 reads of a 2x
// upscale (8 lanes -> 16 columns) land in 8 different 16-byte banks
__device__ __forceinline__ uint32_t ecol(uint32_t c) { return c ^ ((c >> 3) & 1u); }

template <int MODE>
__device__ __forceinline__ void expand_tile_warp(uint8_t* __restrict__ img, size_t pitch, const Tile& t, const pxz_block_desc& d,
                                                 const AxisTab& tx, const AxisTab& ty, const uint32_t* __restrict__ pool,
                                                 float4* strip, const uint8_t* __restrict__ payload, const TapK& k) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t sw = d.w, sh = d.h, dw = t.tw, dh = t.th;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(payload + d.offset);
  uint8_t* dst = img + (size_t)t.y0 * pitch + (size_t)t.x0 * 4;
  const float4* g8 = reinterpret_cast<const float4*>(pool + ty.goff);
  const uint32_t* gleft = pool + ty.goff + 8 * ty.n_out;

  // horizontal: lane = (strip row, group of 4 outputs); fixed for the whole tile
  const uint32_t hr = lane >> 4, ob = lane & 15u;
  const bool hact = ob < tx.nb;
  const uint32_t* blo = pool + tx.boff + 4 * tx.brows_total;
  const uint32_t hn = hact ? __ldg(blo + tx.nb + ob) : 0u, hc0 = hact ? __ldg(blo + ob) : 0u;
  const ulonglong2* hw = reinterpret_cast<const ulonglong2*>(pool + tx.boff) + (hact ? __ldg(blo + 2 * tx.nb + ob) : 0u);
  const uint32_t hox = ob * 4, hvalid = hact ? min(4u, dw - hox) : 0u;
  auto horizontal = [&](uint32_t oy0, uint32_t nrows) {
    if (hact && hr < nrows) {
      const float4* trow = strip + hr * kStripStride;
      Acc4<MODE> acc;
      uint32_t c = 0;
      for (; c + 2 <= hn; c += 2) {
        const float4 p0 = trow[ecol(hc0 + c)], p1 = trow[ecol(hc0 + c + 1)];
        const ulonglong2 w0 = __ldg(hw + c), w1 = __ldg(hw + c + 1);
        acc.step(p0, w0, k);
        acc.step(p1, w1, k);
      }
      if (c < hn) acc.step(trow[ecol(hc0 + c)], __ldg(hw + c), k);
      uint32_t* o = reinterpret_cast<uint32_t*>(dst + (size_t)(oy0 + hr) * pitch) + hox;
      const uint32_t p0 = pack_px<MODE>(acc.out(0)), p1 = pack_px<MODE>(acc.out(1));
      const uint32_t p2 = pack_px<MODE>(acc.out(2)), p3 = pack_px<MODE>(acc.out(3));
      if (hvalid == 4) {
        *reinterpret_cast<uint4*>(o) = make_uint4(p0, p1, p2, p3);  // tile rows are 16-byte aligned
      } else {
        o[0] = p0;
        if (hvalid > 1) o[1] = p1;
        if (hvalid > 2) o[2] = p2;
      }
    }
  };

  if (sw <= 32) {
    // vertical: lane = source column (two row groups when the block is at most 16 wide); 7-row window of converted
    // samples in registers, advanced as the outputs' first tap moves down
    const bool two = sw <= 16;
    const uint32_t x = two ? (lane & 15u) : lane, rg = two ? (lane >> 4) : 0u;
    const bool vact = x < sw;
    u64 wrg[7], wba[7];
    float wbl[7];
    auto conv = [&](uint32_t word, u64& rgp, u64& bap, float& bl) {
      const float4 p = px_to_f4<MODE>(word);
      rgp = pk2(p.x, p.y);
      bap = pk2(p.z, p.w);
      bl = p.z;
    };
#pragma unroll
    for (int j = 0; j < 7; ++j) conv((vact && (uint32_t)j < sh) ? __ldg(src + (size_t)j * sw + x) : 0u, wrg[j], wba[j], wbl[j]);
    uint32_t L = 0;
    uint32_t nextpx = (vact && 7 < sh) ? __ldg(src + (size_t)7 * sw + x) : 0u;
    auto vrow = [&](uint32_t oy, uint32_t srow) {
      const uint32_t left = __ldg(gleft + oy);
      while (L < left) {
#pragma unroll
        for (int j = 0; j < 6; ++j) { wrg[j] = wrg[j + 1]; wba[j] = wba[j + 1]; wbl[j] = wbl[j + 1]; }
        conv(nextpx, wrg[6], wba[6], wbl[6]);
        ++L;
        nextpx = (vact && L + 7 < sh) ? __ldg(src + (size_t)(L + 7) * sw + x) : 0u;
      }
      const float4 wa = __ldg(g8 + 2 * oy), wb = __ldg(g8 + 2 * oy + 1);
      const float w[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
      Acc1<MODE> acc;
#pragma unroll
      for (int j = 0; j < 7; ++j) acc.step(wrg[j], wba[j], wbl[j], w[j], k);
      if (vact) strip[srow * kStripStride + ecol(x)] = acc.out();
    };
    for (uint32_t oy0 = 0; oy0 < dh; oy0 += 2) {
      const uint32_t nrows = min(2u, dh - oy0);
      if (two) {
        if (rg < nrows) vrow(oy0 + rg, rg);
      } else {
        vrow(oy0, 0);
        if (nrows > 1) vrow(oy0 + 1, 1);
      }
      __syncwarp();
      horizontal(oy0, nrows);
      __syncwarp();
    }
  } else {
    // wide source (only one axis was reduced): every tap is read and converted where it is used
    const bool has0 = lane < sw, has1 = lane + 32 < sw;
    const uint32_t* lcnt = pool + ty.off + ty.n_out;
    for (uint32_t oy0 = 0; oy0 < dh; oy0 += 2) {
      const uint32_t nrows = min(2u, dh - oy0);
      for (uint32_t r = 0; r < nrows; ++r) {
        const uint32_t oy = oy0 + r, left = __ldg(gleft + oy), n = __ldg(lcnt + oy);
        const float* w = reinterpret_cast<const float*>(g8 + 2 * oy);
        Acc1<MODE> a0, a1;
        for (uint32_t j = 0; j < n; ++j) {
          const float wj = __ldg(w + j);
          const float4 p0 = px_to_f4<MODE>(has0 ? __ldg(src + (size_t)(left + j) * sw + lane) : 0u);
          const float4 p1 = px_to_f4<MODE>(has1 ? __ldg(src + (size_t)(left + j) * sw + lane + 32) : 0u);
          a0.step(pk2(p0.x, p0.y), pk2(p0.z, p0.w), p0.z, wj, k);
          a1.step(pk2(p1.x, p1.y), pk2(p1.z, p1.w), p1.z, wj, k);
        }
        if (has0) strip[r * kStripStride + ecol(lane)] = a0.out();
        if (has1) strip[r * kStripStride + ecol(lane + 32)] = a1.out();
      }
      __syncwarp();
      horizontal(oy0, nrows);
      __syncwarp();
    }
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(kWarpCtaThreads) k_expand_warp(uint8_t* __restrict__ img, size_t pitch, Geom g,
                                                                 const pxz_block_desc* __restrict__ descs,
                                                                 const uint32_t* __restrict__ tabidx, const uint8_t* __restrict__ payload,
                                                                 const AxisTab* __restrict__ tabs, const uint32_t* __restrict__ pool,
                                                                 uint32_t* counter, float rt_one, float rt_negzero) {
  extern __shared__ float4 s_strip[];
  float4* strip = s_strip + (threadIdx.x >> 5) * kExpandStripPx;
  const TapK k = make_tapk(rt_one, rt_negzero);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ntiles = g.cols * g.rows, total_warps = gridDim.x * kWarpsPerCta;
  constexpr int F = FUSED ? 2 : 0;
  for (;;) {
    const uint32_t b = next_tile(counter, ntiles, total_warps);
    if (b >= ntiles) break;
    const Tile t = tile_of(g, b);
    const pxz_block_desc d = descs[b];
    const uint32_t sw = d.w, sh = d.h, dw = t.tw, dh = t.th;
    if (sw == 0 || sh == 0) continue;  // masked out: the tile keeps what the output image already holds
    uint8_t* dst = img + (size_t)t.y0 * pitch + (size_t)t.x0 * 4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(payload + d.offset);
    if (sw == dw && sh == dh) {
      for (uint32_t r0 = 0; r0 < dh; r0 += 4) {
        uint32_t v0[4], v1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v0[j] = (r0 + j < dh && lane < dw) ? __ldg(src + (size_t)(r0 + j) * dw + lane) : 0u;
          v1[j] = (r0 + j < dh && lane + 32 < dw) ? __ldg(src + (size_t)(r0 + j) * dw + lane + 32) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t* o = reinterpret_cast<uint32_t*>(dst + (size_t)(r0 + j) * pitch);
          if (r0 + j < dh && lane < dw) o[lane] = v0[j];
          if (r0 + j < dh && lane + 32 < dw) o[lane + 32] = v1[j];
        }
      }
      continue;
    }
    if (sw == 1 && sh == 1) {
      // every output has one tap of normalised weight w / w = 1.0 in both passes: the tile is the source pixel
      const uint32_t p = __ldg(src);
      const uint4 v = make_uint4(p, p, p, p);
      const uint32_t qpr = (dw + 3) >> 2;
      for (uint32_t q = lane; q < qpr * dh; q += 32) {
        const uint32_t y = q / qpr, x = (q - y * qpr) << 2;
        uint8_t* o = dst + (size_t)y * pitch + (size_t)x * 4;
        if (x + 4 <= dw) *reinterpret_cast<uint4*>(o) = v;
        else for (uint32_t i = x; i < dw; ++i) reinterpret_cast<uint32_t*>(dst + (size_t)y * pitch)[i] = p;
      }
      continue;
    }
    const uint32_t ti = tabidx[b];
    const AxisTab tx = tabs[ti & 0xFFFFu], ty = tabs[ti >> 16];
    uint32_t aand = 0xFF000000u;
    for (uint32_t i = lane; i < sw * sh; i += 32) aand &= __ldg(src + i);
    const bool opaque = __all_sync(0xffffffffu, (aand & 0xFF000000u) == 0xFF000000u) != 0;
    if (opaque) expand_tile_warp<F>(img, pitch, t, d, tx, ty, pool, strip, payload, k);
    else expand_tile_warp<F | 1>(img, pitch, t, d, tx, ty, pool, strip, payload, k);
  }
}
