// Native driver for BASELINE config C5 on one GPU: N images of 1920x1080 RGBA8, encode + decode through the C ABI with
// device-resident images, T host threads with one context (= one stream) each.  Shows the per-image rate the ABI
// sustains without an interpreter between the calls (tools/sharded_configs.py is bound by its Python loop).
//   g++ -O2 -std=c++17 tools/c5_native.cpp -Iinclude -Lpixlzr-rust_b200 -lpixlzr_b200 -Wl,-rpath,'$ORIGIN/../pixlzr-rust_b200' -lpthread -o tools/c5_native
//   tools/c5_native [images] [threads]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "pixlzr_b200.h"

static void check(pxz_ctx* c, pxz_status st, const char* what) {
  if (st != PXZ_OK) {
    fprintf(stderr, "%s failed: %d %s\n", what, (int)st, c ? pxz_last_error(c) : "");
    exit(1);
  }
}

int main(int argc, char** argv) {
  const int images = argc > 1 ? atoi(argv[1]) : 4096, threads = argc > 2 ? atoi(argv[2]) : 4;
  const uint32_t W = 1920, H = 1080, distinct = 16;
  std::vector<std::vector<uint8_t>> host(distinct, std::vector<uint8_t>((size_t)W * H * 4));
  uint64_t s = 0x5049584C5A52ull;
  for (uint32_t i = 0; i < distinct; ++i) {
    for (uint32_t y = 0; y < H; ++y)
      for (uint32_t x = 0; x < W; ++x) {
        const uint32_t tile = (y / 64) * 31 + x / 64 + i * 7;
        const int amp = 1 << (tile * 2654435761u >> 29);  // 1..128 per 64x64 tile
        uint8_t* p = &host[i][((size_t)y * W + x) * 4];
        for (int c = 0; c < 3; ++c) {
          s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
          const int noise = (int)((s * 2685821657736338717ull) >> 56) * amp / 128 - amp;
          const int v = 128 + (int)(90 * sinf((x + 3 * c * y) / 97.0f + i)) + noise;
          p[c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
        p[3] = 255;
      }
  }
  std::vector<pxz_ctx*> ctx(threads);
  std::vector<std::vector<pxz_image*>> img(threads);
  std::vector<pxz_image*> out(threads);
  for (int t = 0; t < threads; ++t) {
    check(nullptr, pxz_ctx_create(0, &ctx[t]), "pxz_ctx_create");
    img[t].resize(distinct);
    for (uint32_t i = 0; i < distinct; ++i) check(ctx[t], pxz_image_upload(ctx[t], host[i].data(), W, H, 4, (size_t)W * 4, &img[t][i]), "upload");
    check(ctx[t], pxz_image_alloc(ctx[t], W, H, 4, &out[t]), "alloc");
  }
  auto work = [&](int t, int n) {
    for (int k = 0; k < n; ++k) {
      pxz_payload* pl = nullptr;
      check(ctx[t], pxz_shrink(ctx[t], img[t][k % distinct], 64, 64, PXZ_METRIC_OKLAB_MAD, 1.0f, PXZ_LANCZOS3, 0, &pl), "shrink");
      check(ctx[t], pxz_expand_to_image(ctx[t], pl, PXZ_LANCZOS3, out[t]), "expand");
      pxz_payload_free(pl);
    }
    check(ctx[t], pxz_synchronize(ctx[t]), "sync");
  };
  for (int rep = 0; rep < 3; ++rep) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t, images / threads);
    for (auto& x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int done = images / threads * threads;
    if (rep) printf("{\"config\": \"C5 native driver, %d images 1920x1080 RGBA8, %d host threads / contexts\", \"images_per_s\": %.0f, \"MP/s\": %.0f, \"ms\": %.2f}\n",
                    done, threads, done / dt, done * (double)W * H / dt / 1e6, dt * 1e3);
  }
  return 0;
}
