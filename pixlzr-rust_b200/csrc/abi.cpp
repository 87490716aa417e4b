// abi.cpp — implementation of the C ABI in include/pixlzr_b200.h: context / stream / scratch
// management, resample-table caching, and the stream-ordered pipelines
//
//   pxz_shrink : analyse -> [min/max (+ NCCL)] -> guard-band recompute -> plan (scan) -> resample
//   pxz_expand : resample (payload -> tiles of the pitched image)
//
// which replace the bodies of Pixlzr::shrink_by / shrink_directionally / expand / to_image
// (src/data_types/pixlzr.rs:77-205, pixlzr_image.rs:24-74).  No CPU fallback exists.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "pxz_host.h"

using namespace pxz;

// -------------------------------------------------------------------------------------------------
// objects
// -------------------------------------------------------------------------------------------------
namespace {

// index-aligned list of (reduced size, tile size) pairs that a payload's tabidx refers to
struct TabSpec {
  std::vector<uint32_t> n_small, n_tile;
};

struct TabSet {  // device copy of the axis tables of one (spec, filter, direction)
  AxisTab* d_tabs = nullptr;
  uint32_t* d_pool = nullptr;
  uint32_t max_words = 0;  // largest single table (left | count | weights), in 32-bit words
  uint32_t ntabs = 0;
  bool warp_ok = false;    // every table has the form the warp-per-tile kernels need
  bool has_noslide = false;  // direction 0: some table has no slide form (k_shrink_tma leaves those tiles to k_shrink_warp)
};

using TabKey = std::tuple<std::vector<uint32_t>, std::vector<uint32_t>, int, int>;

enum KernelId { K_MAD_FAST = 0, K_MAD_EXACT, K_SOBEL, K_MINMAX, K_PLAN, K_RESAMPLE_DOWN, K_RESAMPLE_UP, K_QOI_ENCODE, K_QOI_DECODE, K_COUNT };
const char* const kKernelNames[K_COUNT] = {"analyze_mad_fast", "mad_exact",     "analyze_sobel", "minmax",
                                           "plan",             "resample_down", "resample_up",   "qoi_encode",
                                           "qoi_decode"};
struct ProfRec {
  int id;
  cudaEvent_t a, b;
};

}  // namespace

// device buffers of a released payload, kept by the context for the next pxz_shrink / pxz_payload_upload so that
// steady-state calls neither touch the allocator nor share blocks with other streams' contexts
struct PayloadBufs {
  pxz_block_desc* d_descs = nullptr;
  uint32_t* d_tabidx = nullptr;
  uint8_t* d_pixels = nullptr;
  uint64_t* d_total = nullptr;
  size_t nblocks = 0;
  uint64_t capacity = 0;
};

struct pxz_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  size_t max_smem_optin = 0;
  std::string err;
  uint64_t launches = 0;
  LevelThresholds thr{};
  StrategyLut strategy{};  // on = 0 unless pxz_ctx_set_strategy installed a table
  GuardBand band{};
  // scratch, grown on demand
  float* d_vx = nullptr;
  float* d_vy = nullptr;
  uint8_t* d_opaque = nullptr;
  uint32_t* d_list = nullptr;  // [values_cap + 1]: banded tile indices, last word = count
  size_t values_cap = 0;
  float* d_minmax = nullptr;
  uint32_t* d_tile_counter = nullptr;  // work counter of the warp-per-tile resample kernels (they leave it at 0)
  int resample_kernels = 0;            // 0 = by tile count, 1 = warp-per-tile (cp.async ring), 2 = CTA-per-tile, 3 = warp-per-tile with the TMA shrink kernel (PXZ_RESAMPLE_KERNELS=warp|cta|tma)
  void* d_scan = nullptr;
  size_t scan_cap = 0;
  uint8_t* d_scratch = nullptr;
  size_t scratch_cap = 0;
  uint64_t* h_total = nullptr;  // pinned
  std::map<TabKey, TabSet> tabs;
  void* comm = nullptr;
  bool fast_resample = false;
  bool rgb_via_rgba = true;            // RGB images run on the RGBA fast kernels (PXZ_RGB_VIA_RGBA=0: the 3-channel kernels)
  int resize_semantics = PXZ_RESIZE_IMAGE_RS;  // which branch of PixlzrBlock::resize (pxz_ctx_set_resize_semantics)
  std::vector<PayloadBufs> payload_cache;  // at most kPayloadCacheMax entries
  // per-kernel timing
  bool profiling = false;
  std::vector<ProfRec> prof_open;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[K_COUNT] = {0};
  uint64_t prof_n[K_COUNT] = {0};
};

struct pxz_image {
  pxz_ctx* ctx;
  uint8_t* d;
  size_t pitch;
  uint32_t w, h, c;
  bool owned;
  uint32_t nimg = 1;  // > 1: a batch, image i at rows [i * h, (i + 1) * h) of the allocation
};

constexpr size_t kPayloadCacheMax = 4;
constexpr size_t kTabCacheMax = 48;  // table sets kept per context (get_tabset)

struct pxz_payload {
  pxz_ctx* ctx;
  Geom g;
  size_t nblocks_cap = 0;
  pxz_block_desc* d_descs = nullptr;
  uint32_t* d_tabidx = nullptr;  // [nblocks_cap] table indices, then the work-order lists (launch_plan) of capacity nblocks_cap
  uint32_t* d_order() const { return d_tabidx + nblocks_cap; }
  uint32_t* d_up() const { return d_order() + order_list_words(nblocks_cap); }  // [nblocks_cap] expand-side indices (strategy)
  // 0: one filter per call | 1: per-block filters, expand indices in d_up() (pxz_shrink) | 2: ... in d_tabidx (upload)
  int strategy = 0;
  uint8_t* d_pixels = nullptr;
  uint64_t capacity = 0;
  uint64_t* d_total = nullptr;
  bool bytes_known = false;
  uint64_t bytes = 0;
  std::shared_ptr<TabSpec> spec;
  uint32_t max_small_px = 0;  // largest reduced block (pixels)
  uint32_t max_small_dim = 0; // largest reduced block side
  uint32_t max_tmp_down = 0;  // largest vertical-pass intermediate when shrinking / expanding (pixels)
  uint32_t max_tmp_up = 0;
};

namespace {

pxz_status fail(pxz_ctx* ctx, pxz_status st, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return st;
}

#define PXZ_CUDA(ctx, expr)                                                                             \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      cudaGetLastError();                                                                               \
      return fail((ctx), _e == cudaErrorMemoryAllocation ? PXZ_E_OOM : PXZ_E_CUDA,                      \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                                  \
    }                                                                                                   \
  } while (0)

// records an event pair around one kernel launch while profiling is on
struct ProfScope {
  pxz_ctx* ctx;
  ProfRec rec;
  bool live = false;
  ProfScope(pxz_ctx* c, int id) : ctx(c) {
    if (!c->profiling) return;
    auto take = [&](cudaEvent_t* e) {
      if (!c->prof_pool.empty()) {
        *e = c->prof_pool.back();
        c->prof_pool.pop_back();
        return true;
      }
      return cudaEventCreate(e) == cudaSuccess;
    };
    rec.id = id;
    if (!take(&rec.a)) return;
    if (!take(&rec.b)) { c->prof_pool.push_back(rec.a); return; }
    cudaEventRecord(rec.a, c->stream);
    live = true;
  }
  ~ProfScope() {
    if (!live) return;
    cudaEventRecord(rec.b, ctx->stream);
    ctx->prof_open.push_back(rec);
  }
};

inline uint32_t ceil_div_u32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + b - 1) / b); }

pxz_status make_geom(pxz_ctx* ctx, uint32_t w, uint32_t h, uint32_t c, uint32_t bw, uint32_t bh, Geom* g, uint32_t nimg = 1) {
  if (w == 0 || h == 0 || nimg == 0) return fail(ctx, PXZ_E_ARG, "empty image");
  if (c != 3 && c != 4) return fail(ctx, PXZ_E_ARG, "channels must be 3 (RGB8) or 4 (RGBA8)");
  if (bw == 0 || bh == 0) return fail(ctx, PXZ_E_ARG, "block size must be >= 1 (the reference divides by it)");
  if (bw > 65535u || bh > 65535u) return fail(ctx, PXZ_E_UNSUPPORTED, "block size above 65535 is not supported");
  g->W = w; g->H = h; g->bw = bw; g->bh = bh; g->C = c;
  g->cols = ceil_div_u32(w, bw);  // == ceil(w as f64 / bw as f64), split.rs:45-46
  g->rows_img = ceil_div_u32(h, bh);
  g->nimg = nimg;
  g->img_rows = h;
  g->trail_w = w % bw;
  g->trail_h = h % bh;
  if ((uint64_t)g->cols * g->rows_img * nimg > 0x7FFFFFFFull) return fail(ctx, PXZ_E_UNSUPPORTED, "more than 2^31 blocks");
  if ((uint64_t)h * nimg > 0xFFFFFFFFull) return fail(ctx, PXZ_E_UNSUPPORTED, "batch taller than 2^32 rows");
  g->rows = g->rows_img * nimg;  // block rows of the whole batch
  return PXZ_OK;
}

pxz_status dev_alloc(pxz_ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  PXZ_CUDA(ctx, cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream));
  return PXZ_OK;
}
void dev_free(pxz_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}

pxz_status ensure_scratch(pxz_ctx* ctx, uint32_t nblocks) {
  if (ctx->values_cap < nblocks) {
    dev_free(ctx, ctx->d_vx);
    dev_free(ctx, ctx->d_vy);
    dev_free(ctx, ctx->d_opaque);
    dev_free(ctx, ctx->d_list);
    ctx->d_vx = ctx->d_vy = nullptr;
    ctx->d_opaque = nullptr;
    ctx->d_list = nullptr;
    ctx->values_cap = 0;
    pxz_status st;
    if ((st = dev_alloc(ctx, (void**)&ctx->d_vx, (size_t)nblocks * 4)) != PXZ_OK) return st;
    if ((st = dev_alloc(ctx, (void**)&ctx->d_vy, (size_t)nblocks * 4)) != PXZ_OK) return st;
    if ((st = dev_alloc(ctx, (void**)&ctx->d_opaque, (size_t)nblocks)) != PXZ_OK) return st;
    if ((st = dev_alloc(ctx, (void**)&ctx->d_list, ((size_t)nblocks + 1) * 4)) != PXZ_OK) return st;
    ctx->values_cap = nblocks;
  }
  const size_t need = plan_scan_state_bytes(nblocks);
  if (ctx->scan_cap < need) {
    dev_free(ctx, ctx->d_scan);
    ctx->d_scan = nullptr;
    ctx->scan_cap = 0;
    pxz_status st;
    if ((st = dev_alloc(ctx, &ctx->d_scan, need)) != PXZ_OK) return st;
    // the plan / work-order kernels find this zeroed and leave it zeroed
    PXZ_CUDA(ctx, cudaMemsetAsync(ctx->d_scan, 0, need, ctx->stream));
    ctx->scan_cap = need;
  }
  return PXZ_OK;
}

// table index layout of payloads produced by pxz_shrink: (axis * 2 + trailing) * 17 + level
std::shared_ptr<TabSpec> spec_for_geom(const Geom& g) {
  auto sp = std::make_shared<TabSpec>();
  const uint32_t n_of[4] = {g.bw, g.trail_w ? g.trail_w : g.bw, g.bh, g.trail_h ? g.trail_h : g.bh};
  for (int a = 0; a < 4; ++a) {
    for (int k = 0; k <= kMaxLevel; ++k) {
      const uint32_t n = n_of[a];
      uint32_t nn = (uint32_t)(((uint64_t)n + ((1ull << k) - 1)) >> k);
      if (nn < 1) nn = 1;
      sp->n_small.push_back(nn);
      sp->n_tile.push_back(n);
    }
  }
  return sp;
}

// direction 0: tile -> reduced (filter_down); 1: reduced -> tile (filter_up)
pxz_status get_tabset(pxz_ctx* ctx, const TabSpec& spec, int filter, int direction, TabSet* out) {
  const bool fir = ctx->resize_semantics == PXZ_RESIZE_FIR;
  TabKey key(spec.n_small, spec.n_tile, filter, direction + (fir ? 2 : 0));
  auto it = ctx->tabs.find(key);
  if (it != ctx->tabs.end()) {
    *out = it->second;
    return PXZ_OK;
  }
  // filter < 0: the tables of all five filters one after the other (index = filter * spec size + i), for payloads
  // whose blocks carry their own filter (pxz_ctx_set_strategy)
  const size_t per_filter = spec.n_small.size();
  const int f0 = filter < 0 ? 0 : filter, f1 = filter < 0 ? 4 : filter;
  std::vector<AxisTab> tabs(per_filter * (size_t)(f1 - f0 + 1));
  std::vector<uint32_t> pool;
  for (int f = f0; f <= f1; ++f) {
    std::map<std::pair<uint32_t, uint32_t>, AxisTab> seen;
    for (size_t i = 0; i < per_filter; ++i) {
      AxisTab& t = tabs[(size_t)(f - f0) * per_filter + i];
      const uint32_t n_in = direction == 0 ? spec.n_tile[i] : spec.n_small[i];
      const uint32_t n_out = direction == 0 ? spec.n_small[i] : spec.n_tile[i];
      auto s = seen.find({n_in, n_out});
      if (s != seen.end()) {
        t = s->second;
        continue;
      }
      if (fir) {
        // FilterType::to_fir_resizing_algorithm (data_types/mod.rs:65-107): Triangle is Hamming when a block shrinks and
        // Bilinear when it grows; the other filters keep their kernel, Nearest stays nearest
        const int alg = f == PXZ_NEAREST ? PXZ_FIR_NEAREST : f == PXZ_TRIANGLE ? (direction == 0 ? PXZ_FIR_HAMMING : PXZ_FIR_BILINEAR)
                        : f == PXZ_CATMULLROM ? PXZ_FIR_CATMULLROM : f == PXZ_GAUSSIAN ? PXZ_FIR_GAUSSIAN : PXZ_FIR_LANCZOS3;
        if (!build_axis_table_fir(n_in, n_out, alg, &pool, &t)) return fail(ctx, PXZ_E_ARG, "bad resample table request");
      } else if (!build_axis_table(n_in, n_out, f, &pool, &t)) {
        return fail(ctx, PXZ_E_ARG, "bad resample table request");
      }
      seen[{n_in, n_out}] = t;
    }
  }
  pool.resize(pool.size() + 16, 0u);  // staged copies are rounded up to 16 bytes
  TabSet ts;
  ts.ntabs = (uint32_t)tabs.size();
  ts.warp_ok = true;
  for (const AxisTab& t : tabs) {
    ts.max_words = std::max(ts.max_words, std::max(2 * t.n_out + t.n_out * t.stride, t.bwords));
    if (direction == 1 && t.goff == 0xFFFFFFFFu) ts.warp_ok = false;
    if (direction == 0 && t.s2words == 0) ts.has_noslide = true;
  }
  // the cache is bounded: payloads built from host descriptors (decoded files, PixlzrBlock::resize size pairs) bring
  // their own size sets, and a long-running decoder must not pile their tables up.  Frees are stream-ordered, so kernels
  // already queued on this context's stream keep their tables.
  if (ctx->tabs.size() >= kTabCacheMax) {
    for (auto& kv : ctx->tabs) {
      dev_free(ctx, kv.second.d_tabs);
      dev_free(ctx, kv.second.d_pool);
    }
    ctx->tabs.clear();
  }
  pxz_status st;
  if ((st = dev_alloc(ctx, (void**)&ts.d_tabs, tabs.size() * sizeof(AxisTab))) != PXZ_OK) return st;
  if ((st = dev_alloc(ctx, (void**)&ts.d_pool, pool.size() * 4)) != PXZ_OK) {
    dev_free(ctx, ts.d_tabs);
    return st;
  }
  // pageable -> device copies are staged by the runtime before returning, so the vectors may die
  cudaError_t ce = cudaMemcpyAsync(ts.d_tabs, tabs.data(), tabs.size() * sizeof(AxisTab), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(ts.d_pool, pool.data(), pool.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    dev_free(ctx, ts.d_tabs);
    dev_free(ctx, ts.d_pool);
    return fail(ctx, PXZ_E_CUDA, std::string("resample table upload: ") + cudaGetErrorString(ce));
  }
  ctx->tabs[key] = ts;
  *out = ts;
  return PXZ_OK;
}

pxz_status run_resample(pxz_ctx* ctx, int direction, uint8_t* img, size_t pitch, const pxz_payload* p, const TabSet& ts,
                        uint32_t max_src_px, uint32_t max_src_dim, uint32_t max_tmp_px, const uint8_t* opaque_flags = nullptr) {
  const Geom& g = p->g;
  const uint32_t nblocks = g.cols * g.rows;
  if (ctx->resize_semantics == PXZ_RESIZE_FIR) {
    // integer convolution, horizontal pass first: its own kernel (one CTA per block, source + u8 intermediate in smem)
    const uint32_t tmp_px = max_src_dim * std::max(g.bw, max_src_dim);  // source rows x widest destination row
    const size_t fsmem = resample_fir_smem_bytes(max_src_px, tmp_px, g.C);
    if (fsmem > ctx->max_smem_optin) return fail(ctx, PXZ_E_UNSUPPORTED, "blocks too large for the fir resize path");
    ProfScope prof(ctx, direction == 0 ? K_RESAMPLE_DOWN : K_RESAMPLE_UP);
    const uint32_t* tabidx = (direction == 1 && p->strategy == 1) ? p->d_up() : p->d_tabidx;
    PXZ_CUDA(ctx, launch_resample_fir(direction, img, pitch, g, p->d_descs, tabidx, p->d_pixels, ts.d_tabs, ts.d_pool, fsmem,
                                      (uint32_t)(((size_t)max_src_px * g.C + 15) & ~(size_t)15), ctx->stream, ctx->sm_count, &ctx->launches));
    return PXZ_OK;
  }
  size_t smem = resample_smem_bytes(max_src_px, max_tmp_px, g.C);
  int grid = resample_grid(ctx->sm_count, nblocks);
  uint8_t* scratch = nullptr;
  size_t per_cta = 0;
  if (smem > ctx->max_smem_optin) {  // tiles too large for shared memory
    per_cta = (smem + 255) & ~(size_t)255;
    grid = std::min<int>(grid, ctx->sm_count * 2);
    const size_t need = per_cta * (size_t)grid;
    if (ctx->scratch_cap < need) {
      dev_free(ctx, ctx->d_scratch);
      ctx->d_scratch = nullptr;
      ctx->scratch_cap = 0;
      pxz_status st;
      if ((st = dev_alloc(ctx, (void**)&ctx->d_scratch, need)) != PXZ_OK) return st;
      ctx->scratch_cap = need;
    }
    scratch = ctx->d_scratch;
  }
  ProfScope prof(ctx, direction == 0 ? K_RESAMPLE_DOWN : K_RESAMPLE_UP);
  const uint32_t* tabidx = (direction == 1 && p->strategy == 1) ? p->d_up() : p->d_tabidx;
  PXZ_CUDA(ctx, launch_resample(direction, img, pitch, g, p->d_descs, tabidx, p->d_pixels, ts.d_tabs, ts.d_pool,
                                ts.ntabs, max_src_px, max_src_dim, max_tmp_px, ts.max_words, scratch, per_cta, grid, ctx->fast_resample, opaque_flags,
                                ctx->resample_kernels != 2 ? ctx->d_tile_counter : nullptr, p->d_order(), (uint32_t)p->nblocks_cap,
                                ts.warp_ok, ctx->resample_kernels, ts.has_noslide, ctx->stream, ctx->sm_count,
                                &ctx->launches));
  return PXZ_OK;
}

void payload_release(pxz_payload* p) {
  if (!p) return;
  pxz_ctx* ctx = p->ctx;
  if (p->d_descs && p->d_tabidx && p->d_pixels && p->d_total && ctx->payload_cache.size() < kPayloadCacheMax) {
    PayloadBufs b;
    b.d_descs = p->d_descs; b.d_tabidx = p->d_tabidx; b.d_pixels = p->d_pixels; b.d_total = p->d_total;
    b.nblocks = p->nblocks_cap; b.capacity = p->capacity;
    ctx->payload_cache.push_back(b);  // stream order makes the reuse safe: later work on this ctx runs after ours
  } else {
    dev_free(ctx, p->d_descs);
    dev_free(ctx, p->d_tabidx);
    dev_free(ctx, p->d_pixels);
    dev_free(ctx, p->d_total);
  }
  delete p;
}

pxz_status ctx_create_common(int device, cudaStream_t stream, bool own, pxz_ctx** out) {
  if (!out) return PXZ_E_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return PXZ_E_CUDA;
  }
  if (device < 0 || device >= n) return PXZ_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return PXZ_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PXZ_E_CUDA;
  pxz_ctx* ctx = new (std::nothrow) pxz_ctx();
  if (!ctx) return PXZ_E_OOM;
  ctx->device = device;
  if (prop.major != 10) {
    // the library only carries sm_100a code; refuse loudly instead of failing at the first launch
    fprintf(stderr, "pixlzr_b200: device %d is sm_%d%d, need sm_100 (B200)\n", device, prop.major, prop.minor);
    delete ctx;
    return PXZ_E_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
  if (own) {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PXZ_E_CUDA; }
    ctx->own_stream = true;
  } else {
    ctx->stream = stream;
  }
  // keep freed blocks in the stream-ordered pool: alloc/free per call must not hit the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thresh = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
  }
  build_level_thresholds(&ctx->thr);
  ctx->band.rel = 2.0e-5f;     // fast arithmetic, relative
  ctx->band.abs_raw = 8.0e-6f; // fast arithmetic, absolute (SFU cube roots; measured <= 4e-6)
  if (const char* e = getenv("PXZ_GUARD_REL")) ctx->band.rel = (float)atof(e);
  if (const char* e = getenv("PXZ_GUARD_ABS")) ctx->band.abs_raw = (float)atof(e);
  if (const char* e = getenv("PXZ_RGB_VIA_RGBA")) ctx->rgb_via_rgba = atoi(e) != 0;
  if (const char* e = getenv("PXZ_RESIZE_SEMANTICS")) ctx->resize_semantics = !strcmp(e, "fir") ? PXZ_RESIZE_FIR : PXZ_RESIZE_IMAGE_RS;
  if (const char* e = getenv("PXZ_RESAMPLE_KERNELS")) ctx->resample_kernels = !strcmp(e, "warp") ? 1 : !strcmp(e, "cta") ? 2 : !strcmp(e, "tma") ? 3 : 0;
  if (cudaMalloc((void**)&ctx->d_minmax, 8 * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_tile_counter, 64) != cudaSuccess || cudaMemset(ctx->d_tile_counter, 0, 64) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_total, 64) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return PXZ_E_OOM;
  }
  *out = ctx;
  return PXZ_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
extern "C" {

int pxz_abi_version(void) { return PXZ_ABI_VERSION; }

int pxz_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

pxz_status pxz_ctx_create(int device, pxz_ctx** out) { return ctx_create_common(device, nullptr, true, out); }

pxz_status pxz_ctx_create_on_stream(int device, void* cuda_stream, pxz_ctx** out) {
  return ctx_create_common(device, (cudaStream_t)cuda_stream, false, out);
}

void pxz_ctx_destroy(pxz_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) nccl_comm_destroy(ctx->comm);
  for (auto& b : ctx->payload_cache) {
    dev_free(ctx, b.d_descs); dev_free(ctx, b.d_tabidx); dev_free(ctx, b.d_pixels); dev_free(ctx, b.d_total);
  }
  ctx->payload_cache.clear();
  for (auto& kv : ctx->tabs) {
    dev_free(ctx, kv.second.d_tabs);
    dev_free(ctx, kv.second.d_pool);
  }
  dev_free(ctx, ctx->d_vx);
  dev_free(ctx, ctx->d_vy);
  dev_free(ctx, ctx->d_opaque);
  dev_free(ctx, ctx->d_list);
  dev_free(ctx, ctx->d_scan);
  dev_free(ctx, ctx->d_scratch);
  cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->prof_open) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
  cudaFree(ctx->d_minmax);
  cudaFree(ctx->d_tile_counter);
  cudaFreeHost(ctx->h_total);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* pxz_last_error(const pxz_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

pxz_status pxz_synchronize(pxz_ctx* ctx) {
  if (!ctx) return PXZ_E_ARG;
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PXZ_OK;
}

uint64_t pxz_launch_count(const pxz_ctx* ctx) { return ctx ? ctx->launches : 0; }

pxz_status pxz_ctx_set_resize_semantics(pxz_ctx* ctx, pxz_resize_semantics semantics) {
  if (!ctx) return PXZ_E_ARG;
  if (semantics != PXZ_RESIZE_IMAGE_RS && semantics != PXZ_RESIZE_FIR) return fail(ctx, PXZ_E_ARG, "unknown resize semantics");
  ctx->resize_semantics = (int)semantics;
  return PXZ_OK;
}

pxz_status pxz_ctx_set_fast_resample(pxz_ctx* ctx, int on) {
  if (!ctx) return PXZ_E_ARG;
  ctx->fast_resample = on != 0;
  return PXZ_OK;
}

// ---- per-block filter pairs (strategies.txt / strategies_by_level.txt) ----------------------------------------
uint32_t pxz_strategy_bucket(float block_value) { return strategy_bucket(block_value); }

pxz_status pxz_strategy_by_level(pxz_strategy* out) {
  if (!out) return PXZ_E_ARG;
  // strategies_by_level.txt:1-12 (= strategies.txt, one line per bucket of width 1/64)
  for (int b = 0; b < PXZ_STRATEGY_BUCKETS; ++b) {
    pxz_filter down = PXZ_LANCZOS3, up = PXZ_LANCZOS3;  // v in [0.0625; 0.703125)
    if (b == 0) { down = PXZ_NEAREST; up = PXZ_NEAREST; }            // v < 0.015625
    else if (b == 1) { down = PXZ_TRIANGLE; up = PXZ_NEAREST; }      // [0.015625; 0.03125)
    else if (b == 2) { down = PXZ_CATMULLROM; up = PXZ_LANCZOS3; }   // [0.03125; 0.046875)
    else if (b == 3) { down = PXZ_LANCZOS3; up = PXZ_CATMULLROM; }   // [0.046875; 0.0625)
    else if (b >= 45) { down = PXZ_NEAREST; up = PXZ_NEAREST; }      // v >= 0.703125
    out->down[b] = (uint8_t)down;
    out->up[b] = (uint8_t)up;
  }
  return PXZ_OK;
}

pxz_status pxz_ctx_set_strategy(pxz_ctx* ctx, const pxz_strategy* strategy) {
  if (!ctx) return PXZ_E_ARG;
  if (!strategy) {
    ctx->strategy.on = 0;
    return PXZ_OK;
  }
  for (int b = 0; b < PXZ_STRATEGY_BUCKETS; ++b)
    if (strategy->down[b] > 4 || strategy->up[b] > 4) return fail(ctx, PXZ_E_ARG, "unknown filter in the strategy table");
  static_assert(PXZ_STRATEGY_BUCKETS == kStrategyBuckets, "bucket count");
  memcpy(ctx->strategy.down, strategy->down, PXZ_STRATEGY_BUCKETS);
  memcpy(ctx->strategy.up, strategy->up, PXZ_STRATEGY_BUCKETS);
  ctx->strategy.on = 1;
  return PXZ_OK;
}

static pxz_status prof_drain(pxz_ctx* ctx) {
  if (ctx->prof_open.empty()) return PXZ_OK;
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->prof_open) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ctx->prof_ms[r.id] += ms;
      ctx->prof_n[r.id] += 1;
    }
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof_open.clear();
  return PXZ_OK;
}

pxz_status pxz_profile_enable(pxz_ctx* ctx, int on) {
  if (!ctx) return PXZ_E_ARG;
  pxz_status st = prof_drain(ctx);
  if (st != PXZ_OK) return st;
  if (on) {
    for (int i = 0; i < K_COUNT; ++i) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
  }
  ctx->profiling = on != 0;
  return PXZ_OK;
}

const char* pxz_profile_kernel_name(int kernel_id) {
  return (kernel_id >= 0 && kernel_id < K_COUNT) ? kKernelNames[kernel_id] : nullptr;
}

pxz_status pxz_profile_read(pxz_ctx* ctx, int kernel_id, double* total_ms, uint64_t* launches) {
  if (!ctx || kernel_id < 0 || kernel_id >= K_COUNT) return PXZ_E_ARG;
  pxz_status st = prof_drain(ctx);
  if (st != PXZ_OK) return st;
  if (total_ms) *total_ms = ctx->prof_ms[kernel_id];
  if (launches) *launches = ctx->prof_n[kernel_id];
  return PXZ_OK;
}

pxz_status pxz_host_alloc(size_t bytes, void** out) {
  if (!out) return PXZ_E_ARG;
  if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return PXZ_E_OOM;
  }
  return PXZ_OK;
}
void pxz_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

pxz_status pxz_grid(uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, uint32_t* cols, uint32_t* rows) {
  if (!cols || !rows || bw == 0 || bh == 0) return PXZ_E_ARG;
  *cols = ceil_div_u32(w, bw);
  *rows = ceil_div_u32(h, bh);
  return PXZ_OK;
}

// ---- images ---------------------------------------------------------------------------------------
pxz_status pxz_image_alloc(pxz_ctx* ctx, uint32_t w, uint32_t h, uint32_t channels, pxz_image** out) {
  return pxz_image_alloc_batch(ctx, w, h, channels, 1, out);
}

pxz_status pxz_image_alloc_batch(pxz_ctx* ctx, uint32_t w, uint32_t h, uint32_t channels, uint32_t n_images, pxz_image** out) {
  if (!ctx || !out) return PXZ_E_ARG;
  *out = nullptr;
  if (w == 0 || h == 0 || n_images == 0) return fail(ctx, PXZ_E_ARG, "empty image");
  if ((uint64_t)h * n_images > 0xFFFFFFFFull) return fail(ctx, PXZ_E_UNSUPPORTED, "batch taller than 2^32 rows");
  if (channels != 3 && channels != 4) return fail(ctx, PXZ_E_ARG, "channels must be 3 or 4");
  cudaSetDevice(ctx->device);
  pxz_image* im = new (std::nothrow) pxz_image();
  if (!im) return fail(ctx, PXZ_E_OOM, "host allocation failed");
  im->ctx = ctx; im->w = w; im->h = h; im->c = channels; im->owned = true; im->nimg = n_images;
  im->pitch = (((size_t)w * channels) + 127) & ~(size_t)127;  // 128-byte rows: every tile row starts 16 B aligned for RGBA
  pxz_status st = dev_alloc(ctx, (void**)&im->d, im->pitch * h * n_images);
  if (st != PXZ_OK) { delete im; return st; }
  *out = im;
  return PXZ_OK;
}

pxz_status pxz_image_upload(pxz_ctx* ctx, const uint8_t* host, uint32_t w, uint32_t h, uint32_t channels, size_t host_pitch,
                            pxz_image** out) {
  return pxz_image_upload_batch(ctx, host, w, h, channels, host_pitch, 1, out);
}

pxz_status pxz_image_upload_batch(pxz_ctx* ctx, const uint8_t* host, uint32_t w, uint32_t h, uint32_t channels, size_t host_pitch,
                                  uint32_t n_images, pxz_image** out) {
  if (!ctx || !host || !out) return PXZ_E_ARG;
  if (host_pitch < (size_t)w * channels) return fail(ctx, PXZ_E_ARG, "host pitch smaller than a row");
  pxz_status st = pxz_image_alloc_batch(ctx, w, h, channels, n_images, out);
  if (st != PXZ_OK) return st;
  // the host images follow each other without a gap (image i at host + i * h * host_pitch): one 2-D copy
  cudaError_t e = cudaMemcpy2DAsync((*out)->d, (*out)->pitch, host, host_pitch, (size_t)w * channels, (size_t)h * n_images,
                                    cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    pxz_image_free(*out);
    *out = nullptr;
    return fail(ctx, PXZ_E_CUDA, std::string("cudaMemcpy2DAsync: ") + cudaGetErrorString(e));
  }
  return PXZ_OK;
}

pxz_status pxz_image_wrap(pxz_ctx* ctx, void* device_ptr, uint32_t w, uint32_t h, uint32_t channels, size_t pitch,
                          pxz_image** out) {
  return pxz_image_wrap_batch(ctx, device_ptr, w, h, channels, pitch, 1, out);
}

pxz_status pxz_image_wrap_batch(pxz_ctx* ctx, void* device_ptr, uint32_t w, uint32_t h, uint32_t channels, size_t pitch,
                                uint32_t n_images, pxz_image** out) {
  if (!ctx || !device_ptr || !out) return PXZ_E_ARG;
  if (w == 0 || h == 0 || n_images == 0 || (channels != 3 && channels != 4) || pitch < (size_t)w * channels ||
      (uint64_t)h * n_images > 0xFFFFFFFFull)
    return fail(ctx, PXZ_E_ARG, "bad image geometry");
  pxz_image* im = new (std::nothrow) pxz_image();
  if (!im) return fail(ctx, PXZ_E_OOM, "host allocation failed");
  im->ctx = ctx; im->d = (uint8_t*)device_ptr; im->pitch = pitch; im->w = w; im->h = h; im->c = channels; im->owned = false;
  im->nimg = n_images;
  *out = im;
  return PXZ_OK;
}

pxz_status pxz_image_info(const pxz_image* img, uint32_t* w, uint32_t* h, uint32_t* channels, size_t* pitch, void** device_ptr) {
  if (!img) return PXZ_E_ARG;
  if (w) *w = img->w;
  if (h) *h = img->h;
  if (channels) *channels = img->c;
  if (pitch) *pitch = img->pitch;
  if (device_ptr) *device_ptr = img->d;
  return PXZ_OK;
}

pxz_status pxz_image_download(pxz_ctx* ctx, const pxz_image* img, uint8_t* host, size_t host_pitch) {
  if (!ctx || !img || !host) return PXZ_E_ARG;
  if (host_pitch < (size_t)img->w * img->c) return fail(ctx, PXZ_E_ARG, "host pitch smaller than a row");
  PXZ_CUDA(ctx, cudaMemcpy2DAsync(host, host_pitch, img->d, img->pitch, (size_t)img->w * img->c, (size_t)img->h * img->nimg,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PXZ_OK;
}

void pxz_image_free(pxz_image* img) {
  if (!img) return;
  if (img->owned) dev_free(img->ctx, img->d);
  delete img;
}

// ---- analysis -----------------------------------------------------------------------------------------
static pxz_status run_analysis(pxz_ctx* ctx, const pxz_image* img, const Geom& g, pxz_metric metric, bool exact_all,
                               const ValueMap* vm_for_band) {
  const uint32_t nblocks = g.cols * g.rows;
  pxz_status st = ensure_scratch(ctx, nblocks);
  if (st != PXZ_OK) return st;
  if (metric == PXZ_METRIC_OKLAB_MAD) {
    if (exact_all) {
      ProfScope prof(ctx, K_MAD_EXACT);
      PXZ_CUDA(ctx, launch_analyze_mad_exact(img->d, img->pitch, g, ctx->d_vx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                             ctx->d_list, ctx->d_list + nblocks, ctx->stream, ctx->sm_count, &ctx->launches));
    } else {
      {
        ProfScope prof(ctx, K_MAD_FAST);
        PXZ_CUDA(ctx, launch_analyze_mad_fast(img->d, img->pitch, g, ctx->d_vx, ctx->d_opaque, ctx->d_list + nblocks, ctx->stream,
                                              ctx->sm_count, &ctx->launches));
      }
      if (vm_for_band) {
        // recompute, in reference order, the tiles whose level could differ from the CPU result
        ProfScope prof(ctx, K_MAD_EXACT);
        PXZ_CUDA(ctx, launch_analyze_mad_exact(img->d, img->pitch, g, ctx->d_vx, ctx->d_vx, ctx->d_opaque, vm_for_band, &ctx->thr,
                                               &ctx->band, ctx->d_minmax, ctx->d_list, ctx->d_list + nblocks, ctx->stream,
                                               ctx->sm_count, &ctx->launches));
      }
    }
  } else if (metric == PXZ_METRIC_SOBEL_DIR) {
    // operations.rs:220-221: `height - 2` / `width - 2` underflow -> the reference panics
    const uint32_t min_w = g.trail_w ? std::min(g.bw, g.trail_w) : std::min(g.bw, g.W);
    const uint32_t min_h = g.trail_h ? std::min(g.bh, g.trail_h) : std::min(g.bh, g.H);
    if (min_w < 2 || min_h < 2)
      return fail(ctx, PXZ_E_ARG, "directional metric needs every block to be at least 2x2 (the reference panics)");
    ProfScope prof(ctx, K_SOBEL);
    PXZ_CUDA(ctx, launch_analyze_sobel(img->d, img->pitch, g, ctx->d_vx, ctx->d_vy, ctx->stream, ctx->sm_count, &ctx->launches));
  } else {
    return fail(ctx, PXZ_E_ARG, "unknown metric");
  }
  return PXZ_OK;
}

pxz_status pxz_analyze(pxz_ctx* ctx, const pxz_image* img, uint32_t bw, uint32_t bh, pxz_metric metric, uint32_t flags,
                       float* host_values_x, float* host_values_y) {
  if (!ctx || !img || !host_values_x) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  Geom g;
  pxz_status st = make_geom(ctx, img->w, img->h, img->c, bw, bh, &g, img->nimg);
  if (st != PXZ_OK) return st;
  st = run_analysis(ctx, img, g, metric, (flags & PXZ_FLAG_EXACT_VALUES) != 0, nullptr);
  if (st != PXZ_OK) return st;
  const size_t n = (size_t)g.cols * g.rows;
  PXZ_CUDA(ctx, cudaMemcpyAsync(host_values_x, ctx->d_vx, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (host_values_y) {
    const float* src = metric == PXZ_METRIC_SOBEL_DIR ? ctx->d_vy : ctx->d_vx;
    PXZ_CUDA(ctx, cudaMemcpyAsync(host_values_y, src, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PXZ_OK;
}

// ---- encode -------------------------------------------------------------------------------------------
static pxz_status payload_new(pxz_ctx* ctx, const Geom& g, uint64_t capacity, pxz_payload** out) {
  pxz_payload* p = new (std::nothrow) pxz_payload();
  if (!p) return fail(ctx, PXZ_E_OOM, "host allocation failed");
  p->ctx = ctx;
  p->g = g;
  const size_t nblocks = (size_t)g.cols * g.rows;
  // a cached set that is large enough (smallest fit)
  int best = -1;
  for (size_t i = 0; i < ctx->payload_cache.size(); ++i) {
    const PayloadBufs& b = ctx->payload_cache[i];
    if (b.nblocks >= nblocks && b.capacity >= capacity && (best < 0 || b.capacity < ctx->payload_cache[best].capacity)) best = (int)i;
  }
  if (best >= 0) {
    const PayloadBufs b = ctx->payload_cache[best];
    ctx->payload_cache.erase(ctx->payload_cache.begin() + best);
    p->d_descs = b.d_descs; p->d_tabidx = b.d_tabidx; p->d_pixels = b.d_pixels; p->d_total = b.d_total;
    p->nblocks_cap = b.nblocks; p->capacity = b.capacity;
    *out = p;
    return PXZ_OK;
  }
  p->capacity = capacity;
  p->nblocks_cap = nblocks;
  pxz_status st;
  if ((st = dev_alloc(ctx, (void**)&p->d_descs, nblocks * sizeof(pxz_block_desc))) != PXZ_OK ||
      (st = dev_alloc(ctx, (void**)&p->d_tabidx, (2 * nblocks + order_list_words(nblocks)) * 4)) != PXZ_OK ||  // table indices | work order lists | expand-side indices
      (st = dev_alloc(ctx, (void**)&p->d_pixels, capacity)) != PXZ_OK ||
      (st = dev_alloc(ctx, (void**)&p->d_total, 8)) != PXZ_OK) {
    // partial allocations must not enter the cache
    dev_free(ctx, p->d_descs); dev_free(ctx, p->d_tabidx); dev_free(ctx, p->d_pixels); dev_free(ctx, p->d_total);
    p->d_descs = nullptr; p->d_tabidx = nullptr; p->d_pixels = nullptr; p->d_total = nullptr;
    delete p;
    return st;
  }
  *out = p;
  return PXZ_OK;
}

// The one exchange of the path (PXZ_FLAG_NORMALISE_GLOBAL with a communicator): ncclMin over {min_x, -max_x, min_y, -max_y,
// ok}.  `ok` is 0 on a healthy rank and -1 on a rank that failed before it got here — such a rank still joins, with
// neutral values, so that its peers do not wait in the all-reduce forever; everybody then learns that a rank failed.
// `local_ok == false`: this rank contributes nothing (an error before the exchange, or a shard without rows).
static pxz_status exchange_minmax(pxz_ctx* ctx, bool local_ok, bool* peers_ok) {
  *peers_ok = true;
  if (!ctx->comm) return PXZ_OK;
  if (local_ok) {
    PXZ_CUDA(ctx, cudaMemsetAsync(ctx->d_minmax + 4, 0, sizeof(float), ctx->stream));
  } else {
    static const float neutral[5] = {INFINITY, INFINITY, INFINITY, INFINITY, -1.0f};
    PXZ_CUDA(ctx, cudaMemcpyAsync(ctx->d_minmax, neutral, sizeof(neutral), cudaMemcpyHostToDevice, ctx->stream));
  }
  std::string err;
  if (nccl_allreduce_min_f32(ctx->comm, ctx->d_minmax, 5, ctx->stream, &err) != 0) return fail(ctx, PXZ_E_NCCL, err);
  float flag = 0.f;
  PXZ_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_minmax + 4, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *peers_ok = !(flag < 0.f);
  return PXZ_OK;
}

pxz_status pxz_comm_join_empty(pxz_ctx* ctx) {
  if (!ctx) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->comm) return fail(ctx, PXZ_E_ARG, "no communicator (pxz_comm_init)");
  // a rank whose shard has no rows: neutral values, but a healthy flag
  static const float neutral[5] = {INFINITY, INFINITY, INFINITY, INFINITY, 0.0f};
  PXZ_CUDA(ctx, cudaMemcpyAsync(ctx->d_minmax, neutral, sizeof(neutral), cudaMemcpyHostToDevice, ctx->stream));
  std::string err;
  if (nccl_allreduce_min_f32(ctx->comm, ctx->d_minmax, 5, ctx->stream, &err) != 0) return fail(ctx, PXZ_E_NCCL, err);
  float flag = 0.f;
  PXZ_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_minmax + 4, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag < 0.f) return fail(ctx, PXZ_E_NCCL, "another rank failed before the min/max exchange");
  return PXZ_OK;
}

// RGB images whose geometry suits the RGBA fast kernels run on them (kernels.cu "RGB images on the RGBA fast kernels")
static bool rgb_on_rgba(const pxz_ctx* ctx, uint32_t w, uint32_t c, uint32_t bw, uint32_t bh) {
  return c == 3 && ctx->rgb_via_rgba && ctx->resize_semantics == PXZ_RESIZE_IMAGE_RS && bw <= 64 && bh <= 64 && bw % 4 == 0 && w % 4 == 0;
}

// payload of the other channel count (3 <-> 4) with the same blocks, tables and work order
static pxz_status payload_rechannel(pxz_ctx* ctx, const pxz_payload* src, uint32_t dc, pxz_payload** out) {
  *out = nullptr;
  Geom g = src->g;
  const uint32_t sc = g.C;
  g.C = dc;
  const uint32_t nblocks = g.cols * g.rows;
  pxz_payload* p = nullptr;
  pxz_status st = payload_new(ctx, g, (uint64_t)g.W * g.H * dc * g.nimg, &p);
  if (st != PXZ_OK) return st;
  p->spec = src->spec;
  p->strategy = src->strategy;
  p->max_small_px = src->max_small_px; p->max_small_dim = src->max_small_dim;
  p->max_tmp_down = src->max_tmp_down; p->max_tmp_up = src->max_tmp_up;
  if (src->bytes_known) { p->bytes = src->bytes / sc * dc; p->bytes_known = true; }
  cudaError_t e = launch_payload_convert(src->d_descs, src->d_pixels, src->d_tabidx, (uint32_t)src->nblocks_cap, src->d_total, p->d_descs,
                                         p->d_pixels, p->d_tabidx, (uint32_t)p->nblocks_cap, p->d_total, nblocks, (int)sc, (int)dc,
                                         ctx->stream, ctx->sm_count, &ctx->launches);
  if (e != cudaSuccess) {
    cudaGetLastError();
    payload_release(p);
    return fail(ctx, PXZ_E_CUDA, std::string("payload convert: ") + cudaGetErrorString(e));
  }
  *out = p;
  return PXZ_OK;
}

pxz_status pxz_shrink(pxz_ctx* ctx, const pxz_image* img, uint32_t bw, uint32_t bh, pxz_metric metric, float factor,
                      pxz_filter filter_down, uint32_t flags, pxz_payload** out) {
  if (!ctx || !img || !out) return PXZ_E_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  if (rgb_on_rgba(ctx, img->w, img->c, bw, bh)) {
    // widen to RGBA (alpha 255: its metric term is exactly +0 and the opaque resample path never touches it), run the
    // 4-channel pipeline, narrow the payload back
    pxz_image* wide = nullptr;
    pxz_status st = pxz_image_alloc_batch(ctx, img->w, img->h, 4, img->nimg, &wide);
    if (st != PXZ_OK) return st;
    cudaError_t e = launch_rgb_widen(img->d, img->pitch, wide->d, wide->pitch, img->w, img->h * img->nimg, 1, ctx->stream, ctx->sm_count,
                                     &ctx->launches);
    pxz_payload* p4 = nullptr;
    if (e != cudaSuccess) {
      cudaGetLastError();
      st = fail(ctx, PXZ_E_CUDA, std::string("rgb widen: ") + cudaGetErrorString(e));
    } else {
      st = pxz_shrink(ctx, wide, bw, bh, metric, factor, filter_down, flags, &p4);
    }
    pxz_image_free(wide);  // stream-ordered: the kernels queued above keep it
    if (st != PXZ_OK) return st;
    st = payload_rechannel(ctx, p4, 3, out);
    payload_release(p4);
    return st;
  }
  const bool normalise = (flags & PXZ_FLAG_NORMALISE_GLOBAL) != 0;
  // Every rank of a communicator must reach the min/max exchange, whatever happens to it before: a rank-local failure
  // (an unknown filter, a trailing block the Sobel metric cannot take, no memory) joins with neutral values and an
  // error flag instead of leaving its peers in the collective.
  auto bail = [&](pxz_status why) -> pxz_status {
    if (normalise && ctx->comm) {
      const std::string msg = ctx->err;
      bool peers_ok = true;
      exchange_minmax(ctx, false, &peers_ok);
      ctx->err = msg;
    }
    return why;
  };
  if ((int)filter_down < 0 || (int)filter_down > 4) return bail(fail(ctx, PXZ_E_ARG, "unknown filter"));
  Geom g;
  pxz_status st = make_geom(ctx, img->w, img->h, img->c, bw, bh, &g, img->nimg);
  if (st != PXZ_OK) return bail(st);
  if (normalise && img->nimg > 1) return bail(fail(ctx, PXZ_E_UNSUPPORTED, "global normalisation is per image: not available on a batch"));
  const uint32_t nblocks = g.cols * g.rows;

  ValueMap vm;
  vm.factor = factor;
  vm.mode = (metric == PXZ_METRIC_SOBEL_DIR) ? 2 : ((flags & PXZ_FLAG_AFTER_IDENTITY) ? 1 : 0);
  vm.normalise = normalise ? 1 : 0;
  vm.extra_thr = NAN;
  if (ctx->strategy.on)
    for (int k = 1; k < kStrategyBuckets; ++k)
      if (ctx->strategy.down[k] != ctx->strategy.down[k - 1] || ctx->strategy.up[k] != ctx->strategy.up[k - 1])
        vm.bucket_edges |= 1ull << (k - 1);

  // with global normalisation every value depends on the exact min and max, so the Oklab path
  // runs in reference order for all blocks
  const bool exact_all = (flags & PXZ_FLAG_EXACT_VALUES) != 0;
  // Global normalisation of the Oklab values on the fast path: the exact extremes come from the few tiles that can hold
  // them, and the guard band (scaled by 1 / range) then decides as in the default mode which tiles need their
  // reference-order value; with PXZ_FLAG_EXACT_VALUES every tile is computed in reference order as before.
  const bool fast_norm = normalise && metric == PXZ_METRIC_OKLAB_MAD && !exact_all;
  st = run_analysis(ctx, img, g, metric, exact_all, (metric == PXZ_METRIC_OKLAB_MAD && !exact_all && !normalise) ? &vm : nullptr);
  if (st != PXZ_OK) return bail(st);

  if (normalise) {
    auto minmax = [&]() -> pxz_status {
      ProfScope prof(ctx, K_MINMAX);
      cudaError_t me = launch_minmax(ctx->d_vx, metric == PXZ_METRIC_SOBEL_DIR ? ctx->d_vy : nullptr, nblocks, ctx->d_minmax,
                                     ctx->stream, &ctx->launches);
      if (me != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, PXZ_E_CUDA, std::string("minmax: ") + cudaGetErrorString(me));
      }
      return PXZ_OK;
    };
    if ((st = minmax()) != PXZ_OK) return bail(st);
    if (fast_norm) {
      {
        ProfScope prof(ctx, K_MAD_EXACT);
        cudaError_t ee = launch_analyze_mad_exact(img->d, img->pitch, g, ctx->d_vx, ctx->d_vx, ctx->d_opaque, nullptr, &ctx->thr, &ctx->band,
                                                  ctx->d_minmax, ctx->d_list, ctx->d_list + nblocks, ctx->stream, ctx->sm_count,
                                                  &ctx->launches);
        if (ee != cudaSuccess) {
          cudaGetLastError();
          return bail(fail(ctx, PXZ_E_CUDA, std::string("extremes: ") + cudaGetErrorString(ee)));
        }
      }
      if ((st = minmax()) != PXZ_OK) return bail(st);  // over the patched values: the exact extremes
    }
    bool peers_ok = true;
    st = exchange_minmax(ctx, true, &peers_ok);
    if (st != PXZ_OK) return st;
    if (!peers_ok) return fail(ctx, PXZ_E_NCCL, "another rank failed before the min/max exchange");
    if (fast_norm) {
      ProfScope prof(ctx, K_MAD_EXACT);
      cudaError_t ee = cudaMemsetAsync(ctx->d_list + nblocks, 0, sizeof(uint32_t), ctx->stream);  // the list counter
      if (ee == cudaSuccess)
        ee = launch_analyze_mad_exact(img->d, img->pitch, g, ctx->d_vx, ctx->d_vx, ctx->d_opaque, &vm, &ctx->thr, &ctx->band, ctx->d_minmax,
                                      ctx->d_list, ctx->d_list + nblocks, ctx->stream, ctx->sm_count, &ctx->launches);
      if (ee != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, PXZ_E_CUDA, std::string("guard band: ") + cudaGetErrorString(ee));
      }
    }
  }

  pxz_payload* p = nullptr;
  st = payload_new(ctx, g, (uint64_t)g.W * g.H * g.C * g.nimg, &p);
  if (st != PXZ_OK) return st;
  p->spec = spec_for_geom(g);
  {
    // shared-memory bounds of the two resample directions for this geometry
    const bool coupled = metric == PXZ_METRIC_OKLAB_MAD;  // both axes share the level
    const uint32_t tw_max = g.bw, th_max = g.bh;
    auto pad8 = [](uint32_t v) { return (v + 7u) & ~7u; };  // intermediate rows are padded to 8 columns
    p->max_small_px = tw_max * th_max;
    p->max_small_dim = std::max(tw_max, th_max);
    if (coupled) {
      p->max_tmp_down = ((th_max + 1) / 2) * pad8(tw_max);  // dh <= ceil(th/2), sw = tw
      p->max_tmp_up = th_max * pad8((tw_max + 1) / 2);      // dh = th, sw <= ceil(tw/2)
      // a 1-px-wide/-high trailing tile keeps that axis while the other shrinks: still within the bounds
    } else {
      p->max_tmp_down = th_max * pad8(tw_max);
      p->max_tmp_up = th_max * pad8(tw_max);
    }
  }
  cudaError_t e;
  {
    ProfScope prof(ctx, K_PLAN);
    StrategyLut lut = ctx->strategy;
    lut.stride = (uint32_t)p->spec->n_small.size();
    p->strategy = lut.on ? 1 : 0;
    e = launch_plan(ctx->d_vx, metric == PXZ_METRIC_SOBEL_DIR ? ctx->d_vy : nullptr, g, vm, ctx->d_minmax, ctx->thr, nullptr,
                    p->d_descs, p->d_tabidx, p->d_total, ctx->d_scan, p->d_order(), (uint32_t)p->nblocks_cap, ctx->stream, &ctx->launches,
                    &lut, p->d_up());
  }
  if (e != cudaSuccess) {
    payload_release(p);
    return fail(ctx, PXZ_E_CUDA, std::string("plan: ") + cudaGetErrorString(e));
  }
  TabSet ts;
  st = get_tabset(ctx, *p->spec, p->strategy ? -1 : (int)filter_down, 0, &ts);
  // the fast Oklab-MAD pass also told which tiles are fully opaque: their alpha channel needs no arithmetic
  const uint8_t* opaque_flags = (metric == PXZ_METRIC_OKLAB_MAD && !exact_all) ? ctx->d_opaque : nullptr;
  if (st == PXZ_OK)
    st = run_resample(ctx, 0, img->d, img->pitch, p, ts, g.bw * g.bh, std::max(g.bw, g.bh), p->max_tmp_down, opaque_flags);
  if (st != PXZ_OK) {
    payload_release(p);
    return st;
  }
  *out = p;
  return PXZ_OK;
}

// host restatement of the plan kernel's scalar path (same threshold table)
pxz_status pxz_reduce_dims(float v0, float v1, uint32_t w, uint32_t h, uint32_t* out_w, uint32_t* out_h, float* stored) {
  if (!out_w || !out_h || w == 0 || h == 0) return PXZ_E_ARG;
  static const LevelThresholds thr = [] {  // function-local static: initialised once, thread-safe
    LevelThresholds t;
    build_level_thresholds(&t);
    return t;
  }();
  auto parse = [](float v) -> float {  // operations.rs:128-138
    if (!signbit(v)) return v;
    float x = 1.0f + v;
    return (x != x) ? 0.0f : (x > 0.0f ? x : 0.0f);
  };
  auto level = [&](float pv) -> uint32_t {
    if (pv != pv) return 0;
    if (pv >= thr.thr[0]) return 0;
    for (int k = 1; k < kThresholds; ++k)
      if (pv >= thr.thr[k]) return (uint32_t)k;
    return kLevelOnePixel;
  };
  auto dim = [](uint32_t n, uint32_t k) -> uint32_t {
    if (k >= 32) return 1;
    uint32_t d = (uint32_t)(((uint64_t)n + ((1ull << k) - 1)) >> k);
    return d < 1 ? 1 : d;
  };
  const float p0 = parse(v0), p1 = parse(v1);
  *out_w = dim(w, level(p0));
  *out_h = dim(h, level(p1));
  if (stored) *stored = (isinf(p0) || isinf(p1)) ? INFINITY : (float)sqrt((double)p0 * (double)p0 + (double)p1 * (double)p1);
  return PXZ_OK;
}

int32_t pxz_resample_table(uint32_t n_in, uint32_t n_out, pxz_filter filter, uint32_t* left, uint32_t* count, float* weights,
                           uint32_t max_taps) {
  if (!left || !count || !weights || n_in == 0 || n_out == 0) return PXZ_E_ARG;
  std::vector<uint32_t> pool;
  AxisTab t;
  if (!build_axis_table(n_in, n_out, (int)filter, &pool, &t)) return PXZ_E_ARG;
  if (t.stride > max_taps) return (int32_t)t.stride;
  for (uint32_t o = 0; o < n_out; ++o) {
    left[o] = pool[t.off + o];
    count[o] = pool[t.off + n_out + o];
    for (uint32_t i = 0; i < max_taps; ++i) {
      uint32_t bits = i < t.stride ? pool[t.off + 2 * n_out + (size_t)o * t.stride + i] : 0u;
      memcpy(&weights[(size_t)o * max_taps + i], &bits, 4);
    }
  }
  return (int32_t)t.stride;
}

// ---- payload ------------------------------------------------------------------------------------------
static pxz_status payload_bytes(pxz_ctx* ctx, const pxz_payload* cp, uint64_t* bytes) {
  pxz_payload* p = const_cast<pxz_payload*>(cp);
  if (!p->bytes_known) {
    PXZ_CUDA(ctx, cudaMemcpyAsync(ctx->h_total, p->d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->bytes = *ctx->h_total;
    p->bytes_known = true;
  }
  *bytes = p->bytes;
  return PXZ_OK;
}

pxz_status pxz_payload_info(pxz_ctx* ctx, const pxz_payload* p, uint32_t* w, uint32_t* h, uint32_t* bw, uint32_t* bh,
                            uint32_t* cols, uint32_t* rows, uint32_t* channels, uint64_t* bytes) {
  if (!ctx || !p) return PXZ_E_ARG;
  if (w) *w = p->g.W;
  if (h) *h = p->g.H;
  if (bw) *bw = p->g.bw;
  if (bh) *bh = p->g.bh;
  if (cols) *cols = p->g.cols;
  if (rows) *rows = p->g.rows_img;  /* per image; pxz_payload_batch_count() images */
  if (channels) *channels = p->g.C;
  if (bytes) return payload_bytes(ctx, p, bytes);
  return PXZ_OK;
}

pxz_status pxz_payload_download(pxz_ctx* ctx, const pxz_payload* p, pxz_block_desc* host_descs, uint8_t* host_pixels) {
  if (!ctx || !p || !host_descs || !host_pixels) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  uint64_t bytes = 0;
  pxz_status st = payload_bytes(ctx, p, &bytes);
  if (st != PXZ_OK) return st;
  const size_t nblocks = (size_t)p->g.cols * p->g.rows;
  PXZ_CUDA(ctx, cudaMemcpyAsync(host_descs, p->d_descs, nblocks * sizeof(pxz_block_desc), cudaMemcpyDeviceToHost, ctx->stream));
  if (bytes) PXZ_CUDA(ctx, cudaMemcpyAsync(host_pixels, p->d_pixels, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PXZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PXZ_OK;
}

// pixels == NULL with copy_pixels == false: the caller fills p->d_pixels on the device (pxz_payload_from_container)
static pxz_status payload_from_descs(pxz_ctx* ctx, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, uint32_t channels,
                                     const pxz_block_desc* descs, const uint8_t* pixels, bool copy_pixels, uint64_t bytes,
                                     pxz_payload** out, uint32_t nimg = 1) {
  *out = nullptr;
  cudaSetDevice(ctx->device);
  Geom g;
  pxz_status st = make_geom(ctx, w, h, channels, bw, bh, &g, nimg);
  if (st != PXZ_OK) return st;
  // the decoder sizes the grid in f32 (encoding/mod.rs:118-119); identical to the integer ceil below 2^24
  const size_t nblocks = (size_t)g.cols * g.rows;
  // distinct (reduced, tile) size pairs -> table indices
  auto spec = std::make_shared<TabSpec>();
  std::map<std::pair<uint32_t, uint32_t>, uint32_t> index;
  std::vector<uint32_t> tabidx(nblocks);
  std::vector<uint8_t> filt(ctx->strategy.on ? nblocks : 0);  // per-block expand filter (pxz_ctx_set_strategy)
  uint32_t max_small = 1, max_tmp_up = 1, max_tmp_down = 1, max_dim = 1;
  auto pad8 = [](uint32_t v) { return (v + 7u) & ~7u; };
  auto idx_of = [&](uint32_t small, uint32_t tile) -> uint32_t {
    auto it = index.find({small, tile});
    if (it != index.end()) return it->second;
    const uint32_t id = (uint32_t)spec->n_small.size();
    spec->n_small.push_back(small);
    spec->n_tile.push_back(tile);
    index[{small, tile}] = id;
    return id;
  };
  for (size_t b = 0; b < nblocks; ++b) {
    const uint32_t by = (uint32_t)((b / g.cols) % g.rows_img), bx = (uint32_t)(b % g.cols);
    // target size as Pixlzr::expand computes it (pixlzr.rs:92-107)
    const uint32_t tw = (bx == g.cols - 1 && g.trail_w) ? g.trail_w : g.bw;
    const uint32_t th = (by == g.rows_img - 1 && g.trail_h) ? g.trail_h : g.bh;
    const pxz_block_desc& d = descs[b];
    if (d.w == 0 || d.h == 0) return fail(ctx, PXZ_E_ARG, "block with zero size");
    const uint64_t sz = (uint64_t)d.w * d.h * channels;
    if (d.offset > bytes || sz > bytes - d.offset) return fail(ctx, PXZ_E_ARG, "block descriptor points outside the payload");
    if (channels == 4 && (d.offset & 3u)) return fail(ctx, PXZ_E_ARG, "RGBA block offsets must be 4-byte aligned");
    const uint32_t ix = idx_of(d.w, tw), iy = idx_of(d.h, th);
    if (ix > 0xFFFFu || iy > 0xFFFFu) return fail(ctx, PXZ_E_UNSUPPORTED, "more than 65536 distinct block sizes");
    tabidx[b] = ix | (iy << 16);
    if (ctx->strategy.on) filt[b] = ctx->strategy.up[strategy_bucket(d.value)];
    max_small = std::max(max_small, (uint32_t)d.w * d.h);
    max_dim = std::max(max_dim, (uint32_t)std::max(d.w, d.h));
    max_tmp_up = std::max(max_tmp_up, th * pad8(d.w));
    max_tmp_down = std::max(max_tmp_down, (uint32_t)d.h * pad8(tw));
  }
  pxz_payload* p = nullptr;
  st = payload_new(ctx, g, bytes, &p);
  if (st != PXZ_OK) return st;
  p->spec = spec;
  if (ctx->strategy.on) {
    const uint32_t per_filter = (uint32_t)spec->n_small.size();
    if (per_filter * 5u > 0x10000u) {
      payload_release(p);
      return fail(ctx, PXZ_E_UNSUPPORTED, "too many distinct block sizes for per-block filters");
    }
    for (size_t b = 0; b < nblocks; ++b) tabidx[b] += (filt[b] * per_filter) * 0x10001u;
    p->strategy = 2;
  }
  p->max_small_px = max_small;
  p->max_small_dim = max_dim;
  p->max_tmp_up = max_tmp_up;
  p->max_tmp_down = max_tmp_down;
  p->bytes = bytes;
  p->bytes_known = true;
  cudaError_t e = cudaMemcpyAsync(p->d_descs, descs, nblocks * sizeof(pxz_block_desc), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_tabidx, tabidx.data(), nblocks * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && bytes && copy_pixels) e = cudaMemcpyAsync(p->d_pixels, pixels, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_total, &p->bytes, 8, cudaMemcpyHostToDevice, ctx->stream);
  // work order of the expand kernel (blocks by cost class)
  if (e == cudaSuccess && ensure_scratch(ctx, (uint32_t)nblocks) != PXZ_OK) e = cudaErrorMemoryAllocation;
  if (e == cudaSuccess) e = launch_class_lists(p->d_descs, g, ctx->d_scan, p->d_order(), (uint32_t)p->nblocks_cap, ctx->stream, &ctx->launches);
  // tabidx is a pageable temporary and `pixels` may be reused by the caller right away
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    payload_release(p);
    return fail(ctx, PXZ_E_CUDA, std::string("payload upload: ") + cudaGetErrorString(e));
  }
  *out = p;
  return PXZ_OK;
}

pxz_status pxz_payload_upload(pxz_ctx* ctx, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, uint32_t channels,
                              const pxz_block_desc* descs, const uint8_t* pixels, uint64_t bytes, pxz_payload** out) {
  if (!ctx || !descs || (!pixels && bytes) || !out) return PXZ_E_ARG;
  return payload_from_descs(ctx, w, h, bw, bh, channels, descs, pixels, true, bytes, out);
}

pxz_status pxz_payload_upload_batch(pxz_ctx* ctx, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, uint32_t channels,
                                    uint32_t n_images, const pxz_block_desc* descs, const uint8_t* pixels, uint64_t bytes,
                                    pxz_payload** out) {
  if (!ctx || !descs || (!pixels && bytes) || !out || n_images == 0) return PXZ_E_ARG;
  return payload_from_descs(ctx, w, h, bw, bh, channels, descs, pixels, true, bytes, out, n_images);
}

uint32_t pxz_image_batch_count(const pxz_image* img) { return img ? img->nimg : 0; }
uint32_t pxz_payload_batch_count(const pxz_payload* p) { return p ? p->g.nimg : 0; }

pxz_status pxz_shrink_batch(pxz_ctx* ctx, const pxz_image* batch, uint32_t bw, uint32_t bh, pxz_metric metric, float factor,
                            pxz_filter filter_down, uint32_t flags, pxz_payload** out) {
  // one launch per stage over the tiles of all images: the batch is one block grid (pxz_internal.h, Geom)
  return pxz_shrink(ctx, batch, bw, bh, metric, factor, filter_down, flags, out);
}

pxz_status pxz_expand_batch(pxz_ctx* ctx, const pxz_payload* p, pxz_filter filter_up, pxz_image* out_batch) {
  return pxz_expand_to_image(ctx, p, filter_up, out_batch);
}

// ---- container stage on the device (qoi_device.cu) ------------------------------------------------------------
pxz_status pxz_payload_to_container(pxz_ctx* ctx, const pxz_payload* p, uint32_t filter_byte, int values_present, uint8_t* host_out,
                                    size_t cap, uint64_t* bytes_out) {
  if (!ctx || !p || !host_out || !bytes_out) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  if (p->g.nimg != 1) return fail(ctx, PXZ_E_UNSUPPORTED, "a container holds one image: not available on a batch");
  uint64_t bytes = 0;
  pxz_status st = payload_bytes(ctx, p, &bytes);
  if (st != PXZ_OK) return st;
  const Geom& g = p->g;
  const uint32_t nblocks = g.cols * g.rows;
  const size_t arena_bytes = qoi_arena_bytes(nblocks, bytes, g.C);
  const size_t out_cap = 26 + (size_t)4 * g.rows + arena_bytes;
  uint8_t *d_arena = nullptr, *d_out = nullptr;
  uint32_t* d_len = nullptr;
  unsigned long long *d_off = nullptr, *d_total = nullptr;
  auto release = [&]() { dev_free(ctx, d_arena); dev_free(ctx, d_out); dev_free(ctx, d_len); dev_free(ctx, d_off); dev_free(ctx, d_total); };
  if ((st = dev_alloc(ctx, (void**)&d_arena, arena_bytes)) != PXZ_OK || (st = dev_alloc(ctx, (void**)&d_out, out_cap)) != PXZ_OK ||
      (st = dev_alloc(ctx, (void**)&d_len, (size_t)nblocks * 4)) != PXZ_OK || (st = dev_alloc(ctx, (void**)&d_off, (size_t)nblocks * 8)) != PXZ_OK ||
      (st = dev_alloc(ctx, (void**)&d_total, 8)) != PXZ_OK) {
    release();
    return st;
  }
  cudaError_t e;
  {
    ProfScope prof(ctx, K_QOI_ENCODE);
    e = launch_qoi_encode(p->d_descs, p->d_pixels, g, values_present, filter_byte, d_arena, d_len, d_off, d_out, d_total, ctx->stream,
                          &ctx->launches);
  }
  unsigned long long total = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && total <= cap) {
    e = cudaMemcpyAsync(host_out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  release();
  if (e != cudaSuccess) return fail(ctx, PXZ_E_CUDA, std::string("container encode: ") + cudaGetErrorString(e));
  *bytes_out = total;
  if (total > cap) return fail(ctx, PXZ_E_ARG, "output buffer too small (size it with pxz_container_bound)");
  return PXZ_OK;
}

pxz_status pxz_payload_from_container(pxz_ctx* ctx, const uint8_t* data, size_t len, int32_t* filter_byte, pxz_payload** out) {
  if (!ctx || !data || !out) return PXZ_E_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  uint32_t w, h, bw, bh, ch;
  int32_t filt = -1;
  uint64_t bytes = 0;
  // the host walks the headers (a few bytes per block); the streams themselves are decoded on the device
  pxz_status st = pxz_container_decode(data, len, &w, &h, &bw, &bh, &filt, &ch, &bytes, nullptr, nullptr);
  if (st != PXZ_OK) return fail(ctx, st, "malformed .pxlzr container");
  Geom g;
  if ((st = make_geom(ctx, w, h, ch, bw, bh, &g)) != PXZ_OK) return st;
  const size_t nblocks = (size_t)g.cols * g.rows;  // == the decoder's f32 grid: pxz_container_decode refuses files where they differ
  std::vector<pxz_block_desc> descs(nblocks);
  st = pxz_container_decode(data, len, &w, &h, &bw, &bh, &filt, &ch, &bytes, descs.data(), nullptr);
  if (st != PXZ_OK) return fail(ctx, st, "malformed .pxlzr container");
  std::vector<unsigned long long> in_off(nblocks);
  std::vector<uint32_t> qlen(nblocks);
  {
    size_t q = 26 + (size_t)4 * g.rows;  // constants.rs:19-20 + the line table; validated by the walk above
    for (size_t b = 0; b < nblocks; ++b) {
      if (q + 13 > len) return fail(ctx, PXZ_E_FORMAT, "malformed .pxlzr container");  // cannot happen after the walk above
      const uint8_t* hd = data + q + 9;
      qlen[b] = (uint32_t)hd[0] << 24 | (uint32_t)hd[1] << 16 | (uint32_t)hd[2] << 8 | hd[3];
      in_off[b] = q + 13;
      q += 13 + (size_t)qlen[b];
      if (q > len) return fail(ctx, PXZ_E_FORMAT, "malformed .pxlzr container");
    }
  }
  for (size_t b = 0; b < nblocks; ++b)
    if (ch == 4 && (descs[b].offset & 3u)) return fail(ctx, PXZ_E_UNSUPPORTED, "unaligned RGBA block");  // cannot happen: w*h*4
  pxz_payload* p = nullptr;
  st = payload_from_descs(ctx, w, h, bw, bh, ch, descs.data(), nullptr, false, bytes, &p);
  if (st != PXZ_OK) return st;
  uint8_t* d_in = nullptr;
  unsigned long long* d_off = nullptr;
  uint32_t* d_len = nullptr;
  int* d_err = nullptr;
  auto release = [&]() { dev_free(ctx, d_in); dev_free(ctx, d_off); dev_free(ctx, d_len); dev_free(ctx, d_err); };
  if ((st = dev_alloc(ctx, (void**)&d_in, len)) != PXZ_OK || (st = dev_alloc(ctx, (void**)&d_off, nblocks * 8)) != PXZ_OK ||
      (st = dev_alloc(ctx, (void**)&d_len, nblocks * 4)) != PXZ_OK || (st = dev_alloc(ctx, (void**)&d_err, 4)) != PXZ_OK) {
    release();
    payload_release(p);
    return st;
  }
  int err = 0;
  cudaError_t e = cudaMemcpyAsync(d_in, data, len, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, in_off.data(), nblocks * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_len, qlen.data(), nblocks * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_err, 0, 4, ctx->stream);
  if (e == cudaSuccess) {
    ProfScope prof(ctx, K_QOI_DECODE);
    e = launch_qoi_decode(d_in, d_off, d_len, p->d_descs, g, p->d_pixels, d_err, ctx->stream, &ctx->launches);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  release();
  if (e != cudaSuccess || err) {
    payload_release(p);
    return e != cudaSuccess ? fail(ctx, PXZ_E_CUDA, std::string("container decode: ") + cudaGetErrorString(e))
                            : fail(ctx, PXZ_E_FORMAT, "truncated QOI stream");
  }
  if (filter_byte) *filter_byte = filt;
  *out = p;
  return PXZ_OK;
}

void pxz_payload_free(pxz_payload* p) { payload_release(p); }

// ---- decode -------------------------------------------------------------------------------------------
pxz_status pxz_expand_to_image(pxz_ctx* ctx, const pxz_payload* p, pxz_filter filter_up, pxz_image* out) {
  if (!ctx || !p || !out) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  if ((int)filter_up < 0 || (int)filter_up > 4) return fail(ctx, PXZ_E_ARG, "unknown filter");
  if (out->w != p->g.W || out->h != p->g.H || out->c != p->g.C || out->nimg != p->g.nimg)
    return fail(ctx, PXZ_E_ARG, "output image geometry mismatch");
  if (rgb_on_rgba(ctx, p->g.W, p->g.C, p->g.bw, p->g.bh) && p->max_small_dim <= 64) {
    pxz_payload* p4 = nullptr;
    pxz_status st = payload_rechannel(ctx, p, 4, &p4);
    if (st != PXZ_OK) return st;
    pxz_image* wide = nullptr;
    st = pxz_image_alloc_batch(ctx, out->w, out->h, 4, out->nimg, &wide);
    if (st == PXZ_OK) st = pxz_expand_to_image(ctx, p4, filter_up, wide);
    if (st == PXZ_OK) {
      cudaError_t e = launch_rgb_widen(wide->d, wide->pitch, out->d, out->pitch, out->w, out->h * out->nimg, 0, ctx->stream, ctx->sm_count,
                                       &ctx->launches);
      if (e != cudaSuccess) {
        cudaGetLastError();
        st = fail(ctx, PXZ_E_CUDA, std::string("rgb narrow: ") + cudaGetErrorString(e));
      }
    }
    pxz_image_free(wide);
    payload_release(p4);
    return st;
  }
  TabSet ts;
  pxz_status st = get_tabset(ctx, *p->spec, p->strategy ? -1 : (int)filter_up, 1, &ts);
  if (st != PXZ_OK) return st;
  return run_resample(ctx, 1, out->d, out->pitch, p, ts, p->max_small_px, p->max_small_dim, p->max_tmp_up);
}

pxz_status pxz_expand(pxz_ctx* ctx, const pxz_payload* p, pxz_filter filter_up, uint8_t* host_out, size_t host_pitch) {
  if (!ctx || !p || !host_out) return PXZ_E_ARG;
  pxz_image* im = nullptr;
  pxz_status st = pxz_image_alloc_batch(ctx, p->g.W, p->g.H, p->g.C, p->g.nimg, &im);
  if (st != PXZ_OK) return st;
  st = pxz_expand_to_image(ctx, p, filter_up, im);
  if (st == PXZ_OK) st = pxz_image_download(ctx, im, host_out, host_pitch);
  pxz_image_free(im);
  return st;
}

// ---- quadtree processing (process/tree.rs:23-109) ---------------------------------------------------------
pxz_status pxz_tree_process(pxz_ctx* ctx, const pxz_image* img, float threshold, uint32_t bw, uint32_t bh, uint32_t min_bw,
                            uint32_t min_bh, pxz_filter filter_down, pxz_filter filter_up, pxz_image* out) {
  if (!ctx || !img || !out) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  if ((int)filter_down < 0 || (int)filter_down > 4 || (int)filter_up < 0 || (int)filter_up > 4) return fail(ctx, PXZ_E_ARG, "unknown filter");
  if (out->w != img->w || out->h != img->h || out->c != img->c) return fail(ctx, PXZ_E_ARG, "output image geometry mismatch");
  if (img->nimg != 1 || out->nimg != 1) return fail(ctx, PXZ_E_UNSUPPORTED, "tree processing takes one image at a time");
  if (bw == 0 || bh == 0) return fail(ctx, PXZ_E_ARG, "block size must be >= 1");
  // unchanged pixels (blocks that reach the minimum size, tree.rs:35-37) come straight from the source
  PXZ_CUDA(ctx, cudaMemcpy2DAsync(out->d, out->pitch, img->d, img->pitch, (size_t)img->w * img->c, img->h, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
  const uint32_t mbw = std::max(min_bw, 4u), mbh = std::max(min_bh, 4u);  // :33-34
  int levels = 0;
  while ((bw >> levels) > mbw && (bh >> levels) > mbh) ++levels;
  if (levels == 0) return PXZ_OK;  // :35-37: image.clone()
  // the reference splits every block relative to its own origin; the levels form global grids only when the halved
  // sizes stay exact, which is what the accelerated path covers
  if ((bw & ((1u << (levels - 1)) - 1)) || (bh & ((1u << (levels - 1)) - 1)))
    return fail(ctx, PXZ_E_UNSUPPORTED, "tree processing needs block sizes divisible by 2^(levels-1)");
  const bool positive = threshold >= 0.0f;  // :38
  const float thr = fabsf(threshold);       // :39

  uint8_t* d_leaf = nullptr;
  uint8_t* d_rec[2] = {nullptr, nullptr};
  Geom gl;
  pxz_status st = make_geom(ctx, img->w, img->h, img->c, bw >> (levels - 1), bh >> (levels - 1), &gl);
  if (st != PXZ_OK) return st;
  const size_t max_blocks = (size_t)gl.cols * gl.rows;
  if ((st = dev_alloc(ctx, (void**)&d_leaf, max_blocks)) != PXZ_OK) return st;
  if ((st = dev_alloc(ctx, (void**)&d_rec[0], max_blocks)) != PXZ_OK || (st = dev_alloc(ctx, (void**)&d_rec[1], max_blocks)) != PXZ_OK) {
    dev_free(ctx, d_leaf); dev_free(ctx, d_rec[0]); dev_free(ctx, d_rec[1]);
    return st;
  }
  uint32_t parent_cols = 0;
  for (int lv = 0; lv < levels && st == PXZ_OK; ++lv) {
    Geom g;
    st = make_geom(ctx, img->w, img->h, img->c, bw >> lv, bh >> lv, &g);
    if (st != PXZ_OK) break;
    ValueMap vm;
    vm.factor = 1.0f; vm.mode = 1; vm.normalise = 0;  // after = identity (tree.rs:97)
    vm.extra_thr = thr;
    st = run_analysis(ctx, img, g, PXZ_METRIC_OKLAB_MAD, false, &vm);
    if (st != PXZ_OK) break;
    cudaError_t e = launch_tree_mask(ctx->d_vx, g, lv == 0 ? nullptr : d_rec[(lv - 1) & 1], parent_cols, thr,
                                     (lv == 0 ? positive : true) ? 1 : 0,  // the recursion receives |threshold| (:70)
                                     d_leaf, d_rec[lv & 1], ctx->stream, &ctx->launches);
    if (e != cudaSuccess) { st = fail(ctx, PXZ_E_CUDA, std::string("tree mask: ") + cudaGetErrorString(e)); break; }
    parent_cols = g.cols;
    pxz_payload* p = nullptr;
    st = payload_new(ctx, g, (uint64_t)g.W * g.H * g.C, &p);
    if (st != PXZ_OK) break;
    p->spec = spec_for_geom(g);
    auto pad8 = [](uint32_t v) { return (v + 7u) & ~7u; };
    p->max_small_px = g.bw * g.bh;
    p->max_small_dim = std::max(g.bw, g.bh);
    p->max_tmp_down = ((g.bh + 1) / 2) * pad8(g.bw);
    p->max_tmp_up = g.bh * pad8((g.bw + 1) / 2);
    vm.extra_thr = NAN;
    e = launch_plan(ctx->d_vx, nullptr, g, vm, ctx->d_minmax, ctx->thr, d_leaf, p->d_descs, p->d_tabidx, p->d_total, ctx->d_scan, p->d_order(),
                    (uint32_t)p->nblocks_cap, ctx->stream, &ctx->launches);
    if (e != cudaSuccess) st = fail(ctx, PXZ_E_CUDA, std::string("plan: ") + cudaGetErrorString(e));
    TabSet ts;
    if (st == PXZ_OK) st = get_tabset(ctx, *p->spec, (int)filter_down, 0, &ts);
    if (st == PXZ_OK) st = run_resample(ctx, 0, img->d, img->pitch, p, ts, g.bw * g.bh, std::max(g.bw, g.bh), p->max_tmp_down);
    if (st == PXZ_OK) st = get_tabset(ctx, *p->spec, (int)filter_up, 1, &ts);
    if (st == PXZ_OK) st = run_resample(ctx, 1, out->d, out->pitch, p, ts, p->max_small_px, p->max_small_dim, p->max_tmp_up);
    payload_release(p);
  }
  dev_free(ctx, d_leaf);
  dev_free(ctx, d_rec[0]);
  dev_free(ctx, d_rec[1]);
  return st;
}

// ---- multi-GPU ----------------------------------------------------------------------------------------
pxz_status pxz_comm_unique_id(uint8_t id[PXZ_COMM_ID_BYTES]) {
  if (!id) return PXZ_E_ARG;
  std::string err;
  if (nccl_get_unique_id(id, &err) != 0) {
    fprintf(stderr, "pixlzr_b200: %s\n", err.c_str());
    return PXZ_E_NCCL;
  }
  return PXZ_OK;
}

pxz_status pxz_comm_init(pxz_ctx* ctx, int nranks, int rank, const uint8_t id[PXZ_COMM_ID_BYTES]) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return PXZ_E_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->comm) {
    nccl_comm_destroy(ctx->comm);
    ctx->comm = nullptr;
  }
  std::string err;
  if (nccl_comm_init(&ctx->comm, nranks, rank, id, &err) != 0) return fail(ctx, PXZ_E_NCCL, err);
  return PXZ_OK;
}

void pxz_comm_destroy(pxz_ctx* ctx) {
  if (ctx && ctx->comm) {
    nccl_comm_destroy(ctx->comm);
    ctx->comm = nullptr;
  }
}

}  // extern "C"
