/*
 * pxz_oracle.h — CPU ORACLE for the pixlzr hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain C++ restatement of the reference's algorithm (crate `pixlzr` v0.3.1,
 * cited as `path:line` relative to the reference tree) used ONLY as the checker in
 * `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs
 * of `bench.py`.  Nothing in the product path (pixlzr-rust_b200/) may include, link or
 * call it.
 *
 * Parity pinning (see tests/test_oracle_golden.py, tests/golden/):
 *   - Oklab-MAD metric, value->dims, stored value: PINNED bit-exact by Big-Ruscher.png -> .pix
 *     (2040/2040 f32 values, 2040/2040 dims), 3-channel path.  4-channel alpha term: parity unpinned.
 *   - image-crate (`image_rs`) resize: PINNED for Lanczos3 downscale (all blocks larger than
 *     1x1 exact, 1x1 blocks within +-1 LSB on exact .5 ties) and Nearest upscale (.pix.png
 *     2040/2040 exact).  Triangle / CatmullRom / Gaussian and non-nearest upscales: parity
 *     unpinned (restated from the `image` 0.25.5 algorithm).
 *   - container v0.0.2 + QOI (qoi 0.4.1 incl. its run-of-1 quirk): PINNED byte-exact by
 *     benches/base.pixlzr and Big-Ruscher.pix.
 *   - `fast_image_resize` branch (the reference's default cargo feature, block.rs:292-333): restated from the
 *     crate's published algorithm (pxo_resize_fir, pxo_set_resize_semantics); PARITY UNPINNED — no fixture was made
 *     with it and its source is not under /root/reference; only block.rs:401-435 constrains it.
 */
#ifndef PXZ_ORACLE_H
#define PXZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FilterType repr(u8), src/data_types/mod.rs:10-30 */
enum { PXO_NEAREST = 0, PXO_TRIANGLE = 1, PXO_CATMULLROM = 2, PXO_GAUSSIAN = 3, PXO_LANCZOS3 = 4 };
/* metric: 0 = Oklab MAD (operations.rs:26-126), 1 = directional Sobel (operations.rs:192-259) */
enum { PXO_METRIC_OKLAB_MAD = 0, PXO_METRIC_SOBEL_DIR = 1 };

typedef struct {
  uint64_t offset; /* byte offset of the block's pixels inside the packed payload */
  float value;     /* stored block value = hypot(v0, v1) after parse_value        */
  uint16_t w, h;   /* reduced block size                                          */
} pxo_block_desc;

/* ---- primitives ------------------------------------------------------------------ */
void pxo_grid(uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, uint32_t* cols, uint32_t* rows);
void pxo_srgb_lut(float out[256]);
float pxo_cbrtf(float x);
void pxo_oklab(uint8_t r, uint8_t g, uint8_t b, float out_lab[3]);
/* raw metric of one block (before `after`); px points at the block's top-left pixel */
float pxo_block_mad(const uint8_t* px, size_t pitch, uint32_t w, uint32_t h, int channels);
int pxo_block_sobel(const uint8_t* px, size_t pitch, uint32_t w, uint32_t h, int channels,
                    float* hz, float* vr);
float pxo_parse_value(float v);
/* level exponent e (level = 2^e, e <= 0; INT32_MIN for level 0) */
int32_t pxo_level_exp(float parsed_v);
void pxo_reduce_dims(float v0, float v1, uint32_t w, uint32_t h, uint32_t* ow, uint32_t* oh,
                     float* stored);
/* PixlzrBlock::resize, image-crate branch (block.rs:273-290). src/dst tightly packed. */
int pxo_resize(const uint8_t* src, uint32_t w, uint32_t h, int channels, uint8_t* dst,
               uint32_t nw, uint32_t nh, int filter);
/* PixlzrBlock::resize, fast_image_resize branch (block.rs:292-333, data_types/mod.rs:65-107): parity unpinned. */
int pxo_resize_fir(const uint8_t* src, uint32_t w, uint32_t h, int channels, uint8_t* dst,
                   uint32_t nw, uint32_t nh, int filter);
/* which branch the drivers (shrink / expand / tree) use: 0 = image crate (default, pinned), 1 = fast_image_resize.
   Process-wide, not thread-safe: set it between calls. */
void pxo_set_resize_semantics(int fir);
int pxo_get_resize_semantics(void);
/* normalised f32 weight table of one axis (image 0.25.5 sample loops): for each output o,
   left[o], count[o] and weights (row stride = max_taps). returns max taps or <0 */
int pxo_axis_weights(uint32_t n, uint32_t nn, int filter, uint32_t* left, uint32_t* count,
                     float* weights, uint32_t max_taps);

/* ---- drivers ------------------------------------------------------------------------ */
/* Per-block raw metric values for the whole image (row-major grid). vy may be NULL for MAD. */
int pxo_analyze(const uint8_t* img, uint32_t w, uint32_t h, int channels, size_t pitch,
                uint32_t bw, uint32_t bh, int metric, float* vx, float* vy, int nthreads);
/* Pixlzr::shrink_by / shrink_directionally (pixlzr.rs:155-205) or process()'s after = id
   (process/mod.rs:107-121, use_factor = 0).  normalise_global: extension, see DESIGN.md.
   descs: cols*rows; payload: capacity >= w*h*channels. returns payload bytes or <0. */
int64_t pxo_shrink(const uint8_t* img, uint32_t w, uint32_t h, int channels, size_t pitch,
                   uint32_t bw, uint32_t bh, int metric, float factor, int use_factor,
                   int filter_down, int normalise_global, pxo_block_desc* descs,
                   uint8_t* payload, int nthreads);
/* Pixlzr::expand + to_image (pixlzr.rs:77-122, pixlzr_image.rs:24-74) */
int pxo_expand(const pxo_block_desc* descs, const uint8_t* payload, uint32_t w, uint32_t h,
               uint32_t bw, uint32_t bh, int channels, int filter_up, uint8_t* out,
               size_t out_pitch, int nthreads);

/* EXTENSION — per-block filter pairs (include/pixlzr_b200.h "strategy").  The reference only logged the experiment
   (strategies.txt:1-118, strategies_by_level.txt:1-12); there is no reference run to pin these against, so they are
   "parity unpinned" beyond the per-block resize they are built from (which is pinned).  *_by_bucket: 65 filter ids,
   indexed by pxo_strategy_bucket(stored block value) = floor(64 * value / sqrt(2)) clamped to [0, 64]. */
uint32_t pxo_strategy_bucket(float value);
int64_t pxo_shrink_strategy(const uint8_t* img, uint32_t w, uint32_t h, int channels, size_t pitch, uint32_t bw,
                            uint32_t bh, int metric, float factor, int use_factor, const uint8_t* down_by_bucket,
                            int normalise_global, pxo_block_desc* descs, uint8_t* payload, int nthreads);
int pxo_expand_strategy(const pxo_block_desc* descs, const uint8_t* payload, uint32_t w, uint32_t h, uint32_t bw,
                        uint32_t bh, int channels, const uint8_t* up_by_bucket, uint8_t* out, size_t out_pitch,
                        int nthreads);

/* tree::process_custom (process/tree.rs:23-83) with before = |x-avg|, after = identity: quadtree of blocks, a block
   whose value is below the threshold is reduced + re-expanded, the others are split again with halved block size
   until the size reaches max(min, 4).  Output has the input's channel count (the reference pastes into an RGBA8
   canvas: alpha 255 is added by the caller for RGB inputs).  parity unpinned: no reference fixture exercises it. */
int pxo_tree_process(const uint8_t* img, uint32_t w, uint32_t h, int channels, size_t pitch, float threshold,
                     uint32_t bw, uint32_t bh, uint32_t min_bw, uint32_t min_bh, int filter_down, int filter_up,
                     uint8_t* out, size_t out_pitch);

/* ---- container (encoding/mod.rs:40-242) + QOI (qoi 0.4.1) --------------------------- */
/* returns bytes written (or needed if out == NULL / cap too small -> negative of needed) */
int64_t pxo_qoi_encode(const uint8_t* px, uint32_t w, uint32_t h, int channels, uint8_t* out,
                       size_t cap);
int pxo_qoi_decode(const uint8_t* data, size_t len, uint32_t* w, uint32_t* h, int* channels,
                   uint8_t* out, size_t cap);
int64_t pxo_container_encode(uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, int filter,
                             int channels, const pxo_block_desc* descs, const uint8_t* payload,
                             const uint8_t* value_present /* may be NULL = all present */,
                             uint8_t* out, size_t cap);
/* pass 1 (descs == NULL): fills header fields + payload size; pass 2: fills descs + payload */
int pxo_container_decode(const uint8_t* data, size_t len, uint32_t* w, uint32_t* h,
                         uint32_t* bw, uint32_t* bh, int* filter, int* channels,
                         uint64_t* payload_bytes, pxo_block_desc* descs, uint8_t* payload);

#ifdef __cplusplus
}
#endif
#endif
