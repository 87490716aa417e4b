/*
 * pxz_oracle.cpp — CPU ORACLE for the pixlzr hot path.  TEST INFRASTRUCTURE ONLY
 * (see pxz_oracle.h for the usage rule and the parity-pinning status of every part).
 *
 * Build: make -C oracle      (g++ -O2 -ffp-contract=off -fopenmp; NO -ffast-math)
 *
 * Every function cites the reference file:line (relative to the reference tree) or, for
 * arithmetic living in un-vendored crates, the crate@version whose published algorithm it
 * restates:  palette 0.7.6 + fast-srgb8 1.0.0 (sRGB->linear->Oklab), image 0.25.5
 * (imageops::resize), qoi 0.4.1 (block codec), glibc < 2.41 cbrtf.
 */
#include "pxz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ===================================================================================
 * sRGB u8 -> linear f32 LUT.  palette 0.7.6 `Srgb::into_linear::<f32>(u8)` ->
 * fast_srgb8::srgb8_to_f32 (table).  Restated as the stepwise-f32 formula that reproduces
 * the reference fixture bit-for-bit: x = u * (1/255)f; x <= 0.04045 -> x/12.92,
 * else powf((x+0.055)/1.055, 2.4) with a correctly rounded powf (double pow, rounded once).
 * =================================================================================== */
struct SrgbLut {
  float v[256];
  SrgbLut() {
    const float inv255 = (float)(1.0 / 255.0);
    for (int u = 0; u < 256; ++u) {
      float x = (float)u * inv255;
      if (x <= 0.04045f) {
        v[u] = x / 12.92f;
      } else {
        float t = (x + 0.055f) / 1.055f;
        v[u] = (float)pow((double)t, (double)2.4f);
      }
    }
  }
};
const SrgbLut g_lut;

/* ===================================================================================
 * cbrtf — the pre-2.41 glibc algorithm (sysdeps/ieee754/flt-32/s_cbrtf.c), which is what
 * Rust's f32::cbrt resolved to when the reference fixture was produced (2040/2040 values
 * only match with this variant).  Own copy so that a newer host libm cannot
 * change the oracle.
 * =================================================================================== */
inline float cbrtf_glibc_old(float x) {
  static const double factor[5] = {
      0.62996052494743658238361, /* 1 / 2^(2/3) */
      0.79370052598409973737585, /* 1 / 2^(1/3) */
      1.0,
      1.2599210498948731647672, /* 2^(1/3) */
      1.5874010519681994747517, /* 2^(2/3) */
  };
  int xe;
  float xm = frexpf(fabsf(x), &xe);
  /* inf, nan, 0 */
  if (xe == 0 && (x == 0.0f || isinf(x) || isnan(x))) return x + x;
  float u = (float)(0.492659620528969547 + (0.697570460207922770 - 0.191502161678719066 * (double)xm) * (double)xm);
  float t2 = u * u * u;
  float ym = (float)((double)u * ((double)t2 + 2.0 * (double)xm) / (2.0 * (double)t2 + (double)xm) * factor[2 + xe % 3]);
  return ldexpf(x > 0.0f ? ym : -ym, xe / 3);
}

/* ===================================================================================
 * linear sRGB -> Oklab, palette 0.7.6 `Oklab::from_color_unclamped(LinSrgb)`:
 * two 3x3 products evaluated left to right in f32 WITHOUT fma contraction
 * (operations.rs:56-59,94-97 call sites).
 * =================================================================================== */
struct Lab { float l, a, b; };

inline Lab oklab_from_srgb8(uint8_t r8, uint8_t g8, uint8_t b8) {
  const float r = g_lut.v[r8], g = g_lut.v[g8], b = g_lut.v[b8];
  const float l = 0.4122214708f * r + 0.5363325363f * g + 0.0514459929f * b;
  const float m = 0.2119034982f * r + 0.6806995451f * g + 0.1073969566f * b;
  const float s = 0.0883024619f * r + 0.2817188376f * g + 0.6299787005f * b;
  const float l_ = cbrtf_glibc_old(l), m_ = cbrtf_glibc_old(m), s_ = cbrtf_glibc_old(s);
  Lab o;
  o.l = 0.2104542553f * l_ + 0.7936177850f * m_ - 0.0040720468f * s_;
  o.a = 1.9779984951f * l_ - 2.4285922050f * m_ + 0.4505937099f * s_;
  o.b = 0.0259040371f * l_ + 0.7827717662f * m_ - 0.8086757660f * s_;
  return o;
}

/* ===================================================================================
 * get_block_variance with before = |x - avg| (operations.rs:26-126, pixlzr.rs:160-161).
 * Sequential f32 running sums in block scan order, channel order (a, b, l[, alpha]).
 * Returns the value BEFORE `after` (i.e. tot / count).
 * =================================================================================== */
float block_mad(const uint8_t* px, size_t pitch, uint32_t w, uint32_t h, int C) {
  const float count = (float)(uint32_t)(w * h); /* operations.rs:51 */
  const float inv255 = (float)(1.0 / 255.0);    /* palette u8->f32 stimulus (alpha) */
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* row = px + (size_t)y * pitch;
    for (uint32_t x = 0; x < w; ++x) {
      const uint8_t* p = row + (size_t)x * C;
      Lab c = oklab_from_srgb8(p[0], p[1], p[2]);
      sum[0] += c.a; /* operations.rs:60-63 / 98-100 */
      sum[1] += c.b;
      sum[2] += c.l;
      if (C == 4) sum[3] += (float)p[3] * inv255;
    }
  }
  for (int i = 0; i < 4; ++i) sum[i] /= count; /* :65-68 / :102-104 */
  float delta[4] = {0.f, 0.f, 0.f, 0.f};
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* row = px + (size_t)y * pitch;
    for (uint32_t x = 0; x < w; ++x) {
      const uint8_t* p = row + (size_t)x * C;
      Lab c = oklab_from_srgb8(p[0], p[1], p[2]);
      delta[0] += fabsf(c.a - sum[0]); /* :80-83 / :116-118 */
      delta[1] += fabsf(c.b - sum[1]);
      delta[2] += fabsf(c.l - sum[2]);
      if (C == 4) delta[3] += fabsf((float)p[3] * inv255 - sum[3]);
    }
  }
  if (C == 4) return (delta[0] + delta[1] + delta[2] + delta[3]) / count; /* :89 */
  return (delta[0] + delta[1] + delta[2]) / count;                         /* :124 */
}

/* ===================================================================================
 * get_block_variance_directionally (operations.rs:192-259).  Pure integer, then one f64
 * divide.  The reference underflows `height - 2` / `width - 2` for blocks thinner than 2
 * (panic); for exactly 2 it divides 0/0 (x86: negative NaN).  The oracle returns -1 for
 * w < 2 || h < 2 and the x86 NaN for == 2.
 * =================================================================================== */
int block_sobel(const uint8_t* px, size_t pitch, uint32_t w, uint32_t h, int C, float* hz, float* vr) {
  if (w < 2 || h < 2) return -1;
  uint64_t sum_hz = 0, sum_vr = 0;
  for (uint32_t y = 0; y + 2 < h; ++y) {
    const uint8_t* r0 = px + (size_t)y * pitch;
    const uint8_t* r1 = r0 + pitch;
    const uint8_t* r2 = r1 + pitch;
    for (uint32_t x = 0; x + 2 < w; ++x) {
      for (int c = 0; c < 3; ++c) { /* alpha ignored, :217-218 */
        const int v00 = r0[(x + 0) * C + c], v01 = r0[(x + 1) * C + c], v02 = r0[(x + 2) * C + c];
        const int v10 = r1[(x + 0) * C + c], v12 = r1[(x + 2) * C + c];
        const int v20 = r2[(x + 0) * C + c], v21 = r2[(x + 1) * C + c], v22 = r2[(x + 2) * C + c];
        const int phz = -v00 - 2 * v01 - v02 + v20 + 2 * v21 + v22; /* :240-241 */
        const int pvr = -v00 - 2 * v10 - v20 + v02 + 2 * v12 + v22; /* :244-245 */
        sum_hz += (uint64_t)abs(phz);
        sum_vr += (uint64_t)abs(pvr);
      }
    }
  }
  const double factor = (double)((uint64_t)(w - 2) * (uint64_t)(h - 2) * 4096ull); /* :158,:253-254 */
  if (factor == 0.0) {
    /* 0/0 on x86-64 SSE2 yields the default (negative) quiet NaN */
    uint32_t bits = 0xFFC00000u;
    float nanv;
    memcpy(&nanv, &bits, 4);
    *hz = nanv;
    *vr = nanv;
    return 0;
  }
  *hz = (float)((double)sum_hz / factor);
  *vr = (float)((double)sum_vr / factor);
  return 0;
}

/* parse_value, operations.rs:128-138 */
float parse_value(float value) {
  if (!signbit(value)) return value;
  float v = 1.0f + value;
  /* f32::max(NaN, 0) = 0 */
  v = (v != v) ? 0.0f : (v > 0.0f ? v : 0.0f);
  /* `(1+value).max(0)` of -0.0: Rust's max may return either zero; x86 maxss(a=-0,b=0) -> 0.0;
     the reference maps a negative zero to 1.0 (:133-137).  +0 here, so value 0. */
  if (!signbit(v)) return v;
  return 1.0f;
}

/* level = exp2(min(0, round(log2 v))), operations.rs:147-148.  Returns the exponent. */
int32_t level_exp(float v) {
  float lg = log2f(v);              /* host libm, as Rust's f32::log2 on Linux */
  float r = roundf(lg);             /* half away from zero */
  float m = (r != r) ? 0.0f : (r < 0.0f ? r : 0.0f); /* f32::min(NaN, 0) = 0 */
  if (isinf(m)) return INT32_MIN;   /* exp2(-inf) = 0 */
  if (m < -200.0f) return -200;     /* exp2 underflows to 0/denormal: dims are 1 anyway */
  return (int32_t)m;
}

uint32_t scaled_dim(uint32_t n, int32_t e) {
  /* (n as f64 * level as f64).max(1).ceil() as u32, operations.rs:150-151 */
  if (e == INT32_MIN) return 1;
  double level = (e < -149) ? 0.0 : (double)ldexpf(1.0f, e); /* exp2f result as f32 (denormal ok) */
  double d = (double)n * level;
  if (!(d > 1.0)) d = 1.0;
  return (uint32_t)ceil(d);
}

/* f32::hypot -> libm hypotf; restated as the double formula pinned by the fixture */
float stored_value(float v0, float v1) {
  if (isinf(v0) || isinf(v1)) return INFINITY;
  return (float)sqrt((double)v0 * (double)v0 + (double)v1 * (double)v1);
}

/* ===================================================================================
 * image 0.25.5 imageops::sample kernels + resize (called through
 * DynamicImage::resize_exact, block.rs:288).  All f32, no fma.
 * =================================================================================== */
const float PI_F = 3.14159274101257324219f;

float k_sinc(float t) {
  float a = t * PI_F;
  if (t == 0.0f) return 1.0f;
  return sinf(a) / a;
}
float k_lanczos3(float x) { return fabsf(x) < 3.0f ? k_sinc(x) * k_sinc(x / 3.0f) : 0.0f; }
float k_triangle(float x) { return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f; }
float k_catmullrom(float x) { /* bc_cubic_spline(x, b = 0, c = 0.5) */
  float a = fabsf(x), k;
  if (a < 1.0f) {
    k = 9.0f * (a * a * a) + -15.0f * (a * a) + 6.0f;
  } else if (a < 2.0f) {
    k = -3.0f * (a * a * a) + 15.0f * (a * a) + -24.0f * a + 12.0f;
  } else {
    k = 0.0f;
  }
  return k / 6.0f;
}
float k_gaussian(float x) { /* gaussian(x, r = 0.5) */
  const float r = 0.5f;
  float norm = 1.0f / (sqrtf(2.0f * PI_F) * r);
  return norm * expf(-(x * x) / (2.0f * (r * r)));
}
float k_box(float) { return 1.0f; }

struct Filter { float (*kernel)(float); float support; };
bool get_filter(int f, Filter* out) {
  switch (f) {
    case PXO_NEAREST: *out = {k_box, 0.0f}; return true;
    case PXO_TRIANGLE: *out = {k_triangle, 1.0f}; return true;
    case PXO_CATMULLROM: *out = {k_catmullrom, 2.0f}; return true;
    case PXO_GAUSSIAN: *out = {k_gaussian, 3.0f}; return true;
    case PXO_LANCZOS3: *out = {k_lanczos3, 3.0f}; return true;
  }
  return false;
}

struct AxisTaps { std::vector<uint32_t> left, count; std::vector<float> w; uint32_t stride; };

/* the per-output-index part of horizontal_sample / vertical_sample */
void axis_taps(uint32_t n, uint32_t nn, const Filter& flt, AxisTaps* t) {
  const float ratio = (float)n / (float)nn;
  const float sratio = ratio < 1.0f ? 1.0f : ratio;
  const float src_support = flt.support * sratio;
  t->left.resize(nn);
  t->count.resize(nn);
  std::vector<std::vector<float>> ws(nn);
  uint32_t maxc = 0;
  for (uint32_t o = 0; o < nn; ++o) {
    float in = ((float)o + 0.5f) * ratio;
    int64_t left = (int64_t)floorf(in - src_support);
    left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)n - 1);
    int64_t right = (int64_t)ceilf(in + src_support);
    right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)n);
    in = in - 0.5f;
    float sum = 0.0f;
    for (int64_t i = left; i < right; ++i) {
      float w = flt.kernel(((float)i - in) / sratio);
      ws[o].push_back(w);
      sum += w;
    }
    for (float& w : ws[o]) w /= sum;
    t->left[o] = (uint32_t)left;
    t->count[o] = (uint32_t)(right - left);
    maxc = std::max(maxc, t->count[o]);
  }
  t->stride = maxc;
  t->w.assign((size_t)nn * maxc, 0.0f);
  for (uint32_t o = 0; o < nn; ++o)
    for (uint32_t i = 0; i < t->count[o]; ++i) t->w[(size_t)o * maxc + i] = ws[o][i];
}

inline uint8_t to_u8(float t) {
  /* NumCast::from(FloatNearest(clamp(t, 0, 255))): round half away from zero */
  if (t < 0.0f) t = 0.0f;
  if (t > 255.0f) t = 255.0f; /* NaN passes through clamp; cast saturates to 0 */
  float r = roundf(t);
  return (r != r) ? 0 : (uint8_t)r;
}

/* imageops::resize = vertical_sample (-> f32 image) then horizontal_sample (-> u8) */
int resize_image_rs(const uint8_t* src, uint32_t w, uint32_t h, int C, uint8_t* dst, uint32_t nw,
                    uint32_t nh, int filter) {
  if (w == 0 || h == 0 || nw == 0 || nh == 0) return -1;
  if (w == nw && h == nh) { /* block.rs:279-281 */
    memcpy(dst, src, (size_t)w * h * C);
    return 0;
  }
  Filter flt;
  if (!get_filter(filter, &flt)) return -1;
  AxisTaps tv, th;
  axis_taps(h, nh, flt, &tv);
  axis_taps(w, nw, flt, &th);
  std::vector<float> tmp((size_t)nh * w * C);
  for (uint32_t oy = 0; oy < nh; ++oy) {
    const float* ws = &tv.w[(size_t)oy * tv.stride];
    const uint32_t left = tv.left[oy], cnt = tv.count[oy];
    for (uint32_t x = 0; x < w; ++x) {
      for (int c = 0; c < C; ++c) {
        float t = 0.0f;
        for (uint32_t i = 0; i < cnt; ++i) t += (float)src[((size_t)(left + i) * w + x) * C + c] * ws[i];
        tmp[((size_t)oy * w + x) * C + c] = t;
      }
    }
  }
  for (uint32_t ox = 0; ox < nw; ++ox) {
    const float* ws = &th.w[(size_t)ox * th.stride];
    const uint32_t left = th.left[ox], cnt = th.count[ox];
    for (uint32_t y = 0; y < nh; ++y) {
      for (int c = 0; c < C; ++c) {
        float t = 0.0f;
        for (uint32_t i = 0; i < cnt; ++i) t += tmp[((size_t)y * w + (left + i)) * C + c] * ws[i];
        dst[((size_t)y * nw + ox) * C + c] = to_u8(t);
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * PixlzrBlock::resize, `fir` branch (block.rs:292-333) — the reference's DEFAULT cargo feature: fast_image_resize
 * 4.2.1 driven by FilterType::to_fir_resizing_algorithm (data_types/mod.rs:65-107).
 *
 * PARITY UNPINNED.  The crate's source is not under /root/reference (Cargo.lock pins fast_image_resize 4.2.1) and no
 * fixture of the reference was produced with this branch; what follows restates the crate's published algorithm (the
 * Pillow-SIMD lineage it documents) from the call site's arguments:
 *   - algorithm: Nearest -> nearest; growing in either axis -> SuperSampling(f, 2), which for an enlargement is a plain
 *     convolution with f (the pre-shrink only happens when the reduction exceeds 2 x 1.2); otherwise Convolution(f),
 *     with Triangle -> Bilinear when growing and -> Hamming when shrinking (mod.rs:75-101);
 *   - coefficients per axis in f64: scale = in / out, filter_scale = max(scale, 1), radius = support * filter_scale,
 *     centre = (o + 0.5) * scale, taps x in [floor(centre - radius) clamped to 0, ceil(centre + radius) clamped to in),
 *     w = f((x + 0.5 - centre) / filter_scale), normalised to sum 1;
 *   - 16-bit fixed point with adaptive precision: the largest p < 22 with round(max_w * 2^(p+1)) < 2^15, coefficient =
 *     round-half-away(w * 2^p); a sample is clip8((2^(p-1) + sum(pixel * coefficient)) >> p);
 *   - HORIZONTAL pass first into a u8 image of (new width x old height), then the vertical pass; an axis whose size
 *     does not change is not convolved;
 *   - U8x4 with a non-nearest algorithm: alpha is pre-multiplied before and divided out after the resize
 *     (ResizeOptions::new() has use_alpha = true): c' = mul_div_255(c, a), back = min(255, (c' * 255 + a / 2) / a).
 * The only constraint the reference's own tests put on it is block.rs:401-435 (constant images stay constant).
 * ------------------------------------------------------------------------------------------------------------------ */
enum FirFilter { FIR_BOX = 0, FIR_BILINEAR = 1, FIR_CATMULLROM = 2, FIR_GAUSSIAN = 3, FIR_LANCZOS3 = 4, FIR_HAMMING = 5 };

double fir_sinc(double x) { return x == 0.0 ? 1.0 : sin(x * M_PI) / (x * M_PI); }
double fir_kernel(int f, double x) {
  switch (f) {
    case FIR_BILINEAR: { x = fabs(x); return x < 1.0 ? 1.0 - x : 0.0; }
    case FIR_HAMMING: {
      x = fabs(x);
      if (x == 0.0) return 1.0;
      if (x >= 1.0) return 0.0;
      x *= M_PI;
      return (0.54 + 0.46 * cos(x)) * sin(x) / x;
    }
    case FIR_CATMULLROM: {
      const double a = -0.5;
      x = fabs(x);
      if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
      if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
      return 0.0;
    }
    case FIR_GAUSSIAN: {
      if (fabs(x) >= 3.0) return 0.0;
      const double r = 0.5;
      return exp(-(x * x) / (2.0 * r * r)) / (sqrt(2.0 * M_PI) * r);
    }
    case FIR_LANCZOS3: return (x >= -3.0 && x < 3.0) ? fir_sinc(x) * fir_sinc(x / 3.0) : 0.0;
  }
  return 0.0;
}
double fir_support(int f) { return f == FIR_CATMULLROM ? 2.0 : (f == FIR_GAUSSIAN || f == FIR_LANCZOS3) ? 3.0 : 1.0; }

struct FirAxis { std::vector<uint32_t> left, count; std::vector<int32_t> k; uint32_t stride; int precision; };

void fir_axis(uint32_t n_in, uint32_t n_out, int f, FirAxis* t) {
  const double scale = (double)n_in / (double)n_out, fscale = scale > 1.0 ? scale : 1.0, radius = fir_support(f) * fscale;
  t->left.assign(n_out, 0); t->count.assign(n_out, 0);
  std::vector<std::vector<double>> w(n_out);
  double maxw = 0.0;
  uint32_t stride = 1;
  for (uint32_t o = 0; o < n_out; ++o) {
    const double centre = ((double)o + 0.5) * scale;
    double lo = floor(centre - radius), hi = ceil(centre + radius);
    if (lo < 0.0) lo = 0.0;
    if (hi > (double)n_in) hi = (double)n_in;
    const uint32_t x0 = (uint32_t)lo, x1 = (uint32_t)hi;
    double ww = 0.0;
    for (uint32_t x = x0; x < x1; ++x) { const double v = fir_kernel(f, ((double)x + 0.5 - centre) / fscale); w[o].push_back(v); ww += v; }
    if (ww != 0.0) for (double& v : w[o]) v /= ww;
    for (double v : w[o]) maxw = std::max(maxw, v);
    t->left[o] = x0; t->count[o] = x1 - x0;
    stride = std::max(stride, x1 - x0);
  }
  int p = 0;
  for (p = 0; p < 22; ++p) { if ((int)lround(maxw * (double)(1 << (p + 1))) >= (1 << 15)) break; }
  t->precision = p; t->stride = stride;
  t->k.assign((size_t)n_out * stride, 0);
  for (uint32_t o = 0; o < n_out; ++o)
    for (size_t i = 0; i < w[o].size(); ++i) t->k[(size_t)o * stride + i] = (int32_t)lround(w[o][i] * (double)(1 << p));
}

inline uint8_t fir_clip8(int32_t v, int p) { v >>= p; return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
inline uint8_t mul_div_255(uint32_t a, uint32_t b) { const uint32_t t = a * b + 128; return (uint8_t)(((t >> 8) + t) >> 8); }

int resize_fir(const uint8_t* src, uint32_t w, uint32_t h, int C, uint8_t* dst, uint32_t nw, uint32_t nh, int filter) {
  if (w == 0 || h == 0 || nw == 0 || nh == 0) return -1;
  if (w == nw && h == nh) { memcpy(dst, src, (size_t)w * h * C); return 0; } /* block.rs:279-281 */
  if (filter == PXO_NEAREST) { /* ResizeAlg::Nearest */
    const double xs = (double)w / nw, ys = (double)h / nh;
    for (uint32_t y = 0; y < nh; ++y) {
      const uint32_t sy = std::min(h - 1, (uint32_t)(ys * 0.5 + ys * y));
      for (uint32_t x = 0; x < nw; ++x) {
        const uint32_t sx = std::min(w - 1, (uint32_t)(xs * 0.5 + xs * x));
        memcpy(dst + ((size_t)y * nw + x) * C, src + ((size_t)sy * w + sx) * C, C);
      }
    }
    return 0;
  }
  const bool upscale = nw > w || nh > h; /* block.rs:302 */
  int f;
  switch (filter) {
    case PXO_TRIANGLE: f = upscale ? FIR_BILINEAR : FIR_HAMMING; break;
    case PXO_CATMULLROM: f = FIR_CATMULLROM; break;
    case PXO_GAUSSIAN: f = FIR_GAUSSIAN; break;
    case PXO_LANCZOS3: f = FIR_LANCZOS3; break;
    default: return -1;
  }
  std::vector<uint8_t> in(src, src + (size_t)w * h * C);
  if (C == 4)
    for (size_t i = 0; i < (size_t)w * h; ++i) {
      const uint32_t a = in[i * 4 + 3];
      for (int c = 0; c < 3; ++c) in[i * 4 + c] = mul_div_255(in[i * 4 + c], a);
    }
  std::vector<uint8_t> tmp;
  const uint8_t* hsrc = in.data();
  if (nw != w) {
    FirAxis tx;
    fir_axis(w, nw, f, &tx);
    tmp.resize((size_t)nw * h * C);
    for (uint32_t y = 0; y < h; ++y)
      for (uint32_t x = 0; x < nw; ++x)
        for (int c = 0; c < C; ++c) {
          int32_t ss = 1 << (tx.precision - 1);
          for (uint32_t i = 0; i < tx.count[x]; ++i) ss += (int32_t)in[((size_t)y * w + tx.left[x] + i) * C + c] * tx.k[(size_t)x * tx.stride + i];
          tmp[((size_t)y * nw + x) * C + c] = fir_clip8(ss, tx.precision);
        }
    hsrc = tmp.data();
  }
  std::vector<uint8_t> out((size_t)nw * nh * C);
  if (nh != h) {
    FirAxis ty;
    fir_axis(h, nh, f, &ty);
    for (uint32_t y = 0; y < nh; ++y)
      for (uint32_t x = 0; x < nw; ++x)
        for (int c = 0; c < C; ++c) {
          int32_t ss = 1 << (ty.precision - 1);
          for (uint32_t i = 0; i < ty.count[y]; ++i) ss += (int32_t)hsrc[((size_t)(ty.left[y] + i) * nw + x) * C + c] * ty.k[(size_t)y * ty.stride + i];
          out[((size_t)y * nw + x) * C + c] = fir_clip8(ss, ty.precision);
        }
  } else {
    memcpy(out.data(), hsrc, out.size());
  }
  if (C == 4)
    for (size_t i = 0; i < (size_t)nw * nh; ++i) {
      const uint32_t a = out[i * 4 + 3];
      for (int c = 0; c < 3; ++c) out[i * 4 + c] = a == 0 ? 0 : (uint8_t)std::min<uint32_t>(255u, (out[i * 4 + c] * 255u + a / 2) / a);
    }
  memcpy(dst, out.data(), out.size());
  return 0;
}

/* which branch of PixlzrBlock::resize the drivers below use: 0 = image crate (pinned), 1 = fast_image_resize (unpinned) */
int g_resize_semantics = 0;
inline int resize_block(const uint8_t* src, uint32_t w, uint32_t h, int C, uint8_t* dst, uint32_t nw, uint32_t nh, int filter) {
  return g_resize_semantics ? resize_fir(src, w, h, C, dst, nw, nh, filter) : resize_image_rs(src, w, h, C, dst, nw, nh, filter);
}

inline uint32_t ceil_div_f64(uint32_t a, uint32_t b) { /* split.rs:45-46 (f64 ceil) */
  return (uint32_t)ceil((double)a / (double)b);
}

/* ===================================================================================
 * QOI, qoi crate 0.4.1 (encode_impl / decode_impl) — spec-order ops plus the crate's
 * run-of-1 -> QOI_OP_INDEX quirk.
 * =================================================================================== */
struct Px { uint8_t r, g, b, a; };
inline bool px_eq(Px x, Px y) { return x.r == y.r && x.g == y.g && x.b == y.b && x.a == y.a; }
inline uint8_t px_hash(Px p) { return (uint8_t)((p.r * 3 + p.g * 5 + p.b * 7 + p.a * 11) % 64); }

struct ByteSink {
  uint8_t* out; size_t cap; size_t n;
  void put(uint8_t b) { if (out && n < cap) out[n] = b; ++n; }
};

int64_t qoi_encode(const uint8_t* data, uint32_t w, uint32_t h, int C, uint8_t* out, size_t cap) {
  ByteSink s{out, cap, 0};
  const uint8_t magic[4] = {'q', 'o', 'i', 'f'};
  for (uint8_t b : magic) s.put(b);
  for (int i = 3; i >= 0; --i) s.put((uint8_t)(w >> (8 * i)));
  for (int i = 3; i >= 0; --i) s.put((uint8_t)(h >> (8 * i)));
  s.put((uint8_t)C);
  s.put(0); /* ColorSpace::Srgb */
  Px index[64];
  memset(index, 0, sizeof(index));
  Px prev{0, 0, 0, 255};
  uint8_t hash_prev = px_hash(prev);
  uint32_t run = 0;
  bool index_allowed = false;
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* p = data + i * C;
    Px px{p[0], p[1], p[2], (uint8_t)(C == 4 ? p[3] : 255)};
    if (px_eq(px, prev)) {
      ++run;
      if (run == 62 || i == n - 1) {
        s.put((uint8_t)(0xC0 | (run - 1)));
        run = 0;
      }
    } else {
      if (run != 0) {
        if (run == 1 && index_allowed) s.put((uint8_t)(0x00 | hash_prev));
        else s.put((uint8_t)(0xC0 | (run - 1)));
        run = 0;
      }
      index_allowed = true;
      hash_prev = px_hash(px);
      if (px_eq(index[hash_prev], px)) {
        s.put((uint8_t)(0x00 | hash_prev));
      } else {
        index[hash_prev] = px;
        if (px.a == prev.a) {
          const uint8_t vr = (uint8_t)(px.r - prev.r), vg = (uint8_t)(px.g - prev.g), vb = (uint8_t)(px.b - prev.b);
          const uint8_t vg32 = (uint8_t)(vg + 32);
          if ((vg32 | 63) == 63) {
            const uint8_t vg_r = (uint8_t)(vr - vg), vg_b = (uint8_t)(vb - vg);
            const uint8_t vr2 = (uint8_t)(vr + 2), vg2 = (uint8_t)(vg + 2), vb2 = (uint8_t)(vb + 2);
            if ((vr2 | vg2 | vb2 | 3) == 3) {
              s.put((uint8_t)(0x40 | (vr2 << 4) | (vg2 << 2) | vb2));
            } else {
              const uint8_t vgr8 = (uint8_t)(vg_r + 8), vgb8 = (uint8_t)(vg_b + 8);
              if ((vgr8 | vgb8 | 15) == 15) {
                s.put((uint8_t)(0x80 | vg32));
                s.put((uint8_t)((vgr8 << 4) | vgb8));
              } else {
                s.put(0xFE); s.put(px.r); s.put(px.g); s.put(px.b);
              }
            }
          } else {
            s.put(0xFE); s.put(px.r); s.put(px.g); s.put(px.b);
          }
        } else {
          s.put(0xFF); s.put(px.r); s.put(px.g); s.put(px.b); s.put(px.a);
        }
      }
      prev = px;
    }
  }
  const uint8_t pad[8] = {0, 0, 0, 0, 0, 0, 0, 1};
  for (uint8_t b : pad) s.put(b);
  if (!out || s.n > cap) return -(int64_t)s.n;
  return (int64_t)s.n;
}

int qoi_decode(const uint8_t* d, size_t len, uint32_t* w, uint32_t* h, int* ch, uint8_t* out, size_t cap) {
  if (len < 14 + 8 || memcmp(d, "qoif", 4) != 0) return -1;
  uint32_t W = (uint32_t)d[4] << 24 | (uint32_t)d[5] << 16 | (uint32_t)d[6] << 8 | d[7];
  uint32_t H = (uint32_t)d[8] << 24 | (uint32_t)d[9] << 16 | (uint32_t)d[10] << 8 | d[11];
  int C = d[12];
  if (C != 3 && C != 4) return -1;
  *w = W; *h = H; *ch = C;
  const size_t n = (size_t)W * H;
  if (!out) return 0;
  if (cap < n * C) return -2;
  Px index[64];
  memset(index, 0, sizeof(index));
  Px px{0, 0, 0, 255};
  size_t p = 14, end = len - 8;
  uint32_t run = 0;
  for (size_t i = 0; i < n; ++i) {
    if (run > 0) {
      --run;
    } else if (p < end) {
      uint8_t b1 = d[p++];
      if (b1 == 0xFE) { px.r = d[p]; px.g = d[p + 1]; px.b = d[p + 2]; p += 3; }
      else if (b1 == 0xFF) { px.r = d[p]; px.g = d[p + 1]; px.b = d[p + 2]; px.a = d[p + 3]; p += 4; }
      else if ((b1 & 0xC0) == 0x00) { px = index[b1]; }
      else if ((b1 & 0xC0) == 0x40) {
        px.r = (uint8_t)(px.r + ((b1 >> 4) & 3) - 2);
        px.g = (uint8_t)(px.g + ((b1 >> 2) & 3) - 2);
        px.b = (uint8_t)(px.b + (b1 & 3) - 2);
      } else if ((b1 & 0xC0) == 0x80) {
        uint8_t b2 = d[p++];
        int vg = (b1 & 0x3f) - 32;
        px.r = (uint8_t)(px.r + vg - 8 + ((b2 >> 4) & 0x0f));
        px.g = (uint8_t)(px.g + vg);
        px.b = (uint8_t)(px.b + vg - 8 + (b2 & 0x0f));
      } else {
        run = (b1 & 0x3f);
      }
      index[px_hash(px)] = px;
    }
    uint8_t* o = out + i * C;
    o[0] = px.r; o[1] = px.g; o[2] = px.b;
    if (C == 4) o[3] = px.a;
  }
  return 0;
}

inline void put_u32be(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
inline uint32_t get_u32be(const uint8_t* d) { return (uint32_t)d[0] << 24 | (uint32_t)d[1] << 16 | (uint32_t)d[2] << 8 | d[3]; }

}  // namespace

/* ===================================================================================== */
extern "C" {

void pxo_grid(uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, uint32_t* cols, uint32_t* rows) {
  *cols = ceil_div_f64(w, bw);
  *rows = ceil_div_f64(h, bh);
}
void pxo_srgb_lut(float out[256]) { memcpy(out, g_lut.v, sizeof(g_lut.v)); }
float pxo_cbrtf(float x) { return cbrtf_glibc_old(x); }
void pxo_oklab(uint8_t r, uint8_t g, uint8_t b, float out[3]) {
  Lab c = oklab_from_srgb8(r, g, b);
  out[0] = c.l; out[1] = c.a; out[2] = c.b;
}
float pxo_block_mad(const uint8_t* px, size_t pitch, uint32_t w, uint32_t h, int channels) {
  return block_mad(px, pitch, w, h, channels);
}
int pxo_block_sobel(const uint8_t* px, size_t pitch, uint32_t w, uint32_t h, int channels, float* hz, float* vr) {
  return block_sobel(px, pitch, w, h, channels, hz, vr);
}
float pxo_parse_value(float v) { return parse_value(v); }
int32_t pxo_level_exp(float parsed_v) { return level_exp(parsed_v); }

void pxo_reduce_dims(float v0, float v1, uint32_t w, uint32_t h, uint32_t* ow, uint32_t* oh, float* stored) {
  /* reduce_image_section, operations.rs:140-156 */
  float p0 = parse_value(v0), p1 = parse_value(v1);
  *ow = scaled_dim(w, level_exp(p0));
  *oh = scaled_dim(h, level_exp(p1));
  if (stored) *stored = stored_value(p0, p1);
}

int pxo_resize(const uint8_t* src, uint32_t w, uint32_t h, int channels, uint8_t* dst, uint32_t nw, uint32_t nh, int filter) {
  return resize_image_rs(src, w, h, channels, dst, nw, nh, filter);
}
int pxo_resize_fir(const uint8_t* src, uint32_t w, uint32_t h, int channels, uint8_t* dst, uint32_t nw, uint32_t nh, int filter) {
  return resize_fir(src, w, h, channels, dst, nw, nh, filter);
}
void pxo_set_resize_semantics(int fir) { g_resize_semantics = fir ? 1 : 0; }
int pxo_get_resize_semantics(void) { return g_resize_semantics; }

int pxo_axis_weights(uint32_t n, uint32_t nn, int filter, uint32_t* left, uint32_t* count, float* weights, uint32_t max_taps) {
  Filter flt;
  if (!get_filter(filter, &flt) || n == 0 || nn == 0) return -1;
  AxisTaps t;
  axis_taps(n, nn, flt, &t);
  if (t.stride > max_taps) return -(int)t.stride;
  for (uint32_t o = 0; o < nn; ++o) {
    left[o] = t.left[o];
    count[o] = t.count[o];
    for (uint32_t i = 0; i < max_taps; ++i) weights[(size_t)o * max_taps + i] = i < t.stride ? t.w[(size_t)o * t.stride + i] : 0.0f;
  }
  return (int)t.stride;
}

int pxo_analyze(const uint8_t* img, uint32_t w, uint32_t h, int C, size_t pitch, uint32_t bw, uint32_t bh,
                int metric, float* vx, float* vy, int nthreads) {
  if (!img || !vx || (C != 3 && C != 4) || bw == 0 || bh == 0 || w == 0 || h == 0) return -1;
  const uint32_t cols = ceil_div_f64(w, bw), rows = ceil_div_f64(h, bh);
  int err = 0;
  const int64_t nb = (int64_t)cols * rows;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int64_t bi = 0; bi < nb; ++bi) {
    const uint32_t bx = (uint32_t)(bi % cols), by = (uint32_t)(bi / cols);
    const uint32_t x0 = bx * bw, y0 = by * bh;
    const uint32_t tw = std::min(bw, w - x0), th = std::min(bh, h - y0); /* split.rs:18-19 */
    const uint8_t* p = img + (size_t)y0 * pitch + (size_t)x0 * C;
    if (metric == PXO_METRIC_OKLAB_MAD) {
      vx[bi] = block_mad(p, pitch, tw, th, C);
      if (vy) vy[bi] = vx[bi];
    } else {
      float hz = 0, vr = 0;
      if (block_sobel(p, pitch, tw, th, C, &hz, &vr) != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
        err = 1;
      }
      vx[bi] = hz;
      if (vy) vy[bi] = vr;
    }
  }
  return err ? -5 : 0;
}

/* EXTENSION (include/pixlzr_b200.h "strategy"; the reference only logged the experiment, strategies.txt:1-118):
   bucket of a block = floor(64 * stored value / sqrt(2)), one f32 multiply, clamped to [0, 64]. */
uint32_t pxo_strategy_bucket(float value) {
  const float t = value * 45.25483322143555f;
  if (!(t > 0.0f)) return 0;
  if (t >= 64.0f) return 64;
  return (uint32_t)t;
}

static int64_t shrink_impl(const uint8_t* img, uint32_t w, uint32_t h, int C, size_t pitch, uint32_t bw, uint32_t bh,
                           int metric, float factor, int use_factor, int filter_down, const uint8_t* down_by_bucket,
                           int normalise_global, pxo_block_desc* descs, uint8_t* payload, int nthreads) {
  if (!descs || !payload) return -1;
  const uint32_t cols = ceil_div_f64(w, bw), rows = ceil_div_f64(h, bh);
  const int64_t nb = (int64_t)cols * rows;
  std::vector<float> vx(nb), vy(nb);
  int rc = pxo_analyze(img, w, h, C, pitch, bw, bh, metric, vx.data(), vy.data(), nthreads);
  if (rc != 0) return rc;
  if (normalise_global) {
    /* EXTENSION (no reference semantics; DESIGN.md "global normalisation"): per metric
       component, v' = (v - min) / (max - min) over all raw block values (0 if max == min),
       NaNs ignored for min/max, applied before `after`. */
    for (int comp = 0; comp < (metric == PXO_METRIC_SOBEL_DIR ? 2 : 1); ++comp) {
      std::vector<float>& v = comp ? vy : vx;
      float mn = INFINITY, mx = -INFINITY;
      for (float x : v) { if (x == x) { mn = std::min(mn, x); mx = std::max(mx, x); } }
      float range = mx - mn;
      for (float& x : v) x = (range > 0.0f) ? (x - mn) / range : 0.0f;
    }
    if (metric == PXO_METRIC_OKLAB_MAD) vy = vx;
  }
  /* value -> dims (serial; the payload offset of a block depends on all earlier blocks) */
  uint64_t off = 0;
  std::vector<uint32_t> ow(nb), oh(nb), tw(nb), th(nb);
  for (int64_t bi = 0; bi < nb; ++bi) {
    const uint32_t bx = (uint32_t)(bi % cols), by = (uint32_t)(bi / cols);
    tw[bi] = std::min(bw, w - bx * bw);
    th[bi] = std::min(bh, h - by * bh);
    float v0, v1;
    if (metric == PXO_METRIC_OKLAB_MAD) {
      /* after = x * factor * BASE_FACTOR (pixlzr.rs:15,162) or identity (process/mod.rs:110) */
      float v = use_factor ? vx[bi] * factor * 10.0f : vx[bi];
      v0 = v1 = v;
    } else {
      v0 = vx[bi] * factor; /* pixlzr.rs:199 */
      v1 = vy[bi] * factor;
    }
    float stored;
    pxo_reduce_dims(v0, v1, tw[bi], th[bi], &ow[bi], &oh[bi], &stored);
    if (ow[bi] > 65535u || oh[bi] > 65535u) return -1;
    descs[bi].offset = off;
    descs[bi].value = stored;
    descs[bi].w = (uint16_t)ow[bi];
    descs[bi].h = (uint16_t)oh[bi];
    off += (uint64_t)ow[bi] * oh[bi] * C;
  }
  int err = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int64_t bi = 0; bi < nb; ++bi) {
    const uint32_t bx = (uint32_t)(bi % cols), by = (uint32_t)(bi / cols);
    /* tight copy of the block = crop_imm (split.rs:24) */
    std::vector<uint8_t> blk((size_t)tw[bi] * th[bi] * C);
    for (uint32_t y = 0; y < th[bi]; ++y)
      memcpy(&blk[(size_t)y * tw[bi] * C], img + (size_t)(by * bh + y) * pitch + (size_t)bx * bw * C, (size_t)tw[bi] * C);
    const int filt = down_by_bucket ? down_by_bucket[pxo_strategy_bucket(descs[bi].value)] : filter_down;
    if (resize_block(blk.data(), tw[bi], th[bi], C, payload + descs[bi].offset, ow[bi], oh[bi], filt) != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      err = 1;
    }
  }
  return err ? -1 : (int64_t)off;
}

int64_t pxo_shrink(const uint8_t* img, uint32_t w, uint32_t h, int C, size_t pitch, uint32_t bw, uint32_t bh,
                   int metric, float factor, int use_factor, int filter_down, int normalise_global,
                   pxo_block_desc* descs, uint8_t* payload, int nthreads) {
  return shrink_impl(img, w, h, C, pitch, bw, bh, metric, factor, use_factor, filter_down, nullptr, normalise_global, descs,
                     payload, nthreads);
}

int64_t pxo_shrink_strategy(const uint8_t* img, uint32_t w, uint32_t h, int C, size_t pitch, uint32_t bw, uint32_t bh,
                            int metric, float factor, int use_factor, const uint8_t* down_by_bucket, int normalise_global,
                            pxo_block_desc* descs, uint8_t* payload, int nthreads) {
  if (!down_by_bucket) return -1;
  return shrink_impl(img, w, h, C, pitch, bw, bh, metric, factor, use_factor, 0, down_by_bucket, normalise_global, descs,
                     payload, nthreads);
}

static int expand_impl(const pxo_block_desc* descs, const uint8_t* payload, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh,
                       int C, int filter_up, const uint8_t* up_by_bucket, uint8_t* out, size_t out_pitch, int nthreads) {
  /* block grid in f32 as Pixlzr::block_grid_width/height do (pixlzr.rs:37-42) */
  const uint32_t cols = (uint32_t)ceilf((float)w / (float)bw), rows = (uint32_t)ceilf((float)h / (float)bh);
  const int64_t nb = (int64_t)cols * rows;
  const uint32_t trail_w = w % bw, trail_h = h % bh; /* pixlzr.rs:83-85 */
  int err = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int64_t bi = 0; bi < nb; ++bi) {
    const uint32_t bx = (uint32_t)(bi % cols), by = (uint32_t)(bi / cols);
    const uint32_t nw = (bx == cols - 1 && trail_w > 0) ? trail_w : bw; /* :103-107 */
    const uint32_t nh = (by == rows - 1 && trail_h > 0) ? trail_h : bh; /* :92-96 */
    std::vector<uint8_t> blk((size_t)nw * nh * C);
    const int filt = up_by_bucket ? up_by_bucket[pxo_strategy_bucket(descs[bi].value)] : filter_up;
    if (resize_block(payload + descs[bi].offset, descs[bi].w, descs[bi].h, C, blk.data(), nw, nh, filt) != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      err = 1;
      continue;
    }
    /* paste, pixlzr_image.rs:43-54 */
    for (uint32_t y = 0; y < nh; ++y)
      memcpy(out + (size_t)(by * bh + y) * out_pitch + (size_t)bx * bw * C, &blk[(size_t)y * nw * C], (size_t)nw * C);
  }
  return err ? -1 : 0;
}

int pxo_expand(const pxo_block_desc* descs, const uint8_t* payload, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh,
               int C, int filter_up, uint8_t* out, size_t out_pitch, int nthreads) {
  return expand_impl(descs, payload, w, h, bw, bh, C, filter_up, nullptr, out, out_pitch, nthreads);
}

int pxo_expand_strategy(const pxo_block_desc* descs, const uint8_t* payload, uint32_t w, uint32_t h, uint32_t bw,
                        uint32_t bh, int C, const uint8_t* up_by_bucket, uint8_t* out, size_t out_pitch, int nthreads) {
  if (!up_by_bucket) return -1;
  return expand_impl(descs, payload, w, h, bw, bh, C, 0, up_by_bucket, out, out_pitch, nthreads);
}

/* process/tree.rs:23-83, literally: `img` is the (sub-)image of this recursion level, tightly packed. */
static std::vector<uint8_t> tree_rec(const std::vector<uint8_t>& img, uint32_t w, uint32_t h, int C, float threshold,
                                     uint32_t bw, uint32_t bh, uint32_t min_bw, uint32_t min_bh, int fdown, int fup) {
  const uint32_t mbw = std::max(min_bw, 4u), mbh = std::max(min_bh, 4u); /* :33-34 */
  if (bw <= mbw || bh <= mbh) return img;                                 /* :35-37 image.clone() */
  const bool is_positive = threshold >= 0.0f;                             /* :38 */
  const float thr = fabsf(threshold);                                     /* :39 */
  std::vector<uint8_t> out((size_t)w * h * C, 0);
  const uint32_t cols = ceil_div_f64(w, bw), rows = ceil_div_f64(h, bh);
  for (uint32_t by = 0; by < rows; ++by) {
    for (uint32_t bx = 0; bx < cols; ++bx) {
      const uint32_t x0 = bx * bw, y0 = by * bh;
      const uint32_t w0 = std::min(bw, w - x0), h0 = std::min(bh, h - y0);
      std::vector<uint8_t> blk((size_t)w0 * h0 * C);
      for (uint32_t y = 0; y < h0; ++y) memcpy(&blk[(size_t)y * w0 * C], &img[((size_t)(y0 + y) * w + x0) * C], (size_t)w0 * C);
      const float value = block_mad(blk.data(), (size_t)w0 * C, w0, h0, C); /* after = identity, :97 */
      std::vector<uint8_t> res;
      if ((value >= thr) ^ is_positive) { /* :56 */
        uint32_t ow, oh;
        pxo_reduce_dims(value, value, w0, h0, &ow, &oh, nullptr);
        std::vector<uint8_t> small((size_t)ow * oh * C);
        resize_block(blk.data(), w0, h0, C, small.data(), ow, oh, fdown);
        res.resize((size_t)w0 * h0 * C);
        resize_block(small.data(), ow, oh, C, res.data(), w0, h0, fup);
      } else {
        /* :68-76 — note that the recursion receives the ABSOLUTE threshold */
        res = tree_rec(blk, w0, h0, C, thr, bw >> 1, bh >> 1, mbw, mbh, fdown, fup);
      }
      for (uint32_t y = 0; y < h0; ++y) memcpy(&out[((size_t)(y0 + y) * w + x0) * C], &res[(size_t)y * w0 * C], (size_t)w0 * C);
    }
  }
  return out;
}

int pxo_tree_process(const uint8_t* img, uint32_t w, uint32_t h, int C, size_t pitch, float threshold, uint32_t bw, uint32_t bh,
                     uint32_t min_bw, uint32_t min_bh, int filter_down, int filter_up, uint8_t* out, size_t out_pitch) {
  if (!img || !out || (C != 3 && C != 4) || w == 0 || h == 0 || bw == 0 || bh == 0) return -1;
  std::vector<uint8_t> tight((size_t)w * h * C);
  for (uint32_t y = 0; y < h; ++y) memcpy(&tight[(size_t)y * w * C], img + (size_t)y * pitch, (size_t)w * C);
  std::vector<uint8_t> res = tree_rec(tight, w, h, C, threshold, bw, bh, min_bw, min_bh, filter_down, filter_up);
  for (uint32_t y = 0; y < h; ++y) memcpy(out + (size_t)y * out_pitch, &res[(size_t)y * w * C], (size_t)w * C);
  return 0;
}

int64_t pxo_qoi_encode(const uint8_t* px, uint32_t w, uint32_t h, int channels, uint8_t* out, size_t cap) {
  return qoi_encode(px, w, h, channels, out, cap);
}
int pxo_qoi_decode(const uint8_t* data, size_t len, uint32_t* w, uint32_t* h, int* channels, uint8_t* out, size_t cap) {
  return qoi_decode(data, len, w, h, channels, out, cap);
}

int64_t pxo_container_encode(uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, int filter, int C,
                             const pxo_block_desc* descs, const uint8_t* payload, const uint8_t* value_present,
                             uint8_t* out, size_t cap) {
  /* Pixlzr::encode_to_vec, encoding/mod.rs:40-89; encode_block :168-200 */
  const uint32_t cols = (uint32_t)ceilf((float)w / (float)bw), rows = (uint32_t)ceilf((float)h / (float)bh);
  std::vector<uint8_t> o;
  const char* magic = "PIXLZR";
  o.insert(o.end(), magic, magic + 6);
  o.push_back(0); o.push_back(0); o.push_back(2);
  o.push_back((uint8_t)filter);
  put_u32be(o, w); put_u32be(o, h); put_u32be(o, bw); put_u32be(o, bh);
  std::vector<std::vector<uint8_t>> blocks((size_t)cols * rows);
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    std::vector<uint8_t>& b = blocks[bi];
    const char* bm = "block";
    b.insert(b.end(), bm, bm + 5);
    float val = (value_present && !value_present[bi]) ? 0.0f : descs[bi].value;
    uint32_t bits;
    memcpy(&bits, &val, 4);
    put_u32be(b, bits);
    int64_t need = -qoi_encode(payload + descs[bi].offset, descs[bi].w, descs[bi].h, C, nullptr, 0);
    std::vector<uint8_t> q((size_t)need);
    qoi_encode(payload + descs[bi].offset, descs[bi].w, descs[bi].h, C, q.data(), q.size());
    put_u32be(b, (uint32_t)(need - 4));
    b.insert(b.end(), q.begin() + 4, q.end());
  }
  for (uint32_t r = 0; r < rows; ++r) {
    size_t sum = 0;
    for (uint32_t c = 0; c < cols; ++c) sum += blocks[(size_t)r * cols + c].size();
    put_u32be(o, (uint32_t)sum);
  }
  for (auto& b : blocks) o.insert(o.end(), b.begin(), b.end());
  if (!out || o.size() > cap) return -(int64_t)o.size();
  memcpy(out, o.data(), o.size());
  return (int64_t)o.size();
}

int pxo_container_decode(const uint8_t* d, size_t len, uint32_t* w, uint32_t* h, uint32_t* bw, uint32_t* bh,
                         int* filter, int* channels, uint64_t* payload_bytes, pxo_block_desc* descs, uint8_t* payload) {
  /* Pixlzr::decode_from_vec, encoding/mod.rs:95-165; decode_block :202-242 */
  if (len < 26 || memcmp(d, "PIXLZR", 6) != 0) return -1;
  const uint32_t ver = (uint32_t)d[6] << 16 | (uint32_t)d[7] << 8 | d[8];
  size_t p = 9;
  *filter = -1;
  if (ver >= 1) *filter = d[p++]; /* "filter" since 0.0.1 */
  *w = get_u32be(d + p); *h = get_u32be(d + p + 4); *bw = get_u32be(d + p + 8); *bh = get_u32be(d + p + 12);
  p += 16;
  if (*bw == 0 || *bh == 0) return -1;
  const uint32_t cols = (uint32_t)ceilf((float)*w / (float)*bw), rows = (uint32_t)ceilf((float)*h / (float)*bh);
  if (ver < 2) return -3; /* line sizes since 0.0.2; older layouts not restated */
  if (len < p + (size_t)rows * 4) return -1;
  size_t total = 0;
  for (uint32_t r = 0; r < rows; ++r) total += get_u32be(d + p + (size_t)r * 4);
  p += (size_t)rows * 4;
  if (p + total != len) return -2; /* assert_eq!(reader.data.len(), ...) :141 */
  uint64_t off = 0;
  int C = 0;
  for (size_t bi = 0; bi < (size_t)cols * rows; ++bi) {
    if (p + 13 > len || memcmp(d + p, "block", 5) != 0) return -1;
    uint32_t bits = get_u32be(d + p + 5);
    float val;
    memcpy(&val, &bits, 4);
    uint32_t qlen = get_u32be(d + p + 9);
    p += 13;
    if (p + qlen > len) return -1;
    std::vector<uint8_t> q(4 + (size_t)qlen);
    memcpy(q.data(), "qoif", 4);
    memcpy(q.data() + 4, d + p, qlen);
    p += qlen;
    uint32_t qw, qh;
    int qc;
    if (qoi_decode(q.data(), q.size(), &qw, &qh, &qc, nullptr, 0) != 0) return -1;
    if (C == 0) C = qc;
    if (qc != C) return -4; /* mixed channel counts are not representable in one payload */
    if (descs) {
      descs[bi].offset = off;
      descs[bi].value = val;
      descs[bi].w = (uint16_t)qw;
      descs[bi].h = (uint16_t)qh;
      if (payload && qoi_decode(q.data(), q.size(), &qw, &qh, &qc, payload + off, (size_t)qw * qh * qc) != 0) return -1;
    }
    off += (uint64_t)qw * qh * qc;
  }
  *channels = C;
  *payload_bytes = off;
  return 0;
}

}  // extern "C"
