"""Diagnostic (GPU): deviation of the fast Oklab-MAD path from the reference-order path, and how
many blocks fall into the guard band.  Run under gpurun."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixlzr_b200 as P
from PIL import Image

N = P.native
ctx = N.Context(0)
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([128 + 96 * np.sin(xx / 9700.0 + seed), 128 + 96 * np.cos(yy / 13100.0), 128 + 64 * np.sin((xx + yy) / 6100.0)], -1)
    amp = rng.choice([0, 1, 2, 4, 8, 16, 32, 64], size=((h + 63) // 64, (w + 63) // 64)).astype(np.float32)
    amp = np.kron(amp, np.ones((64, 64), np.float32))[:h, :w]
    img = base + (rng.random((h, w, 3), dtype=np.float32) - 0.5) * 2 * amp[..., None]
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(np.concatenate([img, np.full((h, w, 1), 255, np.uint8)], -1))


cases = [("Big-Ruscher.png", np.array(Image.open(os.path.join(G, "Big-Ruscher.png")))),
         ("base.png", np.array(Image.open(os.path.join(G, "base.png")))),
         ("image.png", np.array(Image.open(os.path.join(G, "image.png")))),
         ("synth8k", synth(7680, 4320, 1))]
for name, img in cases:
    d = ctx.image_upload(img)
    for bs in (16, 32, 64, 128):
        fast, _ = d.analyze(bs, bs, N.METRIC_OKLAB_MAD, 0)
        exact, _ = d.analyze(bs, bs, N.METRIC_OKLAB_MAD, N.FLAG_EXACT_VALUES)
        diff = np.abs(fast.astype(np.float64) - exact)
        rel = diff / np.maximum(exact, 1e-12)
        big = exact > 1e-3
        print(f"{name:16s} bs={bs:3d} blocks={len(fast):6d} max|d|={diff.max():.3e} p99|d|={np.quantile(diff,0.99):.3e} "
              f"max rel(v>1e-3)={rel[big].max() if big.any() else 0:.3e} mean|d|={diff.mean():.3e}")
    d.free()

# timing of the stages on the 8K image
img = cases[-1][1]
d = ctx.image_upload(img)
out = ctx.image_alloc(7680, 4320, 4)
for filt in (4, 2, 0):
    for it in range(3):
        pl = d.shrink(64, 64, N.METRIC_OKLAB_MAD, 1.0, filt, 0)
        pl.expand_to_image(filt, out)
        pl.free()
    ctx.profile_enable(True)
    t0 = time.time()
    for it in range(10):
        pl = d.shrink(64, 64, N.METRIC_OKLAB_MAD, 1.0, filt, 0)
        pl.expand_to_image(filt, out)
        info = pl.info()
        pl.free()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    print("filter", filt, "payload frac", info["bytes"] / img.size, "wall ms/iter", (time.time() - t0) * 100)
    for k, (ms, n) in prof.items():
        if n:
            print(f"   {k:18s} {ms / n * 1000:9.1f} us x{n}")
# fused multiply-add resample (opt-in)
ctx.set_fast_resample(True)
for it in range(2):
    pl = d.shrink(64, 64, N.METRIC_OKLAB_MAD, 1.0, 4, 0); pl.expand_to_image(4, out); pl.free()
ctx.profile_enable(True)
for it in range(10):
    pl = d.shrink(64, 64, N.METRIC_OKLAB_MAD, 1.0, 4, 0); pl.expand_to_image(4, out); pl.free()
print("fma resample:", {k: round(ms / n * 1000, 1) for k, (ms, n) in ctx.profile_read().items() if n})
ctx.profile_enable(False)
ctx.set_fast_resample(False)
# exact-all timing
ctx.profile_enable(True)
for it in range(3):
    pl = d.shrink(64, 64, N.METRIC_OKLAB_MAD, 1.0, 4, N.FLAG_EXACT_VALUES)
    pl.free()
print("exact-all:", {k: round(ms / n * 1000, 1) for k, (ms, n) in ctx.profile_read().items() if n})
ctx.profile_enable(False)
# sobel timing
ctx.profile_enable(True)
for it in range(3):
    pl = d.shrink(64, 64, N.METRIC_SOBEL_DIR, 8.0, 4, 0)
    pl.free()
print("sobel:", {k: round(ms / n * 1000, 1) for k, (ms, n) in ctx.profile_read().items() if n})
