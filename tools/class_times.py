"""Per-level cost of the resample kernels: 8K RGBA frames whose tiles all carry the same noise amplitude, so
every block lands on the same level.  Prints kernel times per level (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixlzr_b200 as P
N = P.native
ctx = N.Context(0)
W, H = 7680, 4320
rng = np.random.default_rng(0)
out = ctx.image_alloc(W, H, 4)
for amp in (0, 1, 2, 4, 8, 16, 32):
    rgb = np.clip(np.rint(128 + (rng.random((H, W, 3), dtype=np.float32) - 0.5) * 2 * amp), 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(np.concatenate([rgb, np.full((H, W, 1), 255, np.uint8)], -1))
    d = ctx.image_upload(img)
    for filt in (4,):
        for _ in range(2):
            pl = d.shrink(64, 64, 0, 1.0, filt, 0); pl.expand_to_image(filt, out); pl.free()
        ctx.profile_enable(True)
        for _ in range(5):
            pl = d.shrink(64, 64, 0, 1.0, filt, 0); pl.expand_to_image(filt, out); info = pl.info(); pl.free()
        prof = ctx.profile_read(); ctx.profile_enable(False)
        descs_w = None
        print(f"amp {amp:3d} payload {info['bytes'] / img.size:.4f}  " + "  ".join(f"{k}={ms / n * 1e3:7.1f}us" for k, (ms, n) in prof.items() if n))
    d.free()
