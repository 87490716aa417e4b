"""ncu driver: encode+decode of an 8K RGBA frame whose tiles all land on one level (noise amplitude `amp`, see
tools/class_times.py).  python tools/prof_level.py [amp] [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixlzr_b200 as P

N = P.native
amp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H = 7680, 4320
rng = np.random.default_rng(0)
rgb = np.clip(np.rint(128 + (rng.random((H, W, 3), dtype=np.float32) - 0.5) * 2 * amp), 0, 255).astype(np.uint8)
img = np.ascontiguousarray(np.concatenate([rgb, np.full((H, W, 1), 255, np.uint8)], -1))
ctx = N.Context(0)
d = ctx.image_upload(img)
out = ctx.image_alloc(W, H, 4)
for _ in range(iters):
    pl = d.shrink(64, 64, 0, 1.0, 4, 0)
    pl.expand_to_image(4, out)
    pl.free()
ctx.synchronize()
print("ok", ctx.launch_count(), "launches")
