"""Debug: per-warp timeline of the last k_shrink_warp launch (library built with -DPXZ_WARP_STATS):
    PXZ_LIB=.../libstats.so python tools/warp_stats.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pixlzr_b200 as P

N = P.native
ctx = N.Context(0)
img = bench.synth_image_np(0, bench.IMG_W, bench.IMG_H)
d = ctx.image_upload(img)
for _ in range(3):
    pl = d.shrink(64, 64, 0, 1.0, 4, 0)
    ctx.synchronize()
    pl.free()
lib = ctypes.CDLL(os.environ["PXZ_LIB"])
nw = 148 * 3 * 4
buf = np.zeros(nw * 4, np.uint64)
rc = lib.pxz_debug_warp_stats(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(nw * 4))
s = buf.reshape(nw, 4).astype(np.int64)
t0 = s[:, 0].min()
start, end, sw2, drawn = s[:, 0] - t0, s[:, 1] - t0, s[:, 2] - t0, s[:, 3]
print("rc", rc, "warps", nw)
print(f"start  min {start.min() / 1e3:.1f}  max {start.max() / 1e3:.1f} us")
print(f"sweep2 min {sw2[sw2 > 0].min() / 1e3:.1f}  median {np.median(sw2[sw2 > 0]) / 1e3:.1f}  max {sw2.max() / 1e3:.1f} us")
print(f"end    min {end.min() / 1e3:.1f}  median {np.median(end) / 1e3:.1f}  max {end.max() / 1e3:.1f} us")
print(f"drawn  min {drawn.min()} mean {drawn.mean():.1f} max {drawn.max()}")
print("busy fraction (sum of warp lifetimes / (warps * kernel span))", float((end - start).sum()) / (nw * end.max()))
