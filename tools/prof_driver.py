"""Small driver for ncu: a few encode+decode passes over one synthetic 8K RGBA frame (same
workload as bench.py).  python tools/prof_driver.py [iters] [filter_down] [filter_up] [metric]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pixlzr_b200 as P

N = P.native
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fd = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fu = int(sys.argv[3]) if len(sys.argv) > 3 else 4
metric = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = N.Context(0)
img = bench.synth_image_np(0, bench.IMG_W, bench.IMG_H)
d = ctx.image_upload(img)
out = ctx.image_alloc(bench.IMG_W, bench.IMG_H, 4)
for _ in range(iters):
    pl = d.shrink(64, 64, metric, 1.0 if metric == 0 else 8.0, fd, 0)
    pl.expand_to_image(fu, out)
    pl.free()
ctx.synchronize()
print("ok", ctx.launch_count(), "launches")
