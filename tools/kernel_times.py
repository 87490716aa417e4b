"""Per-kernel CUDA-event times on the bench workload (one synthetic 8K RGBA frame, bs 64, Lanczos3 both ways):
    [PXZ_LIB=variant.so] python tools/kernel_times.py [reps] [fast_resample 0/1] [metric 0 = Oklab MAD / 1 = Sobel]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pixlzr_b200 as P

N = P.native
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = N.Context(0)
if len(sys.argv) > 2 and int(sys.argv[2]):
    ctx.set_fast_resample(True)
metric = int(sys.argv[3]) if len(sys.argv) > 3 else 0
factor = 8.0 if metric == 1 else 1.0
img = bench.synth_image_np(0, bench.IMG_W, bench.IMG_H)
d = ctx.image_upload(img)
out = ctx.image_alloc(bench.IMG_W, bench.IMG_H, 4)
for _ in range(3):
    pl = d.shrink(64, 64, metric, factor, 4, 0); pl.expand_to_image(4, out); pl.free()
ctx.profile_enable(True)
for _ in range(reps):
    pl = d.shrink(64, 64, metric, factor, 4, 0); pl.expand_to_image(4, out); pl.free()
prof = ctx.profile_read()
ctx.profile_enable(False)
tot = sum(ms / n for ms, n in prof.values() if n)
print(os.environ.get("PXZ_LIB", "default"), "  ".join(f"{k}={ms / n * 1e3:6.1f}us" for k, (ms, n) in prof.items() if n), f" total={tot * 1e3:6.1f}us")
