"""A few larger seeded cases (thousands of tiles: the library picks the warp-per-tile kernels and the cost-ordered work
lists by itself) against the oracle.  Run under gpurun:  python tests/tools/fuzz_large.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O
import pixlzr_b200 as P

N = P.native
ctx = N.Context(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
bad = 0
for it in range(cases):
    bs = int(rng.choice([16, 32, 32, 64, 64, 64, 48]))
    w, h = int(rng.integers(1200, 4200)) // 4 * 4, int(rng.integers(900, 3000))
    metric = int(rng.choice([0, 0, 1]))
    if metric == 1 and ((w % bs) == 1 or (h % bs) == 1):
        metric = 0  # the reference panics on blocks thinner than 2 px (operations.rs:220-221)
    fd, fu = int(rng.integers(0, 5)), int(rng.integers(0, 5))
    factor = float(rng.choice([0.3, 1.0, 2.0])) * (8.0 if metric else 1.0)
    flags = int(rng.choice([0, 0, N.FLAG_EXACT_VALUES]))
    amp = np.kron(rng.choice([0, 1, 2, 4, 8, 16, 32, 64, 128], size=((h + 63) // 64, (w + 63) // 64)).astype(np.float32), np.ones((64, 64), np.float32))[:h, :w]
    xx = np.arange(w, dtype=np.float32)[None, :]
    img = np.empty((h, w, 4), np.uint8)
    for c in range(3):
        img[..., c] = np.clip(np.rint(128 + 90 * np.sin(xx / (50.0 + 13 * c) + it) + (rng.random((h, w), dtype=np.float32) - 0.5) * amp), 0, 255)
    img[..., 3] = 255
    if it % 4 == 0:
        img[: h // 3, :, 3] = img[: h // 3, :, 0]
    ref = O.shrink(img, bs, bs, metric, factor, fd, nthreads=16)
    d = ctx.image_upload(img)
    pl = d.shrink(bs, bs, metric, factor, fd, flags)
    descs, px = pl.download()
    out = pl.expand(fu)
    pl.free(); d.free()
    ok = (np.array_equal(descs["w"], ref.descs["w"]) and np.array_equal(descs["h"], ref.descs["h"]) and np.array_equal(px, ref.payload)
          and np.array_equal(out, O.expand(ref, fu, nthreads=16)))
    if not ok:
        bad += 1
    print(dict(it=it, w=w, h=h, bs=bs, tiles=((w + bs - 1) // bs) * ((h + bs - 1) // bs), metric=metric, fd=fd, fu=fu, flags=flags), "ok" if ok else "MISMATCH", flush=True)
print(f"{cases} large cases, {bad} mismatches")
