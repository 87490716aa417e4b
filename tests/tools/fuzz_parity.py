"""Longer seeded fuzz of the whole encode / decode path against the oracle (run under gpurun; not part of the test suite):
random image and block shapes, both channel counts, both metrics, every filter pair, flags, both resample kernel families.
    python tests/tools/fuzz_parity.py [iterations] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O
import pixlzr_b200 as P

N = P.native
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def forced(kind):
    os.environ["PXZ_RESAMPLE_KERNELS"] = kind
    c = N.Context(0)
    del os.environ["PXZ_RESAMPLE_KERNELS"]
    return c


ctxs = {"auto": N.Context(0), "warp": forced("warp"), "cta": forced("cta")}
bad = 0
for it in range(iters):
    c = int(rng.choice([3, 4]))
    bw, bh = int(rng.choice([4, 8, 12, 16, 24, 32, 40, 48, 64, 96, 128])), int(rng.choice([4, 8, 16, 20, 32, 48, 64, 80, 128]))
    w, h = int(rng.integers(1, 500)), int(rng.integers(1, 400))
    if c == 4 and rng.random() < 0.7:
        w = max(4, w // 4 * 4)
    fd, fu = int(rng.integers(0, 5)), int(rng.integers(0, 5))
    metric = int(rng.choice([0, 0, 1]))
    tw_min, th_min = (w % bw) or min(bw, w), (h % bh) or min(bh, h)
    if metric == 1 and (min(bw, w, tw_min) < 2 or min(bh, h, th_min) < 2):
        metric = 0
    factor = float(rng.choice([0.05, 0.2, 1.0, 3.0, -0.5])) * (8.0 if metric == 1 else 1.0)
    flags = int(rng.choice([0, 0, N.FLAG_EXACT_VALUES, N.FLAG_NORMALISE_GLOBAL, N.FLAG_AFTER_IDENTITY]))
    if metric == 1 and flags == N.FLAG_AFTER_IDENTITY:
        flags = 0
    kind = str(rng.choice(["auto", "warp", "cta"]))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    amp = rng.choice([0, 1, 2, 4, 8, 16, 32, 64, 128], size=((h + 31) // 32, (w + 31) // 32)).astype(np.float32)
    amp = np.kron(amp, np.ones((32, 32), np.float32))[:h, :w]
    img = 128 + 90 * np.sin(xx / 37.0 + it)[..., None] * np.ones(c) + (rng.random((h, w, c), dtype=np.float32) - 0.5) * amp[..., None]
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if c == 4 and rng.random() < 0.5:
        img[..., 3] = 255
    img = np.ascontiguousarray(img)
    try:
        ref = O.shrink(img, bw, bh, metric, factor, fd, use_factor=0 if flags == N.FLAG_AFTER_IDENTITY else 1,
                       normalise_global=flags == N.FLAG_NORMALISE_GLOBAL, nthreads=8)
        ctx = ctxs[kind]
        d = ctx.image_upload(img)
        pl = d.shrink(bw, bh, metric, factor, fd, flags)
        descs, px = pl.download()
        out = pl.expand(fu)
        if it % 4 == 0:  # container stage on the device: same bytes as the oracle's writer, and back
            file_bytes = pl.to_container(fd, True)
            back, _ = ctx.payload_from_container(file_bytes)
            bd, bp = back.download()
            back.free()
            if file_bytes != O.container_encode(ref, fd) and (flags == N.FLAG_EXACT_VALUES or (flags == N.FLAG_NORMALISE_GLOBAL and metric == 1)) or not np.array_equal(bp, px) \
                    or not np.array_equal(bd["w"], descs["w"]):
                raise RuntimeError("device container mismatch")
        pl.free(); d.free()
        ok = (np.array_equal(descs["w"], ref.descs["w"]) and np.array_equal(descs["h"], ref.descs["h"]) and np.array_equal(descs["offset"], ref.descs["offset"])
              and np.array_equal(px, ref.payload) and np.array_equal(out, O.expand(ref, fu, nthreads=8)))
        if flags == N.FLAG_EXACT_VALUES or metric == 1:
            ok = ok and np.array_equal(descs["value"].view("<u4"), ref.descs["value"].view("<u4"))
    except Exception as e:  # noqa: BLE001
        ok = False
        print("EXC", repr(e)[:200])
    if not ok:
        bad += 1
        print("MISMATCH", dict(it=it, c=c, w=w, h=h, bw=bw, bh=bh, fd=fd, fu=fu, metric=metric, factor=factor, flags=flags, kind=kind), flush=True)
# quadtree processing (process/tree.rs) on a fifth as many cases
tbad = 0
for it in range(iters // 5):
    c = int(rng.choice([3, 4]))
    k = int(rng.integers(1, 4))
    mw, mh = int(rng.choice([4, 8, 12, 16])), int(rng.choice([4, 8, 16]))
    bw, bh = mw << k, mh << k
    w, h = int(rng.integers(8, 400)), int(rng.integers(8, 300))
    if c == 4:
        w = max(4, w // 4 * 4)
    fd, fu = int(rng.integers(0, 5)), int(rng.integers(0, 5))
    thr = float(rng.choice([0.002, 0.01, 0.03, 0.1, -0.01, -0.05]))
    img = np.clip(np.rint(128 + 100 * np.sin(np.mgrid[0:h, 0:w][1] / 23.0 + it)[..., None] * np.ones(c)
                          + (rng.random((h, w, c)) - 0.5) * np.kron(rng.choice([0, 4, 16, 64], size=((h + 15) // 16, (w + 15) // 16)), np.ones((16, 16)))[:h, :w, None]),
                  0, 255).astype(np.uint8)
    if c == 4:
        img[..., 3] = 255 if rng.random() < 0.5 else img[..., 0]
    img = np.ascontiguousarray(img)
    try:
        want = O.tree_process(img, thr, bw, bh, mw, mh, fd, fu)
        got = P.tree_process_custom(img, thr, (bw, bh), (mw, mh), (P.FilterType(fd), P.FilterType(fu)))
        ok = np.array_equal(got[..., :c], want) and (c == 4 or (got[..., 3] == 255).all())
    except Exception as e:  # noqa: BLE001
        ok = False
        print("EXC", repr(e)[:200])
    if not ok:
        tbad += 1
        print("TREE MISMATCH", dict(it=it, c=c, w=w, h=h, bw=bw, bh=bh, mw=mw, mh=mh, fd=fd, fu=fu, thr=thr), flush=True)
# per-block filter pairs (pxz_ctx_set_strategy) on a third as many cases: random tables, continuous factors (bucket edges),
# fast and exact analysis, the expand both from the device payload and from descriptors that went through the host
sbad = 0
for it in range(iters // 3):
    c = int(rng.choice([3, 4]))
    bw, bh = int(rng.choice([8, 16, 24, 32, 40, 64, 128])), int(rng.choice([8, 16, 20, 32, 64, 80]))
    w, h = int(rng.integers(8, 500)), int(rng.integers(8, 400))
    if c == 4 and rng.random() < 0.7:
        w = max(4, w // 4 * 4)
    metric = int(rng.choice([0, 0, 0, 1]))
    tw_min, th_min = (w % bw) or min(bw, w), (h % bh) or min(bh, h)
    if metric == 1 and (min(bw, w, tw_min) < 2 or min(bh, h, th_min) < 2):
        metric = 0
    factor = float(np.exp(rng.uniform(np.log(0.01), np.log(0.6)))) * (20.0 if metric == 1 else 1.0)
    flags = int(rng.choice([0, 0, N.FLAG_EXACT_VALUES, N.FLAG_AFTER_IDENTITY])) if metric == 0 else 0
    kind = str(rng.choice(["auto", "warp", "cta"]))
    if rng.random() < 0.5:
        down, up = rng.integers(0, 5, 65).astype(np.uint8), rng.integers(0, 5, 65).astype(np.uint8)
    else:  # a few ranges
        edges = np.sort(rng.integers(1, 65, 4))
        down, up = np.zeros(65, np.uint8), np.zeros(65, np.uint8)
        for lo, hi in zip(np.r_[0, edges], np.r_[edges, 65]):
            down[lo:hi], up[lo:hi] = rng.integers(0, 5), rng.integers(0, 5)
    amp = 128.0 * 2.0 ** (-9.0 * rng.random(((h + bh - 1) // bh, (w + bw - 1) // bw)))
    amp = np.kron(amp, np.ones((bh, bw)))[:h, :w, None]
    xx = np.mgrid[0:h, 0:w][1]
    img = np.clip(np.rint(128 + 60 * np.sin(xx / 41.0 + it)[..., None] * np.ones(c) + (rng.random((h, w, c)) - 0.5) * 2 * amp), 0, 255).astype(np.uint8)
    if c == 4 and rng.random() < 0.5:
        img[..., 3] = 255
    img = np.ascontiguousarray(img)
    ctx = ctxs[kind]
    try:
        ref = O.shrink_strategy(img, bw, bh, metric, factor, down, use_factor=0 if flags == N.FLAG_AFTER_IDENTITY else 1, nthreads=8)
        want = O.expand_strategy(ref, up, nthreads=8)
        ctx.set_strategy(down, up)
        d = ctx.image_upload(img)
        pl = d.shrink(bw, bh, metric, factor, 3, flags)
        descs, px = pl.download()
        out = pl.expand(3)
        pl2 = ctx.payload_upload(w, h, bw, bh, c, descs, px)
        out2 = pl2.expand(1)
        pl2.free(); pl.free(); d.free()
        ok = (np.array_equal(descs["w"], ref.descs["w"]) and np.array_equal(descs["h"], ref.descs["h"]) and np.array_equal(px, ref.payload)
              and np.array_equal(out, want) and np.array_equal(out2, want))
    except Exception as e:  # noqa: BLE001
        ok = False
        print("EXC", repr(e)[:200])
    finally:
        ctx.set_strategy(None, None)
    if not ok:
        sbad += 1
        print("STRATEGY MISMATCH", dict(it=it, c=c, w=w, h=h, bw=bw, bh=bh, metric=metric, factor=factor, flags=flags, kind=kind), flush=True)
print(f"{iters} cases, {bad} mismatches; {iters // 5} tree cases, {tbad} mismatches; {iters // 3} strategy cases, {sbad} mismatches")
