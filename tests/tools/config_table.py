"""Fills the BASELINE.md results table: for every BASELINE.json config that fits one GPU, encode / decode
throughput on the B200 (device-timed and end-to-end through the C ABI), the CPU oracle on 1 and all host
cores, and the parity of the GPU result against the oracle (grid, dims, values, pixels, PSNR).

    python tests/tools/config_table.py > gpurun_out/config_table.md        (run under gpurun)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle as O  # noqa: E402
import pixlzr_b200 as P  # noqa: E402
from PIL import Image  # noqa: E402

N = P.native
G = os.path.join(ROOT, "tests", "golden")
FN = ["Nearest", "Triangle", "CatmullRom", "Gaussian", "Lanczos3"]
ctx = N.Context(0)
cores = os.cpu_count() or 1


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def gpu_times(img, bs, metric, factor, fd, fu, reps=20):
    h, w, c = img.shape
    d = ctx.image_upload(img)
    out = ctx.image_alloc(w, h, c)
    for _ in range(3):
        pl = d.shrink(bs, bs, metric, factor, fd, 0); pl.expand_to_image(fu, out); pl.free()
    ctx.profile_enable(True)
    for _ in range(reps):
        pl = d.shrink(bs, bs, metric, factor, fd, 0); pl.expand_to_image(fu, out); pl.free()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    enc = sum(ms / n for k, (ms, n) in prof.items() if n and k != "resample_up")
    dec = sum(ms / n for k, (ms, n) in prof.items() if n and k == "resample_up")
    # end to end through host buffers (pageable numpy arrays here; bench.py uses pinned ones)
    t0 = time.perf_counter()
    for _ in range(3):
        d2 = ctx.image_upload(img); pl = d2.shrink(bs, bs, metric, factor, fd, 0); descs, px = pl.download(); pl.free(); d2.free()
        pl2 = ctx.payload_upload(w, h, bs, bs, c, descs, px); got = pl2.expand(fu); pl2.free()
    e2e = (time.perf_counter() - t0) / 3
    pl = d.shrink(bs, bs, metric, factor, fd, 0)
    descs, px = pl.download()
    pl.free(); d.free(); out.free()
    return enc, dec, e2e, descs, px, got


def cpu_times(img, bs, metric, factor, fd, fu):
    res = {}
    for nt in (1, cores):
        O.shrink(img, bs, bs, metric, factor, fd, nthreads=nt)
        t0 = time.perf_counter(); s = O.shrink(img, bs, bs, metric, factor, fd, nthreads=nt); t1 = time.perf_counter()
        o = O.expand(s, fu, nthreads=nt); t2 = time.perf_counter()
        res[nt] = (t1 - t0, t2 - t1)
    return res, s, o


rows = []


def run(name, img, bs, metric, factor, fd, fu):
    mp = img.shape[0] * img.shape[1] / 1e6
    enc, dec, e2e, descs, px, got = gpu_times(img, bs, metric, factor, fd, fu)
    cpu, ref, want = cpu_times(img, bs, metric, factor, fd, fu)
    dims_ok = bool(np.array_equal(descs["w"], ref.descs["w"]) and np.array_equal(descs["h"], ref.descs["h"]))
    vals = float(np.max(np.abs(descs["value"].astype(np.float64) - ref.descs["value"])))
    pix_ok = bool(px.size == ref.payload.size and np.array_equal(px, ref.payload))
    dec_ok = bool(np.array_equal(got, want))
    frac_payload = px.size / img.size
    rows.append(f"| {name} | 1 | enc+dec | {mp / ((enc + dec) * 1e-3):,.0f} (enc {mp / (enc * 1e-3):,.0f} / dec {mp / (dec * 1e-3):,.0f}) | "
                f"{mp / e2e:,.0f} | {mp / sum(cpu[1]):.1f} | {mp / sum(cpu[cores]):.1f} ({cores}) | "
                f"grid ok / dims {'bit-exact' if dims_ok else 'MISMATCH'} / max|dvalue| {vals:.1e} / shrunk px {'bit-exact' if pix_ok else 'DIFF'} / "
                f"decoded px {'bit-exact' if dec_ok else 'DIFF'} (PSNR vs oracle decode {psnr(got, want):.0f} dB; vs source {psnr(got, img):.1f} dB) / payload {100 * frac_payload:.1f}% |")
    print(rows[-1], flush=True)


print("| Config | GPUs | Stage | MP/s device-timed | MP/s end-to-end (pageable host) | CPU 1-thread MP/s | CPU all-core MP/s (N) | Parity |")
print("|---|---|---|---|---|---|---|---|")
base = np.array(Image.open(os.path.join(G, "base.png")))
run("C1 base.png bs64 MAD k=0.25 CatmullRom/CatmullRom", base, 64, 0, 0.25, 2, 2)
run("C1 base.png bs64 MAD k=1 Lanczos3/Lanczos3", base, 64, 0, 1.0, 4, 4)
big = np.array(Image.open(os.path.join(G, "Big-Ruscher.png")))
for bs in (16, 32, 64):
    for fd, fu in [(0, 0), (1, 0), (2, 4), (4, 2), (4, 4)]:
        run(f"C2 Big-Ruscher bs{bs} MAD k=1 {FN[fd]}/{FN[fu]}", big, bs, 0, 1.0, fd, fu)
    run(f"C2 Big-Ruscher bs{bs} Sobel k=8 Lanczos3/Lanczos3", big, bs, 1, 8.0, 4, 4)
img8k = bench.synth_image_np(0, bench.IMG_W, bench.IMG_H)
run("C3 synthetic 8K RGBA bs64 MAD k=1 Lanczos3/Lanczos3", img8k, 64, 0, 1.0, 4, 4)
run("C3 synthetic 8K RGBA bs64 MAD k=1 Nearest/Nearest", img8k, 64, 0, 1.0, 0, 0)
run("C3 synthetic 8K RGBA bs64 Sobel k=8 Lanczos3/Lanczos3", img8k, 64, 1, 8.0, 4, 4)
