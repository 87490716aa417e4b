"""Container stage of one 8K RGBA frame (bench.py's workload): device QOI (pxz_payload_to_container /
pxz_payload_from_container) against the host stage (pxz_container_encode / _decode + payload transfer).
    python tools/container_times.py [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pixlzr_b200 as P

N = P.native
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = N.Context(0)
img = bench.synth_image_np(0, bench.IMG_W, bench.IMG_H)
d = ctx.image_upload(img)
pl = d.shrink(64, 64, 0, 1.0, 4, 0)
info = pl.info()


def timed(fn):
    fn()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, r


def host_encode():
    descs, px = pl.download()
    return N.container_encode(info["w"], info["h"], 64, 64, 4, 4, descs, px)


ctx.profile_enable(True)
t_dev, data = timed(lambda: pl.to_container(4, True))
import torch  # pinned host buffers only
pin = torch.empty(N.lib().pxz_container_bound(info["w"], info["h"], 64, 64, 4, info["bytes"]), dtype=torch.uint8).pin_memory().numpy()
t_pin, n_pin = timed(lambda: pl.to_container_into(pin, 4, True))
assert pin[:n_pin].tobytes() == data
pin_d, pin_p = torch.empty(info["cols"] * info["rows"] * 16, dtype=torch.uint8).pin_memory().numpy().view(N.DESC_DTYPE), torch.empty(info["bytes"], dtype=torch.uint8).pin_memory().numpy()
t_dl, _ = timed(lambda: pl.download_into(pin_d, pin_p))
t_host, data_h = timed(host_encode)
assert data == data_h


def host_decode():
    _, descs, px = N.container_decode(data)
    p = ctx.payload_upload(info["w"], info["h"], 64, 64, 4, descs, px)
    ctx.synchronize()
    p.free()


def dev_decode():
    p, _ = ctx.payload_from_container(data)
    p.free()


t_ddev, _ = timed(dev_decode)
t_dhost, _ = timed(host_decode)
prof = ctx.profile_read()
print("device kernels:", "  ".join(f"{k}={ms / n * 1e3:.0f}us" for k, (ms, n) in prof.items() if n and k.startswith("qoi")))
mp = bench.IMG_W * bench.IMG_H / 1e6
print(f"payload {info['bytes'] / 1e6:.1f} MB -> file {len(data) / 1e6:.1f} MB ({os.cpu_count()} host threads)")
print(f"encode: device {t_dev:.2f} ms ({mp / t_dev * 1e3:.0f} MP/s of source image)   host stage + payload D2H {t_host:.2f} ms")
print(f"encode into a pinned buffer (the C ABI call alone): {t_pin:.2f} ms; raw payload download into pinned memory: {t_dl:.2f} ms")
print(f"decode: device {t_ddev:.2f} ms   host stage + payload H2D {t_dhost:.2f} ms")
