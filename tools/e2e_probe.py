"""Stage timings of the end-to-end path (one worker), and scaling with host threads."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import pixlzr_b200 as P
N = P.native
W, H = bench.IMG_W, bench.IMG_H
img_t = torch.from_numpy(bench.synth_image_np(0, W, H)).pin_memory()
img = img_t.numpy()
nblocks = 120 * 68

class Worker:
    def __init__(self):
        self.ctx = N.Context(0)
        self.pin = [torch.empty(nblocks * 16, dtype=torch.uint8).pin_memory(), torch.empty(W * H * 4, dtype=torch.uint8).pin_memory(),
                    torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()]
        self.descs = self.pin[0].numpy().view(N.DESC_DTYPE); self.pixels = self.pin[1].numpy(); self.out = self.pin[2].numpy()
        self.t = np.zeros(6)
    def run(self, a):
        c = self.ctx; t = [time.perf_counter()]
        im = c.image_upload(a); c.synchronize(); t.append(time.perf_counter())
        pl = im.shrink(64, 64, 0, 1.0, 4, 0); c.synchronize(); t.append(time.perf_counter())
        nbytes = pl.download_into(self.descs, self.pixels); t.append(time.perf_counter())
        pl.free(); im.free()
        pl2 = c.payload_upload(W, H, 64, 64, 4, self.descs, self.pixels[:nbytes]); t.append(time.perf_counter())
        pl2.expand_into(4, self.out); t.append(time.perf_counter())
        pl2.free()
        self.t += np.diff(np.array(t + [time.perf_counter()]))

w = Worker()
for _ in range(3): w.run(img)
w.t[:] = 0
for _ in range(10): w.run(img)
print("stages ms (upload, shrink, download, payload_upload, expand+download, free):", np.round(w.t / 10 * 1e3, 3), "total", round(w.t.sum() / 10 * 1e3, 3))

for nt in (1, 2, 3, 4):
    ws = [Worker() for _ in range(nt)]
    for x in ws: x.run(img)
    def loop(x):
        for _ in range(8): x.run(img)
    th = [threading.Thread(target=loop, args=(x,)) for x in ws]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(nt, "threads: %.2f ms per image, %.1f MP/s" % (dt / (8 * nt) * 1e3, 8 * nt * W * H / 1e6 / dt))
