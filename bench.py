#!/usr/bin/env python3
"""bench.py — encode+decode megapixels/s of the pixlzr hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], "C3"): synthetic 7680x4320 RGBA8 images, 64x64 blocks, metric
Oklab-MAD with k = 1, Lanczos3 down / Lanczos3 up (the reference CLI's defaults, src/bin/main.rs:19,27-36).
One step = encode (analyse -> plan -> shrink into the packed payload) + decode (expand + paste) of a batch
of `--batch` distinct images per rank; images are independent, so ranks share nothing (weak scaling) and
no collective sits on the data path.  `value` is timed with CUDA events on the launching stream with
every input already resident in HBM; `e2e` runs the same work through the C ABI with pinned HOST
buffers, host<->device copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "encode+decode megapixels/sec"
UNIT = "MP/s"
BS = 64
FILTER_DOWN = 4  # Lanczos3
FILTER_UP = 4
FACTOR = 1.0
IMG_W, IMG_H = 7680, 4320
CPU_SAMPLE_ROWS = 1088  # 17 block rows of the 8K frame = 8.36 MP: the bounded CPU sample


# --------------------------------------------------------------------------------------------------
# synthetic input: slow colour ramps + per-64x64-tile uniform noise whose amplitude
# is picked from {0,1,2,4,8,16,32,64} by a tile hash, so that every level 2^0 .. 2^-6 occurs
# --------------------------------------------------------------------------------------------------
def synth_image_np(seed: int, w: int, h: int) -> np.ndarray:
    rng = np.random.default_rng(0x5049584C5A52 ^ seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([128 + 96 * np.sin(xx / 9000.0 + seed), 128 + 96 * np.cos(yy / 7000.0 + 0.3 * seed),
                     128 + 64 * np.sin((xx + yy) / 11000.0)], -1)
    ty, tx = np.mgrid[0:(h + 63) // 64, 0:(w + 63) // 64].astype(np.uint64)
    hsh = (tx * np.uint64(73856093)) ^ (ty * np.uint64(19349663)) ^ np.uint64((seed * 83492791) & 0xFFFFFFFF)
    amp = np.array([0, 1, 2, 4, 8, 16, 32, 64], np.float32)[(hsh >> np.uint64(3)) % np.uint64(8)]
    amp = np.kron(amp, np.ones((64, 64), np.float32))[:h, :w]
    img = base + (rng.random((h, w, 3), dtype=np.float32) - 0.5) * 2 * amp[..., None]
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(np.concatenate([img, np.full((h, w, 1), 255, np.uint8)], -1))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout: float = 3.0):
        """nvidia-smi needs a few hundred ms before its first line: do not start the timed region before it polls."""
        t = time.time()
        while self.proc is not None and not self.lines and time.time() - t < timeout:
            time.sleep(0.01)

    def samples_since(self, t0: float) -> int:
        return sum(1 for (t, _) in self.lines if t >= t0)

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [ln for (t, ln) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [ln for _, ln in self.lines]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU legs (oracle = C++ port of the reference; the Rust reference itself cannot be built here)
# --------------------------------------------------------------------------------------------------
def cpu_encode_decode(O, sample: np.ndarray, threads: int) -> float:
    t0 = time.perf_counter()
    s = O.shrink(sample, BS, BS, O.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, nthreads=threads)
    O.expand(s, FILTER_UP, nthreads=threads)
    return time.perf_counter() - t0


def cpu_baseline(sample: np.ndarray) -> dict:
    import oracle as O

    cores = os.cpu_count() or 1
    mp = sample.shape[0] * sample.shape[1] / 1e6
    cpu_encode_decode(O, sample, cores)  # warm-up (thread pool, page faults)
    best_all = min(cpu_encode_decode(O, sample, cores) for _ in range(5))
    best_one = min(cpu_encode_decode(O, sample, 1) for _ in range(2))
    return {"value": mp / best_all, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"top {sample.shape[0]} rows ({mp:.2f} MP) of image 0 of the same workload, best of 5",
            "one_thread_value": mp / best_one,
            "note": "shrink* is a serial loop in the reference (pixlzr.rs:163-184); the all-core figure is charitable"}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Rust crate cannot
    be built in this image: no cargo/rustc) on the host cores, same config/metric."""
    if rank != 0:
        return
    import oracle as O

    cores = os.cpu_count() or 1
    sample = synth_image_np(0, IMG_W, CPU_SAMPLE_ROWS)
    mp = sample.shape[0] * sample.shape[1] / 1e6
    for _ in range(args.warmup):
        cpu_encode_decode(O, sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_encode_decode(O, sample, cores)
    dt = time.perf_counter() - t0
    val = mp * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1, sample_rows=CPU_SAMPLE_ROWS),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = top {CPU_SAMPLE_ROWS} rows ({mp:.2f} MP) of one 8K frame of the workload"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch: int, sample_rows: int | None = None) -> dict:
    cfg = {
        "workload": "C3 synthetic 7680x4320 RGBA8 (alpha 255), 64x64 blocks, Oklab-MAD k=1, Lanczos3 down / Lanczos3 up, "
                    "encode (analyse+plan+shrink) + decode (expand+paste)",
        "width": IMG_W, "height": IMG_H, "channels": 4, "block": BS, "metric": "oklab_mad", "factor": FACTOR,
        "filter_down": "Lanczos3", "filter_up": "Lanczos3", "images_per_step_per_gpu": batch,
        "sharding": "independent images per rank, no data-path collective",
        "streams_per_gpu": getattr(args, "streams", 1),
        "cache": "inputs larger than L2 (batch x 132.7 MB per step)",
        "resize_semantics": "image_rs (the branch pinned by the reference's fixtures)",
    }
    if sample_rows:
        cfg["cpu_sample_rows"] = sample_rows
    return cfg


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    import pixlzr_b200 as P

    N = P.native
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    batch = args.batch

    with torch.cuda.stream(stream):
        ctx = N.Context(local_rank, cuda_stream=stream.cuda_stream)
        # inputs: host (pinned) for the e2e leg, device-resident copies for the kernel leg
        host_imgs = []
        for i in range(batch):
            a = synth_image_np(rank * 1000 + i, IMG_W, IMG_H)
            t = torch.from_numpy(a).pin_memory()
            host_imgs.append(t)
        dev_imgs = [t.to(dev, non_blocking=True) for t in host_imgs]
        dev_out = torch.empty((IMG_H, IMG_W, 4), dtype=torch.uint8, device=dev)
        stream.synchronize()
        # device-resident leg: `--streams` contexts (one CUDA stream each) take the images of a step in turn, so the
        # latency-bound kernels of one image (guard-band recompute, plan) overlap the throughput kernels of another
        n_streams = max(1, min(args.streams, batch))
        streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
        ctxs = [ctx] + [N.Context(local_rank, cuda_stream=st.cuda_stream) for st in streams[1:]]
        outs = [dev_out] + [torch.empty_like(dev_out) for _ in range(n_streams - 1)]
        wrapped = [ctxs[i % n_streams].image_wrap(t.data_ptr(), IMG_W, IMG_H, 4, IMG_W * 4) for i, t in enumerate(dev_imgs)]
        wrapped_out = [ctxs[k].image_wrap(outs[k].data_ptr(), IMG_W, IMG_H, 4, IMG_W * 4) for k in range(n_streams)]

        def step_device():
            for i, im in enumerate(wrapped):
                pl = im.shrink(BS, BS, N.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, 0)
                pl.expand_to_image(FILTER_UP, wrapped_out[i % n_streams])
                pl.free()

        # payload sizes (for the algorithmic-byte counts), untimed
        payload_bytes, nblocks = [], 0
        for im in wrapped:
            pl = im.shrink(BS, BS, N.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, 0)
            info = pl.info()
            payload_bytes.append(info["bytes"])
            nblocks = info["cols"] * info["rows"]
            pl.free()

        for _ in range(args.warmup):
            step_device()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                               int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
        sampler.start()
        sampler.wait_first()
        launches0 = sum(c.launch_count() for c in ctxs)
        for c in ctxs:
            c.profile_enable(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        ev0.record(stream)               # the timed region starts on the launching stream ...
        for st in streams[1:]:
            st.wait_event(ev0)           # ... and every other stream starts after it
        for _ in range(args.steps):
            step_device()
        t_launched = time.time()         # host time to issue the K steps (launch-bound when close to the device time)
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            stream.wait_event(e)         # the launching stream joins all the others before the end event
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        t_wall1 = time.time()
        if world > 1:
            dist.barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        prof = {}
        for c in ctxs:
            for name, (ms, n) in c.profile_read().items():
                a, b = prof.get(name, (0.0, 0))
                prof[name] = (a + ms, b + n)
            c.profile_enable(False)
        launches = sum(c.launch_count() for c in ctxs) - launches0
        # the timed region is a few ms, nvidia-smi polls every >= 20 ms: the same steps keep running (untimed, outside the
        # launch count) until at least two polls have seen the GPU under this load
        t_keep = time.time()
        while sampler.proc is not None and sampler.samples_since(t_wall0) < 2 and time.time() - t_keep < 1.0:
            step_device()
            torch.cuda.synchronize()
        t_wall1 = time.time()
        clocks = sampler.stop(t_wall0, t_wall1)
        clocks["window"] = "timed region + the same steps continued (untimed) until two nvidia-smi polls"
        # per-kernel durations without cross-stream contention: the same K steps again on the launching stream only
        prof_overlapped = prof
        if n_streams > 1:
            solo = [ctx.image_wrap(t.data_ptr(), IMG_W, IMG_H, 4, IMG_W * 4) for t in dev_imgs]
            ctx.profile_enable(True)
            ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev2.record(stream)
            for _ in range(args.steps):
                for im in solo:
                    pl = im.shrink(BS, BS, N.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, 0)
                    pl.expand_to_image(FILTER_UP, wrapped_out[0])
                    pl.free()
            ev3.record(stream)
            stream.synchronize()
            solo_ms = ev2.elapsed_time(ev3)
            prof = ctx.profile_read()
            ctx.profile_enable(False)
        else:
            solo_ms = elapsed_ms

        # ---- informative: the opt-in fused-multiply-add resample (pixels within +-1 LSB of the reference, the tolerance
        # BASELINE.json states; tests/test_gpu_parity.py checks it).  Single stream, same K steps; not the headline.
        fused_info = None
        if rank == 0:
            im0 = ctx.image_wrap(dev_imgs[0].data_ptr(), IMG_W, IMG_H, 4, IMG_W * 4)
            ctx.set_fast_resample(True)
            for _ in range(2):
                pl = im0.shrink(BS, BS, N.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, 0)
                pl.expand_to_image(FILTER_UP, wrapped_out[0])
                pl.free()
            ctx.profile_enable(True)
            for _ in range(args.steps):
                pl = im0.shrink(BS, BS, N.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, 0)
                pl.expand_to_image(FILTER_UP, wrapped_out[0])
                pl.free()
            stream.synchronize()
            fprof = ctx.profile_read()
            ctx.profile_enable(False)
            ctx.set_fast_resample(False)
            fused_info = {"note": "pxz_ctx_set_fast_resample(1): taps as one fused multiply-add, pixels within +-1 LSB; single stream",
                          "kernel_us": {k: round(ms / n * 1e3, 2) for k, (ms, n) in fprof.items() if n}}
            fused_info["MPps_single_stream"] = round(IMG_W * IMG_H / 1e6 / (sum(fused_info["kernel_us"].values()) * 1e-6), 1)

        # ---- e2e: same work through the C ABI with pinned host buffers ---------------------------------
        # `E2E_WORKERS` host threads, each with its own context (= its own stream) and pinned staging buffers, take
        # the images of the step in turn, so one image's H2D overlaps another's kernels and D2H (PCIe is full duplex).
        img_bytes = IMG_W * IMG_H * 4
        n_workers = max(1, min(args.e2e_workers, batch))
        np_imgs = [t.numpy() for t in host_imgs]

        class Worker:
            def __init__(self):
                self.ctx = N.Context(local_rank)
                self.pin = [torch.empty(nblocks * 16, dtype=torch.uint8).pin_memory(),
                            torch.empty(img_bytes, dtype=torch.uint8).pin_memory(),
                            torch.empty((IMG_H, IMG_W, 4), dtype=torch.uint8).pin_memory()]
                self.descs = self.pin[0].numpy().view(N.DESC_DTYPE)
                self.pixels, self.out = self.pin[1].numpy(), self.pin[2].numpy()
                self.h2d = self.d2h = 0
                self.busy_s = 0.0

            def encode_decode(self, a):
                c = self.ctx
                t_in = time.perf_counter()
                im = c.image_upload(a)                                                  # H2D image
                pl = im.shrink(BS, BS, N.METRIC_OKLAB_MAD, FACTOR, FILTER_DOWN, 0)
                nbytes = pl.download_into(self.descs, self.pixels)                      # D2H descs + payload (encode result)
                pl.free()
                im.free()
                pl2 = c.payload_upload(IMG_W, IMG_H, BS, BS, 4, self.descs, self.pixels[:nbytes])  # H2D payload
                pl2.expand_into(FILTER_UP, self.out)                                    # D2H decoded image
                pl2.free()
                self.h2d += img_bytes + nbytes + nblocks * 16
                self.d2h += nbytes + nblocks * 16 + img_bytes
                self.busy_s += time.perf_counter() - t_in

        workers = [Worker() for _ in range(n_workers)]

        def run_steps(k: int):
            # worker w takes images w, w + n_workers, ... of every step; steps run back to back (no per-step join)
            def loop(wi: int):
                for _ in range(k):
                    for a in np_imgs[wi::n_workers]:
                        workers[wi].encode_decode(a)
            th = [threading.Thread(target=loop, args=(wi,)) for wi in range(n_workers)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        e2e_steps = max(2, min(args.steps, 8))
        run_steps(2)
        for wk in workers:
            wk.h2d = wk.d2h = 0
            wk.busy_s = 0.0
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        run_steps(e2e_steps)
        e2e_s = time.perf_counter() - t0
        h2d = sum(wk.h2d for wk in workers)
        d2h = sum(wk.d2h for wk in workers)
        if os.environ.get("PXZ_BENCH_DEBUG"):
            sys.stderr.write(f"[rank {rank}] e2e {e2e_s * 1e3:.1f} ms for {e2e_steps * batch} images; per-worker busy "
                             f"{[round(wk.busy_s * 1e3, 1) for wk in workers]} ms\n")

    # ---- reduce over ranks: max time -------------------------------------------------------------------
    reason_names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reason_bits = sum(1 << i for i, n in enumerate(reason_names) if n in (clocks.get("reasons") or []))
    times = torch.tensor([elapsed_ms, e2e_s * 1e3, (t_launched - t_wall0) * 1e3, clocks.get("sm_mhz") or -1.0, float(reason_bits)],
                         dtype=torch.float64, device=dev)
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        gathered = [torch.zeros_like(times) for _ in range(world)]
        dist.all_gather(gathered, times)
        per_rank_ms = [float(t[0]) / args.steps for t in gathered]  # diagnostic: which rank sets the max
        # clocks of every rank's GPU: the line reports the slowest one and the union of the throttle reasons
        mhz = [float(t[3]) for t in gathered]
        bits = 0
        for t in gathered:
            bits |= int(t[4])
        clocks["sm_mhz_per_rank"] = mhz
        if all(m > 0 for m in mhz):
            clocks["sm_mhz"] = min(mhz)
        clocks["reasons"] = sorted(set(clocks.get("reasons") or []) | {n for i, n in enumerate(reason_names) if bits >> i & 1})
        head = times[:3].clone()
        dist.all_reduce(head, op=dist.ReduceOp.MAX)
        times[:3] = head
    elapsed_ms, e2e_ms, host_issue_ms = float(times[0]), float(times[1]), float(times[2])
    if world > 1:  # whole-job launch count
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    mp_per_step = world * batch * IMG_W * IMG_H / 1e6
    value = mp_per_step * args.steps / (elapsed_ms / 1e3)
    e2e_value = mp_per_step * e2e_steps / (e2e_ms / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        n_px = IMG_W * IMG_H
        mean_payload = float(np.mean(payload_bytes))
        # algorithmic bytes per launch of every kernel (DESIGN.md "algorithmic bytes")
        algo = {
            "analyze_mad_fast": 4 * n_px + 5 * nblocks,
            "mad_exact": 4 * n_px + 4 * nblocks,
            "plan": 4 * nblocks + 20 * nblocks,
            "resample_down": 4 * n_px + mean_payload + 20 * nblocks,
            "resample_up": mean_payload + 20 * nblocks + 4 * n_px,
        }
        kernels = {}
        for name, (ms, n) in prof.items():
            if n:
                us = ms / n * 1e3
                gbs = algo.get(name, 0) / (us * 1e-6) / 1e9
                kernels[name] = {"us": round(us, 2), "launches": int(n), "algo_GBps": round(gbs, 1),
                                 "frac": round(gbs / peak, 4), "share": round(ms / solo_ms, 4)}
                if n_streams > 1 and name in prof_overlapped and prof_overlapped[name][1]:
                    kernels[name]["us_in_timed_region"] = round(prof_overlapped[name][0] / prof_overlapped[name][1] * 1e3, 2)
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        except OSError:
            pass
        dom = max(kernels, key=lambda k: kernels[k]["us"] * kernels[k]["launches"]) if kernels else None
        roofline = None
        if dom:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["algo_GBps"], "peak": peak,
                        "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic.get(dom),
                        "traffic_source": "profiles/r01_traffic.json (ncu --set full capture of the same workload)" if dom in traffic else None,
                        "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": int(algo[dom]), "avg_launch_us": kernels[dom]["us"],
                        "timing": ("CUDA events around every launch; single-stream pass of the same K steps right after the "
                                   "timed region (in the multi-stream timed region kernels of different images overlap: see "
                                   "kernels[*].us_in_timed_region)") if n_streams > 1 else
                                  "CUDA events around every launch inside the timed region",
                        "single_stream_value_MPps": round(world * batch * IMG_W * IMG_H / 1e6 * args.steps / (solo_ms / 1e3), 1)}
        # whole encode+decode stage against the roofline (image read once + payload written, payload read + image written)
        stage_bytes = (4 * n_px + mean_payload + 16 * nblocks) + (mean_payload + 16 * nblocks + 4 * n_px)
        per_image_s = elapsed_ms / 1e3 / (args.steps * batch)
        stage = {"algorithmic_bytes_per_image": int(stage_bytes), "GBps": round(stage_bytes / per_image_s / 1e9, 1),
                 "frac_of_hbm_peak": round(stage_bytes / per_image_s / 1e9 / peak, 4),
                 "payload_fraction": round(mean_payload / (4 * n_px), 4)}
        cpu = cpu_baseline(np.ascontiguousarray(np_imgs[0][:CPU_SAMPLE_ROWS])) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, batch),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d // e2e_steps),
                    "d2h_bytes_per_step": int(d2h // e2e_steps), "steps": e2e_steps,
                    "host_threads": n_workers,
                    "path": "pxz_image_upload -> pxz_shrink -> pxz_payload_download -> pxz_payload_upload -> pxz_expand, pinned host buffers"},
            "gpu_launches": int(launches),
            # host wall time to issue one step's launches (max over ranks): the step is launch-bound when this nears ms_per_step
            "host_issue_ms_per_step": round(host_issue_ms / args.steps, 4),
            "ms_per_step_per_rank": [round(x, 4) for x in per_rank_ms],
            "fused_resample_mode": fused_info,
            "roofline": roofline,
            "kernels": kernels,
            "encode_decode_stage": stage,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="8K images per rank per step")
    ap.add_argument("--streams", type=int, default=4, help="contexts / CUDA streams of the device-resident leg")
    ap.add_argument("--e2e-workers", type=int, default=4, help="host threads (contexts) of the end-to-end leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun when started plainly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
