"""BASELINE.json's two sharded configurations, measured device-timed on N GPUs of one node (one process per GPU):

  C4: one large synthetic RGBA8 image cut into contiguous block-row shards (generated on the device, shard by shard),
      Oklab-MAD with the global normalisation extension — the library's own 4-float NCCL min all-reduce is the only
      exchange — Lanczos3 down / Lanczos3 up, encode + decode;
  C5: a batch of synthetic 1920x1080 RGBA8 images dealt round-robin to the ranks, no communication.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_configs.py \
        --config c4 [--side 65536]        |      --config c5 [--images 4096]

Prints one JSON line on rank 0.  `checksum` is a sum over all ranks of the descriptor dims and payload bytes: the same
value at every N shows that the sharded result is the single-GPU result (the parity test proper, against the oracle, is
tests/test_multi_gpu.py).  PyTorch is plumbing here: device tensors for the synthetic pixels, torch.distributed for the
rendezvous and the timing barrier."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixlzr_b200 as P  # noqa: E402

N, S = P.native, P.sharding
LANCZOS3 = 4


def synth_rows(y0, y1, width, seed, device, chunk=512):
    """uint8 [y1-y0, width, 4] on the device: smooth base + per-64x64-tile noise whose amplitude is a hash of the tile
    (independent of the sharding), alpha 255."""
    out = torch.empty((y1 - y0, width, 4), dtype=torch.uint8, device=device)
    out[..., 3] = 255
    xx = torch.arange(width, device=device, dtype=torch.float32)
    tx = (torch.arange(width, device=device) // 64).to(torch.int64)
    for g0 in range((y0 // chunk) * chunk, y1, chunk):  # chunks are aligned globally: the pixels do not depend on the sharding
        yy = torch.arange(g0, g0 + chunk, device=device, dtype=torch.float32)[:, None]
        ty = (torch.arange(g0, g0 + chunk, device=device) // 64).to(torch.int64)[:, None]
        h = (ty * 73856093 + tx[None, :] * 19349663 + seed * 83492791) & 0x7FFFFFFF
        amp = torch.tensor([0, 1, 2, 4, 8, 16, 32, 64], device=device, dtype=torch.float32)[(h >> 7) % 8]
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + g0)
        a, b = max(y0, g0), min(y1, g0 + chunk)
        for ch, base in enumerate((128 + 96 * torch.sin(xx[None, :] / 9700.0 + seed), 128 + 96 * torch.cos(yy / 13100.0),
                                   128 + 64 * torch.sin((xx[None, :] + yy) / 6100.0))):
            noise = (torch.rand((chunk, width), device=device, generator=g) - 0.5) * 2 * amp
            full = torch.clamp(torch.round(base + noise), 0, 255).to(torch.uint8)
            out[a - y0:b - y0, :, ch] = full[a - g0:b - g0]
    return out


def timed(fn, reps, device):
    """max over ranks of the mean device time of `fn` (barrier + synchronize on both sides)."""
    fn()
    torch.cuda.synchronize(device)
    dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(device)
    dt = torch.tensor([(time.perf_counter() - t0) / reps], device=device, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


def run_c4(args, rank, world, device):
    side, bs = args.side, 64
    ctx = N.Context(rank)
    S.init_comm(ctx, dist, rank, world, device=device)
    y0, y1 = S.shard_pixel_rows(side, bs, world, rank)
    src = synth_rows(y0, y1, side, 7, device)
    dst = torch.empty_like(src)
    img = ctx.image_wrap(src.data_ptr(), side, y1 - y0, 4, side * 4)
    out = ctx.image_wrap(dst.data_ptr(), side, y1 - y0, 4, side * 4)
    res = {}
    for name, flags in (("normalise_global (reference-order values + NCCL min/max)", N.FLAG_NORMALISE_GLOBAL), ("default", 0)):
        def step():
            pl = img.shrink(bs, bs, N.METRIC_OKLAB_MAD, 1.0, LANCZOS3, flags)
            pl.expand_to_image(LANCZOS3, out)
            ctx.synchronize()
            pl.free()
        dt = timed(step, args.reps, device)
        pl = img.shrink(bs, bs, N.METRIC_OKLAB_MAD, 1.0, LANCZOS3, flags)
        descs, px = pl.download()
        pl.free()
        chk = torch.tensor([int(descs["w"].astype(np.uint64).sum() * 65537 + descs["h"].astype(np.uint64).sum()),
                            int(px.astype(np.uint64).sum()), int(px.size)], device=device, dtype=torch.int64)
        dist.all_reduce(chk)
        res[name] = {"MP/s": round(side * side / dt / 1e6), "ms": round(dt * 1e3, 2), "payload_fraction": round(int(chk[2]) / (side * side * 4), 4),
                     "checksum": [int(chk[0]), int(chk[1])]}
    return {"config": f"C4 synthetic {side}x{side} RGBA8, 64x64 blocks, block-row shards, Oklab-MAD k=1, Lanczos3 / Lanczos3, encode+decode, device-resident",
            "n_gpus": world, "blocks": (side // bs) ** 2, "results": res}


def run_c5(args, rank, world, device):
    w, h, bs = 1920, 1080, 64
    mine = S.round_robin(args.images, world, rank)
    distinct = min(len(mine), 32)
    srcs = [synth_rows(0, h, w, 100 + i, device) for i in mine[:distinct]]
    ctxs = [N.Context(rank) for _ in range(args.streams)]
    imgs = [[c.image_wrap(s.data_ptr(), w, h, 4, w * 4) for s in srcs] for c in ctxs]
    outs_t = [torch.empty((h, w, 4), dtype=torch.uint8, device=device) for _ in ctxs]
    outs = [c.image_wrap(t.data_ptr(), w, h, 4, w * 4) for c, t in zip(ctxs, outs_t)]

    def step():
        pend = [None] * len(ctxs)
        for n, _ in enumerate(mine):
            k = n % len(ctxs)
            if pend[k] is not None:
                pend[k].free()
            pend[k] = imgs[k][n % distinct].shrink(bs, bs, N.METRIC_OKLAB_MAD, 1.0, LANCZOS3, 0)
            pend[k].expand_to_image(LANCZOS3, outs[k])
        for k, c in enumerate(ctxs):
            c.synchronize()
            if pend[k] is not None:
                pend[k].free()
    dt = timed(step, args.reps, device)
    return {"config": f"C5 batch of {args.images} synthetic 1920x1080 RGBA8 images round-robin, 64x64 blocks, Oklab-MAD k=1, Lanczos3 / Lanczos3, "
                      f"encode+decode, device-resident ({distinct} distinct images per rank, {args.streams} streams per GPU)",
            "n_gpus": world, "images_per_s": round(args.images / dt), "MP/s": round(args.images * w * h / dt / 1e6), "ms_per_batch": round(dt * 1e3, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["c4", "c5"], required=True)
    ap.add_argument("--side", type=int, default=32768)
    ap.add_argument("--images", type=int, default=4096)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    res = run_c4(args, rank, world, device) if args.config == "c4" else run_c5(args, rank, world, device)
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
