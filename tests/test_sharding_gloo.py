"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: block-row partitioning, shard merging,
sharing of the communicator id, and that a block-row-sharded run — including the global min/max
exchange of the normalise extension — reproduces the single-process result bit for bit.  The per-rank
compute is played by the oracle here (no GPU); on the GPU box tests/test_gpu_parity.py covers the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_partition_and_round_robin():
    import pixlzr_b200 as P

    S = P.sharding
    assert S.partition_block_rows(1024, 8) == [(i * 128, 128) for i in range(8)]
    parts = S.partition_block_rows(68, 8)
    assert sum(n for _, n in parts) == 68 and parts[0] == (0, 9) and parts[-1] == (60, 8)
    assert [s for s, _ in parts] == [0, 9, 18, 27, 36, 44, 52, 60]
    assert S.partition_block_rows(3, 8)[3:] == [(3, 0)] * 5
    assert S.shard_pixel_rows(4320, 64, 8, 7) == (3840, 4320)  # last shard keeps the 32-px trailing row
    assert S.shard_pixel_rows(100, 64, 4, 3) == (100, 100)      # more ranks than block rows: empty shard
    assert S.round_robin(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((S.round_robin(4096, 8, r) for r in range(8)), [])) == list(range(4096))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import oracle as O
    import pixlzr_b200 as P

    S = P.sharding
    # 1. the 128-byte communicator id reaches every rank unchanged
    raw = S.share_comm_id(dist, lambda: bytes(range(128)), rank)
    assert raw == bytes(range(128))

    # 2. block-row sharded shrink == single-process shrink (plain and with global normalisation)
    rng = np.random.default_rng(42)
    h, w, bs = 330, 200, 32
    amp = np.kron(rng.choice([0, 2, 8, 32, 128], size=((h + 31) // 32, (w + 31) // 32)), np.ones((32, 32)))[:h, :w]
    img = np.clip(128 + (rng.random((h, w, 3)) - 0.5) * amp[..., None], 0, 255).astype(np.uint8)
    y0, y1 = S.shard_pixel_rows(h, bs, world, rank)
    mine = np.ascontiguousarray(img[y0:y1])

    for normalise in (False, True):
        for metric, factor in ((O.METRIC_OKLAB_MAD, 0.7 if not normalise else 0.05), (O.METRIC_SOBEL_DIR, 6.0 if not normalise else 1.0)):
            if normalise:
                # what the library does on the GPU: local {min, -max} then a MIN all-reduce (4 floats)
                vx, vy = O.analyze(mine, bs, bs, metric)
                mm = torch.tensor([np.nanmin(vx), -np.nanmax(vx), np.nanmin(vy), -np.nanmax(vy)], dtype=torch.float32)
                dist.all_reduce(mm, op=dist.ReduceOp.MIN)
                # apply the global range on this shard exactly as the oracle defines the extension
                def norm(v, mn, mx):
                    rg = np.float32(mx) - np.float32(mn)
                    return ((v - np.float32(mn)) / rg).astype(np.float32) if rg > 0 else np.zeros_like(v)
                nx = norm(vx, mm[0].item(), -mm[1].item())
                ny = norm(vy, mm[2].item(), -mm[3].item()) if metric == O.METRIC_SOBEL_DIR else nx
                whole = O.shrink(img, bs, bs, metric, factor, O.LANCZOS3, normalise_global=True)
                cols = O.grid(w, h, bs, bs)[0]
                first = y0 // bs
                for i in range(len(vx)):
                    v0 = nx[i] * np.float32(factor) * np.float32(10) if metric == O.METRIC_OKLAB_MAD else nx[i] * np.float32(factor)
                    v1 = v0 if metric == O.METRIC_OKLAB_MAD else ny[i] * np.float32(factor)
                    tw = min(bs, w - (i % cols) * bs)
                    th = min(bs, (y1 - y0) - (i // cols) * bs)
                    ow, oh, st = O.reduce_dims(float(v0), float(v1), tw, th)
                    g = whole.descs[first * cols + i]
                    assert (ow, oh) == (int(g["w"]), int(g["h"])), (rank, i)
                    assert np.float32(st) == g["value"]
                continue
            local = O.shrink(mine, bs, bs, metric, factor, O.LANCZOS3)
            gathered = [None] * world
            dist.all_gather_object(gathered, (local.descs, local.payload))
            descs, pixels = S.merge_shards(gathered)
            whole = O.shrink(img, bs, bs, metric, factor, O.LANCZOS3)
            assert np.array_equal(descs, whole.descs.astype(descs.dtype)) and np.array_equal(pixels, whole.payload)
            # and the merged payload goes through the product's host container stage unchanged
            a = P.native.container_encode(w, h, bs, bs, 4, 3, descs, pixels, None, nthreads=2)
            assert a == O.container_encode(whole, 4)
            # each rank writes its own shard's file (on its GPU in production: Payload.to_container); stitched = the whole file
            mine_file = O.container_encode(local, 4) if y1 > y0 else b""
            files = [None] * world
            dist.all_gather_object(files, mine_file)
            assert S.merge_shard_containers(files, w, h) == a

    # 2b. the interleaved layout (block row g -> rank g mod world) stitches to the same frame and the same file
    rows = -(-h // bs)
    idx = S.cyclic_block_rows(rows, world, rank)
    local = O.shrink(S.gather_block_rows(img, bs, idx), bs, bs, O.METRIC_OKLAB_MAD, 0.7, O.LANCZOS3)
    gathered = [None] * world
    dist.all_gather_object(gathered, (local.descs, local.payload, O.container_encode(local, 4)))
    whole = O.shrink(img, bs, bs, O.METRIC_OKLAB_MAD, 0.7, O.LANCZOS3)
    descs, pixels = S.merge_shards_cyclic([(d, p) for d, p, _ in gathered], O.grid(w, h, bs, bs)[0])
    assert np.array_equal(descs, whole.descs.astype(descs.dtype)) and np.array_equal(pixels, whole.payload)
    assert S.merge_shard_containers_cyclic([f for _, _, f in gathered], w, h) == O.container_encode(whole, 4)

    # 3. batch round-robin: every image is processed by exactly one rank
    mine_idx = S.round_robin(7, world, rank)
    allidx = [None] * world
    dist.all_gather_object(allidx, mine_idx)
    assert sorted(sum(allidx, [])) == list(range(7))
    dist.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_equals_single_process_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
