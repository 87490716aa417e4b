"""2-GPU test (skipped with fewer devices): one image sharded by block rows across ranks, global min/max of the
normalise extension exchanged by the library's own NCCL all-reduce, merged result bit-identical to the oracle's
single-process result.  Run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import oracle as O
    import pixlzr_b200 as P

    N, S = P.native, P.sharding
    ctx = N.Context(rank)
    S.init_comm(ctx, dist, rank, world, device=torch.device("cuda", rank))

    rng = np.random.default_rng(123)
    h, w, bs = 1000, 1536, 64
    amp = np.kron(rng.choice([0, 2, 8, 32, 128], size=((h + 63) // 64, (w + 63) // 64)), np.ones((64, 64)))[:h, :w]
    rgb = np.clip(128 + (rng.random((h, w, 3)) - 0.5) * amp[..., None], 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], -1))
    y0, y1 = S.shard_pixel_rows(h, bs, world, rank)
    for metric, factor in ((N.METRIC_SOBEL_DIR, 1.0), (N.METRIC_OKLAB_MAD, 0.05)):
        # PXZ_FLAG_EXACT_VALUES: the stored values are compared bit for bit below (without it the Oklab values take the
        # fast path and only dims, offsets and pixels are exact; bench.py's nccl_parity block runs that variant)
        descs, pixels = S.shrink_sharded(ctx, img[y0:y1], bs, bs, metric, factor, O.CATMULLROM, N.FLAG_NORMALISE_GLOBAL | N.FLAG_EXACT_VALUES)
        gathered = [None] * world
        dist.all_gather_object(gathered, (descs, pixels))
        if rank == 0:
            md, mp = S.merge_shards(gathered)
            ref = O.shrink(img, bs, bs, metric, factor, O.CATMULLROM, normalise_global=True, nthreads=8)
            assert np.array_equal(md["w"], ref.descs["w"]) and np.array_equal(md["h"], ref.descs["h"])
            assert np.array_equal(md["value"].view("<u4"), ref.descs["value"].view("<u4"))
            assert np.array_equal(md["offset"], ref.descs["offset"]) and np.array_equal(mp, ref.payload)
    dist.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_block_row_sharding_with_nccl_minmax(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
