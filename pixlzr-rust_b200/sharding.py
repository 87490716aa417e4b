"""Multi-GPU host logic (one process per GPU).  The hot path has no cross-block dependency
(SURVEY 8e): Sobel windows and resize taps never leave a block, so

  * a batch of images is dealt round-robin to the ranks with no communication at all, and
  * one large image is cut into contiguous runs of block rows; each rank shrinks its own rows into a
    shard-relative payload and the host concatenates the shards in rank order (container lines are
    per block row, so a shard never splits a line).

The only collective is the optional global normalisation (PXZ_FLAG_NORMALISE_GLOBAL): one 4-float
ncclMin all-reduce of {min, -max} per metric component, issued by the library on the context's stream
(pxz_comm_init).  torch.distributed is used for plumbing only: sharing the 128-byte NCCL id and
gathering shard payloads.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import _native as N


def partition_block_rows(rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of `rows` block rows over `world` ranks: [(first_row, count)]."""
    if world <= 0:
        raise ValueError("world must be positive")
    base, extra = divmod(rows, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, n))
        start += n
    return out


def shard_pixel_rows(height: int, block_height: int, world: int, rank: int) -> Tuple[int, int]:
    """Pixel-row range [y0, y1) of `rank`'s shard (empty when there are more ranks than block rows)."""
    rows = -(-height // block_height)
    first, count = partition_block_rows(rows, world)[rank]
    y0 = min(height, first * block_height)
    y1 = min(height, (first + count) * block_height)
    return y0, y1


def round_robin(n_items: int, world: int, rank: int) -> List[int]:
    """Image i of a batch goes to GPU i mod world."""
    return list(range(rank, n_items, world))


def merge_shards(parts: Sequence[Tuple[np.ndarray, np.ndarray]]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenates per-rank (descs, pixels) of block-row shards, rebasing the payload offsets."""
    descs_out, pixels_out, base = [], [], 0
    for descs, pixels in parts:
        d = np.array(descs, dtype=N.DESC_DTYPE, copy=True)
        d["offset"] += np.uint64(base)
        descs_out.append(d)
        pixels_out.append(np.asarray(pixels, np.uint8))
        base += int(np.asarray(pixels).size)
    return (np.concatenate(descs_out) if descs_out else np.zeros(0, N.DESC_DTYPE),
            np.concatenate(pixels_out) if pixels_out else np.zeros(0, np.uint8))


def merge_shard_containers(files: Sequence[bytes], width: int, height: int) -> bytes:
    """One .pxlzr file from the files of the block-row shards, in rank order.  Each rank can write its shard's file on
    its own GPU (Payload.to_container): a file is header | one length per block row | the rows' blocks, and a shard is a
    run of whole block rows, so the image's file is the image header, the length tables one after the other and the
    block bytes one after the other.  Empty shards (more ranks than block rows) are passed as b""."""
    tables, bodies, first = [], [], None
    rows_total, prev_partial = 0, False
    for f in files:
        if not f:
            continue
        if len(f) < 26 or f[:6] != b"PIXLZR" or f[6:9] != b"\x00\x00\x02":
            raise ValueError("not a version 0.0.2 .pxlzr shard")
        w, h, bw, bh = (int.from_bytes(f[10 + 4 * i:14 + 4 * i], "big") for i in range(4))
        if first is None:
            first = (f[9], bw, bh)
        if w != width or (f[9], bw, bh) != first:
            raise ValueError("shards disagree on width, block size or filter")
        rows = -(-h // bh)
        if prev_partial:
            raise ValueError("only the last shard may end in a partial block row")
        prev_partial = h % bh != 0
        tables.append(f[26:26 + 4 * rows])
        bodies.append(f[26 + 4 * rows:])
        rows_total += rows
    if first is None:
        raise ValueError("no shard holds any block row")
    if rows_total != -(-height // first[2]):
        raise ValueError("the shards do not cover the image's block rows")
    head = b"PIXLZR\x00\x00\x02" + bytes([first[0]]) + b"".join(int(v).to_bytes(4, "big") for v in (width, height, first[1], first[2]))
    return head + b"".join(tables) + b"".join(bodies)


def share_comm_id(dist, make_id, rank: int, device=None) -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id (make_id()), everyone receives it through a
    torch.distributed broadcast (any backend)."""
    import torch

    buf = torch.zeros(N.COMM_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        assert len(raw) == N.COMM_ID_BYTES
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(ctx: "N.Context", dist, rank: int, world: int, device=None) -> None:
    """Creates the library's NCCL communicator across all ranks of the torch.distributed group."""
    raw = share_comm_id(dist, N.comm_unique_id, rank, device)
    ctx.comm_init(world, rank, raw)


def shrink_sharded(ctx: "N.Context", image_rows: np.ndarray, bw: int, bh: int, metric: int, factor: float,
                   filter_down: int, flags: int = 0):
    """Shrinks this rank's rows (host array); returns (descs, pixels) with shard-relative offsets."""
    img = ctx.image_upload(np.ascontiguousarray(image_rows))
    try:
        pl = img.shrink(bw, bh, metric, factor, int(filter_down), flags)
        try:
            return pl.download()
        finally:
            pl.free()
    finally:
        img.free()
