"""Multi-GPU host logic (one process per GPU).  The hot path has no cross-block dependency
(SURVEY 8e): Sobel windows and resize taps never leave a block, so

  * a batch of images is dealt round-robin to the ranks with no communication at all, and
  * one large image is cut into block-row shards — contiguous runs, or interleaved (block row g to rank
    g mod world: the same level mix on every rank); each rank shrinks its own rows into a shard-relative
    payload and the host puts the shards together again (container lines are per block row, so a shard
    never splits a line).

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


def cyclic_block_rows(rows: int, world: int, rank: int) -> List[int]:
    """Interleaved split: block row g of the image goes to rank g mod world.  The cost of a block row depends on its
    content (the level mix), and content varies slowly down a frame, so contiguous runs can differ by tens of per cent
    (BASELINE config 4 at 8 GPUs: 5.9 ms against 3.9 ms) while every world-th row gives each rank the same mix."""
    if world <= 0:
        raise ValueError("world must be positive")
    return list(range(rank, rows, world))


def gather_block_rows(image: np.ndarray, block_height: int, rows_idx: Sequence[int]) -> np.ndarray:
    """The pixel rows of the given block rows stacked into one array (host images; on a device the rows are generated or
    copied in place): the local image of a rank in the interleaved layout.  Blocks never reach across block rows, so the
    pipeline sees an ordinary image of len(rows_idx) block rows; only the image's last block row may be lower than the
    others, and it is the last row of its owner."""
    h = image.shape[0]
    parts = [image[g * block_height:min(h, (g + 1) * block_height)] for g in rows_idx]
    if not parts:
        return image[:0]
    return np.ascontiguousarray(np.concatenate(parts, axis=0))


def scatter_block_rows(local: np.ndarray, out: np.ndarray, block_height: int, rows_idx: Sequence[int]) -> None:
    """Inverse of gather_block_rows: writes a rank's decoded rows back into the frame."""
    y = 0
    for g in rows_idx:
        n = min(out.shape[0], (g + 1) * block_height) - g * block_height
        out[g * block_height:g * block_height + n] = local[y:y + n]
        y += n


def merge_shards_cyclic(parts: Sequence[Tuple[np.ndarray, np.ndarray]], cols: int) -> Tuple[np.ndarray, np.ndarray]:
    """(descs, pixels) of the whole frame from the per-rank results of the interleaved layout: block row g is local row
    g // world of rank g mod world; its blocks are contiguous in that rank's payload."""
    world = len(parts)
    descs = [np.asarray(d) for d, _ in parts]
    pixels = [np.asarray(p, np.uint8) for _, p in parts]
    local_rows = [len(d) // cols if cols else 0 for d in descs]
    rows = sum(local_rows)
    descs_out, pixels_out, base = [], [], 0
    for g in range(rows):
        r, k = g % world, g // world
        if k >= local_rows[r]:
            raise ValueError("the shards are not an interleaved split of one frame")
        d = np.array(descs[r][k * cols:(k + 1) * cols], dtype=N.DESC_DTYPE, copy=True)
        lo = int(d["offset"][0])
        hi = int(descs[r]["offset"][(k + 1) * cols]) if (k + 1) * cols < len(descs[r]) else int(pixels[r].size)
        d["offset"] = d["offset"] - np.uint64(lo) + np.uint64(base)
        descs_out.append(d)
        pixels_out.append(pixels[r][lo:hi])
        base += hi - lo
    return (np.concatenate(descs_out) if descs_out else np.zeros(0, N.DESC_DTYPE),
            np.concatenate(pixels_out) if pixels_out else np.zeros(0, np.uint8))


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


def merge_shard_containers_cyclic(files: Sequence[bytes], width: int, height: int) -> bytes:
    """The image's .pxlzr file from the shard files of the interleaved layout (files[r] holds block rows r, r + world, ...):
    a file is header | one length per block row | the rows' blocks, so the lines are dealt back in turn."""
    world = len(files)
    tabs, bodies, first, heights = [], [], None, []
    for f in files:
        if not f:
            tabs.append([]); bodies.append([]); heights.append(0)
            continue
        if len(f) < 26 or f[:6] != b"PIXLZR" or f[6:9] != b"\x00\x00\x02":
            raise ValueError("not a version 0.0.2 .pxlzr shard")
        w, h, bw, bh = (int.from_bytes(f[10 + 4 * i:14 + 4 * i], "big") for i in range(4))
        if first is None:
            first = (f[9], bw, bh)
        if w != width or (f[9], bw, bh) != first:
            raise ValueError("shards disagree on width, block size or filter")
        rows = -(-h // bh)
        lens = [int.from_bytes(f[26 + 4 * i:30 + 4 * i], "big") for i in range(rows)]
        off, lines = 26 + 4 * rows, []
        for n in lens:
            lines.append(f[off:off + n])
            off += n
        if off != len(f):
            raise ValueError("shard file length does not match its line table")
        tabs.append(lens); bodies.append(lines); heights.append(h)
    if first is None:
        raise ValueError("no shard holds any block row")
    bh = first[2]
    rows_total = -(-height // bh)
    if sum(len(t) for t in tabs) != rows_total or sum(heights) != height:
        raise ValueError("the shards do not cover the image's block rows")
    table, body = [], []
    for g in range(rows_total):
        r, k = g % world, g // world
        if k >= len(tabs[r]):
            raise ValueError("the shards are not an interleaved split of one frame")
        table.append(tabs[r][k].to_bytes(4, "big"))
        body.append(bodies[r][k])
    head = b"PIXLZR\x00\x00\x02" + bytes([first[0]]) + b"".join(int(v).to_bytes(4, "big") for v in (width, height, first[1], bh))
    return head + b"".join(table) + b"".join(body)


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
