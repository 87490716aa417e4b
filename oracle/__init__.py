"""ctypes binding of the CPU oracle (oracle/pxz_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product path (``pixlzr-rust_b200/``) never
does; it fails loudly when its CUDA library is missing instead of falling back to this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpxz_oracle.so")

NEAREST, TRIANGLE, CATMULLROM, GAUSSIAN, LANCZOS3 = range(5)
METRIC_OKLAB_MAD, METRIC_SOBEL_DIR = 0, 1
FILTER_NAMES = ["Nearest", "Triangle", "CatmullRom", "Gaussian", "Lanczos3"]


class BlockDesc(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("value", C.c_float), ("w", C.c_uint16), ("h", C.c_uint16)]


DESC_DTYPE = np.dtype([("offset", "<u8"), ("value", "<f4"), ("w", "<u2"), ("h", "<u2")])
assert DESC_DTYPE.itemsize == C.sizeof(BlockDesc) == 16


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pxz_oracle.cpp")
    hdr = os.path.join(_HERE, "pxz_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, f32p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.pxo_grid.argtypes = [C.c_uint32] * 4 + [u32p, u32p]
        L.pxo_srgb_lut.argtypes = [f32p]
        L.pxo_cbrtf.argtypes = [C.c_float]
        L.pxo_cbrtf.restype = C.c_float
        L.pxo_oklab.argtypes = [C.c_uint8] * 3 + [f32p]
        L.pxo_block_mad.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int]
        L.pxo_block_mad.restype = C.c_float
        L.pxo_block_sobel.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, f32p, f32p]
        L.pxo_parse_value.argtypes = [C.c_float]
        L.pxo_parse_value.restype = C.c_float
        L.pxo_level_exp.argtypes = [C.c_float]
        L.pxo_level_exp.restype = C.c_int32
        L.pxo_reduce_dims.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32, u32p, u32p, f32p]
        L.pxo_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.pxo_resize_fir.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.pxo_set_resize_semantics.argtypes = [C.c_int]
        L.pxo_set_resize_semantics.restype = None
        L.pxo_axis_weights.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.pxo_analyze.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t, C.c_uint32, C.c_uint32,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.pxo_shrink.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t, C.c_uint32, C.c_uint32,
                                 C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.pxo_shrink.restype = C.c_int64
        L.pxo_expand.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                 C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.pxo_strategy_bucket.argtypes = [C.c_float]
        L.pxo_strategy_bucket.restype = C.c_uint32
        L.pxo_shrink_strategy.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t, C.c_uint32, C.c_uint32,
                                          C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.pxo_shrink_strategy.restype = C.c_int64
        L.pxo_expand_strategy.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.pxo_tree_process.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t, C.c_float, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.pxo_qoi_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
        L.pxo_qoi_encode.restype = C.c_int64
        L.pxo_qoi_decode.argtypes = [C.c_void_p, C.c_size_t, u32p, u32p, C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        L.pxo_container_encode.argtypes = [C.c_uint32] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                              C.c_void_p, C.c_size_t]
        L.pxo_container_encode.restype = C.c_int64
        L.pxo_container_decode.argtypes = [C.c_void_p, C.c_size_t, u32p, u32p, u32p, u32p, C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def grid(w, h, bw, bh):
    c, r = C.c_uint32(), C.c_uint32()
    lib().pxo_grid(w, h, bw, bh, C.byref(c), C.byref(r))
    return c.value, r.value


def srgb_lut() -> np.ndarray:
    out = np.zeros(256, np.float32)
    lib().pxo_srgb_lut(out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def cbrtf(x: float) -> float:
    return lib().pxo_cbrtf(x)


def oklab(r, g, b) -> np.ndarray:
    out = np.zeros(3, np.float32)
    lib().pxo_oklab(r, g, b, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def _check_img(img: np.ndarray):
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] in (3, 4)
    assert img.strides[2] == 1 and img.strides[1] == img.shape[2]
    return img.shape[1], img.shape[0], img.shape[2], img.strides[0]


def block_mad(block: np.ndarray) -> float:
    w, h, c, pitch = _check_img(block)
    return lib().pxo_block_mad(_ptr(block), pitch, w, h, c)


def block_sobel(block: np.ndarray):
    w, h, c, pitch = _check_img(block)
    hz, vr = C.c_float(), C.c_float()
    rc = lib().pxo_block_sobel(_ptr(block), pitch, w, h, c, C.byref(hz), C.byref(vr))
    if rc != 0:
        raise ValueError("block thinner than 2 px: the reference panics")
    return hz.value, vr.value


def parse_value(v: float) -> float:
    return lib().pxo_parse_value(v)


def level_exp(v: float) -> int:
    return lib().pxo_level_exp(v)


def reduce_dims(v0, v1, w, h):
    ow, oh, st = C.c_uint32(), C.c_uint32(), C.c_float()
    lib().pxo_reduce_dims(v0, v1, w, h, C.byref(ow), C.byref(oh), C.byref(st))
    return ow.value, oh.value, st.value


IMAGE_RS, FIR = 0, 1


def set_resize_semantics(sem: int) -> None:
    """Which branch of PixlzrBlock::resize shrink / expand / tree_process use: IMAGE_RS (default, pinned by the reference's
    fixtures) or FIR (fast_image_resize, the reference's default cargo feature; parity unpinned).  Process-wide."""
    lib().pxo_set_resize_semantics(int(sem))


class resize_semantics:
    """with O.resize_semantics(O.FIR): ..."""

    def __init__(self, sem: int):
        self.sem = sem

    def __enter__(self):
        set_resize_semantics(self.sem)

    def __exit__(self, *a):
        set_resize_semantics(IMAGE_RS)


def resize_fir(block: np.ndarray, nw: int, nh: int, filt: int) -> np.ndarray:
    block = np.ascontiguousarray(block)
    h, w, c = block.shape
    out = np.zeros((nh, nw, c), np.uint8)
    if lib().pxo_resize_fir(_ptr(block), w, h, c, _ptr(out), nw, nh, filt) != 0:
        raise ValueError("pxo_resize_fir failed")
    return out


def resize(block: np.ndarray, nw: int, nh: int, filt: int) -> np.ndarray:
    block = np.ascontiguousarray(block)
    h, w, c = block.shape
    out = np.zeros((nh, nw, c), np.uint8)
    rc = lib().pxo_resize(_ptr(block), w, h, c, _ptr(out), nw, nh, filt)
    if rc != 0:
        raise ValueError("pxo_resize failed")
    return out


def axis_weights(n: int, nn: int, filt: int, max_taps: int = 0):
    if max_taps == 0:
        max_taps = max(1, n)
    left = np.zeros(nn, np.uint32)
    count = np.zeros(nn, np.uint32)
    w = np.zeros((nn, max_taps), np.float32)
    rc = lib().pxo_axis_weights(n, nn, filt, _ptr(left), _ptr(count), _ptr(w), max_taps)
    if rc < 0:
        raise ValueError("pxo_axis_weights failed")
    return left, count, w[:, :rc].copy()


def analyze(img: np.ndarray, bw: int, bh: int, metric: int, nthreads: int = 1):
    w, h, c, pitch = _check_img(img)
    cols, rows = grid(w, h, bw, bh)
    vx = np.zeros(cols * rows, np.float32)
    vy = np.zeros(cols * rows, np.float32)
    rc = lib().pxo_analyze(_ptr(img), w, h, c, pitch, bw, bh, metric, _ptr(vx), _ptr(vy), nthreads)
    if rc != 0:
        raise ValueError(f"pxo_analyze failed ({rc})")
    return vx, vy


@dataclass
class Shrunk:
    width: int
    height: int
    block_width: int
    block_height: int
    channels: int
    descs: np.ndarray  # DESC_DTYPE, row-major grid
    payload: np.ndarray  # uint8, packed

    def block(self, i: int) -> np.ndarray:
        d = self.descs[i]
        n = int(d["w"]) * int(d["h"]) * self.channels
        o = int(d["offset"])
        return self.payload[o:o + n].reshape(int(d["h"]), int(d["w"]), self.channels)


def shrink(img: np.ndarray, bw: int, bh: int, metric: int, factor: float, filter_down: int,
           use_factor: bool = True, normalise_global: bool = False, nthreads: int = 1) -> Shrunk:
    w, h, c, pitch = _check_img(img)
    cols, rows = grid(w, h, bw, bh)
    descs = np.zeros(cols * rows, DESC_DTYPE)
    payload = np.zeros(w * h * c, np.uint8)
    n = lib().pxo_shrink(_ptr(img), w, h, c, pitch, bw, bh, metric, factor, int(use_factor), filter_down,
                         int(normalise_global), _ptr(descs), _ptr(payload), nthreads)
    if n < 0:
        raise ValueError(f"pxo_shrink failed ({n})")
    return Shrunk(w, h, bw, bh, c, descs, payload[:n].copy())


def expand(s: Shrunk, filter_up: int, nthreads: int = 1) -> np.ndarray:
    out = np.zeros((s.height, s.width, s.channels), np.uint8)
    rc = lib().pxo_expand(_ptr(s.descs), _ptr(s.payload), s.width, s.height, s.block_width, s.block_height,
                          s.channels, filter_up, _ptr(out), out.strides[0], nthreads)
    if rc != 0:
        raise ValueError("pxo_expand failed")
    return out


STRATEGY_BUCKETS = 65


def strategy_bucket(value: float) -> int:
    """EXTENSION: bucket of a block's stored value, floor(64 * value / sqrt(2)) clamped to [0, 64] (strategies.txt's levels)."""
    return int(lib().pxo_strategy_bucket(float(value)))


def strategy_by_level():
    """(down, up) filter ids per bucket as strategies_by_level.txt:1-12 lists them."""
    down = np.full(STRATEGY_BUCKETS, LANCZOS3, np.uint8)
    up = np.full(STRATEGY_BUCKETS, LANCZOS3, np.uint8)
    down[0], up[0] = NEAREST, NEAREST          # v < 0.015625
    down[1], up[1] = TRIANGLE, NEAREST         # [0.015625; 0.03125)
    down[2], up[2] = CATMULLROM, LANCZOS3      # [0.03125; 0.046875)
    down[3], up[3] = LANCZOS3, CATMULLROM      # [0.046875; 0.0625)
    down[45:], up[45:] = NEAREST, NEAREST      # v >= 0.703125
    return down, up


def shrink_strategy(img: np.ndarray, bw: int, bh: int, metric: int, factor: float, down_by_bucket: np.ndarray,
                    use_factor: bool = True, normalise_global: bool = False, nthreads: int = 1) -> Shrunk:
    """shrink() with the down filter of every block taken from `down_by_bucket[strategy_bucket(block value)]`."""
    w, h, c, pitch = _check_img(img)
    cols, rows = grid(w, h, bw, bh)
    descs = np.zeros(cols * rows, DESC_DTYPE)
    payload = np.zeros(w * h * c, np.uint8)
    lut = np.ascontiguousarray(down_by_bucket, np.uint8)
    assert lut.size == STRATEGY_BUCKETS
    n = lib().pxo_shrink_strategy(_ptr(img), w, h, c, pitch, bw, bh, metric, factor, int(use_factor), _ptr(lut),
                                  int(normalise_global), _ptr(descs), _ptr(payload), nthreads)
    if n < 0:
        raise ValueError(f"pxo_shrink_strategy failed ({n})")
    return Shrunk(w, h, bw, bh, c, descs, payload[:n].copy())


def expand_strategy(s: Shrunk, up_by_bucket: np.ndarray, nthreads: int = 1) -> np.ndarray:
    out = np.zeros((s.height, s.width, s.channels), np.uint8)
    lut = np.ascontiguousarray(up_by_bucket, np.uint8)
    assert lut.size == STRATEGY_BUCKETS
    rc = lib().pxo_expand_strategy(_ptr(s.descs), _ptr(s.payload), s.width, s.height, s.block_width, s.block_height,
                                   s.channels, _ptr(lut), _ptr(out), out.strides[0], nthreads)
    if rc != 0:
        raise ValueError("pxo_expand_strategy failed")
    return out


def tree_process(img: np.ndarray, threshold: float, bw: int, bh: int, min_bw: int = 4, min_bh: int = 4,
                 filter_down: int = LANCZOS3, filter_up: int = NEAREST) -> np.ndarray:
    """tree::process_custom (process/tree.rs:23-83); same channel count as the input."""
    w, h, c, pitch = _check_img(img)
    out = np.zeros((h, w, c), np.uint8)
    rc = lib().pxo_tree_process(_ptr(img), w, h, c, pitch, threshold, bw, bh, min_bw, min_bh, filter_down, filter_up,
                                _ptr(out), out.strides[0])
    if rc != 0:
        raise ValueError("pxo_tree_process failed")
    return out


def from_image(img: np.ndarray, bw: int, bh: int) -> Shrunk:
    """Pixlzr::from_image (pixlzr_image.rs:6-22): tiles as full-size blocks, values absent (0)."""
    w, h, c, _ = _check_img(img)
    cols, rows = grid(w, h, bw, bh)
    descs = np.zeros(cols * rows, DESC_DTYPE)
    parts, off = [], 0
    for by in range(rows):
        for bx in range(cols):
            blk = img[by * bh:min(h, (by + 1) * bh), bx * bw:min(w, (bx + 1) * bw)]
            i = by * cols + bx
            descs[i] = (off, 0.0, blk.shape[1], blk.shape[0])
            parts.append(np.ascontiguousarray(blk).reshape(-1))
            off += blk.size
    return Shrunk(w, h, bw, bh, c, descs, np.concatenate(parts))


def qoi_encode(px: np.ndarray) -> bytes:
    px = np.ascontiguousarray(px)
    h, w, c = px.shape
    need = -lib().pxo_qoi_encode(_ptr(px), w, h, c, None, 0)
    out = np.zeros(need, np.uint8)
    n = lib().pxo_qoi_encode(_ptr(px), w, h, c, _ptr(out), need)
    assert n == need
    return out.tobytes()


def qoi_decode(data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, np.uint8)
    w, h, c = C.c_uint32(), C.c_uint32(), C.c_int()
    if lib().pxo_qoi_decode(_ptr(buf), len(data), C.byref(w), C.byref(h), C.byref(c), None, 0) != 0:
        raise ValueError("bad qoi")
    out = np.zeros((h.value, w.value, c.value), np.uint8)
    if lib().pxo_qoi_decode(_ptr(buf), len(data), C.byref(w), C.byref(h), C.byref(c), _ptr(out), out.size) != 0:
        raise ValueError("bad qoi")
    return out


def container_encode(s: Shrunk, filt: int, values_present: bool = True) -> bytes:
    vp = None
    if not values_present:
        vp = np.zeros(len(s.descs), np.uint8)
    need = -lib().pxo_container_encode(s.width, s.height, s.block_width, s.block_height, filt, s.channels,
                                       _ptr(s.descs), _ptr(s.payload), _ptr(vp) if vp is not None else None, None, 0)
    out = np.zeros(need, np.uint8)
    n = lib().pxo_container_encode(s.width, s.height, s.block_width, s.block_height, filt, s.channels,
                                   _ptr(s.descs), _ptr(s.payload), _ptr(vp) if vp is not None else None,
                                   _ptr(out), need)
    assert n == need
    return out.tobytes()


def container_decode(data: bytes):
    buf = np.frombuffer(data, np.uint8)
    w, h, bw, bh = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    filt, ch, nbytes = C.c_int(), C.c_int(), C.c_uint64()
    args = [_ptr(buf), len(data), C.byref(w), C.byref(h), C.byref(bw), C.byref(bh), C.byref(filt), C.byref(ch),
            C.byref(nbytes)]
    rc = lib().pxo_container_decode(*args, None, None)
    if rc != 0:
        raise ValueError(f"bad container ({rc})")
    cols = int(np.ceil(np.float32(w.value) / np.float32(bw.value)))
    rows = int(np.ceil(np.float32(h.value) / np.float32(bh.value)))
    descs = np.zeros(cols * rows, DESC_DTYPE)
    payload = np.zeros(nbytes.value, np.uint8)
    rc = lib().pxo_container_decode(*args, _ptr(descs), _ptr(payload))
    if rc != 0:
        raise ValueError(f"bad container ({rc})")
    return Shrunk(w.value, h.value, bw.value, bh.value, ch.value, descs, payload), filt.value
