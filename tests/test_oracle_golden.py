"""Pins the CPU oracle (oracle/) against every golden artefact the reference ships for the
hot path.  CPU only.  The golden files were copied / derived from the reference's
own committed fixtures by tests/golden/make_golden.py."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

import oracle as O
from conftest import GOLDEN, load_png


def fhex(s):
    return np.float32(float.fromhex(s))


def test_fixture_files_intact(golden_meta):
    for name, info in golden_meta["files"].items():
        data = open(os.path.join(GOLDEN, name), "rb").read()
        assert hashlib.sha256(data).hexdigest() == info["sha256"], name


def test_srgb_lut(golden_meta):
    lut = O.srgb_lut()
    assert hashlib.sha256(lut.astype("<f4").tobytes()).hexdigest() == golden_meta["kat"]["srgb_lut_sha256_le_f32"]
    assert lut[1] == fhex("0x1.3e4568p-12") and lut[128] == fhex("0x1.ba1516p-3") and lut[255] == 1.0


def test_oklab_kat(golden_meta):
    for key, exp in golden_meta["kat"]["oklab"].items():
        rgb = [int(v) for v in key.split(",")]
        got = O.oklab(*rgb)
        for g, e in zip(got, exp):
            assert g == fhex(e), (key, float(g).hex(), e)


def test_own_cbrtf_equals_old_glibc():
    """Only meaningful on glibc < 2.41 (this image: 2.39), where libm's cbrtf IS the algorithm."""
    libm = ctypes.CDLL("libm.so.6")
    libc = ctypes.CDLL("libc.so.6")
    libc.gnu_get_libc_version.restype = ctypes.c_char_p
    ver = tuple(int(x) for x in libc.gnu_get_libc_version().decode().split(".")[:2])
    if ver >= (2, 41):
        pytest.skip("host glibc has the new cbrtf")
    libm.cbrtf.restype = ctypes.c_float
    libm.cbrtf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.random(5000, dtype=np.float32),
                         (np.float32(10) ** rng.uniform(-7, 0, 5000)).astype(np.float32)])
    for x in xs:
        assert O.cbrtf(float(x)) == libm.cbrtf(float(x))


def test_sobel_kat(golden_meta):
    y, x, c = np.meshgrid(np.arange(8), np.arange(8), np.arange(3), indexing="ij")
    blk = ((((y * 8 + x) * 3 + c) * 37 + 11) % 256).astype(np.uint8)
    hz, vr = O.block_sobel(blk)
    assert [np.float32(hz), np.float32(vr)] == [fhex(s) for s in golden_meta["kat"]["sobel_8x8"]]
    big = load_png("Big-Ruscher.png")
    for bi, exp in golden_meta["kat"]["sobel_big_ruscher_bs32"].items():
        by, bx = divmod(int(bi), 60)
        b = big[by * 32:(by + 1) * 32, bx * 32:(bx + 1) * 32]
        hz, vr = O.block_sobel(b)
        assert [np.float32(hz), np.float32(vr)] == [fhex(s) for s in exp], bi


def test_level_dims_kat(golden_meta):
    for v, exp in golden_meta["kat"]["level_dims_64_56_17"].items():
        got = [O.reduce_dims(float(v), float(v), n, n)[0] for n in (64, 56, 17)]
        assert got == exp, v
        got_h = [O.reduce_dims(float(v), float(v), n, n)[1] for n in (64, 56, 17)]
        assert got_h == exp, v


def test_parse_value_edges():
    assert O.parse_value(0.5) == 0.5
    assert O.parse_value(-0.25) == 0.75
    assert O.parse_value(-1.5) == 0.0
    assert O.parse_value(-1.0) == 0.0
    assert O.parse_value(-0.0) == 1.0  # 1 + (-0) = 1
    assert np.isnan(O.parse_value(float("nan")))
    # NaN (positive) -> level 1 -> full size; +inf -> full size; 0 -> 1 px
    assert O.reduce_dims(float("nan"), float("nan"), 64, 64)[:2] == (64, 64)
    assert O.reduce_dims(float("inf"), 0.0, 64, 64)[:2] == (64, 1)


def test_big_ruscher_pix_values_dims_pixels(golden_meta):
    """Big-Ruscher.png --shrink_by(Lanczos3, 0.125), bs 32--> Big-Ruscher.pix (made by the reference)."""
    img = load_png("Big-Ruscher.png")
    assert img.shape == (1080, 1920, 3)
    gold = np.load(os.path.join(GOLDEN, "Big-Ruscher.pix.blocks.npy"))
    mine = O.shrink(img, 32, 32, O.METRIC_OKLAB_MAD, 0.125, O.LANCZOS3)
    assert len(mine.descs) == 2040 == len(gold)
    assert np.array_equal(mine.descs["value"].view("<u4"), gold[:, 0]), "stored f32 values must be bit-exact"
    assert np.array_equal(mine.descs["w"], gold[:, 1]) and np.array_equal(mine.descs["h"], gold[:, 2])
    ref, filt = O.container_decode(open(os.path.join(GOLDEN, "Big-Ruscher.pix"), "rb").read())
    assert filt == 0 and ref.channels == 3
    assert np.array_equal(mine.payload, ref.payload), "down-sampled pixels must be bit-exact"
    # and therefore the whole file
    assert O.container_encode(mine, 0) == open(os.path.join(GOLDEN, "Big-Ruscher.pix"), "rb").read()


def test_big_ruscher_expand_nearest():
    ref, _ = O.container_decode(open(os.path.join(GOLDEN, "Big-Ruscher.pix"), "rb").read())
    out = O.expand(ref, O.NEAREST)
    assert np.array_equal(out, load_png("Big-Ruscher.pix.png"))


def test_base_pixlzr_tiling_and_container():
    img = load_png("base.png")
    assert img.shape == (1617, 1080, 4) and img[..., 3].min() == 255
    raw = open(os.path.join(GOLDEN, "base.pixlzr"), "rb").read()
    ref, filt = O.container_decode(raw)
    assert (ref.width, ref.height, ref.block_width, ref.block_height, ref.channels) == (1080, 1617, 64, 64, 4)
    tiles = O.from_image(img, 64, 64)
    assert O.grid(1080, 1617, 64, 64) == (17, 26)
    assert np.array_equal(tiles.descs["w"], ref.descs["w"]) and np.array_equal(tiles.descs["h"], ref.descs["h"])
    assert np.array_equal(tiles.payload, ref.payload)
    assert O.container_encode(tiles, 0, values_present=False) == raw


@pytest.mark.parametrize("bs", [8, 64])
def test_identity_roundtrip_image_png(bs):
    """main.rs:299-356: from_image -> (file) -> to_image(Nearest) without shrinking is the identity."""
    img = load_png("image.png")
    assert img.shape == (1170, 1920, 3)
    tiles = O.from_image(img, bs, bs)
    dec, _ = O.container_decode(O.container_encode(tiles, 0, values_present=False))
    assert np.array_equal(O.expand(dec, O.NEAREST), img)


def test_resize_constant_lanczos3():
    """block.rs:401-435"""
    for v in (0, 255):
        r = O.resize(np.full((100, 100, 3), v, np.uint8), 10, 10, O.LANCZOS3)
        assert r.shape == (10, 10, 3) and (r == v).all()


@pytest.mark.parametrize("filt", range(5))
def test_resize_weights_normalised(filt):
    for n, nn in [(64, 8), (64, 1), (17, 3), (8, 64), (3, 56), (64, 64)]:
        left, count, w = O.axis_weights(n, nn, filt)
        assert ((left + count) <= n).all() and (count >= 1).all()
        assert np.allclose(w.sum(axis=1), 1.0, atol=1e-5)


def test_qoi_roundtrip_random():
    rng = np.random.default_rng(3)
    for c in (3, 4):
        for shape in [(1, 1), (5, 7), (64, 64)]:
            img = rng.integers(0, 4, size=shape + (c,), dtype=np.uint8) * 60
            assert np.array_equal(O.qoi_decode(O.qoi_encode(img)), img)


def test_oracle_strategy_reduces_to_plain_filters_and_follows_buckets():
    """EXTENSION (per-block filter pairs): a uniform table is the ordinary path, and a two-filter table equals the
    block-by-block composition of the two plain results, split by pxo_strategy_bucket of the stored value."""
    import numpy as np
    import oracle as O
    rng = np.random.default_rng(12)
    h, w, bs = 96, 160, 32
    amp = np.kron(128.0 * 2.0 ** (-8.0 * rng.random((h // bs, w // bs))), np.ones((bs, bs)))[..., None]
    img = np.clip(128 + (rng.random((h, w, 3)) - 0.5) * 2 * amp, 0, 255).astype(np.uint8)
    plain = {f: O.shrink(img, bs, bs, O.METRIC_OKLAB_MAD, 0.2, f) for f in (O.NEAREST, O.LANCZOS3)}
    uni = O.shrink_strategy(img, bs, bs, O.METRIC_OKLAB_MAD, 0.2, np.full(65, O.LANCZOS3, np.uint8))
    assert np.array_equal(uni.payload, plain[O.LANCZOS3].payload) and np.array_equal(uni.descs, plain[O.LANCZOS3].descs)
    assert np.array_equal(O.expand_strategy(uni, np.full(65, O.TRIANGLE, np.uint8)), O.expand(uni, O.TRIANGLE))
    # buckets below 20 -> Nearest, the others -> Lanczos3
    table = np.where(np.arange(65) < 20, O.NEAREST, O.LANCZOS3).astype(np.uint8)
    mixed = O.shrink_strategy(img, bs, bs, O.METRIC_OKLAB_MAD, 0.2, table)
    buckets = np.array([O.strategy_bucket(v) for v in mixed.descs["value"]])
    assert (buckets < 20).any() and (buckets >= 20).any()
    for i, b in enumerate(buckets):
        src = plain[O.NEAREST] if b < 20 else plain[O.LANCZOS3]
        assert np.array_equal(mixed.block(i), src.block(i)), i
    up = O.expand_strategy(mixed, table)
    full = {f: O.expand(mixed, f) for f in (O.NEAREST, O.LANCZOS3)}
    cols = w // bs
    for i, b in enumerate(buckets):
        y, x = (i // cols) * bs, (i % cols) * bs
        assert np.array_equal(up[y:y + bs, x:x + bs], full[O.NEAREST if b < 20 else O.LANCZOS3][y:y + bs, x:x + bs]), i


def test_fir_branch_constant_images_stay_constant():
    """block.rs:401-435 (test_resize): 100x100 RGB all-0 and all-255 -> 10x10 Lanczos3 stays constant — the only
    constraint the reference puts on its default (fast_image_resize) branch; the oracle's restatement of that branch
    is otherwise PARITY UNPINNED."""
    for v in (0, 255):
        for c in (3, 4):
            blk = np.full((100, 100, c), v, np.uint8)
            for f in (O.NEAREST, O.TRIANGLE, O.CATMULLROM, O.GAUSSIAN, O.LANCZOS3):
                assert (O.resize_fir(blk, 10, 10, f) == v).all(), (v, c, f)
                assert (O.resize_fir(blk[:10, :10], 64, 37, f) == v).all(), (v, c, f)


def test_fir_branch_properties():
    rng = np.random.default_rng(3)
    blk = rng.integers(0, 256, (64, 48, 3), dtype=np.uint8)
    # same size: clone (block.rs:279-281); nearest picks source pixels; the drivers follow the process-wide switch
    assert np.array_equal(O.resize_fir(blk, 48, 64, O.LANCZOS3), blk)
    near = O.resize_fir(blk, 24, 32, O.NEAREST)
    assert np.array_equal(near, blk[1::2, 1::2])
    # opaque RGBA == RGB with alpha 255 appended (pre-multiplication by 255 is the identity)
    rgba = np.concatenate([blk, np.full((64, 48, 1), 255, np.uint8)], -1)
    for f in (O.TRIANGLE, O.CATMULLROM, O.GAUSSIAN, O.LANCZOS3):
        a, b = O.resize_fir(blk, 20, 30, f), O.resize_fir(rgba, 20, 30, f)
        assert np.array_equal(a, b[..., :3]) and (b[..., 3] == 255).all()
    # same kernels as the image-crate branch for CatmullRom / Gaussian / Lanczos3, but a rounded and clipped u8
    # intermediate: close, not equal — SURVEY.md 8c measured up to 9 LSB between the two semantics on hard-edged content
    for f in (O.CATMULLROM, O.GAUSSIAN, O.LANCZOS3):
        dlt = np.abs(O.resize_fir(blk, 24, 32, f).astype(int) - O.resize(blk, 24, 32, f).astype(int))
        assert 0 < dlt.max() <= 12 and dlt.mean() < 1.0
    img = np.ascontiguousarray(np.concatenate([rng.integers(0, 256, (96, 128, 3), dtype=np.uint8), np.full((96, 128, 1), 255, np.uint8)], -1))
    ref = O.shrink(img, 32, 32, O.METRIC_OKLAB_MAD, 0.02, O.LANCZOS3)
    with O.resize_semantics(O.FIR):
        fir = O.shrink(img, 32, 32, O.METRIC_OKLAB_MAD, 0.02, O.LANCZOS3)
    assert np.array_equal(fir.descs["w"], ref.descs["w"]) and np.array_equal(fir.descs["value"], ref.descs["value"])  # the metric does not depend on it
    assert not np.array_equal(fir.payload, ref.payload)
    assert np.array_equal(O.shrink(img, 32, 32, O.METRIC_OKLAB_MAD, 0.02, O.LANCZOS3).payload, ref.payload)  # the switch is back
