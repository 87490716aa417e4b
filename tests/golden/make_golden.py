#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference's committed fixtures.

Run ONCE in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
It copies the reference's *data* fixtures (PNG / .pix files, no source code) and derives
known-answer vectors from them, so that nothing under tests/ needs /root/reference at run
time (it does not exist on the GPU box).  The reference cannot be executed here (no Rust
toolchain), so the golden outputs are the artefacts the reference itself committed:

  Big-Ruscher.png -> Big-Ruscher.pix      produced by the reference with 32x32 blocks,
                                          Oklab-MAD, shrink_by(Lanczos3, k = 0.125)
  Big-Ruscher.pix -> Big-Ruscher.pix.png  produced by the reference with to_image(Nearest)
  benches/base.png -> benches/base.pixlzr produced by from_image(64, 64) + save (no shrink)
  image.png                               input of the identity round-trip tests main.rs:299-356
"""
import hashlib
import json
import os
import shutil
import struct

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "Big-Ruscher.png": "Big-Ruscher.png",
    "Big-Ruscher.pix": "Big-Ruscher.pix",
    "Big-Ruscher.pix.png": "Big-Ruscher.pix.png",
    "image.png": "image.png",
    "base.png": "benches/base.png",
    "base.pixlzr": "benches/base.pixlzr",
}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def parse_pix_header_and_blocks(data):
    """Minimal independent parser of container v0.0.2 (encoding/mod.rs:95-165): returns the
    header and, per block, (value_bits, w, h, channels) WITHOUT decoding QOI."""
    assert data[:6] == b"PIXLZR" and data[6:9] == bytes([0, 0, 2])
    filt = data[9]
    w, h, bw, bh = struct.unpack(">IIII", data[10:26])
    cols = -(-w // bw)
    rows = -(-h // bh)
    p = 26
    lines = struct.unpack(">%dI" % rows, data[p:p + 4 * rows])
    p += 4 * rows
    blocks = []
    for _ in range(cols * rows):
        assert data[p:p + 5] == b"block"
        (vbits,) = struct.unpack(">I", data[p + 5:p + 9])
        (qlen,) = struct.unpack(">I", data[p + 9:p + 13])
        qw, qh = struct.unpack(">II", data[p + 13:p + 21])
        ch = data[p + 21]
        blocks.append((vbits, qw, qh, ch))
        p += 13 + qlen
    assert p == len(data)
    return dict(filter=filt, width=w, height=h, block_width=bw, block_height=bh, cols=cols, rows=rows,
                line_lengths=list(lines)), blocks


def main():
    meta = {"files": {}}
    for dst, src in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        meta["files"][dst] = {"sha256": sha(os.path.join(HERE, dst)), "reference_path": src}

    for name in ("Big-Ruscher.pix", "base.pixlzr"):
        hdr, blocks = parse_pix_header_and_blocks(open(os.path.join(HERE, name), "rb").read())
        arr = np.array(blocks, dtype=np.uint32)
        np.save(os.path.join(HERE, name + ".blocks.npy"), arr)  # columns: value bits, w, h, channels
        hdr["line_lengths"] = hdr["line_lengths"][:8]
        meta[name] = hdr
        meta[name]["values_sha256_be_f32"] = hashlib.sha256(arr[:, 0].astype(">u4").tobytes()).hexdigest()
        sizes, counts = np.unique(arr[:, 1].astype(np.uint64) << 32 | arr[:, 2], return_counts=True)
        meta[name]["size_histogram"] = {"%dx%d" % (int(s) >> 32, int(s) & 0xFFFFFFFF): int(c)
                                        for s, c in zip(sizes, counts)}

    # known-answer vectors of SURVEY 8c (hex floats), kept verbatim for the oracle self-test
    meta["kat"] = {
        "srgb_lut_sha256_le_f32": "37e7a29de559c9a886732b687c9b79ff5e44b13e9ef839714ccd4c5ddb485fb6",
        "oklab": {
            "0,0,0": ["0x0p+0", "0x0p+0", "0x0p+0"],
            "255,255,255": ["0x1p+0", "0x0p+0", "0x1p-24"],
            "255,0,0": ["0x1.41835ep-1", "0x1.cc85p-3", "0x1.01bbb4p-3"],
            "0,255,0": ["0x1.bb9df8p-1", "-0x1.df005cp-3", "0x1.6f9ce8p-3"],
            "0,0,255": ["0x1.cedcaep-2", "-0x1.09e33p-5", "-0x1.3f013cp-2"],
            "0,73,166": ["0x1.b69b4cp-2", "-0x1.07d5cp-5", "-0x1.47047cp-3"],
            "128,128,128": ["0x1.332244p-1", "0x1p-25", "0x1p-25"],
            "12,200,77": ["0x1.73f2f2p-1", "-0x1.6d908cp-3", "0x1.d6255p-4"],
        },
        "sobel_8x8": ["0x1.18e38ep-4", "0x1.1671c8p-3"],
        "sobel_big_ruscher_bs32": {"0": ["0x1.99999ap-14", "0x1.99999ap-14"],
                                   "1000": ["0x1.dddddep-12", "0x1.0eca86p-12"],
                                   "2039": ["0x1.1745d2p-12", "0x1.745d18p-14"]},
        "level_dims_64_56_17": {
            "0": [1, 1, 1], "0.0031": [1, 1, 1], "0.0883": [4, 4, 2], "0.0884": [8, 7, 3],
            "0.1767": [8, 7, 3], "0.1768": [16, 14, 5], "0.3535": [16, 14, 5], "0.3536": [32, 28, 9],
            "0.7071": [32, 28, 9], "0.70711": [64, 56, 17], "1.0": [64, 56, 17], "7.3": [64, 56, 17],
            "-0.25": [64, 56, 17], "-1.5": [1, 1, 1]},
        "parse_shrinking_factor": {"+1": 1.0, "-1": -1.0, "+1/2": 0.5, "-1/2": -0.5, "2": 2.0, "-2": -2.0,
                                   "1/": 1.0, "1/2/": 1.0},
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta["files"], indent=1))


if __name__ == "__main__":
    main()
