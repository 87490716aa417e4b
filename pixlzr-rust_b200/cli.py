"""Command-line driver with the reference CLI's arguments and operation selection (src/bin/main.rs:7-265).

    python pixlzr_b200.py -i in.png -o out.pix -b 64 -k 1/2 -f lanczos3 --force

The operation follows from the two file extensions (main.rs:93-114): `.pix` / `.pixlzr` is the container, anything else an
image; an output without extension is a container.  As in the reference the blocks are only shrunk with `--force`
(main.rs:166-172, 204-210, 226-232, 256-262).  The pixel work runs on the GPU through the C ABI; image files are read and
written with Pillow on the host, the container with the library's host stage.
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import List, Optional, Tuple

import numpy as np

from . import api

FILTERS = {"nearest": api.FilterType.Nearest, "triangle": api.FilterType.Triangle, "catmull-rom": api.FilterType.CatmullRom,
           "gaussian": api.FilterType.Gaussian, "lanczos3": api.FilterType.Lanczos3}


def build_parser() -> argparse.ArgumentParser:
    """clap definition of main.rs:7-43 (same flags, defaults and value names)."""
    p = argparse.ArgumentParser(prog="pixlzr", description="Pixlzr - A rust lib and CLI for the pixlzr image format")
    p.add_argument("-i", "--input", required=True, help="The input image file")
    p.add_argument("-o", "--output", required=True, help="The output image file")
    p.add_argument("-b", "--block-width", type=int, default=64, help="The width of each block")
    p.add_argument("--block-height", type=int, default=None, help="The height of each block")
    p.add_argument("-k", "--shrinking-factor", default="1",
                   help="The shrinking factor: [+|-][1/][D][.D]  If negative, is passed through max(0, 1 - x).")
    p.add_argument("-f", "--filter", choices=sorted(FILTERS), default="lanczos3",
                   help="The filter used when resizing the image blocks")
    p.add_argument("-d", "--direction-wise", type=lambda s: {"true": True, "false": False}[s.lower()], default=None,
                   metavar="true|false", help="Direction-wise scan")
    p.add_argument("--force", action="store_true", help="If image-2-image, force shrinking?")
    return p


def parse_args(argv: Optional[List[str]] = None):
    """`-k -1/2`: the reference's clap option has allow_hyphen_values (main.rs:28-33); argparse needs `--opt=value`."""
    argv = list(sys.argv[1:] if argv is None else argv)
    out, i = [], 0
    while i < len(argv):
        if argv[i] in ("-k", "--shrinking-factor") and i + 1 < len(argv):
            out.append("--shrinking-factor=" + argv[i + 1])
            i += 2
        else:
            out.append(argv[i])
            i += 1
    return build_parser().parse_args(out)


def kind_of(path: str, default: str) -> str:
    """main.rs:93-114: "pix" for .pix / .pixlzr (any case), "image" for another extension, `default` without one."""
    ext = os.path.splitext(path)[1]
    if not ext:
        return default
    return "pix" if ext[1:].lower() in ("pix", "pixlzr") else "image"


def operation(input_path: str, output_path: str) -> Tuple[str, str]:
    return kind_of(input_path, "image"), kind_of(output_path, "pix")


def _open_image(path: str) -> np.ndarray:
    from PIL import Image
    im = Image.open(path)
    if im.mode not in ("RGB", "RGBA"):
        im = im.convert("RGBA" if "A" in im.getbands() or "transparency" in im.info else "RGB")
    return np.ascontiguousarray(np.array(im))


def _save_image(path: str, pixels: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(pixels).save(path)


def _maybe_shrink(pix: api.Pixlzr, args, shrink_by: float) -> None:
    if not args.force:
        return
    if args.direction_wise:
        pix.shrink_directionally(FILTERS[args.filter], shrink_by)
    else:
        pix.shrink_by(FILTERS[args.filter], shrink_by)


def run(args) -> None:
    bw = args.block_width
    bh = args.block_height if args.block_height is not None else bw
    filt = FILTERS[args.filter]
    shrink_by = api.parse_shrinking_factor(args.shrinking_factor)
    src, dst = operation(args.input, args.output)
    # the two plain conversions run with the container stage on the device (same bytes / pixels as the general path)
    if src == "image" and dst == "pix" and args.force:
        with open(args.output, "wb") as f:
            f.write(api.Pixlzr.encode_image_to_vec(_open_image(args.input), bw, bh, filt, shrink_by, bool(args.direction_wise)))
        return
    # pix -> image: shrink_by leaves a decoded file alone (every block carries a value, pixlzr.rs:168-170), so --force only
    # matters with -d true: shrink_directionally has no such check (pixlzr.rs:187-205) and re-shrinks the decoded blocks —
    # that case takes the general path below
    if src == "pix" and dst == "image" and not (args.force and args.direction_wise):
        with open(args.input, "rb") as f:
            _save_image(args.output, api.Pixlzr.decode_vec_to_image(f.read(), filt))
        return
    if src == "image":                                   # image_to_pix / image_to_image (main.rs:142-214)
        pix = api.Pixlzr.from_image(_open_image(args.input), bw, bh)
    elif dst == "image":                                 # pix_to_image (main.rs:216-240)
        pix = api.Pixlzr.open(args.input)
    else:                                                # pix_to_pix re-tiles the decoded image (main.rs:242-265)
        pix = api.Pixlzr.from_image(api.Pixlzr.open(args.input).to_image(filt), bw, bh)
    _maybe_shrink(pix, args, shrink_by)
    if dst == "pix":
        pix.save(args.output)
    else:
        _save_image(args.output, pix.to_image(filt))


def main(argv: Optional[List[str]] = None) -> int:
    args = parse_args(argv)
    try:
        run(args)
    except FileNotFoundError as e:
        print(f"Could not open the image [ {e.filename} ]", file=sys.stderr)
        return 1
    return 0
