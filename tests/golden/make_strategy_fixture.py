"""Generates tests/golden/strategies.json from the reference's experiment logs (run in the build container, where
/root/reference is mounted; the tests only read the committed JSON).

strategies.txt holds one line per value bucket of width 1/64 that occurred in the author's runs:
    Level(3)\t[0.046875; 0.0625  )\thas best strategy (down Lanczos3 and up CatmullRom)
strategies_by_level.txt summarises them as ranges."""
import json
import os
import re

REF = os.environ.get("PXZ_REFERENCE", "/root/reference")
FILTERS = {"Nearest": 0, "Triangle": 1, "CatmullRom": 2, "Gaussian": 3, "Lanczos3": 4}


def main():
    per_bucket = {}
    pat = re.compile(r"Level\((\d+)\)\s*\[([0-9.]+)\s*;\s*([0-9.]+)\s*\)\s*has best strategy \(down (\w+) and up (\w+)\)")
    with open(os.path.join(REF, "strategies.txt")) as f:
        for line in f:
            m = pat.match(line.strip())
            if m:
                b, lo, hi = int(m.group(1)), float(m.group(2)), float(m.group(3))
                assert lo == b / 64 and hi == (b + 1) / 64, line
                per_bucket[b] = [FILTERS[m.group(4)], FILTERS[m.group(5)]]
    ranges = []
    with open(os.path.join(REF, "strategies_by_level.txt")) as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    for cond, pair in zip(lines[0::2], lines[1::2]):
        m = re.match(r"\(down (\w+), up (\w+)\)", pair)
        nums = [float(x.replace("_", "")) for x in re.findall(r"[0-9][0-9_.]*", cond)]
        if cond.startswith("v <="):
            lo, hi = 0.0, nums[0]
        elif cond.startswith("v >="):
            lo, hi = nums[0], None
        else:
            lo, hi = nums
        ranges.append({"lo": lo, "hi": hi, "down": FILTERS[m.group(1)], "up": FILTERS[m.group(2)]})
    out = {"source": "strategies.txt / strategies_by_level.txt of the reference", "bucket_width": 1 / 64,
           "per_bucket": {str(k): v for k, v in sorted(per_bucket.items())}, "ranges": ranges}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "strategies.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(per_bucket), "buckets,", len(ranges), "ranges")


if __name__ == "__main__":
    main()
