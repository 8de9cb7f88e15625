#!/usr/bin/env python3
"""Regenerate ookiedokie_b200/data/{filters,devices}/*.json from the reference tree.

The filter and device JSON files are *input formats* of the receive path
(SURVEY.md section 1, L0): FIR tap values and device timing tables, not code.
This script re-emits them in a normalised form (json.dumps, 1-space indent)
so the product tree carries the same numbers without carrying the reference's
files.  Python's float repr round-trips every double exactly, so the taps that
reach `(float) json_number_value(tap)` (reference src/fir.c:224) are identical.

fs64_fs8.json does not exist in the reference (only filters/fs64_fs8.mat,
32 float64 taps == stage 2 of fs128_fs16_dec4.json); it is authored here as a
single stage without a "decimation" key (=> 1, reference src/fir.c:139-157).

Run in the build container only (needs /root/reference); outputs are committed.
"""
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ookiedokie_b200", "data")


def emit(path, obj):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(obj, f, indent=1)
        f.write("\n")
    print("wrote", os.path.relpath(path))


def main():
    for name in ("fs32_fs4", "fs128_fs16_dec4"):
        emit(f"{OUT}/filters/{name}.json", json.load(open(f"{REF}/filters/{name}.json")))
    for name in ("p3l-nexa2012", "unknown-remote1"):
        emit(f"{OUT}/devices/{name}.json", json.load(open(f"{REF}/devices/{name}.json")))
    for name in ("unity1", "unity16"):
        emit(f"{OUT}/filters/{name}.json", json.load(open(f"{REF}/src/test/filters/{name}.json")))

    import scipy.io
    taps = scipy.io.loadmat(f"{REF}/filters/fs64_fs8.mat")["fs64_fs8"].ravel()
    emit(f"{OUT}/filters/fs64_fs8.json", {"filter": {
        "comment": "Authored from filters/fs64_fs8.mat: pass band Fs/64, stop band Fs/8, no decimation",
        "stages": [{"taps": [float(t) for t in taps]}]}})


if __name__ == "__main__":
    main()
