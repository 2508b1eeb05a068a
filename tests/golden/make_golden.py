#!/usr/bin/env python
"""Regenerates tests/golden/known_answers.json.

Two kinds of entries:
  "survey"  integer known answers recorded in SURVEY.md §4 from an independent behavioural model of
            the reference (coefficient-stream SHA-256, bytes_read, MCU/block counts, spot pixels).
            They are constants typed in below — the oracle must reproduce them, not define them.
  "oracle"  outputs of the oracle itself (oracle/jpeg_oracle.c) on the reference fixtures and on a few
            images of the synthetic corpus, recorded so that any later change of the oracle, the encoder
            or the generator is caught.  The reference is Rust and cannot be run in this image, so no
            entry comes from the reference binary (parity vs the binary is unpinned).
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ffi as O  # noqa: E402
from jpeg_rust_b200 import synth  # noqa: E402

SURVEY = {
    "lena.jpeg": {"layout": "REF", "ext": 0, "mcus": 2048, "nblocks": [4096, 2048, 2048], "bytes_read": 90694,
                  "scan_len": 90696, "coef_sha256": "ba5ce1b7b3b108bb2bf354a742b7c1217138d58cefd78abe347e05255c060c06",
                  "first_block": [87, 3, 4, -3, -1, 2, 0, 1, 0, 1, 0, 0],
                  "pixels": {"0,0": [224, 138, 127], "256,256": [180, 66, 73]}, "means": [179.24, 99.06, 104.90]},
    "2x2-chroma.jpeg": {"layout": "REF", "ext": 0, "mcus": 1763, "nblocks": [7052, 1763, 1763], "bytes_read": 144537,
                        "scan_len": 145021,
                        "coef_sha256": "04cf33d3a2401666bcf9972782886d8e4418ac2eff240308ed2706b83f6379b0",
                        "first_block": [165, -6, -4, -2, 0, -2, 4, 0, 2, -2, -1, 1]},
    "2x2-chroma.jpeg#spec": {"layout": "SPEC", "ext": 0, "mcus": 1786, "nblocks": [7144, 1786, 1786],
                             "bytes_read": 145019,
                             "coef_sha256": "d29cc5cc8e09c5def6a31f652905c99f00e9db477b5c91d1666a01c02fd839a7"},
    "lena-bw.jpeg": {"layout": "REF", "ext": 0, "mcus": 4096, "nblocks": [4096], "bytes_read": 21494, "scan_len": 21496,
                     "coef_sha256": "0fa4cc6820aa8a2d1d2448b5ad5f6b036ff007c17ef8b0369c5207655775dea2",
                     "first_block": [13, 1, 1, 0], "pixels": {"0,0": [157, 157, 157]}, "means": [116.55, 116.55, 116.55]},
    "huff_simple0.jpg": {"layout": "REF", "ext": 1, "mcus": 2, "nblocks": [2, 2, 2], "bytes_read": 8, "scan_len": 10,
                         "coef_sha256": "5b042c5ca7ff9a10af546a630d36b235ee4d23e3bea5d6f0f74197fb770efece"},
}

SYNTH = [  # (index, width, height, subsampling, quality, restart_interval)
    (0, 1920, 1080, "420", 85, 0), (1, 640, 480, "422", 85, 0), (2, 320, 240, "444", 85, 0),
    (3, 200, 100, "gray", 85, 0), (4, 250, 131, "420", 85, 0), (5, 333, 200, "440", 50, 0),
    (6, 640, 480, "444", 85, 7), (7, 250, 131, "420", 95, 3),
]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out = {"survey": SURVEY, "oracle": {"fixtures": {}, "synthetic": []}}
    for name in ("lena.jpeg", "2x2-chroma.jpeg", "lena-bw.jpeg", "huff_simple0.jpg"):
        data = open(os.path.join(HERE, "fixtures", name), "rb").read()
        ext = O.EXT_SKIP_APPN if name == "huff_simple0.jpg" else O.EXT_NONE
        e = {}
        for lname, layout in (("REF", O.LAYOUT_REF), ("SPEC", O.LAYOUT_SPEC)):
            r = O.decode(data, layout=layout, ext=ext)
            e[lname] = {"status": r.status, "rgb_sha256": sha(r.rgb.tobytes()), "coef_sha256": sha(r.coefficient_stream()),
                        "bytes_read": r.bytes_read, "mcus": r.mcus_read}
        out["oracle"]["fixtures"][name] = e
    for idx, w, h, sub, q, ri in SYNTH:
        data, gt = synth.synth_jpeg(idx, w, h, sub, q, ri, want_coefs=True)
        r = O.decode(data, layout=O.LAYOUT_SPEC, ext=O.EXT_DRI if ri else O.EXT_NONE)
        assert r.status == 0
        gt_stream = b"".join(c.astype("<i2").tobytes() for c in gt)
        assert gt_stream == r.coefficient_stream(), "oracle disagrees with the encoder's ground truth"
        out["oracle"]["synthetic"].append({"index": idx, "width": w, "height": h, "subsampling": sub, "quality": q,
                                           "restart_interval": ri, "file_sha256": sha(data), "file_len": len(data),
                                           "coef_sha256": sha(gt_stream), "rgb_sha256_spec": sha(r.rgb.tobytes()),
                                           "bytes_read": r.bytes_read})
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote known_answers.json")


if __name__ == "__main__":
    main()
