#!/usr/bin/env python3
"""Generates the committed golden fixtures (run in the BUILD container only).

Two sources, neither of which exists on the GPU box, hence the committed outputs:
  * cv2 4.13.0 (Python wheel in this container): the OpenCV primitives the reference's
    extractor calls -- cv::resize, cv::GaussianBlur, cv::FAST, cv::fastAtan2 -- on seeded
    synthetic inputs  ->  cv2_vectors.npz
  * the reference's own src/ORBextractor.cc compiled unmodified against oracle/cvshim
    (oracle/_ref/libref_orb.so, needs /root/reference): full operator() outputs on the
    BASELINE shapes  ->  reference_extract.json (digests, counts, leading records)

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cv2  # noqa: E402

import oracle  # noqa: E402
from orb_slam_system_b200.synth import synth_frame  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cv2_vectors():
    out = {}
    # resize chains of the three BASELINE shapes: digest of every level
    digests = {}
    for (h, w) in [(376, 1241), (480, 640), (480, 752)]:
        img = synth_frame(h, w, frame=11)
        inv = oracle.tables(1000)["inv_scale"]
        cur = img
        for l in range(2, 8):
            dw = int(np.rint(np.float32(w) * inv[l]))
            dh = int(np.rint(np.float32(h) * inv[l]))
            cur = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
            digests[f"resize_{h}x{w}_L{l}"] = sha(cur)
        digests[f"blur_{h}x{w}"] = sha(cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    # one small resize + blur kept in full
    small = synth_frame(61, 83, frame=12)
    out["small_img"] = small
    out["small_resize_51x69"] = cv2.resize(small, (69, 51), interpolation=cv2.INTER_LINEAR)
    out["small_blur"] = cv2.GaussianBlur(small, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    # FAST on 24 cells cut from a KITTI-shape frame, thresholds 20 and 7
    img = synth_frame(376, 1241, frame=13)
    rng = np.random.default_rng(13)
    rects, kp20, kp7 = [], [], []
    for i in range(24):
        h, w = int(rng.integers(7, 48)), int(rng.integers(7, 48))
        y0, x0 = int(rng.integers(0, 376 - h)), int(rng.integers(0, 1241 - w))
        rects.append((x0, y0, w, h))
        cell = np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])
        for t, dst in ((20, kp20), (7, kp7)):
            fd = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            for p in fd.detect(cell):
                dst.append((i, int(p.pt[0]), int(p.pt[1]), int(p.response)))
    out["fast_rects"] = np.array(rects, np.int32)
    out["fast_kp20"] = np.array(kp20, np.int32).reshape(-1, 4)
    out["fast_kp7"] = np.array(kp7, np.int32).reshape(-1, 4)
    # fastAtan2
    xs = (rng.normal(size=2000) * 3000).astype(np.float32)
    ys = (rng.normal(size=2000) * 3000).astype(np.float32)
    xs[:8] = [0, 1, -1, 0, 5, -5, 7, 0]
    ys[:8] = [0, 0, 0, 1, 5, 5, -7, -1]
    out["atan2_xy"] = np.stack([xs, ys], 1)
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for x, y in zip(xs, ys)], np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_vectors.npz"), **out)
    return digests


CASES = [
    # name, rows, cols, nfeatures, variant, right, frame  (BASELINE configs 1-3 and retry/degenerate shapes)
    ("tum_640x480", 480, 640, 1000, 0, 0, 0),
    ("euroc_left_752x480", 480, 752, 1200, 0, 0, 0),
    ("euroc_right_752x480", 480, 752, 1200, 0, 1, 0),
    ("kitti_left_1241x376", 376, 1241, 2000, 0, 0, 0),
    ("kitti_right_1241x376", 376, 1241, 2000, 0, 1, 0),
    ("kitti_lowcontrast", 376, 1241, 2000, 1, 0, 1),
    ("kitti03_1242x375", 375, 1242, 2000, 0, 0, 2),
    ("small_300x200", 200, 300, 500, 0, 0, 0),
]


def reference_extract():
    assert oracle.ref_lib() is not None, "needs /root/reference to build oracle/_ref"
    res = {}
    for name, rows, cols, nf, variant, right, frame in CASES:
        img = synth_frame(rows, cols, frame=frame, variant=variant, right=right)
        k, d = oracle.ref_extract(img, nfeatures=nf, cap=16 * nf)
        per_level = np.bincount(k["octave"], minlength=8).tolist()
        res[name] = dict(rows=rows, cols=cols, nfeatures=nf, variant=variant, right=right, frame=frame, count=len(k),
                         per_level=per_level, keypoints_sha256=sha(k), descriptors_sha256=sha(d),
                         int_fields_sha256=sha(np.stack([k["x"], k["y"], k["size"], k["response"]], 1)) + ":" + sha(k["octave"]),
                         first_keypoints=[[float(v) for v in (p["x"], p["y"], p["size"], p["angle"], p["response"])] + [int(p["octave"])]
                                          for p in k[:4]],
                         first_descriptor=d[0].tolist() if len(d) else [])
    return res


if __name__ == "__main__":
    digests = cv2_vectors()
    ref = reference_extract()
    json.dump({"cv2_version": cv2.__version__, "cv2_digests": digests, "reference_extract": ref,
               "note": "reference_extract = /root/reference/src/ORBextractor.cc compiled unmodified against oracle/cvshim, -O2 -ffp-contract=off"},
              open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))
