#!/usr/bin/env python3
"""Golden vectors for the image-ingest models (run in the BUILD container only; cv2 4.13.0 is the source).

cv::cvtColor RGB/BGR(A) -> gray (reference src/Tracking.cc:118-126) and cv::remap INTER_LINEAR with the CV_32FC1 maps of
cv::initUndistortRectifyMap (reference Examples/Stereo/stereo_euroc.cc:97-98, :136-137) on seeded inputs
->  cv2_ingest_vectors.npz.   python tests/golden/make_ingest_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def rectify_maps(cols, rows):
    """EuRoC-like left camera (Examples/Stereo/EuRoC.yaml LEFT.K / LEFT.D / LEFT.R / LEFT.P scaled to the test size)."""
    s = cols / 752.0
    K = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * s], [0, 0, 1]])
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
    R = np.array([[0.999966347530033, -0.001422739138722922, 0.008079580483432283],
                  [0.001365741834644127, 0.9999741760894847, 0.007055629199258132],
                  [-0.008089410156878961, -0.007044357138835809, 0.9999424675829176]])
    P = np.array([[435.2046959714599 * s, 0, 367.4517211914062 * s], [0, 435.2046959714599 * s, 252.2008514404297 * s], [0, 0, 1]])
    return cv2.initUndistortRectifyMap(K, D, R, P, (cols, rows), cv2.CV_32F)


def main():
    rng = np.random.default_rng(2024)
    out = {}
    color = rng.integers(0, 256, (72, 100, 3), dtype=np.uint8)
    color[:8] = rng.integers(0, 2, (8, 100, 3), dtype=np.uint8) * 255  # saturated rows
    out["color"] = color
    out["gray_rgb"] = cv2.cvtColor(color, cv2.COLOR_RGB2GRAY)
    out["gray_bgr"] = cv2.cvtColor(color, cv2.COLOR_BGR2GRAY)
    color4 = rng.integers(0, 256, (40, 52, 4), dtype=np.uint8)
    out["color4"] = color4
    out["gray_rgba"] = cv2.cvtColor(color4, cv2.COLOR_RGBA2GRAY)
    out["gray_bgra"] = cv2.cvtColor(color4, cv2.COLOR_BGRA2GRAY)
    # rectification maps on a 188x120 raw frame (EuRoC / 4)
    gray = rng.integers(0, 256, (120, 188), dtype=np.uint8)
    m1, m2 = rectify_maps(188, 120)
    out["raw_gray"] = gray
    out["map_x"], out["map_y"] = m1, m2
    out["rect_gray"] = cv2.remap(gray, m1, m2, cv2.INTER_LINEAR)
    rawc = rng.integers(0, 256, (120, 188, 3), dtype=np.uint8)
    out["raw_color"] = rawc
    out["rect_color"] = cv2.remap(rawc, m1, m2, cv2.INTER_LINEAR)
    # adversarial maps: out of range, exact integers, exact 1/64 fractions (rounding of the 1/32 grid), negative
    mx = (rng.random((64, 96), dtype=np.float32) * 260 - 40).astype(np.float32)
    my = (rng.random((64, 96), dtype=np.float32) * 180 - 30).astype(np.float32)
    mx[:16] = np.round(mx[:16])
    my[:16:2] = np.round(my[:16:2])
    mx[16:32] = np.round(mx[16:32] * 64) / 64
    my[16:32] = np.round(my[16:32] * 64) / 64
    out["adv_map_x"], out["adv_map_y"] = mx, my
    out["adv_gray"] = cv2.remap(gray, mx, my, cv2.INTER_LINEAR)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "cv2_ingest_vectors.npz"), **out)
    print("wrote cv2_ingest_vectors.npz")


if __name__ == "__main__":
    main()
