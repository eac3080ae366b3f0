"""Pins the oracle's image-ingest models (cv::cvtColor to gray, cv::remap INTER_LINEAR) against cv2 4.13.0:
the committed vectors of tests/golden/make_ingest_golden.py, and cv2 itself where it is importable.
Reference call sites: src/Tracking.cc:118-126, Examples/Stereo/stereo_euroc.cc:97-98, :136-137."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_ingest_vectors.npz"))


def test_gray_matches_cv2_vectors():
    assert np.array_equal(oracle.cvt_gray(G["color"]), G["gray_rgb"])
    assert np.array_equal(oracle.cvt_gray(G["color"], bgr=True), G["gray_bgr"])
    assert np.array_equal(oracle.cvt_gray(G["color4"]), G["gray_rgba"])
    assert np.array_equal(oracle.cvt_gray(G["color4"], bgr=True), G["gray_bgra"])


def test_gray_variant3_is_the_14_bit_form():
    c = G["color"].astype(np.int64)
    want = ((c[..., 0] * 4899 + c[..., 1] * 9617 + c[..., 2] * 1868 + (1 << 13)) >> 14).astype(np.uint8)
    assert np.array_equal(oracle.cvt_gray(G["color"], variant=3), want)
    # the two fixed-point forms agree except for rounding ties
    d = np.abs(oracle.cvt_gray(G["color"]).astype(int) - want.astype(int))
    assert d.max() <= 1


def test_remap_matches_cv2_vectors():
    assert np.array_equal(oracle.remap_linear(G["raw_gray"], G["map_x"], G["map_y"]), G["rect_gray"])
    assert np.array_equal(oracle.remap_linear(G["raw_color"], G["map_x"], G["map_y"]), G["rect_color"])
    assert np.array_equal(oracle.remap_linear(G["raw_gray"], G["adv_map_x"], G["adv_map_y"]), G["adv_gray"])


def test_remap_identity_and_outside():
    img = G["raw_gray"]
    r, c = img.shape
    xs, ys = np.meshgrid(np.arange(c, dtype=np.float32), np.arange(r, dtype=np.float32))
    assert np.array_equal(oracle.remap_linear(img, xs, ys), img)  # weights {32767, 0, 0, 1}: still the identity
    assert not oracle.remap_linear(img, xs + 10 * c, ys).any()    # BORDER_CONSTANT 0


def test_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)
    assert np.array_equal(oracle.cvt_gray(img), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    mx = (rng.random((70, 110), dtype=np.float32) * 170 - 20).astype(np.float32)
    my = (rng.random((70, 110), dtype=np.float32) * 120 - 15).astype(np.float32)
    assert np.array_equal(oracle.remap_linear(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
