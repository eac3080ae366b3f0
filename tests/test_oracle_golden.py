"""CPU tests: the oracle against the committed golden fixtures (cv2 4.13.0 primitives and the
reference's own ORBextractor.cc compiled against oracle/cvshim), and -- when they are available
in this container -- against cv2 and oracle/_ref live."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from orb_slam_system_b200.synth import synth_frame

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
VEC = np.load(os.path.join(HERE, "golden", "cv2_vectors.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_synth_matches_oracle_generator():
    for args in [(376, 1241, 7, 0, 0, 0), (100, 131, 9, 2, 1, 1), (64, 80, 7, 5, 1, 0), (50, 70, 3, 1, 0, 1)]:
        assert (synth_frame(*args) == oracle.synth_frame(*args)).all()


def test_ctor_tables():
    # quotas and scale factors of the fork (D1: level 1 has scale 1.0)
    t = oracle.tables(2000, 1.2, 8)
    assert t["features_per_level"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.tables(1000)["features_per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert oracle.tables(1200)["features_per_level"].tolist() == [261, 217, 181, 151, 126, 105, 87, 72]
    assert t["scale"][0] == 1.0 and t["scale"][1] == 1.0 and abs(t["scale"][2] - 1.2) < 1e-6
    assert abs(t["scale"][7] - 1.2 ** 6) < 1e-5
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_resize_and_blur_digests_cv2():
    inv = oracle.tables(1000)["inv_scale"]
    for (h, w) in [(376, 1241), (480, 640), (480, 752)]:
        img = synth_frame(h, w, frame=11)
        cur = img
        for l in range(2, 8):
            dw = int(np.rint(np.float32(w) * inv[l]))
            dh = int(np.rint(np.float32(h) * inv[l]))
            cur = oracle.resize(cur, dh, dw)
            assert sha(cur) == GOLD["cv2_digests"][f"resize_{h}x{w}_L{l}"], (h, w, l)
        assert sha(oracle.blur7(img)) == GOLD["cv2_digests"][f"blur_{h}x{w}"]
    small = VEC["small_img"]
    assert (oracle.resize(small, 51, 69) == VEC["small_resize_51x69"]).all()
    assert (oracle.blur7(small) == VEC["small_blur"]).all()


def test_fast_cells_cv2():
    img = synth_frame(376, 1241, frame=13)
    for t, key in ((20, "fast_kp20"), (7, "fast_kp7")):
        want = VEC[key]
        for i, (x0, y0, w, h) in enumerate(VEC["fast_rects"]):
            got = oracle.fast(img[y0:y0 + h, x0:x0 + w], t)
            exp = want[want[:, 0] == i][:, 1:]
            assert got.shape == exp.shape and (got == exp).all(), (t, i)


def test_fast_atan2_cv2():
    xy, deg = VEC["atan2_xy"], VEC["atan2_deg"]
    got = np.array([oracle.fast_atan2(float(y), float(x)) for x, y in xy], np.float32)
    assert (got == deg).all()


@pytest.mark.parametrize("name", sorted(GOLD["reference_extract"]))
def test_extract_matches_reference_golden(name):
    g = GOLD["reference_extract"][name]
    img = synth_frame(g["rows"], g["cols"], frame=g["frame"], variant=g["variant"], right=g["right"])
    k, d = oracle.extract(img, nfeatures=g["nfeatures"], cap=16 * g["nfeatures"])
    assert len(k) == g["count"]
    assert np.bincount(k["octave"], minlength=8).tolist() == g["per_level"]
    assert sha(k) == g["keypoints_sha256"]
    assert sha(d) == g["descriptors_sha256"]


def test_known_answer_counts_survey_a8():
    info = {}
    oracle.extract(synth_frame(376, 1241), nfeatures=2000, cap=32000, info=info)
    assert info["candidates"].tolist() == [7111, 7111, 9303, 8657, 6734, 4691, 3124, 1921]
    assert info["kept"].tolist() == [766, 766, 768, 768, 768, 192, 192, 256]
    info = {}
    k, d = oracle.extract(synth_frame(376, 1241, variant=1), nfeatures=2000, cap=32000, info=info)
    assert info["retry_cells"] > 100  # the low-contrast variant exercises the minThFAST retry
    assert (d[:, 23:] == 0).all() and (d[:, 22] & 0xC0 == 0).all()  # bits 182..255 are always 0


def test_live_reference_when_built():
    if not os.path.exists("/root/reference/src/ORBextractor.cc") and oracle.ref_lib() is None:
        pytest.skip("reference sources and oracle/_ref absent")
    for (h, w, nf, var, right, fr) in [(376, 1241, 2000, 0, 1, 5), (480, 640, 1000, 1, 0, 6), (140, 400, 300, 0, 0, 7), (90, 90, 300, 0, 0, 0)]:
        img = synth_frame(h, w, frame=fr, variant=var, right=right)
        k1, d1 = oracle.extract(img, nfeatures=nf, cap=16 * nf)
        k2, d2 = oracle.ref_extract(img, nfeatures=nf, cap=16 * nf)
        assert k1.tobytes() == k2.tobytes() and (d1 == d2).all()


def test_live_cv2_when_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    img = synth_frame(200, 260, frame=21)
    assert (cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101) == oracle.blur7(img)).all()
    for (dh, dw) in [(167, 217), (139, 181), (200, 260), (99, 77)]:
        assert (cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) == oracle.resize(img, dh, dw)).all()
    fd = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for _ in range(10):
        h, w = int(rng.integers(5, 60)), int(rng.integers(5, 60))
        y0, x0 = int(rng.integers(0, 200 - h)), int(rng.integers(0, 260 - w))
        cell = np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])
        a = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in fd.detect(cell)], np.int32).reshape(-1, 3)
        b = oracle.fast(cell, 7)
        assert a.shape == b.shape and (a == b).all()


def test_octree_edge_cases():
    # empty, single key, quota already met by the roots, all-singleton stop, ties on response
    assert len(oracle.octree(np.zeros((0, 3), np.float32), 16, 316, 16, 116, 50)) == 0
    assert oracle.octree(np.array([[10, 10, 5]], np.float32), 16, 316, 16, 116, 50).tolist() == [0]
    pts = np.array([[10, 10, 5], [200, 50, 7], [250, 90, 7], [251, 91, 7]], np.float32)
    out = oracle.octree(pts, 16, 316, 16, 116, 1)
    assert sorted(out.tolist()) == [0, 1] or len(out) >= 1
    out = oracle.octree(pts, 16, 316, 16, 116, 100)
    assert sorted(out.tolist()) == [0, 1, 2, 3]


def test_matcher_scan_semantics():
    # first minimum wins, a duplicate of the minimum counts as second (ORBmatcher.cc:49-55)
    q = np.zeros((1, 32), np.uint8)
    t = np.zeros((4, 32), np.uint8)
    t[0, 0] = 0b111
    t[1, 0] = 0b1
    t[2, 0] = 0b1
    t[3, 0] = 0b11
    bi, bd, sd = oracle.match_all(q, t)
    assert (bi[0], bd[0], sd[0]) == (1, 1, 1)
    bi, bd, sd = oracle.match_all(q, t[:0])
    assert (bi[0], bd[0], sd[0]) == (-1, 2147483647, 2147483647)
    off = np.array([0, 4], np.int32)
    cand = np.array([3, 2, 1, 0], np.int32)
    bi, bd, sd = oracle.match_csr(q, t, off, cand, tie_last=False)
    assert (bi[0], bd[0], sd[0]) == (2, 1, 1)
    bi, bd, sd = oracle.match_csr(q, t, off, cand, tie_last=True, max_dist=50)
    assert (bi[0], bd[0]) == (1, 1)
    a = np.arange(32, dtype=np.uint8)
    b = a[::-1].copy()
    assert oracle.distance(a, b) == int(np.unpackbits(a ^ b).sum())
