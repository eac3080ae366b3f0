"""Image ingest fused into the level-0 load: raw frame -> cv::remap -> cv::cvtColor gray -> extractor,
through the C ABI, against the oracle's models (pinned to cv2 4.13 in tests/test_ingest_oracle.py) followed by the
oracle extractor.  Reference: src/Tracking.cc:118-126, Examples/Stereo/stereo_euroc.cc:97-98, :136-137."""
import os

import numpy as np
import pytest

import oracle
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200._lib import ORB_ERR_INVALID, OrbError
from test_gpu_extract import compare

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_ingest_vectors.npz"))


def radial_maps(rows, cols, k1=-0.12, shift=(3.25, -2.5), dst=None):
    """Smooth undistortion-like float32 maps (inputs of the test, not a model of initUndistortRectifyMap)."""
    dr, dc = dst or (rows, cols)
    xs, ys = np.meshgrid(np.arange(dc, dtype=np.float64), np.arange(dr, dtype=np.float64))
    xn, yn = (xs - dc / 2) / (0.6 * cols), (ys - dr / 2) / (0.6 * cols)
    f = 1 + k1 * (xn * xn + yn * yn)
    return ((xn * f * 0.6 * cols + cols / 2 + shift[0]).astype(np.float32),
            (yn * f * 0.6 * cols + rows / 2 + shift[1]).astype(np.float32))


def color_frames(n, rows, cols, ch, first=0):
    """Synthetic textured colour frames: three differently seeded planes of the benchmark generator."""
    out = np.zeros((n, rows, cols, ch), np.uint8)
    for f in range(n):
        for c in range(ch):
            out[f, :, :, c] = oracle.synth_frame(rows, cols, frame=first + 7 * f + c, seed=7 + c)
    return out


def check_frames(ex, results, grays, nf, tag):
    for f, ((kg, dg), gray) in enumerate(zip(results, grays)):
        ko, do = oracle.extract(gray, nfeatures=nf, cap=16 * nf)
        info = compare(kg, dg, ko, do, tag=f"{tag}[{f}]")
        assert info["desc_bit_mismatch"] == 0 and info["n"] > 0
        assert np.array_equal(ex.pyramid_level(0, frame=f), gray), f"{tag}[{f}]: level 0 is not the reference's mImGray"


@pytest.mark.parametrize("ch,bgr,variant", [(3, False, 4), (3, True, 4), (4, False, 4), (4, True, 3), (3, False, 3)])
def test_gray_conversion_then_extract(ch, bgr, variant):
    rows, cols, nf, n = 240, 320, 500, 3
    raw = color_frames(n, rows, cols, ch)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=n)
    ex.set_ingest((rows, cols, ch), bgr=bgr, gray_variant=variant)
    res = ex.ingest_extract_batch(raw)
    check_frames(ex, res, [oracle.cvt_gray(raw[f], bgr=bgr, variant=variant) for f in range(n)], nf, f"gray{ch}")
    ex.close()


def test_rectify_gray_euroc_shape():
    rows, cols, nf = 480, 752, 1200   # BASELINE config 2 shape, the stereo_euroc.cc path (gray camera + remap)
    raw = np.stack([oracle.synth_frame(rows, cols, frame=f, right=f & 1) for f in range(2)])
    mx, my = radial_maps(rows, cols)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=2)
    ex.set_ingest((rows, cols), maps=(mx, my))
    res = ex.ingest_extract_batch(raw)
    check_frames(ex, res, [oracle.remap_linear(raw[f], mx, my) for f in range(2)], nf, "rectify")
    ex.close()


def test_rectify_colour_changes_size_and_strides():
    srows, scols, ch, nf, n = 200, 280, 3, 400, 4
    drows, dcols = 176, 250   # maps smaller than the raw frame, dcols % 4 != 0
    raw_padded = np.zeros((n, srows + 3, scols + 5, ch), np.uint8)
    raw_padded[:, :srows, :scols] = color_frames(n, srows, scols, ch, first=40)
    raw = raw_padded[:, :srows, :scols]   # padded rows and frames
    mx, my = radial_maps(srows, scols, k1=0.2, shift=(-6.5, 9.125), dst=(drows, dcols))
    mx[:3] -= 40.0                        # some taps outside the source: BORDER_CONSTANT 0
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=n)
    ex.set_ingest((srows, scols, ch), maps=(mx, my), bgr=True)
    from orb_slam_system_b200._lib import check, lib, ptr, KP_DTYPE
    cap = ex.keypoint_bound(drows, dcols)
    kps, desc, counts = np.zeros((n, cap), KP_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32)
    check(lib().orb_ingest_extract_batch(ex._h, n, ptr(raw_padded), raw.strides[1], raw.strides[0], ptr(kps), ptr(desc), cap, ptr(counts)))
    res = [(kps[f, :counts[f]], desc[f, :counts[f]]) for f in range(n)]
    grays = [oracle.cvt_gray(oracle.remap_linear(np.ascontiguousarray(raw[f]), mx, my), bgr=True) for f in range(n)]
    check_frames(ex, res, grays, nf, "rectify-colour")
    ex.close()


def test_level0_equals_cv2_vectors():
    """GPU ingest against cv2 4.13 outputs directly (committed vectors): remap of a colour frame, then gray."""
    raw = G["raw_color"][None]
    ex = ORBextractor(300, 1.2, 2, 20, 7)   # two levels: the vectors are small images
    ex.set_ingest(raw.shape[1:], maps=(G["map_x"], G["map_y"]))
    ex.ingest_extract_batch(raw)
    assert np.array_equal(ex.pyramid_level(0), oracle.cvt_gray(G["rect_color"]))
    ex.set_ingest(G["raw_gray"].shape, maps=(G["adv_map_x"], G["adv_map_y"]))
    ex.ingest_extract_batch(G["raw_gray"][None])
    assert np.array_equal(ex.pyramid_level(0), G["adv_gray"])
    ex.set_ingest(G["color"].shape, bgr=True)
    ex.ingest_extract_batch(G["color"][None])
    assert np.array_equal(ex.pyramid_level(0), G["gray_bgr"])
    ex.close()


def test_async_and_device_forms_agree():
    import torch
    rows, cols, ch, nf, n = 240, 320, 3, 500, 32
    raw = color_frames(n, rows, cols, ch, first=100)
    mx, my = radial_maps(rows, cols, k1=-0.05)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=n)
    ex.set_ingest((rows, cols, ch), maps=(mx, my))
    base = ex.ingest_extract_batch(raw)
    cap = ex.keypoint_bound(rows, cols)
    # two batches in flight on pinned buffers
    pin = [dict(raw=torch.from_numpy(raw).pin_memory(), kps=torch.zeros((n, cap, 28), dtype=torch.uint8).pin_memory(),
                desc=torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory(), counts=torch.zeros(n, dtype=torch.int32).pin_memory())
           for _ in range(2)]
    tickets = [ex.submit_ingest_pinned(p["raw"], p["kps"], p["desc"], p["counts"], cap) for p in pin]
    for t in tickets:
        ex.wait_batch(t)
    for p in pin:
        for f in range(n):
            c = int(p["counts"][f])
            assert c == len(base[f][0])
            assert np.array_equal(p["kps"][f, :c].numpy().view(base[f][0].dtype).ravel(), base[f][0])
            assert np.array_equal(p["desc"][f, :c].numpy(), base[f][1])
    # device-resident raw frames
    d_raw = torch.from_numpy(raw).cuda()
    d_kps = torch.zeros((n, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.ingest_extract_batch_device(d_raw, d_kps, d_desc, d_counts, cap)
    ex.sync()
    cnt = d_counts.cpu().numpy()
    for f in range(n):
        assert cnt[f] == len(base[f][0])
        assert np.array_equal(d_kps[f, :cnt[f]].cpu().numpy().view(base[f][0].dtype).ravel(), base[f][0])
        assert np.array_equal(d_desc[f, :cnt[f]].cpu().numpy(), base[f][1])
    ex.close()


def test_ingest_errors():
    from orb_slam_system_b200._lib import lib, ptr, KP_DTYPE
    ex = ORBextractor(300, 1.2, 8, 20, 7)
    raw = np.zeros((1, 100, 120, 3), np.uint8)
    kps, desc, counts = np.zeros((1, 64), KP_DTYPE), np.zeros((1, 64, 32), np.uint8), np.zeros(1, np.int32)

    def call():
        return lib().orb_ingest_extract_batch(ex._h, 1, ptr(raw), 360, 36000, ptr(kps), ptr(desc), 64, ptr(counts))

    assert call() == ORB_ERR_INVALID           # no configuration yet
    with pytest.raises(OrbError) as e:
        ex.set_ingest((100, 100, 2))           # channels must be 1, 3 or 4
    assert e.value.code == ORB_ERR_INVALID
    ex.set_ingest((100, 120, 3))
    assert lib().orb_ingest_extract_batch(ex._h, 1, ptr(raw), 359, 36000, ptr(kps), ptr(desc), 64, ptr(counts)) == ORB_ERR_INVALID
    assert call() == 0 and counts[0] == 0      # a black frame: no keypoints, like the reference
    ex.set_ingest(None)
    assert call() == ORB_ERR_INVALID           # cleared again
    ex.close()
