"""Parity of the CUDA Hamming searches (through the C ABI) against the CPU oracle: bit-exact
best index / best distance / second distance for brute-force, windowed (CSR) and stereo scans."""
import numpy as np
import pytest

import oracle
from orb_slam_system_b200 import ORBextractor, ORBmatcher

pytestmark = pytest.mark.gpu


def rand_desc(rng, n, live_bits=182, flip_from=None, nflip=20):
    """Descriptors shaped like the fork's: bits >= live_bits are zero."""
    bits = rng.integers(0, 2, size=(n, 256), dtype=np.uint8)
    bits[:, live_bits:] = 0
    if flip_from is not None:
        src = flip_from[rng.integers(0, len(flip_from), size=n)]
        bits = np.unpackbits(src, axis=1, bitorder="little")
        for i in range(n):
            idx = rng.choice(live_bits, size=rng.integers(0, nflip), replace=False)
            bits[i, idx] ^= 1
    return np.packbits(bits, axis=1, bitorder="little")


@pytest.mark.parametrize("nq,nt,live", [(1, 1, 256), (7, 300, 256), (2000, 2000, 182), (513, 1025, 182), (300, 1, 256), (64, 0, 256)])
def test_match_all(nq, nt, live):
    rng = np.random.default_rng(nq * 7 + nt)
    t = rand_desc(rng, nt, live)
    q = rand_desc(rng, nq, live, flip_from=t if nt else None) if nt else rand_desc(rng, nq, live)
    # force ties: duplicate some train rows so "first minimum" and "second == best" paths are hit
    if nt > 10:
        t[nt // 2] = t[3]
        t[nt - 1] = t[3]
    m = ORBmatcher(0.6, True)
    bi, bd, sd = m.match_all(q, t)
    oi, od, os_ = oracle.match_all(q, t)
    assert (bi == oi).all() and (bd == od).all() and (sd == os_).all()
    m.close()


def test_match_all_batch_ragged():
    rng = np.random.default_rng(5)
    P, Q, T = 5, 700, 900
    nq = np.array([700, 1, 0, 333, 256], np.int32)
    nt = np.array([900, 17, 50, 0, 256], np.int32)
    t = rand_desc(rng, P * T, 182).reshape(P, T, 32)
    q = rand_desc(rng, P * Q, 182, flip_from=t.reshape(-1, 32)).reshape(P, Q, 32)
    m = ORBmatcher()
    bi, bd, sd = m.match_all_batch(q, nq, t, nt)
    for p in range(P):
        oi, od, os_ = oracle.match_all(q[p, :nq[p]], t[p, :nt[p]])
        assert (bi[p, :nq[p]] == oi).all() and (bd[p, :nq[p]] == od).all() and (sd[p, :nq[p]] == os_).all()
    m.close()


@pytest.mark.parametrize("tie_last", [False, True])
def test_match_csr(tie_last):
    rng = np.random.default_rng(11 + tie_last)
    nq, nt = 1500, 2200
    t = rand_desc(rng, nt, 182)
    q = rand_desc(rng, nq, 182, flip_from=t, nflip=60)
    t[100:140] = t[100]  # many equal distances -> first/last tie rules differ
    lens = rng.integers(0, 90, size=nq)
    lens[::97] = 0  # empty candidate lists
    offsets = np.zeros(nq + 1, np.int32)
    offsets[1:] = np.cumsum(lens)
    cand = rng.integers(0, nt, size=offsets[-1]).astype(np.int32)
    cand[: 40] = np.arange(100, 140)
    m = ORBmatcher()
    bi, bd, sd = m.match_csr(q, t, offsets, cand, tie_last=tie_last, max_dist=50)
    oi, od, os_ = oracle.match_csr(q, t, offsets, cand, tie_last=tie_last, max_dist=50)
    assert (bi == oi).all() and (bd == od).all() and (sd == os_).all()
    m.close()


def test_descriptor_distance_unit():
    rng = np.random.default_rng(3)
    a, b = rand_desc(rng, 1, 256)[0], rand_desc(rng, 1, 256)[0]
    assert ORBmatcher.DescriptorDistance(a, b) == oracle.distance(a, b) == int(np.unpackbits(a ^ b).sum())
    assert ORBmatcher.DescriptorDistance(a, a) == 0


def test_bruteforce_ratio_config4():
    # BASELINE config 4: 2000 x 2000 rows drawn from extractor output, ratio 0.6
    img_l = oracle.synth_frame(376, 1241, frame=0)
    img_r = oracle.synth_frame(376, 1241, frame=0, right=1)
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    _, dl = ex(img_l)
    _, dr = ex(img_r)
    q, t = dl[:2000], dr[:2000]
    m = ORBmatcher(0.6, True)
    got = m.BruteForceRatio(q, t)
    oi, od, os_ = oracle.match_all(q, t)
    ok = (od <= 100) & (od.astype(np.float32) <= np.float32(0.6) * os_.astype(np.float32))
    want = np.where(ok, oi, -1)
    assert (got == want).all()
    assert (got >= 0).sum() > 50  # the synthetic stereo pair really matches
    ex.close()
    m.close()


@pytest.mark.parametrize("rows,cols,nf", [(480, 752, 1200), (376, 1241, 2000)])
def test_stereo_match(rows, cols, nf):
    # BASELINE config 2: left+right extraction, then the Hamming part of ComputeStereoMatches
    img_l = oracle.synth_frame(rows, cols, frame=1)
    img_r = oracle.synth_frame(rows, cols, frame=1, right=1)
    ex = ORBextractor(nf, 1.2, 8, 20, 7)
    kl, dl = ex(img_l)
    kr, dr = ex(img_r)
    bf, fx = 47.90639384423901, 435.2046959714599  # EuRoC.yaml Camera.bf, Camera.fx
    m = ORBmatcher()
    br, bd = m.stereo_match(kl, dl, kr, dr, ex.GetScaleFactors(), rows, bf, fx)
    orr, od = oracle.stereo_match(kl, dl, kr, dr, ex.GetScaleFactors(), rows, bf, fx)
    assert (br == orr).all() and (bd == od).all()
    assert (br >= 0).sum() > 100
    ex.close()
    m.close()


@pytest.mark.parametrize("rows,cols,nf,frame", [(480, 752, 1200, 1), (376, 1241, 2000, 0), (376, 1241, 2000, 5)])
def test_compute_stereo_matches_whole(rows, cols, nf, frame):
    # Frame::ComputeStereoMatches end to end: Hamming search, SAD refinement on the device pyramids, parabola,
    # disparity gate, median cut -> mvuRight / mvDepth identical to the oracle (float ops replayed in order)
    img_l = oracle.synth_frame(rows, cols, frame=frame)
    img_r = oracle.synth_frame(rows, cols, frame=frame, right=1)
    bf, fx = 47.90639384423901, 435.2046959714599
    exl, exr = ORBextractor(nf, 1.2, 8, 20, 7), ORBextractor(nf, 1.2, 8, 20, 7)  # one instance per camera (Tracking.cc:76-82)
    kl, dl = exl(img_l)
    kr, dr = exr(img_r)
    m = ORBmatcher()
    ur, dep = m.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, bf, fx)
    our, odep = oracle.compute_stereo_matches(img_l, img_r, kl, dl, kr, dr, bf, fx)
    assert ur.tobytes() == our.tobytes() and dep.tobytes() == odep.tobytes()
    assert (ur >= 0).sum() > 500
    # the pair inside one batch of a single extractor (frames 0 and 1)
    exb = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=2)
    res = exb.extract_batch(np.stack([img_l, img_r]))
    (kl2, dl2), (kr2, dr2) = res[0], res[1]
    ur2, dep2 = m.ComputeStereoMatches(exb, exb, kl2, dl2, kr2, dr2, bf, fx, frame_left=0, frame_right=1)
    assert ur2.tobytes() == our.tobytes() and dep2.tobytes() == odep.tobytes()
    for e in (exl, exr, exb):
        e.close()
    m.close()
