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


@pytest.mark.parametrize("nt", [1, 63, 64, 65, 128, 129, 640])
def test_match_all_extreme_rows(nt):
    """All-zero and all-one descriptors on both sides: a'.b reaches -256 and +256, the ends of the 16-bit key range the
    tensor-core kernel packs (distance codes 0 and 512), on every slab / tile boundary of the train set."""
    rng = np.random.default_rng(nt)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t[::5] = 0
    t[1::7] = 255
    q[::3] = 0
    q[1::4] = 255
    t[nt - 1] = 255 if nt % 2 else 0
    m = ORBmatcher()
    got = m.match_all(q, t)
    want = oracle.match_all(q, t)
    for g, w in zip(got, want):
        assert (g == w).all()
    m.close()


@pytest.mark.parametrize("env", [
    {"ORB_B200_MMA_BK": "0"},                                  # k_match_mma3 without the bias K-step
    {"ORB_B200_MMA_KIND": "f8"},                               # e4m3 / f32 twin
    {"ORB_B200_MMA_VARIANT": "30"},                            # k_match_mma2, 8 epilogue warps
    {"ORB_B200_MMA_VARIANT": "20"},                            # k_match_mma2, 16 epilogue warps
    {"ORB_B200_MMA_VARIANT": "10"},                            # k_match_mma, the first form
    {"ORB_B200_MMA_VARIANT": "10", "ORB_B200_MMA_KIND": "f8"},
    {"ORB_B200_MATCH": "popc"},                                # the POPC kernel
])
def test_comparator_kernels_stay_exact(env, monkeypatch):
    """The kernels DESIGN.md keeps as comparators (and the POPC kernel) give the oracle's result on a ragged batch."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(21)
    P, Q, T = 4, 520, 777
    nq = np.array([520, 257, 0, 31], np.int32)
    nt = np.array([777, 129, 40, 0], np.int32)
    t = rand_desc(rng, P * T, 182).reshape(P, T, 32)
    q = rand_desc(rng, P * Q, 182, flip_from=t.reshape(-1, 32)).reshape(P, Q, 32)
    t[1, :64] = rng.integers(0, 256, (64, 32), dtype=np.uint8)  # 256 live bits in one train tile: the 8-K-step path
    m = ORBmatcher()
    bi, bd, sd = m.match_all_batch(q, nq, t, nt)
    for p in range(P):
        oi, od, os_ = oracle.match_all(q[p, :nq[p]], t[p, :nt[p]])
        assert (bi[p, :nq[p]] == oi).all() and (bd[p, :nq[p]] == od).all() and (sd[p, :nq[p]] == os_).all(), (env, p)
    m.close()


def test_match_all_batch_many_items_per_cta(monkeypatch):
    """More (pair, 256-query) items than SMs, ragged and empty pairs in between: the persistent kernel's running barrier
    phases, A-buffer hand-over and skipped items, against the popcount kernel on every pair and the oracle on a few."""
    rng = np.random.default_rng(11)
    P, Q, T = 230, 600, 700
    nq = rng.integers(0, Q + 1, P).astype(np.int32)
    nt = rng.integers(0, T + 1, P).astype(np.int32)
    nq[[3, 77]] = 0
    nt[[5, 78, 200]] = 0
    nq[[9, 100]] = Q
    nt[[9, 101]] = T
    nt[[10, 150]] = [1, 129]
    t = rand_desc(rng, T, 182)[rng.integers(0, T, (P, T))]
    q = rand_desc(rng, Q, 182)[rng.integers(0, Q, (P, Q))]
    q[:, ::3] = t[:, :Q:3]  # exact hits and, through the repeated rows, ties
    m = ORBmatcher()
    monkeypatch.setenv("ORB_B200_MATCH", "mma")
    got = m.match_all_batch(q, nq, t, nt)
    monkeypatch.setenv("ORB_B200_MATCH", "popc")
    want = m.match_all_batch(q, nq, t, nt)
    for p in range(P):
        for g, w in zip(got, want):
            assert (g[p, :nq[p]] == w[p, :nq[p]]).all(), p
    for p in (0, 9, 10, 78, 150, 229):
        o = oracle.match_all(q[p, :nq[p]], t[p, :nt[p]])
        for g, w in zip(got, o):
            assert (g[p, :nq[p]] == w).all(), p
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


@pytest.mark.parametrize("rows,cols,nf,npairs", [(480, 752, 1200, 3), (376, 1241, 2000, 2)])
def test_compute_stereo_matches_batch(rows, cols, nf, npairs):
    # Frame::ComputeStereoMatches for all stereo pairs of one extractor batch (four launches over all pairs, median cut on
    # the device) against the oracle pair by pair: host buffers and the device-resident form
    import torch
    from orb_slam_system_b200 import KP_DTYPE
    bf, fx = 47.90639384423901, 435.2046959714599
    frames = np.stack([oracle.synth_frame(rows, cols, frame=20 + f // 2, right=f & 1) for f in range(2 * npairs)])
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=2 * npairs)
    cap = ex.keypoint_bound(rows, cols)
    kps = np.zeros((2 * npairs, cap), KP_DTYPE)
    desc = np.zeros((2 * npairs, cap, 32), np.uint8)
    counts = np.zeros(2 * npairs, np.int32)
    ex.extract_batch_pinned(frames, kps, desc, counts, cap)
    m = ORBmatcher()
    ur, dep, status = m.ComputeStereoMatchesBatch(ex, kps, desc, counts, cap, bf, fx)
    assert (status == 0).all()
    want = []
    for p in range(npairs):
        cl, cr = counts[2 * p], counts[2 * p + 1]
        our, odep = oracle.compute_stereo_matches(frames[2 * p], frames[2 * p + 1], kps[2 * p, :cl], desc[2 * p, :cl], kps[2 * p + 1, :cr],
                                                  desc[2 * p + 1, :cr], bf, fx)
        want.append((our, odep))
        assert ur[p, :cl].tobytes() == our.tobytes() and dep[p, :cl].tobytes() == odep.tobytes(), p
        assert (our >= 0).sum() > 500
    # device resident: the outputs of extract_batch_device go in untouched
    pitch = (cols + 63) // 64 * 64
    d_in = torch.zeros((2 * npairs, rows, pitch), dtype=torch.uint8, device="cuda")
    d_in[:, :, :cols] = torch.from_numpy(frames).cuda()
    d_k = torch.zeros((2 * npairs, cap, 28), dtype=torch.uint8, device="cuda")
    d_d = torch.zeros((2 * npairs, cap, 32), dtype=torch.uint8, device="cuda")
    d_c = torch.zeros((2 * npairs,), dtype=torch.int32, device="cuda")
    d_ur = torch.full((npairs, cap), -7.0, dtype=torch.float32, device="cuda")
    d_dep = torch.full((npairs, cap), -7.0, dtype=torch.float32, device="cuda")
    d_st = torch.full((npairs,), 9, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_batch_device(d_in[:, :, :cols], d_k, d_d, d_c, cap)
    m.ComputeStereoMatchesBatchDevice(ex, d_k, d_d, d_c, cap, bf, fx, d_ur, d_dep, d_st)
    m.sync()
    assert (d_st.cpu().numpy() == 0).all()
    for p in range(npairs):
        cl = counts[2 * p]
        assert d_ur[p, :cl].cpu().numpy().tobytes() == want[p][0].tobytes() and d_dep[p, :cl].cpu().numpy().tobytes() == want[p][1].tobytes(), p
    # a pair on which the reference faults (a right keypoint whose row band leaves the image) is refused, the others are not
    bad = kps.copy()
    bad[1, 0]["y"] = 0.5
    ur2, dep2, st2 = m.ComputeStereoMatchesBatch(ex, bad, desc, counts, cap, bf, fx)
    assert st2[0] != 0 and (st2[1:] == 0).all()
    assert (ur2[0, :counts[0]] == -1).all()
    assert ur2[1, :counts[2]].tobytes() == want[1][0].tobytes()
    ex.close()
    m.close()
