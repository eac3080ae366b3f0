"""Pins the oracle's restatements of Frame::ComputeStereoMatches (src/Frame.cc:446-619) and of Frame::AssignFeaturesToGrid +
Frame::GetFeaturesInArea (:210-225, :307-372) against the reference's OWN src/Frame.cc, compiled unmodified behind stand-in
objects (oracle/Makefile refframe, oracle/mshim/frame_objects.h).  The compiled reference needs /root/reference; digests of its
outputs are committed in tests/golden/reference_frame.json and checked against the oracle on every box and against the CUDA
library on the GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from search_cases import make_frame, projected_queries

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_frame.json")
HAVE_REF = oracle.ref_frame_lib() is not None
ADAPTER_FRAME = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libadapter_frame.so")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_frame.so needs /root/reference to build")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:24]


def case_stereo(impl, rows, cols, nf, bf, fx, frame):
    il = oracle.synth_frame(rows, cols, frame=frame)
    ir = oracle.synth_frame(rows, cols, frame=frame, right=1)
    if impl == "adapter":
        # the reference's own stereo Frame constructor (src/Frame.cc:41-97, compiled unmodified) with Frame::ComputeStereoMatches
        # supplied by orb_slam_system_b200/adapter/Frame_stereo_b200.cc (oracle/Makefile adapterframe)
        import ctypes as C
        L = C.CDLL(ADAPTER_FRAME)
        cap = 16 * nf
        ur = np.zeros(cap, np.float32)
        dep = np.zeros(cap, np.float32)
        kl = np.zeros(cap, oracle.KP_DTYPE)
        dl = np.zeros((cap, 32), np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        n = L.adpf_stereo_frame(p(il), p(ir), rows, cols, nf, C.c_float(1.2), 8, 20, 7, C.c_float(bf), C.c_float(fx), p(ur), p(dep), p(kl), p(dl), cap)
        assert n >= 0, n
        return int((ur[:n] >= 0).sum()), ur[:n].copy(), dep[:n].copy()
    if impl == "gpu":
        from orb_slam_system_b200 import ORBextractor, ORBmatcher
        exl, exr = ORBextractor(nf, 1.2, 8, 20, 7), ORBextractor(nf, 1.2, 8, 20, 7)
        kl, dl = exl(il)
        kr, dr = exr(ir)
        m = ORBmatcher(0.6, True)
        ur, dep = m.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, bf, fx)
        m.close(); exl.close(); exr.close()
        return int((ur >= 0).sum()), ur, dep
    kl, dl = oracle.extract(il, nfeatures=nf, cap=16 * nf)
    kr, dr = oracle.extract(ir, nfeatures=nf, cap=16 * nf)
    fn = oracle.ref_compute_stereo_matches if impl == "reference" else oracle.compute_stereo_matches
    ur, dep = fn(il, ir, kl, dl, kr, dr, bf, fx)
    return int((ur >= 0).sum()), ur, dep


def case_windows(impl, seed, clustered):
    rng = np.random.default_rng(seed)
    rows, cols = 480, 752
    F = make_frame(rng, 1500, cols=cols, rows=rows, clustered=clustered)
    nq = 400
    _, q, u, v, level, _ = projected_queries(rng, F, nq, outside=0.1)
    r = rng.choice([2.5, 4.0, 10.0, 37.5, 100.0], nq).astype(np.float32)
    lo = np.where(rng.random(nq) < 0.3, -1, level - 1).astype(np.int32)
    hi = np.where(rng.random(nq) < 0.3, -1, level + 1).astype(np.int32)
    if impl == "gpu":
        from orb_slam_system_b200 import ORBmatcher
        m = ORBmatcher(0.6, True)
        off, cand, _ = m.window_search(F, q, u, v, r, lo, hi)
        m.close()
    elif impl == "reference":
        off, cand = oracle.ref_frame_features_in_area(F.keys_un, rows, cols, u, v, r, lo, hi)
    else:
        off, cand = oracle.features_in_area(F, u, v, r, lo, hi)
    return int(off[-1]), off.astype(np.int32), cand.astype(np.int32)


CASES = {
    "stereo_euroc": lambda impl: case_stereo(impl, 480, 752, 1200, 47.90639384423901, 435.2046959714599, 3),   # EuRoC.yaml
    "stereo_kitti": lambda impl: case_stereo(impl, 376, 1241, 2000, 386.1448, 718.856, 5),                     # KITTI00-02.yaml
    "windows_uniform": lambda impl: case_windows(impl, 1, False),
    "windows_clustered": lambda impl: case_windows(impl, 2, True),
}


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_compiled_reference(name):
    o, r = CASES[name](None), CASES[name]("reference")
    assert o[0] == r[0] and o[0] > 500, (name, o[0], r[0])
    for a, b in zip(o[1:], r[1:]):
        assert a.tobytes() == b.tobytes(), name   # floats compared bit for bit


@needs_ref
def test_keyframe_windows_equal_compiled_reference():
    """KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:549-588) of the compiled reference against the oracle's walk without level test."""
    for seed, clustered in ((3, False), (4, True)):
        rng = np.random.default_rng(seed)
        rows, cols = 480, 752
        F = make_frame(rng, 1500, cols=cols, rows=rows, clustered=clustered)
        _, _, u, v, _, _ = projected_queries(rng, F, 400, outside=0.1)
        r = rng.choice([2.5, 4.0, 10.0, 37.5, 100.0], 400).astype(np.float32)
        ref = oracle.ref_keyframe_features_in_area(F.keys_un, rows, cols, u, v, r)
        assert ref is not None
        off, cand = oracle.features_in_area(F, u, v, r, None, None)
        assert np.array_equal(off, ref[0]) and np.array_equal(cand, ref[1]) and len(cand) > 3000


def test_oracle_matches_committed_reference_digests():
    gold = json.load(open(GOLDEN))
    for name in sorted(CASES):
        o = CASES[name](None)
        assert gold[name]["count"] == o[0], name
        assert gold[name]["sha256_24"] == digest(*o[1:]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_committed_reference_digests(name):
    gold = json.load(open(GOLDEN))
    g = CASES[name]("gpu")
    assert gold[name]["count"] == g[0], name
    assert gold[name]["sha256_24"] == digest(*g[1:]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["stereo_euroc", "stereo_kitti"])
def test_adapter_frame_stereo_matches_committed_reference_digests(name):
    """Frame::ComputeStereoMatches through the drop-in member function (adapter/Frame_stereo_b200.cc) called by the reference's
    own stereo Frame constructor: mvuRight / mvDepth against the digests of the compiled reference."""
    if not os.path.exists(ADAPTER_FRAME):
        pytest.skip("oracle/_ref/libadapter_frame.so not built (oracle/Makefile adapterframe)")
    from orb_slam_system_b200 import kernel_launch_count
    gold = json.load(open(GOLDEN))
    before = kernel_launch_count()
    a = CASES[name]("adapter")
    assert kernel_launch_count() > before  # the member function ran on the GPU, not the reference's body
    assert gold[name]["count"] == a[0], name
    assert gold[name]["sha256_24"] == digest(*a[1:]), name


if __name__ == "__main__":  # regenerate the golden digests from the compiled reference
    assert HAVE_REF
    out = {}
    for name in sorted(CASES):
        r = CASES[name]("reference")
        out[name] = {"count": r[0], "sha256_24": digest(*r[1:])}
    json.dump({"source": "reference src/Frame.cc compiled unmodified behind oracle/mshim/frame_objects.h (oracle/Makefile refframe)", **out},
              open(GOLDEN, "w"), indent=1)
    print("wrote", GOLDEN)
