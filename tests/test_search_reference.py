"""Pins the oracle's restatements of the ORBmatcher search methods against the reference's OWN src/ORBmatcher.cc, compiled
unmodified against stand-in objects (oracle/Makefile refmatch, oracle/mshim).  The compiled reference needs /root/reference,
so these tests run in the build container; digests of its outputs are committed in tests/golden/reference_match.json and
checked against the oracle everywhere (test_oracle_matches_committed_reference_digests)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from search_cases import SCALE, bow_pair, make_frame, noisy_copy, projected_queries, rand_desc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_match.json")
HAVE_REF = oracle.ref_match_lib() is not None
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_match.so needs /root/reference to build")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:24]


# ---- seeded cases: every call returns the flattened result of one method for impl = None (oracle), "reference" (the
# compiled src/ORBmatcher.cc), "gpu" (the CUDA library through its host mirror) or "adapter" (the drop-in
# ORB_SLAM2::ORBmatcher class of orb_slam_system_b200/adapter, compiled against the same stand-in objects and driven
# through the very bridge that drives the compiled reference: oracle/Makefile adaptermatch) ----
def _gpu(nnratio, check_ori):
    from orb_slam_system_b200 import ORBmatcher
    return ORBmatcher(nnratio, check_ori)


def case_projection_map(impl, seed, th, stereo, nnratio, some_unobserved):
    rng = np.random.default_rng(seed)
    F = make_frame(rng, 1800, stereo=stereo)
    src, q, u, v, level, _ = projected_queries(rng, F, 1200, max_flips=60)
    occ = (rng.random(F.N) < 0.15).astype(np.uint8)
    vc = rng.choice([0.9, 0.9985], 1200).astype(np.float32)
    xr = (u - rng.uniform(0, 30, 1200)).astype(np.float32)
    qo = (rng.random(1200) < 0.8).astype(np.uint8) if some_unobserved else None
    if impl == "gpu":
        m = _gpu(nnratio, True)
        n, fq = m.SearchByProjection(F, occ, q, u, v, xr, level, vc, th, q_observed=qo)
        m.close()
        return n, fq, occ
    n, fq = oracle.search_by_projection_map(F, occ, q, u, v, xr, level, vc, th, nnratio, qo, impl=impl)
    return n, fq, occ


def case_projection_last(impl, seed, mode, check_ori, th):
    rng = np.random.default_rng(seed)
    Cur = make_frame(rng, 2000)
    nq = 1300
    src, q, u, v, level, angle = projected_queries(rng, Cur, nq, max_flips=70, outside=0.0)
    # rot = last.angle - cur.angle must stay >= 0 in the reference (no wrap: a negative bin indexes rotHist out of bounds, D9)
    angle = (Cur.keys_un["angle"][src] % 200 + 100 + rng.uniform(0, 20, nq)).astype(np.float32)
    Cur.keys_un["angle"] = Cur.keys_un["angle"] % 100
    # the reference drops projections outside the image bounds before the search (:763): same filter on the inputs
    keep = (u >= Cur.mnMinX) & (u <= Cur.mnMinX + 64.0 / Cur.mfGridElementWidthInv) & (v >= Cur.mnMinY) & \
           (v <= Cur.mnMinY + 48.0 / Cur.mfGridElementHeightInv)
    q, u, v, level, angle = q[keep], u[keep], v[keep], level[keep], angle[keep]
    claimed = (rng.random(Cur.N) < 0.1).astype(np.uint8)
    if impl == "gpu":
        m = _gpu(0.9, check_ori)
        n, fq = m.SearchByProjectionLast(Cur, claimed, q, u, v, level, angle, th, forward=mode == "forward", backward=mode == "backward")
        m.close()
        return n, fq, claimed
    n, fq = oracle.search_by_projection_last(Cur, claimed, q, u, v, level, angle, th, mode == "forward", mode == "backward", check_ori,
                                             impl=impl)
    return n, fq, claimed


def case_bow(impl, seed, nnratio, check_ori):
    rng = np.random.default_rng(seed)
    F1, F2, fv1, fv2, has1, has2 = bow_pair(rng)
    kf1 = dict(desc=F1.desc, keys=F1.keys_un, has_mp=has1, featvec=fv1)
    kf2 = dict(desc=F2.desc, keys=F2.keys_un, has_mp=has2, featvec=fv2, sigma2=SCALE * SCALE)
    if impl == "gpu":
        from orb_slam_system_b200 import FeatureVector
        m = _gpu(nnratio, check_ori)
        r = m.SearchByBoW(F1.desc, F1.keys_un["angle"], has1, FeatureVector(fv1), F2.desc, F2.keys_un["angle"], has2, FeatureVector(fv2))
        m.close()
        return r
    return oracle.search_by_bow_kf(kf1, kf2, nnratio, check_ori, impl=impl)


def case_triangulation(impl, seed, check_ori):
    rng = np.random.default_rng(seed)
    F1, F2, fv1, fv2, has1, has2 = bow_pair(rng, n_nodes=25)
    has1 = (rng.random(F1.N) < 0.3).astype(np.uint8)   # most features are free to triangulate
    has2 = (rng.random(F2.N) < 0.3).astype(np.uint8)
    kf1 = dict(desc=F1.desc, keys=F1.keys_un, has_mp=has1, featvec=fv1)
    kf2 = dict(desc=F2.desc, keys=F2.keys_un, has_mp=has2, featvec=fv2, sigma2=SCALE * SCALE)
    # an essentially horizontal epipolar geometry: the copies in KF2 sit within a few pixels of their source rows
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32) + rng.normal(0, 1e-5, (3, 3)).astype(np.float32)
    if impl == "gpu":
        from orb_slam_system_b200 import FeatureVector
        m = _gpu(0.6, check_ori)
        r = m.SearchForTriangulation(F1.keys_un, F1.desc, has1, FeatureVector(fv1), F2.keys_un, F2.desc, has2, FeatureVector(fv2), F12,
                                     (SCALE * SCALE).astype(np.float32))
        m.close()
        return r
    return oracle.search_for_triangulation(kf1, kf2, F12, check_ori, impl=impl)


def case_initialization(impl, seed, window, nnratio, check_ori):
    rng = np.random.default_rng(seed)
    F2 = make_frame(rng, 1800)
    src = rng.integers(0, F2.N, 1500)
    k1 = F2.keys_un[src].copy()
    k1["octave"][::2] = 0
    k1["angle"] = np.mod(k1["angle"] + rng.normal(0, 6, len(k1)), 360).astype(np.float32)
    k1["angle"][k1["angle"] >= 360] = 0
    d1 = noisy_copy(rng, F2.desc[src], 35)
    prev = np.stack([k1["x"] + rng.normal(0, 4, len(k1)), k1["y"] + rng.normal(0, 4, len(k1))], 1).astype(np.float32).copy()
    if impl == "gpu":
        m = _gpu(nnratio, check_ori)
        n, m12 = m.SearchForInitialization(k1, d1, F2, prev, window)
        m.close()
        return n, m12, prev
    n, m12 = oracle.search_for_initialization(k1, d1, F2, prev, window, nnratio, check_ori, impl=impl)
    return n, m12, prev


def _inside(F, u, v):
    """The reference drops projections outside the image before it searches (IsInImage / the mnMin..mnMax test)."""
    return (u >= F.mnMinX) & (u < F.mnMinX + 64.0 / F.mfGridElementWidthInv) & (v >= F.mnMinY) & (v < F.mnMinY + 48.0 / F.mfGridElementHeightInv)


def case_reloc(impl, seed, check_ori, th, orb_dist):
    rng = np.random.default_rng(seed)
    Cur = make_frame(rng, 2000)
    src, q, u, v, level, angle = projected_queries(rng, Cur, 1400, max_flips=70, outside=0.0)
    keep = _inside(Cur, u, v)
    q, u, v, level, angle = q[keep], u[keep], v[keep], level[keep], angle[keep]
    claimed = (rng.random(Cur.N) < 0.1).astype(np.uint8)
    if impl == "gpu":
        m = _gpu(0.9, check_ori)
        n, fq = m.SearchByProjectionReloc(Cur, claimed, q, u, v, level, angle, th, orb_dist)
        m.close()
        return n, fq, claimed
    n, fq = oracle.search_by_projection_reloc(Cur, claimed, q, u, v, level, angle, th, orb_dist, check_ori, impl=impl)
    return n, fq, claimed


def case_loop(impl, seed, th):
    rng = np.random.default_rng(seed)
    KF = make_frame(rng, 1800)
    src, q, u, v, level, _ = projected_queries(rng, KF, 1200, max_flips=40, outside=0.0)
    keep = _inside(KF, u, v)
    q, u, v, level = q[keep], u[keep], v[keep], level[keep]
    radius = (np.float32(th) * SCALE[level]).astype(np.float32)
    claimed = (rng.random(KF.N) < 0.15).astype(np.uint8)
    if impl == "gpu":
        m = _gpu(0.75, True)
        n, fq = m.SearchByProjectionKF(KF, claimed, q, u, v, radius)
        m.close()
        return n, fq, claimed
    if impl in ("reference", "adapter"):
        n, fq = oracle.ref_search_by_projection_loop(KF, claimed, q, u, v, radius, impl=impl)
    else:
        n, fq = oracle.search_kf_window(KF, claimed, q, u, v, radius, None, 50)
    return n, fq, claimed


def case_sim3(impl, seed, th):
    rng = np.random.default_rng(seed)
    KF2 = make_frame(rng, 1800)
    src, q, u, v, level, _ = projected_queries(rng, KF2, 1200, max_flips=60, outside=0.0)
    keep = _inside(KF2, u, v)
    q, u, v, level = q[keep], u[keep], v[keep], level[keep]
    radius = (np.float32(th) * SCALE[level]).astype(np.float32)
    if impl == "gpu":
        m = _gpu(0.75, True)
        r = m.SearchBySim3(KF2, q, u, v, radius, level)
        m.close()
        return r
    if impl in ("reference", "adapter"):
        return oracle.ref_search_by_sim3(KF2, q, u, v, level, th, impl=impl)
    return oracle.search_kf_window(KF2, None, q, u, v, radius, level, 100)


def case_fuse(impl, seed, th, variant):
    rng = np.random.default_rng(seed)
    KF = make_frame(rng, 1800)
    src, q, u, v, level, _ = projected_queries(rng, KF, 1200, max_flips=40, outside=0.0)
    keep = _inside(KF, u, v)
    q, u, v, level = q[keep], u[keep], v[keep], level[keep]
    radius = (np.float32(th) * SCALE[level]).astype(np.float32)
    if impl == "gpu":
        m = _gpu(0.6, True)
        r = m.FuseSearch(KF, q, u, v, radius, level)
        m.close()
        return r
    if impl in ("reference", "adapter"):
        return oracle.ref_fuse_search(KF, q, u, v, level, th, variant, impl=impl)
    return oracle.search_kf_window(KF, None, q, u, v, radius, level, 50)


CASES = {
    "projection_map_mono": lambda impl: case_projection_map(impl, 11, 1.0, False, 0.8, False),
    "projection_map_stereo_th3": lambda impl: case_projection_map(impl, 12, 3.0, True, 0.8, True),
    "projection_map_ratio06": lambda impl: case_projection_map(impl, 13, 5.0, True, 0.6, True),
    "projection_last_normal": lambda impl: case_projection_last(impl, 21, "normal", True, 7.0),
    "projection_last_forward": lambda impl: case_projection_last(impl, 22, "forward", True, 15.0),
    "projection_last_backward_noori": lambda impl: case_projection_last(impl, 23, "backward", False, 7.0),
    "reloc_ori": lambda impl: case_reloc(impl, 61, True, 10.0, 100),
    "reloc_noori_tight": lambda impl: case_reloc(impl, 62, False, 3.0, 64),
    "loop_th10": lambda impl: case_loop(impl, 71, 10),
    "sim3_th75": lambda impl: case_sim3(impl, 81, 7.5),
    "fuse_th3": lambda impl: case_fuse(impl, 91, 3.0, 0),
    "fuse_scw_th4": lambda impl: case_fuse(impl, 92, 4.0, 1),
    "bow_ori": lambda impl: case_bow(impl, 31, 0.75, True),
    "bow_noori_tight": lambda impl: case_bow(impl, 32, 0.6, False),
    "triangulation": lambda impl: case_triangulation(impl, 41, False),
    "triangulation_ori": lambda impl: case_triangulation(impl, 42, True),
    "initialization_w100": lambda impl: case_initialization(impl, 51, 100, 0.9, True),
    "initialization_w30_noori": lambda impl: case_initialization(impl, 52, 30, 0.9, False),
}


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_compiled_reference(name):
    o = CASES[name](None)
    r = CASES[name]("reference")
    assert o[0] == r[0], f"{name}: match count {o[0]} (oracle) vs {r[0]} (reference)"
    assert o[0] > 20, f"{name}: the case does not exercise the method ({o[0]} matches)"
    for a, b in zip(o[1:], r[1:]):
        assert np.array_equal(a, b), name


@needs_ref
def test_descriptor_distance_equals_compiled_reference():
    rng = np.random.default_rng(5)
    a, b = rand_desc(rng, 300, live_bits=256), rand_desc(rng, 300, live_bits=256)
    for i in range(300):
        want = int(np.unpackbits(a[i] ^ b[i]).sum())
        assert oracle.ref_descriptor_distance(a[i], b[i]) == want == oracle.distance(a[i], b[i])


def test_oracle_matches_committed_reference_digests():
    """The digests were written from the COMPILED REFERENCE (python tests/test_search_reference.py); the oracle must
    reproduce them wherever the suite runs, including boxes without /root/reference."""
    gold = json.load(open(GOLDEN))
    for name in sorted(CASES):
        o = CASES[name](None)
        assert gold[name]["matches"] == int(o[0]), name
        assert gold[name]["sha256_24"] == digest(*o[1:]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_committed_reference_digests(name):
    """The CUDA search methods against the digests of the compiled reference, without the oracle in between."""
    gold = json.load(open(GOLDEN))
    g = CASES[name]("gpu")
    assert gold[name]["matches"] == int(g[0]), name
    assert gold[name]["sha256_24"] == digest(*g[1:]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_adapter_class_matches_committed_reference_digests(name):
    """The drop-in ORB_SLAM2::ORBmatcher (adapter/ORBmatcher.h + ORBmatcher_b200.cc: the reference's signatures, Frame& /
    KeyFrame* / MapPoint* arguments, gather and write-back around the orb_search_* calls) run through the bridge that
    runs the reference's own class, against the digests of the compiled reference."""
    if oracle.adapter_match_lib() is None:
        pytest.skip("oracle/_ref/libadapter_match.so not built (oracle/Makefile adaptermatch)")
    gold = json.load(open(GOLDEN))
    a = CASES[name]("adapter")
    assert gold[name]["matches"] == int(a[0]), name
    assert gold[name]["sha256_24"] == digest(*a[1:]), name


@pytest.mark.gpu
def test_adapter_descriptor_distance():
    if oracle.adapter_match_lib() is None:
        pytest.skip("oracle/_ref/libadapter_match.so not built")
    rng = np.random.default_rng(6)
    a, b = rand_desc(rng, 40, live_bits=256), rand_desc(rng, 40, live_bits=256)
    fn = oracle.adapter_match_lib().refm_descriptor_distance
    for i in range(40):
        assert fn(a[i].ctypes.data, b[i].ctypes.data) == int(np.unpackbits(a[i] ^ b[i]).sum())


if __name__ == "__main__":  # regenerate the golden digests from the compiled reference
    assert HAVE_REF
    out = {}
    for name in sorted(CASES):
        r = CASES[name]("reference")
        out[name] = {"matches": int(r[0]), "sha256_24": digest(*r[1:])}
    json.dump({"source": "reference src/ORBmatcher.cc compiled unmodified against oracle/mshim (oracle/Makefile refmatch)", **out},
              open(GOLDEN, "w"), indent=1)
    print("wrote", GOLDEN)
