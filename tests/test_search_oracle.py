"""CPU tests of the oracle's search-method restatements (reference src/ORBmatcher.cc, src/Frame.cc:307-372):
window generation against an independent numpy restatement, and method-level invariants."""
import numpy as np

import oracle
from search_cases import SCALE, bow_pair, make_frame, numpy_features_in_area, projected_queries


def test_features_in_area_matches_numpy_restatement():
    rng = np.random.default_rng(101)
    for clustered, mn in [(False, 0.0), (True, 0.0), (False, -13.5)]:
        F = make_frame(rng, 700, clustered=clustered, min_x=mn, min_y=mn / 2)
        _, _, u, v, level, _ = projected_queries(rng, F, 120, outside=0.1)
        r = rng.choice([2.5, 4.0, 10.0, 37.5, 100.0], 120).astype(np.float32)
        lo = np.where(rng.random(120) < 0.3, -1, level - 1).astype(np.int32)
        hi = np.where(rng.random(120) < 0.3, -1, level + 1).astype(np.int32)
        off, cand = oracle.features_in_area(F, u, v, r, lo, hi)
        for i in range(120):
            want = numpy_features_in_area(F, u[i], v[i], r[i], int(lo[i]), int(hi[i]))
            assert list(cand[off[i]:off[i + 1]]) == want, i
        assert off[-1] > 200


def test_search_by_projection_map_properties():
    rng = np.random.default_rng(7)
    F = make_frame(rng, 1500, stereo=True)
    src, q, u, v, level, _ = projected_queries(rng, F, 800)
    occ = (rng.random(F.N) < 0.2).astype(np.uint8)
    occ0 = occ.copy()
    vc = rng.choice([0.9, 0.9985], 800).astype(np.float32)
    xr = (u - 10).astype(np.float32)
    n, fq = oracle.search_by_projection_map(F, occ, q, u, v, xr, level, vc, 1.0, 0.8)
    assert n == (fq >= 0).sum() > 100
    got = fq[fq >= 0]
    assert len(set(got)) == len(got)          # greedy state: a feature is claimed once
    assert not occ0[got].any()                # never an occupied feature
    assert occ[got].all()


def test_search_for_initialization_is_one_to_one():
    rng = np.random.default_rng(9)
    F2 = make_frame(rng, 1800)
    src = rng.integers(0, F2.N, 1500)
    k1 = F2.keys_un[src].copy()
    k1["octave"][::3] = 0
    from search_cases import noisy_copy
    d1 = noisy_copy(rng, F2.desc[src], 30)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32).copy()
    n, m12 = oracle.search_for_initialization(k1, d1, F2, prev, 100, 0.9, True)
    got = m12[m12 >= 0]
    assert n == len(got) > 50 and len(set(got)) == len(got)
    assert (k1["octave"][m12 >= 0] == 0).all()


def test_bow_and_triangulation_shapes():
    rng = np.random.default_rng(3)
    F1, F2, fv1, fv2, has1, has2 = bow_pair(rng)
    kf1 = dict(desc=F1.desc, keys=F1.keys_un, has_mp=has1, featvec=fv1)
    kf2 = dict(desc=F2.desc, keys=F2.keys_un, has_mp=has2, featvec=fv2, sigma2=SCALE * SCALE)
    n, m = oracle.search_by_bow_kf(kf1, kf2, 0.75, True)
    assert n == (m >= 0).sum() > 20
    assert has1[m >= 0].all() and has2[m[m >= 0]].all()
    F12 = np.array([[0, -1e-4, 0.02], [1e-4, 0, -0.03], [-0.02, 0.03, 0]], np.float32)
    n2, m2 = oracle.search_for_triangulation(kf1, kf2, F12, False)
    assert n2 == (m2 >= 0).sum()
    assert not has1[m2 >= 0].any() and not has2[m2[m2 >= 0]].any()
