"""The C++ drop-in adapter (orb_slam_system_b200/adapter: ORB_SLAM2::ORBextractor with the reference's
signatures, over the C ABI) against the reference's own class compiled from its sources -- both driven
through the same C bridge (oracle/cvshim/ref_bridge.cpp), as a maintainer's test of the swap would."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from orb_slam_system_b200.synth import synth_frame

pytestmark = pytest.mark.gpu

ADAPTER = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libadapter_orb.so")


def _load():
    if not os.path.exists(ADAPTER):
        pytest.skip("oracle/_ref/libadapter_orb.so not built")
    return C.CDLL(ADAPTER)


def _extract(lib, img, nf):
    cap = 16 * nf
    kps = np.zeros(cap, oracle.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    rc = lib.ref_extract(nf, C.c_float(1.2), 8, 20, 7, img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], img.strides[0],
                         kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, C.byref(cnt))
    assert rc == 0
    return kps[:cnt.value], desc[:cnt.value]


@pytest.mark.parametrize("rows,cols,nf", [(480, 640, 1000), (376, 1241, 2000)])
def test_adapter_equals_reference_class(rows, cols, nf):
    lib = _load()
    img = synth_frame(rows, cols, frame=3)
    ka, da = _extract(lib, img, nf)
    if oracle.ref_lib() is not None:
        kr, dr = oracle.ref_extract(img, nfeatures=nf, cap=16 * nf)  # the reference's own ORBextractor.cc
    else:
        kr, dr = oracle.extract(img, nfeatures=nf, cap=16 * nf)
    assert len(ka) == len(kr)
    for fld in ("x", "y", "size", "response", "octave", "class_id"):
        assert (ka[fld] == kr[fld]).all(), fld
    assert np.abs(ka["angle"] - kr["angle"]).max() <= 1e-3
    assert (da == dr).all()


def test_adapter_pyramid_has_reference_border():
    lib = _load()
    img = synth_frame(376, 1241, frame=4)
    for level in (0, 1, 3, 7):
        r, c = C.c_int(0), C.c_int(0)
        buf = np.zeros(img.size, np.uint8)
        rc = lib.ref_pyramid_level(C.c_float(1.2), 8, img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], img.strides[0], level,
                                   buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(r), C.byref(c))
        assert rc == 0
        got = buf[: r.value * c.value].reshape(r.value, c.value)
        assert (got == oracle.pyramid_level(img, level)).all()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_adapter_matcher_search_methods():
    """orb_b200::Matcher (adapter/orb_match_b200.h) -- the C++ a maintainer pastes into ORBmatcher.cc -- driven through its
    C bridge: SearchByBoW with DBoW2-shaped maps and SearchForInitialization with cv::KeyPoint / cv::Point2f containers."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from search_cases import bow_pair, make_frame, noisy_copy
    lib = _load()
    if not hasattr(lib, "adapter_search_by_bow"):
        pytest.skip("adapter built without the matcher bridge")
    rng = np.random.default_rng(8)
    F1, F2, fv1, fv2, has1, has2 = bow_pair(rng)
    n1, o1, i1 = oracle._featvec_csr(fv1)
    n2, o2, i2 = oracle._featvec_csr(fv2)
    a1 = np.ascontiguousarray(F1.keys_un["angle"], np.float32)
    a2 = np.ascontiguousarray(F2.keys_un["angle"], np.float32)
    out = np.zeros(F1.N, np.int32)
    lib.adapter_search_by_bow.restype = C.c_int
    n = lib.adapter_search_by_bow(_p(F1.desc), _p(a1), _p(has1), F1.N, _p(F2.desc), _p(a2), _p(has2), F2.N, _p(n1), _p(o1), _p(i1), len(n1),
                                  _p(n2), _p(o2), _p(i2), len(n2), C.c_float(0.75), 1, _p(out))
    kf1 = dict(desc=F1.desc, keys=F1.keys_un, has_mp=has1, featvec=fv1)
    kf2 = dict(desc=F2.desc, keys=F2.keys_un, has_mp=has2, featvec=fv2)
    no, oo = oracle.search_by_bow_kf(kf1, kf2, 0.75, True)
    assert n == no and (out == oo).all() and n > 20

    G = make_frame(rng, 1500)
    src = rng.integers(0, G.N, 1200)
    k1 = G.keys_un[src].copy()
    k1["octave"][::2] = 0
    d1 = noisy_copy(rng, G.desc[src], 40)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32).copy()
    pa, po = prev.copy(), prev.copy()
    m12 = np.zeros(len(k1), np.int32)
    lib.adapter_search_for_initialization.restype = C.c_int
    n = lib.adapter_search_for_initialization(_p(k1), _p(d1), len(k1), _p(G.keys_un), _p(G.desc), G.N, C.c_float(G.mnMinX),
                                              C.c_float(G.mnMinY), C.c_float(G.mfGridElementWidthInv), C.c_float(G.mfGridElementHeightInv),
                                              _p(pa), 100, C.c_float(0.9), 1, _p(m12))
    no, oo = oracle.search_for_initialization(k1, d1, G, po, 100, 0.9, True)
    assert n == no and (m12 == oo).all() and (pa == po).all() and n > 50
