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
