"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(orb_slam_system_b200) never does.

The oracle restates reference src/ORBextractor.cc, src/ORBmatcher.cc:896-908
and src/Frame.cc:446-529; see oracle/orb_oracle.h for the pinning story.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborb_oracle.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libref_orb.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def build(force=False):
    """Compile liborb_oracle.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("orb_oracle.cpp", "orb_oracle.h")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, _LIB])
    return _LIB


def build_ref():
    """Compile the reference's own ORBextractor.cc against oracle/cvshim (needs /root/reference)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return _REF_LIB if os.path.exists(_REF_LIB) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_extract_many.restype = C.c_double
        _lib.orc_extract_many.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_longlong)]
        _lib.orc_synth_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                         C.c_uint64, C.c_int, C.c_int]
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


DEFAULT = dict(nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)


def synth_frame(rows, cols, seed=7, frame=0, variant=0, right=0):
    out = np.empty((rows, cols), np.uint8)
    lib().orc_synth_frame(_p(out), rows, cols, cols, seed, frame, variant, right)
    return out


def tables(nfeatures=1000, scaleFactor=1.2, nlevels=8):
    sc = np.zeros(nlevels, np.float32)
    inv = np.zeros(nlevels, np.float32)
    s2 = np.zeros(nlevels, np.float32)
    is2 = np.zeros(nlevels, np.float32)
    nf = np.zeros(nlevels, np.int32)
    um = np.zeros(16, np.int32)
    lib().orc_tables(nfeatures, C.c_float(scaleFactor), nlevels, _p(sc), _p(inv), _p(s2), _p(is2),
                     _p(nf), _p(um))
    return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, features_per_level=nf, umax=um)


def extract(img, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, cap=None,
            info=None):
    """Oracle ORBextractor::operator(). Returns (keypoints[KP_DTYPE], descriptors[K,32])."""
    img = _img(img)
    cap = cap or 8 * max(nfeatures, 64)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    ncand = np.zeros(nlevels, np.int32)
    nkept = np.zeros(nlevels, np.int32)
    nretry = C.c_int(0)
    rc = lib().orc_extract(nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST, minThFAST,
                           _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps), _p(desc),
                           cap, C.byref(cnt), _p(ncand), _p(nkept), C.byref(nretry))
    if rc != 0:
        raise RuntimeError(f"oracle extract failed rc={rc}")
    if info is not None:
        info.update(candidates=ncand, kept=nkept, retry_cells=nretry.value)
    n = cnt.value
    return kps[:n].copy(), desc[:n].copy()


def pyramid_level(img, level, scaleFactor=1.2, nlevels=8):
    img = _img(img)
    dst = np.zeros(img.size, np.uint8)
    r = C.c_int(0)
    c = C.c_int(0)
    rc = lib().orc_pyramid(C.c_float(scaleFactor), nlevels, _p(img), img.shape[0], img.shape[1],
                           img.strides[0], level, _p(dst), dst.size, C.byref(r), C.byref(c))
    assert rc == 0
    return dst[: r.value * c.value].reshape(r.value, c.value).copy()


def resize(img, drows, dcols):
    img = _img(img)
    dst = np.zeros((drows, dcols), np.uint8)
    lib().orc_resize(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(dst), drows, dcols, dcols)
    return dst


def blur7(img):
    img = _img(img)
    dst = np.zeros_like(img)
    lib().orc_blur7(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(dst), dst.strides[0])
    return dst


def fast(img, threshold):
    """cv::FAST(img, threshold, nms=True) model: int32 [n,3] rows (x, y, score), row-major order."""
    img = _img(img)
    cap = max(16, img.size // 4 + 16)
    out = np.zeros((cap, 3), np.int32)
    n = lib().orc_fast(_p(img), img.shape[0], img.shape[1], img.strides[0], threshold, _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def octree(xyr, minX, maxX, minY, maxY, N):
    """DistributeOctTree: xyr float32 [n,3] (x, y, response) -> kept indices in reference order."""
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    out = np.zeros(max(1, len(xyr)), np.int32)
    k = lib().orc_octree(_p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(out), len(out))
    if k < 0:
        raise RuntimeError("octree does not terminate in the reference")
    return out[:k].copy()


def ic_angle(img, x, y):
    img = _img(img)
    return float(lib().orc_ic_angle(_p(img), img.shape[0], img.shape[1], img.strides[0], x, y))


def describe(blurred, x, y, angle):
    blurred = _img(blurred)
    d = np.zeros(32, np.uint8)
    lib().orc_describe(_p(blurred), blurred.shape[0], blurred.shape[1], blurred.strides[0], x, y,
                       C.c_float(angle), _p(d))
    return d


def distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_distance(_p(a), _p(b)))


def match_all(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    bi = np.zeros(len(q), np.int32)
    bd = np.zeros(len(q), np.int32)
    sd = np.zeros(len(q), np.int32)
    lib().orc_match_all(_p(q), len(q), _p(t), len(t), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def match_csr(q, t, offsets, cand, tie_last=False, max_dist=50):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    cand = np.ascontiguousarray(cand, np.int32)
    assert len(offsets) == len(q) + 1
    bi = np.zeros(len(q), np.int32)
    bd = np.zeros(len(q), np.int32)
    sd = np.zeros(len(q), np.int32)
    lib().orc_match_csr(_p(q), len(q), _p(t), len(t), _p(offsets), _p(cand), int(tie_last),
                        int(max_dist), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def stereo_match(kl, dl, kr, dr, scale, rows, bf, fx):
    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8)
    dr = np.ascontiguousarray(dr, np.uint8)
    scale = np.ascontiguousarray(scale, np.float32)
    br = np.zeros(len(kl), np.int32)
    bd = np.zeros(len(kl), np.int32)
    rc = lib().orc_stereo_match(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(scale),
                                len(scale), rows, C.c_float(bf), C.c_float(fx), _p(br), _p(bd))
    if rc != 0:
        raise RuntimeError("stereo row band out of range (reference UB)")
    return br, bd


def extract_many(rows, cols, nframes, nthreads, first_frame=0, seed=7, **params):
    p = dict(DEFAULT)
    p.update(params)
    tot = C.c_longlong(0)
    secs = lib().orc_extract_many(p["nfeatures"], C.c_float(p["scaleFactor"]), p["nlevels"],
                                  p["iniThFAST"], p["minThFAST"], rows, cols, seed, first_frame,
                                  nframes, nthreads, C.byref(tot))
    return secs, tot.value


# ---- the reference's own extractor (oracle/_ref/libref_orb.so), when built --------------------
_ref = None


def ref_lib():
    """CDLL of the reference ORBextractor.cc compiled against oracle/cvshim, or None."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_LIB) and os.path.exists("/root/reference/src/ORBextractor.cc"):
            build_ref()
        if not os.path.exists(_REF_LIB):
            return None
        _ref = C.CDLL(_REF_LIB)
        _ref.ref_extract_many.restype = C.c_double
        _ref.ref_extract_many.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_longlong)]
    return _ref


def ref_extract(img, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, cap=None):
    img = _img(img)
    cap = cap or 8 * max(nfeatures, 64)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    rc = ref_lib().ref_extract(nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST, minThFAST,
                               _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps),
                               _p(desc), cap, C.byref(cnt))
    if rc != 0:
        raise RuntimeError(f"reference extract failed rc={rc}")
    return kps[: cnt.value].copy(), desc[: cnt.value].copy()


def ref_extract_many(rows, cols, nframes, nthreads, first_frame=0, seed=7, **params):
    p = dict(DEFAULT)
    p.update(params)
    tot = C.c_longlong(0)
    secs = ref_lib().ref_extract_many(p["nfeatures"], C.c_float(p["scaleFactor"]), p["nlevels"],
                                      p["iniThFAST"], p["minThFAST"], rows, cols, seed, first_frame,
                                      nframes, nthreads, C.byref(tot))
    return secs, tot.value
